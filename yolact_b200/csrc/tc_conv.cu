// Implicit-GEMM convolution on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces cuDNN nn.Conv2d + BatchNorm2d + ReLU (+ residual add) of the reference's backbone /
// FPN / protonet / prediction head (backbone.py:37-57,126-139; yolact.py:311-361,133-212;
// utils/functions.py:163-213) and the SGEMM of DCNv2 (dcn_v2_cuda.cu:149-163).
//
// GEMM view (no im2col buffer anywhere):
//   D[M = output pixels of one spatial tile, N = Cout tile] = sum over taps (r,s) and 64-channel
//   chunks of  A_tap[M, 64] * W_tap[N, 64]^T,   fp16 operands, fp32 accumulation in TMEM.
// A operand: the activation tensor is NHWC fp16; a 4-D TMA tensor map {C, W, H, B} loads the box
//   {64 channels, tw, th, 1} whose origin is shifted by the tap offset.  Out-of-bounds rows /
//   columns (the conv's zero padding) are zero-filled by TMA itself.  tw*th <= 128 rows land in
//   shared memory as 128-byte rows with the 128B swizzle == the canonical K-major SWIZZLE_128B
//   UMMA layout, so the MMA consumes them with no repacking.
//   Stride-2 convs use one tensor map per input phase (py,px) (a strided *view* of the same
//   buffer: element strides doubled), and the tap table selects (phase map, dx, dy).
// B operand: weights packed [tap][Cout][Cin] fp16, 3-D map, box {64, BN, 1}.
// Pipeline: warp 0 = TMA producer, warp 1 = TMEM alloc + single-thread tcgen05.mma issue,
//   warps 2..5 = epilogue (tcgen05.ld -> +bias -> +residual -> activation -> 16-byte stores).
//   `stages`-deep mbarrier ring (full/empty), tcgen05.commit releases shared memory slots.
// CTA pairs (PAIR = true): a cluster of two CTAs on one TPC computes two adjacent M tiles of the same N tile
//   with ONE tcgen05.mma.cta_group::2 (M = 256).  Each CTA loads its own A tile and only HALF of the weight
//   tile, so the bytes a CTA pulls from L2 per k-block drop from 16 KB + BN*128 to 16 KB + BN*64 -- the conv
//   stack at batch 8 is L2->SM bandwidth bound, not tensor bound.  The even CTA issues
//   the MMAs; both CTAs' TMA loads signal ITS full barrier; tcgen05.commit multicasts to both CTAs' barriers.
// Chain kernel (tc_chain_kernel, further down): a run of consecutive layers -- the whole ResNet trunk after the
//   max-pool -- in ONE persistent cooperative launch; the layers' parameter blocks live in global memory, the work
//   units of all layers are dealt round-robin to the CTAs, dependencies are tracked per M tile with counters.
#include <atomic>
#include <vector>
#include "tc_common.cuh"

namespace yb {

using namespace tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;              // fp16 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_THREADS = 192;         // 2 + 4*H warps: producer, MMA issuer, H epilogue groups (H = 1 here)
constexpr int MAX_TAPS = 9;
constexpr int MAX_STAGES = 8;

struct alignas(64) TcParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmY;      // output tile store  (epi_tma)
  CUtensorMap tmR;      // residual tile load (epi_tma && residual)
  CUtensorMap tmY32;    // output tile store in 32-channel boxes, 64-byte swizzle (chain kernel: 8 KB staging tiles)
  int epi_tma;          // 1: fp16 NHWC output goes smem -> TMA store, residual comes in by TMA
  int out_off;          // byte offset of the 2 x 16 KB epilogue staging tiles inside dynamic smem
  int res_off;          // byte offset of the residual staging buffers inside dynamic smem
  int m_tiles, n_tiles; // tile = m_tile * n_tiles + n_tile
  int acc_stages;       // TMEM accumulator buffers (2 when a CTA processes several tiles)
  int nseg;             // > 0: output channels are split over several fp32 tensors (fused prediction head)
  int seg_begin[3], seg_end[3], seg_ps[3], seg_act[3];
  long long seg_bs[3];
  float* seg_y[3];
  int tmem_cols;        // power of two >= acc_stages * BN
  int pdl;              // launched with programmatic stream serialization
  int pair;             // 1: CTA pairs (cluster of 2, tcgen05 cta_group::2, UMMA M = 256)
  int nb;               // batch extent of the tile grid (1 when flattened): tiles with b >= nb are padding
  int ntaps, kchunks, stages;
  int tap_map[MAX_TAPS], tap_dx[MAX_TAPS], tap_dy[MAX_TAPS];
  int tw, th, tiles_x, tiles_y;
  int Ho, Wo, Cout;
  int a_box_bytes;      // tw*th*128
  uint32_t idesc;
  // epilogue
  void* y;
  const float* bias;
  const __half* residual;
  long long y_batch_stride;
  long long res_batch_stride;
  int y_pix_stride;
  int y_f32;
  int act;
  int vec_ok;
  int res_after_act;
  // split precision (SPLIT kernels): activations are [hi(C) | lo(C)] fp16 pairs per pixel, weights [.. hi(Cin) | lo(Cin)]
  // scaled by a power of two; the accumulator is multiplied by out_scale (its exact inverse) before the bias
  int split;
  int cin;              // logical input channels: the lo plane starts at channel `cin` of the A / B tensor maps
  float out_scale;
  // stream-K (sk != 0): the units' k-block iterations are dealt out in equal contiguous shares, one share per CTA /
  // cluster; a share that ends inside a unit leaves an fp32 partial tile in sk_ws (one slot of 128 x BN floats per
  // CTA) and raises sk_flags[2 * slot + epilogue group]; the CTA holding the unit's FIRST k-blocks adds the partials
  // and runs the epilogue
  int sk;
  float* sk_ws;
  int* sk_flags;
  // split plans with two epilogue groups on a flattened (1x1, stride 1) layout read the residual straight from global
  // memory: its staging tiles (2 x 32 KB) would leave a single pipeline stage beside the two output staging buffers
  int res_direct;
  int bn;               // N tile (the kernel template's BN; the chain kernel reads it per layer)
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// Persistent: CTA c processes tiles c, c + gridDim.x, ...  (tile = m_tile * n_tiles + n_tile).
// Three concurrent pipelines: smem ring (TMA producer <-> MMA), TMEM accumulator ring of
// `acc_stages` buffers (MMA <-> epilogue), and the epilogue's own double-buffered staging tiles,
// so the loads of tile i+1, the MMAs of tile i and the stores of tile i-1 overlap.
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// activation with a compile-time selector (the generic apply_act() is a runtime switch: far too
// expensive inside the 64-element epilogue loops)
template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_LEAKY) return v > 0.f ? v : 0.1f * v;
  if (ACT == ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ __half2 pack_sat(float a, float b) {
  // saturate to the fp16 range instead of producing inf
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  return __floats2half2_rn(a, b);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One 64-channel chunk of one accumulator row: (* out_scale) +bias (smem, broadcast), +residual (swizzled smem
// tile), activation, fp16 pack, swizzled 16-byte stores into the output staging tile.  SPLIT: the residual is
// hi + lo from two tiles and the result is written as hi / lo = rn(v - hi) into two staging tiles.
// res_g (res_direct plans): this row's 64 residual channels in GLOBAL memory (hi plane; the lo plane res_lo_off halfs
// further) instead of a staged tile.
template <int ACT, bool RES_AFTER, bool SPLIT>
__device__ __forceinline__ void epi_chunk(const uint32_t* r0, const uint32_t* r1, const float* sbias,
                                          const uint8_t* res_tile, uint8_t* out_tile, uint32_t row, uint32_t sw,
                                          float out_scale, const __half* res_g = nullptr, int res_lo_off = 0) {
#pragma unroll
  for (int j8 = 0; j8 < 8; ++j8) {
    const float4 b0 = *reinterpret_cast<const float4*>(sbias + j8 * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(sbias + j8 * 8 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j8 * 8 + j;
      const float a = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]);
      v[j] = SPLIT ? __fmaf_rn(a, out_scale, bb[j]) : a + bb[j];
      if (RES_AFTER) v[j] = act_t<ACT>(v[j]);
    }
    const uint32_t off = row * 128u + (((uint32_t)j8 ^ sw) << 4);
    if (res_tile || res_g) {
      const uint4 raw = res_tile ? *reinterpret_cast<const uint4*>(res_tile + off)
                                 : __ldcg(reinterpret_cast<const uint4*>(res_g + j8 * 8));   // L2: inside a chain other SMs wrote it during this launch
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        v[2 * j] += f.x;
        v[2 * j + 1] += f.y;
      }
      if (SPLIT) {
        const uint4 rawl = res_tile ? *reinterpret_cast<const uint4*>(res_tile + A_STAGE_BYTES + off)
                                    : __ldcg(reinterpret_cast<const uint4*>(res_g + res_lo_off + j8 * 8));
        const __half2* l2 = reinterpret_cast<const __half2*>(&rawl);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = lo2_to_f32(l2[j]);   // lo plane: 2^11-scaled
          v[2 * j] += f.x;
          v[2 * j + 1] += f.y;
        }
      }
    }
    if (!RES_AFTER) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_t<ACT>(v[j]);
    }
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = pack_sat(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(out_tile + off) = o;
    if (SPLIT) {
      uint4 ol;
      __half2* l2 = reinterpret_cast<__half2*>(&ol);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 hf = __half22float2(o2[j]);
        // hi saturates at +-65504 (pack_sat); the residual is taken against the clamped value, so a saturated
        // element gets lo = 0 and hi + lo stays finite
        const float c0 = fminf(fmaxf(v[2 * j], -65504.f), 65504.f), c1 = fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f);
        l2[j] = lo2_from_f32(c0 - hf.x, c1 - hf.y);   // lo plane: residual * 2^11 (see common.cuh)
      }
      *reinterpret_cast<uint4*>(out_tile + A_STAGE_BYTES + off) = ol;
    }
  }
}

// Chain-kernel variant: a 32-channel chunk of one accumulator row (r: the combined fp32 accumulator) into an 8 KB staging
// tile of 128 rows x 64 bytes with the 64-byte TMA swizzle (16-byte unit index ^= (row >> 1) & 3); the lo plane's tile
// follows the hi plane's.  The residual always comes from global memory (res_g: this row's 32 channels, hi plane).
// (Requesting the residual of all of a group's chunks up front, into registers, was measured SLOWER -- 4.01 vs 3.69 ms
// for the 105-layer trunk chain, profiles/r2_call17_summary.txt: 168 registers, and the burst of loads delays the TMA.)
constexpr int CHUNK32_BYTES = BLOCK_M * 32 * 2;   // 8 KB
template <int ACT, bool RES_AFTER, bool SPLIT>
__device__ __forceinline__ void epi_chunk32(const uint32_t* r, const float* sbias, uint8_t* out_tile, uint32_t row,
                                            float out_scale, const __half* res_g, int res_lo_off) {
  const uint32_t sw = (row >> 1) & 3u;
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    const float4 b0 = *reinterpret_cast<const float4*>(sbias + j8 * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(sbias + j8 * 8 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = __uint_as_float(r[j8 * 8 + j]);
      v[j] = SPLIT ? __fmaf_rn(a, out_scale, bb[j]) : a + bb[j];
      if (RES_AFTER) v[j] = act_t<ACT>(v[j]);
    }
    if (res_g) {
      const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(res_g + j8 * 8));   // L2: other SMs wrote it during this launch
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        v[2 * j] += f.x;
        v[2 * j + 1] += f.y;
      }
      if (SPLIT) {
        const uint4 rawl = __ldcg(reinterpret_cast<const uint4*>(res_g + res_lo_off + j8 * 8));
        const __half2* l2 = reinterpret_cast<const __half2*>(&rawl);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = lo2_to_f32(l2[j]);
          v[2 * j] += f.x;
          v[2 * j + 1] += f.y;
        }
      }
    }
    if (!RES_AFTER) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_t<ACT>(v[j]);
    }
    const uint32_t off = row * 64u + (((uint32_t)j8 ^ sw) << 4);
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = pack_sat(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(out_tile + off) = o;
    if (SPLIT) {
      uint4 ol;
      __half2* l2 = reinterpret_cast<__half2*>(&ol);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 hf = __half22float2(o2[j]);
        const float c0 = fminf(fmaxf(v[2 * j], -65504.f), 65504.f), c1 = fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f);
        l2[j] = lo2_from_f32(c0 - hf.x, c1 - hf.y);
      }
      *reinterpret_cast<uint4*>(out_tile + CHUNK32_BYTES + off) = ol;
    }
  }
}

struct TileCoord {
  int b, x0, y0, n0;
};
__device__ __forceinline__ TileCoord decode_mn(const TcParams& p, int m, int nt, int BN) {
  TileCoord t;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  t.b = m / p.tiles_y;
  t.x0 = tx * p.tw;
  t.y0 = ty * p.th;
  t.n0 = nt * BN;
  return t;
}
// Work unit u of a CTA: one (M tile, N tile).  Single CTAs walk tiles; a CTA pair walks (M-tile pair, N tile)
// units, CTA `rank` taking M tile 2*group + rank (a padding tile when the M-tile count is odd: its loads are
// zero-filled and its stores clipped by the tensor maps, b >= nb marks it for the direct epilogue).
template <bool PAIR>
__device__ __forceinline__ TileCoord decode_unit(const TcParams& p, int u, int rank, int BN) {
  const int nt = u % p.n_tiles;
  const int mg = u / p.n_tiles;
  return decode_mn(p, PAIR ? 2 * mg + rank : mg, nt, BN);
}

// Work of one CTA (cluster): a list of segments (unit, [k0, k1)).  Without stream-K: whole units unit0, unit0 + step, ...
// With stream-K: the contiguous iteration range [cid * I / G, (cid + 1) * I / G) of I = units * k-blocks, cut at unit
// boundaries -- so a CTA's first segment may be the TAIL of a unit (k0 > 0: dumped as a partial) and its last one the
// HEAD of a unit (k0 == 0, k1 < num_kb: this CTA collects the partials and finishes the unit).
struct WorkIter {
  int num_kb, num_units, ustep, sk, u;
  long long it, it_hi;
  __device__ __forceinline__ static long long share_lo(long long total, int c, int G) { return total * c / G; }
  __device__ __forceinline__ void init(int sk_, int num_kb_, int num_units_, int unit0, int ustep_) {
    sk = sk_;
    num_kb = num_kb_;
    num_units = num_units_;
    ustep = ustep_;
    u = unit0;
    const long long total = (long long)num_units_ * num_kb_;
    it = share_lo(total, unit0, ustep_);
    it_hi = share_lo(total, unit0 + 1, ustep_);
  }
  __device__ __forceinline__ bool next(int& uu, int& k0, int& k1) {
    if (!sk) {
      if (u >= num_units) return false;
      uu = u;
      k0 = 0;
      k1 = num_kb;
      u += ustep;
      return true;
    }
    if (it >= it_hi) return false;
    uu = (int)(it / num_kb);
    k0 = (int)(it - (long long)uu * num_kb);
    const long long rem = it_hi - it;
    k1 = (rem < (long long)(num_kb - k0)) ? (int)(k0 + rem) : num_kb;
    it += k1 - k0;
    return true;
  }
};

template <int BN, bool PAIR, int H, bool SPLIT>
__global__ void __launch_bounds__(64 + 128 * H)
tc_conv_kernel(const __grid_constant__ TcParams p) {
  // SPLIT (YB_PREC_F16X3): every operand has a hi and a (2^11-scaled) lo fp16 plane; a k-block stages A_hi, A_lo, W_hi,
  // W_lo and issues A_hi*W_hi into the tile's first accumulator and A_lo*W_hi + A_hi*W_lo into its second one (BN TMEM
  // columns further); the epilogue combines acc_hi + 2^-11 * acc_lo (the lo*lo term is below fp32 resolution).
  constexpr int NPL = SPLIT ? 2 : 1;
  constexpr int B_PLANE_BYTES = (PAIR ? BN / 2 : BN) * BLOCK_K * 2;   // a pair CTA stages half of the weight tile
  constexpr int A_BYTES = NPL * A_STAGE_BYTES;
  constexpr int B_STAGE_BYTES = NPL * B_PLANE_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;

  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint64_t res_full_bar[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float sbias[H * BN];   // one copy per epilogue group

  // 1024-byte alignment required by SWIZZLE_128B (host adds 1024 bytes of slack)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_base = smem + p.out_off;   // epilogue staging tiles: [buffer][plane] x 16 KB
  uint8_t* res_base = smem + p.res_off;   // residual tiles, same layout (only when residual && epi_tma)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int num_kb = p.ntaps * p.kchunks;
  // work units of this CTA: unit0, unit0 + ustep, ... < num_units
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int unit0 = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
  const int ustep = PAIR ? (int)cluster_nclusters_x() : (int)gridDim.x;
  const int num_units = (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  const int acc_stages = p.acc_stages;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], (PAIR ? 2 : 1) * H);   // every epilogue group (of both CTAs of a pair) drains it
      mbar_init(&res_full_bar[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
    if (p.epi_tma) tma_prefetch_desc(&p.tmY);
  }
  if (warp == 1) tmem_alloc_dyn(&s_tmem_base, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // Programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream, and so do
  // the WEIGHT tiles of this CTA's first k-blocks (constants: they do not depend on the previous layer); the
  // activations are only touched after griddepcontrol.wait.  Dependents are released at once: they can become resident
  // (and do the same) as soon as an SM has room -- which needs a plan that leaves room (plan->pdl_friendly).
  uint32_t pre = 0;   // producer thread only: k-blocks whose weight tile is already in flight
  if (p.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!PAIR && !SPLIT && !p.sk && warp == 0 && lane == 0 && unit0 < num_units) {   // (a stream-K walk starts mid-unit)
      const TileCoord t0 = decode_unit<PAIR>(p, unit0, rank, BN);
      const int npre = min(stages, num_kb);
      for (int kb = 0; kb < npre; ++kb) {
        const int tap = kb / p.kchunks;
        const int kc = kb - tap * p.kchunks;
        mbar_expect_tx(&full_bar[kb], (uint32_t)p.a_box_bytes + (uint32_t)B_STAGE_BYTES);
        tma_load_3d(smem + (size_t)kb * STAGE_BYTES + A_BYTES, &p.tmB, &full_bar[kb], kc * BLOCK_K, t0.n0, tap);
      }
      pre = (uint32_t)npre;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // a pair's leader expects the bytes of BOTH CTAs (each: its A tile + its half of the weight tile, per plane)
      const uint32_t tx_bytes = ((uint32_t)p.a_box_bytes + (uint32_t)B_PLANE_BYTES) * (uint32_t)NPL * (PAIR ? 2u : 1u);
      const uint32_t leader_full = PAIR ? mapa_u32(smem_u32(&full_bar[0]), 0) : 0u;
      uint32_t kbg = 0;
      WorkIter wi;
      wi.init(p.sk, num_kb, num_units, unit0, ustep);
      int u, k0, k1;
      while (wi.next(u, k0, k1)) {
        const TileCoord tc_ = decode_unit<PAIR>(p, u, rank, BN);
        for (int kb = k0; kb < k1; ++kb, ++kbg) {
          const uint32_t s = kbg % (uint32_t)stages;
          const uint32_t it = kbg / (uint32_t)stages;
          mbar_wait(&empty_bar[s], (it & 1u) ^ 1u);
          const int tap = kb / p.kchunks;
          const int kc = kb - tap * p.kchunks;
          uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const CUtensorMap* ma = &p.tmA[p.tap_map[tap]];
          const int ax = tc_.x0 + p.tap_dx[tap], ay = tc_.y0 + p.tap_dy[tap];
          if (PAIR) {
            if (rank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);
            const uint32_t fb = leader_full + s * (uint32_t)sizeof(uint64_t);
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
              tma_load_4d_pair(sa + pl * A_STAGE_BYTES, ma, fb, kc * BLOCK_K + pl * p.cin, ax, ay, tc_.b);
              tma_load_3d_pair(sb + pl * B_PLANE_BYTES, &p.tmB, fb, kc * BLOCK_K + pl * p.cin, tc_.n0 + rank * (BN / 2), tap);
            }
          } else if (!SPLIT && kbg < pre) {
            // PDL: this stage was armed and its weight tile requested before the dependency wait
            tma_load_4d(sa, ma, &full_bar[s], kc * BLOCK_K, ax, ay, tc_.b);
          } else {
            mbar_expect_tx(&full_bar[s], tx_bytes);
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
              tma_load_4d(sa + pl * A_STAGE_BYTES, ma, &full_bar[s], kc * BLOCK_K + pl * p.cin, ax, ay, tc_.b);
              tma_load_3d(sb + pl * B_PLANE_BYTES, &p.tmB, &full_bar[s], kc * BLOCK_K + pl * p.cin, tc_.n0, tap);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && rank == 0) {   // in a pair only the even CTA issues (for both)
      uint32_t kbg = 0, t = 0;
      WorkIter wi;
      wi.init(p.sk, num_kb, num_units, unit0, ustep);
      int u, k0, k1;
      for (; wi.next(u, k0, k1); ++t) {
        const uint32_t acc = t % (uint32_t)acc_stages;
        const uint32_t use = t / (uint32_t)acc_stages;
        mbar_wait(&tmem_empty_bar[acc], (use & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)(NPL * BN);
        for (int kb = k0; kb < k1; ++kb, ++kbg) {
          const uint32_t s = kbg % (uint32_t)stages;
          const uint32_t it = kbg / (uint32_t)stages;
          mbar_wait(&full_bar[s], it & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = make_sw128_desc(sa);
          const uint64_t db = make_sw128_desc(sb);
          auto mma = [&](uint64_t a, uint64_t b, uint32_t accum) {
            if (PAIR) umma_f16_pair(tmem_d, a, b, p.idesc, accum); else umma_f16(tmem_d, a, b, p.idesc, accum);
          };
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the >>4 field
            mma(da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (kb > k0 || k > 0) ? 1u : 0u);
          }
          if (SPLIT) {
            const uint64_t dal = make_sw128_desc(sa + A_STAGE_BYTES);
            const uint64_t dbl = make_sw128_desc(sb + B_PLANE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)   // A_lo * W_hi -> second accumulator
              if (PAIR) umma_f16_pair(tmem_d + BN, dal + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb > k0 || k > 0) ? 1u : 0u);
              else umma_f16(tmem_d + BN, dal + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb > k0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)   // A_hi * W_lo -> second accumulator
              if (PAIR) umma_f16_pair(tmem_d + BN, da + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), p.idesc, 1u);
              else umma_f16(tmem_d + BN, da + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), p.idesc, 1u);
          }
          // frees this smem slot (in both CTAs of a pair) once the MMAs above have read it
          if (PAIR) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
        }
        // accumulator complete (both halves)
        if (PAIR) umma_commit_pair(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue: H groups of 4 warps (TMEM lane quadrant = warp % 4) =====================
    // H == 1: warps 2..5 process every 64-channel chunk of the tile, double-buffered staging tiles.
    // H == 2: warps 2..5 take the even chunks and warps 6..9 the odd ones, one staging tile (and one
    //         residual tile) per group: two chunks are in flight at once, which halves the epilogue a CTA
    //         with a single tile cannot hide behind its main loop.
    const int quad = warp & 3;
    const int hgrp = (H == 2) ? ((warp - 2) >> 2) : 0;       // epilogue group of this warp
    const int row = quad * 32 + lane;  // accumulator row == tile-local output pixel
    const bool issuer = (warp == 2 + 4 * hgrp && lane == 0);
    const bool has_res = (p.residual != nullptr) && !p.res_direct;   // residual staged by TMA (res_direct: read from global)
    const uint32_t sw = (uint32_t)(row & 7);
    float* my_bias = sbias + hgrp * BN;
    constexpr int NBUF = (H == 2 || SPLIT) ? 1 : 2;           // staging / residual buffers per group (a buffer = NPL tiles)
    constexpr int BUF_BYTES = NPL * A_STAGE_BYTES;
    auto group_sync = [&]() {                                 // the 128 threads of this group
      if (hgrp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };

    const uint32_t leader_tmem_empty = PAIR ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    auto release_acc = [&](uint32_t acc) {   // this group's 128 threads have read the accumulator
      if (PAIR && rank != 0)
        mbar_arrive_cluster(leader_tmem_empty + acc * (uint32_t)sizeof(uint64_t));
      else
        mbar_arrive(&tmem_empty_bar[acc]);
    };
    // residual chunk stream (epi_tma only): this group's chunks, in order and across tiles; chunk number gq of
    // the group lives in residual tile (H == 2 ? hgrp : gq & 1); the issuer keeps NBUF chunks in flight.
    // (pf_tile, pf_c) = next chunk to fetch.
    WorkIter pfi;                    // walks ahead over the segments that run the epilogue proper (k0 == 0)
    pfi.init(p.sk, num_kb, num_units, unit0, ustep);
    int pf_tile = -1, pf_c = hgrp;
    auto pf_advance = [&]() {
      int uu, a, b;
      pf_tile = -1;
      while (pfi.next(uu, a, b))
        if (a == 0) {
          pf_tile = uu;
          break;
        }
    };
    pf_advance();
    uint32_t pf_g = 0;
    auto prefetch_res = [&]() {
      while (pf_tile >= 0) {
        const TileCoord tcp = decode_unit<PAIR>(p, pf_tile, rank, BN);
        const int nch = (min(BN, p.Cout - tcp.n0) + 63) >> 6;
        if (pf_c >= nch) {            // this group has no (more) chunks in that tile
          pf_c = hgrp;
          pf_advance();
          continue;
        }
        const uint32_t buf = (NBUF == 1) ? (uint32_t)hgrp : (pf_g & 1u);
        mbar_expect_tx(&res_full_bar[buf], (uint32_t)p.a_box_bytes * (uint32_t)NPL);
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl)
          tma_load_4d(res_base + buf * BUF_BYTES + pl * A_STAGE_BYTES, &p.tmR, &res_full_bar[buf],
                      tcp.n0 + pf_c * 64 + pl * p.Cout, tcp.x0, tcp.y0, tcp.b);
        ++pf_g;
        pf_c += H;
        return;
      }
    };
    if (p.epi_tma && has_res && issuer) {
      for (int i = 0; i < NBUF; ++i) prefetch_res();
    }

    uint32_t t = 0, g = 0;  // local segment counter, staged-chunk counter of this group
    WorkIter wi;
    wi.init(p.sk, num_kb, num_units, unit0, ustep);
    int u, k0, k1;
    // stream-K bookkeeping: this CTA's workspace slot, and (for the head of a split unit) the slots it collects
    const int sk_slot = PAIR ? 2 * unit0 + rank : unit0;
    for (; wi.next(u, k0, k1); ++t) {
      const TileCoord tc_ = decode_unit<PAIR>(p, u, rank, BN);
      const bool sk_dump = (k0 > 0);                       // tail / middle of a unit: leave a partial tile
      const bool sk_head = (k0 == 0 && k1 < num_kb);       // head of a split unit: add the partials, then finish
      int sk_parts = 0;
      if (sk_head) {
        const long long total = (long long)num_units * num_kb, unit_end = (long long)(u + 1) * num_kb;
        while (unit0 + 1 + sk_parts < ustep && WorkIter::share_lo(total, unit0 + 1 + sk_parts, ustep) < unit_end) ++sk_parts;
      }
      const int x0 = tc_.x0, y0 = tc_.y0, b = tc_.b, n0 = tc_.n0;
      const uint32_t acc = t % (uint32_t)acc_stages;
      const uint32_t use = t / (uint32_t)acc_stages;
      const uint32_t tmem_acc = tmem_base + acc * (uint32_t)(NPL * BN) + ((uint32_t)(quad * 32) << 16);
      // bias of this tile's BN output channels -> this group's copy in shared memory (broadcast float4 reads)
      {
        const int et = (threadIdx.x - 64) & 127;   // 0..127 within the group
        for (int j = et; j < BN; j += 128) my_bias[j] = (p.bias && n0 + j < p.Cout) ? __ldg(p.bias + n0 + j) : 0.f;
      }
      mbar_wait(&tmem_full_bar[acc], use & 1u);
      tc_fence_after();
      group_sync();   // bias visible; the group's readers of the previous tile's bias are done

      if (sk_dump) {
        // ---- stream-K partial: the combined fp32 accumulator of this group's chunks goes to the CTA's workspace slot
        const int nchunks = (min(BN, p.Cout - n0) + 63) >> 6;
        float* wrow = p.sk_ws + ((size_t)sk_slot * BLOCK_M + (size_t)row) * BN;
#pragma unroll 1
        for (int c = hgrp; c < nchunks; c += H) {
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(tmem_acc + (uint32_t)(c * 64), r0);
          tmem_ld32_nowait(tmem_acc + (uint32_t)(c * 64 + 32), r1);
          if (SPLIT) {
            uint32_t q[32];
            tmem_ld32_nowait(tmem_acc + (uint32_t)(BN + c * 64), q);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r0[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q[j]), YB_LO_INV, __uint_as_float(r0[j])));
            tmem_ld32_nowait(tmem_acc + (uint32_t)(BN + c * 64 + 32), q);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r1[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q[j]), YB_LO_INV, __uint_as_float(r1[j])));
          }
          tmem_ld_wait();
          uint4* w4 = reinterpret_cast<uint4*>(wrow + c * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w4[j] = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
            w4[8 + j] = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
          }
        }
        tc_fence_before();
        __threadfence();     // the partial is visible device-wide before the flag
        group_sync();
        if (issuer) {
          release_acc(acc);
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + 2 * sk_slot + hgrp), "r"(1) : "memory");
        }
      } else if (p.epi_tma) {
        // ---- staged epilogue: TMEM -> regs -> (+bias, +residual from smem, act) -> swizzled smem tile
        //      -> one TMA store per 64-channel chunk (full 128-byte lines, OOB rows clipped by hardware)
        const int nchunks = (min(BN, p.Cout - n0) + 63) >> 6;
        if (sk_head) {   // the partial tiles of this unit's other k ranges (the following CTAs / clusters left them first thing)
          if (issuer)
            for (int j = 1; j <= sk_parts; ++j) {
              const int* f = p.sk_flags + 2 * (PAIR ? 2 * (unit0 + j) + rank : unit0 + j) + hgrp;
              int v;
              do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              } while (v == 0);
            }
          group_sync();
        }
        const int c_last = ((nchunks - 1 - hgrp) / H) * H + hgrp;   // this group's last chunk (< hgrp: none)
        if (nchunks <= hgrp) {   // nothing to read for this group in this tile
          tc_fence_before();
          if (issuer) release_acc(acc);
        }
#pragma unroll 1
        for (int c = hgrp; c < nchunks; c += H, ++g) {
          const uint32_t buf = (NBUF == 1) ? (uint32_t)hgrp : (g & 1u);
          uint8_t* out_tile = out_base + buf * BUF_BYTES;
          uint8_t* res_tile = res_base + buf * BUF_BYTES;
          if (g >= (uint32_t)NBUF) {
            if (issuer) bulk_wait_read<NBUF - 1>();  // the store that last used this staging tile has read it
            group_sync();
          }
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(tmem_acc + (uint32_t)(c * 64), r0);
          tmem_ld32_nowait(tmem_acc + (uint32_t)(c * 64 + 32), r1);
          if (SPLIT) {   // + 2^-11 * (A_lo*W_hi + A_hi*W_lo), from the tile's second accumulator
            uint32_t q[32];
            tmem_ld32_nowait(tmem_acc + (uint32_t)(BN + c * 64), q);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r0[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q[j]), YB_LO_INV, __uint_as_float(r0[j])));
            tmem_ld32_nowait(tmem_acc + (uint32_t)(BN + c * 64 + 32), q);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r1[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q[j]), YB_LO_INV, __uint_as_float(r1[j])));
          }
          tmem_ld_wait();
          if (sk_head) {
            for (int j = 1; j <= sk_parts; ++j) {
              const int slot = PAIR ? 2 * (unit0 + j) + rank : unit0 + j;
              const float4* w4 = reinterpret_cast<const float4*>(p.sk_ws + ((size_t)slot * BLOCK_M + (size_t)row) * BN + c * 64);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 a = __ldcg(w4 + i), b = __ldcg(w4 + 8 + i);
                r0[4 * i] = __float_as_uint(__uint_as_float(r0[4 * i]) + a.x);
                r0[4 * i + 1] = __float_as_uint(__uint_as_float(r0[4 * i + 1]) + a.y);
                r0[4 * i + 2] = __float_as_uint(__uint_as_float(r0[4 * i + 2]) + a.z);
                r0[4 * i + 3] = __float_as_uint(__uint_as_float(r0[4 * i + 3]) + a.w);
                r1[4 * i] = __float_as_uint(__uint_as_float(r1[4 * i]) + b.x);
                r1[4 * i + 1] = __float_as_uint(__uint_as_float(r1[4 * i + 1]) + b.y);
                r1[4 * i + 2] = __float_as_uint(__uint_as_float(r1[4 * i + 2]) + b.z);
                r1[4 * i + 3] = __float_as_uint(__uint_as_float(r1[4 * i + 3]) + b.w);
              }
            }
          }
          if (has_res) mbar_wait(&res_full_bar[buf], (NBUF == 1) ? (g & 1u) : ((g >> 1) & 1u));
          const int nbase = n0 + c * 64;
          const float* sb = my_bias + c * 64;
          const uint8_t* rt = has_res ? res_tile : nullptr;
          // res_direct (flattened 1x1 layouts only): pixel = x0 + row, both planes of the pixel are contiguous
          const __half* rg = nullptr;
          if (p.res_direct && x0 + row < p.Wo) rg = p.residual + (size_t)(x0 + row) * (size_t)(NPL * p.Cout) + nbase;
          switch (p.act) {
            case ACT_RELU: epi_chunk<ACT_RELU, false, SPLIT>(r0, r1, sb, rt, out_tile, (uint32_t)row, sw, p.out_scale, rg, p.Cout); break;
            case ACT_LEAKY:
              if (p.res_after_act)
                epi_chunk<ACT_LEAKY, true, SPLIT>(r0, r1, sb, rt, out_tile, (uint32_t)row, sw, p.out_scale, rg, p.Cout);
              else
                epi_chunk<ACT_LEAKY, false, SPLIT>(r0, r1, sb, rt, out_tile, (uint32_t)row, sw, p.out_scale, rg, p.Cout);
              break;
            case ACT_TANH: epi_chunk<ACT_TANH, false, SPLIT>(r0, r1, sb, rt, out_tile, (uint32_t)row, sw, p.out_scale, rg, p.Cout); break;
            default: epi_chunk<ACT_NONE, false, SPLIT>(r0, r1, sb, rt, out_tile, (uint32_t)row, sw, p.out_scale, rg, p.Cout); break;
          }
          fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
          if (c == c_last) tc_fence_before();
          group_sync();
          if (issuer) {
            if (c == c_last) release_acc(acc);  // all 128 threads of the group have read their rows
            tma_store_4d(&p.tmY, out_tile, nbase, x0, y0, b);
            if (SPLIT) tma_store_4d(&p.tmY, out_tile + A_STAGE_BYTES, nbase + p.Cout, x0, y0, b);
            bulk_commit();
            if (has_res) prefetch_res();   // the group is done reading res_tile[buf]: refill it
          }
        }
        if (sk_head && issuer)   // every thread of the group has read the partials (barrier above): re-arm the flags
          for (int j = 1; j <= sk_parts; ++j)
            asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + 2 * (PAIR ? 2 * (unit0 + j) + rank : unit0 + j) + hgrp),
                         "r"(0)
                         : "memory");
      } else {
        // ---- direct epilogue (fp32 / unaligned outputs: the head tensors): each warp transposes its
        //      32 rows x 32 columns through shared memory so that a store instruction writes 32
        //      consecutive channels of ONE pixel (coalesced), not one channel of 32 pixels.
        float* tbuf = reinterpret_cast<float*>(out_base) + (hgrp * 4 + quad) * (32 * 33);
        const int ncols = min(BN, p.Cout - n0);
#pragma unroll 1
        for (int c0 = hgrp * 32; c0 < ncols; c0 += 32 * H) {
          uint32_t r[32];
          tmem_ld32(tmem_acc + (uint32_t)c0, r);
          if (SPLIT) {
            uint32_t q[32];
            tmem_ld32(tmem_acc + (uint32_t)(BN + c0), q);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q[j]), YB_LO_INV, __uint_as_float(r[j])));
          }
          const int nbase = n0 + c0;
          const int nvalid = min(32, p.Cout - nbase);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          const float bias_l = my_bias[c0 + lane];
          int act = p.act;
          const int raa = p.res_after_act;
          // fused prediction head: this lane's channel belongs to one of up to 3 output tensors
          float* seg_base = nullptr;
          long long seg_bs = 0;
          int seg_ps = 0, seg_c = 0;
          if (p.nseg > 0) {
            const int n = nbase + lane;
#pragma unroll
            for (int sg = 0; sg < 3; ++sg)
              if (sg < p.nseg && n >= p.seg_begin[sg] && n < p.seg_end[sg]) {
                seg_base = p.seg_y[sg];
                seg_bs = p.seg_bs[sg];
                seg_ps = p.seg_ps[sg];
                seg_c = n - p.seg_begin[sg];
                act = p.seg_act[sg];
              }
          }
          for (int rr = 0; rr < 32; ++rr) {
            const int trow = quad * 32 + rr;
            const int ty_ = trow / p.tw, tx_ = trow - ty_ * p.tw;
            const int oy = y0 + ty_, ox = x0 + tx_;
            if (trow >= p.tw * p.th || oy >= p.Ho || ox >= p.Wo || b >= p.nb) continue;   // warp-uniform
            if (lane >= nvalid) continue;
            const long long pix = (long long)oy * p.Wo + ox;
            float v = SPLIT ? __fmaf_rn(tbuf[rr * 33 + lane], p.out_scale, bias_l) : tbuf[rr * 33 + lane] + bias_l;
            float rsd = 0.f;
            if (has_res) {
              const __half* rp = p.residual + ((long long)b * p.res_batch_stride + pix * p.Cout) * NPL + nbase + lane;
              rsd = __half2float(rp[0]);
              if (SPLIT) rsd += lo_to_f32(rp[p.Cout]);
            }
            if (!raa) v += rsd;
            v = (act == ACT_RELU) ? fmaxf(v, 0.f) : (act == ACT_TANH) ? tanhf(v) : (act == ACT_LEAKY) ? (v > 0.f ? v : 0.1f * v) : v;
            if (raa) v += rsd;
            if (p.nseg > 0) {
              if (seg_base) seg_base[(long long)b * seg_bs + pix * seg_ps + seg_c] = v;
              continue;
            }
            const long long o = (long long)b * p.y_batch_stride + pix * p.y_pix_stride + nbase + lane;
            if (p.y_f32) {
              reinterpret_cast<float*>(p.y)[o] = v;
            } else if (SPLIT) {   // y_pix_stride counts halfs and already includes both planes
              __half hi, lo;
              split_f32(v, hi, lo);
              reinterpret_cast<__half*>(p.y)[o] = hi;
              reinterpret_cast<__half*>(p.y)[o + p.Cout] = lo;
            } else {
              reinterpret_cast<__half*>(p.y)[o] = from_f32<__half>(v);
            }
          }
        }
        tc_fence_before();
        group_sync();
        if (issuer) release_acc(acc);
      }
    }
    if (p.epi_tma && issuer) bulk_wait_read<0>();  // smem must outlive the bulk reads
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // no CTA leaves while its peer's MMAs / signals may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// chain kernel: a run of consecutive convolutions (the 1x1 -> 3x3 -> 1x1 (+residual) bottleneck blocks of one ResNet
// stage, backbone.py:37-57) in ONE persistent launch
// ---------------------------------------------------------------------------------------------
// Every layer of the chain is an ordinary plan of the shape <BN = 128, single CTA, two epilogue groups, staged TMA
// epilogue, residual read from global memory>; their parameter blocks (tensor maps included) sit in an array in global
// memory.  The work units of all layers form ONE list (layer-major, then M tile, then N tile) that is dealt round-robin
// to the 148 CTAs, so a CTA moves from its last tile of layer L straight to its first tile of layer L + 1: no launch,
// no prologue (barriers, TMEM allocation, descriptor fetch), no tail where most SMs idle -- the three pipelines of a
// CTA (TMA ring, TMEM accumulators, staging tiles) simply keep running across the layer boundary.
// Dependencies are tracked per M tile instead of by a grid-wide barrier: the epilogue groups bump done[layer][m tile]
// once their stores of a tile have completed; before the first ACTIVATION load of a tile the producer warp waits until
// every tile of the layer that produces its input which overlaps the input rows the tile reads (its own rows mapped
// through stride / padding / kernel height), and every tile of the layer that produces its residual which overlaps its
// own rows, has been finished by all its N tiles (the weight tiles of the first k-blocks are requested before the
// wait).  Layers may differ in resolution (stride-2 layers), N tile (64 / 128) and tile shape.  All dependencies point
// backwards in the unit list and every CTA walks its units in list order, so the earliest unfinished unit can always
// run: no deadlock as long as the CTAs are co-resident (cooperative launch, one CTA per SM).
// Every activation has its own buffer (engine.cu alloc_act: nothing is recycled inside a forward pass), so there are
// no write-after-read hazards to order.
struct ChainLayer {
  int ubase;               // position of this layer's unit 0 in the chain's unit list
  int units;               // m_tiles * n_tiles
  int flat;                // 1: tile m = pixels [m * tw, m * tw + tw) of the flattened B*H*W axis; 0: tiles_x x tiles_y tiles per image
  int W, H, rows;          // OUTPUT image size; rows = B * H  (global row r = b * H + y, pixel = r * W + x)
  int tw, th, tiles_x, tiles_y;
  int done_off;            // this layer's counters: done[done_off + m_tile]
  int target;              // a finished M tile: n_tiles * 2 (each epilogue group of each N tile adds 1)
  int dep_a;               // chain layer that writes this layer's input (-1: a tensor complete before the launch)
  int dep_r;               // chain layer that writes this layer's residual (-1: none / outside / implied by dep_a's own dependencies)
  int stride, pad, kh;     // output row y reads input rows [y * stride - pad, y * stride - pad + kh)
};

// The dependency arithmetic, shared by the kernel and a host mirror (yb_debug_chain_deps: CPU property tests).
// Global output rows [r0, r1] of M tile m of layer ci
__host__ __device__ __forceinline__ void chain_rows_of_tile(const ChainLayer& ci, int m, int& r0, int& r1) {
  if (ci.flat) {
    const int lo = m * ci.tw;
    int hi = lo + ci.tw;
    if (hi > ci.rows * ci.W) hi = ci.rows * ci.W;
    r0 = lo / ci.W;
    r1 = (hi - 1) / ci.W;
  } else {
    const int ty = (m / ci.tiles_x) % ci.tiles_y, b = m / (ci.tiles_x * ci.tiles_y);
    int y1 = ty * ci.th + ci.th;
    if (y1 > ci.H) y1 = ci.H;
    r0 = b * ci.H + ty * ci.th;
    r1 = b * ci.H + y1 - 1;
  }
}
// Rows [ra, rb] of the input tensor (written by layer pa: its image height differs under a stride) that output rows
// [r0, r1] of layer ci read
__host__ __device__ __forceinline__ void chain_input_rows(const ChainLayer& ci, const ChainLayer& pa, int r0, int r1, int& ra, int& rb) {
  const int b0 = r0 / ci.H, y0 = r0 - b0 * ci.H, b1 = r1 / ci.H, y1 = r1 - b1 * ci.H;
  int lo = y0 * ci.stride - ci.pad, hi = y1 * ci.stride - ci.pad + ci.kh - 1;
  if (lo < 0) lo = 0;
  if (hi > pa.H - 1) hi = pa.H - 1;
  ra = b0 * pa.H + lo;
  rb = b1 * pa.H + hi;
}
// M tiles of layer `pi` that overlap its global rows [ra, rb]
__host__ __device__ __forceinline__ void chain_tiles_of_rows(const ChainLayer& pi, int ra, int rb, int& ia, int& ib) {
  if (pi.flat) {
    ia = (ra * pi.W) / pi.tw;
    ib = ((rb + 1) * pi.W - 1) / pi.tw;
  } else {
    const int ba = ra / pi.H, ya = ra - ba * pi.H, bb = rb / pi.H, yb = rb - bb * pi.H;
    ia = (ba * pi.tiles_y + ya / pi.th) * pi.tiles_x;
    ib = (bb * pi.tiles_y + yb / pi.th) * pi.tiles_x + pi.tiles_x - 1;
  }
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// H epilogue groups of 4 warps; each group owns an 8 KB (x 2 planes) staging buffer: H = 2 leaves room for three
// split-precision stages, H = 4 for two
__host__ __device__ constexpr int chain_stages(bool split, int h) { return split ? (h == 4 ? 2 : 3) : 6; }
__device__ __forceinline__ bool mbar_test_wait(uint32_t addr, uint32_t parity) {   // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}

// first unit of a layer that belongs to CTA `cta` when the chain's unit list is dealt round-robin over G CTAs
__device__ __forceinline__ int chain_first_unit(int cta, int ubase, int G) {
  int r = (cta - ubase) % G;
  return r < 0 ? r + G : r;
}

template <bool SPLIT, int H>
__global__ void __launch_bounds__(64 + 128 * H)
tc_chain_kernel(const TcParams* __restrict__ layers, const ChainLayer* __restrict__ info, int nl, int* __restrict__ done,
                long long* __restrict__ stats) {
  // stats (diagnostics, YB_CHAIN_STATS=1; null otherwise): per CTA 8 counters of SM cycles spent waiting --
  // [0] producer: dependency counters, [1] producer: free ring slot, [2] MMA issuer: operands, [3] MMA issuer: free
  // accumulator, [4] epilogue group 0: accumulator, [5] epilogue group 0: store completion before a signal, [6] whole kernel
  const long long t_kernel0 = stats ? clock64() : 0;
  long long w0 = 0, w1 = 0;
  constexpr int BN = 128;
  constexpr int NPL = SPLIT ? 2 : 1;
  constexpr int B_PLANE_BYTES = BN * BLOCK_K * 2;
  constexpr int A_BYTES = NPL * A_STAGE_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + NPL * B_PLANE_BYTES;
  constexpr int BUF_BYTES = NPL * CHUNK32_BYTES;
  // the chain's own shared-memory layout (the plans' stage counts / offsets are not used): ring, then one staging
  // buffer per epilogue group; two TMEM accumulator buffers.  The epilogue moves 32-channel chunks (8 KB tiles), which
  // leaves room for a THIRD split-precision stage (3 x 64 KB + 32 KB).  Measured (profiles/r2_call14_summary.txt): the
  // 105-layer trunk chain 2.14 -> 1.92 ms at batch 2, unchanged (3.59 ms) at batch 8, where the MMA issuer's operand
  // waits are set by the L2 -> SM rate rather than by the bytes a CTA keeps in flight
  constexpr int stages = chain_stages(SPLIT, H);

  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float sbias[H * BN];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* out_base = smem + stages * STAGE_BYTES;
  const int G = (int)gridDim.x;
  const int cta = (int)blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], H);
    }
    fence_barrier_init();
    tma_prefetch_desc(&layers[0].tmB);
    tma_prefetch_desc(&layers[0].tmA[0]);
    tma_prefetch_desc(&layers[0].tmY32);
  }
  if (warp == 1) tmem_alloc_dyn(&s_tmem_base, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (lane 0 issues; the whole warp polls the dependency counters) =====================
    uint32_t kbg = 0;
    for (int L = 0; L < nl; ++L) {
      const TcParams& p = layers[L];
      const ChainLayer ci = info[L];
      const int num_kb = p.ntaps * p.kchunks;
      const int bn = p.bn;
      const uint32_t tx_bytes = ((uint32_t)p.a_box_bytes + (uint32_t)(bn * BLOCK_K * 2)) * (uint32_t)NPL;
      if (lane == 0 && L + 1 < nl) {   // the next layer's descriptors: fetched long before their first use
        tma_prefetch_desc(&layers[L + 1].tmB);
        tma_prefetch_desc(&layers[L + 1].tmA[0]);
      }
      const bool nodeps = (ci.dep_a < 0 && ci.dep_r < 0);
      // the counters M tile m of this layer waits for: entry i of the list (input tiles first, then residual tiles);
      // returns the length of the list
      auto dep_of = [&](int m, int i, const int*& c, int& tgt) -> int {
        int r0, r1;
        chain_rows_of_tile(ci, m, r0, r1);
        int ia = 0, ib = -1, ja = 0, jb = -1, ta = 0, tr = 0;
        const int *ca = done, *cr = done;
        if (ci.dep_a >= 0) {   // input rows, in the producer's row numbering (its image height differs under a stride)
          const ChainLayer pa = info[ci.dep_a];
          int ra, rb;
          chain_input_rows(ci, pa, r0, r1, ra, rb);
          chain_tiles_of_rows(pa, ra, rb, ia, ib);
          ca = done + pa.done_off;
          ta = pa.target;
        }
        if (ci.dep_r >= 0) {   // the residual has this layer's geometry
          const ChainLayer pr = info[ci.dep_r];
          chain_tiles_of_rows(pr, r0, r1, ja, jb);
          cr = done + pr.done_off;
          tr = pr.target;
        }
        const int na = ib - ia + 1, nr = jb - ja + 1;
        if (i < na) {
          c = ca + ia + i;
          tgt = ta;
        } else {
          c = cr + ja + (i - na);
          tgt = tr;
        }
        return na + nr;
      };
      // (Reading the NEXT tile's counters ahead of time, so that a satisfied check costs no L2 round trip in front of the
      // tile, was tried: the producer's waiting share fell from 30 % to 18 %, the kernel time did not move -- 3.745 vs
      // 3.740 ms, profiles/r2_call19_summary.txt: the MMA issuer waits for the operands themselves.  Removed.)
      for (int u = chain_first_unit(cta, ci.ubase, G); u < ci.units; u += G) {
        const TileCoord tc_ = decode_unit<false>(p, u, 0, bn);
        bool ready = nodeps;
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const uint32_t s = kbg % (uint32_t)stages;
          const uint32_t it = kbg / (uint32_t)stages;
          const int tap = kb / p.kchunks;
          const int kc = kb - tap * p.kchunks;
          uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (lane == 0) {
            const long long t0 = stats ? clock64() : 0;
            mbar_wait(&empty_bar[s], (it & 1u) ^ 1u);
            if (stats) w1 += clock64() - t0;
            mbar_expect_tx(&full_bar[s], tx_bytes);
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl)   // weights: constants, no dependency
              tma_load_3d(sb + pl * B_PLANE_BYTES, &p.tmB, &full_bar[s], kc * BLOCK_K + pl * p.cin, tc_.n0, tap);
          }
          if (!ready) {
            const long long tdep0 = stats ? clock64() : 0;
            __syncwarp();
            const int* c = done;
            int tgt = 0;
            const int n = dep_of(u / p.n_tiles, lane, c, tgt);
            for (int i = lane; i < n; i += 32) {
              if (i != lane) dep_of(u / p.n_tiles, i, c, tgt);
#ifdef YB_WATCHDOG
              const long long t0 = clock64();
#endif
              while (ld_acquire_gpu(c) < tgt) {
#ifdef YB_WATCHDOG
                if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
#endif
              }
            }
            __syncwarp();
            if (stats) w0 += clock64() - tdep0;
            fence_proxy_async_all();   // the tiles were written through the async proxy (TMA stores) and are read through it
            ready = true;
          }
          if (lane == 0) {
            const CUtensorMap* ma = &p.tmA[p.tap_map[tap]];
            const int ax = tc_.x0 + p.tap_dx[tap], ay = tc_.y0 + p.tap_dy[tap];
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl)
              tma_load_4d(sa + pl * A_STAGE_BYTES, ma, &full_bar[s], kc * BLOCK_K + pl * p.cin, ax, ay, tc_.b);
          }
        }
      }
    }
    if (stats && lane == 0) {
      stats[cta * 8 + 0] = w0;
      stats[cta * 8 + 1] = w1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      uint32_t kbg = 0, t = 0;
      for (int L = 0; L < nl; ++L) {
        const TcParams& p = layers[L];
        const int num_kb = p.ntaps * p.kchunks;
        const uint32_t idesc = p.idesc;
        const int ubase = info[L].ubase, units = info[L].units;
        for (int u = chain_first_unit(cta, ubase, G); u < units; u += G, ++t) {
          const uint32_t acc = t & 1u;
          const uint32_t use = t >> 1;
          long long t0 = stats ? clock64() : 0;
          mbar_wait(&tmem_empty_bar[acc], (use & 1u) ^ 1u);
          if (stats) w1 += clock64() - t0;
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * (uint32_t)(NPL * BN);
          for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
            const uint32_t s = kbg % (uint32_t)stages;
            const uint32_t it = kbg / (uint32_t)stages;
            t0 = stats ? clock64() : 0;
            mbar_wait(&full_bar[s], it & 1u);
            if (stats) w0 += clock64() - t0;
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
            const uint32_t sb = sa + A_BYTES;
            const uint64_t da = make_sw128_desc(sa);
            const uint64_t db = make_sw128_desc(sb);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (SPLIT) {
              const uint64_t dal = make_sw128_desc(sa + A_STAGE_BYTES);
              const uint64_t dbl = make_sw128_desc(sb + B_PLANE_BYTES);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)   // A_lo * W_hi -> second accumulator
                umma_f16(tmem_d + BN, dal + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)   // A_hi * W_lo -> second accumulator
                umma_f16(tmem_d + BN, da + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), idesc, 1u);
            }
            umma_commit(&empty_bar[s]);
          }
          umma_commit(&tmem_full_bar[acc]);
        }
      }
      if (stats) {
        stats[cta * 8 + 2] = w0;
        stats[cta * 8 + 3] = w1;
      }
    }
  } else {
    // ===================== epilogue: H groups of 4 warps, group g takes the 32-channel chunks g, g + H, ... =====================
    const int quad = warp & 3;
    const int hgrp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const bool issuer = (warp == 2 + 4 * hgrp && lane == 0);
    float* my_bias = sbias + hgrp * BN;
    uint8_t* out_tile = out_base + hgrp * BUF_BYTES;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(hgrp + 1) : "memory"); };
    // issuer only: counter to bump once this thread's outstanding stores (the previous tile's) have completed
    int* pending = nullptr;
    auto flush_pending = [&]() {
      const long long t0 = stats ? clock64() : 0;
      bulk_wait_all();            // the stores have been performed, not just read out of shared memory
      if (stats) w1 += clock64() - t0;
      fence_proxy_async_all();
      red_release_gpu_add(pending, 1);
      pending = nullptr;
    };
    uint32_t t = 0, g = 0;
    for (int L = 0; L < nl; ++L) {
      const TcParams& p = layers[L];
      const int ubase = info[L].ubase, units = info[L].units, done_off = info[L].done_off;
      const int act = p.act, raa = p.res_after_act, Cout = p.Cout, res_direct = p.res_direct, Wo = p.Wo;
      const float out_scale = p.out_scale;
      const float* bias = p.bias;
      const __half* residual = p.residual;
      const int n_tiles = p.n_tiles;
      const int bn = p.bn;
      for (int u = chain_first_unit(cta, ubase, G); u < units; u += G, ++t) {
        const TileCoord tc_ = decode_unit<false>(p, u, 0, bn);
        const int x0 = tc_.x0, y0 = tc_.y0, b = tc_.b, n0 = tc_.n0;
        const uint32_t acc = t & 1u;
        const uint32_t use = t >> 1;
        const uint32_t tmem_acc = tmem_base + acc * (uint32_t)(NPL * BN) + ((uint32_t)(quad * 32) << 16);
        {
          const int et = (threadIdx.x - 64) & 127;
          for (int j = et; j < bn; j += 128) my_bias[j] = (bias && n0 + j < Cout) ? __ldg(bias + n0 + j) : 0.f;
        }
        // The previous tile's completion signal waits for its stores; that is free when this tile's accumulator is
        // not ready yet (the group would idle anyway).  When it IS ready the signal is deferred until the first chunk
        // of this tile has been computed (by then the stores have long landed) -- safe, because a ready accumulator
        // means this tile depends on nothing that could be waiting for the deferred signal.
        if (issuer && pending && !mbar_test_wait(smem_u32(&tmem_full_bar[acc]), use & 1u)) flush_pending();
        const long long tacc0 = stats ? clock64() : 0;
        mbar_wait(&tmem_full_bar[acc], use & 1u);
        if (stats) w0 += clock64() - tacc0;
        tc_fence_after();
        group_sync();
        const int nchunks = (min(bn, Cout - n0) + 31) >> 5;   // 32-channel chunks: 4, or 2 for a 64-wide N tile
        const int c_last = ((nchunks - 1 - hgrp) / H) * H + hgrp;
        if (nchunks <= hgrp) {   // nothing for this group, which only hands the accumulator back
          tc_fence_before();
          if (issuer) {
            mbar_arrive(&tmem_empty_bar[acc]);
            if (pending) flush_pending();
          }
        }
#pragma unroll 1
        for (int c = hgrp; c < nchunks; c += H, ++g) {
          if (g >= 1u) {
            if (issuer) bulk_wait_read<0>();   // the store that last used this group's staging tile has read it
            group_sync();
          }
          uint32_t r[32];
          tmem_ld32_nowait(tmem_acc + (uint32_t)(c * 32), r);
          if (SPLIT) {
            uint32_t q[32];
            tmem_ld32_nowait(tmem_acc + (uint32_t)(BN + c * 32), q);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q[j]), YB_LO_INV, __uint_as_float(r[j])));
          } else {
            tmem_ld_wait();
          }
          const int nbase = n0 + c * 32;
          const float* sb = my_bias + c * 32;
          const __half* rg = nullptr;
          if (res_direct && x0 + row < Wo) rg = residual + (size_t)(x0 + row) * (size_t)(NPL * Cout) + nbase;
          switch (act) {
            case ACT_RELU: epi_chunk32<ACT_RELU, false, SPLIT>(r, sb, out_tile, (uint32_t)row, out_scale, rg, Cout); break;
            case ACT_LEAKY:
              if (raa)
                epi_chunk32<ACT_LEAKY, true, SPLIT>(r, sb, out_tile, (uint32_t)row, out_scale, rg, Cout);
              else
                epi_chunk32<ACT_LEAKY, false, SPLIT>(r, sb, out_tile, (uint32_t)row, out_scale, rg, Cout);
              break;
            default: epi_chunk32<ACT_NONE, false, SPLIT>(r, sb, out_tile, (uint32_t)row, out_scale, rg, Cout); break;
          }
          fence_proxy_async();
          if (issuer && pending) flush_pending();   // deferred signal of the previous tile
          if (c == c_last) tc_fence_before();
          group_sync();
          if (issuer) {
            if (c == c_last) mbar_arrive(&tmem_empty_bar[acc]);
            tma_store_4d(&p.tmY32, out_tile, nbase, x0, y0, b);
            if (SPLIT) tma_store_4d(&p.tmY32, out_tile + CHUNK32_BYTES, nbase + Cout, x0, y0, b);
            bulk_commit();
          }
        }
        if (issuer) {
          if (nchunks > hgrp) pending = done + done_off + u / n_tiles;
          else red_release_gpu_add(done + done_off + u / n_tiles, 1);   // nothing stored: counts at once
        }
      }
    }
    if (issuer) {
      if (pending) flush_pending();
      bulk_wait_read<0>();
      if (stats && hgrp == 0) {
        stats[cta * 8 + 4] = w0;
        stats[cta * 8 + 5] = w1;
        stats[cta * 8 + 6] = clock64() - t_kernel0;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

namespace {
PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  YB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
  YB_REQUIRE(qres == cudaDriverEntryPointSuccess && ptr, "cuTensorMapEncodeTiled not available from the driver");
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

}  // namespace

void tc::encode_map_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    YB_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map: box dim out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    YB_REQUIRE(strides_bytes[i] % 16 == 0, "tensor map: stride must be a multiple of 16 bytes");
  }
  YB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base must be 16-byte aligned");
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                               gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(YB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
}


struct TcConvPlan {
  TcParams prm;
  int split = 0;
  int sk = 0;           // stream-K requested (effective once a workspace is attached: prm.sk)
  int BN = 128;
  int pair = 0;
  int epi_groups = 1;   // H: 4-warp epilogue groups per CTA
  int pdl_friendly = 0; // sized so that two CTAs (this kernel's and the next layer's) fit on one SM
  int flat = 0;         // 1x1 / stride 1 / dense: all pixels of the batch on one axis
  int B = 0, Ho = 0, Wo = 0;   // the problem's output geometry (prm.Ho / prm.Wo are the flattened view's)
  int Hi = 0, Wi = 0, stride = 1, pad = 0, KH = 1;
  dim3 grid;
  size_t smem_bytes = 0;
};

bool tc_conv_supported(const ConvProblem& p) {
  if (p.x_nchw_f32) return false;
  if (p.Cin % BLOCK_K != 0) return false;
  if (p.KH * p.KW > MAX_TAPS) return false;
  if (p.stride != 1 && p.stride != 2) return false;
  if (p.Cout < 1) return false;
  return true;
}

// spatial tile tw x th <= 128 accumulator rows with the best fill (ties: the wider tile)
static void choose_tile(int Wov, int Hov, int& best_tw, int& best_th) {
  best_tw = 1;
  best_th = 1;
  double best_eff = -1.0;
  for (int tw = 1; tw <= std::min(Wov, 128); ++tw) {
    int th = std::min(Hov, 128 / tw);
    if (th < 1) continue;
    if (tw > 256 || th > 256) continue;
    long long tiles = (long long)ceil_div(Wov, tw) * ceil_div(Hov, th);
    double eff = (double)Wov * Hov / ((double)tiles * 128.0);
    if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && tw > best_tw)) {
      best_eff = eff;
      best_tw = tw;
      best_th = th;
    }
  }
}

// Host mirror of the chain kernel's dependency arithmetic for ONE consumer layer (a k x k conv, stride, pad, on a
// B x Hin x Win input written by a producer layer that is flattened (1x1 stride 1) or not): the tilings both layers
// would get and, for consumer M tile m, the inclusive range of producer M tiles the kernel would wait for.
// out[0..5] = consumer flat, tw, th, tiles_x, tiles_y, m_tiles; out[6..11] = producer flat, tw, th, tiles_x, tiles_y,
// m_tiles; out[12], out[13] = first, last producer tile.  No device needed.
void tc_chain_debug_deps(int B, int Hin, int Win, int k, int stride, int pad, int producer_flat, int m, int32_t* out) {
  YB_REQUIRE(B >= 1 && Hin >= 1 && Win >= 1 && (k == 1 || k == 3) && (stride == 1 || stride == 2) && pad >= 0 && out,
             "chain_debug_deps: bad argument");
  const int Ho = (Hin + 2 * pad - k) / stride + 1, Wo = (Win + 2 * pad - k) / stride + 1;
  YB_REQUIRE(Ho >= 1 && Wo >= 1, "chain_debug_deps: empty output");
  auto layer = [&](int flat, int H, int W, int s, int p, int kh) {
    ChainLayer c = {};
    c.flat = flat;
    c.W = W;
    c.H = H;
    c.rows = B * H;
    const int Wov = flat ? B * H * W : W, Hov = flat ? 1 : H;
    choose_tile(Wov, Hov, c.tw, c.th);
    c.tiles_x = ceil_div(Wov, c.tw);
    c.tiles_y = ceil_div(Hov, c.th);
    c.stride = s;
    c.pad = p;
    c.kh = kh;
    return c;
  };
  const ChainLayer ci = layer((k == 1 && stride == 1 && pad == 0) ? 1 : 0, Ho, Wo, stride, pad, k);
  const ChainLayer pa = layer(producer_flat ? 1 : 0, Hin, Win, 1, 0, 1);
  const int cm = ci.tiles_x * ci.tiles_y * (ci.flat ? 1 : B), pm = pa.tiles_x * pa.tiles_y * (pa.flat ? 1 : B);
  YB_REQUIRE(m >= 0 && m < cm, "chain_debug_deps: tile out of range");
  int r0, r1, ra, rb, ia, ib;
  chain_rows_of_tile(ci, m, r0, r1);
  chain_input_rows(ci, pa, r0, r1, ra, rb);
  chain_tiles_of_rows(pa, ra, rb, ia, ib);
  const int32_t v[14] = {ci.flat, ci.tw, ci.th, ci.tiles_x, ci.tiles_y, cm, pa.flat, pa.tw, pa.th, pa.tiles_x, pa.tiles_y, pm, ia, ib};
  for (int i = 0; i < 14; ++i) out[i] = v[i];
}

TcConvPlan* tc_conv_plan_create(const ConvProblem& p, const __half* w_packed, int bn_override, int stages_override,
                                int grid_override, int pair_override, int epi_override, int pdl_override, int sk_override,
                                int chain_override) {
  YB_REQUIRE(tc_conv_supported(p), "tc_conv: unsupported problem");
  YB_REQUIRE(p.Ho == (p.H + 2 * p.pad - p.KH) / p.stride + 1, "tc_conv: bad Ho");
  YB_REQUIRE(p.Wo == (p.W + 2 * p.pad - p.KW) / p.stride + 1, "tc_conv: bad Wo");
  auto* plan = new TcConvPlan();
  TcParams& q = plan->prm;
  memset(&q, 0, sizeof(q));
  const int s = p.stride;
  q.ntaps = p.KH * p.KW;
  q.kchunks = p.Cin / BLOCK_K;
  const int split = p.split ? 1 : 0;
  const int npl = split ? 2 : 1;       // fp16 planes per operand / activation
  plan->split = split;
  q.split = split;
  q.cin = p.Cin;
  q.out_scale = split ? p.out_scale : 1.f;

  // ---- geometry: can the whole problem be flattened into one pixel axis? (1x1, stride 1, dense out)
  const bool dense_out = (p.y_batch_stride == (int64_t)p.Ho * p.Wo * p.y_pix_stride);
  YB_REQUIRE(!split || p.y_f32 || p.nseg > 0 || p.y_pix_stride >= 2 * p.Cout, "tc_conv: split outputs need 2*Cout halfs per pixel");
  const bool flat = (q.ntaps == 1 && s == 1 && p.pad == 0 && dense_out);
  int Bv = p.B, Hov = p.Ho, Wov = p.Wo;
  if (flat) {
    Wov = p.B * p.Ho * p.Wo;
    Hov = 1;
    Bv = 1;
  }
  // ---- spatial tile (tw x th <= 128) with the best fill
  int best_tw = 1, best_th = 1;
  choose_tile(Wov, Hov, best_tw, best_th);
  plan->flat = flat ? 1 : 0;
  plan->B = p.B;
  plan->Ho = p.Ho;
  plan->Wo = p.Wo;
  plan->Hi = p.H;
  plan->Wi = p.W;
  plan->stride = p.stride;
  plan->pad = p.pad;
  plan->KH = p.KH;
  q.tw = best_tw;
  q.th = best_th;
  q.tiles_x = ceil_div(Wov, q.tw);
  q.tiles_y = ceil_div(Hov, q.th);
  q.Ho = Hov;
  q.Wo = Wov;
  q.Cout = p.Cout;
  q.a_box_bytes = q.tw * q.th * BLOCK_K * 2;
  const long long m_tiles = (long long)q.tiles_x * q.tiles_y * Bv;

  // ---- epilogue mode: fp16 NHWC outputs with 16-byte aligned rows go through smem + TMA store
  const int esz = p.y_f32 ? 4 : 2;
  const int vec_elems = 16 / esz;
  bool vec_ok = (p.y_pix_stride % vec_elems == 0) && (p.y_batch_stride % vec_elems == 0) &&
                ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
  if (p.residual) vec_ok = vec_ok && (p.Cout % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
  q.vec_ok = vec_ok ? 1 : 0;
  q.epi_tma = (!p.y_f32 && vec_ok && p.Cout % 8 == 0 && p.nseg == 0) ? 1 : 0;
  // split: a 64-channel hi box must not run into the lo plane of the same pixel (nothing clips it there)
  if (split && q.epi_tma && p.Cout % 64 != 0) q.epi_tma = 0;
  // the staged epilogue moves 64-channel boxes: a CTA must own at least 64 output channels
  const int bn_min = q.epi_tma ? 64 : 32;

  // ---- N tile: minimise waves * (BN + fixed cost)
  {
    const int cands[4] = {256, 128, 64, 32};
    double best = 1e30;
    int bn_best = 32;
    for (int i = 0; i < 4; ++i) {
      int bn = cands[i];
      if (bn < bn_min) continue;
      if (bn > bn_min && bn >= 2 * p.Cout) continue;  // more than half the tile would be padding
      long long ctas = m_tiles * ceil_div(p.Cout, bn);
      long long waves = (ctas + 147) / 148;
      double cost = (double)waves * (bn + 64);
      if (cost < best - 1e-9) {
        best = cost;
        bn_best = bn;
      }
    }
    plan->BN = bn_best;
    if (bn_override == 32 || bn_override == 64 || bn_override == 128 || bn_override == 256)
      plan->BN = std::max(bn_override, bn_min);
  }
  // ---- CTA pairs: two adjacent M tiles per (cluster of 2), each CTA stages half of the weight tile
  const int pair = (pair_override > 0 && m_tiles >= 2 && plan->BN >= 64) ? 1 : 0;
  int epi_req = (epi_override == 2 && plan->BN >= 64) ? 2 : 1;
  if (split) {
    // every stage and every epilogue buffer is twice as large: step down (second epilogue group first, then the
    // N tile) until at least two pipeline stages fit
    auto stages_for = [&](int bn, int hg) {
      const int sb = 2 * (A_STAGE_BYTES + (pair ? bn / 2 : bn) * BLOCK_K * 2);
      const int tiles = 2 * hg;
      const int ob = std::max(tiles * A_STAGE_BYTES, hg == 2 ? 36 * 1024 : 2 * A_STAGE_BYTES);
      // two groups on a flattened layout read the residual from global memory (no staging tiles), see res_direct
      const int rb = (q.epi_tma && p.residual && !(hg == 2 && flat)) ? tiles * A_STAGE_BYTES : 0;
      return (221 * 1024 - ob - rb) / sb;
    };
    while (stages_for(plan->BN, epi_req) < 2) {
      if (epi_req == 2) epi_req = 1;
      else if (plan->BN > std::max(bn_min, pair ? 64 : 32)) plan->BN /= 2;
      else break;
    }
  }
  const int BN = plan->BN;
  q.bn = BN;
  plan->pair = pair;
  q.pair = pair;
  q.nb = Bv;
  const int stage_bytes = npl * (A_STAGE_BYTES + (pair ? BN / 2 : BN) * BLOCK_K * 2);
  // ---- persistent grid + shared-memory layout: [pipeline stages][2 x 16 KB out tiles][2 x 16 KB residual tiles]
  q.m_tiles = (int)m_tiles;
  q.n_tiles = ceil_div(p.Cout, BN);
  const int num_tiles = (pair ? (q.m_tiles + 1) / 2 : q.m_tiles) * q.n_tiles;   // work units (pairs of M tiles when paired)
  int grid = std::min(num_tiles, grid_override > 0 ? (pair ? std::max(1, grid_override / 2) : grid_override) : (pair ? 74 : 148));
  if (pair) grid = std::min(grid, 74);   // one cluster per TPC: 148 SMs = 74 CTA pairs, one CTA per SM
  // stream-K: one share of the k-block iterations per SM (cluster), whatever the unit count -- worth it only when the
  // units do not fill whole waves; needs the staged (TMA) epilogue and co-resident CTAs (<= one per SM)
  const long long total_it = (long long)num_tiles * q.ntaps * q.kchunks;
  int sk = (sk_override > 0 && q.epi_tma && !(pdl_override > 0)) ? 1 : 0;
  if (sk) {
    const int gsk = std::min<long long>(grid_override > 0 ? grid : (pair ? 74 : 148), total_it);
    if (gsk <= 1 || num_tiles % gsk == 0) sk = 0;   // whole waves already: nothing to balance
    else grid = gsk;
  }
  plan->sk = sk;
  q.acc_stages = (grid < num_tiles || sk) ? 2 : 1;
  if (q.acc_stages * npl * BN > 512) q.acc_stages = 1;   // split: two accumulators per tile (2 * BN columns)
  int tmem_cols = 32;
  while (tmem_cols < q.acc_stages * npl * BN) tmem_cols *= 2;
  // Pairs allocate all of TMEM (and > half of the shared memory, below): one CTA per SM, so both CTAs of a pair
  // get the SAME accumulator address, which the single cta_group::2 MMA requires.
  if (pair) tmem_cols = 512;
  q.tmem_cols = tmem_cols;
  const int tiles_per_cta = sk ? 2 : ceil_div(num_tiles, grid);
  // two epilogue groups (8 warps) work on two 64-channel chunks at once; pointless for a single-chunk tile
  // PDL-friendly plan: <= ~108 KB of shared memory, one epilogue group (192 threads x 122 registers) and <= 256 TMEM
  // columns, so a CTA of the NEXT layer can become resident beside it and overlap its prologue + first weight tiles
  const int pdlf = (pdl_override > 0 && !pair && !split) ? 1 : 0;
  plan->pdl_friendly = pdlf;
  if (pdlf && q.acc_stages * BN > 256) q.acc_stages = 1;
  if (pdlf) {   // (never a split plan)
    int tc = 32;
    while (tc < q.acc_stages * BN) tc *= 2;
    q.tmem_cols = tc;
  }
  plan->epi_groups = (epi_req == 2 && !pdlf) ? 2 : 1;
  // staging: 2 x 16 KB tiles (one per group when there are two); the direct (fp32) epilogue needs a padded
  // 32x33 float transpose buffer per epilogue warp
  // (split: one buffer of two tiles per group)
  const int epi_tiles = split ? 2 * plan->epi_groups : 2;
  const int out_bytes = std::max(epi_tiles * A_STAGE_BYTES, plan->epi_groups == 2 ? 36 * 1024 : 2 * A_STAGE_BYTES);
  q.res_direct = ((split || chain_override > 0) && plan->epi_groups == 2 && p.residual && flat && q.epi_tma) ? 1 : 0;
  const int res_bytes = (q.epi_tma && p.residual && !q.res_direct) ? epi_tiles * A_STAGE_BYTES : 0;
  int stages = std::min(MAX_STAGES, ((pdlf ? 108 : (split ? 221 : 200)) * 1024 - out_bytes - res_bytes) / stage_bytes);
  if (stages < 1 && pdlf) {   // does not fit in half an SM: an ordinary plan
    plan->pdl_friendly = 0;
    stages = std::min(MAX_STAGES, (200 * 1024 - out_bytes - res_bytes) / stage_bytes);
  }
  YB_REQUIRE(stages >= 1, "tc_conv: tile does not fit in shared memory");
  if (stages_override > 0) stages = std::min(stages, stages_override);
  stages = std::max(1, std::min(stages, q.ntaps * q.kchunks * tiles_per_cta));
  q.stages = stages;
  q.out_off = stages * stage_bytes;
  q.res_off = q.out_off + out_bytes;
  plan->smem_bytes = (size_t)q.res_off + res_bytes + 1024;
  if (plan->pdl_friendly && plan->smem_bytes > (size_t)112 * 1024) plan->pdl_friendly = 0;   // does not fit twice: plain plan
  q.pdl = plan->pdl_friendly;
  if (pair) plan->smem_bytes = std::max(plan->smem_bytes, (size_t)120 * 1024);   // at most one pair CTA per SM
  plan->grid = dim3((unsigned)(pair ? 2 * grid : grid), 1, 1);

  // ---- instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major, N, M=128 (256 for a CTA pair)
  q.idesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) |
            ((uint32_t)((pair ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);

  // ---- A tensor maps
  const __half* x = reinterpret_cast<const __half*>(p.x);
  const uint64_t CinP = (uint64_t)npl * p.Cin;    // halfs per input pixel (both planes)
  const uint64_t CoutP = (uint64_t)npl * p.Cout;  // halfs per residual pixel
  if (flat) {
    uint64_t dims[4] = {CinP, (uint64_t)Wov, 1, 1};
    uint64_t str[3] = {CinP * 2, (uint64_t)Wov * CinP * 2, (uint64_t)Wov * CinP * 2};
    uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)q.tw, (uint32_t)q.th, 1};
    encode_map_f16(&q.tmA[0], x, 4, dims, str, box);
    q.tap_map[0] = 0;
    q.tap_dx[0] = 0;
    q.tap_dy[0] = 0;
  } else {
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < p.KH; ++r)
      for (int c = 0; c < p.KW; ++c) {
        int qy = r - p.pad, qx = c - p.pad;
        int py = ((qy % s) + s) % s, px = ((qx % s) + s) % s;
        int tap = r * p.KW + c;
        q.tap_map[tap] = py * s + px;
        q.tap_dy[tap] = floordiv(qy - py, s);
        q.tap_dx[tap] = floordiv(qx - px, s);
        used[py * s + px] = true;
      }
    for (int py = 0; py < s; ++py)
      for (int px = 0; px < s; ++px) {
        if (!used[py * s + px]) continue;
        int Hv = (p.H - py + s - 1) / s, Wv = (p.W - px + s - 1) / s;
        YB_REQUIRE(Hv >= 1 && Wv >= 1, "tc_conv: empty phase view");
        const __half* base = x + ((size_t)py * p.W + px) * CinP;
        uint64_t dims[4] = {CinP, (uint64_t)Wv, (uint64_t)Hv, (uint64_t)p.B};
        uint64_t str[3] = {(uint64_t)s * CinP * 2, (uint64_t)s * p.W * CinP * 2,
                           (uint64_t)p.H * p.W * CinP * 2};
        uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)q.tw, (uint32_t)q.th, 1};
        encode_map_f16(&q.tmA[py * s + px], base, 4, dims, str, box);
      }
  }
  // ---- B tensor map: [tap][Cout][Cin]
  {
    uint64_t dims[3] = {CinP, (uint64_t)p.Cout, (uint64_t)q.ntaps};
    uint64_t str[2] = {CinP * 2, (uint64_t)p.Cout * CinP * 2};
    uint32_t box[3] = {(uint32_t)BLOCK_K, (uint32_t)(pair ? BN / 2 : BN), 1};   // a pair CTA loads half a weight tile
    encode_map_f16(&q.tmB, w_packed, 3, dims, str, box);
  }
  // ---- epilogue
  q.y = p.y;
  q.bias = p.bias;
  q.residual = reinterpret_cast<const __half*>(p.residual);
  q.y_f32 = p.y_f32;
  q.act = p.act;
  q.res_after_act = p.res_after_act;
  q.nseg = p.nseg;
  for (int i = 0; i < p.nseg && i < 3; ++i) {
    q.seg_begin[i] = p.seg_begin[i];
    q.seg_end[i] = p.seg_end[i];
    q.seg_ps[i] = p.seg_ps[i];
    q.seg_act[i] = p.seg_act[i];
    q.seg_bs[i] = p.seg_bs[i];
    q.seg_y[i] = p.seg_y[i];
  }
  q.y_pix_stride = p.y_pix_stride;
  if (flat) {
    q.y_batch_stride = 0;
    q.res_batch_stride = 0;
  } else {
    q.y_batch_stride = p.y_batch_stride;
    q.res_batch_stride = (long long)p.Ho * p.Wo * p.Cout;   // pixels * Cout; the kernel scales by the plane count
  }
  if (q.epi_tma) {
    // output / residual tile maps: same geometry as the accumulator tile, 64-channel boxes
    uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)q.tw, (uint32_t)q.th, 1};
    uint32_t box32[4] = {32u, (uint32_t)q.tw, (uint32_t)q.th, 1};
    if (flat) {
      uint64_t dims[4] = {CoutP, (uint64_t)Wov, 1, 1};
      uint64_t ystr[3] = {(uint64_t)p.y_pix_stride * 2, (uint64_t)Wov * p.y_pix_stride * 2, (uint64_t)Wov * p.y_pix_stride * 2};
      encode_map_f16(&q.tmY, p.y, 4, dims, ystr, box);
      encode_map_f16(&q.tmY32, p.y, 4, dims, ystr, box32, 64);
      if (p.residual) {
        uint64_t rstr[3] = {CoutP * 2, (uint64_t)Wov * CoutP * 2, (uint64_t)Wov * CoutP * 2};
        encode_map_f16(&q.tmR, p.residual, 4, dims, rstr, box);
      }
    } else {
      uint64_t dims[4] = {CoutP, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.B};
      uint64_t ystr[3] = {(uint64_t)p.y_pix_stride * 2, (uint64_t)p.Wo * p.y_pix_stride * 2, (uint64_t)p.y_batch_stride * 2};
      encode_map_f16(&q.tmY, p.y, 4, dims, ystr, box);
      encode_map_f16(&q.tmY32, p.y, 4, dims, ystr, box32, 64);
      if (p.residual) {
        uint64_t rstr[3] = {CoutP * 2, (uint64_t)p.Wo * CoutP * 2, (uint64_t)p.Ho * p.Wo * CoutP * 2};
        encode_map_f16(&q.tmR, p.residual, 4, dims, rstr, box);
      }
    }
  }
  return plan;
}

void tc_conv_plan_destroy(TcConvPlan* plan) { delete plan; }
void tc_conv_plan_set_pdl(TcConvPlan* plan, int enable) { plan->prm.pdl = (enable && !plan->pair) ? 1 : 0; }
int tc_conv_plan_bn(const TcConvPlan* plan) { return plan->BN; }
int tc_conv_plan_stages(const TcConvPlan* plan) { return plan->prm.stages; }
int tc_conv_plan_grid(const TcConvPlan* plan) { return (int)plan->grid.x; }
int tc_conv_plan_pair(const TcConvPlan* plan) { return plan->pair; }
int tc_conv_plan_epi_groups(const TcConvPlan* plan) { return plan->epi_groups; }
int tc_conv_plan_pdl_friendly(const TcConvPlan* plan) { return plan->pdl_friendly; }
int tc_conv_plan_sk(const TcConvPlan* plan) { return plan->sk; }
size_t tc_conv_sk_workspace_bytes() { return (size_t)148 * BLOCK_M * 256 * sizeof(float) + 148 * 2 * sizeof(int); }
// ws: tc_conv_sk_workspace_bytes() of device memory whose LAST 148 * 2 ints (the flags) are zero; kernels that share a
// workspace must be stream-ordered (each launch leaves the flags zero again)
void tc_conv_plan_set_sk_workspace(TcConvPlan* plan, void* ws) {
  plan->prm.sk = (plan->sk && ws) ? 1 : 0;
  plan->prm.sk_ws = reinterpret_cast<float*>(ws);
  plan->prm.sk_flags = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + (size_t)148 * BLOCK_M * 256 * sizeof(float));
}

template <int BN, bool PAIR, int H, bool SPLIT>
static void launch_bn(const TcConvPlan* plan, cudaStream_t stream) {
  static PerDeviceOnce attr;
  if (attr.first())
    YB_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<BN, PAIR, H, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(224 * 1024)));
  constexpr int THREADS = 64 + 128 * H;
  if (plan->prm.pdl || PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = plan->grid;
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (PAIR) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    } else {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    YB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_conv_kernel<BN, PAIR, H, SPLIT>, plan->prm));
  } else {
    tc_conv_kernel<BN, PAIR, H, SPLIT><<<plan->grid, THREADS, plan->smem_bytes, stream>>>(plan->prm);
  }
}

template <int BN, bool PAIR, bool SPLIT>
static void launch_h(const TcConvPlan* plan, cudaStream_t stream) {
  if (plan->epi_groups == 2)
    launch_bn<BN, PAIR, (BN >= 64 ? 2 : 1), SPLIT>(plan, stream);
  else
    launch_bn<BN, PAIR, 1, SPLIT>(plan, stream);
}

template <bool SPLIT>
static void launch_s(const TcConvPlan* plan, cudaStream_t stream) {
  if (plan->pair) {
    switch (plan->BN) {
      case 256: launch_h<256, true, SPLIT>(plan, stream); break;
      case 128: launch_h<128, true, SPLIT>(plan, stream); break;
      case 64: launch_h<64, true, SPLIT>(plan, stream); break;
      default: YB_REQUIRE(false, "tc_conv: bad BN for a CTA pair");
    }
  } else {
    switch (plan->BN) {
      case 256: launch_h<256, false, SPLIT>(plan, stream); break;
      case 128: launch_h<128, false, SPLIT>(plan, stream); break;
      case 64: launch_h<64, false, SPLIT>(plan, stream); break;
      case 32: launch_h<32, false, SPLIT>(plan, stream); break;
      default: YB_REQUIRE(false, "tc_conv: bad BN");
    }
  }
}

// ---------------------------------------------------------------------------------------------
// chains
// ---------------------------------------------------------------------------------------------
struct TcChain {
  int nl = 0, split = 0, grid = 0, n_done = 0;
  int groups = 2;   // epilogue groups per CTA (2: three split stages, 4: two)
  size_t smem_bytes = 0;
  TcParams* d_layers = nullptr;
  ChainLayer* d_info = nullptr;
  int* d_done = nullptr;
  long long* d_stats = nullptr;   // diagnostics (YB_CHAIN_STATS=1)
};

// a plan the chain kernel can run: its one tile shape (BN = 128, two epilogue groups, staged epilogue, two accumulator
// buffers) and nothing the chain does not implement (pairs, stream-K, staged residual, partial N tiles)
bool tc_conv_plan_chainable(const TcConvPlan* pl) {
  const TcParams& q = pl->prm;
  return (pl->BN == 128 || pl->BN == 64) && !pl->pair && pl->epi_groups == 2 && !pl->sk && !pl->pdl_friendly && q.epi_tma &&
         q.nseg == 0 && q.Cout % pl->BN == 0 && (!q.residual || q.res_direct) && (q.act == ACT_RELU || q.act == ACT_NONE || q.act == ACT_LEAKY);
}

// dep_a[i] / dep_r[i]: index (< i) of the plan that writes plan i's input / residual, -1 for a tensor that is complete
// before the launch (or, for the residual, one whose completion the input dependency already implies)
TcChain* tc_chain_create(const std::vector<const TcConvPlan*>& plans, const std::vector<int>& dep_a, const std::vector<int>& dep_r,
                         int groups) {
  YB_REQUIRE(groups == 2 || groups == 4, "tc_chain: 2 or 4 epilogue groups");
  YB_REQUIRE(!plans.empty() && plans.size() == dep_a.size() && plans.size() == dep_r.size(), "tc_chain: empty chain");
  const TcConvPlan* p0 = plans[0];
  std::vector<TcParams> lp;
  std::vector<ChainLayer> li;
  int ubase = 0, done_off = 0;
  for (size_t i = 0; i < plans.size(); ++i) {
    const TcConvPlan* pl = plans[i];
    YB_REQUIRE(tc_conv_plan_chainable(pl), "tc_chain: plan not chainable");
    YB_REQUIRE(pl->split == p0->split, "tc_chain: the layers of a chain share one precision mode");
    YB_REQUIRE(pl->B == p0->B, "tc_chain: the layers of a chain share one batch size");
    YB_REQUIRE(dep_a[i] < (int)i && dep_r[i] < (int)i, "tc_chain: dependencies point backwards");
    YB_REQUIRE(dep_r[i] < 0 || (plans[dep_r[i]]->Ho == pl->Ho && plans[dep_r[i]]->Wo == pl->Wo), "tc_chain: residual geometry");
    YB_REQUIRE(dep_a[i] < 0 || (plans[dep_a[i]]->Ho == pl->Hi && plans[dep_a[i]]->Wo == pl->Wi), "tc_chain: input geometry");
    ChainLayer c = {};
    c.ubase = ubase;
    c.units = pl->prm.m_tiles * pl->prm.n_tiles;
    c.flat = pl->flat;
    c.W = pl->Wo;
    c.H = pl->Ho;
    c.rows = pl->B * pl->Ho;
    c.tw = pl->prm.tw;
    c.th = pl->prm.th;
    c.tiles_x = pl->prm.tiles_x;
    c.tiles_y = pl->prm.tiles_y;
    c.done_off = done_off;
    c.target = groups * pl->prm.n_tiles;
    c.dep_a = dep_a[i];
    c.dep_r = dep_r[i];
    c.stride = pl->stride;
    c.pad = pl->pad;
    c.kh = pl->KH;
    ubase += c.units;
    done_off += pl->prm.m_tiles;
    lp.push_back(pl->prm);
    li.push_back(c);
  }
  auto* ch = new TcChain();
  ch->nl = (int)plans.size();
  ch->split = p0->split;
  ch->groups = groups;
  ch->n_done = done_off;
  {
    const int npl = p0->split ? 2 : 1;
    ch->smem_bytes = (size_t)chain_stages(p0->split != 0, groups) * npl * (A_STAGE_BYTES + 128 * BLOCK_K * 2) +
                     (size_t)groups * npl * CHUNK32_BYTES + 1024;
  }
  int dev = 0, sms = 148;
  YB_CHECK_CUDA(cudaGetDevice(&dev));
  YB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ch->grid = std::min(sms, ubase);   // one CTA per SM (the shared memory of a CTA sees to that): all co-resident
  try {
    YB_CHECK_CUDA(cudaMalloc(&ch->d_layers, lp.size() * sizeof(TcParams)));
    YB_CHECK_CUDA(cudaMalloc(&ch->d_info, li.size() * sizeof(ChainLayer)));
    YB_CHECK_CUDA(cudaMalloc(&ch->d_done, (size_t)done_off * sizeof(int)));
    YB_CHECK_CUDA(cudaMemcpy(ch->d_layers, lp.data(), lp.size() * sizeof(TcParams), cudaMemcpyHostToDevice));
    YB_CHECK_CUDA(cudaMemcpy(ch->d_info, li.data(), li.size() * sizeof(ChainLayer), cudaMemcpyHostToDevice));
  } catch (...) {
    tc_chain_destroy(ch);
    throw;
  }
  return ch;
}
void tc_chain_destroy(TcChain* ch) {
  if (!ch) return;
  cudaFree(ch->d_layers);
  cudaFree(ch->d_info);
  cudaFree(ch->d_done);
  cudaFree(ch->d_stats);
  delete ch;
}
int tc_chain_layers(const TcChain* ch) { return ch->nl; }
int tc_chain_groups(const TcChain* ch) { return ch->groups; }

template <bool SPLIT, int H>
static void launch_chain_t(const TcChain* ch, cudaStream_t stream) {
  static PerDeviceOnce attr;
  if (attr.first())
    YB_CHECK_CUDA(cudaFuncSetAttribute(tc_chain_kernel<SPLIT, H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)ch->smem_bytes));   // (one size per <SPLIT, H>)
  // Cooperative launch: the grid starts only when ALL its CTAs can be resident at once.  The tile dependencies make CTAs
  // wait for each other, so a partially scheduled grid (two chains from different streams sharing the SMs) could deadlock.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ch->grid);
  cfg.blockDim = dim3(64 + 128 * H);
  cfg.dynamicSmemBytes = ch->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  YB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_chain_kernel<SPLIT, H>, (const TcParams*)ch->d_layers, (const ChainLayer*)ch->d_info,
                                   ch->nl, ch->d_done, ch->d_stats));
}

// diagnostics: one launch with the wait counters on, averaged over the CTAs -> stderr
void tc_chain_print_stats(TcChain* ch, const char* name) {
  const int G = ch->grid;
  YB_CHECK_CUDA(cudaMalloc(&ch->d_stats, (size_t)G * 8 * sizeof(long long)));
  YB_CHECK_CUDA(cudaMemset(ch->d_stats, 0, (size_t)G * 8 * sizeof(long long)));
  launch_tc_chain(ch, 0, nullptr);
  YB_CHECK_CUDA(cudaDeviceSynchronize());
  std::vector<long long> hs((size_t)G * 8);
  YB_CHECK_CUDA(cudaMemcpy(hs.data(), ch->d_stats, hs.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  cudaFree(ch->d_stats);
  ch->d_stats = nullptr;
  double avg[8] = {}, mx[8] = {};
  for (int c = 0; c < G; ++c)
    for (int k = 0; k < 8; ++k) {
      avg[k] += (double)hs[(size_t)c * 8 + k] / G;
      mx[k] = std::max(mx[k], (double)hs[(size_t)c * 8 + k]);
    }
  const double tot = avg[6] > 0 ? avg[6] : 1.0;
  fprintf(stderr,
          "[yolact_b200] chain stats %s: kernel %.0f kcyc/CTA; share of it spent waiting (avg over CTAs / max): producer deps %.1f%% / %.1f%%, "
          "producer ring slot %.1f%%, MMA operands %.1f%% / %.1f%%, MMA accumulator %.1f%%, epilogue accumulator %.1f%% / %.1f%%, epilogue store "
          "completion %.1f%%\n",
          name, tot / 1e3, 100 * avg[0] / tot, 100 * mx[0] / tot, 100 * avg[1] / tot, 100 * avg[2] / tot, 100 * mx[2] / tot, 100 * avg[3] / tot,
          100 * avg[4] / tot, 100 * mx[4] / tot, 100 * avg[5] / tot);
}

// Can a chain launch be captured into a CUDA graph and replayed on this driver?  (One trial per process.)
bool tc_chain_graph_ok(const TcChain* ch) {
  static std::atomic<int> cached{-1};   // (executors of different handles may be built from different threads)
  if (cached.load() >= 0) return cached.load() != 0;
  bool ok = false;
  cudaStream_t s = nullptr;
  cudaGraph_t g = nullptr;
  cudaGraphExec_t ge = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) {
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      try {
        launch_tc_chain(ch, s, nullptr);
      } catch (const Error&) {
      }
      if (cudaStreamEndCapture(s, &g) == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess)
        ok = (cudaGraphLaunch(ge, s) == cudaSuccess) && (cudaStreamSynchronize(s) == cudaSuccess);
    }
  }
  if (ge) cudaGraphExecDestroy(ge);
  if (g) cudaGraphDestroy(g);
  if (s) cudaStreamDestroy(s);
  cudaGetLastError();
  cached.store(ok ? 1 : 0);
  return ok;
}
void launch_tc_chain(const TcChain* ch, cudaStream_t stream, LaunchCounter* lc) {
  YB_CHECK_CUDA(cudaMemsetAsync(ch->d_done, 0, (size_t)ch->n_done * sizeof(int), stream));   // (a memset node in the captured graph)
  if (ch->split) {
    if (ch->groups == 4) launch_chain_t<true, 4>(ch, stream); else launch_chain_t<true, 2>(ch, stream);
  } else {
    if (ch->groups == 4) launch_chain_t<false, 4>(ch, stream); else launch_chain_t<false, 2>(ch, stream);
  }
  YB_CHECK_LAUNCH();
  if (lc) lc->n++;
}

void launch_tc_conv(const TcConvPlan* plan, cudaStream_t stream, LaunchCounter* lc) {
  if (plan->split) launch_s<true>(plan, stream); else launch_s<false>(plan, stream);
  YB_CHECK_LAUNCH();
  if (lc) lc->n++;
}

}  // namespace yb
