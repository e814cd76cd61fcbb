// Fused DCNv2 (modulated deformable 3x3 convolution, deformable_groups = 1) on tcgen05: deformable gather ->
// shared-memory A stage -> tensor-core contraction -> bias + ReLU, ONE kernel, no column buffer.
//
// Reference: DCN.forward (external/DCNv2/dcn_v2.py:118-128), modulated_deformable_im2col_gpu_kernel
// (external/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:125-195, bilinear sampler :25-54) followed by the batched SGEMM + bias
// of dcn_v2_cuda_forward (src/cuda/dcn_v2_cuda.cu:123-163).  The reference materialises the sampled columns
// [B, 9*C, Ho*Wo] fp32 in HBM between the two; round 1 of this repo wrote them as fp16 (88 MB per layer at 69x69) and ran
// a 1x1 tensor-core conv over them.  Here they never leave the SM:
//
//   GEMM view   D[128 output pixels, BN couts] = sum over taps t (9) and 64-channel chunks kc of
//               A_t,kc[128, 64] * W[BN, t*C + kc*64 .. +64]^T,
//               A_t,kc[m, c] = mask(m,t) * bilinear(x, pos(m,t))[kc*64 + c]          (fp32 math, reference op order)
//   warp 0      TMA producer of the weight tiles (B operand, [Cout][9*C] K-major)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit frees the stage
//   warps 2..9  gather: warp w owns 16 tile rows.  Once per tap, lanes 0..15 each compute the sampling geometry of ONE
//               row (4 corner offsets, 4 bilinear weights, the modulation mask) from the 27 offset/mask channels; for
//               every 64-channel chunk the geometry reaches the lanes that need it by WARP SHUFFLE (8 lanes share a
//               row: one 16-byte vector of 8 channels each), the four corners are fetched with 16-byte loads, blended
//               (fp32 in the split mode, packed half2 in the fp16 mode) and written as one swizzled 16-byte piece of the K-major SWIZZLE_128B A tile -- the layout
//               tcgen05.mma consumes directly.  Geometry is computed once per (pixel, tap) instead of once per
//               (pixel, tap, 8 channels) as in the per-thread gather.
//   epilogue    the same 8 warps: TMEM -> registers -> (* out_scale) + bias -> ReLU -> one full 128-byte line per
//               thread and 64-channel chunk (even chunks: warps 2..5, odd chunks: warps 6..9).
//   SPLIT       (YB_PREC_F16X3) x and y are [hi(C) | lo(C)] pairs, the sample is hi + lo, the A stage holds a hi (fp16)
//               and a lo (2^11-scaled fp16) tile, the weights [Cout][hi(9C) | lo(9C)], three MMA passes per k-block.
#include "tc_common.cuh"

namespace yb {

using namespace tc;

namespace {

constexpr int DM = 128;                 // tile rows (output pixels)
constexpr int DK = 64;                  // channels per k-block (one 128-byte swizzle row)
constexpr int A_TILE = DM * DK * 2;     // 16 KB
constexpr int GATHER_WARPS = 8;
constexpr int DTHREADS = 64 + 32 * GATHER_WARPS;
constexpr int MAX_DSTAGES = 6;

struct alignas(64) DcnParams {
  CUtensorMap tmW;
  const __half* x;
  const float* om;      // [M, 27]: 18 offsets (dh, dw interleaved per tap), 9 mask logits / masks
  const float* bias;
  __half* y;
  int B, H, W, C, Ho, Wo, Cout;
  long long M;
  int stride, pad, dil;
  int act, mask_logits;
  int stages, kchunks;
  uint32_t idesc;
  float out_scale;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(DTHREADS)
dcn_tc_kernel(const __grid_constant__ DcnParams p) {
  constexpr int NPL = SPLIT ? 2 : 1;
  constexpr int B_PLANE = BN * DK * 2;
  constexpr int STAGE = NPL * (A_TILE + B_PLANE);

  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t a_full[MAX_DSTAGES], b_full[MAX_DSTAGES], empty_bar[MAX_DSTAGES], tmem_full;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int num_kb = 9 * p.kchunks;
  const long long m0 = (long long)blockIdx.x * DM;
  const int n0 = blockIdx.y * BN;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&a_full[s], GATHER_WARPS);
      mbar_init(&b_full[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmW);
  }
  if (warp == 1) tmem_alloc<NPL * BN>(&s_tmem);   // split: second accumulator for the lo cross terms (common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ===================== weight tiles (TMA) =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages, it = kb / stages;
        mbar_wait(&empty_bar[s], (uint32_t)(it & 1) ^ 1u);
        const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
        uint8_t* sb = smem + (size_t)s * STAGE + NPL * A_TILE;
        mbar_expect_tx(&b_full[s], (uint32_t)(NPL * B_PLANE));
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl)
          tma_load_3d(sb + pl * B_PLANE, &p.tmW, &b_full[s], tap * p.C + kc * DK + pl * 9 * p.C, n0, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages, it = kb / stages;
        mbar_wait(&b_full[s], (uint32_t)(it & 1));
        mbar_wait(&a_full[s], (uint32_t)(it & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE);
        const uint32_t sb = sa + NPL * A_TILE;
        const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sb);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
          umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        if (SPLIT) {
          const uint64_t dal = make_sw128_desc(sa + A_TILE), dbl = make_sw128_desc(sb + B_PLANE);
#pragma unroll
          for (int k = 0; k < DK / 16; ++k)   // A_lo * W_hi -> second accumulator
            umma_f16(tmem_base + BN, dal + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < DK / 16; ++k)   // A_hi * W_lo -> second accumulator
            umma_f16(tmem_base + BN, da + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), p.idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full);
    }
  } else {
    // ===================== gather warps =====================
    const int gw = warp - 2;                        // 0..7: rows [16*gw, 16*gw + 16)
    const int PS = NPL * p.C;                       // halfs per input pixel
    // geometry owner: lanes 0..15 <-> row 16*gw + lane
    const long long gm = m0 + gw * 16 + (lane & 15);
    const bool gvalid = (lane < 16) && (gm < p.M);
    int gb = 0, gho = 0, gwo = 0;
    if (gvalid) {
      gwo = (int)(gm % p.Wo);
      const long long t = gm / p.Wo;
      gho = (int)(t % p.Ho);
      gb = (int)(t / p.Ho);
    }
    const float* gom = p.om + gm * 27;
    const int img_base = gb * p.H * p.W;
    // this lane's items: 4 per k-block; item i -> row 16*gw + 4*i + (lane >> 3), 16-byte piece (lane & 7)
    const int piece = lane & 7;
    const int src_sub = lane >> 3;                  // + 4*i = geometry owner lane

    // Corner offsets are ALWAYS valid addresses (an out-of-range corner points at this image's pixel 0 and carries
    // weight 0, which is what the reference's `if (h_low >= 0 && ...) v = ...` amounts to): the corner loads of a
    // k-block are then unconditional and the compiler issues them back to back -- with a branch per corner every load
    // waited for the previous one (measured: 3.4 us per k-block, 124 us for a 35x35 layer).
    // The modulation mask is folded into the four bilinear weights (one rounding of difference to the reference's
    // (sum) * mask, far below the fp16 / split resolution).
    int ao[4][4];        // [item][corner] pixel offsets of this lane's 4 rows, refreshed once per tap by shuffle
    float aw[4][4];      // [item][corner] weight * mask (split mode)
    __half2 hw[4][4];    // the same as broadcast half2 pairs (fp16 mode)
    int cur_tap = -1;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages, it = kb / stages;
      const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
      if (tap != cur_tap) {
        cur_tap = tap;
        int o00 = img_base, o01 = img_base, o10 = img_base, o11 = img_base;
        float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
        if (gvalid) {
          // dcn_v2_im2col_cuda.cu:151-189: h_im = h_in + i*dil + offset_h, w_im likewise; inside test (-1, H) x (-1, W)
          const int i = tap / 3, j = tap - i * 3;
          const float hh = __fadd_rn((float)(gho * p.stride - p.pad + i * p.dil), __ldg(gom + 2 * tap));
          const float ww = __fadd_rn((float)(gwo * p.stride - p.pad + j * p.dil), __ldg(gom + 2 * tap + 1));
          const float mv = __ldg(gom + 18 + tap);
          const float msk = p.mask_logits ? __fdiv_rn(1.f, __fadd_rn(1.f, expf(-mv))) : mv;
          if (hh > -1.f && ww > -1.f && hh < (float)p.H && ww < (float)p.W) {
            const int hl = (int)floorf(hh), wl = (int)floorf(ww);
            const int hhi = hl + 1, whi = wl + 1;
            const float lh = __fsub_rn(hh, (float)hl), lw = __fsub_rn(ww, (float)wl);
            const float uh = __fsub_rn(1.f, lh), uw = __fsub_rn(1.f, lw);
            if (hl >= 0 && wl >= 0) {
              o00 = img_base + hl * p.W + wl;
              w00 = __fmul_rn(__fmul_rn(uh, uw), msk);
            }
            if (hl >= 0 && whi <= p.W - 1) {
              o01 = img_base + hl * p.W + whi;
              w01 = __fmul_rn(__fmul_rn(uh, lw), msk);
            }
            if (hhi <= p.H - 1 && wl >= 0) {
              o10 = img_base + hhi * p.W + wl;
              w10 = __fmul_rn(__fmul_rn(lh, uw), msk);
            }
            if (hhi <= p.H - 1 && whi <= p.W - 1) {
              o11 = img_base + hhi * p.W + whi;
              w11 = __fmul_rn(__fmul_rn(lh, lw), msk);
            }
          }
        }
        // geometry owner lane -> the 8 lanes that gather that row (item i: row 16*gw + 4*i + (lane >> 3))
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int src = 4 * i + src_sub;
          ao[i][0] = __shfl_sync(0xffffffffu, o00, src);
          ao[i][1] = __shfl_sync(0xffffffffu, o01, src);
          ao[i][2] = __shfl_sync(0xffffffffu, o10, src);
          ao[i][3] = __shfl_sync(0xffffffffu, o11, src);
          aw[i][0] = __shfl_sync(0xffffffffu, w00, src);
          aw[i][1] = __shfl_sync(0xffffffffu, w01, src);
          aw[i][2] = __shfl_sync(0xffffffffu, w10, src);
          aw[i][3] = __shfl_sync(0xffffffffu, w11, src);
          if (!SPLIT) {
#pragma unroll
            for (int c = 0; c < 4; ++c) hw[i][c] = __float2half2_rn(aw[i][c]);
          }
        }
      }
      mbar_wait(&empty_bar[s], (uint32_t)(it & 1) ^ 1u);   // the MMAs that read this stage have completed
      uint8_t* sa = smem + (size_t)s * STAGE;
      const __half* xc = p.x + kc * DK + piece * 8;
      if (!SPLIT) {
        // fp16 mode: the samples are rounded to fp16 anyway -- blend in packed half2 arithmetic (4 HFMA2 per corner)
        uint4 raw[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) raw[i][c] = ldg_nc16(xc + (size_t)ao[i][c] * PS);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 oh;
          __half2* o2 = reinterpret_cast<__half2*>(&oh);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const __half2 w2 = hw[i][c];
            const __half2* v2 = reinterpret_cast<const __half2*>(&raw[i][c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = (c == 0) ? __hmul2(w2, v2[j]) : __hfma2(w2, v2[j], o2[j]);
          }
          const int row = gw * 16 + 4 * i + src_sub;
          const uint32_t off = (uint32_t)row * 128u + (((uint32_t)piece ^ ((uint32_t)row & 7u)) << 4);
          *reinterpret_cast<uint4*>(sa + off) = oh;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 rh[4], rl[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const __half* px = xc + (size_t)ao[i][c] * PS;
            rh[c] = ldg_nc16(px);
            rl[c] = ldg_nc16(px + p.C);
          }
          uint4 oh, ol;
          __half2* oh2 = reinterpret_cast<__half2*>(&oh);
          __half2* ol2 = reinterpret_cast<__half2*>(&ol);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float r0 = 0.f, r1 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float2 fh = __half22float2(reinterpret_cast<const __half2*>(&rh[c])[j]);
              const float2 fl = lo2_to_f32(reinterpret_cast<const __half2*>(&rl[c])[j]);
              r0 = __fmaf_rn(aw[i][c], fh.x + fl.x, r0);
              r1 = __fmaf_rn(aw[i][c], fh.y + fl.y, r1);
            }
            oh2[j] = __floats2half2_rn(r0, r1);   // |sample| <= max |x|: no saturation needed
            const float2 hf = __half22float2(oh2[j]);
            ol2[j] = lo2_from_f32(r0 - hf.x, r1 - hf.y);
          }
          const int row = gw * 16 + 4 * i + src_sub;
          const uint32_t off = (uint32_t)row * 128u + (((uint32_t)piece ^ ((uint32_t)row & 7u)) << 4);
          *reinterpret_cast<uint4*>(sa + off) = oh;
          *reinterpret_cast<uint4*>(sa + A_TILE + off) = ol;
        }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
    }

    // ===================== epilogue =====================
    mbar_wait(&tmem_full, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int hgrp = (warp - 2) >> 2;               // warps 2..5: even 64-channel chunks, warps 6..9: odd chunks
    const int row = quad * 32 + lane;
    const long long m = m0 + row;
    const int nchunks = (min(BN, p.Cout - n0) + 63) >> 6;
    __half* yrow = p.y + m * (long long)(NPL * p.Cout) + n0;
    for (int c = hgrp; c < nchunks; c += 2) {
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 64), r0);
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 64 + 32), r1);
      if (SPLIT) {
        uint32_t q2[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + c * 64), q2);
#pragma unroll
        for (int j = 0; j < 32; ++j) r0[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q2[j]), YB_LO_INV, __uint_as_float(r0[j])));
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + c * 64 + 32), q2);
#pragma unroll
        for (int j = 0; j < 32; ++j) r1[j] = __float_as_uint(__fmaf_rn(__uint_as_float(q2[j]), YB_LO_INV, __uint_as_float(r1[j])));
      }
      if (m >= p.M) continue;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 o, ol;
        __half2* o2 = reinterpret_cast<__half2*>(&o);
        __half2* ol2 = reinterpret_cast<__half2*>(&ol);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = q * 8 + 2 * j;
          const int ch = n0 + c * 64 + col;
          float v0 = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]);
          float v1 = __uint_as_float(col + 1 < 32 ? r0[col + 1] : r1[col + 1 - 32]);
          if (SPLIT) {
            v0 *= p.out_scale;
            v1 *= p.out_scale;
          }
          v0 = apply_act(v0 + ((p.bias && ch < p.Cout) ? __ldg(p.bias + ch) : 0.f), p.act);
          v1 = apply_act(v1 + ((p.bias && ch + 1 < p.Cout) ? __ldg(p.bias + ch + 1) : 0.f), p.act);
          const float c0 = fminf(fmaxf(v0, -65504.f), 65504.f), c1 = fminf(fmaxf(v1, -65504.f), 65504.f);
          o2[j] = __floats2half2_rn(c0, c1);
          if (SPLIT) {
            const float2 hf = __half22float2(o2[j]);
            ol2[j] = lo2_from_f32(c0 - hf.x, c1 - hf.y);
          }
        }
        const int cbase = c * 64 + q * 8;
        if (n0 + cbase + 8 <= p.Cout) {   // Cout % 8 == 0 is required by the launcher
          *reinterpret_cast<uint4*>(yrow + cbase) = o;
          if (SPLIT) *reinterpret_cast<uint4*>(yrow + p.Cout + cbase) = ol;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<NPL * BN>(tmem_base);
  }
}

template <int BN, bool SPLIT>
void launch_dcn_variant(const DcnParams& prm, size_t smem, dim3 grid, cudaStream_t stream) {
  static PerDeviceOnce attr;
  if (attr.first())
    YB_CHECK_CUDA(cudaFuncSetAttribute(dcn_tc_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024)));
  dcn_tc_kernel<BN, SPLIT><<<grid, DTHREADS, smem, stream>>>(prm);
}

}  // namespace

struct DcnTcPlan {
  DcnParams prm;
  int BN = 256;
  int split = 0;
  size_t smem = 0;
  dim3 grid;
};

bool dcn_tc_supported(int C, int Cout) { return C % 64 == 0 && Cout % 8 == 0 && Cout >= 8; }

// w_packed: [Cout][9*C] fp16, k = tap*C + c  (split: [Cout][hi(9C) | lo(9C)] of w / out_scale)
DcnTcPlan* dcn_tc_plan_create(const __half* x, const float* om, const __half* w_packed, const float* bias, __half* y, int B,
                              int H, int W, int C, int Ho, int Wo, int Cout, int stride, int pad, int dil, int act,
                              int mask_logits, int split, float out_scale, int bn_override) {
  YB_REQUIRE(dcn_tc_supported(C, Cout), "dcn_tc: needs C % 64 == 0 and Cout % 8 == 0");
  auto* plan = new DcnTcPlan();
  DcnParams& q = plan->prm;
  memset(&q, 0, sizeof(q));
  plan->split = split ? 1 : 0;
  const int npl = split ? 2 : 1;
  int BN = Cout >= 256 ? 256 : (Cout > 64 ? 128 : 64);
  if (bn_override == 64 || bn_override == 128 || bn_override == 256) BN = bn_override;
  plan->BN = BN;
  q.x = x;
  q.om = om;
  q.bias = bias;
  q.y = y;
  q.B = B;
  q.H = H;
  q.W = W;
  q.C = C;
  q.Ho = Ho;
  q.Wo = Wo;
  q.Cout = Cout;
  q.M = (long long)B * Ho * Wo;
  q.stride = stride;
  q.pad = pad;
  q.dil = dil;
  q.act = act;
  q.mask_logits = mask_logits;
  q.kchunks = C / DK;
  q.out_scale = split ? out_scale : 1.f;
  q.idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(DM >> 4) << 24);
  const int stage = npl * (A_TILE + BN * DK * 2);
  int stages = std::min(MAX_DSTAGES, (221 * 1024) / stage);
  stages = std::max(1, std::min(stages, 9 * q.kchunks));
  q.stages = stages;
  plan->smem = (size_t)stages * stage + 1024;
  plan->grid = dim3((unsigned)((q.M + DM - 1) / DM), (unsigned)ceil_div(Cout, BN), 1);
  const uint64_t KP = (uint64_t)npl * 9 * C;
  uint64_t dims[3] = {KP, (uint64_t)Cout, 1};
  uint64_t str[2] = {KP * 2, KP * (uint64_t)Cout * 2};
  uint32_t box[3] = {(uint32_t)DK, (uint32_t)BN, 1};
  encode_map_f16(&q.tmW, w_packed, 3, dims, str, box);
  return plan;
}

void dcn_tc_plan_destroy(DcnTcPlan* plan) { delete plan; }

void launch_dcn_tc(const DcnTcPlan* plan, cudaStream_t stream, LaunchCounter* lc) {
  const DcnParams& q = plan->prm;
  if (plan->split) {
    switch (plan->BN) {
      case 256: launch_dcn_variant<256, true>(q, plan->smem, plan->grid, stream); break;
      case 128: launch_dcn_variant<128, true>(q, plan->smem, plan->grid, stream); break;
      default: launch_dcn_variant<64, true>(q, plan->smem, plan->grid, stream); break;
    }
  } else {
    switch (plan->BN) {
      case 256: launch_dcn_variant<256, false>(q, plan->smem, plan->grid, stream); break;
      case 128: launch_dcn_variant<128, false>(q, plan->smem, plan->grid, stream); break;
      default: launch_dcn_variant<64, false>(q, plan->smem, plan->grid, stream); break;
    }
  }
  YB_CHECK_LAUNCH();
  if (lc) lc->n++;
}

}  // namespace yb
