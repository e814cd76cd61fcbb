// Shared helpers for the yolact_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

#include "../../include/yolact_b200.h"

namespace yb {

// ---- error plumbing -------------------------------------------------------------------------
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define YB_CHECK_CUDA(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      throw ::yb::Error(YB_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + \
                                         " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                          \
  } while (0)

#define YB_REQUIRE(cond, msg)                                                                  \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      throw ::yb::Error(YB_ERR_INVALID, std::string(msg) + " [" #cond "] (" + __FILE__ + ":" + \
                                            std::to_string(__LINE__) + ")");                   \
    }                                                                                          \
  } while (0)

// Checks the launch itself (not completion); cheap and capture-safe.
#define YB_CHECK_LAUNCH() YB_CHECK_CUDA(cudaGetLastError())

// ---- enums shared by kernels ----------------------------------------------------------------
enum ActKind : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_LEAKY = 3 };

template <typename T>
struct DType;
template <>
struct DType<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
};
template <>
struct DType<__half> {
  static constexpr int kVec = 8;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) {
  // saturate instead of producing inf: fp16 max is 65504
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __float2half_rn(v);
}

// Split-precision operands (YB_PREC_F16X3): a value is stored as a pair hi + lo.
//   * activations: hi = rn_fp16(v), lo = rn_BF16(v - hi).  The lo plane is bfloat16 on purpose: lo is ~2^-12 |v|, which
//     for |v| < 0.25 is an fp16 SUBNORMAL, and the tensor core flushes fp16 subnormal inputs to zero (measured: head
//     tensors 5e-5 of range off with fp16 lo planes, the arithmetic's own bound is 1e-7).  bfloat16 has the exponent
//     range of fp32; its 8 significand bits on top of hi's 11 give >= 19 bits.  tcgen05.mma kind::f16 takes the A and
//     B formats independently, so the A_lo * W_hi pass runs as bf16 x fp16.
//   * weights: w * 2^e = hi + lo, both fp16 (e puts the layer's largest weight just below 2^14, so lo is a normal
//     fp16 number for every weight above 2^-17 of the maximum), see engine.cu split_exponent.
// A pixel of a split NHWC tensor is [hi(C) | lo(C)], i.e. 2*C 16-bit elements; buffers are typed __half.
__device__ __forceinline__ __half lo_from_f32(float r) {     // fp32 residual -> bf16 bits carried in a __half slot
  const __nv_bfloat16 b = __float2bfloat16_rn(r);
  return *reinterpret_cast<const __half*>(&b);
}
__device__ __forceinline__ float lo_to_f32(__half h) {
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&h));
}
__device__ __forceinline__ __half2 lo2_from_f32(float r0, float r1) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(r0, r1);
  return *reinterpret_cast<const __half2*>(&b);
}
__device__ __forceinline__ float2 lo2_to_f32(__half2 h) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h));
}
// activation split (hi saturates at +-65504; the residual of a saturated value is dropped)
__device__ __forceinline__ void split_f32(float v, __half& hi, __half& lo) {
  const float c = fminf(fmaxf(v, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = lo_from_f32(c - __half2float(hi));
}
// weight split: both halves fp16 (the caller pre-scales by a power of two)
__device__ __forceinline__ void split_w_f32(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_TANH: return tanhf(v);
    case ACT_LEAKY: return v > 0.f ? v : 0.1f * v;
    default: return v;
  }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute applies to the CURRENT device: a function-local static of this type remembers, per device,
// whether the attribute of that kernel instantiation has been raised (one process may drive several GPUs).
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Launch counter (per handle); every launcher takes one of these.
struct LaunchCounter {
  int64_t n = 0;
};

}  // namespace yb
