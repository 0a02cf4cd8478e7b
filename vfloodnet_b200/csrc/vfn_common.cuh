// Shared helpers for libvfn_sm100a.so (error plumbing, bf16 split, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vfn.h"

namespace vfn {

void set_error(const char* fmt, ...);

// measurement hooks (vfn_profile_* in include/vfn.h): CUDA events on the launching stream around selected kernels
enum ProfKind { PROF_READ_A = 0, PROF_READ_B = 1, PROF_MATCH = 2, PROF_COMPACT = 3, PROF_MERGE = 4, PROF_APPEND = 5,
                PROF_URR = 6, PROF_KINDS = 8 };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st, double work);
void count_launches(int n);

#define VFN_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      ::vfn::set_error(__VA_ARGS__);          \
      return VFN_E_ARG;                       \
    }                                         \
  } while (0)

#define VFN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::vfn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VFN_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

#define VFN_LAUNCH_OK() VFN_CUDA_OK(cudaPeekAtLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bf16 hi/lo split: x ~= hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi)   (16 mantissa bits kept)
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(h);
  __nv_bfloat16 l = __float2bfloat16_rn(r);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}

// tf32 hi/lo split: hi keeps sign, exponent and the top 10 mantissa bits (exactly representable in tf32, whatever
// the tensor core does with the discarded bits), lo = x - hi is exact in fp32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void store_nk4(float* nkh, float* nkl, int64_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(nkh + off) = h;
  *reinterpret_cast<float4*>(nkl + off) = l;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; red must hold >= 32 floats; all threads get the result. blockDim.x multiple of 32.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

}  // namespace vfn
