// Shared helpers for libvfn_sm100a.so (error plumbing, bf16 split, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vfn.h"

namespace vfn {

void set_error(const char* fmt, ...);

// measurement hooks (vfn_profile_* in include/vfn.h): CUDA events on the launching stream around selected kernels
enum ProfKind { PROF_READ_A = 0, PROF_READ_B = 1, PROF_MATCH = 2, PROF_COMPACT = 3, PROF_MERGE = 4, PROF_APPEND = 5,
                PROF_URR = 6, PROF_KV = 7, PROF_KINDS = 8 };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st, double work);
void count_launches(int n);

// several (d, n) -> (n, d) preparation jobs in one launch (vfn_bank.cu)
struct PrepJob {
  const float* src; int d; int64_t n;
  float* raw; float* normed; uint16_t* hi; uint16_t* lo; float scale; int split_normed;
  int src_em;      // 0: src is (d, n) dimension-major (the reference's layout); 1: src is already (n, d) entry-major
  int64_t n_pad;   // hi / lo: rows [n, n_pad) are zero-filled by the same launch (operand arrays padded to whole tiles)
};
int launch_prep(const PrepJob* jobs, int n_jobs, cudaStream_t st);

// per-object arguments of the update's bookkeeping kernels; the launch_* helpers run one launch for all objects
struct UpdObj {
  vfn_bank bank;
  const float *ck, *cv, *nck, *ncv;            // (hw, d) entry-major candidates: raw / normalised
  const int32_t* match_idx; const float* match_corr;
  int32_t *merge_q, *merge_slot, *run_off, *append_q, *counts, *h_counts;
  void* plan_ws;
  int32_t* n_live;                              // plan: stage n_live[1] = n_live[0] + |A| (NULL: no device count)
  const int32_t* sel; const int32_t* n_sel_dev; int64_t n_sel;   // append: rows to ingest
};
int launch_plan(const UpdObj* o, int n_obj, int64_t hw, float thres_close, cudaStream_t st);
int launch_merge(const UpdObj* o, int n_obj, int64_t hw, float update_rate, cudaStream_t st);
int launch_append(const UpdObj* o, int n_obj, float info0, float info1, cudaStream_t st);
// clamp over rows [0, banks[c].n); banks with n_live also commit their live count: n_live[0] = commit[c] when
// commit && commit[c] >= 0, else the count staged by the plan kernel (n_live[1])
int launch_clamp(const vfn_bank* banks, int n_obj, cudaStream_t st, const int64_t* commit = nullptr);

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

#define VFN_LAUNCH_OK() VFN_CUDA_OK(cudaGetLastError())   /* reads AND clears: a failed launch fails its own call only */

// Programmatic dependent launch: the kernel is queued behind its predecessor on the stream with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs can be placed on the SMs while the predecessor drains;
// every kernel launched this way calls pdl_wait() before it touches memory (griddepcontrol.wait = the predecessor has
// completed and its writes are visible) and pdl_trigger() right after, which lets ITS successor be placed early too.
// vfn_debug_set_pdl(0) falls back to plain launches (the device calls are no-ops then).
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// true the first time it is called for the CURRENT device with this flag array: per-device one-time setup
// (cudaFuncSetAttribute is per device; a process may drive banks on several GPUs)
static inline bool first_use_on_device(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- tensor-core operand formats --------------------------------------------------------------
// fp16 hi/lo split: x ~= hi + lo with hi = f16_rn(x), lo = f16_rn(x - hi): ~22 significant bits.  satfinite: a value
// beyond the fp16 range clamps (finite garbage for |x| > 6.5e4) instead of producing inf/NaN inside an MMA.
__device__ __forceinline__ uint16_t f16_sat(float x) {
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ float f16_to_f32(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ void split_f16(float x, uint16_t& hi, uint16_t& lo) {
  hi = f16_sat(x);
  lo = f16_sat(x - f16_to_f32(hi));
}
// two floats -> packed fp8 pair, a in the LOW byte (memory order a, b)
__device__ __forceinline__ uint16_t e4m3x2(float a, float b) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint16_t e5m2x2(float a, float b) {
  uint16_t r;
  asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(r) : "f"(b), "f"(a));
  return r;
}
// normalised keys are stored x16 in fp16 hi/lo so that the lo part stays in the normal range (|nk| <= 1)
constexpr float NK_SCALE = 16.f;

// keys: fp16 hi + fp16 lo (3-pass product, ~2^-22)
__device__ __forceinline__ void store_key_ops4(uint16_t* kh, uint16_t* kl, int64_t off, float4 v) {
  uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
  split_f16(v.x, h0, l0); split_f16(v.y, h1, l1); split_f16(v.z, h2, l2); split_f16(v.w, h3, l3);
  *reinterpret_cast<uint2*>(kh + off) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
  *reinterpret_cast<uint2*>(kl + off) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
}
// values: fp16 hi, e4m3 of the value (partner of the P-residual pass), e5m2 of the residual v - hi
__device__ __forceinline__ void store_val_ops4(uint16_t* vh, uint8_t* v8, uint8_t* vl, int64_t off, float4 v) {
  const uint16_t h0 = f16_sat(v.x), h1 = f16_sat(v.y), h2 = f16_sat(v.z), h3 = f16_sat(v.w);
  *reinterpret_cast<uint2*>(vh + off) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
  *reinterpret_cast<uint32_t*>(v8 + off) = (uint32_t)e4m3x2(v.x, v.y) | ((uint32_t)e4m3x2(v.z, v.w) << 16);
  *reinterpret_cast<uint32_t*>(vl + off) =
      (uint32_t)e5m2x2(v.x - f16_to_f32(h0), v.y - f16_to_f32(h1)) |
      ((uint32_t)e5m2x2(v.z - f16_to_f32(h2), v.w - f16_to_f32(h3)) << 16);
}
// normalised key: exact fp32 copy (re-score / SIMT match) + fp16 hi/lo of 16*nk (tensor-core match), if present
__device__ __forceinline__ void store_nk4(float* nk, uint16_t* nkh, uint16_t* nkl, int64_t off, float4 v) {
  *reinterpret_cast<float4*>(nk + off) = v;
  if (nkh) store_key_ops4(nkh, nkl, off, make_float4(v.x * NK_SCALE, v.y * NK_SCALE, v.z * NK_SCALE, v.w * NK_SCALE));
}

// live slot count of a bank: device-resident when the bank carries n_live (vfn.h), else the host-tracked n
__device__ __forceinline__ int64_t live_n(const vfn_bank& b) {
  return b.n_live ? (int64_t)*reinterpret_cast<const volatile int32_t*>(b.n_live) : b.n;
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
