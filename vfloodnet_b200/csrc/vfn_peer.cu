// Exchange steps of the sharded feature bank as kernels over PEER memory: every rank's partial results
// live in buffers that all GPUs of the node map (CUDA IPC / symmetric memory, set up by the host), and the combine
// kernels read their peers' partials straight over NVLink - no library collective on the data path.
//   * lse_combine_peers:    (m, l) of every rank's local slots -> global log-sum-exp        (split-memory softmax)
//   * reduce_peers:         sum of the ranks' partial readouts, in rank order (deterministic, bit-identical on all ranks)
//   * match_combine_peers:  arg-max of the cosine match across shards, ties -> lowest global sequence id
// The host brackets each kernel with a cross-GPU barrier (producers done / consumers done).
#include "vfn_common.cuh"

#include <algorithm>

namespace vfn {

constexpr int PEER_MAX = 16;
struct PeerPtrs { const void* p[PEER_MAX]; };

// peer loads must not be served from a stale L1 line of a previous frame: ld.global.relaxed.sys / volatile semantics
__device__ __forceinline__ float2 ld_peer_f2(const float2* p) {
  float2 v;
  asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ longlong2 ld_peer_i2(const longlong2* p) {
  longlong2 v;
  asm volatile("ld.volatile.global.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

__global__ void lse_combine_peers_kernel(const __grid_constant__ PeerPtrs peers, int n_parts, int64_t rows,
                                         float* __restrict__ lse) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows) return;
  float2 v[PEER_MAX];
  float M = -INFINITY;
#pragma unroll
  for (int s = 0; s < PEER_MAX; ++s)
    if (s < n_parts) {
      v[s] = ld_peer_f2(reinterpret_cast<const float2*>(peers.p[s]) + idx);
      M = fmaxf(M, v[s].x);
    }
  float L = 0.f;
  if (M > -INFINITY) {
#pragma unroll
    for (int s = 0; s < PEER_MAX; ++s)
      if (s < n_parts && v[s].x > -INFINITY) L += v[s].y * expf(v[s].x - M);       // rank order: same value on every rank
  }
  lse[idx] = (M > -INFINITY) ? M + logf(L) : -INFINITY;
}

// out[i] = sum_r peers[r][i] for i in [first4, first4 + count4) float4 elements, summed in rank order 0..n-1
__global__ void reduce_peers_kernel(const __grid_constant__ PeerPtrs peers, int n_parts, int64_t first4, int64_t count4,
                                    float4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride) {
    float4 v[PEER_MAX];
#pragma unroll
    for (int s = 0; s < PEER_MAX; ++s)
      if (s < n_parts) v[s] = ld_peer_f4(reinterpret_cast<const float4*>(peers.p[s]) + first4 + i);   // all loads in flight
    float4 a = v[0];
#pragma unroll
    for (int s = 1; s < PEER_MAX; ++s)
      if (s < n_parts) { a.x += v[s].x; a.y += v[s].y; a.z += v[s].z; a.w += v[s].w; }
    out[first4 + i] = a;
  }
}

// out[i] = peers[r][i] for i in the slice rank r reduced (slice4 float4 per rank, the last rank takes the rest): the
// all-gather half of a two-shot reduction.  The local slice is read through peers[self] like the others.
__global__ void gather_peers_kernel(const __grid_constant__ PeerPtrs peers, int n_parts, int64_t slice4, int64_t total4,
                                    float4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int r = (int)min((int64_t)(n_parts - 1), i / slice4);
    out[i] = ld_peer_f4(reinterpret_cast<const float4*>(peers.p[r]) + i);
  }
}

// per query: (corr, seq) pairs of every rank -> best corr, lowest sequence id among the ranks that reach it
// (FeatureBank.py:67: arg-max ties -> lowest index; shards keep insertion order, so lowest index == lowest sequence id)
// pair layout: longlong2 {x = float bits of corr in the low 32 bits, y = sequence id}
__global__ void match_combine_peers_kernel(const __grid_constant__ PeerPtrs peers, int n_parts, int64_t rows,
                                           float* __restrict__ best_corr, int64_t* __restrict__ best_seq) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows) return;
  float best = -INFINITY;
  int64_t seq = INT64_MAX;
  bool nan = false;
  for (int s = 0; s < n_parts; ++s) {
    const longlong2 v = ld_peer_i2(reinterpret_cast<const longlong2*>(peers.p[s]) + idx);
    const float c = __int_as_float((int)(v.x & 0xffffffffll));
    if (c != c) nan = true;                                    // torch.max propagates NaN: such a query goes nowhere
    if (c > best || (c == best && v.y < seq)) { best = c; seq = v.y; }
  }
  best_corr[idx] = nan ? __int_as_float(0x7fc00000) : best;
  best_seq[idx] = nan ? INT64_MAX : seq;
}

// pack the local match result for the exchange: pair[q] = {corr bits, sequence id of the matched local slot}
__global__ void match_pack_kernel(const float* __restrict__ corr, const int32_t* __restrict__ idx,
                                  const int64_t* __restrict__ seq_of_slot, int64_t n_local, int64_t rows,
                                  longlong2* __restrict__ pair) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= rows) return;
  longlong2 v;
  if (n_local > 0) {
    v.x = (long long)(unsigned int)__float_as_int(corr[q]);
    v.y = seq_of_slot[idx[q]];
  } else {
    v.x = (long long)(unsigned int)__float_as_int(-INFINITY);
    v.y = INT64_MAX;
  }
  pair[q] = v;
}

static int fill_peers(PeerPtrs* pp, const void* const* h_peers, int n) {
  if (!h_peers || n < 1 || n > PEER_MAX) { set_error("peer exchange: 1..%d parts, got %d", PEER_MAX, n); return VFN_E_ARG; }
  for (int i = 0; i < PEER_MAX; ++i) pp->p[i] = i < n ? h_peers[i] : nullptr;
  for (int i = 0; i < n; ++i)
    if (!h_peers[i]) { set_error("peer exchange: NULL peer pointer %d", i); return VFN_E_ARG; }
  return VFN_OK;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

int vfn_lse_combine_peers(const void* const* h_peer_ml, int32_t n_parts, int64_t n_rows, float* d_lse, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, h_peer_ml, n_parts)) return rc;
  VFN_CHECK_ARG(d_lse && n_rows >= 1, "lse_combine_peers: bad args");
  lse_combine_peers_kernel<<<(unsigned)cdiv(n_rows, 256), 256, 0, as_stream(stream)>>>(pp, n_parts, n_rows, d_lse);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_reduce_peers(const void* const* h_peer_part, int32_t n_parts, int64_t first, int64_t count, float* d_out,
                     void* stream) {
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, h_peer_part, n_parts)) return rc;
  VFN_CHECK_ARG(d_out && first >= 0 && count >= 0 && first % 4 == 0 && count % 4 == 0,
                "reduce_peers: first/count must be multiples of 4 floats");
  if (count == 0) return VFN_OK;
  int sms = 0, dev = 0;
  VFN_CUDA_OK(cudaGetDevice(&dev));
  VFN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t c4 = count / 4;
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(c4, 256), (int64_t)sms * 8);
  reduce_peers_kernel<<<grid, 256, 0, as_stream(stream)>>>(pp, n_parts, first / 4, c4, reinterpret_cast<float4*>(d_out));
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_gather_peers(const void* const* h_peer_buf, int32_t n_parts, int32_t self, int64_t slice, int64_t total,
                     float* d_out, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, h_peer_buf, n_parts)) return rc;
  VFN_CHECK_ARG(d_out && slice > 0 && total > 0 && slice % 4 == 0 && total % 4 == 0 && self >= 0 && self < n_parts,
                "gather_peers: bad args");
  int sms = 0, dev = 0;
  VFN_CUDA_OK(cudaGetDevice(&dev));
  VFN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t t4 = total / 4;
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(t4, 256), (int64_t)sms * 8);
  gather_peers_kernel<<<grid, 256, 0, as_stream(stream)>>>(pp, n_parts, slice / 4, t4, reinterpret_cast<float4*>(d_out));
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_match_pack(const float* d_corr, const int32_t* d_idx, const int64_t* d_seq_of_slot, int64_t n_local, int64_t hw,
                   void* d_pair, void* stream) {
  VFN_CHECK_ARG(d_pair && hw >= 1 && (n_local == 0 || (d_corr && d_idx && d_seq_of_slot)), "match_pack: bad args");
  match_pack_kernel<<<(unsigned)cdiv(hw, 256), 256, 0, as_stream(stream)>>>(d_corr, d_idx, d_seq_of_slot, n_local, hw,
                                                                          reinterpret_cast<longlong2*>(d_pair));
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_match_combine_peers(const void* const* h_peer_pair, int32_t n_parts, int64_t hw, float* d_best_corr,
                            int64_t* d_best_seq, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, h_peer_pair, n_parts)) return rc;
  VFN_CHECK_ARG(d_best_corr && d_best_seq && hw >= 1, "match_combine_peers: bad args");
  match_combine_peers_kernel<<<(unsigned)cdiv(hw, 256), 256, 0, as_stream(stream)>>>(pp, n_parts, hw, d_best_corr, d_best_seq);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

}  // extern "C"
