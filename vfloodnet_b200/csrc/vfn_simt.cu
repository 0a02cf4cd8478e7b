// fp32 SIMT implementation of the dense parts of the hot path (affinity scores -> LSE / arg-max, and the
// softmax readout).  Exact fp32 arithmetic, any d_key/d_val multiple of 8.  It is (a) the decision-exact path
// for the cosine match, (b) the small-shape / non-128x512 path of the read, (c) the on-GPU cross-check of the
// tcgen05 kernels.  Also hosts the pieces shared with the tcgen05 path: partial combine, LSE combine, usage-count
// finalize, and the vfn_memread / vfn_bank_match dispatch.
// Reference behaviour restated: AFB_URR.py:136-178 (read), FeatureBank.py:63-68 (match).
#include "vfn_common.cuh"
#include "vfn_tc.cuh"

namespace vfn {

constexpr int TM = 128;   // bank slots per tile
constexpr int TN = 128;   // queries per tile
constexpr int TK = 16;    // reduction chunk
constexpr int LDS_ = 132; // padded smem leading dimension (multiple of 4)
constexpr int ST_THREADS = 256;
constexpr int MAX_OBJ = 8;

struct BankSet {
  vfn_bank b[MAX_OBJ];
};

enum ScoreMode { MODE_LSE = 0, MODE_MATCH = 1 };

// Loads rows [r0, r0+128) x cols [k0, k0+16) of a row-major (n_rows, d) matrix into registers (2 float4 / thread).
// A2 != nullptr: the matrix is stored as an exact two-term split (A + A2), e.g. the bank's nkh + nkl.
__device__ __forceinline__ void load_chunk(const float* __restrict__ A, const float* __restrict__ A2, int64_t n_rows,
                                           int d, int64_t r0, int k0, float4 (&reg)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int idx = threadIdx.x + ST_THREADS * r;
    const int row = idx >> 2, c4 = idx & 3;
    const int64_t gr = r0 + row;
    const int k = k0 + c4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < n_rows && k < d) {
      v = *reinterpret_cast<const float4*>(A + gr * d + k);
      if (A2) {
        const float4 w = *reinterpret_cast<const float4*>(A2 + gr * d + k);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;     // exact: lo was defined as value - hi
      }
    }
    reg[r] = v;
  }
}
__device__ __forceinline__ void store_chunk_t(float (*S)[LDS_], const float4 (&reg)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int idx = threadIdx.x + ST_THREADS * r;
    const int row = idx >> 2, c4 = idx & 3;
    S[c4 * 4 + 0][row] = reg[r].x;
    S[c4 * 4 + 1][row] = reg[r].y;
    S[c4 * 4 + 2][row] = reg[r].z;
    S[c4 * 4 + 3][row] = reg[r].w;
  }
}

// acc[a][b] = sum_k A[i_a][k] * B[j_b][k] for the thread's 8 rows (i) and 8 columns (j); sequential fp32 FMA over k.
__device__ __forceinline__ void tile_scores(const float* __restrict__ A, const float* __restrict__ A2, int64_t n_a,
                                            const float* __restrict__ B,
                                            int64_t n_b, int d, int64_t i0, int64_t j0, float (*As)[LDS_],
                                            float (*Bs)[LDS_], float (&acc)[8][8]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  float4 ra[2], rb[2];
  load_chunk(A, A2, n_a, d, i0, 0, ra);
  load_chunk(B, nullptr, n_b, d, j0, 0, rb);
  for (int k0 = 0; k0 < d; k0 += TK) {
    __syncthreads();
    store_chunk_t(As, ra);
    store_chunk_t(Bs, rb);
    __syncthreads();
    if (k0 + TK < d) {
      load_chunk(A, A2, n_a, d, i0, k0 + TK, ra);
      load_chunk(B, nullptr, n_b, d, j0, k0 + TK, rb);
    }
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
  }
}

__device__ __forceinline__ int row_of(int ty, int a) { return (a < 4) ? ty * 4 + a : 64 + ty * 4 + (a - 4); }
__device__ __forceinline__ int col_of(int tx, int b) { return (b < 4) ? tx * 4 + b : 64 + tx * 4 + (b - 4); }

// ------------------------------------------------------------------------------------------------
// score kernel: grid (q_tiles, n_split, obj_n).  MODE_LSE: per-query (max, sum exp) of s = <K_i,q_j>/sqrt(d)
// over the split's slots; MODE_MATCH: per-query (max, argmax) of <A_i, B_j>, ties -> lowest slot.
// partial layout: part[(obj * n_split + split) * hw + j] = {v0, v1}
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) simt_score_kernel(BankSet banks, const float* __restrict__ Q, int64_t hw,
                                                                int n_split, float2* __restrict__ part) {
  __shared__ __align__(16) float As[TK][LDS_];
  __shared__ __align__(16) float Bs[TK][LDS_];
  __shared__ float red0[16][TN];
  __shared__ float red1[16][TN];
  const vfn_bank bk = banks.b[blockIdx.z];
  const int d = bk.d_key;
  const float* A = (MODE == MODE_MATCH) ? bk.nk : bk.keys;
  const float* A2 = nullptr;
  const int64_t n = bk.n;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t j0 = (int64_t)blockIdx.x * TN;
  const int64_t tiles = (n + TM - 1) / TM;
  const int64_t t_begin = tiles * blockIdx.y / n_split, t_end = tiles * (blockIdx.y + 1) / n_split;
  const float sqrt_d = sqrtf((float)d);
  float run0 = -INFINITY;   // running max
  float run1 = 0.f;         // running sum-exp (LSE) or arg index as int bits (MATCH)
  int run_idx = 0x7fffffff;
  float acc[8][8];
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t i0 = t * TM;
    tile_scores(A, A2, n, Q, hw, d, i0, j0, As, Bs, acc);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float m = -INFINITY;
      int mi = 0x7fffffff;
      float l = 0.f;
      float v[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int64_t gi = i0 + row_of(ty, a);
        float s = (MODE == MODE_LSE) ? __fdiv_rn(acc[a][b], sqrt_d) : acc[a][b];
        if (gi >= n) s = -INFINITY;
        v[a] = s;
        if (s > m) { m = s; mi = (int)gi; }
      }
      if (MODE == MODE_LSE) {
        if (m > -INFINITY) {
#pragma unroll
          for (int a = 0; a < 8; ++a) l += expf(v[a] - m);
        }
        red1[ty][col_of(tx, b)] = l;
      } else {
        red1[ty][col_of(tx, b)] = __int_as_float(mi);
      }
      red0[ty][col_of(tx, b)] = m;
    }
    __syncthreads();
    if (threadIdx.x < TN) {
      const int j = threadIdx.x;
      if (MODE == MODE_LSE) {
        float m = run0;
#pragma unroll
        for (int y = 0; y < 16; ++y) m = fmaxf(m, red0[y][j]);
        if (m > -INFINITY) {
          float l = run1 * expf(run0 - m);
#pragma unroll
          for (int y = 0; y < 16; ++y) {
            const float pm = red0[y][j];
            if (pm > -INFINITY) l += red1[y][j] * expf(pm - m);
          }
          run0 = m;
          run1 = l;
        }
      } else {
#pragma unroll
        for (int y = 0; y < 16; ++y) {
          const float pm = red0[y][j];
          const int pi = __float_as_int(red1[y][j]);
          if (pm > run0 || (pm == run0 && pi < run_idx)) { run0 = pm; run_idx = pi; }
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < TN) {
    const int64_t j = j0 + threadIdx.x;
    if (j < hw) {
      float2 o;
      o.x = run0;
      o.y = (MODE == MODE_LSE) ? run1 : __int_as_float(run_idx);
      part[((int64_t)blockIdx.z * n_split + blockIdx.y) * hw + j] = o;
    }
  }
}

// (m, l) partials -> natural-log LSE = M + log(sum_s l_s * exp(m_s - M)).  rows = obj_n*hw; parts strided by `rows`
// for every object: part index ((obj*n_split + s)*hw + j).
// n_split_dev (optional): the split chosen on the device (banks with live counts, vfn_tc.cu split_plan_kernel)
__global__ void lse_combine_kernel(const float2* __restrict__ part, int n_split, const int32_t* __restrict__ n_split_dev,
                                   int64_t hw, int obj_n, float* __restrict__ lse, float* __restrict__ lse_copy) {
  pdl_wait();
  pdl_trigger();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hw * obj_n) return;
  if (n_split_dev) n_split = *n_split_dev;
  const int64_t obj = idx / hw, j = idx % hw;
  float M = -INFINITY;
  for (int s = 0; s < n_split; ++s) M = fmaxf(M, part[((int64_t)obj * n_split + s) * hw + j].x);
  float L = 0.f;
  if (M > -INFINITY)
    for (int s = 0; s < n_split; ++s) {
      const float2 p = part[((int64_t)obj * n_split + s) * hw + j];
      if (p.x > -INFINITY) L += p.y * expf(p.x - M);
    }
  const float r = (M > -INFINITY) ? M + logf(L) : -INFINITY;
  lse[idx] = r;
  if (lse_copy) lse_copy[idx] = r;
}

// generic stacked (n_parts, rows, 2) -> lse (rows): used by the sharded-bank host path after all-gather
__global__ void lse_combine_flat_kernel(const float2* __restrict__ part, int n_parts, int64_t rows,
                                        float* __restrict__ lse) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows) return;
  float M = -INFINITY;
  for (int s = 0; s < n_parts; ++s) M = fmaxf(M, part[(int64_t)s * rows + idx].x);
  float L = 0.f;
  if (M > -INFINITY)
    for (int s = 0; s < n_parts; ++s) {
      const float2 p = part[(int64_t)s * rows + idx];
      if (p.x > -INFINITY) L += p.y * expf(p.x - M);
    }
  lse[idx] = (M > -INFINITY) ? M + logf(L) : -INFINITY;
}

__global__ void ml_merge_kernel(const float2* __restrict__ part, int n_split, const int32_t* __restrict__ n_split_dev,
                                int64_t hw, int obj_n,
                                float2* __restrict__ ml) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hw * obj_n) return;
  if (n_split_dev) n_split = *n_split_dev;
  const int64_t obj = idx / hw, j = idx % hw;
  float M = -INFINITY;
  for (int s = 0; s < n_split; ++s) M = fmaxf(M, part[((int64_t)obj * n_split + s) * hw + j].x);
  float L = 0.f;
  if (M > -INFINITY)
    for (int s = 0; s < n_split; ++s) {
      const float2 p = part[((int64_t)obj * n_split + s) * hw + j];
      if (p.x > -INFINITY) L += p.y * expf(p.x - M);
    }
  ml[idx] = make_float2(M, L);
}

__global__ void match_reduce_kernel(const float2* __restrict__ part, int n_split, int64_t hw,
                                    int32_t* __restrict__ idx_out, float* __restrict__ corr_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= hw) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int s = 0; s < n_split; ++s) {
    const float2 p = part[(int64_t)s * hw + j];
    const int pi = __float_as_int(p.y);
    if (p.x > best || (p.x == best && pi < bi)) { best = p.x; bi = pi; }
  }
  idx_out[j] = bi;
  corr_out[j] = best;
}

// ------------------------------------------------------------------------------------------------
// readout kernel: grid (q_tiles, n_split, obj_n * n_chunk) ; chunk = 128 value channels.
// p_ij = exp(s_ij - lse_j); usage counts (chunk 0 only); partial O[c][j] = sum_i V[i][c] p_ij
// partial layout: po[((obj*n_split + split) * d_val + c) * hw + j]
// ------------------------------------------------------------------------------------------------
struct ReadoutSmem {
  float As[TK][LDS_];
  float Bs[TK][LDS_];
  float Vs[TK][LDS_];
  float Ps[TM][LDS_];
};

__global__ void __launch_bounds__(ST_THREADS) simt_readout_kernel(BankSet banks, const float* __restrict__ Q, int64_t hw,
                                                                  int n_split, int n_chunk,
                                                                  const float* __restrict__ lse, float thres_valid,
                                                                  int do_count, float* __restrict__ po) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ReadoutSmem& sm = *reinterpret_cast<ReadoutSmem*>(smem_raw);
  const int obj = blockIdx.z / n_chunk, chunk = blockIdx.z % n_chunk;
  const vfn_bank bk = banks.b[obj];
  const int d = bk.d_key, dv = bk.d_val;
  const int64_t n = bk.n;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t j0 = (int64_t)blockIdx.x * TN;
  const int c0 = chunk * 128;
  const int64_t tiles = (n + TM - 1) / TM;
  const int64_t t_begin = tiles * blockIdx.y / n_split, t_end = tiles * (blockIdx.y + 1) / n_split;
  const float sqrt_d = sqrtf((float)d);
  float lse_j[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int64_t j = j0 + col_of(tx, b);
    lse_j[b] = (j < hw) ? lse[(int64_t)obj * hw + j] : INFINITY;
  }
  float acc[8][8], o[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) o[a][b] = 0.f;
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t i0 = t * TM;
    tile_scores(bk.keys, nullptr, n, Q, hw, d, i0, j0, sm.As, sm.Bs, acc);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int r = row_of(ty, a);
      const int64_t gi = i0 + r;
      int c = 0;
      float pv[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        float p = 0.f;
        if (gi < n && lse_j[b] < INFINITY) p = expf(__fdiv_rn(acc[a][b], sqrt_d) - lse_j[b]);
        pv[b] = p;
        c += (p > thres_valid) ? 1 : 0;      // AFB_URR.py:165 strict >
      }
      *reinterpret_cast<float4*>(&sm.Ps[r][tx * 4]) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      *reinterpret_cast<float4*>(&sm.Ps[r][64 + tx * 4]) = make_float4(pv[4], pv[5], pv[6], pv[7]);
      if (do_count && chunk == 0) {
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (tx == 0 && c > 0 && gi < n) atomicAdd(&bk.cnt[gi], c);
      }
    }
    // O[c][j] += sum_i V[i][c0+c] * P[i][j]
    for (int k0 = 0; k0 < TM; k0 += TK) {
      __syncthreads();   // Ps complete (first iter) / Vs free (later iters)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int idx = threadIdx.x + ST_THREADS * r;   // 16 rows x 32 float4
        const int row = idx >> 5, c4 = idx & 31;
        const int64_t gi = i0 + k0 + row;
        const int c = c0 + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < n && c < dv) v = *reinterpret_cast<const float4*>(bk.values + gi * dv + c);
        *reinterpret_cast<float4*>(&sm.Vs[row][c4 * 4]) = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sm.Vs[kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sm.Vs[kk][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.Ps[k0 + kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&sm.Ps[k0 + kk][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) o[a][b] = fmaf(av[a], bv[b], o[a][b]);
      }
    }
    __syncthreads();
  }
  float* dst = po + ((int64_t)obj * n_split + blockIdx.y) * dv * hw;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int c = c0 + row_of(ty, a);
    if (c >= dv) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int64_t j = j0 + col_of(tx, b);
      if (j < hw) dst[(int64_t)c * hw + j] = o[a][b];
    }
  }
}

// info[:,1] += log(cnt+1) ; cnt = 0 for slot i of one bank                                  (AFB_URR.py:174)
__device__ __forceinline__ void finalize_count(const vfn_bank& bk, int64_t i) {
  const int c = bk.cnt[i];
  if (c == 0) return;            // log(0 + 1) = 0: info unchanged (most slots of a large bank; skips the fp64 log)
  bk.cnt[i] = 0;
  // bank_cnt + 1 is exact in fp32; log evaluated in double and rounded once
  bk.info[2 * i + 1] += (float)log((double)((float)c + 1.0f));
}

// out[obj][c][j] = sum_s po[obj][s][c][j] (c < dv) ; out[obj][dv + c][j] = q_out[c][j]     (AFB_URR.py:159,176)
// counts != 0: the same launch also folds the usage counts of phase B into info (finalize_count) - they are complete
// when this kernel starts, and a separate 7-block launch cost as much as its gap on the stream
// V = 4: 128-bit accesses (plane % 4 == 0 and 16-byte aligned bases), same per-element summation order as V = 1
template <int V>
__global__ void __launch_bounds__(256) combine_out_kernel(const float* __restrict__ po, int n_split,
                                                          const int32_t* __restrict__ n_split_dev,
                                                          int64_t plane /* dv*hw */, int obj_n,
                                                          const float* __restrict__ q_out, float* __restrict__ out,
                                                          int with_qout, BankSet banks, int counts) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = plane * obj_n;
  if (n_split_dev) n_split = *n_split_dev;
  if (counts) {
    for (int o = 0; o < obj_n; ++o) {
      const vfn_bank bk = banks.b[o];
      const int64_t n = live_n(bk);
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        finalize_count(bk, i);
    }
  }
  if (V == 4) {
    // grid-stride over float4 positions: a few CTAs per SM, each thread keeps up to four partial planes in flight
    for (int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x * 4) {
      const int64_t obj = idx / plane, r = idx % plane;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* base = po + (int64_t)obj * n_split * plane + r;
      int k = 0;
      for (; k + 4 <= n_split; k += 4) {          // added in plane order
        const float4 a = __ldcs(reinterpret_cast<const float4*>(base + (k + 0) * plane));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(base + (k + 1) * plane));
        const float4 c = __ldcs(reinterpret_cast<const float4*>(base + (k + 2) * plane));
        const float4 d = __ldcs(reinterpret_cast<const float4*>(base + (k + 3) * plane));
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
        s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
        s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
      }
      for (; k < n_split; ++k) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(base + k * plane));
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      }
      if (with_qout) {
        *reinterpret_cast<float4*>(out + obj * 2 * plane + r) = s;
        *reinterpret_cast<float4*>(out + obj * 2 * plane + plane + r) = __ldg(reinterpret_cast<const float4*>(q_out + r));
      } else {
        *reinterpret_cast<float4*>(out + obj * plane + r) = s;
      }
    }
  } else {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t obj = idx / plane, r = idx % plane;
    float s = 0.f;
    for (int k = 0; k < n_split; ++k) s += po[((int64_t)obj * n_split + k) * plane + r];
    if (with_qout) {
      out[obj * 2 * plane + r] = s;
      out[obj * 2 * plane + plane + r] = q_out[r];
    } else {
      out[obj * plane + r] = s;
    }
  }
}

static void launch_combine_out(const float* po, int n_split, const int32_t* n_split_dev, int64_t plane, int obj_n,
                               const float* q_out, float* out, int with_qout, cudaStream_t st, const BankSet& banks,
                               int counts) {
  const bool vec = plane % 4 == 0 && ((uintptr_t)po % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                   (!with_qout || (uintptr_t)q_out % 16 == 0);
  const int64_t need = cdiv(plane * obj_n / 4, 256);
  if (vec)
    launch_pdl(combine_out_kernel<4>, dim3((unsigned)(need < 148 * 4 ? need : 148 * 4)), dim3(256), 0, st, po, n_split,
               n_split_dev, plane, obj_n, q_out, out, with_qout, banks, counts);
  else
    launch_pdl(combine_out_kernel<1>, dim3((unsigned)cdiv(plane * obj_n, 256)), dim3(256), 0, st, po, n_split, n_split_dev,
               plane, obj_n, q_out, out, with_qout, banks, counts);
}

// ------------------------------------------------------------------------------------------------
// host-side planning
// ------------------------------------------------------------------------------------------------
struct ReadPlan {
  int obj_n;
  int64_t hw, n_max;
  int d_key, d_val;
  int q_tiles, split_a, split_b, n_chunk;
  bool tc;
  const int32_t *dev_a, *dev_b;      // splits chosen on the device (banks with live counts), else NULL
  size_t off_q, off_part, off_lse, off_po, off_tc, total;
};

static int pick_split(int64_t tiles, int base_ctas, int target) {
  int s = (int)cdiv(target, base_ctas > 0 ? base_ctas : 1);
  if (s < 1) s = 1;
  if (s > tiles) s = (int)(tiles > 0 ? tiles : 1);
  if (s > 64) s = 64;
  return s;
}

static ReadPlan make_read_plan(int obj_n, int64_t n_max, int64_t hw, int d_key, int d_val, int impl) {
  ReadPlan p{};
  p.obj_n = obj_n; p.hw = hw; p.n_max = n_max; p.d_key = d_key; p.d_val = d_val;
  p.tc = (impl != 1) && tc_shapes_ok(d_key, d_val);
  p.q_tiles = (int)cdiv(hw, TN);
  p.n_chunk = (int)cdiv(d_val, 128);
  const int64_t tiles = cdiv(n_max > 0 ? n_max : 1, TM);
  if (p.tc) {
    tc_pick_splits(obj_n, n_max, hw, &p.split_a, &p.split_b);
  } else {
    p.split_a = pick_split(tiles, p.q_tiles * obj_n, 148 * 4);
    p.split_b = pick_split(tiles, p.q_tiles * obj_n * p.n_chunk, 148 * 2);
  }
  size_t o = 0;
  p.off_q = o;    o += align_up((size_t)hw * d_key * sizeof(float), 256);
  p.off_part = o; o += align_up((size_t)obj_n * p.split_a * hw * sizeof(float2), 256);
  p.off_lse = o;  o += align_up((size_t)obj_n * hw * sizeof(float), 256);
  p.off_po = o;   o += align_up((size_t)obj_n * p.split_b * d_val * hw * sizeof(float), 256);
  p.off_tc = o;   o += p.tc ? tc_workspace_bytes(obj_n, hw) : 0;
  p.total = o;
  return p;
}

static int check_banks(const vfn_bank* banks, int obj_n, int64_t* n_max, BankSet* set) {
  VFN_CHECK_ARG(banks && obj_n >= 1 && obj_n <= MAX_OBJ, "obj_n=%d out of range (1..%d)", obj_n, MAX_OBJ);
  *n_max = 0;
  for (int i = 0; i < obj_n; ++i) {
    VFN_CHECK_ARG(banks[i].d_key == banks[0].d_key && banks[i].d_val == banks[0].d_val, "banks differ in dims");
    VFN_CHECK_ARG(banks[i].d_key % 8 == 0 && banks[i].d_val % 8 == 0 && banks[i].d_key > 0, "dims must be multiples of 8");
    VFN_CHECK_ARG(banks[i].n >= 1 && banks[i].n <= banks[i].cap, "bank %d: n=%lld invalid", i, (long long)banks[i].n);
    if (banks[i].n > *n_max) *n_max = banks[i].n;
    set->b[i] = banks[i];
  }
  return VFN_OK;
}

static int run_phase_a(const BankSet& set, ReadPlan& p, const float* q_in_dm, char* ws, cudaStream_t st, int q_em = 0) {
  float* Q = reinterpret_cast<float*>(ws + p.off_q);
  float2* part = reinterpret_cast<float2*>(ws + p.off_part);
  if (p.tc) return tc_phase_a(set.b, p.obj_n, q_in_dm, p.hw, p.split_a, part, ws + p.off_tc, st, &p.split_a, &p.dev_a, q_em);
  for (int o = 0; o < p.obj_n; ++o)
    VFN_CHECK_ARG(!set.b[o].n_live || set.b[o].n_min == set.b[o].n,
                  "the fp32 SIMT read needs exact bank sizes (bank %d was passed with bounds)", o);
  {
    PrepJob jb{q_in_dm, p.d_key, p.hw, Q, nullptr, nullptr, nullptr, 1.f, 0, q_em};
    if (int rc = launch_prep(&jb, 1, st)) return rc;
  }
  dim3 grid(p.q_tiles, p.split_a, p.obj_n);
  double work = 0;
  for (int o = 0; o < p.obj_n; ++o) work += 2.0 * p.d_key * (double)set.b[o].n * (double)p.hw;
  prof_begin(PROF_READ_A, st);
  simt_score_kernel<MODE_LSE><<<grid, ST_THREADS, 0, st>>>(set, Q, p.hw, p.split_a, part);
  prof_end(PROF_READ_A, st, work);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

static int run_phase_b(const BankSet& set, ReadPlan& p, const float* q_in_dm, int q_em, const float* lse,
                       float thres_valid, int update_bank, char* ws, cudaStream_t st) {
  float* Q = reinterpret_cast<float*>(ws + p.off_q);
  float* po = reinterpret_cast<float*>(ws + p.off_po);
  if (p.tc) return tc_phase_b(set.b, p.obj_n, q_in_dm, q_em, p.hw, p.split_b, lse, thres_valid, update_bank, po,
                              ws + p.off_tc, st, &p.split_b, &p.dev_b);
  for (int o = 0; o < p.obj_n; ++o)
    VFN_CHECK_ARG(!set.b[o].n_live || set.b[o].n_min == set.b[o].n,
                  "the fp32 SIMT read needs exact bank sizes (bank %d was passed with bounds)", o);
  static bool attr_set[64] = {false};
  if (first_use_on_device(attr_set))
    VFN_CUDA_OK(cudaFuncSetAttribute(simt_readout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ReadoutSmem)));
  dim3 grid(p.q_tiles, p.split_b, p.obj_n * p.n_chunk);
  double work = 0;
  for (int o = 0; o < p.obj_n; ++o) work += 2.0 * p.d_val * (double)set.b[o].n * (double)p.hw;
  prof_begin(PROF_READ_B, st);
  simt_readout_kernel<<<grid, ST_THREADS, sizeof(ReadoutSmem), st>>>(set, Q, p.hw, p.split_b, p.n_chunk, lse,
                                                                     thres_valid, update_bank, po);
  prof_end(PROF_READ_B, st, work);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

size_t vfn_memread_workspace_bytes(int32_t obj_n, int64_t n_max, int64_t hw, int32_t d_key, int32_t d_val) {
  if (obj_n < 1 || hw < 1 || d_key < 1 || d_val < 1) return 0;
  // sized for whichever implementation needs more
  ReadPlan a = make_read_plan(obj_n, n_max, hw, d_key, d_val, 1);
  ReadPlan b = make_read_plan(obj_n, n_max, hw, d_key, d_val, 0);
  return a.total > b.total ? a.total : b.total;
}

int vfn_memread_phase_a(const vfn_bank* banks, int32_t obj_n, const float* d_q_in_dm, int64_t hw, float* d_ml,
                        void* d_ws, size_t ws_bytes, int32_t impl, void* stream) {
  BankSet set;
  int64_t n_max;
  if (int rc = check_banks(banks, obj_n, &n_max, &set)) return rc;
  VFN_CHECK_ARG(d_q_in_dm && d_ml && d_ws && hw > 0, "memread_phase_a: bad args");
  const int q_em = (impl & VFN_Q_IN_EM) ? 1 : 0;
  impl &= 0xff;
  if (impl == 2 && !tc_shapes_ok(set.b[0].d_key, set.b[0].d_val)) {
    set_error("tcgen05 read needs d_key=128, d_val=512");
    return VFN_E_UNSUPPORTED;
  }
  ReadPlan p = make_read_plan(obj_n, n_max, hw, set.b[0].d_key, set.b[0].d_val, impl);
  if (ws_bytes < p.total) { set_error("memread: workspace %zu < %zu", ws_bytes, p.total); return VFN_E_CAPACITY; }
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(d_ws);
  if (int rc = run_phase_a(set, p, d_q_in_dm, ws, st, q_em)) return rc;
  const int64_t rows = hw * obj_n;
  ml_merge_kernel<<<(unsigned)cdiv(rows, 256), 256, 0, st>>>(reinterpret_cast<float2*>(ws + p.off_part), p.split_a,
                                                             p.dev_a, hw, obj_n, reinterpret_cast<float2*>(d_ml));
  VFN_LAUNCH_OK();
  return VFN_OK;
}

int vfn_lse_combine(const float* d_ml, int32_t n_parts, int64_t n_rows, float* d_lse, void* stream) {
  VFN_CHECK_ARG(d_ml && d_lse && n_parts >= 1 && n_rows >= 1, "lse_combine: bad args");
  lse_combine_flat_kernel<<<(unsigned)cdiv(n_rows, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(d_ml), n_parts, n_rows, d_lse);
  VFN_LAUNCH_OK();
  return VFN_OK;
}

int vfn_memread_phase_b(const vfn_bank* banks, int32_t obj_n, const float* d_q_in_dm, int64_t hw, const float* d_lse,
                        float thres_valid, int32_t update_bank, float* d_partial_out, void* d_ws, size_t ws_bytes,
                        int32_t impl, void* stream) {
  BankSet set;
  int64_t n_max;
  if (int rc = check_banks(banks, obj_n, &n_max, &set)) return rc;
  VFN_CHECK_ARG(d_q_in_dm && d_lse && d_partial_out && d_ws && hw > 0, "memread_phase_b: bad args");
  const int q_em = (impl & VFN_Q_IN_EM) ? 1 : 0;
  impl &= 0xff;
  ReadPlan p = make_read_plan(obj_n, n_max, hw, set.b[0].d_key, set.b[0].d_val, impl);
  if (ws_bytes < p.total) { set_error("memread: workspace %zu < %zu", ws_bytes, p.total); return VFN_E_CAPACITY; }
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(d_ws);
  // phase A of the same call sequence left Q in the workspace (fp32 SIMT kernels; the tcgen05 kernels read d_q_in_dm)
  if (int rc = run_phase_b(set, p, d_q_in_dm, q_em, d_lse, thres_valid, update_bank, ws, st)) return rc;
  const int64_t plane = (int64_t)set.b[0].d_val * hw;
  launch_combine_out(reinterpret_cast<float*>(ws + p.off_po), p.split_b, p.dev_b, plane, obj_n, nullptr, d_partial_out, 0,
                     st, set, update_bank);
  VFN_LAUNCH_OK();
  return VFN_OK;
}

int vfn_memread(const vfn_bank* banks, int32_t obj_n, const float* d_q_in_dm, const float* d_q_out_dm, int64_t hw,
                float thres_valid, int32_t update_bank, float* d_out, float* d_lse, void* d_ws, size_t ws_bytes,
                int32_t impl, void* stream) {
  BankSet set;
  int64_t n_max;
  if (int rc = check_banks(banks, obj_n, &n_max, &set)) return rc;
  VFN_CHECK_ARG(d_q_in_dm && d_q_out_dm && d_out && d_ws && hw > 0, "memread: bad args");
  const int q_em = (impl & VFN_Q_IN_EM) ? 1 : 0;
  impl &= 0xff;
  if (impl == 2 && !tc_shapes_ok(set.b[0].d_key, set.b[0].d_val)) {
    set_error("tcgen05 read needs d_key=128, d_val=512");
    return VFN_E_UNSUPPORTED;
  }
  ReadPlan p = make_read_plan(obj_n, n_max, hw, set.b[0].d_key, set.b[0].d_val, impl);
  if (ws_bytes < p.total) { set_error("memread: workspace %zu < %zu", ws_bytes, p.total); return VFN_E_CAPACITY; }
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(d_ws);
  if (int rc = run_phase_a(set, p, d_q_in_dm, ws, st, q_em)) return rc;
  float* lse = reinterpret_cast<float*>(ws + p.off_lse);
  const int64_t rows = hw * obj_n;
  launch_pdl(lse_combine_kernel, dim3((unsigned)cdiv(rows, 256)), dim3(256), 0, st,
             reinterpret_cast<const float2*>(ws + p.off_part), p.split_a, p.dev_a, hw, obj_n, lse, d_lse);
  VFN_LAUNCH_OK();
  if (int rc = run_phase_b(set, p, d_q_in_dm, q_em, lse, thres_valid, update_bank, ws, st)) return rc;
  const int64_t plane = (int64_t)set.b[0].d_val * hw;
  launch_combine_out(reinterpret_cast<float*>(ws + p.off_po), p.split_b, p.dev_b, plane, obj_n, d_q_out_dm, d_out, 1, st,
                     set, update_bank);
  VFN_LAUNCH_OK();
  count_launches(2);
  return VFN_OK;
}

size_t vfn_bank_match_workspace_bytes(int64_t n, int64_t hw) {
  (void)n;
  const size_t simt = align_up((size_t)64 * hw * sizeof(float2), 256);
  const size_t tc = tc_match_workspace_bytes(1, hw);
  return simt > tc ? simt : tc;
}

int vfn_bank_match(const vfn_bank* bank, const float* d_nck_em, int64_t hw, int32_t* d_match_idx, float* d_match_corr,
                   void* d_ws, size_t ws_bytes, int32_t impl, void* stream) {
  BankSet set;
  int64_t n_max;
  if (int rc = check_banks(bank, 1, &n_max, &set)) return rc;
  VFN_CHECK_ARG(d_nck_em && d_match_idx && d_match_corr && d_ws && hw > 0, "match: bad args");
  cudaStream_t st = as_stream(stream);
  if (ws_bytes < vfn_bank_match_workspace_bytes(n_max, hw)) { set_error("match: workspace too small"); return VFN_E_CAPACITY; }
  int split = 0;
  const bool tc = (impl != 1) && set.b[0].d_key == 128 && set.b[0].nkh != nullptr && vfn_device_is_sm100();
  if (impl == 2 && !tc) { set_error("tcgen05 match needs d_key = 128 on an sm_100 device"); return VFN_E_UNSUPPORTED; }
  if (tc) return tc_match(&set.b[0], 1, &d_nck_em, hw, reinterpret_cast<char*>(d_ws), 0, &d_match_idx, &d_match_corr, st);
  {
    const int q_tiles = (int)cdiv(hw, TN);
    split = pick_split(cdiv(n_max, TM), q_tiles, 148 * 4);
    dim3 grid(q_tiles, split, 1);
    prof_begin(PROF_MATCH, st);
    simt_score_kernel<MODE_MATCH><<<grid, ST_THREADS, 0, st>>>(set, d_nck_em, hw, split,
                                                               reinterpret_cast<float2*>(d_ws));
    prof_end(PROF_MATCH, st, 2.0 * set.b[0].d_key * (double)n_max * (double)hw);
    count_launches(1);
  }
  count_launches(1);
  match_reduce_kernel<<<(unsigned)cdiv(hw, 256), 256, 0, st>>>(reinterpret_cast<float2*>(d_ws), split, hw, d_match_idx,
                                                               d_match_corr);
  VFN_LAUNCH_OK();
  return VFN_OK;
}

}  // extern "C"
