// Feature-bank bookkeeping kernels (bandwidth-bound): candidate preparation, append, merge (segmented
// mean in fixed order), LFU eviction planning, order-preserving compaction, info clamp.
// Reference behaviour restated: video_module/model/FeatureBank.py:27-143 (see include/vfn.h per entry point).
#include "vfn_common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace vfn {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// measurement hooks
// ------------------------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int kind; double work; };
// The library may be driven from several host threads (one per rank in the thread-rank sharded tests): the launch counter
// is atomic, the open event of a bracket is per thread, and the shared record list / event pool sit behind one mutex.
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;
static thread_local cudaEvent_t g_open[PROF_KINDS];
static std::atomic<long long> g_launches{0};

static cudaEvent_t get_event() {   // caller holds g_prof_mu
  if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof_on) return;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_open[kind] = get_event();
  }
  cudaEventRecord(g_open[kind], st);
}
void prof_end(int kind, cudaStream_t st, double work) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  g_prof.push_back({g_open[kind], e, kind, work});
}
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// prep_rows: (d, n) dimension-major [or (n, d) entry-major when src_em]  ->  (n, d) entry-major raw / L2-normalised /
// fp16 hi+lo.
// One launch serves several independent jobs (blockIdx.y): all candidate tensors of a frame in one go.
// hi/lo receive the split of raw*scale, or of normalised*scale when `split_normed` is set.
// ------------------------------------------------------------------------------------------------
constexpr int PREP_MAX_JOBS = 16;
constexpr int PREP_CHUNK = 128;
struct PrepJobs { PrepJob j[PREP_MAX_JOBS]; };

__global__ void __launch_bounds__(256) prep_rows_kernel(const __grid_constant__ PrepJobs jobs) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  __shared__ float part[8][32];
  __shared__ float denom[32];
  __shared__ float stage[32][129];      // entry-major sources: 128 columns of the CTA's 32 rows, staged for the norm
  const PrepJob& jb = jobs.j[blockIdx.y];
  const float* __restrict__ src = jb.src;
  const int d = jb.d;
  const int64_t n = jb.n;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * 32;
  // jobs without a norm (the read's query operand) are cut into 32-row chunks: four times as many CTAs for a job that
  // would otherwise run on 51; normalising jobs keep PREP_CHUNK rows per CTA (every CTA re-reads the whole column for
  // the norm)
  const int chunk = (jb.normed || jb.split_normed) ? PREP_CHUNK : 32;
  if ((int)blockIdx.z * chunk >= d) return;
  if (q0 >= n) {
    // zero padding of the operand arrays beyond the last row (rows [n, n_pad)); blocks past n_pad have nothing to do
    if (jb.hi && q0 < jb.n_pad) {
      const int k_end = min(d, (int)(blockIdx.z + 1) * chunk);
      for (int r = ty; r < 32; r += 8) {
        const int64_t qq = q0 + r;
        if (qq >= jb.n_pad) break;
        for (int k = blockIdx.z * chunk + tx; k < k_end; k += 32) {
          jb.hi[qq * d + k] = 0;
          if (jb.lo) jb.lo[qq * d + k] = 0;
        }
      }
    }
    return;
  }
  const int64_t q = q0 + tx;
  if (jb.normed || jb.split_normed) {
    // same summation order for both source layouts (8 strided partial sums per row, then their sum in order)
    float ss = 0.f;
    if (jb.src_em) {
      // rows are contiguous: stage 128 columns at a time with coalesced loads, then every thread walks ITS row in the
      // same k order as below (ty, ty + 8, ...; the chain carries over the stages), so both layouts give the same bits
      for (int kc = 0; kc < d; kc += 128) {
        for (int r = ty; r < 32; r += 8) {
          const int64_t qq = q0 + r;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kc + tx + 32 * j;
            stage[r][tx + 32 * j] = (qq < n && k < d) ? src[qq * d + k] : 0.f;
          }
        }
        __syncthreads();
        const int kend = min(128, d - kc);
        for (int k = ty; k < kend; k += 8) {
          const float v = stage[tx][k];
          ss = fmaf(v, v, ss);
        }
        __syncthreads();
      }
    } else if (q < n)
      for (int k = ty; k < d; k += 8) {
        float v = src[(int64_t)k * n + q];
        ss = fmaf(v, v, ss);
      }
    part[ty][tx] = ss;
    __syncthreads();
    if (ty == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += part[i][tx];
      denom[tx] = fmaxf(sqrtf(t), 1e-12f);   // NF.normalize eps
    }
    __syncthreads();
  }
  // blockIdx.z takes PREP_CHUNK of the d rows (the norm above is over all of them: cheap re-reads from L2), so that a
  // 512-channel job spreads over four times as many CTAs as a 128-channel one
  const int k_end = min(d, (int)(blockIdx.z + 1) * chunk);
  for (int k0 = blockIdx.z * chunk; k0 < k_end; k0 += 32) {
    if (jb.src_em) {
      // entry-major source (vfn_keyvalue's output): rows are already contiguous in k; the tile is filled transposed
      for (int r = ty; r < 32; r += 8) {
        const int64_t qq = q0 + r;
        const int k = k0 + tx;
        tile[tx][r] = (k < d && qq < n) ? src[qq * d + k] : 0.f;
      }
    } else {
      for (int r = ty; r < 32; r += 8) {
        int k = k0 + r;
        tile[r][tx] = (k < d && q < n) ? src[(int64_t)k * n + q] : 0.f;
      }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      int64_t qq = q0 + r;
      int k = k0 + tx;
      if (qq >= n && qq < jb.n_pad && k < d && jb.hi) {      // pad rows inside the last row block
        jb.hi[qq * d + k] = 0;
        if (jb.lo) jb.lo[qq * d + k] = 0;
      }
      if (qq < n && k < d) {
        float v = tile[tx][r];
        int64_t o = qq * d + k;
        if (jb.raw) jb.raw[o] = v;
        float nv = 0.f;
        if (jb.normed || jb.split_normed) nv = v / denom[r];
        if (jb.normed) jb.normed[o] = nv;
        if (jb.hi) {
          uint16_t h, l;
          split_f16((jb.split_normed ? nv : v) * jb.scale, h, l);
          jb.hi[o] = h;
          if (jb.lo) jb.lo[o] = l;
        }
      }
    }
    __syncthreads();
  }
}

int launch_prep(const PrepJob* jobs, int n_jobs, cudaStream_t st) {
  VFN_CHECK_ARG(n_jobs >= 1 && n_jobs <= PREP_MAX_JOBS, "prep: too many jobs");
  PrepJobs pj;
  int64_t n_max = 0;
  int d_max = 0;
  for (int i = 0; i < n_jobs; ++i) {
    pj.j[i] = jobs[i];
    if (jobs[i].n > n_max) n_max = jobs[i].n;
    if (jobs[i].hi && jobs[i].n_pad > n_max) n_max = jobs[i].n_pad;
    if (jobs[i].d > d_max) d_max = jobs[i].d;
  }
  if (n_max == 0) return VFN_OK;
  int z_max = 1;
  for (int i = 0; i < n_jobs; ++i) {
    const int chunk = (jobs[i].normed || jobs[i].split_normed) ? PREP_CHUNK : 32;
    z_max = max(z_max, (int)cdiv(jobs[i].d, chunk));
  }
  dim3 block(32, 8), grid((unsigned)cdiv(n_max, 32), n_jobs, (unsigned)z_max);
  VFN_CUDA_OK(launch_pdl(prep_rows_kernel, grid, block, 0, st, pj));
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

// ------------------------------------------------------------------------------------------------
// row helpers: one block owns one bank slot
// ------------------------------------------------------------------------------------------------
// append: grid.x = upper bound on selected rows; block = 128 threads; row i -> slot bank.n + i
struct AppendObj {
  vfn_bank bank; const float *ck, *cv, *nck; const int32_t *sel, *n_sel_dev; int64_t n_sel_upper;
};
struct AppendSet { AppendObj o[4]; };
// grid (rows, objects)
__global__ void __launch_bounds__(128) append_rows_kernel(const __grid_constant__ AppendSet set, float info0,
                                                          float info1) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const AppendObj& ao = set.o[blockIdx.y];
  const vfn_bank& bank = ao.bank;
  const float* __restrict__ ck = ao.ck;
  const float* __restrict__ cv = ao.cv;
  const float* __restrict__ nck = ao.nck;
  const int32_t* __restrict__ sel = ao.sel;
  const int32_t* __restrict__ n_sel_dev = ao.n_sel_dev;
  const int64_t n_sel_upper = ao.n_sel_upper;
  const int64_t n_sel = n_sel_dev ? (int64_t)*n_sel_dev : n_sel_upper;
  const int64_t n_base = live_n(bank);          // the live count is only advanced by the clamp kernel that follows
  const int dk4 = bank.d_key >> 2, dv4 = bank.d_val >> 2;
  for (int64_t i = blockIdx.x; i < n_sel; i += gridDim.x) {
    const int64_t s = sel ? (int64_t)sel[i] : i;
    const int64_t dst = n_base + i;
    const float4* ks = reinterpret_cast<const float4*>(ck + s * bank.d_key);
    const float4* vs = reinterpret_cast<const float4*>(cv + s * bank.d_val);
    float ss = 0.f;
    for (int f = threadIdx.x; f < dk4; f += blockDim.x) {
      float4 v = ks[f];
      reinterpret_cast<float4*>(bank.keys + dst * bank.d_key)[f] = v;
      if (bank.kh) store_key_ops4(bank.kh, bank.kl, dst * bank.d_key + 4 * f, v);
      if (nck) store_nk4(bank.nk, bank.nkh, bank.nkl, dst * bank.d_key + 4 * f,
                         reinterpret_cast<const float4*>(nck + s * bank.d_key)[f]);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    if (!nck) {   // derive the normalised key here (init_bank / append API path)
      float tot = block_sum(ss, red);
      float den = fmaxf(sqrtf(tot), 1e-12f);
      for (int f = threadIdx.x; f < dk4; f += blockDim.x) {
        float4 v = ks[f];
        store_nk4(bank.nk, bank.nkh, bank.nkl, dst * bank.d_key + 4 * f,
                  make_float4(v.x / den, v.y / den, v.z / den, v.w / den));
      }
    }
    for (int f = threadIdx.x; f < dv4; f += blockDim.x) {
      float4 v = vs[f];
      reinterpret_cast<float4*>(bank.values + dst * bank.d_val)[f] = v;
      if (bank.vh) store_val_ops4(bank.vh, bank.v8, bank.vl, dst * bank.d_val + 4 * f, v);
    }
    if (threadIdx.x == 0) {
      bank.info[dst * 2 + 0] = info0;
      bank.info[dst * 2 + 1] = info1;
      bank.cnt[dst] = 0;
    }
  }
}

// append, AFB-URR dims with tensor-core operand arrays and normalised candidates supplied (the update path): one WARP
// per row, all six 16-byte loads of the row (raw key, normalised key, four value quarters) in flight before the first
// store; the same element-wise conversions as append_rows_kernel, so the rows are bit-identical.
__global__ void __launch_bounds__(128) append_rows_warp_kernel(const __grid_constant__ AppendSet set, float info0,
                                                               float info1) {
  pdl_wait();
  pdl_trigger();
  const AppendObj& ao = set.o[blockIdx.y];
  const vfn_bank& bank = ao.bank;
  const int64_t n_sel = ao.n_sel_dev ? (int64_t)*ao.n_sel_dev : ao.n_sel_upper;
  const int64_t n_base = live_n(bank);
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 4;
  for (int64_t i = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); i < n_sel; i += warps) {
    const int64_t s = ao.sel ? (int64_t)ao.sel[i] : i;
    const int64_t d = n_base + i;
    const float4 k = __ldg(reinterpret_cast<const float4*>(ao.ck + s * 128) + lane);
    const float4 nk = __ldg(reinterpret_cast<const float4*>(ao.nck + s * 128) + lane);
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(ao.cv + s * 512) + lane + 32 * j);
    reinterpret_cast<float4*>(bank.keys + d * 128)[lane] = k;
    store_key_ops4(bank.kh, bank.kl, d * 128 + 4 * lane, k);
    store_nk4(bank.nk, bank.nkh, bank.nkl, d * 128 + 4 * lane, nk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      reinterpret_cast<float4*>(bank.values + d * 512)[lane + 32 * j] = v[j];
      store_val_ops4(bank.vh, bank.v8, bank.vl, d * 512 + 4 * (lane + 32 * j), v[j]);
    }
    if (lane == 0) {
      bank.info[d * 2 + 0] = info0;
      bank.info[d * 2 + 1] = info1;
      bank.cnt[d] = 0;
    }
  }
}

__global__ void __launch_bounds__(128) refresh_rows_kernel(vfn_bank bank, int64_t first, int64_t count) {
  __shared__ float red[32];
  const int dk4 = bank.d_key >> 2, dv4 = bank.d_val >> 2;
  for (int64_t i = blockIdx.x; i < count; i += gridDim.x) {
    const int64_t u = first + i;
    const float4* ks = reinterpret_cast<const float4*>(bank.keys + u * bank.d_key);
    float ss = 0.f;
    for (int f = threadIdx.x; f < dk4; f += blockDim.x) {
      float4 v = ks[f];
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    float den = fmaxf(sqrtf(block_sum(ss, red)), 1e-12f);
    for (int f = threadIdx.x; f < dk4; f += blockDim.x) {
      float4 v = ks[f];
      store_nk4(bank.nk, bank.nkh, bank.nkl, u * bank.d_key + 4 * f,
                make_float4(v.x / den, v.y / den, v.z / den, v.w / den));
      if (bank.kh) store_key_ops4(bank.kh, bank.kl, u * bank.d_key + 4 * f, v);
    }
    if (bank.vh)
      for (int f = threadIdx.x; f < dv4; f += blockDim.x)
        store_val_ops4(bank.vh, bank.v8, bank.vl, u * bank.d_val + 4 * f,
                       reinterpret_cast<const float4*>(bank.values + u * bank.d_val)[f]);
  }
}

// ------------------------------------------------------------------------------------------------
// plan: classify, order-preserving compaction of append set, sort of (slot,q) merge pairs, run detection.
// One CTA of 1024 threads; hw <= 65536.
// ------------------------------------------------------------------------------------------------
constexpr int PLAN_THREADS = 1024;
constexpr int PLAN_SMEM_KEYS = 4096;

// exclusive scan of one int per thread over the block; returns exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot /*[33]*/, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int t = (lane < nw) ? warp_tot[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;          // exclusive warp offsets
    if (lane == 31) warp_tot[32] = ti;  // block total
  }
  __syncthreads();
  *total = warp_tot[32];
  return warp_tot[wid] + inc - v;
}

__device__ __forceinline__ void plan_body(const int32_t* __restrict__ match_idx, const float* __restrict__ match_corr,
                                          int hw, float thres, int32_t* __restrict__ merge_q,
                                          int32_t* __restrict__ merge_slot, int32_t* __restrict__ run_off,
                                          int32_t* __restrict__ append_q, int32_t* __restrict__ counts,
                                          int32_t* __restrict__ h_counts, unsigned long long* __restrict__ gkeys,
                                          int32_t* __restrict__ n_live) {
  __shared__ unsigned long long skeys[PLAN_SMEM_KEYS];
  __shared__ int wt[33];
  const int tid = threadIdx.x;
  int off_m = 0, off_a = 0;
  for (int base = 0; base < hw; base += PLAN_THREADS) {
    const int q = base + tid;
    float c = (q < hw) ? match_corr[q] : 0.f;
    const int fm = (q < hw) && (c > thres);     // FeatureBank.py:71  strict >
    const int fa = (q < hw) && (c <= thres);    // FeatureBank.py:100 ; NaN goes nowhere
    int tot_m, tot_a;
    const int pm = block_excl_scan(fm, wt, &tot_m);
    const int pa = block_excl_scan(fa, wt, &tot_a);
    if (fm) gkeys[off_m + pm] = ((unsigned long long)(uint32_t)match_idx[q] << 32) | (uint32_t)q;
    if (fa) append_q[off_a + pa] = q;
    off_m += tot_m;
    off_a += tot_a;
  }
  const int n_merge = off_m;
  int npad = 1;
  while (npad < n_merge) npad <<= 1;
  for (int i = n_merge + tid; i < npad; i += PLAN_THREADS) gkeys[i] = ~0ull;
  __syncthreads();
  unsigned long long* buf = gkeys;
  if (npad <= PLAN_SMEM_KEYS) {
    for (int i = tid; i < npad; i += PLAN_THREADS) skeys[i] = gkeys[i];
    buf = skeys;
    __syncthreads();
  }
  // bitonic sort ascending by (slot, q): unique keys
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += PLAN_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          unsigned long long a = buf[i], b = buf[l];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) {
            buf[i] = b;
            buf[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  // unpack + run starts (unique touched slots ascending == torch.unique order, FeatureBank.py:73)
  int off_r = 0;
  for (int base = 0; base < n_merge; base += PLAN_THREADS) {
    const int i = base + tid;
    int fr = 0;
    if (i < n_merge) {
      unsigned long long kv = buf[i];
      const int slot = (int)(kv >> 32);
      merge_slot[i] = slot;
      merge_q[i] = (int)(kv & 0xffffffffu);
      fr = (i == 0) || ((int)(buf[i - 1] >> 32) != slot);
    }
    int tot_r;
    const int pr = block_excl_scan(fr, wt, &tot_r);
    if (fr) run_off[off_r + pr] = i;
    off_r += tot_r;
  }
  if (tid == 0) {
    run_off[off_r] = n_merge;
    counts[0] = n_merge;
    counts[1] = off_r;
    counts[2] = off_a;
    // live count after the append that follows (the clamp kernel at the end of the update commits it to n_live[0])
    const int n_next = n_live ? n_live[0] + off_a : 0;
    if (n_live) n_live[1] = n_next;
    counts[3] = n_next;
    if (h_counts) {
      h_counts[0] = n_merge;
      h_counts[1] = off_r;
      h_counts[2] = off_a;
      h_counts[3] = n_next;
    }
  }
}

struct PlanObj {
  const int32_t* match_idx; const float* match_corr; int32_t *merge_q, *merge_slot, *run_off, *append_q, *counts, *h_counts;
  unsigned long long* gkeys;
  int32_t* n_live;
};
struct PlanSet { PlanObj o[4]; };
// one CTA per object (blockIdx.x)
__global__ void __launch_bounds__(PLAN_THREADS) plan_kernel(const __grid_constant__ PlanSet set, int hw, float thres) {
  pdl_wait();
  pdl_trigger();
  const PlanObj& p = set.o[blockIdx.x];
  plan_body(p.match_idx, p.match_corr, hw, thres, p.merge_q, p.merge_slot, p.run_off, p.append_q, p.counts, p.h_counts,
            p.gkeys, p.n_live);
}

// ------------------------------------------------------------------------------------------------
// merge: one block per run of equal slot.  192 threads, one float4 of the [key | value] row per thread.
// ------------------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 192;

struct MergeObj {
  vfn_bank bank; const float *nck, *ncv; const int32_t *merge_q, *merge_slot, *run_off, *counts;
};
struct MergeSet { MergeObj o[4]; };
// grid (runs, objects)
__global__ void __launch_bounds__(MERGE_THREADS) merge_runs_kernel(const __grid_constant__ MergeSet set, float omr,
                                                                   float r) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const MergeObj& mo = set.o[blockIdx.y];
  const vfn_bank& bank = mo.bank;
  const float* __restrict__ nck = mo.nck;
  const float* __restrict__ ncv = mo.ncv;
  const int32_t* __restrict__ merge_q = mo.merge_q;
  const int32_t* __restrict__ merge_slot = mo.merge_slot;
  const int32_t* __restrict__ run_off = mo.run_off;
  const int32_t* __restrict__ counts = mo.counts;
  const int n_runs = counts[1];
  const int dk4 = bank.d_key >> 2, dv4 = bank.d_val >> 2;
  const int f = threadIdx.x;
  const bool is_key = f < dk4;
  const bool is_val = !is_key && (f - dk4) < dv4;
  for (int run = blockIdx.x; run < n_runs; run += gridDim.x) {
    const int b = run_off[run], e = run_off[run + 1];
    const int64_t u = merge_slot[b];
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (is_key) x = reinterpret_cast<const float4*>(bank.keys + u * bank.d_key)[f];
    if (is_val) x = reinterpret_cast<const float4*>(bank.values + u * bank.d_val)[f - dk4];
    const float sq = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w)));
    const float ssk = block_sum(is_key ? sq : 0.f, red);
    const float ssv = block_sum(is_val ? sq : 0.f, red);
    const float mag = sqrtf(is_key ? ssk : ssv);            // .norm(p=2, dim=0)   FeatureBank.py:65,89
    const float den = fmaxf(mag, 1e-12f);                   // NF.normalize        FeatureBank.py:63,87
    // scatter_mean: sequential sum in ascending candidate order, then / count  (FeatureBank.py:78,92)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (is_key || is_val) {
      for (int m = b; m < e; ++m) {
        const int64_t q = merge_q[m];
        const float4 s = is_key ? reinterpret_cast<const float4*>(nck + q * bank.d_key)[f]
                                : reinterpret_cast<const float4*>(ncv + q * bank.d_val)[f - dk4];
        acc.x = __fadd_rn(acc.x, s.x); acc.y = __fadd_rn(acc.y, s.y);
        acc.z = __fadd_rn(acc.z, s.z); acc.w = __fadd_rn(acc.w, s.w);
      }
    }
    const float cntf = (float)(e - b);
    float4 nw;
    {
      // mag * ((1-r) * x/|x| + r * mean), each op rounded separately like the reference's ATen chain (:81-84)
      float mx = __fdiv_rn(acc.x, cntf), my = __fdiv_rn(acc.y, cntf), mz = __fdiv_rn(acc.z, cntf), mw = __fdiv_rn(acc.w, cntf);
      float nx = __fdiv_rn(x.x, den), ny = __fdiv_rn(x.y, den), nz = __fdiv_rn(x.z, den), nwv = __fdiv_rn(x.w, den);
      nw.x = __fmul_rn(mag, __fadd_rn(__fmul_rn(omr, nx), __fmul_rn(r, mx)));
      nw.y = __fmul_rn(mag, __fadd_rn(__fmul_rn(omr, ny), __fmul_rn(r, my)));
      nw.z = __fmul_rn(mag, __fadd_rn(__fmul_rn(omr, nz), __fmul_rn(r, mz)));
      nw.w = __fmul_rn(mag, __fadd_rn(__fmul_rn(omr, nwv), __fmul_rn(r, mw)));
    }
    const float sq2 = fmaf(nw.x, nw.x, fmaf(nw.y, nw.y, fmaf(nw.z, nw.z, nw.w * nw.w)));
    const float ssk2 = block_sum(is_key ? sq2 : 0.f, red);
    if (is_key) {
      reinterpret_cast<float4*>(bank.keys + u * bank.d_key)[f] = nw;
      const float d2 = fmaxf(sqrtf(ssk2), 1e-12f);
      store_nk4(bank.nk, bank.nkh, bank.nkl, u * bank.d_key + 4 * f,
                make_float4(nw.x / d2, nw.y / d2, nw.z / d2, nw.w / d2));
      if (bank.kh) store_key_ops4(bank.kh, bank.kl, u * bank.d_key + 4 * f, nw);
    }
    if (is_val) {
      reinterpret_cast<float4*>(bank.values + u * bank.d_val)[f - dk4] = nw;
      if (bank.vh) store_val_ops4(bank.vh, bank.v8, bank.vl, u * bank.d_val + 4 * (f - dk4), nw);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// evict plan: single CTA threshold search (FeatureBank.py:121-138)
// ------------------------------------------------------------------------------------------------
constexpr int EV_THREADS = 1024;

__device__ __forceinline__ void block_min_count(float& mn, int& cnt, int& nanflag, float* sf, int* si) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    nanflag |= __shfl_xor_sync(0xffffffffu, nanflag, o);
  }
  __syncthreads();
  if (lane == 0) { sf[wid] = mn; si[wid] = cnt; si[32 + wid] = nanflag; }
  __syncthreads();
  mn = sf[lane]; cnt = si[lane]; nanflag = si[32 + lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    nanflag |= __shfl_xor_sync(0xffffffffu, nanflag, o);
  }
}

__global__ void __launch_bounds__(EV_THREADS) evict_plan_kernel(const float* __restrict__ info, int64_t n,
                                                                float frame_idx, double class_budget,
                                                                int64_t request_n, int32_t* __restrict__ plan,
                                                                int32_t* __restrict__ h_plan,
                                                                float* __restrict__ lfu) {
  __shared__ float sf[32];
  __shared__ int si[64];
  const int tid = threadIdx.x;
  float mn = INFINITY;
  int cnt = 0, nanflag = 0;
  for (int64_t i = tid; i < n; i += EV_THREADS) {
    const float2 in = reinterpret_cast<const float2*>(info)[i];
    const float age = __fsub_rn(frame_idx, in.x);     // frame_idx - info[:,0]        (:121)
    const float v = __fdiv_rn(in.y, age);             // info[:,1] / age              (:122)
    lfu[i] = v;
    if (v != v) nanflag = 1;
    mn = fminf(mn, v);
  }
  block_min_count(mn, cnt, nanflag, sf, si);
  int status = 0, T = 0, kept = 0, it = 0;
  if (nanflag || !isfinite(mn) || n == 0) {
    status = 2;                                       // int(nan)/int(inf)/min(empty) raise in the reference
  } else {
    T = (int)truncf(mn) + 1;                          // int(LFU.min()) + 1           (:123)
    // no iteration cap (the reference has none): T strictly increases and LFU <= 1e5, so the search terminates;
    // only the first 64 thresholds are kept for inspection, `it` counts all of them
    for (;; ++it) {
      float m2 = INFINITY;
      int c2 = 0, nf = 0;
      const float Tf = (float)T;
      for (int64_t i = tid; i < n; i += EV_THREADS) {
        const float v = lfu[i];
        if (v > Tf) { ++c2; m2 = fminf(m2, v); }      // strict >                     (:127)
      }
      block_min_count(m2, c2, nf, sf, si);
      kept = c2;
      if (tid == 0 && it < 64) plan[4 + it] = T;
      const double balance = (class_budget - (double)kept) - (double)request_n;   // (:134)
      if (balance < 0) {
        if (kept == 0) { status = 1; ++it; break; }   // LFU.min() of an empty tensor raises
        T = (int)truncf(m2) + 1;                      // (:136)
      } else {
        ++it;
        break;
      }
    }
  }
  if (tid == 0) {
    plan[0] = status; plan[1] = kept; plan[2] = it; plan[3] = T;
    if (h_plan) {
      h_plan[0] = status; h_plan[1] = kept; h_plan[2] = it; h_plan[3] = T;
      for (int k = 0; k < it && k < 64; ++k) h_plan[4 + k] = plan[4 + k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compaction (order preserving, out of place)
// ------------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256;

__global__ void __launch_bounds__(CP_THREADS) compact_count_kernel(const float* __restrict__ lfu, int64_t n,
                                                                   const int32_t* __restrict__ plan,
                                                                   int32_t* __restrict__ block_cnt) {
  const float Tf = (float)plan[3];
  const int64_t i = (int64_t)blockIdx.x * CP_THREADS + threadIdx.x;
  const int keep = (i < n) && (lfu[i] > Tf);
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_cnt[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) compact_scan_kernel(int32_t* __restrict__ block_cnt, int nb) {
  __shared__ int wt[33];
  int carry = 0;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < nb) ? block_cnt[i] : 0;
    int tot;
    const int ex = block_excl_scan(v, wt, &tot);
    if (i < nb) block_cnt[i] = carry + ex;
    carry += tot;
    __syncthreads();
  }
}

__device__ __forceinline__ void warp_copy16(void* dst, const void* src, int n16, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = lane; i < n16; i += 32) d[i] = __ldg(s + i);
}

__global__ void __launch_bounds__(CP_THREADS) compact_move_kernel(vfn_bank src, vfn_bank dst,
                                                                  const float* __restrict__ lfu,
                                                                  const int32_t* __restrict__ plan,
                                                                  const int32_t* __restrict__ block_off) {
  __shared__ int wt[33];
  __shared__ int list[CP_THREADS];
  const float Tf = (float)plan[3];
  const int64_t i = (int64_t)blockIdx.x * CP_THREADS + threadIdx.x;
  const int keep = (i < src.n) && (lfu[i] > Tf);
  int tot;
  const int pos = block_excl_scan(keep, wt, &tot);
  if (keep) list[pos] = threadIdx.x;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t base_dst = block_off[blockIdx.x];
  const int dk = src.d_key, dv = src.d_val;
  if (src.kh && dk == 128 && dv == 512) {
    // AFB-URR dims with tensor-core operand arrays: a warp reads the fp32 masters of a row (keys, nk, values: 3080 B,
    // all loads in flight before the first store) and re-derives the seven fp16/fp8 operand arrays with the same
    // element-wise conversions that wrote them (bit-identical), instead of copying another 3584 B per row.
    // Two rows per warp iteration keep 12 independent 16-byte loads per lane in flight.
    for (int e = 2 * wid; e < tot; e += 2 * (CP_THREADS / 32)) {
      const bool two = e + 1 < tot;
      const int64_t s0 = (int64_t)blockIdx.x * CP_THREADS + list[e];
      const int64_t s1 = two ? (int64_t)blockIdx.x * CP_THREADS + list[e + 1] : s0;
      const int64_t d0 = base_dst + e, d1 = d0 + 1;
      float4 k[2], nk[2], v[2][4];
      float2 inf[2];
      k[0] = __ldg(reinterpret_cast<const float4*>(src.keys + s0 * 128) + lane);
      k[1] = __ldg(reinterpret_cast<const float4*>(src.keys + s1 * 128) + lane);
      nk[0] = __ldg(reinterpret_cast<const float4*>(src.nk + s0 * 128) + lane);
      nk[1] = __ldg(reinterpret_cast<const float4*>(src.nk + s1 * 128) + lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[0][j] = __ldg(reinterpret_cast<const float4*>(src.values + s0 * 512) + lane + 32 * j);
        v[1][j] = __ldg(reinterpret_cast<const float4*>(src.values + s1 * 512) + lane + 32 * j);
      }
      if (lane < 2) inf[0] = __ldg(reinterpret_cast<const float2*>(src.info) + (lane ? s1 : s0));
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
        const int64_t d = r ? d1 : d0;
        reinterpret_cast<float4*>(dst.keys + d * 128)[lane] = k[r];
        store_key_ops4(dst.kh, dst.kl, d * 128 + 4 * lane, k[r]);
        store_nk4(dst.nk, dst.nkh, dst.nkl, d * 128 + 4 * lane, nk[r]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          reinterpret_cast<float4*>(dst.values + d * 512)[lane + 32 * j] = v[r][j];
          store_val_ops4(dst.vh, dst.v8, dst.vl, d * 512 + 4 * (lane + 32 * j), v[r][j]);
        }
      }
      if (lane == 0 || (lane == 1 && two)) {
        reinterpret_cast<float2*>(dst.info)[lane ? d1 : d0] = inf[0];
        dst.cnt[lane ? d1 : d0] = 0;
      }
    }
    return;
  }
  for (int e = wid; e < tot; e += CP_THREADS / 32) {
    const int64_t s = (int64_t)blockIdx.x * CP_THREADS + list[e];
    const int64_t d = base_dst + e;
    warp_copy16(dst.keys + d * dk, src.keys + s * dk, dk / 4, lane);
    warp_copy16(dst.values + d * dv, src.values + s * dv, dv / 4, lane);
    warp_copy16(dst.nk + d * dk, src.nk + s * dk, dk / 4, lane);
    if (src.kh) {
      warp_copy16(dst.nkh + d * dk, src.nkh + s * dk, dk / 8, lane);
      warp_copy16(dst.nkl + d * dk, src.nkl + s * dk, dk / 8, lane);
      warp_copy16(dst.kh + d * dk, src.kh + s * dk, dk / 8, lane);
      warp_copy16(dst.kl + d * dk, src.kl + s * dk, dk / 8, lane);
      warp_copy16(dst.vh + d * dv, src.vh + s * dv, dv / 8, lane);
      warp_copy16(dst.v8 + d * dv, src.v8 + s * dv, dv / 16, lane);
      warp_copy16(dst.vl + d * dv, src.vl + s * dv, dv / 16, lane);
    }
    if (lane == 0) {
      reinterpret_cast<float2*>(dst.info)[d] = reinterpret_cast<const float2*>(src.info)[s];
      dst.cnt[d] = 0;
    }
  }
}

struct ClampSet { float* info[4]; int64_t n[4]; int32_t* n_live[4]; int64_t commit[4]; };
// grid (rows / 256, objects).  Last kernel of an update: also commits the bank's device-resident live count (nothing in
// this kernel reads it; rows are bounded by the host-side upper bound).
__global__ void clamp_info_kernel(const __grid_constant__ ClampSet set) {
  pdl_wait();
  pdl_trigger();
  float* __restrict__ info = set.info[blockIdx.y];
  const int64_t n = set.n[blockIdx.y];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && set.n_live[blockIdx.y]) {
    int32_t* nl = set.n_live[blockIdx.y];
    const int32_t v = set.commit[blockIdx.y] >= 0 ? (int32_t)set.commit[blockIdx.y] : nl[1];
    nl[0] = v;
    nl[1] = v;
  }
  if (i < n) {
    float v = info[2 * i + 1];
    v = v < 0.f ? 0.f : (v > 1e5f ? 1e5f : v);   // torch.clamp(.,0,1e5), NaN preserved (FeatureBank.py:115)
    info[2 * i + 1] = v;
  }
}

__global__ void set_live_kernel(int32_t* n_live, int32_t n) { n_live[0] = n; n_live[1] = n; }

static int check_bank(const vfn_bank* b) {
  VFN_CHECK_ARG(b != nullptr, "bank is NULL");
  VFN_CHECK_ARG(b->d_key > 0 && b->d_val > 0 && b->d_key % 8 == 0 && b->d_val % 8 == 0,
                "d_key/d_val must be positive multiples of 8 (got %d, %d)", b->d_key, b->d_val);
  VFN_CHECK_ARG(b->keys && b->values && b->info && b->nk && b->cnt, "bank has NULL arrays");
  VFN_CHECK_ARG(b->n >= 0 && b->n <= b->cap, "bank n=%lld exceeds cap=%lld", (long long)b->n, (long long)b->cap);
  VFN_CHECK_ARG((b->kh == nullptr) == (b->kl == nullptr) && (b->kh == nullptr) == (b->vh == nullptr) &&
                    (b->kh == nullptr) == (b->vl == nullptr) && (b->kh == nullptr) == (b->v8 == nullptr) &&
                    (b->kh == nullptr) == (b->nkh == nullptr) && (b->kh == nullptr) == (b->nkl == nullptr),
                "tensor-core operand arrays must be all set or all NULL");
  return VFN_OK;
}

int launch_plan(const UpdObj* o, int n_obj, int64_t hw, float thres_close, cudaStream_t st) {
  VFN_CHECK_ARG(n_obj >= 1 && n_obj <= 4, "plan: 1..4 objects per launch");
  VFN_CHECK_ARG(hw > 0 && hw <= 65536, "plan: hw=%lld out of range (1..65536)", (long long)hw);
  PlanSet set;
  for (int c = 0; c < n_obj; ++c) {
    VFN_CHECK_ARG(o[c].match_idx && o[c].match_corr && o[c].merge_q && o[c].merge_slot && o[c].run_off && o[c].append_q &&
                      o[c].counts && o[c].plan_ws, "plan: NULL argument");
    set.o[c] = PlanObj{o[c].match_idx, o[c].match_corr, o[c].merge_q, o[c].merge_slot, o[c].run_off, o[c].append_q,
                       o[c].counts, o[c].h_counts, reinterpret_cast<unsigned long long*>(o[c].plan_ws), o[c].n_live};
  }
  VFN_CUDA_OK(launch_pdl(plan_kernel, dim3(n_obj), dim3(PLAN_THREADS), 0, st, set, (int)hw, thres_close));
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int launch_merge(const UpdObj* o, int n_obj, int64_t hw, float update_rate, cudaStream_t st) {
  VFN_CHECK_ARG(n_obj >= 1 && n_obj <= 4, "merge: 1..4 objects per launch");
  MergeSet set;
  for (int c = 0; c < n_obj; ++c) {
    if (int rc = check_bank(&o[c].bank)) return rc;
    VFN_CHECK_ARG(o[c].nck && o[c].ncv && o[c].merge_q && o[c].merge_slot && o[c].run_off && o[c].counts, "merge: bad args");
    if ((o[c].bank.d_key + o[c].bank.d_val) / 4 > MERGE_THREADS) {
      set_error("merge: d_key + d_val = %d exceeds %d", o[c].bank.d_key + o[c].bank.d_val, MERGE_THREADS * 4);
      return VFN_E_UNSUPPORTED;
    }
    set.o[c] = MergeObj{o[c].bank, o[c].nck, o[c].ncv, o[c].merge_q, o[c].merge_slot, o[c].run_off, o[c].counts};
  }
  // (1 - update_rate) is evaluated in double by Python and rounded to fp32 when it meets the tensor
  const float omr = (float)(1.0 - (double)update_rate);
  dim3 grid((unsigned)(hw < 148 * 8 ? hw : 148 * 8), n_obj);
  prof_begin(PROF_MERGE, st);
  VFN_CUDA_OK(launch_pdl(merge_runs_kernel, grid, dim3(MERGE_THREADS), 0, st, set, omr, update_rate));
  prof_end(PROF_MERGE, st, 0.0);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int launch_append(const UpdObj* o, int n_obj, float info0, float info1, cudaStream_t st) {
  VFN_CHECK_ARG(n_obj >= 1 && n_obj <= 4, "append: 1..4 objects per launch");
  AppendSet set;
  int64_t n_max = 0;
  double bytes = 0;
  for (int c = 0; c < n_obj; ++c) {
    if (int rc = check_bank(&o[c].bank)) return rc;
    VFN_CHECK_ARG(o[c].ck && o[c].cv && o[c].n_sel >= 0, "append_rows: bad args");
    if (o[c].bank.n + o[c].n_sel > o[c].bank.cap) {
      set_error("append_rows: n=%lld + %lld exceeds cap=%lld", (long long)o[c].bank.n, (long long)o[c].n_sel,
                (long long)o[c].bank.cap);
      return VFN_E_CAPACITY;
    }
    set.o[c] = AppendObj{o[c].bank, o[c].ck, o[c].cv, o[c].nck, o[c].sel, o[c].n_sel_dev, o[c].n_sel};
    if (o[c].n_sel > n_max) n_max = o[c].n_sel;
    if (!o[c].n_sel_dev) bytes += 2.0 * 4.0 * (o[c].bank.d_key + o[c].bank.d_val + 2) * (double)o[c].n_sel;
  }
  if (n_max == 0) return VFN_OK;
  dim3 grid((unsigned)(n_max < 148 * 16 ? n_max : 148 * 16), n_obj);
  bool warp_rows = true;        // every object: 128/512 dims, operand arrays, normalised candidates given
  for (int c = 0; c < n_obj; ++c)
    warp_rows = warp_rows && o[c].bank.d_key == 128 && o[c].bank.d_val == 512 && o[c].bank.kh && o[c].bank.nkh && o[c].nck;
  prof_begin(PROF_APPEND, st);
  if (warp_rows) {
    const int64_t blocks = cdiv(n_max, 4);
    dim3 gw((unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), n_obj);
    VFN_CUDA_OK(launch_pdl(append_rows_warp_kernel, gw, dim3(128), 0, st, set, info0, info1));
  } else {
    VFN_CUDA_OK(launch_pdl(append_rows_kernel, grid, dim3(128), 0, st, set, info0, info1));
  }
  // algorithmic bytes: read + write of the appended rows (keys, values, info)
  prof_end(PROF_APPEND, st, bytes);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int launch_clamp(const vfn_bank* banks, int n_obj, cudaStream_t st, const int64_t* commit) {
  VFN_CHECK_ARG(n_obj >= 1 && n_obj <= 4, "clamp: 1..4 objects per launch");
  ClampSet set;
  int64_t n_max = 0;
  bool any_live = false;
  for (int c = 0; c < n_obj; ++c) {
    set.info[c] = banks[c].info;
    set.n[c] = banks[c].n;
    set.n_live[c] = banks[c].n_live;
    set.commit[c] = commit ? commit[c] : -1;
    any_live = any_live || banks[c].n_live != nullptr;
    if (banks[c].n > n_max) n_max = banks[c].n;
  }
  if (n_max == 0 && any_live) n_max = 1;          // the commit must still happen
  if (n_max == 0) return VFN_OK;
  dim3 grid((unsigned)cdiv(n_max, 256), n_obj);
  VFN_CUDA_OK(launch_pdl(clamp_info_kernel, grid, dim3(256), 0, st, set));
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

int vfn_version(void) { return VFN_VERSION; }
const char* vfn_last_error(void) { return g_err; }

int vfn_bank_set_live(const vfn_bank* bank, int64_t n, void* stream) {
  VFN_CHECK_ARG(bank != nullptr && n >= 0 && n <= bank->cap && n < (1ll << 31), "set_live: bad args");
  if (!bank->n_live) return VFN_OK;
  set_live_kernel<<<1, 1, 0, as_stream(stream)>>>(bank->n_live, (int32_t)n);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { g_ev_pool.push_back(r.a); g_ev_pool.push_back(r.b); }
  g_prof.clear();
  if (on) {   // pre-create events so that recording inside a timed region never calls cudaEventCreate
    while (g_ev_pool.size() < 16384) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      g_ev_pool.push_back(e);
    }
    g_prof.reserve(8192);
  }
  g_prof_on = on != 0;
  return VFN_OK;
}

int vfn_profile_collect(double* h_out, int32_t n_kinds) {
  VFN_CHECK_ARG(h_out && n_kinds >= 1 && n_kinds <= PROF_KINDS, "profile_collect: bad args");
  for (int k = 0; k < n_kinds * 3; ++k) h_out[k] = 0.0;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    VFN_CUDA_OK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    VFN_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.kind < n_kinds) { h_out[3 * r.kind] += 1.0; h_out[3 * r.kind + 1] += ms; h_out[3 * r.kind + 2] += r.work; }
  }
  return VFN_OK;
}

int64_t vfn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int vfn_profile_add_work(int32_t kind, double work) {
  VFN_CHECK_ARG(kind >= 0 && kind < PROF_KINDS, "profile_add_work: bad kind");
  if (!g_prof_on) return VFN_OK;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto it = g_prof.rbegin(); it != g_prof.rend(); ++it)
    if (it->kind == kind) { it->work += work; break; }
  return VFN_OK;
}

int vfn_abi_sizeof_bank(void) { return (int)sizeof(vfn_bank); }
int vfn_abi_sizeof_update_io(void) { return (int)sizeof(vfn_update_io); }

int vfn_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int vfn_prep_rows(const float* d_src_dm, int32_t d, int64_t n, float* d_raw_em, float* d_normed_em, uint16_t* d_hi_em,
                  uint16_t* d_lo_em, float scale, void* stream) {
  VFN_CHECK_ARG(d_src_dm && d > 0 && n >= 0, "prep_rows: bad args");
  if (n == 0) return VFN_OK;
  PrepJob jb{d_src_dm, d, n, d_raw_em, d_normed_em, d_hi_em, d_lo_em, scale, d_normed_em ? 1 : 0};
  return launch_prep(&jb, 1, as_stream(stream));
}

int vfn_bank_append_rows(const vfn_bank* bank, const float* d_ck_em, const float* d_cv_em, const float* d_nck_em,
                         const int32_t* d_sel, int64_t n_sel_upper, const int32_t* d_n_sel, float info0, float info1,
                         void* stream) {
  VFN_CHECK_ARG(bank != nullptr, "bank is NULL");
  UpdObj o{};
  o.bank = *bank; o.ck = d_ck_em; o.cv = d_cv_em; o.nck = d_nck_em; o.sel = d_sel; o.n_sel_dev = d_n_sel; o.n_sel = n_sel_upper;
  return launch_append(&o, 1, info0, info1, as_stream(stream));
}

int vfn_bank_refresh(const vfn_bank* bank, int64_t first, int64_t count, void* stream) {
  if (int rc = check_bank(bank)) return rc;
  VFN_CHECK_ARG(first >= 0 && count >= 0 && first + count <= bank->cap, "refresh: range out of bounds");
  if (count == 0) return VFN_OK;
  const unsigned grid = (unsigned)(count < 148 * 16 ? count : 148 * 16);
  refresh_rows_kernel<<<grid, 128, 0, as_stream(stream)>>>(*bank, first, count);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

size_t vfn_bank_plan_workspace_bytes(int64_t hw) {
  int64_t npad = 1;
  while (npad < hw) npad <<= 1;
  return (size_t)npad * sizeof(unsigned long long);
}

int vfn_bank_plan(const int32_t* d_match_idx, const float* d_match_corr, int64_t hw, float thres_close,
                  int32_t* d_merge_q, int32_t* d_merge_slot, int32_t* d_run_off, int32_t* d_append_q, int32_t* d_counts,
                  int32_t* h_counts, void* d_ws, size_t ws_bytes, void* stream) {
  if (ws_bytes < vfn_bank_plan_workspace_bytes(hw) || !d_ws) {
    set_error("plan: workspace too small");
    return VFN_E_CAPACITY;
  }
  UpdObj o{};
  o.match_idx = d_match_idx; o.match_corr = d_match_corr; o.merge_q = d_merge_q; o.merge_slot = d_merge_slot;
  o.run_off = d_run_off; o.append_q = d_append_q; o.counts = d_counts; o.h_counts = h_counts; o.plan_ws = d_ws;
  return launch_plan(&o, 1, hw, thres_close, as_stream(stream));
}

int vfn_bank_merge(const vfn_bank* bank, const float* d_nck_em, const float* d_ncv_em, const int32_t* d_merge_q,
                   const int32_t* d_merge_slot, const int32_t* d_run_off, const int32_t* d_counts, int64_t hw,
                   float update_rate, void* stream) {
  VFN_CHECK_ARG(bank != nullptr && hw > 0, "merge: bad args");
  UpdObj o{};
  o.bank = *bank; o.nck = d_nck_em; o.ncv = d_ncv_em; o.merge_q = const_cast<int32_t*>(d_merge_q);
  o.merge_slot = const_cast<int32_t*>(d_merge_slot); o.run_off = const_cast<int32_t*>(d_run_off);
  o.counts = const_cast<int32_t*>(d_counts);
  return launch_merge(&o, 1, hw, update_rate, as_stream(stream));
}

int vfn_bank_evict_plan(const vfn_bank* bank, float frame_idx, double class_budget, int64_t request_n, int32_t* d_plan,
                        int32_t* h_plan, float* d_lfu_scratch, void* stream) {
  if (int rc = check_bank(bank)) return rc;
  VFN_CHECK_ARG(d_plan && d_lfu_scratch, "evict_plan: NULL argument");
  evict_plan_kernel<<<1, EV_THREADS, 0, as_stream(stream)>>>(bank->info, bank->n, frame_idx, class_budget, request_n,
                                                             d_plan, h_plan, d_lfu_scratch);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

size_t vfn_bank_compact_workspace_bytes(int64_t n) { return (size_t)(cdiv(n, CP_THREADS) + 1) * sizeof(int32_t); }

int vfn_bank_compact(const vfn_bank* src, const vfn_bank* dst, const float* d_lfu, const int32_t* d_plan, void* d_ws,
                     size_t ws_bytes, void* stream) {
  if (int rc = check_bank(src)) return rc;
  if (int rc = check_bank(dst)) return rc;
  VFN_CHECK_ARG(src->d_key == dst->d_key && src->d_val == dst->d_val, "compact: dim mismatch");
  VFN_CHECK_ARG((src->kh == nullptr) == (dst->kh == nullptr), "compact: operand arrays mismatch");
  VFN_CHECK_ARG(d_lfu && d_plan && d_ws, "compact: NULL argument");
  if (ws_bytes < vfn_bank_compact_workspace_bytes(src->n)) {
    set_error("compact: workspace too small");
    return VFN_E_CAPACITY;
  }
  if (src->n == 0) return VFN_OK;
  const int nb = (int)cdiv(src->n, CP_THREADS);
  int32_t* block_cnt = reinterpret_cast<int32_t*>(d_ws);
  cudaStream_t st = as_stream(stream);
  compact_count_kernel<<<nb, CP_THREADS, 0, st>>>(d_lfu, src->n, d_plan, block_cnt);
  compact_scan_kernel<<<1, 1024, 0, st>>>(block_cnt, nb);
  prof_begin(PROF_COMPACT, st);
  compact_move_kernel<<<nb, CP_THREADS, 0, st>>>(*src, *dst, d_lfu, d_plan, block_cnt);
  // work is filled in by the host (kept count is only known after the plan is read back): see vfn_profile_add_work
  prof_end(PROF_COMPACT, st, 0.0);
  VFN_LAUNCH_OK();
  count_launches(3);
  return VFN_OK;
}

int vfn_bank_clamp_info(const vfn_bank* bank, int64_t n, void* stream) {
  if (int rc = check_bank(bank)) return rc;
  VFN_CHECK_ARG(n >= 0 && n <= bank->cap, "clamp_info: n out of range");
  vfn_bank b = *bank;
  b.n = n;
  return launch_clamp(&b, 1, as_stream(stream));
}

}  // extern "C"
