// tcgen05 / TMEM / TMA implementation of the memory read (Matcher.forward, AFB_URR.py:136-178) and of the cosine
// match (FeatureBank.py:63-68) for sm_100a.
//
//   phase A  S^T[j,i] = <q_j, k_i> * log2(e)/sqrt(128)   -> per-query running (max, sum 2^(s-max)) over the slots
//   phase B  S^T recomputed, P = 2^(S - lse2_j) (already normalised: no online rescale), usage counts
//            cnt_i += [P_ij > thres], O^T[j,c] += sum_i P_ij V_ic with the fp32 accumulator resident in TMEM.
//   match    corr[j,i] = <nck_j, nk_i> -> per-query candidate slots within a band of the maximum -> exact fp32 re-score.
//
// Orientation (all kernels): TMEM lanes = queries / candidates (M = 128), columns = bank slots (S) / value channels (O).
// A operands come from TMEM (Q for the S-MMA, P for the O-MMA), B operands from shared memory via TMA (128B swizzle):
//   K tiles  K-major  [slots x 128 d]   fp16 hi, fp16 lo
//   V tiles  MN-major [slots x 256 ch]  fp16 hi, e4m3(value), e5m2(value - hi)
// Precision (scripts/precision_study.py):
//   affinity / match scores: fp16 hi/lo splits, hi*hi + lo*hi + hi*lo (kind::f16, 3 passes, ~2^-21 relative):
//     logits and LSE are fp32-grade, so the usage-count threshold P > 1e-3 decides like the reference;
//   readout: P' = 256*P = hi(fp16) + lo, O' = hi*Vhi (kind::f16) + e4m3(lo)*e4m3(V) + e4m3(P')*e5m2(V - Vhi)
//     (kind::f8f6f4, double rate): 2 bf16-equivalent passes instead of 3, readout error ~1e-4 (tolerance 1e-3).
// Work split: persistent CTAs (one per SM), work items = (slot split x object x query tile[, channel half]) dealt
// round-robin; per-item partials are combined in a fixed order (vfn_simt.cu), deterministic.
#include "vfn_tc.cuh"
#include "vfn_ptx.cuh"

namespace vfn {


constexpr int DK = 128, DV = 512;
constexpr int QT = 128;                 // queries per tile (TMEM lanes)
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..19 sixteen epilogue / softmax warps
// (warp w owns TMEM lanes 32*(w%4)..+32; the four warps sharing a lane quarter split the columns)
constexpr int TC_THREADS = 640;
constexpr int EPI_WARPS = 16, EPI_THREADS = EPI_WARPS * 32;
constexpr int TC_MAX_OBJ = 4;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

struct TcMaps {
  CUtensorMap kh[TC_MAX_OBJ], kl[TC_MAX_OBJ], vh[TC_MAX_OBJ], v8[TC_MAX_OBJ], vl[TC_MAX_OBJ];
};
constexpr int TC_MAX_SPLIT = 24;
struct TcArgs {
  int obj_n, hw, q_tiles, pieces;       // pieces = partial slots per combo
  int n[TC_MAX_OBJ];                    // live slots per object (an upper bound when n_live[obj] is set)
  int tiles[TC_MAX_OBJ];                // slot tiles per object for this phase (same)
  const int32_t* n_live[TC_MAX_OBJ];    // device-resident live count (vfn_bank::n_live) or NULL
  int32_t* pieces_dev;                  // with live counts: `pieces` is chosen on the device (tc_pieces) from the plan_*
                                        // parameters and published here for the combine kernels that follow; else NULL
  int plan_tile, plan_combos, plan_overhead, plan_G, plan_chain_max;
  const uint16_t* qh;                   // (rows, 128) fp16 hi of the A operand (q * log2e/sqrt(d), or 16 * normalised candidate)
  const uint16_t* ql;
  long long a_obj_stride;               // elements between objects in qh/ql (0: one query set for all objects)
  int32_t* cnt[TC_MAX_OBJ];
  float band;                           // match: candidate band in the (scaled) score domain
  float* dbg;
  long long* tstamp;                    // vfn_debug_set_tstamp: per cluster and item {start, first O issued, last O issued, end, tiles}
};

// Work split of a launch.  Host-tracked sizes: args.pieces.  Device-resident live counts: every warp of every CTA
// evaluates the host's best_split() on the live counts (lane s-1 takes split s; integer costs, ties -> smallest s, the
// first minimum of the host loop), so a launch issued against bounds partitions the work exactly like one issued after
// reading the sizes back; CTA 0 publishes the choice for the combine kernels.  Must be called by full warps, after
// pdl_wait().  (A separate 1-warp kernel did this before: three extra launches per frame.)
__device__ __forceinline__ int tc_pieces(const TcArgs& a) {
  if (!a.pieces_dev) return a.pieces;
  const int lane = threadIdx.x & 31;
  long long tmin = 0x7fffffffffffffffll, tmax = 0;
  for (int o = 0; o < a.obj_n; ++o) {
    const long long n = *reinterpret_cast<const volatile int32_t*>(a.n_live[o]);
    const long long t = (n + a.plan_tile - 1) / a.plan_tile;
    tmin = t < tmin ? t : tmin;
    tmax = t > tmax ? t : tmax;
  }
  int s_min = 1;
  if (a.plan_chain_max > 0) s_min = (int)((tmax * a.plan_tile + a.plan_chain_max - 1) / a.plan_chain_max);
  if (s_min < 1) s_min = 1;
  if (s_min > TC_MAX_SPLIT) s_min = TC_MAX_SPLIT;
  if (s_min > tmin) s_min = (int)(tmin > 1 ? tmin : 1);
  const int sp = lane + 1;
  long long cost = 0x7fffffffffffffffll;
  if (sp >= s_min && sp <= TC_MAX_SPLIT && sp <= tmin) {
    const long long rounds = ((long long)a.plan_combos * sp + a.plan_G - 1) / a.plan_G;
    cost = rounds * ((tmax + sp - 1) / sp + a.plan_overhead);
  }
  int best = sp;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long oc = __shfl_xor_sync(0xffffffffu, cost, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (oc < cost || (oc == cost && ob < best)) { cost = oc; best = ob; }
  }
  const int pieces = (cost == 0x7fffffffffffffffll) ? s_min : best;
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.pieces_dev = pieces;
  return pieces;
}

// A operand (128 rows x 128 d, fp16 hi and lo) -> TMEM columns [col_h, col_h+64) and [col_l, col_l+64).
// Epilogue warp (quarter q, column group cg): cg 0,1 -> hi halves, cg 2,3 -> lo halves; 32 columns (64 fp16) each.
__device__ __forceinline__ void load_a_operand(const TcArgs& args, int obj, int qt, uint32_t tmem, uint32_t col_h,
                                               uint32_t col_l, int warp, int lane) {
  const int quarter = warp & 3, cg = (warp - 4) >> 2;
  const int row = (quarter << 5) + lane;
  const uint32_t taddr = tmem + (((uint32_t)quarter * 32u) << 16) + (cg < 2 ? col_h : col_l) + (uint32_t)(cg & 1) * 32u;
  uint32_t v[32];
  {
    const uint16_t* base = (cg < 2 ? args.qh : args.ql) + (size_t)obj * args.a_obj_stride;
    const uint4* src = reinterpret_cast<const uint4*>(base + ((size_t)qt * QT + row) * DK + (cg & 1) * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 x = src[i];
      v[4 * i + 0] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
  }
  tmem_st32(taddr, v);
  tmem_wait_st();
}

// ------------------------------------------------------------------------------------------------
// score scan: phase A of the read (MODE_LSE) and the cosine match (MODE_MATCH).
//   TMEM: A hi [0,64) | A lo [64,128) | three S buffers of 128 columns at 128 + 128 b
//   smem: three 64 KB stages, each one 128-slot tile of the B operand: [hi: 2 boxes of 64 d | lo: 2 boxes]
// Every epilogue warp processes its 32 columns of EVERY tile; a buffer is released as soon as its values are in
// registers, so the MMA stream runs up to two tiles ahead of the epilogue.
// ------------------------------------------------------------------------------------------------
constexpr int SC_TILE = 128;
constexpr int SC_STAGES = 3, SC_BUFS = 3;
constexpr int SC_STAGE_BYTES = SC_TILE * DK * 2 * 2;   // hi + lo = 64 KB
constexpr int SC_AUX_BYTES = EPI_THREADS * 4 * 8;      // match: 4-entry candidate ring per epilogue thread; LSE: (m,l) exchange
constexpr int SC_SMEM = SC_STAGES * SC_STAGE_BYTES + 1024 + 256 + SC_AUX_BYTES;
constexpr uint32_t TS_AH = 0, TS_AL = 64, TS_S = 128;
constexpr int MODE_LSE = 0, MODE_MATCH = 1;
constexpr int MATCH_RING = 4;           // near-tie candidates kept per (item, query, column group)
constexpr int MATCH_CAND = 4 * MATCH_RING;   // candidate entries per (item, query)

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_scan_kernel(const __grid_constant__ TcMaps maps, TcArgs args,
                                                                float2* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SC_STAGES * SC_STAGE_BYTES);
  uint64_t* k_full = bars;                   // [3]
  uint64_t* k_empty = bars + 3;              // [3]
  uint64_t* s_full = bars + 6;               // [3]
  uint64_t* s_empty = bars + 9;              // [3]
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + 12);
  float2* aux = reinterpret_cast<float2*>(smem + SC_STAGES * SC_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_p, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;

  // work items = (split, object, query tile), dealt round-robin to the persistent CTAs: in any round all CTAs
  // stream the same few slot ranges, so the tiles are served from L2 (profiles/r1a_summary.md).
  const int n_combos = args.obj_n * args.q_tiles;
  const int pieces = tc_pieces(args);
  const int n_items = n_combos * pieces;
  uint32_t tile_ctr = 0;        // tiles streamed so far by this CTA (all roles agree)
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int piece = item / n_combos;
    const int combo = item - piece * n_combos;
    const int obj = combo / args.q_tiles;
    const int qt = combo - obj * args.q_tiles;
    const int n_obj = args.n_live[obj] ? *reinterpret_cast<const volatile int32_t*>(args.n_live[obj]) : args.n[obj];
    const int tiles_o = args.n_live[obj] ? (n_obj + SC_TILE - 1) / SC_TILE : args.tiles[obj];
    const int t0 = (int)((long long)tiles_o * piece / pieces);
    const int t1 = (int)((long long)tiles_o * (piece + 1) / pieces);
    const int ntile = t1 - t0;
    const bool first_item = (item == (int)blockIdx.x);

    if (warp >= 4) load_a_operand(args, obj, qt, tmem, TS_AH, TS_AL, warp, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      if (lane == 0) {
        for (int t = 0; t < ntile; ++t) {
          const uint32_t c = tile_ctr + t, st = c % SC_STAGES, ph = (c / SC_STAGES) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          // the match reads only the hi half of the bank operand (two passes, see SCAN_PASSES)
          mbar_arrive_expect_tx(&k_full[st], MODE == MODE_MATCH ? SC_STAGE_BYTES / 2 : SC_STAGE_BYTES);
          uint8_t* dst = kst + st * SC_STAGE_BYTES;
          const int row0 = (t0 + t) * SC_TILE;
          tma_load_2d(dst, &maps.kh[obj], &k_full[st], 0, row0);
          tma_load_2d(dst + 16384, &maps.kh[obj], &k_full[st], 64, row0);
          if (MODE != MODE_MATCH) {
            tma_load_2d(dst + 32768, &maps.kl[obj], &k_full[st], 0, row0);
            tma_load_2d(dst + 49152, &maps.kl[obj], &k_full[st], 64, row0);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // the whole warp runs the loop (uniform control flow and operands); one elected lane issues
      constexpr uint32_t idesc = make_idesc(128, SC_TILE, FMT_F16, FMT_F16, 0, 0);
      for (int t = 0; t < ntile; ++t) {
        const uint32_t c = tile_ctr + t, st = c % SC_STAGES, ph = (c / SC_STAGES) & 1;
        mbar_wait(&s_empty[st], ph ^ 1);
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint32_t kbase = smem_u32(kst + st * SC_STAGE_BYTES);
        const uint32_t d_t = tmem + TS_S + st * SC_TILE;
        if (elect_one()) {
          // passes: (Ah,Bh) (Al,Bh) (Ah,Bl)
#pragma unroll
          for (int pass = 0; pass < (MODE == MODE_MATCH ? 2 : 3); ++pass) {
            const uint32_t a_col = (pass == 1) ? TS_AL : TS_AH;
            const uint32_t kb = kbase + ((pass == 2) ? 32768u : 0u);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t bd = make_sdesc(kb + (ks >> 2) * 16384u + (ks & 3) * 32u, 16, 1024);
              mma_ts(d_t, tmem + a_col + ks * 8, bd, idesc, (pass | ks) ? 1u : 0u);
            }
          }
          tc_commit(&k_empty[st]);
          tc_commit(&s_full[st]);
        }
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int quarter = warp & 3, cg = (warp - 4) >> 2;
      const int row = (quarter << 5) + lane;
      const int et = threadIdx.x - 128;                       // 0..511
      const uint32_t tlane = tmem + (((uint32_t)quarter * 32u) << 16) + TS_S + (uint32_t)cg * 32u;
      float m_run = -INFINITY, l_run = 0.f;
      int cnt = 0;
      float2* ring = aux + et * MATCH_RING;
      const bool row_live = qt * QT + row < args.hw;
      for (int t = 0; t < ntile; ++t) {
        const uint32_t c = tile_ctr + t, st = c % SC_BUFS, ph = (c / SC_BUFS) & 1;
        mbar_wait(&s_full[st], ph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tlane + st * SC_TILE, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);             // values are in registers: the buffer is free
        if (args.dbg && blockIdx.x == 0 && first_item && t == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) args.dbg[row * SC_TILE + cg * 32 + i] = __uint_as_float(v[i]);
        }
        const int slot0 = (t0 + t) * SC_TILE + cg * 32;
        const int lim = n_obj - slot0;                        // valid slots in this thread's 32 columns
        if (lim <= 0) continue;
        if (lim < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= lim) v[i] = __float_as_uint(-INFINITY);
        }
        float cm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) cm = fmax3(cm, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        if (MODE == MODE_LSE) {
          const float m_new = fmaxf(m_run, cm);
          if (m_new > -INFINITY) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              a0 += ex2(__uint_as_float(v[i]) - m_new);
              a1 += ex2(__uint_as_float(v[i + 1]) - m_new);
            }
            l_run = l_run * ex2(m_run - m_new) + (a0 + a1);
            m_run = m_new;
          }
        } else {
          // candidates = every slot whose score is within `band` of the running maximum at its time; a jump of the
          // maximum by more than the band invalidates everything before it
          // (rows beyond hw are zero A rows: every slot ties at score 0 and would take the slow path for nothing)
          if (cm > m_run + args.band) cnt = 0;
          const float m_new = fmaxf(m_run, cm);
          const float thr = row_live ? m_new - args.band : INFINITY;
          // once the running maximum is established almost no tile holds a score inside the band: the 32 compares
          // below only run for the lanes' tiles whose own maximum reaches it
          if (cm >= thr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float sv = __uint_as_float(v[i]);
              if (sv >= thr) {
                ring[cnt & (MATCH_RING - 1)] = make_float2(sv, __int_as_float(slot0 + i));
                ++cnt;
              }
            }
          }
          m_run = m_new;
        }
      }
      const int j = qt * QT + row;
      if (MODE == MODE_LSE) {
        // combine the four column groups' statistics and publish the piece
        aux[cg * QT + row] = make_float2(m_run, l_run);
        named_bar_sync(1, EPI_THREADS);
        if (cg == 0) {
          float m = m_run;
#pragma unroll
          for (int g = 1; g < 4; ++g) m = fmaxf(m, aux[g * QT + row].x);
          float l = 0.f;
          if (m > -INFINITY) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float2 o = aux[g * QT + row];
              if (o.x > -INFINITY) l += o.y * ex2(o.x - m);
            }
          }
          if (j < args.hw) part[((size_t)obj * pieces + piece) * args.hw + j] = make_float2(m * LN2, l);
        }
        named_bar_sync(1, EPI_THREADS);                      // aux is reused by the next item
      } else {
        if (j < args.hw) {
          float2 e[MATCH_RING];
          const float thr = m_run - args.band;
#pragma unroll
          for (int r = 0; r < MATCH_RING; ++r) {
            const float2 x = ring[r];
            e[r] = (r < cnt && x.x >= thr) ? x : make_float2(-INFINITY, __int_as_float(0x7fffffff));
          }
          if (cnt > MATCH_RING) e[0] = make_float2(m_run, __int_as_float(-1));     // overflow: exact scan needed
          float4* dst = reinterpret_cast<float4*>(
              part + (((size_t)obj * pieces + piece) * args.hw + j) * MATCH_CAND + cg * MATCH_RING);
          dst[0] = make_float4(e[0].x, e[0].y, e[1].x, e[1].y);
          dst[1] = make_float4(e[2].x, e[2].y, e[3].x, e[3].y);
        }
      }
    }
    tile_ctr += ntile;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// score scan on CTA pairs (cta_group::2, M = 256): the two CTAs of a cluster take adjacent query tiles of the same
// (object, slot split) and SHARE every 128-slot B tile - each CTA TMA-loads 64 of its slots, the pair's tensor cores
// read both halves.  The single-CTA kernel pulls 64 KB per 1536-clk tile from L2 (43 B/clk/SM x 148 SMs = the whole
// L2 slice throughput, B300_MICROARCH.md "LTS throughput cap"); the pair halves that and leaves room for six stages.
// Everything per CTA (TMEM layout, epilogue) is as in tc_scan_kernel; only the leader issues MMAs, its "full" barriers
// count both CTAs' bytes, commits are multicast, the peer's epilogue warps free S buffers with remote arrives.
// ------------------------------------------------------------------------------------------------
constexpr int SP_STAGES = 6;
constexpr int SP_STAGE_BYTES = 64 * DK * 2 * 2;        // this CTA's 64 slots, hi + lo: 32 KB
constexpr int SP_SMEM = SP_STAGES * SP_STAGE_BYTES + 1024 + 256 + SC_AUX_BYTES;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    tc_scan_pair_kernel(const __grid_constant__ TcMaps maps, TcArgs args, float2* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_STAGES * SP_STAGE_BYTES);
  uint64_t* k_full = bars;                       // [SP_STAGES]  (used on the leader: bytes of both CTAs' loads)
  uint64_t* k_empty = bars + SP_STAGES;          // [SP_STAGES]
  uint64_t* s_full = bars + 2 * SP_STAGES;       // [3]
  uint64_t* s_empty = bars + 2 * SP_STAGES + 3;  // [3]  (used on the leader: 16 local + 16 remote warps)
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + 2 * SP_STAGES + 6);
  float2* aux = reinterpret_cast<float2*>(smem + SP_STAGES * SP_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SP_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_base_p, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;
  // programmatic dependent launch: barriers, TMEM and the cluster handshake above overlap the predecessor's tail;
  // nothing before this point reads global memory
  pdl_wait();
  pdl_trigger();

  // work items = (split, object, query-tile pair), dealt round-robin to the persistent clusters
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int qpairs = (args.q_tiles + 1) >> 1;
  const int n_combos = args.obj_n * qpairs;
  const int pieces = tc_pieces(args);
  const int n_items = n_combos * pieces;
  uint32_t tile_ctr = 0;        // tiles streamed so far by this cluster (all roles agree)
  for (int item = cluster_id; item < n_items; item += n_clusters) {
    const int piece = item / n_combos;
    const int combo = item - piece * n_combos;
    const int obj = combo / qpairs;
    const int qt = (combo - obj * qpairs) * 2 + (int)rank;
    const int n_obj = args.n_live[obj] ? *reinterpret_cast<const volatile int32_t*>(args.n_live[obj]) : args.n[obj];
    const int tiles_o = args.n_live[obj] ? (n_obj + SC_TILE - 1) / SC_TILE : args.tiles[obj];
    const int t0 = (int)((long long)tiles_o * piece / pieces);
    const int t1 = (int)((long long)tiles_o * (piece + 1) / pieces);
    const int ntile = t1 - t0;
    const bool first_item = (item == cluster_id);

    if (warp >= 4) load_a_operand(args, obj, qt, tmem, TS_AH, TS_AL, warp, lane);
    tc_fence_before();
    cluster_sync_all();      // both CTAs' A tiles are in TMEM, both epilogues of the previous item are done
    tc_fence_after();

    if (warp == 0) {
      if (lane == 0) {
        for (int t = 0; t < ntile; ++t) {
          const uint32_t c = tile_ctr + t, st = c % SP_STAGES, ph = (c / SP_STAGES) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&k_full[st], MODE == MODE_MATCH ? SP_STAGE_BYTES : 2 * SP_STAGE_BYTES);
          const uint32_t kf = mapa_u32(smem_u32(&k_full[st]), 0);
          uint8_t* dst = kst + st * SP_STAGE_BYTES;
          const int row0 = (t0 + t) * SC_TILE + (int)rank * 64;     // this CTA's 64 of the tile's 128 slots
          tma_load_2d_pair(dst, &maps.kh[obj], kf, 0, row0);
          tma_load_2d_pair(dst + 8192, &maps.kh[obj], kf, 64, row0);
          if (MODE != MODE_MATCH) {                                 // the match reads only the hi half (SCAN_PASSES)
            tma_load_2d_pair(dst + 16384, &maps.kl[obj], kf, 0, row0);
            tma_load_2d_pair(dst + 24576, &maps.kl[obj], kf, 64, row0);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (leader && elect_one()) {   // ONE thread runs the whole issue loop of the item (see tc_phase_b_pair_kernel)
        constexpr uint32_t idesc = make_idesc(256, SC_TILE, FMT_F16, FMT_F16, 0, 0);
        for (int t = 0; t < ntile; ++t) {
          const uint32_t c = tile_ctr + t, st = c % SP_STAGES, ph = (c / SP_STAGES) & 1;
          const uint32_t sb = c % SC_BUFS, phs = (c / SC_BUFS) & 1;
          mbar_wait(&s_empty[sb], phs ^ 1);
          mbar_wait(&k_full[st], ph);
          tc_fence_after();
          const uint32_t kbase = smem_u32(kst + st * SP_STAGE_BYTES);
          const uint32_t d_t = tmem + TS_S + sb * SC_TILE;
          {
            // passes: (Ah,Bh) (Al,Bh) (Ah,Bl); each CTA's smem holds its 64 slots of the 128-slot B tile
#pragma unroll
            for (int pass = 0; pass < (MODE == MODE_MATCH ? 2 : 3); ++pass) {
              const uint32_t a_col = (pass == 1) ? TS_AL : TS_AH;
              const uint32_t kb = kbase + ((pass == 2) ? 16384u : 0u);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t bd = make_sdesc(kb + (ks >> 2) * 8192u + (ks & 3) * 32u, 16, 1024);
                mma_ts_pair(d_t, tmem + a_col + ks * 8, bd, idesc, (pass | ks) ? 1u : 0u);
              }
            }
            tc_commit_pair(&k_empty[st]);
            tc_commit_pair(&s_full[sb]);
          }
        }
      }
      __syncwarp();
    } else if (warp >= 4) {
      const int quarter = warp & 3, cg = (warp - 4) >> 2;
      const int row = (quarter << 5) + lane;
      const int et = threadIdx.x - 128;                       // 0..511
      const uint32_t tlane = tmem + (((uint32_t)quarter * 32u) << 16) + TS_S + (uint32_t)cg * 32u;
      float m_run = -INFINITY, l_run = 0.f;
      int cnt = 0;
      float2* ring = aux + et * MATCH_RING;
      const bool row_live = qt * QT + row < args.hw;
      for (int t = 0; t < ntile; ++t) {
        const uint32_t c = tile_ctr + t, st = c % SC_BUFS, ph = (c / SC_BUFS) & 1;
        mbar_wait(&s_full[st], ph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tlane + st * SC_TILE, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&s_empty[st]), 0));   // values are in registers: free the buffer (leader's barrier)
        if (args.dbg && blockIdx.x == 0 && first_item && t == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) args.dbg[row * SC_TILE + cg * 32 + i] = __uint_as_float(v[i]);
        }
        const int slot0 = (t0 + t) * SC_TILE + cg * 32;
        const int lim = n_obj - slot0;                        // valid slots in this thread's 32 columns
        if (lim <= 0) continue;
        if (lim < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= lim) v[i] = __float_as_uint(-INFINITY);
        }
        float cm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) cm = fmax3(cm, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        if (MODE == MODE_LSE) {
          const float m_new = fmaxf(m_run, cm);
          if (m_new > -INFINITY) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              a0 += ex2(__uint_as_float(v[i]) - m_new);
              a1 += ex2(__uint_as_float(v[i + 1]) - m_new);
            }
            l_run = l_run * ex2(m_run - m_new) + (a0 + a1);
            m_run = m_new;
          }
        } else {
          // candidates = every slot whose score is within `band` of the running maximum at its time; a jump of the
          // maximum by more than the band invalidates everything before it
          // (rows beyond hw are zero A rows: every slot ties at score 0 and would take the slow path for nothing)
          if (cm > m_run + args.band) cnt = 0;
          const float m_new = fmaxf(m_run, cm);
          const float thr = row_live ? m_new - args.band : INFINITY;
          // once the running maximum is established almost no tile holds a score inside the band: the 32 compares
          // below only run for the lanes' tiles whose own maximum reaches it
          if (cm >= thr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float sv = __uint_as_float(v[i]);
              if (sv >= thr) {
                ring[cnt & (MATCH_RING - 1)] = make_float2(sv, __int_as_float(slot0 + i));
                ++cnt;
              }
            }
          }
          m_run = m_new;
        }
      }
      const int j = qt * QT + row;
      if (MODE == MODE_LSE) {
        // combine the four column groups' statistics and publish the piece
        aux[cg * QT + row] = make_float2(m_run, l_run);
        named_bar_sync(1, EPI_THREADS);
        if (cg == 0) {
          float m = m_run;
#pragma unroll
          for (int g = 1; g < 4; ++g) m = fmaxf(m, aux[g * QT + row].x);
          float l = 0.f;
          if (m > -INFINITY) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float2 o = aux[g * QT + row];
              if (o.x > -INFINITY) l += o.y * ex2(o.x - m);
            }
          }
          if (j < args.hw) part[((size_t)obj * pieces + piece) * args.hw + j] = make_float2(m * LN2, l);
        }
        named_bar_sync(1, EPI_THREADS);                      // aux is reused by the next item
      } else {
        if (j < args.hw) {
          float2 e[MATCH_RING];
          const float thr = m_run - args.band;
#pragma unroll
          for (int r = 0; r < MATCH_RING; ++r) {
            const float2 x = ring[r];
            e[r] = (r < cnt && x.x >= thr) ? x : make_float2(-INFINITY, __int_as_float(0x7fffffff));
          }
          if (cnt > MATCH_RING) e[0] = make_float2(m_run, __int_as_float(-1));     // overflow: exact scan needed
          float4* dst = reinterpret_cast<float4*>(
              part + (((size_t)obj * pieces + piece) * args.hw + j) * MATCH_CAND + cg * MATCH_RING);
          dst[0] = make_float4(e[0].x, e[0].y, e[1].x, e[1].y);
          dst[1] = make_float4(e[2].x, e[2].y, e[3].x, e[3].y);
        }
      }
    }
    tile_ctr += ntile;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
}


// ------------------------------------------------------------------------------------------------
// phase B
//   TMEM: O^T accumulator [0,256) | Q hi [256,320) | Q lo [320,384) | two S/P buffers of 64 columns at 384 + 64 b
//   smem: K stages 32 KB [kh: 2 boxes of 64 d | kl: 2 boxes], V stages 64 KB [vh: 4 boxes of 64 ch | v8: 2 boxes of
//         128 ch | vl: 2 boxes], 64 slots per tile
//   MMA order  S(0) S(1) | O(0) S(2) | O(1) S(3) | ...: softmax(t) has a whole tile period (O(t-1) + S(t+1)) of cover.
//   Softmax warps: parity = tile & 1 picks the group of 8 warps (= S/P buffer), half picks 32 of the 64 slots.
//   P layout per 32-slot half (32 columns, written in place over S): [P hi fp16: 16 | e4m3(P - hi): 8 | e4m3(P): 8]
//   with P scaled by 256 (the scale is removed in the epilogue).
// ------------------------------------------------------------------------------------------------
constexpr int B_TILE = 64;
constexpr int B_KSTAGE_BYTES = B_TILE * DK * 2 * 2;                 // 32 KB
constexpr int B_VSTAGE_BYTES = B_TILE * 256 * (2 + 1 + 1);          // 64 KB
constexpr int B_SMEM = 2 * B_KSTAGE_BYTES + 2 * B_VSTAGE_BYTES + 1024 + 256 + 96 * 64 * 4;   // + alignment, barriers, item counts
constexpr uint32_t TM_O = 0, TM_QH = 256, TM_QL = 320, TM_S = 384;
constexpr float P_SCALE_LOG2 = 8.f, P_SCALE = 256.f;

// 16 logits -> P' = 2^(s - lse2 + 8): packed fp16 hi (8 regs), e4m3 residual (4), e4m3 P' (4).
// MASK: slots >= lim are zeroed (only the last tile of an object needs it).  COUNT: returns the threshold bits,
// bit (15 - i) set iff P'_i > thres_s (sign bit of thres_s - P' shifted in: one FADD + one SHF per element).
template <bool MASK, bool COUNT>
__device__ __forceinline__ uint32_t softmax_chunk16(const uint32_t (&s)[16], float lse2m, float thres_s, int lim,
                                                    uint32_t (&hi)[8], uint32_t (&lo)[4], uint32_t (&p8)[4]) {
  uint32_t bits = 0u;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float p0 = ex2(__uint_as_float(s[i]) - lse2m), p1 = ex2(__uint_as_float(s[i + 1]) - lse2m);
    if (MASK) {
      p0 = (i < lim) ? p0 : 0.f;
      p1 = (i + 1 < lim) ? p1 : 0.f;
    }
    if (COUNT) {
      bits = __funnelshift_l(__float_as_uint(thres_s - p0), bits, 1);
      bits = __funnelshift_l(__float_as_uint(thres_s - p1), bits, 1);
    }
    const uint32_t h = f16x2_rn(p0, p1);
    const float2 hf = f16x2_to_f32(h);
    hi[i >> 1] = h;
    const uint32_t l = e4m3x2(p0 - hf.x, p1 - hf.y), q = e4m3x2(p0, p1);
    if ((i & 2) == 0) { lo[i >> 2] = l; p8[i >> 2] = q; }
    else { lo[i >> 2] |= l << 16; p8[i >> 2] |= q << 16; }
  }
  return bits;
}

// usage counts of one item are accumulated in shared memory (CNT_TILES tiles x 64 slots) and flushed to the bank's
// counters with global atomics once per chunk of CNT_TILES tiles, not once per tile
constexpr int CNT_TILES = 96;
constexpr int CNT_BYTES = CNT_TILES * B_TILE * 4;

__global__ void __launch_bounds__(TC_THREADS, 1) tc_phase_b_kernel(const __grid_constant__ TcMaps maps, TcArgs args,
                                                                   const float* __restrict__ lse, float thres,
                                                                   int do_count, float* __restrict__ po) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;
  uint8_t* vst = smem + 2 * B_KSTAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(vst + 2 * B_VSTAGE_BYTES);
  uint64_t* k_full = bars;            // [2]
  uint64_t* k_empty = bars + 2;       // [2]
  uint64_t* v_full = bars + 4;        // [2]
  uint64_t* v_empty = bars + 6;       // [2]
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* p_full = bars + 10;       // [2]
  uint64_t* o_full = bars + 12;       // [1]
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + 13);
  int* cnt_item = reinterpret_cast<int*>(bars + 32);   // [CNT_TILES][64] usage counts of the current item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], EPI_WARPS);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < CNT_TILES * B_TILE; i += TC_THREADS) cnt_item[i] = 0;
  if (warp == 2) tmem_alloc(tmem_base_p, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;

  // work items = (split, object, query tile, channel half), round-robin over the persistent CTAs
  const int cpo = args.q_tiles * 2;
  const int n_combos = args.obj_n * cpo;
  const int pieces = tc_pieces(args);
  const int n_items = n_combos * pieces;
  uint32_t k_it = 0;            // tiles streamed so far (K/V stage = k_it & 1)
  uint32_t buf_it[2] = {0, 0};  // uses of each S/P buffer so far
  uint32_t seg_it = 0;          // items finished (o_full phase)
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int piece = item / n_combos;
    const int combo = item - piece * n_combos;
    const int obj = combo / cpo;
    const int cidx = combo - obj * cpo;                // qt * 2 + half
    const int qt = cidx >> 1, half = cidx & 1;
    const int n_obj = args.n_live[obj] ? *reinterpret_cast<const volatile int32_t*>(args.n_live[obj]) : args.n[obj];
    const int tiles_o = args.n_live[obj] ? (n_obj + B_TILE - 1) / B_TILE : args.tiles[obj];
    const int t0 = (int)((long long)tiles_o * piece / pieces);
    const int t1 = (int)((long long)tiles_o * (piece + 1) / pieces);
    const int ntile = t1 - t0;
    const bool first_item = (item == (int)blockIdx.x);

    if (warp >= 4) load_a_operand(args, obj, qt, tmem, TM_QH, TM_QL, warp, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      if (lane == 0) {
        for (int t = 0; t < ntile; ++t) {
          const uint32_t kit = k_it + t, st = kit & 1, ph = (kit >> 1) & 1;
          const int row0 = (t0 + t) * B_TILE;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], B_KSTAGE_BYTES);
          uint8_t* kd = kst + st * B_KSTAGE_BYTES;
          tma_load_2d(kd, &maps.kh[obj], &k_full[st], 0, row0);
          tma_load_2d(kd + 8192, &maps.kh[obj], &k_full[st], 64, row0);
          tma_load_2d(kd + 16384, &maps.kl[obj], &k_full[st], 0, row0);
          tma_load_2d(kd + 24576, &maps.kl[obj], &k_full[st], 64, row0);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], B_VSTAGE_BYTES);
          uint8_t* vd = vst + st * B_VSTAGE_BYTES;
#pragma unroll
          for (int g = 0; g < 4; ++g) tma_load_2d(vd + g * 8192, &maps.vh[obj], &v_full[st], half * 256 + g * 64, row0);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            tma_load_2d(vd + 32768 + g * 8192, &maps.v8[obj], &v_full[st], half * 256 + g * 128, row0);
            tma_load_2d(vd + 49152 + g * 8192, &maps.vl[obj], &v_full[st], half * 256 + g * 128, row0);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // the whole warp runs the loop (uniform control flow and operands); one elected lane issues
      constexpr uint32_t idesc_s = make_idesc(128, B_TILE, FMT_F16, FMT_F16, 0, 0);
      constexpr uint32_t idesc_o = make_idesc(128, 256, FMT_F16, FMT_F16, 0, 1);
      constexpr uint32_t idesc_o8 = make_idesc(128, 256, FMT_E4M3, FMT_E4M3, 0, 1);    // e4m3(P - hi) x e4m3(V)
      constexpr uint32_t idesc_ol = make_idesc(128, 256, FMT_E4M3, FMT_E5M2, 0, 1);    // e4m3(P) x e5m2(V - Vhi)
      auto issue_s = [&](int t) {
        const uint32_t kit = k_it + t, st = kit & 1, ph = (kit >> 1) & 1;
        const int b = t & 1;
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint32_t kbase = smem_u32(kst + st * B_KSTAGE_BYTES);
        const uint32_t d_t = tmem + TM_S + (uint32_t)b * 64;
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_col = (pass == 1) ? TM_QL : TM_QH;
            const uint32_t kb = kbase + ((pass == 2) ? 16384u : 0u);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t bd = make_sdesc(kb + (ks >> 2) * 8192u + (ks & 3) * 32u, 16, 1024);
              mma_ts(d_t, tmem + a_col + ks * 8, bd, idesc_s, (pass | ks) ? 1u : 0u);
            }
          }
          tc_commit(&k_empty[st]);
          tc_commit(&s_full[b]);
        }
        __syncwarp();
      };
      if (ntile > 0) issue_s(0);
      if (ntile > 1) issue_s(1);
      for (int t = 0; t < ntile; ++t) {
        const uint32_t kit = k_it + t, st = kit & 1, ph = (kit >> 1) & 1;
        const int b = t & 1;
        mbar_wait(&p_full[b], buf_it[b] & 1);
        ++buf_it[b];
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint32_t vbase = smem_u32(vst + st * B_VSTAGE_BYTES);
        const uint32_t pcol = tmem + TM_S + (uint32_t)b * 64;
        if (elect_one()) {
          // main pass: fp16 P hi x fp16 V hi, 16 slots per instruction
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bd = make_sdesc(vbase + ks * 2048u, 8192, 1024);
            mma_ts(tmem + TM_O, pcol + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, bd, idesc_o, (t | ks) ? 1u : 0u);
          }
          // correction passes (fp8, 32 slots per instruction)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t b8 = make_sdesc(vbase + 32768u + h * 4096u, 8192, 1024);
            mma_ts_f8(tmem + TM_O, pcol + (uint32_t)h * 32u + 16u, b8, idesc_o8, 1u);
            const uint64_t bl = make_sdesc(vbase + 49152u + h * 4096u, 8192, 1024);
            mma_ts_f8(tmem + TM_O, pcol + (uint32_t)h * 32u + 24u, bl, idesc_ol, 1u);
          }
          tc_commit(&v_empty[st]);
        }
        __syncwarp();
        if (t + 2 < ntile) issue_s(t + 2);   // in-order MMA pipe: S(t+2) overwrites buffer b only after O(t) read it
      }
      if (elect_one()) tc_commit(o_full);
      __syncwarp();
    } else if (warp >= 4) {
      // every softmax warp works on every tile: slot group sg = 16 of the tile's 64 slots, for its 32 query rows
      const int quarter = warp & 3, sg = (warp - 4) >> 2;
      const int hsel = sg >> 1, sub = sg & 1;
      const int row = (quarter << 5) + lane;
      const int et = threadIdx.x - 128;                          // 0..511
      const uint32_t tlane = tmem + (((uint32_t)quarter * 32u) << 16);
      const int j = qt * QT + row;
      const float lse2m = (j < args.hw) ? (lse[(size_t)obj * args.hw + j] * LOG2E - P_SCALE_LOG2) : INFINITY;
      const float thres_s = thres * P_SCALE;
      const bool counting = do_count && (half == 0);
      auto flush_counts = [&](int tile_first, int n_tiles) {
        named_bar_sync(9, EPI_THREADS);
        for (int idx = et; idx < n_tiles * B_TILE; idx += EPI_THREADS) {
          const int c = cnt_item[idx];
          if (c) {
            atomicAdd(&args.cnt[obj][(size_t)(t0 + tile_first) * B_TILE + idx], c);
            cnt_item[idx] = 0;
          }
        }
        named_bar_sync(9, EPI_THREADS);
      };
      int chunk0 = 0;                                            // first tile of the current count chunk
      for (int t = 0; t < ntile; ++t) {
        const int b = t & 1;
        mbar_wait(&s_full[b], buf_it[b] & 1);
        ++buf_it[b];
        tc_fence_after();
        const int slot0 = (t0 + t) * B_TILE + sg * 16;
        const uint32_t pb = tlane + TM_S + (uint32_t)b * 64 + (uint32_t)hsel * 32;   // this half's 32-column S / P region
        uint32_t sv[16];
        tmem_ld16(pb + (uint32_t)sub * 16, sv);
        tmem_wait_ld();
        if (args.dbg && blockIdx.x == 0 && first_item && t == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) args.dbg[row * B_TILE + sg * 16 + i] = __uint_as_float(sv[i]);
        }
        // P is written in place over S and the fp8 operands of the two 16-slot groups of a half interleave: all four
        // warps of this lane quarter must hold their logits in registers before any of them stores
        tc_fence_before();
        named_bar_sync(5 + quarter, 128);
        tc_fence_after();
        uint32_t hi[8], lo[4], p8[4];
        uint32_t bits;
        const int lim = n_obj - slot0;
        if (lim >= 16) {
          bits = counting ? softmax_chunk16<false, true>(sv, lse2m, thres_s, lim, hi, lo, p8)
                          : softmax_chunk16<false, false>(sv, lse2m, thres_s, lim, hi, lo, p8);
        } else {
          bits = softmax_chunk16<true, true>(sv, lse2m, thres_s, lim, hi, lo, p8);
        }
        tmem_st8(pb + (uint32_t)sub * 8, hi);
        tmem_st4(pb + 16 + (uint32_t)sub * 4, lo);
        tmem_st4(pb + 24 + (uint32_t)sub * 4, p8);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
        if (counting) {
          // bit (15 - i) <-> slot slot0 + i of this row; rows beyond hw have lse = +inf, P = 0: no bits
          int* cdst = cnt_item + (t - chunk0) * B_TILE + sg * 16;
          if (__any_sync(0xffffffffu, __popc(bits) > 2)) {
            int c_mine = 0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const unsigned bal = __ballot_sync(0xffffffffu, (bits >> c) & 1u);
              if (lane == c) c_mine = __popc(bal);
            }
            if (c_mine) atomicAdd(&cdst[15 - lane], c_mine);
          } else {
            while (bits) {
              const int c = 31 - __clz((int)bits);
              bits &= ~(1u << c);
              atomicAdd(&cdst[15 - c], 1);
            }
          }
          if (t - chunk0 + 1 == CNT_TILES && t + 1 < ntile) {
            flush_counts(chunk0, CNT_TILES);
            chunk0 = t + 1;
          }
        }
      }
      if (counting && ntile > chunk0) flush_counts(chunk0, ntile - chunk0);
      // epilogue: O^T (128 queries x 256 channels) * 2^-8 -> partial buffer; group g takes channels [64 g, 64 g + 64)
      mbar_wait(o_full, seg_it & 1);
      tc_fence_after();
      if (ntile > 0) {
        float* dst = po + (((size_t)obj * pieces + piece) * DV + half * 256 + sg * 64) * (size_t)args.hw;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(tlane + TM_O + (uint32_t)sg * 64 + ch * 32, v);
          tmem_wait_ld();
          if (j < args.hw) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)(ch * 32 + i) * args.hw + j] = __uint_as_float(v[i]) * (1.f / P_SCALE);
          }
        }
      }
    }
    k_it += ntile;
    if (warp < 4 && warp != 1) { buf_it[0] += (ntile + 1) >> 1; buf_it[1] += ntile >> 1; }
    ++seg_it;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// phase B on CTA pairs (cta_group::2, M = 256): the two CTAs of a cluster take two adjacent query tiles of the same
// (object, slot split, channel half) and SHARE every K / V tile - each CTA TMA-loads half of it (K: 32 of the 64
// slots, V: 128 of the 256 channels) and the pair's tensor cores read both halves.  Per CTA this halves the L2 -> smem
// traffic and the smem bandwidth per MMA (the single-CTA kernel needs ~118 B/clk of shared memory, profiles/r1b).
// Everything per CTA (TMEM layout, softmax, epilogue) is as in tc_phase_b_kernel; only the leader CTA (rank 0) issues
// MMAs, its "full" barriers count the bytes of both CTAs' loads, and commits are multicast to both CTAs' barriers.
// ------------------------------------------------------------------------------------------------
constexpr int P_STAGES = 4;
constexpr int P_KSTAGE_BYTES = 32 * DK * 2 * 2;                 // this CTA's 32 slots, hi + lo: 16 KB
constexpr int P_VSTAGE_BYTES = B_TILE * 128 * (2 + 1 + 1);      // 64 slots x this CTA's 128 channels: 32 KB
constexpr int P_SMEM = P_STAGES * (P_KSTAGE_BYTES + P_VSTAGE_BYTES) + 1024 + 256 + 96 * 64 * 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    tc_phase_b_pair_kernel(const __grid_constant__ TcMaps maps, TcArgs args, const float* __restrict__ lse, float thres,
                           int do_count, float* __restrict__ po) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;
  uint8_t* vst = smem + P_STAGES * P_KSTAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(vst + P_STAGES * P_VSTAGE_BYTES);
  uint64_t* k_full = bars;             // [4]  (used on the leader)
  uint64_t* k_empty = bars + 4;        // [4]
  uint64_t* v_full = bars + 8;         // [4]  (used on the leader)
  uint64_t* v_empty = bars + 12;       // [4]
  uint64_t* s_full = bars + 16;        // [2]
  uint64_t* p_full = bars + 18;        // [2]  (used on the leader: 8 local + 8 remote warps)
  uint64_t* o_full = bars + 20;        // [1]
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + 21);
  int* cnt_item = reinterpret_cast<int*>(bars + 32);   // [CNT_TILES][64] usage counts of the current item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 2 * EPI_WARPS); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < CNT_TILES * B_TILE; i += TC_THREADS) cnt_item[i] = 0;
  if (warp == 2) tmem_alloc_pair(tmem_base_p, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;
  pdl_wait();        // programmatic dependent launch: the prologue above overlaps the predecessor's tail
  pdl_trigger();

  // work items = (split, object, query-tile pair, channel half), round-robin over the clusters
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int qpairs = (args.q_tiles + 1) >> 1;
  const int cpo = qpairs * 2;
  const int n_combos = args.obj_n * cpo;
  const int pieces = tc_pieces(args);
  const int n_items = n_combos * pieces;
  uint32_t k_it = 0;            // tiles streamed so far (stage = k_it & 3)
  uint32_t buf_it[2] = {0, 0};  // uses of each S/P buffer so far
  uint32_t seg_it = 0;          // items finished (o_full phase)
  for (int item = cluster_id; item < n_items; item += n_clusters) {
    const int piece = item / n_combos;
    const int combo = item - piece * n_combos;
    const int obj = combo / cpo;
    const int cidx = combo - obj * cpo;                // qp * 2 + half
    const int qt = (cidx >> 1) * 2 + (int)rank, half = cidx & 1;
    const int n_obj = args.n_live[obj] ? *reinterpret_cast<const volatile int32_t*>(args.n_live[obj]) : args.n[obj];
    const int tiles_o = args.n_live[obj] ? (n_obj + B_TILE - 1) / B_TILE : args.tiles[obj];
    const int t0 = (int)((long long)tiles_o * piece / pieces);
    const int t1 = (int)((long long)tiles_o * (piece + 1) / pieces);
    const int ntile = t1 - t0;
    const bool first_item = (item == cluster_id);

    if (warp >= 4) load_a_operand(args, obj, qt, tmem, TM_QH, TM_QL, warp, lane);
    tc_fence_before();
    cluster_sync_all();      // both CTAs' Q tiles are in TMEM, both epilogues of the previous item are done
    tc_fence_after();

    if (warp == 0) {
      if (lane == 0) {
        for (int t = 0; t < ntile; ++t) {
          const uint32_t kit = k_it + t, st = kit & 3, ph = (kit >> 2) & 1;
          const int row0 = (t0 + t) * B_TILE;
          mbar_wait(&k_empty[st], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&k_full[st], 2 * P_KSTAGE_BYTES);
          const uint32_t kf = mapa_u32(smem_u32(&k_full[st]), 0);
          uint8_t* kd = kst + st * P_KSTAGE_BYTES;
          const int krow = row0 + (int)rank * 32;
          tma_load_2d_pair(kd, &maps.kh[obj], kf, 0, krow);
          tma_load_2d_pair(kd + 4096, &maps.kh[obj], kf, 64, krow);
          tma_load_2d_pair(kd + 8192, &maps.kl[obj], kf, 0, krow);
          tma_load_2d_pair(kd + 12288, &maps.kl[obj], kf, 64, krow);
          mbar_wait(&v_empty[st], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&v_full[st], 2 * P_VSTAGE_BYTES);
          const uint32_t vf = mapa_u32(smem_u32(&v_full[st]), 0);
          uint8_t* vd = vst + st * P_VSTAGE_BYTES;
          const int ch0 = half * 256 + (int)rank * 128;
          tma_load_2d_pair(vd, &maps.vh[obj], vf, ch0, row0);
          tma_load_2d_pair(vd + 8192, &maps.vh[obj], vf, ch0 + 64, row0);
          tma_load_2d_pair(vd + 16384, &maps.v8[obj], vf, ch0, row0);
          tma_load_2d_pair(vd + 24576, &maps.vl[obj], vf, ch0, row0);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ONE elected thread runs the whole issue loop of the item (waits, MMAs, commits).  With an elect + __syncwarp around
      // every MMA group the steady tile period was 2092 clk; like this it is 1902 (tests/debug_item_times.py) against
      // an MMA floor of 1795 (tests/microbench/mma_rate.cu mix): every instruction the issuing warp spends between two
      // MMAs shows up in the tile period.
      if (leader && elect_one()) {
        constexpr uint32_t idesc_s = make_idesc(256, B_TILE, FMT_F16, FMT_F16, 0, 0);
        constexpr uint32_t idesc_o = make_idesc(256, 256, FMT_F16, FMT_F16, 0, 1);
        constexpr uint32_t idesc_o8 = make_idesc(256, 256, FMT_E4M3, FMT_E4M3, 0, 1);
        constexpr uint32_t idesc_ol = make_idesc(256, 256, FMT_E4M3, FMT_E5M2, 0, 1);
        auto issue_s = [&](int t) {
          const uint32_t kit = k_it + t, st = kit & 3, ph = (kit >> 2) & 1;
          const int b = t & 1;
          mbar_wait(&k_full[st], ph);
          tc_fence_after();
          const uint32_t kbase = smem_u32(kst + st * P_KSTAGE_BYTES);
          const uint32_t d_t = tmem + TM_S + (uint32_t)b * 64;
          {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a_col = (pass == 1) ? TM_QL : TM_QH;
              const uint32_t kb = kbase + ((pass == 2) ? 8192u : 0u);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t bd = make_sdesc(kb + (ks >> 2) * 4096u + (ks & 3) * 32u, 16, 1024);
                mma_ts_pair(d_t, tmem + a_col + ks * 8, bd, idesc_s, (pass | ks) ? 1u : 0u);
              }
            }
            tc_commit_pair(&k_empty[st]);
            tc_commit_pair(&s_full[b]);
          }
        };
        long long* ts = args.tstamp ? args.tstamp + ((size_t)cluster_id * 64 + (size_t)(item / n_clusters)) * 8 : nullptr;
        if (ts && item / n_clusters < 64) { ts[0] = clock64(); ts[4] = ntile; }
        if (ntile > 0) issue_s(0);
        if (ntile > 1) issue_s(1);
        for (int t = 0; t < ntile; ++t) {
          const uint32_t kit = k_it + t, st = kit & 3, ph = (kit >> 2) & 1;
          const int b = t & 1;
          mbar_wait(&p_full[b], buf_it[b] & 1);
          if (ts && item / n_clusters < 64) { if (t == 0) ts[1] = clock64(); if (t == ntile - 1) ts[2] = clock64(); }
          ++buf_it[b];
          mbar_wait(&v_full[st], ph);
          tc_fence_after();
          const uint32_t vbase = smem_u32(vst + st * P_VSTAGE_BYTES);
          const uint32_t pcol = tmem + TM_S + (uint32_t)b * 64;
          {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t bd = make_sdesc(vbase + ks * 2048u, 8192, 1024);
              mma_ts_pair(tmem + TM_O, pcol + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, bd, idesc_o,
                          (t | ks) ? 1u : 0u);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t b8 = make_sdesc(vbase + 16384u + h * 4096u, 8192, 1024);
              mma_ts_f8_pair(tmem + TM_O, pcol + (uint32_t)h * 32u + 16u, b8, idesc_o8, 1u);
              const uint64_t bl = make_sdesc(vbase + 24576u + h * 4096u, 8192, 1024);
              mma_ts_f8_pair(tmem + TM_O, pcol + (uint32_t)h * 32u + 24u, bl, idesc_ol, 1u);
            }
            tc_commit_pair(&v_empty[st]);
          }
          if (t + 2 < ntile) issue_s(t + 2);
        }
        tc_commit_pair(o_full);
      }
      __syncwarp();
    } else if (warp >= 4) {
      // every softmax warp works on every tile: slot group sg = 16 of the tile's 64 slots, for its 32 query rows
      const int quarter = warp & 3, sg = (warp - 4) >> 2;
      const int hsel = sg >> 1, sub = sg & 1;
      const int row = (quarter << 5) + lane;
      const int et = threadIdx.x - 128;                          // 0..511
      const uint32_t tlane = tmem + (((uint32_t)quarter * 32u) << 16);
      const int j = qt * QT + row;
      const float lse2m = (j < args.hw) ? (lse[(size_t)obj * args.hw + j] * LOG2E - P_SCALE_LOG2) : INFINITY;
      const float thres_s = thres * P_SCALE;
      const bool counting = do_count && (half == 0);
      const uint32_t pf_leader0 = mapa_u32(smem_u32(&p_full[0]), 0), pf_leader1 = mapa_u32(smem_u32(&p_full[1]), 0);
      auto flush_counts = [&](int tile_first, int n_tiles) {
        named_bar_sync(9, EPI_THREADS);
        for (int idx = et; idx < n_tiles * B_TILE; idx += EPI_THREADS) {
          const int c = cnt_item[idx];
          if (c) {
            atomicAdd(&args.cnt[obj][(size_t)(t0 + tile_first) * B_TILE + idx], c);
            cnt_item[idx] = 0;
          }
        }
        named_bar_sync(9, EPI_THREADS);
      };
      int chunk0 = 0;                                            // first tile of the current count chunk
      for (int t = 0; t < ntile; ++t) {
        const int b = t & 1;
        mbar_wait(&s_full[b], buf_it[b] & 1);
        ++buf_it[b];
        tc_fence_after();
        const int slot0 = (t0 + t) * B_TILE + sg * 16;
        const uint32_t pb = tlane + TM_S + (uint32_t)b * 64 + (uint32_t)hsel * 32;   // this half's 32-column S / P region
        uint32_t sv[16];
        tmem_ld16(pb + (uint32_t)sub * 16, sv);
        tmem_wait_ld();
        if (args.dbg && blockIdx.x == 0 && first_item && t == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) args.dbg[row * B_TILE + sg * 16 + i] = __uint_as_float(sv[i]);
        }
        // P is written in place over S and the fp8 operands of the two 16-slot groups of a half interleave: all four
        // warps of this lane quarter must hold their logits in registers before any of them stores
        tc_fence_before();
        named_bar_sync(5 + quarter, 128);
        tc_fence_after();
        uint32_t hi[8], lo[4], p8[4];
        uint32_t bits;
        const int lim = n_obj - slot0;
        if (lim >= 16) {
          bits = counting ? softmax_chunk16<false, true>(sv, lse2m, thres_s, lim, hi, lo, p8)
                          : softmax_chunk16<false, false>(sv, lse2m, thres_s, lim, hi, lo, p8);
        } else {
          bits = softmax_chunk16<true, true>(sv, lse2m, thres_s, lim, hi, lo, p8);
        }
        tmem_st8(pb + (uint32_t)sub * 8, hi);
        tmem_st4(pb + 16 + (uint32_t)sub * 4, lo);
        tmem_st4(pb + 24 + (uint32_t)sub * 4, p8);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(b ? pf_leader1 : pf_leader0);
        if (counting) {
          // bit (15 - i) <-> slot slot0 + i of this row; rows beyond hw have lse = +inf, P = 0: no bits
          int* cdst = cnt_item + (t - chunk0) * B_TILE + sg * 16;
          if (__any_sync(0xffffffffu, __popc(bits) > 2)) {
            int c_mine = 0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const unsigned bal = __ballot_sync(0xffffffffu, (bits >> c) & 1u);
              if (lane == c) c_mine = __popc(bal);
            }
            if (c_mine) atomicAdd(&cdst[15 - lane], c_mine);
          } else {
            while (bits) {
              const int c = 31 - __clz((int)bits);
              bits &= ~(1u << c);
              atomicAdd(&cdst[15 - c], 1);
            }
          }
          if (t - chunk0 + 1 == CNT_TILES && t + 1 < ntile) {
            flush_counts(chunk0, CNT_TILES);
            chunk0 = t + 1;
          }
        }
      }
      if (counting && ntile > chunk0) flush_counts(chunk0, ntile - chunk0);
      mbar_wait(o_full, seg_it & 1);
      tc_fence_after();
      if (ntile > 0) {
        float* dst = po + (((size_t)obj * pieces + piece) * DV + half * 256 + sg * 64) * (size_t)args.hw;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(tlane + TM_O + (uint32_t)sg * 64 + ch * 32, v);
          tmem_wait_ld();
          if (j < args.hw) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)(ch * 32 + i) * args.hw + j] = __uint_as_float(v[i]) * (1.f / P_SCALE);
          }
        }
      }
    }
    k_it += ntile;
    ++seg_it;
    tc_fence_before();
    __syncthreads();
    if (args.tstamp && leader && threadIdx.x == 0 && item / n_clusters < 64)
      args.tstamp[((size_t)cluster_id * 64 + (size_t)(item / n_clusters)) * 8 + 3] = clock64();
    tc_fence_after();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// Exact fp32 re-score of the tensor-core candidates (FeatureBank.py:66-68): one warp per candidate query.
// The fp16x3 scores carry the tensor core's truncating accumulation (measured: ~4e-6 systematic bias), so every slot
// whose approximate score lies within MATCH_BAND of the approximate maximum is re-evaluated with the sequential fp32
// FMA chain the SIMT kernel uses (bit-identical values and ordering, ties -> lowest slot).  An overflow marker (more
// than four near-ties in one column group of one item) sends that query to an exact scan of all slots.
// ------------------------------------------------------------------------------------------------
// The scan runs TWO passes for the match, (Ah + Al) x Bh: the candidate is exact to ~2^-22, the bank operand is the fp16
// rounding of 16 * nk, so |approximate - exact| <= 2^-11 * sum |a_i b_i| <= 2^-11 (unit vectors) + the accumulate
// truncation.  Any slot that can be the true arg-max lies within twice that bound of the approximate maximum: the band.
// (Three passes with a 2e-5 band cost a third more MMA work - and energy, which is what these kernels are bound by -
// for the same exact result after the re-score.)
constexpr float MATCH_BAND = 1.05e-3f;

__device__ __forceinline__ float exact_dot128(const float* __restrict__ nk, const float* __restrict__ q, int64_t slot) {
  const float4* a = reinterpret_cast<const float4*>(nk + slot * DK);
  const float4* b = reinterpret_cast<const float4*>(q);
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < DK / 4; ++k) {
    const float4 av = a[k], bv = b[k];
    acc = fmaf(av.x, bv.x, acc);
    acc = fmaf(av.y, bv.y, acc);
    acc = fmaf(av.z, bv.z, acc);
    acc = fmaf(av.w, bv.w, acc);
  }
  return acc;
}

struct RescoreArgs {
  int obj_n, pieces, hw;
  const int32_t* pieces_dev;
  int n[TC_MAX_OBJ];
  const int32_t* n_live[TC_MAX_OBJ];
  const float* nk[TC_MAX_OBJ];
  const float* nck[TC_MAX_OBJ];
  int32_t* idx_out[TC_MAX_OBJ];
  float* corr_out[TC_MAX_OBJ];
  float band;
};

__global__ void __launch_bounds__(256) match_rescore_kernel(const float2* __restrict__ part, RescoreArgs a) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int obj = blockIdx.y;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= a.hw) return;
  const int n = a.n_live[obj] ? *reinterpret_cast<const volatile int32_t*>(a.n_live[obj]) : a.n[obj];
  const float* nk = a.nk[obj];
  const int pieces = a.pieces_dev ? *reinterpret_cast<const volatile int32_t*>(a.pieces_dev) : a.pieces;
  const int n_cand = pieces * MATCH_CAND;
  const float* q = a.nck[obj] + (size_t)j * DK;
  auto entry = [&](int c) {
    const int pc = c / MATCH_CAND, r = c - pc * MATCH_CAND;
    return part[(((size_t)obj * pieces + pc) * a.hw + j) * MATCH_CAND + r];
  };
  // approximate maximum over all pieces (overflow markers carry their group's maximum)
  float amax = -INFINITY;
  for (int c = lane; c < n_cand; c += 32) amax = fmaxf(amax, entry(c).x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float band = amax - a.band;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  int overflow = 0;
  for (int c = lane; c < n_cand; c += 32) {
    const float2 e = entry(c);
    const int slot = __float_as_int(e.y);
    if (e.x >= band) {
      if (slot < 0) { overflow = 1; continue; }
      if (slot >= n) continue;
      const float v = exact_dot128(nk, q, slot);
      if (v > best || (v == best && slot < bidx)) { best = v; bidx = slot; }
    }
  }
  overflow = __any_sync(0xffffffffu, overflow);
  if (overflow) {
    best = -INFINITY;
    bidx = 0x7fffffff;
    for (int slot = lane; slot < n; slot += 32) {
      const float v = exact_dot128(nk, q, slot);
      if (v > best) { best = v; bidx = slot; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  if (lane == 0) {
    a.idx_out[obj][j] = bidx;
    a.corr_out[obj][j] = best;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// SM count of the CURRENT device (the caller's context: the Python host enters the bank's device around every call),
// cached per device ordinal - a process may drive banks on several GPUs
static int num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

static float* g_dbg = nullptr;
static long long* g_tstamp = nullptr;   // vfn_debug_set_tstamp()
static int g_pair = 3;    // bit 0: CTA-pair (cta_group::2) phase B, bit 1: CTA-pair scan (phase A, match); vfn_debug_set_pair()

bool tc_shapes_ok(int d_key, int d_val) { return d_key == DK && d_val == DV; }

// Banks with a device-resident live count (vfn_bank::n_live): `n` is an upper bound, `n_min` a lower bound; the tensor
// maps then span the whole slab (rows in [live, cap) hold finite stale operands, masked by the kernels) and the work
// partition is chosen so that every piece keeps a tile even at the lower bound.
static int64_t map_rows(const vfn_bank& b) { return b.n_live ? b.cap : b.n; }
static int64_t n_low(const vfn_bank& b) { return b.n_live ? b.n_min : b.n; }

constexpr int B_CHAIN_MAX = 8192;      // slots accumulated into one TMEM readout accumulator (see tc_phase_b)

// number of slot splits per (object, query tile[, half]) combo: minimise rounds x (tiles per item + fixed per-item
// overhead) for `combos` combos dealt round-robin to G persistent CTAs; every item keeps at least one tile.
static int best_split(int combos, int64_t tiles_min, int64_t tiles_max, int overhead_tiles, int G = 0, int s_min = 1) {
  if (G <= 0) G = num_sms();
  if (s_min > TC_MAX_SPLIT) s_min = TC_MAX_SPLIT;
  if (s_min > tiles_min) s_min = (int)(tiles_min > 1 ? tiles_min : 1);
  int best = s_min;
  double best_cost = 1e30;
  for (int s = s_min; s <= TC_MAX_SPLIT && s <= tiles_min; ++s) {
    const int64_t rounds = cdiv((int64_t)combos * s, G);
    const double cost = (double)rounds * (double)(cdiv(tiles_max, s) + overhead_tiles);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

// all banks of a call carry a live count, or none does
static int live_mode(const vfn_bank* banks, int obj_n, bool* live) {
  *live = banks[0].n_live != nullptr;
  for (int o = 1; o < obj_n; ++o)
    VFN_CHECK_ARG((banks[o].n_live != nullptr) == *live, "banks mix device-resident and host-tracked sizes");
  return VFN_OK;
}

static void set_split_plan(TcArgs* a, int tile, int combos, int overhead, int G, int chain_max, int32_t* cell) {
  a->plan_tile = tile; a->plan_combos = combos; a->plan_overhead = overhead; a->plan_G = G; a->plan_chain_max = chain_max;
  a->pieces_dev = cell;
}

void tc_pick_splits(int obj_n, int64_t n_max, int64_t hw, int* split_a, int* split_b) {
  // upper bounds used to size the workspace; the per-launch choice (<= these) is made in tc_phase_a / tc_phase_b
  (void)obj_n; (void)n_max; (void)hw;
  *split_a = TC_MAX_SPLIT;
  *split_b = TC_MAX_SPLIT;
}

// rows padded to a whole number of query-tile PAIRS (the pair kernels read two adjacent tiles)
static size_t a_operand_bytes(int64_t hw) { return align_up((size_t)cdiv(hw, 2 * QT) * 2 * QT * DK * sizeof(uint16_t), 256); }
int64_t tc_operand_rows(int64_t hw) { return cdiv(hw, 2 * QT) * 2 * QT; }

// [Q hi | Q lo | 256 B: device-chosen splits {phase A, phase B}]
size_t tc_workspace_bytes(int obj_n, int64_t hw) {
  (void)obj_n;
  return 2 * a_operand_bytes(hw) + 256;
}
static int32_t* tc_plan_cell(char* ws_tc, int64_t hw) { return reinterpret_cast<int32_t*>(ws_tc + 2 * a_operand_bytes(hw)); }

static int set_attrs() {
  static bool attr[64] = {false};
  if (first_use_on_device(attr)) {
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_scan_kernel<MODE_LSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_scan_kernel<MODE_MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_scan_pair_kernel<MODE_LSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_scan_pair_kernel<MODE_MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_phase_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_phase_b_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
  }
  return VFN_OK;
}

static int fill_args(const vfn_bank* banks, int obj_n, int64_t hw, int pieces, int tile, char* ws_tc, TcMaps* maps,
                     TcArgs* a, bool need_v, int k_box_rows = 0) {
  if (k_box_rows <= 0) k_box_rows = tile;
  VFN_CHECK_ARG(obj_n <= TC_MAX_OBJ, "tcgen05 read supports at most %d objects", TC_MAX_OBJ);
  a->obj_n = obj_n; a->hw = (int)hw; a->q_tiles = (int)cdiv(hw, QT); a->pieces = pieces;
  a->qh = reinterpret_cast<const uint16_t*>(ws_tc);
  a->ql = reinterpret_cast<const uint16_t*>(ws_tc + a_operand_bytes(hw));
  a->a_obj_stride = 0;
  a->pieces_dev = nullptr;
  a->band = 0.f;
  a->dbg = g_dbg;
  a->tstamp = g_tstamp;
  for (int o = 0; o < obj_n; ++o) {
    VFN_CHECK_ARG(banks[o].kh && banks[o].vh, "bank %d has no tensor-core operand arrays", o);
    VFN_CHECK_ARG(banks[o].n < (1ll << 31), "bank too large");
    VFN_CHECK_ARG(!banks[o].n_live || (banks[o].n_min >= 1 && banks[o].n_min <= banks[o].n), "bank %d: bad n_min", o);
    a->n[o] = (int)banks[o].n;
    a->tiles[o] = (int)cdiv(banks[o].n, tile);
    a->n_live[o] = banks[o].n_live;
    a->cnt[o] = banks[o].cnt;
    const int64_t rows = map_rows(banks[o]);
    if (int rc = make_map(&maps->kh[o], banks[o].kh, rows, DK, k_box_rows, 2)) return rc;
    if (int rc = make_map(&maps->kl[o], banks[o].kl, rows, DK, k_box_rows, 2)) return rc;
    if (need_v) {
      if (int rc = make_map(&maps->vh[o], banks[o].vh, rows, DV, tile, 2)) return rc;
      if (int rc = make_map(&maps->v8[o], banks[o].v8, rows, DV, tile, 1)) return rc;
      if (int rc = make_map(&maps->vl[o], banks[o].vl, rows, DV, tile, 1)) return rc;
    }
  }
  for (int o = obj_n; o < TC_MAX_OBJ; ++o) a->n_live[o] = nullptr;
  return VFN_OK;
}

int tc_phase_a(const vfn_bank* banks, int obj_n, const float* q_in_dm, int64_t hw, int split_a, float2* part,
               char* ws_tc, cudaStream_t st, int* pieces_out, const int32_t** pieces_dev_out, int q_em) {
  if (int rc = set_attrs()) return rc;
  bool live;
  if (int rc = live_mode(banks, obj_n, &live)) return rc;
  TcMaps maps;
  TcArgs a = {};
  int64_t tmin = INT64_MAX, tmax = 0;
  for (int o = 0; o < obj_n; ++o) {
    const int64_t t = cdiv(banks[o].n, SC_TILE), tl = cdiv(n_low(banks[o]), SC_TILE);
    tmin = tl < tmin ? tl : tmin;
    tmax = t > tmax ? t : tmax;
  }
  const bool pair = (g_pair & 2) && (num_sms() % 2 == 0);
  const int pieces = pair ? best_split(obj_n * (int)cdiv(hw, 2 * QT), tmin, tmax, 2, num_sms() / 2)
                          : best_split(obj_n * (int)cdiv(hw, QT), tmin, tmax, 2);
  if (pieces > split_a) { set_error("phase A: split %d exceeds workspace bound %d", pieces, split_a); return VFN_E_CAPACITY; }
  *pieces_out = pieces;
  if (int rc = fill_args(banks, obj_n, hw, pieces, SC_TILE, ws_tc, &maps, &a, false, pair ? 64 : SC_TILE)) return rc;
  *pieces_dev_out = nullptr;
  if (live) {
    int32_t* cell = tc_plan_cell(ws_tc, hw);
    set_split_plan(&a, SC_TILE, obj_n * (int)cdiv(hw, pair ? 2 * QT : QT), 2, pair ? num_sms() / 2 : num_sms(), 0, cell);
    *pieces_dev_out = cell;
  }
  // Q hi/lo of q * log2(e)/sqrt(d): logits land in the log2 domain; the pad rows (up to a whole tile pair) are zeroed
  // by the same launch (a separate memset cost a stream gap per read; forming the operand inside the tensor kernels from
  // the fp32 tensor was tried and lost: 64 dependent-latency loads per thread at every item start, phase B + 4 %)
  {
    PrepJob jb{q_in_dm, DK, hw, nullptr, nullptr, const_cast<uint16_t*>(a.qh), const_cast<uint16_t*>(a.ql),
               LOG2E / sqrtf((float)DK), 0, q_em, tc_operand_rows(hw)};
    if (int rc = launch_prep(&jb, 1, st)) return rc;
  }
  double work = 0;
  for (int o = 0; o < obj_n; ++o) work += 2.0 * DK * (double)banks[o].n * (double)hw;
  prof_begin(PROF_READ_A, st);
  if (pair)
    VFN_CUDA_OK(launch_pdl(tc_scan_pair_kernel<MODE_LSE>, dim3(num_sms()), dim3(TC_THREADS), SP_SMEM, st, maps, a, part));
  else
    tc_scan_kernel<MODE_LSE><<<num_sms(), TC_THREADS, SC_SMEM, st>>>(maps, a, part);
  prof_end(PROF_READ_A, st, work);
  VFN_LAUNCH_OK();
  count_launches(1);      // + 1 counted by launch_prep
  return VFN_OK;
}

int tc_phase_b(const vfn_bank* banks, int obj_n, const float* q_in_dm, int q_em, int64_t hw, int split_b,
               const float* lse, float thres_valid, int update_bank, float* po, char* ws_tc, cudaStream_t st,
               int* pieces_out, const int32_t** pieces_dev_out) {
  if (int rc = set_attrs()) return rc;
  bool live;
  if (int rc = live_mode(banks, obj_n, &live)) return rc;
  TcMaps maps;
  TcArgs a = {};
  const bool pair = (g_pair & 1) && (num_sms() % 2 == 0);
  const int tile = B_TILE;
  int64_t tmin = INT64_MAX, tmax = 0;
  for (int o = 0; o < obj_n; ++o) {
    const int64_t t = cdiv(banks[o].n, tile), tl = cdiv(n_low(banks[o]), tile);
    tmin = tl < tmin ? tl : tmin;
    tmax = t > tmax ? t : tmax;
  }
  const int combos = pair ? obj_n * 2 * (int)cdiv(hw, 2 * QT) : obj_n * 2 * (int)cdiv(hw, QT);
  // The tensor core accumulates with truncation (a systematic -2^-25 relative per accumulation step, measured through
  // the match scores and through constant-value banks in tests/test_gpu_fullsize.py): one TMEM accumulator must not
  // run over more than B_CHAIN_MAX slots (512 + 256 accumulation steps -> bias < 5e-5 relative); longer banks are cut
  // into at least that many pieces, whose partials are added in fp32 round-to-nearest by combine_out_kernel.
  const int s_min = (int)cdiv(tmax * tile, B_CHAIN_MAX);
  const int pieces = best_split(combos, tmin, tmax, 4, pair ? num_sms() / 2 : num_sms(), s_min);
  if (pieces > split_b) { set_error("phase B: split %d exceeds workspace bound %d", pieces, split_b); return VFN_E_CAPACITY; }
  *pieces_out = pieces;
  if (int rc = fill_args(banks, obj_n, hw, pieces, tile, ws_tc, &maps, &a, true, pair ? 32 : B_TILE)) return rc;
  (void)q_in_dm; (void)q_em;      // the operand arrays were left in ws_tc by phase A of the same read
  *pieces_dev_out = nullptr;
  if (live) {
    int32_t* cell = tc_plan_cell(ws_tc, hw) + 1;
    set_split_plan(&a, tile, combos, 4, pair ? num_sms() / 2 : num_sms(), B_CHAIN_MAX, cell);
    *pieces_dev_out = cell;
  }
  double work = 0;
  for (int o = 0; o < obj_n; ++o) work += 2.0 * DV * (double)banks[o].n * (double)hw;
  prof_begin(PROF_READ_B, st);
  if (pair)
    VFN_CUDA_OK(launch_pdl(tc_phase_b_pair_kernel, dim3(num_sms()), dim3(TC_THREADS), P_SMEM, st, maps, a, lse, thres_valid, update_bank, po));
  else
    tc_phase_b_kernel<<<num_sms(), TC_THREADS, B_SMEM, st>>>(maps, a, lse, thres_valid, update_bank, po);
  prof_end(PROF_READ_B, st, work);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

// [candidate partials | per object: cand hi, cand lo | 256 B: device-chosen split]
size_t tc_match_workspace_bytes(int obj_n, int64_t hw) {
  return align_up((size_t)obj_n * TC_MAX_SPLIT * hw * MATCH_CAND * sizeof(float2), 256) + 2 * (size_t)obj_n * a_operand_bytes(hw) + 256;
}

// fp32 (hw, 128) entry-major normalised candidates -> fp16 hi/lo of 16x (single-object API path)
__global__ void split_rows_kernel(const float* __restrict__ src, int64_t n, float scale, uint16_t* __restrict__ hi,
                                  uint16_t* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint16_t h, l;
  split_f16(src[i] * scale, h, l);
  hi[i] = h;
  lo[i] = l;
}

// cand_split != 0: the fp16 hi/lo of 16 * normalised candidates are already in the workspace (written by the fused
// preparation kernel at tc_match_cand_hi/lo); otherwise they are derived here from nck.
uint16_t* tc_match_cand_hi(char* ws, int obj_n, int64_t hw, int obj) {
  return reinterpret_cast<uint16_t*>(ws + align_up((size_t)obj_n * TC_MAX_SPLIT * hw * MATCH_CAND * sizeof(float2), 256) +
                                     (size_t)obj * a_operand_bytes(hw));
}
uint16_t* tc_match_cand_lo(char* ws, int obj_n, int64_t hw, int obj) {
  return tc_match_cand_hi(ws, obj_n, hw, obj_n) + (size_t)obj * a_operand_bytes(hw) / sizeof(uint16_t);
}

int tc_match(const vfn_bank* banks, int obj_n, const float* const* nck_em, int64_t hw, char* ws, int cand_split,
             int32_t* const* idx_out, float* const* corr_out, cudaStream_t st) {
  if (int rc = set_attrs()) return rc;
  VFN_CHECK_ARG(obj_n >= 1 && obj_n <= TC_MAX_OBJ, "tcgen05 match supports at most %d objects", TC_MAX_OBJ);
  TcMaps maps;
  TcArgs a = {};
  RescoreArgs r;
  int64_t tmin = INT64_MAX, tmax = 0;
  double work = 0;
  const bool pair = (g_pair & 2) && (num_sms() % 2 == 0);
  for (int o = 0; o < obj_n; ++o) {
    VFN_CHECK_ARG(banks[o].d_key == DK && banks[o].n < (1ll << 31) && banks[o].nkh, "tcgen05 match needs d_key = 128");
    VFN_CHECK_ARG(!banks[o].n_live || (banks[o].n_min >= 1 && banks[o].n_min <= banks[o].n), "bank %d: bad n_min", o);
    const int64_t t = cdiv(banks[o].n, SC_TILE), tl = cdiv(n_low(banks[o]), SC_TILE);
    tmin = tl < tmin ? tl : tmin;
    tmax = t > tmax ? t : tmax;
    if (int rc = make_map(&maps.kh[o], banks[o].nkh, map_rows(banks[o]), DK, pair ? 64 : SC_TILE, 2)) return rc;
    if (int rc = make_map(&maps.kl[o], banks[o].nkl, map_rows(banks[o]), DK, pair ? 64 : SC_TILE, 2)) return rc;
    a.n[o] = (int)banks[o].n; a.tiles[o] = (int)t; a.cnt[o] = nullptr; a.n_live[o] = banks[o].n_live;
    r.n_live[o] = banks[o].n_live;
    r.n[o] = (int)banks[o].n; r.nk[o] = banks[o].nk; r.nck[o] = nck_em[o]; r.idx_out[o] = idx_out[o]; r.corr_out[o] = corr_out[o];
    work += 2.0 * DK * (double)banks[o].n * (double)hw;
  }
  for (int o = obj_n; o < TC_MAX_OBJ; ++o) { a.n_live[o] = nullptr; r.n_live[o] = nullptr; }
  const int pieces = pair ? best_split(obj_n * (int)cdiv(hw, 2 * QT), tmin, tmax, 2, num_sms() / 2)
                          : best_split(obj_n * (int)cdiv(hw, QT), tmin, tmax, 2);
  bool live;
  if (int rc = live_mode(banks, obj_n, &live)) return rc;
  a.obj_n = obj_n; a.hw = (int)hw; a.q_tiles = (int)cdiv(hw, QT); a.pieces = pieces;
  a.pieces_dev = nullptr;
  r.pieces_dev = nullptr;
  if (live) {
    int32_t* cell = reinterpret_cast<int32_t*>(ws + tc_match_workspace_bytes(obj_n, hw) - 256);
    set_split_plan(&a, SC_TILE, obj_n * (int)cdiv(hw, pair ? 2 * QT : QT), 2, pair ? num_sms() / 2 : num_sms(), 0, cell);
    r.pieces_dev = cell;
  }
  a.qh = tc_match_cand_hi(ws, obj_n, hw, 0);
  a.ql = tc_match_cand_lo(ws, obj_n, hw, 0);
  a.a_obj_stride = (long long)(a_operand_bytes(hw) / sizeof(uint16_t));
  a.band = MATCH_BAND * NK_SCALE * NK_SCALE;
  a.dbg = g_dbg;
  a.tstamp = nullptr;
  if (!cand_split) {
    VFN_CUDA_OK(cudaMemsetAsync(const_cast<uint16_t*>(a.qh), 0, 2 * (size_t)obj_n * a_operand_bytes(hw), st));
    for (int o = 0; o < obj_n; ++o) {
      const int64_t ne = hw * DK;
      split_rows_kernel<<<(unsigned)cdiv(ne, 256), 256, 0, st>>>(nck_em[o], ne, NK_SCALE, tc_match_cand_hi(ws, obj_n, hw, o),
                                                                tc_match_cand_lo(ws, obj_n, hw, o));
    }
    count_launches(obj_n);
  }
  float2* part = reinterpret_cast<float2*>(ws);
  prof_begin(PROF_MATCH, st);
  if (pair)
    VFN_CUDA_OK(launch_pdl(tc_scan_pair_kernel<MODE_MATCH>, dim3(num_sms()), dim3(TC_THREADS), SP_SMEM, st, maps, a, part));
  else
    tc_scan_kernel<MODE_MATCH><<<num_sms(), TC_THREADS, SC_SMEM, st>>>(maps, a, part);
  prof_end(PROF_MATCH, st, work);
  r.obj_n = obj_n; r.pieces = pieces; r.hw = (int)hw; r.band = a.band;
  dim3 grid((unsigned)cdiv(hw, 8), obj_n);
  VFN_CUDA_OK(launch_pdl(match_rescore_kernel, grid, dim3(256), 0, st, (const float2*)part, r));
  VFN_LAUNCH_OK();
  count_launches(2);
  return VFN_OK;
}

}  // namespace vfn

extern "C" int vfn_debug_set_pair(int32_t mask) {
  vfn::g_pair = mask & 7;
  return VFN_OK;
}

extern "C" int vfn_debug_set_tstamp(long long* d_ptr) {
  vfn::g_tstamp = d_ptr;
  return VFN_OK;
}

extern "C" int vfn_debug_set_dump(float* d_ptr) {
  vfn::g_dbg = d_ptr;
  return VFN_OK;
}
