// tcgen05 / TMEM / TMA implementation of the memory read (Matcher.forward, AFB_URR.py:136-178) for sm_100a.
//
//   phase A  S^T[j,i] = <q_j, k_i> * log2(e)/sqrt(128)   -> per-query running (max, sum 2^(s-max)) over the slots
//   phase B  S^T recomputed, P = 2^(S - lse2_j) (already normalised: no online rescale), usage counts
//            cnt_i += [P_ij > thres], O^T[j,c] += sum_i P_ij V_ic with the fp32 accumulator resident in TMEM.
//
// Orientation (both phases): TMEM lanes = queries (M = 128), columns = bank slots (S) / value channels (O).
// A operands come from TMEM (Q for the S-MMA, P for the O-MMA), B operands from shared memory via TMA:
//   K tiles  K-major  [slots x 128 d]   bf16, 128B swizzle, two 64-d boxes per piece
//   V tiles  MN-major [slots x 256 ch]  bf16, 128B swizzle, four 64-channel boxes per piece
// Precision: operands are bf16 hi+lo splits (x ~ hi + lo, 16 mantissa bits); every product uses
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (3 MMA passes), which keeps logits to ~2e-5 and the
// readout far inside the 1e-3 tolerance (plain bf16 would give ~7e-3).
// Work split: persistent CTAs (one per SM), static stream-K partition of the (object, query tile[, channel half])
// x slot-tile space, so every CTA gets the same number of tile units; per-CTA partials are combined by
// lse_combine_kernel / combine_out_kernel (vfn_simt.cu) in a fixed order (deterministic).
#include "vfn_tc.cuh"

#include <cuda.h>

namespace vfn {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a broken pipeline traps (launch error) instead of hanging the GPU
  uint32_t done = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, 128B swizzle, Blackwell version bit
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor kind::f16: bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// instruction descriptor kind::tf32: tf32 x tf32 -> f32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int DK = 128, DV = 512;
constexpr int QT = 128;                 // queries per tile (TMEM lanes)
constexpr int TC_THREADS = 384;         // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 softmax WG0, 8-11 softmax WG1
constexpr int TC_MAX_OBJ = 4;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// TMEM columns
constexpr uint32_t TM_O = 0;            // phase B: O^T accumulator, 256 columns
constexpr uint32_t TM_QH = 256, TM_QL = 320;   // Q hi / lo as A operand: 64 columns each (128 bf16)
constexpr uint32_t TM_S = 384;          // phase B: 2 buffers x 64 columns ; phase A uses [0,256) as 2 x 128

struct TcMaps {
  CUtensorMap kh[TC_MAX_OBJ], kl[TC_MAX_OBJ], vh[TC_MAX_OBJ], vl[TC_MAX_OBJ];
};
struct TcArgs {
  int obj_n, hw, q_tiles, pieces;       // pieces = partial slots per combo (P_MAX)
  int n[TC_MAX_OBJ];                    // live slots per object
  int tiles[TC_MAX_OBJ];                // slot tiles per object for this phase
  const uint16_t* qh;                   // (q_tiles*128, 128) bf16 hi of q * log2e/sqrt(d)
  const uint16_t* ql;
  int32_t* cnt[TC_MAX_OBJ];
  float* dbg;
};


// ------------------------------------------------------------------------------------------------
// phase A
// ------------------------------------------------------------------------------------------------
constexpr int A_TILE = 128;             // slots per S tile
constexpr int A_STAGES = 3;
constexpr int A_STAGE_BYTES = A_TILE * DK * 2 * 2;   // hi + lo = 64 KB
constexpr int A_SMEM = A_STAGES * A_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(TC_THREADS, 1) tc_phase_a_kernel(const __grid_constant__ TcMaps maps, TcArgs args,
                                                                   float2* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;                                           // A_STAGES x 64 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_STAGES * A_STAGE_BYTES);
  uint64_t* k_full = bars;                 // [A_STAGES]
  uint64_t* k_empty = bars + A_STAGES;     // [A_STAGES]
  uint64_t* s_full = bars + 2 * A_STAGES;  // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(s_empty + 2);
  __shared__ float2 ml_x[QT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_p, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;

  // work items = (split, object, query tile), dealt round-robin to the persistent CTAs: in any round all CTAs
  // stream the same few slot ranges, so the K tiles are served from L2 (ncu: 14x DRAM re-reads with a per-CTA
  // contiguous partition, profiles/r1_*).
  const int n_combos = args.obj_n * args.q_tiles;
  const int n_items = n_combos * args.pieces;
  uint32_t k_it = 0;            // tiles streamed so far (producer & MMA agree)
  uint32_t buf_it[2] = {0, 0};  // uses of each S buffer so far
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int piece = item / n_combos;
    const int combo = item - piece * n_combos;
    const int obj = combo / args.q_tiles;
    const int qt = combo - obj * args.q_tiles;
    const int tiles_o = args.tiles[obj];
    const int t0 = (int)((long long)tiles_o * piece / args.pieces);
    const int t1 = (int)((long long)tiles_o * (piece + 1) / args.pieces);
    const int n_obj = args.n[obj];
    const int ntile = t1 - t0;
    const bool first_item = (item == (int)blockIdx.x);

    // (1) Q tile -> TMEM (WG0: hi, WG1: lo); one row (query) per thread
    if (warp >= 4) {
      const int wg = (warp - 4) >> 2;
      const int row = ((warp & 3) << 5) + lane;
      const uint16_t* src = (wg == 0 ? args.qh : args.ql) + ((size_t)qt * QT + row) * DK;
      const uint32_t tbase = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (wg == 0 ? TM_QH : TM_QL);
      uint32_t v[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = reinterpret_cast<const uint4*>(src)[h * 8 + i];
          v[4 * i + 0] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
        }
        tmem_st32(tbase + h * 32, v);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // (2) roles
    if (warp == 0) {
      for (int t = 0; t < ntile; ++t, ++k_it) {
        const uint32_t st = k_it % A_STAGES, ph = (k_it / A_STAGES) & 1;
        if (lane == 0) {
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], A_STAGE_BYTES);
          uint8_t* dst = kst + st * A_STAGE_BYTES;
          const int row0 = (t0 + t) * A_TILE;
          tma_load_2d(dst, &maps.kh[obj], &k_full[st], 0, row0);
          tma_load_2d(dst + 16384, &maps.kh[obj], &k_full[st], 64, row0);
          tma_load_2d(dst + 32768, &maps.kl[obj], &k_full[st], 0, row0);
          tma_load_2d(dst + 49152, &maps.kl[obj], &k_full[st], 64, row0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc(128, A_TILE, 0, 0);
      for (int t = 0; t < ntile; ++t, ++k_it) {
        const uint32_t st = k_it % A_STAGES, ph = (k_it / A_STAGES) & 1;
        const int b = t & 1;
        if (lane == 0) {
          mbar_wait(&s_empty[b], (buf_it[b] & 1) ^ 1);
          mbar_wait(&k_full[st], ph);
          tc_fence_after();
          const uint32_t kbase = smem_u32(kst + st * A_STAGE_BYTES);
          const uint32_t d_t = tmem + (uint32_t)b * A_TILE;
          // passes: (Qh,Kh) (Ql,Kh) (Qh,Kl)
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_col = (pass == 1) ? TM_QL : TM_QH;
            const uint32_t kb = kbase + ((pass == 2) ? 32768u : 0u);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t bd = make_sdesc(kb + (ks >> 2) * 16384u + (ks & 3) * 32u, 16, 1024);
              mma_ts(d_t, tmem + a_col + ks * 8, bd, idesc, (pass | ks) ? 1u : 0u);
            }
          }
          tc_commit(&k_empty[st]);
          tc_commit(&s_full[b]);
        }
        __syncwarp();
        ++buf_it[b];
      }
    } else if (warp >= 4) {
      const int wg = (warp - 4) >> 2;
      const int row = ((warp & 3) << 5) + lane;
      const uint32_t tlane = tmem + (((uint32_t)(warp & 3) * 32u) << 16);
      float m_run = -INFINITY, l_run = 0.f;
      for (int t = wg; t < ntile; t += 2) {
        mbar_wait(&s_full[wg], buf_it[wg] & 1);
        ++buf_it[wg];
        tc_fence_after();
        const int slot0 = (t0 + t) * A_TILE;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld32(tlane + (uint32_t)wg * A_TILE + ch * 32, v);
          tmem_wait_ld();
          if (args.dbg && blockIdx.x == 0 && first_item && t == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) args.dbg[row * A_TILE + ch * 32 + i] = __uint_as_float(v[i]);
          }
          float cm = -INFINITY;
          const int lim = n_obj - (slot0 + ch * 32);     // valid slots in this chunk
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float s = __uint_as_float(v[i]);
            s = (i < lim) ? s : -INFINITY;
            v[i] = __float_as_uint(s);
            cm = fmaxf(cm, s);
          }
          const float m_new = fmaxf(m_run, cm);
          if (m_new > -INFINITY) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(v[i]) - m_new);
            l_run = l_run * ex2(m_run - m_new) + acc;
            m_run = m_new;
          }
        }
        tc_fence_before();
        mbar_arrive(&s_empty[wg]);
      }
      // (3) combine the two warpgroups' statistics and publish the piece
      if (wg == 1) ml_x[row] = make_float2(m_run, l_run);
      named_bar_sync(1, 256);
      if (wg == 0) {
        const float2 o = ml_x[row];
        const float m = fmaxf(m_run, o.x);
        float l = 0.f;
        if (m > -INFINITY) l = l_run * ex2(m_run - m) + o.y * ex2(o.x - m);
        const int j = qt * QT + row;
        if (j < args.hw) part[((size_t)obj * args.pieces + piece) * args.hw + j] = make_float2(m * LN2, l);
      }
    } else {
      // warps 2,3 idle during the tile loop
    }
    // keep role-local counters consistent for warps that did not run the loops
    if (warp != 0 && warp != 1) k_it += ntile;
    if (warp != 1 && warp < 4) { buf_it[0] += (ntile + 1) >> 1; buf_it[1] += ntile >> 1; }
    if (warp >= 4) { const int wgx = (warp - 4) >> 2; buf_it[wgx ^ 1] += (wgx ^ 1) == 0 ? (ntile + 1) >> 1 : ntile >> 1; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// phase B
// ------------------------------------------------------------------------------------------------
constexpr int B_TILE = 64;                              // slots per tile
constexpr int B_KSTAGES = 2, B_VSTAGES = 2;
constexpr int B_KSTAGE_BYTES = B_TILE * DK * 2 * 2;     // 32 KB (hi + lo)
constexpr int B_VSTAGE_BYTES = B_TILE * 256 * 2 * 2;    // 64 KB (hi + lo, 256 channels)
constexpr int B_SMEM = B_KSTAGES * B_KSTAGE_BYTES + B_VSTAGES * B_VSTAGE_BYTES + 1024 + 1024;

__global__ void __launch_bounds__(TC_THREADS, 1) tc_phase_b_kernel(const __grid_constant__ TcMaps maps, TcArgs args,
                                                                   const float* __restrict__ lse, float thres,
                                                                   int do_count, float* __restrict__ po) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;
  uint8_t* vst = smem + B_KSTAGES * B_KSTAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(vst + B_VSTAGES * B_VSTAGE_BYTES);
  uint64_t* k_full = bars;            // [2]
  uint64_t* k_empty = bars + 2;       // [2]
  uint64_t* v_full = bars + 4;        // [2]
  uint64_t* v_empty = bars + 6;       // [2]
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* p_full = bars + 10;       // [2]
  uint64_t* o_full = bars + 12;       // [1]
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + 13);
  int* cnt_s = reinterpret_cast<int*>(bars + 16);   // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 256);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 128) cnt_s[threadIdx.x] = 0;
  if (warp == 2) tmem_alloc(tmem_base_p, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;

  // work items = (split, object, query tile, channel half), round-robin over the persistent CTAs (see phase A)
  const int cpo = args.q_tiles * 2;
  const int n_combos = args.obj_n * cpo;
  const int n_items = n_combos * args.pieces;
  uint32_t k_it = 0;            // tiles streamed so far
  uint32_t buf_it[2] = {0, 0};  // uses of each S/P buffer
  uint32_t seg_it = 0;          // items finished (o_full phase)
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int piece = item / n_combos;
    const int combo = item - piece * n_combos;
    const int obj = combo / cpo;
    const int cidx = combo - obj * cpo;                // qt * 2 + half
    const int qt = cidx >> 1, half = cidx & 1;
    const int tiles_o = args.tiles[obj];
    const int t0 = (int)((long long)tiles_o * piece / args.pieces);
    const int t1 = (int)((long long)tiles_o * (piece + 1) / args.pieces);
    const int n_obj = args.n[obj];
    const int ntile = t1 - t0;
    const bool first_item = (item == (int)blockIdx.x);

    if (warp >= 4) {
      const int wg = (warp - 4) >> 2;
      const int row = ((warp & 3) << 5) + lane;
      const uint16_t* src = (wg == 0 ? args.qh : args.ql) + ((size_t)qt * QT + row) * DK;
      const uint32_t tbase = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (wg == 0 ? TM_QH : TM_QL);
      uint32_t v[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = reinterpret_cast<const uint4*>(src)[h * 8 + i];
          v[4 * i + 0] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
        }
        tmem_st32(tbase + h * 32, v);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      for (int t = 0; t < ntile; ++t, ++k_it) {
        const uint32_t st = k_it & 1, ph = (k_it >> 1) & 1;
        if (lane == 0) {
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
          for (int g = 0; g < 4; ++g) {
            tma_load_2d(vd + g * 8192, &maps.vh[obj], &v_full[st], half * 256 + g * 64, row0);
            tma_load_2d(vd + 32768 + g * 8192, &maps.vl[obj], &v_full[st], half * 256 + g * 64, row0);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc_s = make_idesc(128, B_TILE, 0, 0);
      constexpr uint32_t idesc_o = make_idesc(128, 256, 0, 1);
      if (lane == 0) {
        // software pipeline: S(0); for t: { S(t+1); wait P(t); O(t) }
        auto issue_s = [&](int t, uint32_t kit) {
          const uint32_t st = kit & 1, ph = (kit >> 1) & 1;
          const int b = t & 1;
          mbar_wait(&k_full[st], ph);
          tc_fence_after();
          const uint32_t kbase = smem_u32(kst + st * B_KSTAGE_BYTES);
          const uint32_t d_t = tmem + TM_S + (uint32_t)b * 64;
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
        };
        if (ntile > 0) issue_s(0, k_it);
        for (int t = 0; t < ntile; ++t) {
          if (t + 1 < ntile) issue_s(t + 1, k_it + t + 1);
          const uint32_t kit = k_it + t;
          const uint32_t st = kit & 1, ph = (kit >> 1) & 1;
          const int b = t & 1;
          mbar_wait(&p_full[b], buf_it[b] & 1);
          ++buf_it[b];
          mbar_wait(&v_full[st], ph);
          tc_fence_after();
          const uint32_t vbase = smem_u32(vst + st * B_VSTAGE_BYTES);
          const uint32_t pcol = tmem + TM_S + (uint32_t)b * 64;
          // passes: (Ph,Vh) (Ph,Vl) (Pl,Vh).  P layout inside the 64-column buffer (written by the two softmax
          // warpgroups, 32 slots each): [Ph(0-31) | Pl(0-31) | Ph(32-63) | Pl(32-63)], 16 columns per block
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t lo_off = (pass == 2) ? 16u : 0u;
            const uint32_t vb = vbase + ((pass == 1) ? 32768u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t bd = make_sdesc(vb + ks * 2048u, 8192, 1024);
              const uint32_t a_col = pcol + (uint32_t)(ks >> 1) * 32u + lo_off + (uint32_t)(ks & 1) * 8u;
              mma_ts(tmem + TM_O, a_col, bd, idesc_o, (t | pass | ks) ? 1u : 0u);
            }
          }
          tc_commit(&v_empty[st]);
        }
        tc_commit(o_full);
      } else {
        for (int t = 0; t < ntile; ++t) ++buf_it[t & 1];
      }
      k_it += ntile;
      __syncwarp();
    } else if (warp >= 4) {
      const int wg = (warp - 4) >> 2;
      const int row = ((warp & 3) << 5) + lane;
      const int wtid = threadIdx.x - 128 - wg * 128;       // 0..127 within the warpgroup
      const uint32_t tlane = tmem + (((uint32_t)(warp & 3) * 32u) << 16);
      const int j = qt * QT + row;
      const float lse2 = (j < args.hw) ? lse[(size_t)obj * args.hw + j] * LOG2E : INFINITY;
      int* mycnt = cnt_s + wg * 32;
      const bool counting = do_count && (half == 0);
      // both warpgroups work on every tile: WG0 takes slots [0,32) of the tile, WG1 slots [32,64).  Halving the
      // per-tile softmax latency is what keeps the tensor pipe fed (S(t+1) is only 768 clk of cover).
      for (int t = 0; t < ntile; ++t) {
        const int b = t & 1;
        mbar_wait(&s_full[b], buf_it[b] & 1);
        ++buf_it[b];
        tc_fence_after();
        const int slot0 = (t0 + t) * B_TILE + wg * 32;
        const uint32_t sb = tlane + TM_S + (uint32_t)b * 64 + (uint32_t)wg * 32;
        uint32_t s0[32];
        tmem_ld32(sb, s0);
        tmem_wait_ld();
        if (args.dbg && blockIdx.x == 0 && first_item && t == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) args.dbg[row * B_TILE + wg * 32 + i] = __uint_as_float(s0[i]);
        }
        const int lim = n_obj - slot0;
        uint32_t hi_w[16], lo_w[16];
        uint32_t bits = 0u;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2(__uint_as_float(s0[i]) - lse2), p1 = ex2(__uint_as_float(s0[i + 1]) - lse2);
          p0 = (i < lim) ? p0 : 0.f;
          p1 = (i + 1 < lim) ? p1 : 0.f;
          bits |= (uint32_t)(p0 > thres) << i;
          bits |= (uint32_t)(p1 > thres) << (i + 1);
          const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          const __nv_bfloat162 l = __floats2bfloat162_rn(p0 - __low2float(h), p1 - __high2float(h));
          hi_w[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          lo_w[i >> 1] = *reinterpret_cast<const uint32_t*>(&l);
        }
        tmem_st16(sb, hi_w);          // P hi: 32 bf16 = 16 columns
        tmem_st16(sb + 16, lo_w);     // P lo
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[b]);
        if (counting) {
          if (j >= args.hw) bits = 0u;
          const int dense = __any_sync(0xffffffffu, __popc(bits) > 4);
          if (dense) {
            int c_mine = 0;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) {
              const unsigned bal = __ballot_sync(0xffffffffu, (bits >> c) & 1u);
              if (lane == c) c_mine = __popc(bal);
            }
            if (c_mine) atomicAdd(&mycnt[lane], c_mine);
          } else {
            while (bits) {
              const int c = __ffs((int)bits) - 1;
              bits &= bits - 1;
              atomicAdd(&mycnt[c], 1);
            }
          }
          named_bar_sync(1 + wg, 128);
          if (wtid < 32) {
            const int c = mycnt[wtid];
            if (c) {
              atomicAdd(&args.cnt[obj][slot0 + wtid], c);
              mycnt[wtid] = 0;
            }
          }
          named_bar_sync(1 + wg, 128);
        }
      }
      // epilogue: O^T (128 queries x 256 channels) -> partial buffer, WG0 channels [0,128), WG1 [128,256)
      mbar_wait(o_full, seg_it & 1);
      tc_fence_after();
      if (ntile > 0) {
        float* dst = po + (((size_t)obj * args.pieces + piece) * DV + half * 256 + wg * 128) * (size_t)args.hw;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld32(tlane + TM_O + (uint32_t)wg * 128 + ch * 32, v);
          tmem_wait_ld();
          if (j < args.hw) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)(ch * 32 + i) * args.hw + j] = __uint_as_float(v[i]);
          }
        }
      }
    }
    if (warp != 0 && warp != 1) k_it += ntile;
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
// cosine match on the tensor cores (FeatureBank.py:63-68): corr = <nk_i, nck_j> as 3xTF32
// (hi*hi + lo*hi + hi*lo, fp32 accumulate: ~2^-21 relative, i.e. fp32-GEMM-grade), per-query arg-max with
// ties -> lowest slot.  Same pipeline as phase A; lanes = candidates (queries), columns = bank slots.
// ------------------------------------------------------------------------------------------------
constexpr int M_TILE = 64;
constexpr int M_STAGES = 3;
constexpr int M_STAGE_BYTES = M_TILE * DK * 4 * 2;      // fp32 hi + lo = 64 KB
constexpr int M_SMEM = M_STAGES * M_STAGE_BYTES + 1024 + 256;
constexpr uint32_t TMM_QH = 0, TMM_QL = 128, TMM_S = 256;

constexpr int MATCH_TOPK = 4;   // approximate candidates kept per (piece, query) for the exact fp32 re-score

// insert (v, i) into a descending top-4 list; the caller scans slots in ascending order, so strict '>' keeps the
// lowest slot first among equal values
__device__ __forceinline__ void top4_insert(float (&tv)[MATCH_TOPK], int (&ti)[MATCH_TOPK], float v, int i) {
  if (v > tv[3]) {
    tv[3] = v; ti[3] = i;
#pragma unroll
    for (int r = 3; r > 0; --r) {
      if (tv[r] > tv[r - 1]) {
        const float fv = tv[r]; tv[r] = tv[r - 1]; tv[r - 1] = fv;
        const int fi = ti[r]; ti[r] = ti[r - 1]; ti[r - 1] = fi;
      }
    }
  }
}
// general insert with explicit (value desc, slot asc) order, for merging lists that were not scanned in slot order
__device__ __forceinline__ void top4_merge(float (&tv)[MATCH_TOPK], int (&ti)[MATCH_TOPK], float v, int i) {
  if (v > tv[3] || (v == tv[3] && i < ti[3])) {
    tv[3] = v; ti[3] = i;
#pragma unroll
    for (int r = 3; r > 0; --r) {
      if (tv[r] > tv[r - 1] || (tv[r] == tv[r - 1] && ti[r] < ti[r - 1])) {
        const float fv = tv[r]; tv[r] = tv[r - 1]; tv[r - 1] = fv;
        const int fi = ti[r]; ti[r] = ti[r - 1]; ti[r - 1] = fi;
      }
    }
  }
}

struct MatchArgs {
  int hw, q_tiles, pieces, n, tiles;
  const float* nck;        // (hw, 128) normalised candidates, entry-major
  float* dbg;
};

__global__ void __launch_bounds__(TC_THREADS, 1) tc_match_kernel(const __grid_constant__ CUtensorMap map_h,
                                                                 const __grid_constant__ CUtensorMap map_l,
                                                                 MatchArgs args, float2* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kst = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M_STAGES * M_STAGE_BYTES);
  uint64_t* k_full = bars;
  uint64_t* k_empty = bars + M_STAGES;
  uint64_t* s_full = bars + 2 * M_STAGES;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(s_empty + 2);
  __shared__ float2 best_x[QT][MATCH_TOPK];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < M_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_p, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;

  const int n_items = args.q_tiles * args.pieces;
  uint32_t k_it = 0;
  uint32_t buf_it[2] = {0, 0};
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int piece = item / args.q_tiles;
    const int qt = item - piece * args.q_tiles;
    const int t0 = (int)((long long)args.tiles * piece / args.pieces);
    const int t1 = (int)((long long)args.tiles * (piece + 1) / args.pieces);
    const int ntile = t1 - t0;

    // (1) candidate tile -> TMEM as tf32 hi (WG0) / lo (WG1); one candidate per lane, 128 columns each
    if (warp >= 4) {
      const int wg = (warp - 4) >> 2;
      const int row = ((warp & 3) << 5) + lane;
      const int j = qt * QT + row;
      const float* src = args.nck + (size_t)j * DK;
      const uint32_t tbase = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (wg == 0 ? TMM_QH : TMM_QL);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < args.hw) x = reinterpret_cast<const float4*>(src)[ch * 8 + i];
          float h, l;
          split_tf32(x.x, h, l); v[4 * i + 0] = __float_as_uint(wg == 0 ? h : l);
          split_tf32(x.y, h, l); v[4 * i + 1] = __float_as_uint(wg == 0 ? h : l);
          split_tf32(x.z, h, l); v[4 * i + 2] = __float_as_uint(wg == 0 ? h : l);
          split_tf32(x.w, h, l); v[4 * i + 3] = __float_as_uint(wg == 0 ? h : l);
        }
        tmem_st32(tbase + ch * 32, v);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      for (int t = 0; t < ntile; ++t, ++k_it) {
        const uint32_t st = k_it % M_STAGES, ph = (k_it / M_STAGES) & 1;
        if (lane == 0) {
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], M_STAGE_BYTES);
          uint8_t* dst = kst + st * M_STAGE_BYTES;
          const int row0 = (t0 + t) * M_TILE;
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            tma_load_2d(dst + blk * 8192, &map_h, &k_full[st], blk * 32, row0);
            tma_load_2d(dst + 32768 + blk * 8192, &map_l, &k_full[st], blk * 32, row0);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc_tf32(128, M_TILE);
      for (int t = 0; t < ntile; ++t, ++k_it) {
        const uint32_t st = k_it % M_STAGES, ph = (k_it / M_STAGES) & 1;
        const int b = t & 1;
        if (lane == 0) {
          mbar_wait(&s_empty[b], (buf_it[b] & 1) ^ 1);
          mbar_wait(&k_full[st], ph);
          tc_fence_after();
          const uint32_t kbase = smem_u32(kst + st * M_STAGE_BYTES);
          const uint32_t d_t = tmem + TMM_S + (uint32_t)b * M_TILE;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_col = (pass == 1) ? TMM_QL : TMM_QH;
            const uint32_t kb = kbase + ((pass == 2) ? 32768u : 0u);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
              const uint64_t bd = make_sdesc(kb + (ks >> 2) * 8192u + (ks & 3) * 32u, 16, 1024);
              mma_ts_tf32(d_t, tmem + a_col + ks * 8, bd, idesc, (pass | ks) ? 1u : 0u);
            }
          }
          tc_commit(&k_empty[st]);
          tc_commit(&s_full[b]);
        }
        __syncwarp();
        ++buf_it[b];
      }
    } else if (warp >= 4) {
      const int wg = (warp - 4) >> 2;
      const int row = ((warp & 3) << 5) + lane;
      const uint32_t tlane = tmem + (((uint32_t)(warp & 3) * 32u) << 16);
      float tv[MATCH_TOPK];
      int ti[MATCH_TOPK];
#pragma unroll
      for (int r = 0; r < MATCH_TOPK; ++r) { tv[r] = -INFINITY; ti[r] = 0x7fffffff; }
      for (int t = wg; t < ntile; t += 2) {
        mbar_wait(&s_full[wg], buf_it[wg] & 1);
        ++buf_it[wg];
        tc_fence_after();
        const int slot0 = (t0 + t) * M_TILE;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(tlane + TMM_S + (uint32_t)wg * M_TILE + ch * 32, v);
          tmem_wait_ld();
          if (args.dbg && blockIdx.x == 0 && item == (int)blockIdx.x && t == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) args.dbg[row * M_TILE + ch * 32 + i] = __uint_as_float(v[i]);
          }
          const int lim = args.n - (slot0 + ch * 32);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float sv = __uint_as_float(v[i]);
            if (i < lim) top4_insert(tv, ti, sv, slot0 + ch * 32 + i);
          }
        }
        tc_fence_before();
        mbar_arrive(&s_empty[wg]);
      }
      if (wg == 1) {
#pragma unroll
        for (int r = 0; r < MATCH_TOPK; ++r) best_x[row][r] = make_float2(tv[r], __int_as_float(ti[r]));
      }
      named_bar_sync(1, 256);
      if (wg == 0) {
#pragma unroll
        for (int r = 0; r < MATCH_TOPK; ++r) {
          const float2 o = best_x[row][r];
          top4_merge(tv, ti, o.x, __float_as_int(o.y));
        }
        const int j = qt * QT + row;
        if (j < args.hw) {
#pragma unroll
          for (int r = 0; r < MATCH_TOPK; ++r)
            part[((size_t)piece * args.hw + j) * MATCH_TOPK + r] = make_float2(tv[r], __int_as_float(ti[r]));
        }
      }
    }
    if (warp != 0 && warp != 1) k_it += ntile;
    if (warp != 1 && warp < 4) { buf_it[0] += (ntile + 1) >> 1; buf_it[1] += ntile >> 1; }
    if (warp >= 4) { const int wgx = (warp - 4) >> 2; buf_it[wgx ^ 1] += (wgx ^ 1) == 0 ? (ntile + 1) >> 1 : ntile >> 1; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// Exact fp32 re-score of the tensor-core candidates: one warp per candidate query.  The 3xTF32 scores carry the
// tensor core's truncating accumulation (measured: ~4e-6 systematic bias), so every slot whose approximate score is
// within MATCH_BAND of the approximate maximum is re-evaluated with the sequential fp32 FMA chain the SIMT kernel uses
// (bit-identical values and ordering).  If some piece's whole top-4 lies inside the band there may be hidden
// candidates: that query falls back to an exact scan of all slots (rare: needs >= 4 near-duplicate slots).
constexpr float MATCH_BAND = 2e-5f;

__device__ __forceinline__ float exact_dot128(const float* __restrict__ nkh, const float* __restrict__ nkl,
                                              const float* __restrict__ q, int64_t slot) {
  const float4* h = reinterpret_cast<const float4*>(nkh + slot * DK);
  const float4* l = reinterpret_cast<const float4*>(nkl + slot * DK);
  const float4* b = reinterpret_cast<const float4*>(q);
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < DK / 4; ++k) {
    const float4 hv = h[k], lv = l[k], bv = b[k];
    acc = fmaf(hv.x + lv.x, bv.x, acc);
    acc = fmaf(hv.y + lv.y, bv.y, acc);
    acc = fmaf(hv.z + lv.z, bv.z, acc);
    acc = fmaf(hv.w + lv.w, bv.w, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(256) match_rescore_kernel(const float2* __restrict__ part, int pieces, int hw, int n,
                                                            const float* __restrict__ nkh,
                                                            const float* __restrict__ nkl,
                                                            const float* __restrict__ nck,
                                                            int32_t* __restrict__ idx_out,
                                                            float* __restrict__ corr_out) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= hw) return;
  const int n_cand = pieces * MATCH_TOPK;
  const float* q = nck + (size_t)j * DK;
  // approximate maximum over all pieces
  float amax = -INFINITY;
  for (int c = lane; c < n_cand; c += 32) {
    const int pc = c / MATCH_TOPK, r = c - pc * MATCH_TOPK;
    amax = fmaxf(amax, part[((size_t)pc * hw + j) * MATCH_TOPK + r].x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float band = amax - MATCH_BAND;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  int overflow = 0;
  for (int c = lane; c < n_cand; c += 32) {
    const int pc = c / MATCH_TOPK, r = c - pc * MATCH_TOPK;
    const float2 e = part[((size_t)pc * hw + j) * MATCH_TOPK + r];
    const int slot = __float_as_int(e.y);
    if (e.x >= band && slot < n) {
      if (r == MATCH_TOPK - 1) overflow = 1;
      const float v = exact_dot128(nkh, nkl, q, slot);
      if (v > best || (v == best && slot < bidx)) { best = v; bidx = slot; }
    }
  }
  overflow = __any_sync(0xffffffffu, overflow);
  if (overflow) {
    best = -INFINITY;
    bidx = 0x7fffffff;
    for (int slot = lane; slot < n; slot += 32) {
      const float v = exact_dot128(nkh, nkl, q, slot);
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
    idx_out[j] = bidx;
    corr_out[j] = best;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D tensor (rows, cols) row-major, bf16 (elem_bytes 2) or fp32 (4); box = (box_rows, 128 B of columns), 128B swizzle,
// OOB rows -> 0
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int cols, int box_rows, int elem_bytes = 2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return VFN_E_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return VFN_E_CUDA; }
  return VFN_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static float* g_dbg = nullptr;

bool tc_shapes_ok(int d_key, int d_val) { return d_key == DK && d_val == DV; }

constexpr int TC_MAX_SPLIT = 24;

// number of slot splits per (object, query tile[, half]) combo: minimise rounds x (tiles per item + fixed per-item
// overhead) for `combos` combos dealt round-robin to G persistent CTAs; every item keeps at least one tile.
static int best_split(int combos, int64_t tiles_min, int64_t tiles_max, int overhead_tiles) {
  const int G = num_sms();
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= TC_MAX_SPLIT && s <= tiles_min; ++s) {
    const int64_t rounds = cdiv((int64_t)combos * s, G);
    const double cost = (double)rounds * (double)(cdiv(tiles_max, s) + overhead_tiles);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

void tc_pick_splits(int obj_n, int64_t n_max, int64_t hw, int* split_a, int* split_b) {
  // upper bounds used to size the workspace; the per-launch choice (<= these) is made in tc_phase_a / tc_phase_b
  (void)obj_n; (void)n_max; (void)hw;
  *split_a = TC_MAX_SPLIT;
  *split_b = TC_MAX_SPLIT;
}

size_t tc_workspace_bytes(int obj_n, int64_t hw) {
  (void)obj_n;
  const size_t rows = (size_t)cdiv(hw, QT) * QT;
  return 2 * align_up(rows * DK * sizeof(uint16_t), 256);
}

static int fill_args(const vfn_bank* banks, int obj_n, int64_t hw, int pieces, int tile, char* ws_tc, TcMaps* maps,
                     TcArgs* a, bool need_v) {
  VFN_CHECK_ARG(obj_n <= TC_MAX_OBJ, "tcgen05 read supports at most %d objects", TC_MAX_OBJ);
  const size_t rows = (size_t)cdiv(hw, QT) * QT;
  a->obj_n = obj_n; a->hw = (int)hw; a->q_tiles = (int)cdiv(hw, QT); a->pieces = pieces;
  a->qh = reinterpret_cast<const uint16_t*>(ws_tc);
  a->ql = reinterpret_cast<const uint16_t*>(ws_tc + align_up(rows * DK * sizeof(uint16_t), 256));
  a->dbg = g_dbg;
  for (int o = 0; o < obj_n; ++o) {
    VFN_CHECK_ARG(banks[o].kh && banks[o].vh, "bank %d has no bf16 operand arrays", o);
    VFN_CHECK_ARG(banks[o].n < (1ll << 31), "bank too large");
    a->n[o] = (int)banks[o].n;
    a->tiles[o] = (int)cdiv(banks[o].n, tile);
    a->cnt[o] = banks[o].cnt;
    if (int rc = make_map(&maps->kh[o], banks[o].kh, banks[o].n, DK, tile)) return rc;
    if (int rc = make_map(&maps->kl[o], banks[o].kl, banks[o].n, DK, tile)) return rc;
    if (need_v) {
      if (int rc = make_map(&maps->vh[o], banks[o].vh, banks[o].n, DV, tile)) return rc;
      if (int rc = make_map(&maps->vl[o], banks[o].vl, banks[o].n, DV, tile)) return rc;
    }
  }
  return VFN_OK;
}

int tc_phase_a(const vfn_bank* banks, int obj_n, const float* q_in_dm, int64_t hw, int split_a, float2* part,
               char* ws_tc, cudaStream_t st, int* pieces_out) {
  static bool attr = false;
  if (!attr) {
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_phase_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_phase_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    attr = true;
  }
  TcMaps maps;
  TcArgs a;
  int64_t tmin = INT64_MAX, tmax = 0;
  for (int o = 0; o < obj_n; ++o) {
    const int64_t t = cdiv(banks[o].n, A_TILE);
    tmin = t < tmin ? t : tmin;
    tmax = t > tmax ? t : tmax;
  }
  const int pieces = best_split(obj_n * (int)cdiv(hw, QT), tmin, tmax, 3);
  if (pieces > split_a) { set_error("phase A: split %d exceeds workspace bound %d", pieces, split_a); return VFN_E_CAPACITY; }
  *pieces_out = pieces;
  if (int rc = fill_args(banks, obj_n, hw, pieces, A_TILE, ws_tc, &maps, &a, false)) return rc;
  const size_t rows = (size_t)a.q_tiles * QT;
  // Q hi/lo of q * log2(e)/sqrt(d): logits land in the log2 domain; pad rows zeroed
  VFN_CUDA_OK(cudaMemsetAsync(ws_tc, 0, 2 * align_up(rows * DK * sizeof(uint16_t), 256), st));
  const float scale = LOG2E / sqrtf((float)DK);
  if (int rc = vfn_prep_rows(q_in_dm, DK, hw, nullptr, nullptr, const_cast<uint16_t*>(a.qh),
                             const_cast<uint16_t*>(a.ql), scale, st))
    return rc;
  double work = 0;
  for (int o = 0; o < obj_n; ++o) work += 2.0 * DK * (double)banks[o].n * (double)hw;
  prof_begin(PROF_READ_A, st);
  tc_phase_a_kernel<<<num_sms(), TC_THREADS, A_SMEM, st>>>(maps, a, part);
  prof_end(PROF_READ_A, st, work);
  VFN_LAUNCH_OK();
  count_launches(2);
  return VFN_OK;
}

int tc_phase_b(const vfn_bank* banks, int obj_n, int64_t hw, int split_b, const float* lse, float thres_valid,
               int update_bank, float* po, char* ws_tc, cudaStream_t st, int* pieces_out) {
  TcMaps maps;
  TcArgs a;
  int64_t tmin = INT64_MAX, tmax = 0;
  for (int o = 0; o < obj_n; ++o) {
    const int64_t t = cdiv(banks[o].n, B_TILE);
    tmin = t < tmin ? t : tmin;
    tmax = t > tmax ? t : tmax;
  }
  const int pieces = best_split(obj_n * 2 * (int)cdiv(hw, QT), tmin, tmax, 4);
  if (pieces > split_b) { set_error("phase B: split %d exceeds workspace bound %d", pieces, split_b); return VFN_E_CAPACITY; }
  *pieces_out = pieces;
  if (int rc = fill_args(banks, obj_n, hw, pieces, B_TILE, ws_tc, &maps, &a, true)) return rc;
  double work = 0;
  for (int o = 0; o < obj_n; ++o) work += 2.0 * DV * (double)banks[o].n * (double)hw;
  prof_begin(PROF_READ_B, st);
  tc_phase_b_kernel<<<num_sms(), TC_THREADS, B_SMEM, st>>>(maps, a, lse, thres_valid, update_bank, po);
  prof_end(PROF_READ_B, st, work);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int tc_match(const vfn_bank* bank, const float* nck_em, int64_t hw, int max_pieces, float2* part, int32_t* idx_out,
             float* corr_out, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    VFN_CUDA_OK(cudaFuncSetAttribute(tc_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M_SMEM));
    attr = true;
  }
  VFN_CHECK_ARG(bank->d_key == DK && bank->n < (1ll << 31), "tcgen05 match needs d_key = 128");
  CUtensorMap mh, ml;
  if (int rc = make_map(&mh, bank->nkh, bank->n, DK, M_TILE, 4)) return rc;
  if (int rc = make_map(&ml, bank->nkl, bank->n, DK, M_TILE, 4)) return rc;
  MatchArgs a;
  a.hw = (int)hw; a.q_tiles = (int)cdiv(hw, QT); a.n = (int)bank->n; a.tiles = (int)cdiv(bank->n, M_TILE);
  a.nck = nck_em; a.dbg = g_dbg;
  int pieces = best_split(a.q_tiles, a.tiles, a.tiles, 3);
  if (pieces > max_pieces) pieces = max_pieces;
  a.pieces = pieces;
  prof_begin(PROF_MATCH, st);
  tc_match_kernel<<<num_sms(), TC_THREADS, M_SMEM, st>>>(mh, ml, a, part);
  prof_end(PROF_MATCH, st, 2.0 * DK * (double)bank->n * (double)hw);
  match_rescore_kernel<<<(unsigned)cdiv(hw, 8), 256, 0, st>>>(part, pieces, (int)hw, (int)bank->n, bank->nkh, bank->nkl,
                                                              nck_em, idx_out, corr_out);
  VFN_LAUNCH_OK();
  count_launches(2);
  return VFN_OK;
}

}  // namespace vfn

extern "C" int vfn_debug_set_dump(float* d_ptr) {
  vfn::g_dbg = d_ptr;
  return VFN_OK;
}
