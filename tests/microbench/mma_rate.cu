// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per instruction for the shapes the read kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
// One CTA (or CTA pair) per SM; one thread issues `reps` back-to-back MMAs into the same accumulator and commits;
// operands are zero-filled smem / TMEM (timing does not depend on the values).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sdesc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int af, int bf, int amn, int bmn) {
  return (1u << 4) | ((uint32_t)af << 7) | ((uint32_t)bf << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
enum { K_F16 = 0, K_F8 = 1, K_TF32 = 2 };

template <int KIND, int PAIR, int SS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d, uint32_t idesc) {
  if (SS) {
    if (KIND == K_F16) {
      if (PAIR) asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc) : "memory");
      else asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc) : "memory");
    } else {
      if (PAIR) asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc) : "memory");
      else asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc) : "memory");
    }
  } else {
    if (KIND == K_F16) {
      if (PAIR) asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc) : "memory");
      else asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc) : "memory");
    } else {
      if (PAIR) asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc) : "memory");
      else asm volatile("{.reg .pred p; setp.ne.b32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc) : "memory");
    }
  }
}

// B layout: bmn = 0: K-major SW128 (rows = N, 128 B of K per row); bmn = 1: MN-major SW128 (rows = K, 128 B of N per row)
template <int KIND, int PAIR, int SS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int bmn, int reps, int same_b, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_p;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_p)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_p)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_p;
  // zero the A region of TMEM (columns 256..383) so that no NaN patterns are multiplied
  {
    const uint32_t taddr = tmem + (((threadIdx.x >> 5) * 32u) << 16) + 256;
    for (int c = 0; c < 128; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(0) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int M = PAIR ? 256 : 128;
  const uint32_t idesc = make_idesc(M, N, KIND == K_F8 ? 0 : 0, 0, 0, bmn);
  const uint32_t a_smem = smem_u32(smem), b_smem = smem_u32(smem + 65536);
  if (threadIdx.x == 0 && rank == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int ks = same_b ? 0 : (r & 3);
      // K-major: advance 32 B inside the 128 B row per k-step; MN-major: advance 16 (f16) / 32 (f8) rows of 128 B
      const uint64_t bd = bmn ? make_sdesc(b_smem + ks * (KIND == K_F8 ? 4096u : 2048u), 8192, 1024)
                              : make_sdesc(b_smem + ks * 32u, 16, 1024);
      const uint64_t ad = make_sdesc(a_smem + ks * 32u, 16, 1024);
      mma<KIND, PAIR, SS>(tmem, tmem + 256 + ks * 8, ad, bd, idesc);
    }
    if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{.reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0; selp.u32 %0, 1, 0, P1;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  else __syncthreads();
  if (threadIdx.x < 32) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// Phase-B-shaped instruction mix on a CTA pair: per "tile" n_s S-MMAs (f16, K-major B of N_s slots, D = S buffer,
// A = Q in TMEM) followed by n_o O-MMAs (N = 256, MN-major B, D = O accumulator, A = the S buffer: half of them f16,
// half f8f6f4 when mix_f8).  stream != 0: every instruction reads a different B tile (5 distinct 16 KB windows) instead
// of the same one.  Prints cycles per tile.
template <int n_s, int N_s, int n_o, int mix_f8, int stream, int MP = 256>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mix_kernel(int tiles, int rnd, int commits, const uint8_t* fill_src, int fill_bytes, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];   // commit targets inside the tile loop (nobody waits on them)
  __shared__ uint64_t fbar;
  __shared__ volatile int stop_fill;
  __shared__ long long fill_count;
  __shared__ uint32_t tmem_p;
  // rnd: operands are pseudo-random finite fp16 / fp8 bit patterns (|x| < 2) instead of zeros - does the data matter?
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    uint4 v;
    h ^= h >> 15; h *= 2246822519u; v.x = rnd ? (h & 0xbbffbbffu) & 0xb7ffb7ffu : 0u;
    h ^= h >> 13; h *= 3266489917u; v.y = rnd ? (h & 0xb7ffb7ffu) : 0u;
    h ^= h >> 16; h *= 668265263u;  v.z = rnd ? (h & 0xb7ffb7ffu) : 0u;
    h ^= h >> 15; h *= 374761393u;  v.w = rnd ? (h & 0xb7ffb7ffu) : 0u;
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  uint32_t rank = 0;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fbar)));
    stop_fill = 0;
    fill_count = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_p)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_p;
  {
    const uint32_t taddr = tmem + (((threadIdx.x >> 5) * 32u) << 16) + 256;
    for (int c = 0; c < 256; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(rnd ? (((threadIdx.x * 2654435761u + c * 40503u) >> 3) & 0x37ff37ffu) : 0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // MP = 256: each CTA holds 128 rows (the production kernels); MP = 128: the "2x2" layout, 64 rows per CTA, the shape that
  // would keep 64 queries x 512 channels in 256 TMEM columns (S computed once for all channels)
  const uint32_t idesc_s = make_idesc(MP, N_s, 0, 0, 0, 0);
  const uint32_t idesc_o = make_idesc(MP, 256, 0, 0, 0, 1);
  const uint32_t b_smem = smem_u32(smem + 65536);
  // background fill (both CTAs): warp 1 streams fill_bytes-sized bulk copies from global memory into the unused A region
  // of shared memory, back to back, while the MMA loop runs - the TMA fills of the production kernels
  if (fill_src && threadIdx.x == 32) {
    uint32_t ph = 0;
    long long n = 0;
    const uint8_t* src = fill_src + (size_t)blockIdx.x * 65536;
    while (!stop_fill) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fbar)), "r"(fill_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                   "l"(src + (n & 3) * 16384), "r"(fill_bytes), "r"(smem_u32(&fbar))
                   : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{.reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1;}" : "=r"(done) : "r"(smem_u32(&fbar)), "r"(ph) : "memory");
      ph ^= 1;
      ++n;
    }
    fill_count = n;
  }
  if (threadIdx.x == 0 && rank == 0) {
    long long t0 = clock64();
    const int fences = commits / 10;
    commits %= 10;
    for (int t = 0; t < tiles; ++t) {
      if (fences & 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (fences & 2) { uint32_t d_; asm volatile("{.reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 1; selp.u32 %0, 1, 0, P1;}" : "=r"(d_) : "r"(smem_u32(&bar2[3])) : "memory"); }
      // fully unrolled, every descriptor a compile-time offset from b_smem: the issuing thread spends 2-3 instructions
      // per MMA, as the production kernels do
#pragma unroll
      for (int i = 0; i < n_s; ++i) {
        const uint32_t win = stream ? (uint32_t)(i % 5) * 16384u : 0u;
        const uint64_t bd = make_sdesc(b_smem + win + (i & 3) * 32u, 16, 1024);
        mma<K_F16, 1, 0>(tmem + 384, tmem + 256 + (i & 7) * 8, 0, bd, idesc_s);
      }
      if (fences & 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (fences & 2) { uint32_t d_; asm volatile("{.reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 1; selp.u32 %0, 1, 0, P1;}" : "=r"(d_) : "r"(smem_u32(&bar2[3])) : "memory"); }
      // commits as the production kernel issues them: after the S group (K stage free, S ready), after the O group
      if (commits >= 1) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar2[0])), "h"((uint16_t)3) : "memory");
      if (commits >= 3) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar2[1])), "h"((uint16_t)3) : "memory");
#pragma unroll
      for (int i = 0; i < n_o; ++i) {
        const uint32_t win = stream ? (uint32_t)((i + 2) % 5) * 16384u : 0u;
        if (mix_f8 && i >= n_o / 2) {
          const uint64_t bd = make_sdesc(b_smem + win, 8192, 1024);
          mma<K_F8, 1, 0>(tmem, tmem + 384 + (i & 3) * 8, 0, bd, idesc_o);
        } else {
          const uint64_t bd = make_sdesc(b_smem + win + (i & 3) * 2048u, 8192, 1024);
          mma<K_F16, 1, 0>(tmem, tmem + 384 + (i & 3) * 8, 0, bd, idesc_o);
        }
      }
      if (commits >= 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar2[2])), "h"((uint16_t)3) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{.reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0; selp.u32 %0, 1, 0, P1;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  if (threadIdx.x == 0) stop_fill = 1;
  if (threadIdx.x == 0 && rank != 0) { /* the peer stops when the leader's cluster barrier arrives */ }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int n_s, int N_s, int n_o, int mix_f8, int stream, int MP = 256>
void run_mix(const char* name, int rnd = 0, int commits = 0, int fill_bytes = 0) {
  long long* d;
  cudaMalloc(&d, 8);
  const int tiles = getenv("MIX_TILES") ? atoi(getenv("MIX_TILES")) : 256, smem = 160 * 1024 + 2048;
  const int launches = getenv("MIX_LAUNCHES") ? atoi(getenv("MIX_LAUNCHES")) : 2;
  auto k = mix_kernel<n_s, N_s, n_o, mix_f8, stream, MP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  static uint8_t* fsrc = nullptr;
  if (!fsrc) { cudaMalloc(&fsrc, 148 * 65536 + 65536); cudaMemset(fsrc, 1, 148 * 65536 + 65536); }
  for (int it = 0; it < launches; ++it) k<<<148, 128, smem>>>(tiles, rnd, commits, fill_bytes ? fsrc : nullptr, fill_bytes, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double model = n_s * (N_s >= 96 ? N_s / 2.0 : 45.6) + n_o * 128.0;
  if (fill_bytes) printf("[+ background bulk fills of %d B] ", fill_bytes);
  printf("mix %-22s %s S: %2d x N=%3d  O: %2d x N=256 %s %s  %8.1f cyc/tile  (sum of isolated rates %7.1f)  %s\n", name, rnd ? "random" : "zeros ", n_s, N_s, n_o,
         mix_f8 ? "f16+f8" : "f16   ", stream ? "stream B" : "same B  ", (double)h / tiles, model,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

template <int KIND, int PAIR, int SS>
void run(const char* name, int N, int bmn, int same_b = 0) {
  long long* d;
  cudaMalloc(&d, 8);
  const int reps = 4096, smem = 160 * 1024 + 2048;
  auto k = rate_kernel<KIND, PAIR, SS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) cudaLaunchKernelEx(&cfg, k, N, bmn, reps, same_b, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const int M = PAIR ? 256 : 128, K = KIND == K_F8 ? 32 : 16;
  const double cyc = (double)h / reps;
  const double ideal = (double)128 * N * K / (KIND == K_F8 ? 8192.0 : 4096.0);   // per SM: 128 rows
  printf("%-34s M=%3d N=%3d K=%2d %s  %8.1f cyc/instr  (floor %6.1f, %5.1f%%)  %s\n", name, M, N, K, bmn ? "B MN-major" : "B K-major ", cyc,
         ideal, 100.0 * ideal / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'h') {   // the M = 128 cta_group::2 (2x2 layout) question: do 16-clk S-MMAs hide behind the O-MMAs?
    run_mix<24, 64, 8, 1, 1>("M=256: 24 S + 8 O (today)", 1, 3);
    run_mix<24, 64, 16, 1, 1, 128>("M=128: 24 S + 16 O", 1, 3);
    run_mix<24, 64, 16, 1, 1, 128>("M=128: 24 S + 16 O, 0 commits", 1, 0);
    run_mix<24, 64, 0, 0, 1, 128>("M=128: S only", 1);
    run_mix<0, 64, 16, 1, 1, 128>("M=128: O only", 1);
    run_mix<24, 128, 32, 1, 1, 128>("M=128: 24 S(N=128) + 32 O", 1, 3);
    run_mix<12, 128, 16, 1, 1, 128>("M=128: 12 S(N=128) + 16 O", 1, 3);
    run_mix<24, 64, 16, 1, 1, 128>("M=128 + fills 16K", 1, 3, 16384);
    return 0;
  }
  if (argc > 1) {   // phase-B-shaped mixes only
    run_mix<24, 64, 8, 1, 1>("64: 3 commits+fences", 1, 13);
    run_mix<24, 64, 8, 1, 1>("64: 3 commits+try_wait", 1, 23);
    run_mix<24, 64, 8, 1, 1>("64: commits+fence+wait", 1, 33);
    run_mix<24, 64, 8, 1, 1>("64-slot + fills", 1, 3, 16384);
    run_mix<24, 64, 8, 1, 1>("64-slot + fills", 1, 3, 4096);
    run_mix<0, 64, 8, 1, 1>("O only + fills", 1, 0, 16384);
    run_mix<24, 64, 0, 0, 1>("S only + fills", 1, 0, 16384);
    run_mix<24, 64, 8, 1, 1>("64-slot, 0 commits", 1, 0);
    run_mix<24, 64, 8, 1, 1>("64-slot, 1 commit", 1, 1);
    run_mix<24, 64, 8, 1, 1>("64-slot, 2 commits", 1, 2);
    run_mix<24, 64, 8, 1, 1>("64-slot, 3 commits", 1, 3);
    run_mix<24, 96, 12, 1, 1>("96-slot, 3 commits", 1, 3);
    run_mix<24, 64, 8, 1, 1>("64-slot tile", 1);
    run_mix<24, 96, 12, 1, 1>("96-slot tile", 1);
    run_mix<0, 64, 8, 1, 1>("O only f16+f8", 1);
    run_mix<24, 64, 0, 0, 1>("S only N=64", 1);
    run_mix<24, 64, 8, 1, 0>("64-slot tile");
    run_mix<24, 96, 12, 1, 0>("96-slot tile");
    run_mix<24, 64, 8, 1, 1>("64-slot tile");
    run_mix<24, 96, 12, 1, 1>("96-slot tile");
    run_mix<24, 128, 16, 1, 1>("128-slot tile");
    run_mix<24, 64, 8, 0, 1>("64-slot, f16 only");
    run_mix<24, 64, 0, 0, 1>("S only N=64");
    run_mix<24, 96, 0, 0, 1>("S only N=96");
    run_mix<0, 64, 8, 0, 1>("O only f16");
    run_mix<0, 64, 8, 1, 1>("O only f16+f8");
    run_mix<8, 64, 8, 1, 1>("8 S + 8 O (1-pass S)");
    return 0;
  }
  for (int N : {64, 128, 256}) run<K_F16, 0, 0>("TS f16 1-CTA", N, 0);
  for (int N : {64, 128, 256}) run<K_F16, 0, 0>("TS f16 1-CTA same B", N, 0, 1);
  for (int N : {64, 128, 256}) run<K_F16, 0, 1>("SS f16 1-CTA", N, 0);
  for (int N : {128, 256}) run<K_F16, 0, 0>("TS f16 1-CTA", N, 1);
  for (int N : {128, 256}) run<K_F8, 0, 0>("TS f8  1-CTA", N, 1);
  for (int N : {64, 128, 256}) run<K_F8, 0, 0>("TS f8  1-CTA", N, 0);
  for (int N : {64, 128, 256}) run<K_F16, 1, 0>("TS f16 pair", N, 0);
  // intermediate widths (is the N = 64 penalty a fixed per-instruction cost?): candidates for 96-slot S tiles
  for (int N : {32, 48, 80, 96, 112, 160, 192}) run<K_F16, 1, 0>("TS f16 pair", N, 0);
  for (int N : {96, 192}) run<K_F16, 1, 1>("SS f16 pair", N, 0);
  for (int N : {96, 192}) run<K_F8, 1, 0>("TS f8  pair", N, 1);
  for (int N : {64, 128, 256}) run<K_F16, 1, 1>("SS f16 pair", N, 0);
  for (int N : {128, 256}) run<K_F16, 1, 0>("TS f16 pair", N, 1);
  for (int N : {128, 256}) run<K_F8, 1, 0>("TS f8  pair", N, 1);
  return 0;
}
