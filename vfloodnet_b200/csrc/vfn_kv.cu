// KeyValue head (reference: KeyValue.forward, video_module/model/AFB_URR.py:94-111): the two 3x3 / pad 1 convolutions
// 1024 -> 128 (Key) and 1024 -> 512 (Value) on the stride-16 feature map r4, as ONE tcgen05 implicit GEMM for sm_100a
// that writes keys and values entry-major - the layout of the read's query operand and of the bank update's candidate
// rows.
//
//   D[o, n] = sum_{tap = (ky, kx)} sum_c  Xp[o + ky*Wp + kx, c] * Wt[tap][n][c]            (+ bias[n] in the combine)
//
// Xp is the input re-laid as a zero-padded raster, one row of C channels per padded pixel ((h+2) x (w+2) rows per
// image, images stacked), so that the im2col operand of a tap is a plain 2-D box of the SAME matrix at a row offset:
// output "raster row" o = b*Hp*Wp + y*Wp + x reads padded pixel (y + ky, x + kx).  Rows with x >= w or y >= h are
// computed and dropped (10 % at 480p, 2 % at 1080p).  M = raster rows, N = 640 output channels, K = 9 * C.
//
// Precision: fp16 hi/lo splits of both operands (x * 2^k, w * 2^j with the exponents taken from the tensors' absolute
// maxima, on the device), three MMA passes hi*hi + lo*hi + hi*lo with fp32 accumulation (kind::f16): fp32-grade products
// (~2^-22), i.e. the result of a true-fp32 convolution up to summation order - NOT TF32.  passes = 1 keeps only hi*hi
// (11 significant bits per operand: the TF32 class of the reference's default cuDNN math).
//
// Kernel: CTA pairs (cta_group::2, M = 256 = 2 x 128 raster rows, N = 320 = two MMAs of N = 160 per k-step), both
// operands from shared memory via TMA (128B swizzle): two A slots of 2 x 17 KB (130 raster rows, hi / lo, shared by the
// three kx taps of a kernel row), three stages of 40 KB with this CTA's half of one tap's weight tile; fp32 accumulator
// 128 lanes x 320 columns in TMEM.  Work items = (row-tile pair, channel
// half, K split); one cluster per item; per-item partials are summed in fixed order by the combine kernel, which also
// applies the scales and the bias and writes the requested layouts.
#include "vfn_ptx.cuh"

namespace vfn {

constexpr int KV_MT = 128;                 // raster rows per CTA (TMEM lanes)
constexpr int KV_NH = 160;                 // N of one MMA
constexpr int KV_NT = 2 * KV_NH;           // output channels per work item
constexpr int KV_KC = 64;                  // input channels per K chunk (128 B of fp16: one swizzle row)
constexpr int KV_B_BYTES = (KV_NH / 2) * 128;        // 10 KB: this CTA's 80 of the 160 weight rows, hi or lo
constexpr int KV_THREADS = 384;            // warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue
constexpr int KV_MAX_SPLIT = 8;
constexpr int KV_CHAIN_MAX = 48;           // K chunks accumulated into one TMEM accumulator (see kv_geom)
constexpr int KV_X_EXP = 13, KV_W_EXP = 12;   // scaled operands: |x| <= 2^13, |w| <= 2^12

// the three kx taps of a kernel row read raster rows r, r+1, r+2: ONE box of 130 rows per (ky, channel chunk) serves all
// three (two A slots of hi + lo; three stages of weight tiles, one tap each)
constexpr int KR_A_ROWS = KV_MT + 2;
constexpr int KR_A_BYTES = 17 * 1024;                // 130 rows x 128 B = 16640 B, padded to whole 1024-byte swizzle atoms
constexpr int KR_A_SLOT = 2 * KR_A_BYTES;            // hi + lo
constexpr int KR_A_SLOTS = 2;
constexpr int KR_B_STAGE = 4 * KV_B_BYTES;           // 40 KB: this CTA's half of the weight tile of one tap, hi + lo
constexpr int KR_B_STAGES = 3;
constexpr int KR_SMEM = KR_A_SLOTS * KR_A_SLOT + KR_B_STAGES * KR_B_STAGE + 1024 + 256;

struct KvMaps { CUtensorMap wh, wl, xh_rows, xl_rows; };
struct KvArgs {
  int Wp, c_chunks, n_chunks, c_out;       // padded raster width; C / 64; 9 * C / 64; 640
  int n_ntiles, split, passes, m_pad;      // c_out / 320; K splits; 1 or 3; rows of one partial slab
  int n_groups;                            // 3 * C / 64 (ky, channel chunk) groups of three taps
  float* part;                             // [split][m_pad][c_out]
};

// packed weights: [256 B header | wh: 9 * c_out * C fp16 | wl: same | bias: c_out fp32]
// header: u32[0] = bits of max |w|, f32[1] = 2^-j (inverse weight scale), i32[2] = C, i32[3] = c_out
static size_t kv_w_elems(int c_in, int c_out) { return (size_t)9 * c_out * c_in; }
static size_t kv_packed_bytes(int c_in, int c_out) {
  return 256 + align_up(2 * kv_w_elems(c_in, c_out) * sizeof(uint16_t), 256) + (size_t)c_out * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kv_absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ cell) {
  pdl_wait();
  pdl_trigger();
  float m = 0.f;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(cell, __float_as_uint(m));   // non-negative floats order like their bits
}

// power-of-two scale that brings max |x| (given as float bits) to at most 2^target
__device__ __forceinline__ float kv_scale_from(uint32_t absmax_bits, int target) {
  const float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);                         // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, target - e);
}

// x (B, C, h, w) fp32  ->  xh / xl [(B * Hp * Wp) rows][C] fp16 hi / lo of x * scale, zero halo.
// grid (B * Hp, C / 64, ceil(Wp / 32)), block (32, 8)
__global__ void __launch_bounds__(256) kv_pack_input_kernel(const float* __restrict__ x, int C, int h, int w,
                                                           const uint32_t* __restrict__ absmax_cell,
                                                           float* __restrict__ inv_scale_cell, uint16_t* __restrict__ xh,
                                                           uint16_t* __restrict__ xl) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[64][33];
  const int Hp = h + 2, Wp = w + 2;
  const int b = blockIdx.x / Hp, yp = blockIdx.x - b * Hp;
  const int c0 = blockIdx.y * 64, xp0 = blockIdx.z * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float scale = kv_scale_from(*absmax_cell, KV_X_EXP);
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tx == 0 && ty == 0) *inv_scale_cell = 1.f / scale;
  const int y = yp - 1, xx = xp0 + tx - 1;
  const bool inside = y >= 0 && y < h && xx >= 0 && xx < w;
  for (int c = ty; c < 64; c += 8)
    tile[c][tx] = inside ? x[(((int64_t)b * C + c0 + c) * h + y) * w + xx] * scale : 0.f;
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int xp = xp0 + r;
    if (xp >= Wp) break;
    const int64_t row = ((int64_t)b * Hp + yp) * Wp + xp;
    uint16_t h0, l0, h1, l1;
    split_f16(tile[2 * tx][r], h0, l0);
    split_f16(tile[2 * tx + 1][r], h1, l1);
    const int64_t off = row * C + c0 + 2 * tx;
    *reinterpret_cast<uint32_t*>(xh + off) = (uint32_t)h0 | ((uint32_t)h1 << 16);
    *reinterpret_cast<uint32_t*>(xl + off) = (uint32_t)l0 | ((uint32_t)l1 << 16);
  }
}

// weights (n, C, 3, 3) of Key (n < dk) and Value -> wh / wl [tap][n][C] fp16 hi / lo of w * scale; bias -> packed bias
__global__ void __launch_bounds__(256) kv_pack_weights_kernel(const float* __restrict__ wk, const float* __restrict__ bk,
                                                             const float* __restrict__ wv, const float* __restrict__ bv,
                                                             int C, int dk, int c_out, uint32_t* __restrict__ hdr,
                                                             uint16_t* __restrict__ wh, uint16_t* __restrict__ wl,
                                                             float* __restrict__ bias) {
  const float scale = kv_scale_from(hdr[0], KV_W_EXP);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (n, c)
  if (i == 0) {
    reinterpret_cast<float*>(hdr)[1] = 1.f / scale;
    hdr[2] = (uint32_t)C;
    hdr[3] = (uint32_t)c_out;
  }
  if (i < c_out) bias[i] = i < dk ? (bk ? bk[i] : 0.f) : (bv ? bv[i - dk] : 0.f);
  if (i >= (int64_t)c_out * C) return;
  const int n = (int)(i / C), c = (int)(i - (int64_t)n * C);
  const float* src = (n < dk ? wk + ((int64_t)n * C + c) * 9 : wv + ((int64_t)(n - dk) * C + c) * 9);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    uint16_t hh, ll;
    split_f16(src[tap] * scale, hh, ll);
    const int64_t o = ((int64_t)tap * c_out + n) * C + c;
    wh[o] = hh;
    wl[o] = ll;
  }
}

// ------------------------------------------------------------------------------------------------
// the GEMM: one cluster (CTA pair) per (row-tile pair, channel half, K split)
// ------------------------------------------------------------------------------------------------
// The A operand of a (ky, channel chunk) group is loaded ONCE (130 raster rows) and used for kx = 0, 1, 2 through the
// descriptor's start address (+128 B per raster row): the 128B swizzle of TMA and of the tensor core are both functions
// of the shared-memory ADDRESS, so a start address inside a 1024-byte atom needs no further treatment (the descriptor's
// matrix-base-offset field must stay 0: setting it to the row offset gives wrong products - measured, r2z).  A CTA pulls
// 33 + 3 x 40 KB per three taps from L2 instead of 3 x 72 KB with one box per tap.  K splits are cut at group boundaries.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KV_THREADS, 1)
    kv_gemm_pair_kernel(const __grid_constant__ KvMaps maps, KvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* abuf = smem;
  uint8_t* bbuf = smem + KR_A_SLOTS * KR_A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + KR_B_STAGES * KR_B_STAGE);
  uint64_t* a_full = bars;                          // [2]
  uint64_t* a_empty = bars + 2;                     // [2]
  uint64_t* b_full = bars + 4;                      // [3]
  uint64_t* b_empty = bars + 7;                     // [3]
  uint64_t* acc_full = bars + 10;
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < KR_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < KR_B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_base_p, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_p;
  pdl_wait();
  pdl_trigger();

  const int item = blockIdx.x >> 1;
  const int sp = item % a.split;
  const int nt = (item / a.split) % a.n_ntiles;
  const int mp = item / (a.split * a.n_ntiles);
  const int g_begin = (int)((long long)a.n_groups * sp / a.split);
  const int g_end = (int)((long long)a.n_groups * (sp + 1) / a.split);
  const int ng = g_end - g_begin;
  const int m0 = (mp * 2 + (int)rank) * KV_MT;
  const int n0 = nt * KV_NT;
  const bool three = a.passes == 3;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&maps.xh_rows); prefetch_tmap(&maps.wh);
      if (three) { prefetch_tmap(&maps.xl_rows); prefetch_tmap(&maps.wl); }
      const uint32_t a_bytes = (three ? 2u : 1u) * KR_A_ROWS * 128u;
      const uint32_t b_bytes = three ? KR_B_STAGE : KR_B_STAGE / 2;
      uint32_t ci = 0;
      for (int gi = 0; gi < ng; ++gi) {
        const int g = g_begin + gi, ky = g / a.c_chunks, cc = g - ky * a.c_chunks, col = cc * KV_KC;
        const uint32_t slot = gi % KR_A_SLOTS, pha = (gi / KR_A_SLOTS) & 1;
        mbar_wait(&a_empty[slot], pha ^ 1);
        if (leader) mbar_arrive_expect_tx(&a_full[slot], 2 * a_bytes);
        const uint32_t af = mapa_u32(smem_u32(&a_full[slot]), 0);
        uint8_t* ad = abuf + slot * KR_A_SLOT;
        tma_load_2d_pair(ad, &maps.xh_rows, af, col, m0 + ky * a.Wp);
        if (three) tma_load_2d_pair(ad + KR_A_BYTES, &maps.xl_rows, af, col, m0 + ky * a.Wp);
        for (int kx = 0; kx < 3; ++kx, ++ci) {
          const uint32_t st = ci % KR_B_STAGES, ph = (ci / KR_B_STAGES) & 1;
          mbar_wait(&b_empty[st], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&b_full[st], 2 * b_bytes);
          const uint32_t bf = mapa_u32(smem_u32(&b_full[st]), 0);
          const int row_b = (ky * 3 + kx) * a.c_out + n0 + (int)rank * (KV_NH / 2);
          uint8_t* bd = bbuf + st * KR_B_STAGE;
          tma_load_2d_pair(bd, &maps.wh, bf, col, row_b);
          tma_load_2d_pair(bd + KV_B_BYTES, &maps.wh, bf, col, row_b + KV_NH);
          if (three) {
            tma_load_2d_pair(bd + 2 * KV_B_BYTES, &maps.wl, bf, col, row_b);
            tma_load_2d_pair(bd + 3 * KV_B_BYTES, &maps.wl, bf, col, row_b + KV_NH);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(2 * KV_MT, KV_NH, FMT_F16, FMT_F16, 0, 0);
      uint32_t ci = 0;
      for (int gi = 0; gi < ng; ++gi) {
        const uint32_t slot = gi % KR_A_SLOTS, pha = (gi / KR_A_SLOTS) & 1;
        mbar_wait(&a_full[slot], pha);
        const uint32_t abase = smem_u32(abuf + slot * KR_A_SLOT);
        for (int kx = 0; kx < 3; ++kx, ++ci) {
          const uint32_t st = ci % KR_B_STAGES, ph = (ci / KR_B_STAGES) & 1;
          mbar_wait(&b_full[st], ph);
          tc_fence_after();
          const uint32_t bbase = smem_u32(bbuf + st * KR_B_STAGE);
          for (int pass = 0; pass < (three ? 3 : 1); ++pass) {
            const uint32_t ab = abase + (pass == 1 ? (uint32_t)KR_A_BYTES : 0u) + (uint32_t)kx * 128u;
            const uint32_t bb = bbase + (pass == 2 ? 2u * KV_B_BYTES : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = make_sdesc(ab + ks * 32u, 16, 1024);
              const uint32_t acc = (ci | pass | ks) ? 1u : 0u;
              mma_ss_pair(tmem, ad, make_sdesc(bb + ks * 32u, 16, 1024), idesc, acc);
              mma_ss_pair(tmem + KV_NH, ad, make_sdesc(bb + KV_B_BYTES + ks * 32u, 16, 1024), idesc, acc);
            }
          }
          tc_commit_pair(&b_empty[st]);
        }
        tc_commit_pair(&a_empty[slot]);
      }
      tc_commit_pair(acc_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // TMEM -> registers -> this warp's staging tile in shared memory (the operand buffers are free: acc_full means every
    // MMA of the pair has completed) -> global rows.  A lane owns one accumulator ROW, so direct stores would put 32
    // rows x 16 B into every store instruction (10 k clk per item); from the tile a warp writes 512 contiguous bytes.
    constexpr int PITCH = KV_NH + 4;       // floats; 164 mod 32 = 4: the float4 stores of 8 lanes cover all 32 banks
    float* tile = reinterpret_cast<float*>(smem) + (size_t)(warp - 4) * 32 * PITCH;
    const uint32_t tl = tmem + (((uint32_t)quarter * 32u) << 16) + (uint32_t)half * KV_NH;
#pragma unroll 1
    for (int g = 0; g < KV_NH / 32; ++g) {
      uint32_t v[32];
      tmem_ld32(tl + g * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(tile + lane * PITCH + g * 32 + 4 * j) =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                        __uint_as_float(v[4 * j + 3]));
    }
    __syncwarp();
    float* dst = a.part + ((size_t)sp * a.m_pad + m0 + (quarter << 5)) * a.c_out + n0 + half * KV_NH;
#pragma unroll 4
    for (int i = lane; i < 32 * (KV_NH / 4); i += 32) {
      const int row = i / (KV_NH / 4), c4 = i - row * (KV_NH / 4);
      *reinterpret_cast<float4*>(dst + (size_t)row * a.c_out + 4 * c4) =
          *reinterpret_cast<const float4*>(tile + row * PITCH + 4 * c4);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// combine: out[pos, n] = (sum_s part[s][o(pos)][n]) * 2^-(k+j) + bias[n], in the layouts asked for:
//   *_em (B, h*w, d) entry-major;  *_dm (B, d, h*w) = the reference's KeyValue output layout
// grid (ceil(B*h*w / 32), c_out / 32), block (32, 8)
// ------------------------------------------------------------------------------------------------
struct KvOut { float *key_em, *val_em, *key_dm, *val_dm; };

__global__ void __launch_bounds__(256) kv_combine_kernel(const float* __restrict__ part, int split, int m_pad, int c_out,
                                                        int B, int h, int w, const float* __restrict__ inv_sx,
                                                        const float* __restrict__ inv_sw, const float* __restrict__ bias,
                                                        int dk, KvOut out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int hw = h * w, total = B * hw, Wp = w + 2, HpWp = (h + 2) * Wp;
  const int dv = c_out - dk;
  const float inv = (*inv_sx) * (*inv_sw);
  const int n = blockIdx.y * 32 + tx;
  const float bn = bias[n];
  const bool is_key = n < dk;            // dk % 32 == 0: a 32-channel block is all key or all value
  // all (row, split) loads of a thread are issued before the first add: 4 x 8 independent loads in flight
  float pv[4][KV_MAX_SPLIT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = blockIdx.x * 32 + ty + 8 * i;
    const int pc = pos < total ? pos : total - 1;
    const int b = pc / hw, p = pc - b * hw, y = p / w, x = p - y * w;
    const size_t o = (size_t)b * HpWp + (size_t)y * Wp + x;
#pragma unroll
    for (int s = 0; s < KV_MAX_SPLIT; ++s) pv[i][s] = s < split ? part[((size_t)s * m_pad + o) * c_out + n] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 8 * i, pos = blockIdx.x * 32 + r;
    float acc = pv[i][0];                      // splits in ascending order (x + 0 is exact for the unused ones)
#pragma unroll
    for (int s = 1; s < KV_MAX_SPLIT; ++s) acc += pv[i][s];
    const float v = acc * inv + bn;
    if (pos < total) {
      if (is_key) { if (out.key_em) out.key_em[(size_t)pos * dk + n] = v; }
      else if (out.val_em) out.val_em[(size_t)pos * dv + (n - dk)] = v;
    }
    tile[r][tx] = v;
  }
  float* dm = blockIdx.y * 32 < dk ? out.key_dm : out.val_dm;
  if (!dm) return;
  __syncthreads();
  const int pos = blockIdx.x * 32 + tx;
  if (pos >= total) return;
  const int b = pos / hw, p = pos - b * hw;
  const int d = blockIdx.y * 32 < dk ? dk : dv, nbase = blockIdx.y * 32 < dk ? blockIdx.y * 32 : blockIdx.y * 32 - dk;
  for (int r = ty; r < 32; r += 8) dm[((size_t)b * d + nbase + r) * hw + p] = tile[tx][r];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int kv_num_sms() {
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  return v > 0 ? v : 148;
}

struct KvGeom {
  int Hp, Wp, rows;          // padded raster: rows = B * Hp * Wp
  int m_ext, n_mpairs, m_pad, n_ntiles, n_chunks, split;
};

static KvGeom kv_geom(int B, int c_in, int h, int w, int c_out, int n_sm) {
  KvGeom g;
  g.Hp = h + 2; g.Wp = w + 2; g.rows = B * g.Hp * g.Wp;
  g.m_ext = (B - 1) * g.Hp * g.Wp + (h - 1) * g.Wp + w;
  g.n_mpairs = (int)cdiv(g.m_ext, 2 * KV_MT);
  g.m_pad = g.n_mpairs * 2 * KV_MT;
  g.n_ntiles = c_out / KV_NT;
  g.n_chunks = 9 * (c_in / KV_KC);
  // K split: minimise rounds x (chunks per item + fixed per-item cost) over the clusters the device runs at once
  const int G = n_sm / 2 > 0 ? n_sm / 2 : 1, base = g.n_mpairs * g.n_ntiles;
  int best = 1;
  long long best_cost = -1;
  // The tensor core accumulates with truncation (a systematic ~2^-25 relative per accumulation step): one
  // TMEM accumulator runs over at most KV_CHAIN_MAX chunks (576 steps, ~1e-5 relative; measured 7e-6 at 29 chunks and
  // 3e-5 at 144); the partial slabs are added in fp32 round-to-nearest by the combine kernel.
  const int s_min = (int)cdiv(g.n_chunks, KV_CHAIN_MAX) < KV_MAX_SPLIT ? (int)cdiv(g.n_chunks, KV_CHAIN_MAX) : KV_MAX_SPLIT;
  best = s_min;
  const int n_groups = g.n_chunks / 3;       // the K range is cut at (ky, channel chunk) groups of three taps
  for (int s = s_min; s <= KV_MAX_SPLIT && s <= n_groups; ++s) {
    const long long rounds = cdiv((int64_t)base * s, G);
    const long long cost = rounds * (3 * cdiv(n_groups, s) + 6) + 2 * s;    // + 2 s: the combine reads s partial slabs
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  g.split = best;
  return g;
}

// the tensor map of the raster always spans at least one row tile (maps with fewer rows than their box are avoided)
static int kv_map_rows(const KvGeom& g) { return g.rows > KR_A_ROWS ? g.rows : KR_A_ROWS; }
static size_t kv_x_bytes(const KvGeom& g, int c_in) { return align_up((size_t)kv_map_rows(g) * c_in * sizeof(uint16_t), 1024); }

static int kv_check_dims(int B, int c_in, int h, int w, int dk, int dv) {
  VFN_CHECK_ARG(B >= 1 && h >= 1 && w >= 1, "keyvalue: bad shape");
  VFN_CHECK_ARG(c_in >= KV_KC && c_in % KV_KC == 0, "keyvalue: input channels must be a multiple of %d", KV_KC);
  VFN_CHECK_ARG(dk > 0 && dv > 0 && dk % 32 == 0 && dv % 32 == 0 && (dk + dv) % KV_NT == 0,
                "keyvalue: d_key and d_val must be multiples of 32 and d_key + d_val a multiple of %d", KV_NT);
  VFN_CHECK_ARG((int64_t)B * (h + 2) * (w + 2) < (1ll << 30), "keyvalue: feature map too large");
  return VFN_OK;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

size_t vfn_kv_packed_weights_bytes(int32_t c_in, int32_t d_key, int32_t d_val) {
  if (c_in < 1 || d_key < 1 || d_val < 1) return 0;
  return kv_packed_bytes(c_in, d_key + d_val);
}

int vfn_kv_pack_weights(const float* d_wk, const float* d_bk, const float* d_wv, const float* d_bv, int32_t c_in,
                        int32_t d_key, int32_t d_val, void* d_packed, void* stream) {
  VFN_CHECK_ARG(d_wk && d_wv && d_packed, "kv_pack_weights: null pointer");
  if (int rc = kv_check_dims(1, c_in, 1, 1, d_key, d_val)) return rc;
  cudaStream_t st = as_stream(stream);
  const int c_out = d_key + d_val;
  char* base = reinterpret_cast<char*>(d_packed);
  uint32_t* hdr = reinterpret_cast<uint32_t*>(base);
  uint16_t* wh = reinterpret_cast<uint16_t*>(base + 256);
  uint16_t* wl = wh + kv_w_elems(c_in, c_out);
  float* bias = reinterpret_cast<float*>(base + 256 + align_up(2 * kv_w_elems(c_in, c_out) * sizeof(uint16_t), 256));
  VFN_CUDA_OK(cudaMemsetAsync(hdr, 0, 256, st));
  kv_absmax_kernel<<<148, 256, 0, st>>>(d_wk, (int64_t)d_key * c_in * 9, hdr);
  kv_absmax_kernel<<<148, 256, 0, st>>>(d_wv, (int64_t)d_val * c_in * 9, hdr);
  const int64_t n = (int64_t)c_out * c_in;
  kv_pack_weights_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(d_wk, d_bk, d_wv, d_bv, c_in, d_key, c_out, hdr, wh, wl, bias);
  VFN_LAUNCH_OK();
  count_launches(3);
  return VFN_OK;
}

size_t vfn_keyvalue_workspace_bytes(int32_t B, int32_t c_in, int32_t h, int32_t w, int32_t d_key, int32_t d_val) {
  if (B < 1 || c_in < KV_KC || h < 1 || w < 1 || d_key < 1 || d_val < 1) return 0;
  // the K split (number of partial slabs) follows the SM count of the current device, as in vfn_keyvalue itself
  KvGeom g = kv_geom(B, c_in, h, w, d_key + d_val, kv_num_sms());
  return 256 + 2 * kv_x_bytes(g, c_in) + (size_t)g.split * g.m_pad * (d_key + d_val) * sizeof(float);
}

int vfn_keyvalue(const float* d_x, int32_t B, int32_t c_in, int32_t h, int32_t w, const void* d_packed, int32_t d_key,
                 int32_t d_val, int32_t passes, float* d_key_em, float* d_val_em, float* d_key_dm, float* d_val_dm,
                 void* d_ws, size_t ws_bytes, void* stream) {
  VFN_CHECK_ARG(d_x && d_packed && d_ws, "keyvalue: null pointer");
  VFN_CHECK_ARG(passes == 1 || passes == 3, "keyvalue: passes must be 1 (hi*hi only, TF32 class) or 3 (fp32 grade)");
  VFN_CHECK_ARG(d_key_em || d_key_dm || d_val_em || d_val_dm, "keyvalue: no output requested");
  if (int rc = kv_check_dims(B, c_in, h, w, d_key, d_val)) return rc;
  if (!vfn_device_is_sm100()) { set_error("keyvalue: needs an sm_100 device (tcgen05)"); return VFN_E_UNSUPPORTED; }
  const int n_sm = kv_num_sms();
  VFN_CHECK_ARG(n_sm % 2 == 0, "keyvalue: CTA pairs need an even SM count");
  const int c_out = d_key + d_val;
  const KvGeom g = kv_geom(B, c_in, h, w, c_out, n_sm);
  if (ws_bytes < vfn_keyvalue_workspace_bytes(B, c_in, h, w, d_key, d_val)) {
    set_error("keyvalue: workspace %zu < %zu", ws_bytes, vfn_keyvalue_workspace_bytes(B, c_in, h, w, d_key, d_val));
    return VFN_E_CAPACITY;
  }
  static bool attr[64] = {false};
  if (first_use_on_device(attr))
    VFN_CUDA_OK(cudaFuncSetAttribute(kv_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KR_SMEM));
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(d_ws);
  uint32_t* cells = reinterpret_cast<uint32_t*>(ws);                    // [0] bits of max |x|, [1] 2^-k as float
  uint16_t* xh = reinterpret_cast<uint16_t*>(ws + 256);
  uint16_t* xl = reinterpret_cast<uint16_t*>(ws + 256 + kv_x_bytes(g, c_in));
  float* part = reinterpret_cast<float*>(ws + 256 + 2 * kv_x_bytes(g, c_in));
  const char* pk = reinterpret_cast<const char*>(d_packed);
  const uint16_t* wh = reinterpret_cast<const uint16_t*>(pk + 256);
  const uint16_t* wl = wh + kv_w_elems(c_in, c_out);
  const float* bias = reinterpret_cast<const float*>(pk + 256 + align_up(2 * kv_w_elems(c_in, c_out) * sizeof(uint16_t), 256));
  const float* inv_sw = reinterpret_cast<const float*>(pk) + 1;

  VFN_CUDA_OK(cudaMemsetAsync(cells, 0, 256, st));
  if (g.rows < KR_A_ROWS) VFN_CUDA_OK(cudaMemsetAsync(xh, 0, 2 * kv_x_bytes(g, c_in), st));   // rows the packing never writes
  const int64_t nx = (int64_t)B * c_in * h * w;
  kv_absmax_kernel<<<4 * n_sm, 256, 0, st>>>(d_x, nx, cells);
  dim3 pg((unsigned)(B * g.Hp), (unsigned)(c_in / 64), (unsigned)cdiv(g.Wp, 32));
  VFN_CUDA_OK(launch_pdl(kv_pack_input_kernel, pg, dim3(32, 8), 0, st, d_x, (int)c_in, (int)h, (int)w,
                         (const uint32_t*)cells, reinterpret_cast<float*>(cells) + 1, xh, xl));

  KvMaps maps;
  if (int rc = make_map(&maps.wh, wh, (int64_t)9 * c_out, c_in, KV_NH / 2, 2)) return rc;
  if (int rc = make_map(&maps.wl, wl, (int64_t)9 * c_out, c_in, KV_NH / 2, 2)) return rc;
  if (int rc = make_map(&maps.xh_rows, xh, kv_map_rows(g), c_in, KR_A_ROWS, 2)) return rc;
  if (int rc = make_map(&maps.xl_rows, xl, kv_map_rows(g), c_in, KR_A_ROWS, 2)) return rc;
  KvArgs a;
  a.Wp = g.Wp; a.c_chunks = c_in / KV_KC; a.n_chunks = g.n_chunks; a.c_out = c_out;
  a.n_ntiles = g.n_ntiles; a.split = g.split; a.passes = passes; a.m_pad = g.m_pad; a.part = part;
  a.n_groups = g.n_chunks / 3;
  const int items = g.n_mpairs * g.n_ntiles * g.split;
  prof_begin(PROF_KV, st);
  VFN_CUDA_OK(launch_pdl(kv_gemm_pair_kernel, dim3(2 * items), dim3(KV_THREADS), KR_SMEM, st, maps, a));
  prof_end(PROF_KV, st, 2.0 * 9.0 * c_in * c_out * (double)B * h * w);
  KvOut out{d_key_em, d_val_em, d_key_dm, d_val_dm};
  dim3 cg((unsigned)cdiv((int64_t)B * h * w, 32), (unsigned)(c_out / 32));
  VFN_CUDA_OK(launch_pdl(kv_combine_kernel, cg, dim3(32, 8), 0, st, (const float*)part, g.split, g.m_pad, c_out, (int)B,
                         (int)h, (int)w, (const float*)(reinterpret_cast<float*>(cells) + 1), inv_sw, bias, (int)d_key, out));
  VFN_LAUNCH_OK();
  count_launches(4);
  return VFN_OK;
}

}  // extern "C"
