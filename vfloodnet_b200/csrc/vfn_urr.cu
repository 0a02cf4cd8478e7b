// Uncertain-region refinement: the non-convolution parts of Decoder.forward (AFB_URR.py:214-237,
// myutils/data.py:42-48).  Bandwidth-bound streaming kernels; bs = 1 (inference).
#include "vfn_common.cuh"

namespace vfn {

// bilinear x2, align_corners=False (ATen area_pixel_compute_source_index with scale 0.5)
__device__ __forceinline__ void src_index(int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = 0.5f * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ pl, int win, int y0, int y1, int x0, int x1, float ly0,
                                        float ly1, float lx0, float lx1) {
  return ly0 * (lx0 * pl[y0 * win + x0] + lx1 * pl[y0 * win + x1]) +
         ly1 * (lx0 * pl[y1 * win + x0] + lx1 * pl[y1 * win + x1]);
}

// r1_local = (box7(r1*seg) / 49) / (avg + 1e-8)   (AFB_URR.py:227-228).  Both URR-local kernels go through this one
// function so that they stay bit-identical to each other.  1/49 and the reciprocal are 1-2 ulp operations
// (MUFU.RCP + FMUL): 3e-7 relative against the reference's two IEEE divisions, inside the 1e-5 bar of the URR tests -
// the two correctly rounded fp32 divisions were 40 % of the streaming kernel's instructions (16 FCHK/CALL slow-path
// blocks per row: profiles/r2_urr_local.md).
__device__ __forceinline__ float urr_ratio(float tot, float av) {
  return __fdividef(tot * (1.f / 49.f), av + 1e-8f);
}

// stage 1: p (obj,2,h/2,w/2) -> p_up (obj,2,h,w), seg (obj,h,w) object-normalised fg prob, unc (h,w)
constexpr int URR_MAX_OBJ = 8;
__global__ void urr_seg_kernel(const float* __restrict__ p, int obj_n, int h, int w, float* __restrict__ p_up,
                               float* __restrict__ seg, float* __restrict__ unc) {
  pdl_wait();
  pdl_trigger();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const int hin = h >> 1, win = w >> 1;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(y, hin, y0, y1, ly0, ly1);
  src_index(x, win, x0, x1, lx0, lx1);
  float fg[URR_MAX_OBJ];
  float mx = -INFINITY;
  for (int o = 0; o < obj_n; ++o) {
    const float* pl = p + (int64_t)o * 2 * hin * win;
    const float a = bilerp(pl, win, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
    const float b = bilerp(pl + hin * win, win, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
    p_up[((int64_t)o * 2 + 0) * h * w + (int64_t)y * w + x] = a;
    p_up[((int64_t)o * 2 + 1) * h * w + (int64_t)y * w + x] = b;
    const float m = fmaxf(a, b);
    const float e0 = expf(a - m), e1 = expf(b - m);
    fg[o] = e1 / (e0 + e1);                       // softmax(p, dim=1)[:, 1]           AFB_URR.py:217
    mx = fmaxf(mx, fg[o]);
  }
  float sum = 0.f;
  for (int o = 0; o < obj_n; ++o) { fg[o] = expf(fg[o] - mx); sum += fg[o]; }
  float t1 = -INFINITY, t2 = -INFINITY;
  for (int o = 0; o < obj_n; ++o) {
    const float s = fg[o] / sum;                  // object-level softmax                AFB_URR.py:219
    seg[(int64_t)o * h * w + (int64_t)y * w + x] = s;
    if (s > t1) { t2 = t1; t1 = s; } else if (s > t2) { t2 = s; }
  }
  unc[(int64_t)y * w + x] = expf(1.f - t1 / (t2 + 1e-8f));   // myutils/data.py:45-47
}

// stage 2: conf = 7x7 max of seg (-inf pad), avg = 7x7 mean of seg (zero pad, /49)
__global__ void urr_window_kernel(const float* __restrict__ seg, int obj_n, int h, int w, float* __restrict__ conf,
                                  float* __restrict__ avg) {
  pdl_wait();
  pdl_trigger();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, o = blockIdx.z;
  if (x >= w) return;
  const float* pl = seg + (int64_t)o * h * w;
  float mx = -INFINITY, sum = 0.f;
  for (int dy = -3; dy <= 3; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= h) continue;
    for (int dx = -3; dx <= 3; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= w) continue;
      const float v = pl[yy * w + xx];
      mx = fmaxf(mx, v);
      sum += v;
    }
  }
  conf[(int64_t)o * h * w + (int64_t)y * w + x] = mx;
  avg[(int64_t)o * h * w + (int64_t)y * w + x] = sum / 49.f;
}

// stages 1 + 2 in one launch (w, h any size; obj_n <= URR_MAX_OBJ): a CTA owns a 32 x 16 pixel tile, evaluates stage 1
// on the tile plus its 3-pixel halo into shared memory (the halo's stage-1 values are recomputed, 1.6x of a cheap
// stage), writes p_up / seg / unc for its own pixels and takes the 7x7 windows from shared memory in the order of
// urr_window_kernel (dy outer, dx inner): the same bits as the two separate launches, one stream gap fewer per frame.
constexpr int SW_W = 32, SW_H = 16, SW_HW = SW_W + 6, SW_HH = SW_H + 6;
__global__ void __launch_bounds__(256) urr_segwin_kernel(const float* __restrict__ p, int obj_n, int h, int w,
                                                        float* __restrict__ p_up, float* __restrict__ seg,
                                                        float* __restrict__ unc, float* __restrict__ conf,
                                                        float* __restrict__ avg) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sseg[URR_MAX_OBJ][SW_HH][SW_HW + 1];
  const int x0t = blockIdx.x * SW_W, y0t = blockIdx.y * SW_H;
  const int hin = h >> 1, win = w >> 1;
  const int64_t plane = (int64_t)h * w;
  for (int i = threadIdx.x; i < SW_HH * SW_HW; i += blockDim.x) {
    const int ly = i / SW_HW, lx = i - ly * SW_HW;
    const int y = y0t + ly - 3, x = x0t + lx - 3;
    if (y < 0 || y >= h || x < 0 || x >= w) continue;          // never read: the windows skip pixels outside the image
    const bool own = ly >= 3 && ly < 3 + SW_H && lx >= 3 && lx < 3 + SW_W;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(y, hin, y0, y1, ly0, ly1);
    src_index(x, win, x0, x1, lx0, lx1);
    float fg[URR_MAX_OBJ];
    float mx = -INFINITY;
    for (int o = 0; o < obj_n; ++o) {
      const float* pl = p + (int64_t)o * 2 * hin * win;
      const float a = bilerp(pl, win, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
      const float b = bilerp(pl + hin * win, win, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
      if (own) {
        p_up[((int64_t)o * 2 + 0) * plane + (int64_t)y * w + x] = a;
        p_up[((int64_t)o * 2 + 1) * plane + (int64_t)y * w + x] = b;
      }
      const float m = fmaxf(a, b);
      const float e0 = expf(a - m), e1 = expf(b - m);
      fg[o] = e1 / (e0 + e1);                       // softmax(p, dim=1)[:, 1]           AFB_URR.py:217
      mx = fmaxf(mx, fg[o]);
    }
    float sum = 0.f;
    for (int o = 0; o < obj_n; ++o) { fg[o] = expf(fg[o] - mx); sum += fg[o]; }
    float t1 = -INFINITY, t2 = -INFINITY;
    for (int o = 0; o < obj_n; ++o) {
      const float s = fg[o] / sum;                  // object-level softmax                AFB_URR.py:219
      sseg[o][ly][lx] = s;
      if (own) seg[(int64_t)o * plane + (int64_t)y * w + x] = s;
      if (s > t1) { t2 = t1; t1 = s; } else if (s > t2) { t2 = s; }
    }
    if (own) unc[(int64_t)y * w + x] = expf(1.f - t1 / (t2 + 1e-8f));   // myutils/data.py:45-47
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SW_H * SW_W * obj_n; i += blockDim.x) {
    const int o = i / (SW_H * SW_W), r = i - o * (SW_H * SW_W);
    const int ty = r / SW_W, tx = r - ty * SW_W;
    const int y = y0t + ty, x = x0t + tx;
    if (y >= h || x >= w) continue;
    float mx = -INFINITY, sum = 0.f;
    for (int dy = -3; dy <= 3; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = -3; dx <= 3; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const float v = sseg[o][ty + 3 + dy][tx + 3 + dx];
        mx = fmaxf(mx, v);
        sum += v;
      }
    }
    conf[(int64_t)o * plane + (int64_t)y * w + x] = mx;
    avg[(int64_t)o * plane + (int64_t)y * w + x] = sum / 49.f;
  }
}

// stage 3: local_match[o][ch] = r1[ch] ; local_match[o][C+ch] = box7(r1[ch]*seg[o])/49 / (avg[o] + 1e-8)
// CTA = one channel x one 16-row x 128-column tile, ALL objects (r1 is read once and shared by the objects, the
// reference's `expand`).  Stage 1 puts r1 and r1*seg[o] (with a 3-pixel halo) in shared memory; stage 2 gives each
// thread one column: horizontal 7-tap from shared memory, vertical 7-tap from a register ring - no further barriers.
constexpr int UT_W = 128, UT_H = 16, UHALO = 3, UW = UT_W + 2 * UHALO, UHH = UT_H + 2 * UHALO;
constexpr int URR_LOCAL_MAX_OBJ = 2;     // objects per pass (more objects: several passes over the tile)
constexpr int URR_LOCAL_THREADS = UT_W * URR_LOCAL_MAX_OBJ;
__global__ void __launch_bounds__(URR_LOCAL_THREADS) urr_local_kernel(const float* __restrict__ r1, int64_t r1_obj_stride,
                                                                      int c_n, int obj_n, int h, int w,
                                                                      const float* __restrict__ seg,
                                                                      const float* __restrict__ avg,
                                                                      float* __restrict__ lm) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sr1[UHH][UW];
  __shared__ float prod[URR_LOCAL_MAX_OBJ][UHH][UW];
  const int ch = blockIdx.z;
  const int x0 = blockIdx.x * UT_W, y0 = blockIdx.y * UT_H;
  const int64_t plane = (int64_t)h * w;
  for (int ob = 0; ob < obj_n; ob += URR_LOCAL_MAX_OBJ) {
    const int no = min(URR_LOCAL_MAX_OBJ, obj_n - ob);
    // with a per-object r1 (r1_obj_stride != 0) only one object per pass can share the tile
    const int npass = r1_obj_stride ? 1 : no;
    for (int sub = 0; sub < no; sub += npass) {
      const float* rp = r1 + (int64_t)(ob + sub) * r1_obj_stride + (int64_t)ch * plane;
      const float* sp = seg + (int64_t)(ob + sub) * plane;
      __syncthreads();
#pragma unroll 4
      for (int i = threadIdx.x; i < UHH * UW; i += URR_LOCAL_THREADS) {
        const int ly = i / UW, lx = i - ly * UW;
        const int yy = y0 + ly - UHALO, xx = x0 + lx - UHALO;
        const bool in = (yy >= 0 && yy < h && xx >= 0 && xx < w);
        const int64_t off = (int64_t)yy * w + xx;
        const float v = in ? __ldg(rp + off) : 0.f;
        const float s0 = in ? __ldg(sp + off) : 0.f;
        const float s1 = (in && npass > 1) ? __ldg(sp + plane + off) : 0.f;
        sr1[ly][lx] = v;
        prod[0][ly][lx] = v * s0;                         // r1 * rough_seg                   AFB_URR.py:226
        prod[1][ly][lx] = v * s1;
      }
      __syncthreads();
      // one thread per (column, object): horizontal 7-tap from shared memory, vertical 7-tap from a register ring
      const int lx = threadIdx.x & (UT_W - 1), k = threadIdx.x >> 7, xx = x0 + lx;
      if (xx < w && k < npass) {
        const int o = ob + sub + k;
        float* out_raw = lm + ((int64_t)o * 2 * c_n + ch) * plane;
        float* out_loc = lm + ((int64_t)o * 2 * c_n + c_n + ch) * plane;
        const float* av = avg + (int64_t)o * plane;
        float ring[7];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          float s = 0.f;
#pragma unroll
          for (int t = 0; t < 7; ++t) s += prod[k][r][lx + t];
          ring[r] = s;
        }
        float avv[UT_H];
#pragma unroll
        for (int ly = 0; ly < UT_H; ++ly) avv[ly] = (y0 + ly < h) ? __ldg(av + (int64_t)(y0 + ly) * w + xx) : 1.f;
#pragma unroll
        for (int ly = 0; ly < UT_H; ++ly) {
          float s = 0.f;
#pragma unroll
          for (int t = 0; t < 7; ++t) s += prod[k][ly + 6][lx + t];
          ring[(ly + 6) % 7] = s;
          const int yy = y0 + ly;
          if (yy < h) {
            float tot = 0.f;
#pragma unroll
            for (int r = 0; r < 7; ++r) tot += ring[(ly + r) % 7];       // rows ly .. ly+6 of the tile, top to bottom
            const int64_t off = (int64_t)yy * w + xx;
            out_raw[off] = sr1[ly + UHALO][lx + UHALO];
            out_loc[off] = urr_ratio(tot, avv[ly]);                       // AFB_URR.py:227-228
          }
        }
      }
    }
  }
}

// stage 3, streaming form (w % 4 == 0): same arithmetic as urr_local_kernel, organised for HBM bandwidth.
// A warp owns 30 float4 columns (lanes 1..30; lanes 0 and 31 only carry the 3-pixel horizontal halo) of ONE channel and
// streams down a band of rows: per input row it loads r1 and seg[o] as float4, forms r1*seg, gets the horizontal 7-tap
// sums with six warp shuffles per object, and keeps the last seven row sums in a register ring for the vertical 7-tap.
// No shared memory, no barriers; every global access is a 128-bit coalesced load/store; r1 is read once for all objects
// of a pass (the reference's `expand`) and written once per object into [r1 ; r1_local].
constexpr int UL_COLS = 30;          // float4 columns per warp
constexpr int UL_WARPS = 4;
template <int NO>
__global__ void __launch_bounds__(UL_WARPS * 32) urr_local_stream_kernel(
    const float* __restrict__ r1, int64_t r1_obj_stride, int c_n, int obj_n, int h, int w, int band,
    const float* __restrict__ seg, const float* __restrict__ avg, float* __restrict__ lm) {
  pdl_wait();
  pdl_trigger();
  // the warps of a CTA take DIFFERENT channels of the same (band, column range): their seg / avg loads coincide and are
  // served by L1 after the first warp's miss (seg and avg are re-read by every channel: 4 planes x 64 channels)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w4 = w >> 2;
  const int x4 = blockIdx.x * UL_COLS + lane - 1;
  const bool col_ok = (x4 >= 0 && x4 < w4);
  const bool out_lane = col_ok && lane >= 1 && lane <= UL_COLS;
  const int ogroups = (obj_n + NO - 1) / NO;            // object group fastest: the re-reads of an r1 row are close in time
  const int ch = (blockIdx.z / ogroups) * UL_WARPS + warp, ob = (blockIdx.z % ogroups) * NO;
  if (ch >= c_n) return;
  const int no = min(NO, obj_n - ob);
  const int y0 = blockIdx.y * band, y1 = min(h, y0 + band);
  const int64_t plane4 = (int64_t)h * w4;
  const float4* rp = reinterpret_cast<const float4*>(r1 + (int64_t)ob * r1_obj_stride) + (int64_t)ch * plane4 + x4;
  const float4* sp = reinterpret_cast<const float4*>(seg) + (int64_t)ob * plane4 + x4;
  const float4* ap = reinterpret_cast<const float4*>(avg) + (int64_t)ob * plane4 + x4;
  float4* out = reinterpret_cast<float4*>(lm) + ((int64_t)ob * 2 * c_n + ch) * plane4 + x4;
  const int64_t obj_out4 = (int64_t)2 * c_n * plane4, loc4 = (int64_t)c_n * plane4;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // ring of the last row sums: EIGHT slots for a seven-row window, and the row loop unrolled by eight, so that the ring
  // slot (u), the window order ((u + 2) % 8 .. u, oldest first) and the double-buffered prefetch (u & 1) are all
  // compile-time indices: nothing rotates through registers (a rolled loop spent 100 MOVs per row on that)
  float4 ring[NO][8];
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int r = 0; r < 8; ++r) ring[o][r] = z4;

  // running row pointers (advanced by one row per iteration): the 64-bit index products of every load and store of a
  // row were a quarter of the kernel's instructions
  const int y_end = y1 + UHALO;
  int yin = y0 - UHALO;
  const float4* rrow = rp + (int64_t)yin * w4;            // row `yin` of r1 (may point outside the plane: never read then)
  const float4* srow[NO];
  const float4* arow[NO];
  float4* orow[NO];                                       // raw copy of row `yin`; the local row is 3 rows up, + loc4
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    srow[o] = sp + (int64_t)o * plane4 + (int64_t)yin * w4;
    arow[o] = ap + (int64_t)o * plane4 + (int64_t)(yin - UHALO) * w4;
    orow[o] = out + (int64_t)o * obj_out4 + (int64_t)yin * w4;
  }
  const int64_t loc_off = loc4 - (int64_t)UHALO * w4;
  auto load_row = [&](int y, int ahead, float4& rv, float4 (&sv)[NO]) {      // row y = yin + ahead
    const bool ok = col_ok && y >= 0 && y < h;
    rv = ok ? __ldg(rrow + (int64_t)ahead * w4) : z4;
#pragma unroll
    for (int o = 0; o < NO; ++o) sv[o] = (ok && o < no) ? __ldg(srow[o] + (int64_t)ahead * w4) : z4;
  };
  float4 rbuf[2], sbuf[2][NO];
  load_row(yin, 0, rbuf[0], sbuf[0]);
  while (yin < y_end) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (yin < y_end) {                                               // warp-uniform; only the last group stops early
        load_row(yin + 1 < y_end ? yin + 1 : -1, 1, rbuf[(u + 1) & 1], sbuf[(u + 1) & 1]);   // prefetch the next row
        const float4 rv = rbuf[u & 1];
        const int yout = yin - UHALO;
        const bool emit = out_lane && yout >= y0;
        float4 av[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o)
          av[o] = (emit && o < no) ? __ldg(arow[o]) : z4;
        if (out_lane && yin >= y0 && yin < y1) {     // the raw copy [r1 ; .] of the row just loaded (AFB_URR.py:231)
#pragma unroll
          for (int o = 0; o < NO; ++o)
            if (o < no) __stcs(orow[o], rv);
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          // __fmul_rn: the products must round before the tap sums (no FMA contraction), as in the tiled kernel
          const float4 sv = sbuf[u & 1][o];
          const float4 c = make_float4(__fmul_rn(rv.x, sv.x), __fmul_rn(rv.y, sv.y), __fmul_rn(rv.z, sv.z),
                                       __fmul_rn(rv.w, sv.w));   // AFB_URR.py:226
          const float ly = __shfl_up_sync(0xffffffffu, c.y, 1), lz = __shfl_up_sync(0xffffffffu, c.z, 1),
                      lw = __shfl_up_sync(0xffffffffu, c.w, 1);
          const float rx = __shfl_down_sync(0xffffffffu, c.x, 1), ry = __shfl_down_sync(0xffffffffu, c.y, 1),
                      rz = __shfl_down_sync(0xffffffffu, c.z, 1);
          float4 hs;                                   // 7 taps, left to right (the order of urr_local_kernel)
          hs.x = (((((ly + lz) + lw) + c.x) + c.y) + c.z) + c.w;
          hs.y = (((((lz + lw) + c.x) + c.y) + c.z) + c.w) + rx;
          hs.z = (((((lw + c.x) + c.y) + c.z) + c.w) + rx) + ry;
          hs.w = (((((c.x + c.y) + c.z) + c.w) + rx) + ry) + rz;
          ring[o][u] = hs;
        }
        if (emit) {
#pragma unroll
          for (int o = 0; o < NO; ++o) {
            if (o < no) {
              float4 tot = ring[o][(u + 2) % 8];       // rows top to bottom, as in the tiled kernel
#pragma unroll
              for (int r = 3; r <= 8; ++r) {
                const float4 a = ring[o][(u + r) % 8];
                tot.x += a.x; tot.y += a.y; tot.z += a.z; tot.w += a.w;
              }
              const float4 res = make_float4(urr_ratio(tot.x, av[o].x), urr_ratio(tot.y, av[o].y),
                                             urr_ratio(tot.z, av[o].z), urr_ratio(tot.w, av[o].w));
              __stcs(orow[o] + loc_off, res);
            }
          }
        }
        ++yin;
        rrow += w4;
#pragma unroll
        for (int o = 0; o < NO; ++o) { srow[o] += w4; arow[o] += w4; orow[o] += w4; }
      }
    }
  }
}

// post: prob[o][Y][X] = softmax_c( bilinear_x2( p_up + unc * (conf * q_local) ) )[1]      AFB_URR.py:233-237
__global__ void urr_post_kernel(const float* __restrict__ p_up, const float* __restrict__ unc,
                                const float* __restrict__ conf, const float* __restrict__ ql, int obj_n, int h, int w,
                                float* __restrict__ prob) {
  pdl_wait();
  pdl_trigger();
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, o = blockIdx.z;
  const int H = 2 * h, W = 2 * w;
  if (X >= W) return;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(Y, h, y0, y1, ly0, ly1);
  src_index(X, w, x0, x1, lx0, lx1);
  const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
  float v[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int64_t off = (int64_t)ys[a] * w + xs[b];
      const float u = unc[off];
      const float cf = conf[(int64_t)o * h * w + off];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t po = ((int64_t)o * 2 + c) * h * w + off;
        const float q = cf * ql[po];
        v[c][a][b] = p_up[po] + u * q;
      }
    }
  float r[2];
#pragma unroll
  for (int c = 0; c < 2; ++c)
    r[c] = ly0 * (lx0 * v[c][0][0] + lx1 * v[c][0][1]) + ly1 * (lx0 * v[c][1][0] + lx1 * v[c][1][1]);
  const float m = fmaxf(r[0], r[1]);
  const float e0 = expf(r[0] - m), e1 = expf(r[1] - m);
  prob[(int64_t)o * H * W + (int64_t)Y * W + X] = e1 / (e0 + e1);
}

}  // namespace vfn

using namespace vfn;

namespace vfn { int g_pdl = 1; }
static int g_urr_stream = 1;   // vfn_debug_set_urr_stream(0): tiled shared-memory kernel (cross-check in tests/)

extern "C" {

int vfn_debug_set_pdl(int32_t on) {
  vfn::g_pdl = on ? 1 : 0;
  return VFN_OK;
}

int vfn_debug_set_urr_stream(int32_t on) {
  g_urr_stream = on;      // 0 tiled, 1 streaming (two objects per warp), 2 streaming (one object per warp)
  return VFN_OK;
}

int vfn_urr_pre(const float* d_p, const float* d_r1, int64_t r1_obj_stride, int32_t obj_n, int32_t c, int32_t h,
                int32_t w, float* d_p_up, float* d_seg, float* d_unc, float* d_conf, float* d_avg,
                float* d_local_match, void* stream) {
  VFN_CHECK_ARG(d_p && d_r1 && d_p_up && d_seg && d_unc && d_conf && d_avg && d_local_match, "urr_pre: NULL argument");
  VFN_CHECK_ARG(obj_n >= 1 && obj_n <= URR_MAX_OBJ, "urr_pre: obj_n=%d out of range", obj_n);
  VFN_CHECK_ARG(h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "urr_pre: h,w must be even");
  VFN_CHECK_ARG(c > 0, "urr_pre: channels must be positive");
  cudaStream_t st = as_stream(stream);
  if (g_urr_stream == 0) {       // the two separate launches: cross-check of the fused stage 1 + 2 kernel in tests/
    dim3 g1((unsigned)cdiv(w, 128), h);
    launch_pdl(urr_seg_kernel, g1, dim3(128), 0, st, d_p, obj_n, h, w, d_p_up, d_seg, d_unc);
    dim3 g2((unsigned)cdiv(w, 128), h, obj_n);
    launch_pdl(urr_window_kernel, g2, dim3(128), 0, st, d_seg, obj_n, h, w, d_conf, d_avg);
    count_launches(1);
  } else {
    dim3 g12((unsigned)cdiv(w, SW_W), (unsigned)cdiv(h, SW_H));
    launch_pdl(urr_segwin_kernel, g12, dim3(256), 0, st, d_p, obj_n, h, w, d_p_up, d_seg, d_unc, d_conf, d_avg);
  }
  dim3 g3((unsigned)cdiv(w, UT_W), (unsigned)cdiv(h, UT_H), c);
  prof_begin(PROF_URR, st);
  if (w % 4 == 0 && g_urr_stream) {
    // streaming kernel: bands of rows sized so that the grid holds a few CTAs per SM; halo cost 6 / band input rows
    // band height: one wave of CTAs (4 resident per SM at 119 registers) where that keeps bands >= 8 rows
    // objects per warp: 2 shares one r1 load between two objects (128 registers, 4 CTAs per SM); 1 gives each object
    // its own warp (80 registers, 6 CTAs per SM; the second read of r1 is an L2 hit: r1 is 26 MB)
    const int no = (r1_obj_stride == 0 && obj_n >= 2 && g_urr_stream != 2) ? 2 : 1;
    const int resident = no == 2 ? 4 : 6;
    const int64_t per_band = cdiv(w / 4, UL_COLS) * cdiv(c, UL_WARPS) * cdiv(obj_n, no);
    int sms = 0, dev = 0;
    VFN_CUDA_OK(cudaGetDevice(&dev));
    VFN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t n_bands = (int64_t)sms * resident / per_band;
    if (n_bands < 1) n_bands = 1;
    int band = (int)cdiv(h, n_bands);
    if (band < 8) band = 8;
    dim3 gs((unsigned)cdiv(w / 4, UL_COLS), (unsigned)cdiv(h, band), (unsigned)(cdiv(c, UL_WARPS) * cdiv(obj_n, no)));
    if (no == 2)
      launch_pdl(urr_local_stream_kernel<2>, gs, dim3(UL_WARPS * 32), 0, st, d_r1, r1_obj_stride, c, obj_n, h, w, band, d_seg, d_avg, d_local_match);
    else
      launch_pdl(urr_local_stream_kernel<1>, gs, dim3(UL_WARPS * 32), 0, st, d_r1, r1_obj_stride, c, obj_n, h, w, band, d_seg, d_avg, d_local_match);
  } else {
    launch_pdl(urr_local_kernel, g3, dim3(URR_LOCAL_THREADS), 0, st, d_r1, r1_obj_stride, c, obj_n, h, w, d_seg, d_avg, d_local_match);
  }
  // algorithmic bytes: read r1 once, write [r1 ; r1_local] per object, + small planes
  prof_end(PROF_URR, st, 4.0 * (double)h * w * ((r1_obj_stride ? obj_n : 1) * (double)c + obj_n * (2.0 * c + 8.0)));
  VFN_LAUNCH_OK();
  count_launches(2);
  return VFN_OK;
}

int vfn_urr_post(const float* d_p_up, const float* d_unc, const float* d_conf, const float* d_q_local, int32_t obj_n,
                 int32_t h, int32_t w, float* d_prob, void* stream) {
  VFN_CHECK_ARG(d_p_up && d_unc && d_conf && d_q_local && d_prob, "urr_post: NULL argument");
  VFN_CHECK_ARG(obj_n >= 1 && h > 0 && w > 0, "urr_post: bad shape");
  dim3 g((unsigned)cdiv(2 * w, 128), 2 * h, obj_n);
  launch_pdl(urr_post_kernel, g, dim3(128), 0, as_stream(stream), d_p_up, d_unc, d_conf, d_q_local, obj_n, h, w, d_prob);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

}  // extern "C"
