// Frame-loop tail on the device: what the reference does to every frame after fb.update,
// on the host, with a full-resolution mask copied back first:
//   test_video_seg.py:114-115   pred = argmax(TF.resize(pred_mask, ori_size, BICUBIC)[0], dim=0)      (torch + D2H)
//   myutils/data.py:19-39       postprocessing_pred: 8-connected components (cv2 CCL_GRANA), largest one kept (CPU)
//   estimation/reference_tracking.py:190-204   first water pixel below each key point -> level in pixels   (CPU)
// Here: one interpolation+argmax kernel, a union-find labelling (label-equivalence, atomicMin roots) whose ids are
// ordered like cv2's labels so that ties between equally large components resolve identically, and a column scan.
// All HBM-bound byte/integer work on 1-8 M pixels; nothing returns to the host but the levels (and the mask on request).
#include "vfn_common.cuh"

namespace vfn {

__device__ __forceinline__ int pix_id(int x, int y, int wb) { return (((y >> 1) * wb + (x >> 1)) << 2) | ((y & 1) << 1) | (x & 1); }

// ------------------------------------------------------------------------------------------------------------------
// bicubic resize + argmax
// ------------------------------------------------------------------------------------------------------------------
// antialias = 1: Pillow-style cubic (a = -0.5), window [center - support, center + support) truncated at the borders
//                and re-normalised (what TF.resize does on tensors in torchvision >= 0.17; F.interpolate(antialias=True))
// antialias = 0: classic cubic convolution (a = -0.75), indices clamped (torchvision 0.9.1, the reference's README pin)
__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}

struct AxisAA {
  int lo, size;
  float center_off;   // lo - center
  float invscale, total;
};

__device__ __forceinline__ AxisAA axis_aa(int i, int in_size, float scale) {
  AxisAA r;
  const float support = (scale >= 1.f) ? 2.f * scale : 2.f;
  // every step rounded on its own (no FMA contraction): near the right border `center` is ~10^3 and half an ulp of an
  // unrounded product moves the weights by 1e-5 against ATen's separately rounded float arithmetic
  const float center = __fmul_rn(scale, i + 0.5f);
  r.lo = max((int)__fadd_rn(__fsub_rn(center, support), 0.5f), 0);
  r.size = min((int)__fadd_rn(__fadd_rn(center, support), 0.5f), in_size) - r.lo;
  r.center_off = __fsub_rn((float)r.lo, center);
  r.invscale = (scale >= 1.f) ? 1.f / scale : 1.f;
  float t = 0.f;
  for (int j = 0; j < r.size; ++j) t += cubic_aa(__fmul_rn(__fadd_rn(__fadd_rn((float)j, r.center_off), 0.5f), r.invscale));
  r.total = t;
  return r;
}
__device__ __forceinline__ float axis_w(const AxisAA& a, int j) {
  const float w = cubic_aa(__fmul_rn(__fadd_rn(__fadd_rn((float)j, a.center_off), 0.5f), a.invscale));
  return a.total != 0.f ? w / a.total : w;
}

__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  float x = t + 1.f;
  c[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  c[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  c[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

constexpr int TROWS = 4;   // rows per CTA: a CTA owns a 256-column x TROWS-row tile

// one axis of the antialias window for the up-scaling case (<= 4 taps): normalised weights
struct Axis4 { int lo, size; float w[4]; };
__device__ __forceinline__ Axis4 axis4(int i, int in_size, float scale) {
  const AxisAA a = axis_aa(i, in_size, scale);
  Axis4 r;
  r.lo = a.lo;
  r.size = a.size;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.w[j] = j < a.size ? axis_w(a, j) : 0.f;
  return r;
}

// forest initialisation of one row span from the CTA's foreground ballots (see cc_init_kernel)
__device__ __forceinline__ int run_parent(bool fg, unsigned bits, const unsigned* fg_bits, int warp, int lane, int x,
                                          int x0, int y, int wb, bool left_of_span_fg) {
  const int id = pix_id(x, y, wb);
  if (!fg) return -1;
  int start = x0;
  unsigned zeros = ~bits & ((1u << lane) - 1u);
  int k = warp;
  while (true) {
    if (zeros) { start = x0 + 32 * k + (32 - __clz(zeros)); break; }
    if (--k < 0) break;
    zeros = ~fg_bits[k];
  }
  if (start < x) return pix_id(start, y, wb);
  return (x == x0 && left_of_span_fg) ? pix_id(x0 - 1, y, wb) : id;
}

// FUSE: also initialise the labelling forest (lab / size) from the arg-max just computed (vfn_frame_tail)
template <bool AA, bool FUSE>
__global__ void __launch_bounds__(256) resize_argmax_kernel(const float* __restrict__ src, int obj_n, int h, int w, int H,
                                                            int W, float scale_y, float scale_x,
                                                            uint8_t* __restrict__ pred, int wb, int* __restrict__ lab,
                                                            int* __restrict__ size) {
  __shared__ Axis4 ys[TROWS];
  __shared__ float wy_tab[TROWS][8];
  __shared__ unsigned fg_bits[TROWS][8];
  const int x0 = blockIdx.x * 256, X = x0 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t plane = (size_t)h * w;
  const bool in = X < W;
  const bool fast = !AA || (scale_x < 1.f && scale_y < 1.f);   // up-scaling: windows of <= 4 taps
  int cls[TROWS];
  if (AA && fast) {
    // Up-scaling: the TROWS output rows of the tile read at most 8 distinct input rows.  Each input row is filtered
    // horizontally once (4 taps) and every output row then takes its 4-tap vertical combination from a per-tile weight
    // table that is zero outside the row's window: x + 0*v leaves the sum of the real taps bit-identical to ATen's
    // "horizontal, then vertical" order, and there are no predicates in the inner loops.  Taps beyond a window
    // truncated by the image border get weight 0 and a clamped (valid) address.
    if (threadIdx.x < TROWS) {
      Axis4 a = axis4(min(Y0 + (int)threadIdx.x, H - 1), h, scale_y);
      ys[threadIdx.x] = a;
    }
    __syncthreads();
    const int row_lo = ys[0].lo;
    const int n_rows = min(ys[min(TROWS, H - Y0) - 1].lo + 4, h) - row_lo;      // <= 8 for scale < 1
    if (threadIdx.x < TROWS * 8) {
      const int r = threadIdx.x >> 3, k = threadIdx.x & 7, j = row_lo + k - ys[r].lo;
      wy_tab[r][k] = (j >= 0 && j < ys[r].size) ? ys[r].w[j] : 0.f;
    }
    __syncthreads();
    const Axis4 ax = axis4(min(X, W - 1), w, scale_x);
    int xo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xo[j] = min(ax.lo + j, w - 1);
    float best[TROWS];
#pragma unroll
    for (int r = 0; r < TROWS; ++r) { cls[r] = 0; best[r] = 0.f; }
    for (int c = 0; c < obj_n; ++c) {
      const float* p = src + c * plane + (size_t)row_lo * w;
      float out[TROWS];
#pragma unroll
      for (int r = 0; r < TROWS; ++r) out[r] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < n_rows) {                                   // uniform over the CTA
          const float* q = p + k * w;
          float v = __ldg(q + xo[0]) * ax.w[0];
          v += __ldg(q + xo[1]) * ax.w[1];
          v += __ldg(q + xo[2]) * ax.w[2];
          v += __ldg(q + xo[3]) * ax.w[3];
#pragma unroll
          for (int r = 0; r < TROWS; ++r) out[r] += v * wy_tab[r][k];
        }
      }
#pragma unroll
      for (int r = 0; r < TROWS; ++r)
        if (c == 0 || out[r] > best[r]) { best[r] = out[r]; cls[r] = c; }
    }
  } else if (AA) {                         // down-scaling: wide windows, weights recomputed on the fly
    const AxisAA ax = axis_aa(min(X, W - 1), w, scale_x);
    for (int r = 0; r < TROWS; ++r) {
      cls[r] = 0;
      if (Y0 + r >= H || !in) continue;
      const AxisAA ay = axis_aa(Y0 + r, h, scale_y);
      float best = 0.f;
      for (int c = 0; c < obj_n; ++c) {
        const float* p = src + c * plane + (size_t)ay.lo * w + ax.lo;
        float out = 0.f;
        for (int jy = 0; jy < ay.size; ++jy) {
          float v = __ldg(p + (size_t)jy * w) * axis_w(ax, 0);
          for (int jx = 1; jx < ax.size; ++jx) v += __ldg(p + (size_t)jy * w + jx) * axis_w(ax, jx);
          const float wyj = axis_w(ay, jy);
          out = jy == 0 ? v * wyj : out + v * wyj;
        }
        if (c == 0 || out > best) { best = out; cls[r] = c; }
      }
    }
  } else {
    const float rx = __fsub_rn(__fmul_rn(scale_x, min(X, W - 1) + 0.5f), 0.5f);
    const float fx = floorf(rx);
    const int ix = (int)fx;
    float cx[4];
    cubic_coeffs(rx - fx, cx);
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = min(max(ix - 1 + j, 0), w - 1);
    for (int r = 0; r < TROWS; ++r) {
      cls[r] = 0;
      if (Y0 + r >= H || !in) continue;
      const float ry = __fsub_rn(__fmul_rn(scale_y, Y0 + r + 0.5f), 0.5f);
      const float fy = floorf(ry);
      const int iy = (int)fy;
      float cy[4];
      cubic_coeffs(ry - fy, cy);
      float best = 0.f;
      for (int c = 0; c < obj_n; ++c) {
        const float* p = src + c * plane;
        float rows[4];
#pragma unroll
        for (int jy = 0; jy < 4; ++jy) {
          const float* q = p + (size_t)min(max(iy - 1 + jy, 0), h - 1) * w;
          rows[jy] = __ldg(q + xs[0]) * cx[0] + __ldg(q + xs[1]) * cx[1] + __ldg(q + xs[2]) * cx[2] + __ldg(q + xs[3]) * cx[3];
        }
        const float out = rows[0] * cy[0] + rows[1] * cy[1] + rows[2] * cy[2] + rows[3] * cy[3];
        if (c == 0 || out > best) { best = out; cls[r] = c; }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < TROWS; ++r)
    if (in && Y0 + r < H) pred[(size_t)(Y0 + r) * W + X] = (uint8_t)cls[r];
  if (!FUSE) return;
  // forest initialisation.  Whether a run continues from the previous span is only known to the CTA on the left, so a
  // span's first pixel always links to itself here and cc_merge_kernel joins it with its left neighbour.
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    const bool fg = in && Y0 + r < H && cls[r] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    if (lane == 0) fg_bits[r][warp] = bits;
    cls[r] = (int)bits;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    if (!in || Y0 + r >= H) continue;
    const unsigned bits = (unsigned)cls[r];
    const bool fg = (bits >> lane) & 1u;
    const int id = pix_id(X, Y0 + r, wb);
    lab[id] = run_parent(fg, bits, fg_bits[r], warp, lane, X, x0, Y0 + r, wb, false);
    size[id] = 0;
  }
}

// window + normalised weights of every output column (wx[0..W)) and row (wy[0..H)): they depend on the sizes only, so
// the per-pixel kernel reads 24 bytes instead of evaluating 8 cubics and 4 divisions per thread
__global__ void __launch_bounds__(256) tail_weights_kernel(int h, int w, int H, int W, float scale_y, float scale_x,
                                                           Axis4* __restrict__ wx, Axis4* __restrict__ wy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) wx[i] = axis4(i, w, scale_x);
  else if (i < W + H) wy[i - W] = axis4(i - W, h, scale_y);
}

// antialias up-scaling with tabulated weights; always initialises the labelling forest (vfn_frame_tail).
// Same arithmetic as resize_argmax_kernel<true, *>'s up-scaling branch.
__global__ void __launch_bounds__(256, 3) resize_argmax_tab_kernel(const float* __restrict__ src, int obj_n, int h, int w,
                                                                int H, int W, const Axis4* __restrict__ wx,
                                                                const Axis4* __restrict__ wy, uint8_t* __restrict__ pred,
                                                                int wb, int* __restrict__ lab, int* __restrict__ size) {
  __shared__ float wy_tab[TROWS][8];
  __shared__ unsigned fg_bits[TROWS][8];
  __shared__ int s_rows[2];
  const int x0 = blockIdx.x * 256, X = x0 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = min(TROWS, H - Y0);
  if (threadIdx.x < TROWS * 8) {
    const int r = threadIdx.x >> 3, k = threadIdx.x & 7;
    const int row_lo = wy[Y0].lo;
    float v = 0.f;
    if (r < n_out) {
      const Axis4 a = wy[Y0 + r];
      const int j = row_lo + k - a.lo;
      if (j >= 0 && j < a.size) v = a.w[j];
      if (r == n_out - 1 && k == 0) { s_rows[0] = row_lo; s_rows[1] = min(a.lo + 4, h) - row_lo; }
    }
    wy_tab[r][k] = v;
  }
  __syncthreads();
  const int row_lo = s_rows[0], n_rows = s_rows[1];
  const bool in = X < W;
  const Axis4 ax = wx[min(X, W - 1)];
  int xo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) xo[j] = min(ax.lo + j, w - 1);
  float best[TROWS];
  int cls[TROWS];
#pragma unroll
  for (int r = 0; r < TROWS; ++r) { cls[r] = 0; best[r] = 0.f; }
  const size_t plane = (size_t)h * w;
  for (int c = 0; c < obj_n; ++c) {
    const float* p = src + c * plane + (size_t)row_lo * w;
    float out[TROWS];
#pragma unroll
    for (int r = 0; r < TROWS; ++r) out[r] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < n_rows) {                                   // uniform over the CTA
        const float* q = p + k * w;
        float v = __ldg(q + xo[0]) * ax.w[0];
        v += __ldg(q + xo[1]) * ax.w[1];
        v += __ldg(q + xo[2]) * ax.w[2];
        v += __ldg(q + xo[3]) * ax.w[3];
#pragma unroll
        for (int r = 0; r < TROWS; ++r) out[r] += v * wy_tab[r][k];
      }
    }
#pragma unroll
    for (int r = 0; r < TROWS; ++r)
      if (c == 0 || out[r] > best[r]) { best[r] = out[r]; cls[r] = c; }
  }
  unsigned bits[TROWS];
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    const bool fg = in && r < n_out && cls[r] != 0;
    if (in && r < n_out) pred[(size_t)(Y0 + r) * W + X] = (uint8_t)cls[r];
    bits[r] = __ballot_sync(0xffffffffu, fg);
    if (lane == 0) fg_bits[r][warp] = bits[r];
  }
  __syncthreads();
  if (!in) return;
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    if (r < n_out) {
      const int id = pix_id(X, Y0 + r, wb);
      lab[id] = run_parent((bits[r] >> lane) & 1u, bits[r], fg_bits[r], warp, lane, X, x0, Y0 + r, wb, false);
      size[id] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// largest 8-connected component
// ------------------------------------------------------------------------------------------------------------------
// Pixel ids are block-raster: id = ((y/2) * ceil(W/2) + x/2) * 4 + (y&1)*2 + (x&1).  A component's root is its minimum
// id, i.e. it encodes the first 2x2 block the component touches - the order in which cv2's CCL_GRANA numbers its labels
// (two components never share a 2x2 block under 8-connectivity), so "largest, ties to the lowest label"
// (myutils/data.py:28-37) becomes "largest, ties to the lowest root id".
__device__ __forceinline__ int uf_find(const int* lab, int i) {
  int p;
  while ((p = __ldcg(lab + i)) != i) i = p;
  return i;
}
__device__ int g_tail_flags = 0;   // vfn_debug_set_tail: bit 1 = no per-CTA size aggregation (cross-check)
// Lock-free union: the larger root is linked under the smaller one with a CAS that only succeeds while it still is a
// root, so pointers of non-roots never change and "x is an ancestor of y" stays true for ever.  (A first version linked
// with atomicMin, which can re-point a non-root at a root of a set that is not united yet; with path compression on top
// that lost unions - tests/debug_tail_bisect.py.  Compression itself was measured and dropped: the forest of a frame is
// shallow, the kernel's critical path is the handful of dependent L2 round trips per union, profiles/r1k.)
__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
  while (true) {
    a = uf_find(lab, a);
    b = uf_find(lab, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    if (atomicCAS(lab + b, b, a) == b) return;      // b was still a root: linked under the smaller root
  }
}

// parent = first pixel of the horizontal run inside this CTA's 256-pixel span (found from the warps' ballots, so a find
// never walks along a run); parents always carry a smaller id.  Runs crossing a span border are joined by
// cc_merge_kernel.  Sizes zeroed.
__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ pred, int H, int W, int wb,
                                                      int* __restrict__ lab, int* __restrict__ size) {
  __shared__ unsigned fg_bits[TROWS][8];
  const int x0 = blockIdx.x * 256, x = x0 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned bits[TROWS];
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    const bool fg = x < W && Y0 + r < H && pred[(size_t)(Y0 + r) * W + x] != 0;
    bits[r] = __ballot_sync(0xffffffffu, fg);
    if (lane == 0) fg_bits[r][warp] = bits[r];
  }
  __syncthreads();
  if (x >= W) return;
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    if (Y0 + r >= H) break;
    const int id = pix_id(x, Y0 + r, wb);
    lab[id] = run_parent((bits[r] >> lane) & 1u, bits[r], fg_bits[r], warp, lane, x, x0, Y0 + r, wb, false);
    size[id] = 0;
  }
}

// join every run with the row above and, at span borders, with its continuation on the left.  N is enough when it is
// set (NW / NE then sit in N's run); a pixel whose W and NW are both set is already joined through them.
__global__ void __launch_bounds__(256) cc_merge_kernel(const uint8_t* __restrict__ pred, int H, int W, int wb,
                                                       int* __restrict__ lab) {
  const int x = blockIdx.x * 256 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  if (x >= W) return;
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    const int y = Y0 + r;
    if (y >= H) break;
    const uint8_t* row = pred + (size_t)y * W;
    if (!row[x]) continue;
    const bool wl = x > 0 && row[x - 1];
    const int id = pix_id(x, y, wb);
    if (threadIdx.x == 0 && wl) uf_union(lab, id, pix_id(x - 1, y, wb));
    if (y == 0) continue;
    const uint8_t* up = row - W;
    const bool n = up[x], nw = x > 0 && up[x - 1], ne = x + 1 < W && up[x + 1];
    if (n) {
      if (!(wl && nw)) uf_union(lab, id, pix_id(x, y - 1, wb));
    } else {
      if (nw && !wl) uf_union(lab, id, pix_id(x - 1, y - 1, wb));
      if (ne) uf_union(lab, id, pix_id(x + 1, y - 1, wb));
    }
  }
}

// flatten + component sizes.  Only the first pixel of a run walks to the root (the other pixels of the run point at it
// and it sits in the same CTA).  The CTA elects one root (its first foreground pixel's) and counts that one in shared
// memory - for a mask made of a few water bodies that is nearly every pixel of the tile; other roots go to global
// memory with one atomic per distinct root in a warp.
__global__ void __launch_bounds__(256) cc_count_kernel(const uint8_t* __restrict__ pred, int H, int W, int wb,
                                                       int* __restrict__ lab, int* __restrict__ size) {
  __shared__ int cand, cand_cnt;
  const int x = blockIdx.x * 256 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  if (threadIdx.x == 0) { cand = -1; cand_cnt = 0; }
  int root[TROWS], id[TROWS];
  bool start[TROWS];
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    root[r] = -1;
    id[r] = 0;
    start[r] = false;
    if (x < W && Y0 + r < H) {
      const uint8_t* row = pred + (size_t)(Y0 + r) * W;
      if (row[x]) {
        id[r] = pix_id(x, Y0 + r, wb);
        start[r] = threadIdx.x == 0 || !row[x - 1];
        if (start[r]) {
          root[r] = uf_find(lab, id[r]);
          lab[id[r]] = root[r];
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < TROWS; ++r)
    if (x < W && Y0 + r < H && id[r] && !start[r]) {
      root[r] = __ldcg(lab + __ldcg(lab + id[r]));      // my run start (same CTA, flattened above) -> its root
      lab[id[r]] = root[r];
    }
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    const unsigned any = __ballot_sync(0xffffffffu, root[r] >= 0);
    if (any && (threadIdx.x & 31) == __ffs(any) - 1 && *(volatile int*)&cand < 0) atomicCAS(&cand, -1, root[r]);
  }
  __syncthreads();
  const int c = (g_tail_flags & 2) ? -2 : cand;
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    const unsigned peers = __match_any_sync(0xffffffffu, root[r]);
    if (root[r] >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) {
      if (root[r] == c) atomicAdd(&cand_cnt, __popc(peers));
      else atomicAdd(size + root[r], __popc(peers));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && cand_cnt && c >= 0) atomicAdd(size + c, cand_cnt);
}

// stats: [0] foreground pixels, [1] components, [2] size of the kept component, [3] its root id (-1: none)
__global__ void __launch_bounds__(256) cc_select_kernel(int H, int W, int wb, const int* __restrict__ lab,
                                                        const int* __restrict__ size, unsigned long long* best,
                                                        int* __restrict__ stats) {
  __shared__ unsigned long long s_key[8];
  __shared__ int s_fg[8], s_root[8];
  const int x = blockIdx.x * 256 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  unsigned long long key = 0;
  int n_fg = 0, n_root = 0;
#pragma unroll
  for (int r = 0; r < TROWS; ++r) {
    if (x < W && Y0 + r < H) {
      const int id = pix_id(x, Y0 + r, wb);
      const int l = lab[id];
      n_fg += l >= 0;
      if (l == id) {
        ++n_root;
        const unsigned long long k = ((unsigned long long)(unsigned)size[id] << 32) | (0xffffffffu - (unsigned)id);
        key = k > key ? k : key;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
    n_fg += __shfl_xor_sync(0xffffffffu, n_fg, o);
    n_root += __shfl_xor_sync(0xffffffffu, n_root, o);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_key[warp] = key; s_fg[warp] = n_fg; s_root[warp] = n_root; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      key = s_key[k] > key ? s_key[k] : key;
      n_fg += s_fg[k];
      n_root += s_root[k];
    }
    if (n_fg) atomicAdd(stats + 0, n_fg);
    if (n_root) {
      atomicAdd(stats + 1, n_root);
      atomicMax(best, key);
    }
  }
}

// mask = (label == kept root); an EMPTY prediction gives an all-ones mask, as the reference's loop does
// (label_cnt == 1: max_label stays 0 and `labels == 0` is true everywhere, myutils/data.py:27-37)
__global__ void __launch_bounds__(256) cc_write_kernel(int H, int W, int wb, const int* __restrict__ lab,
                                                       const unsigned long long* __restrict__ best,
                                                       uint8_t* __restrict__ mask, int* __restrict__ stats) {
  const int x = blockIdx.x * 256 + threadIdx.x, Y0 = blockIdx.y * TROWS;
  const unsigned long long b = *best;
  const int keep = b ? (int)(0xffffffffu - (unsigned)(b & 0xffffffffu)) : -1;
  if (x == 0 && Y0 == 0) {
    stats[2] = (int)(b >> 32);
    stats[3] = keep;
  }
  if (x >= W) return;
#pragma unroll
  for (int r = 0; r < TROWS; ++r)
    if (Y0 + r < H) mask[(size_t)(Y0 + r) * W + x] = b ? (lab[pix_id(x, Y0 + r, wb)] == keep) : 1;
}

// ------------------------------------------------------------------------------------------------------------------
// water level: first pixel with the water label strictly below the key point, in its column
// ------------------------------------------------------------------------------------------------------------------
// level[t] is in/out: a column without water keeps the previous frame's estimate (reference_tracking.py:188);
// a level of exactly 1 is recorded as NaN (:198-199).
__global__ void __launch_bounds__(256) waterlevel_kernel(const uint8_t* __restrict__ mask, int H, int W,
                                                         const int32_t* __restrict__ key_pts, int water_id,
                                                         float* __restrict__ level) {
  __shared__ int first;
  const int t = blockIdx.x;
  const int kx = key_pts[2 * t], ky = key_pts[2 * t + 1];
  if (threadIdx.x == 0) first = INT_MAX;
  __syncthreads();
  if (kx >= 0 && kx < W) {
    for (int base = max(ky + 1, 0); base < H; base += blockDim.x) {
      const int y = base + threadIdx.x;
      const bool hit = y < H && mask[(size_t)y * W + kx] == water_id;
      if (__syncthreads_or(hit)) {
        if (hit) atomicMin(&first, y);
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && first != INT_MAX) {
    const int d = first - ky;
    level[t] = d == 1 ? __int_as_float(0x7fc00000) : (float)d;
  }
}

}  // namespace vfn

using namespace vfn;

namespace {
struct TailWs { int* lab; int* size; unsigned long long* best; Axis4* wx; Axis4* wy; };
size_t tail_ws_layout(int H, int W, char* base, TailWs* out) {
  const size_t n = (size_t)4 * ((H + 1) / 2) * ((W + 1) / 2);
  size_t off = 0;
  if (out) out->lab = reinterpret_cast<int*>(base + off);
  off += align_up(n * sizeof(int), 256);
  if (out) out->size = reinterpret_cast<int*>(base + off);
  off += align_up(n * sizeof(int), 256);
  if (out) out->best = reinterpret_cast<unsigned long long*>(base + off);
  off += 256;
  if (out) out->wx = reinterpret_cast<Axis4*>(base + off);
  off += align_up((size_t)W * sizeof(Axis4), 256);
  if (out) out->wy = reinterpret_cast<Axis4*>(base + off);
  off += align_up((size_t)H * sizeof(Axis4), 256);
  return off;
}
}  // namespace

extern "C" {

int vfn_debug_set_tail(int32_t flags) {
  VFN_CUDA_OK(cudaMemcpyToSymbol(g_tail_flags, &flags, sizeof(int)));
  return VFN_OK;
}

size_t vfn_tail_workspace_bytes(int32_t H, int32_t W) {
  if (H <= 0 || W <= 0) return 0;
  return tail_ws_layout(H, W, nullptr, nullptr);
}

static int launch_resize(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                         int32_t antialias, uint8_t* d_pred, int* lab, int* size, cudaStream_t st) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;   // area_pixel_compute_scale, align_corners=False
  const int wb = (W + 1) / 2;
  dim3 g((unsigned)cdiv(W, 256), (unsigned)cdiv(H, TROWS));
  if (antialias) {
    if (lab) resize_argmax_kernel<true, true><<<g, 256, 0, st>>>(d_pred_mask, obj_n, h, w, H, W, sy, sx, d_pred, wb, lab, size);
    else resize_argmax_kernel<true, false><<<g, 256, 0, st>>>(d_pred_mask, obj_n, h, w, H, W, sy, sx, d_pred, wb, lab, size);
  } else {
    if (lab) resize_argmax_kernel<false, true><<<g, 256, 0, st>>>(d_pred_mask, obj_n, h, w, H, W, sy, sx, d_pred, wb, lab, size);
    else resize_argmax_kernel<false, false><<<g, 256, 0, st>>>(d_pred_mask, obj_n, h, w, H, W, sy, sx, d_pred, wb, lab, size);
  }
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

// labelling from an initialised forest (init_done) or from d_pred
static int launch_components(const uint8_t* d_pred, int32_t H, int32_t W, uint8_t* d_mask, int32_t* d_stats,
                             const TailWs& ws, bool init_done, cudaStream_t st) {
  const int wb = (W + 1) / 2;
  dim3 g((unsigned)cdiv(W, 256), (unsigned)cdiv(H, TROWS));
  VFN_CUDA_OK(cudaMemsetAsync(ws.best, 0, sizeof(unsigned long long), st));
  VFN_CUDA_OK(cudaMemsetAsync(d_stats, 0, 4 * sizeof(int32_t), st));
  if (!init_done) cc_init_kernel<<<g, 256, 0, st>>>(d_pred, H, W, wb, ws.lab, ws.size);
  cc_merge_kernel<<<g, 256, 0, st>>>(d_pred, H, W, wb, ws.lab);
  cc_count_kernel<<<g, 256, 0, st>>>(d_pred, H, W, wb, ws.lab, ws.size);
  cc_select_kernel<<<g, 256, 0, st>>>(H, W, wb, ws.lab, ws.size, ws.best, d_stats);
  cc_write_kernel<<<g, 256, 0, st>>>(H, W, wb, ws.lab, ws.best, d_mask, d_stats);
  VFN_LAUNCH_OK();
  count_launches(init_done ? 4 : 5);
  return VFN_OK;
}

static int check_resize_args(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                             const uint8_t* d_pred) {
  VFN_CHECK_ARG(d_pred_mask && d_pred, "tail_resize_argmax: NULL argument");
  VFN_CHECK_ARG(obj_n >= 1 && obj_n <= 255 && h > 0 && w > 0 && H > 0 && W > 0 && (int64_t)H * W < (1ll << 29) &&
                    (int64_t)h * w < (1ll << 30),
                "tail_resize_argmax: bad shape obj_n=%d (%d,%d)->(%d,%d)", obj_n, h, w, H, W);
  return VFN_OK;
}

int vfn_tail_resize_argmax(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                           int32_t antialias, uint8_t* d_pred, void* stream) {
  if (int rc = check_resize_args(d_pred_mask, obj_n, h, w, H, W, d_pred)) return rc;
  return launch_resize(d_pred_mask, obj_n, h, w, H, W, antialias, d_pred, nullptr, nullptr, as_stream(stream));
}

int vfn_tail_largest_component(const uint8_t* d_pred, int32_t H, int32_t W, uint8_t* d_mask, int32_t* d_stats,
                               void* d_ws, size_t ws_bytes, void* stream) {
  VFN_CHECK_ARG(d_pred && d_mask && d_stats && d_ws, "tail_largest_component: NULL argument");
  VFN_CHECK_ARG(H > 0 && W > 0 && (int64_t)H * W < (1ll << 29), "tail_largest_component: bad shape (%d,%d)", H, W);
  TailWs ws;
  if (ws_bytes < tail_ws_layout(H, W, static_cast<char*>(d_ws), &ws)) {
    set_error("tail_largest_component: workspace %zu < %zu bytes", ws_bytes, tail_ws_layout(H, W, nullptr, nullptr));
    return VFN_E_CAPACITY;
  }
  return launch_components(d_pred, H, W, d_mask, d_stats, ws, false, as_stream(stream));
}

int vfn_tail_waterlevel(const uint8_t* d_mask, int32_t H, int32_t W, const int32_t* d_key_pts, int32_t n_pts,
                        int32_t water_label_id, float* d_level, void* stream) {
  VFN_CHECK_ARG(d_mask && H > 0 && W > 0 && n_pts >= 0, "tail_waterlevel: bad argument");
  if (n_pts == 0) return VFN_OK;
  VFN_CHECK_ARG(d_key_pts && d_level, "tail_waterlevel: NULL argument");
  waterlevel_kernel<<<n_pts, 256, 0, as_stream(stream)>>>(d_mask, H, W, d_key_pts, water_label_id, d_level);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_frame_tail(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                   int32_t antialias, const int32_t* d_key_pts, int32_t n_pts, int32_t water_label_id,
                   uint8_t* d_pred, uint8_t* d_mask, int32_t* d_stats, float* d_level, void* d_ws, size_t ws_bytes,
                   void* stream) {
  if (int rc = check_resize_args(d_pred_mask, obj_n, h, w, H, W, d_pred)) return rc;
  VFN_CHECK_ARG(d_mask && d_stats && d_ws, "frame_tail: NULL argument");
  TailWs ws;
  if (ws_bytes < tail_ws_layout(H, W, static_cast<char*>(d_ws), &ws)) {
    set_error("frame_tail: workspace %zu < %zu bytes", ws_bytes, tail_ws_layout(H, W, nullptr, nullptr));
    return VFN_E_CAPACITY;
  }
  cudaStream_t st = as_stream(stream);
  // the resize kernel also initialises the labelling forest from the arg-max it holds in registers
  int rc = VFN_OK;
  if (antialias && h < H && w < W) {
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    tail_weights_kernel<<<(unsigned)cdiv(H + W, 256), 256, 0, st>>>(h, w, H, W, sy, sx, ws.wx, ws.wy);
    dim3 g((unsigned)cdiv(W, 256), (unsigned)cdiv(H, TROWS));
    resize_argmax_tab_kernel<<<g, 256, 0, st>>>(d_pred_mask, obj_n, h, w, H, W, ws.wx, ws.wy, d_pred, (W + 1) / 2, ws.lab,
                                                ws.size);
    VFN_LAUNCH_OK();
    count_launches(2);
  } else {
    rc = launch_resize(d_pred_mask, obj_n, h, w, H, W, antialias, d_pred, ws.lab, ws.size, st);
  }
  if (rc) return rc;
  rc = launch_components(d_pred, H, W, d_mask, d_stats, ws, true, st);
  if (rc) return rc;
  return vfn_tail_waterlevel(d_mask, H, W, d_key_pts, n_pts, water_label_id, d_level, stream);
}

}  // extern "C"
