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

template <bool AA>
__global__ void __launch_bounds__(256) resize_argmax_kernel(const float* __restrict__ src, int obj_n, int h, int w, int H,
                                                            int W, float scale_y, float scale_x,
                                                            uint8_t* __restrict__ pred) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  if (X >= W) return;
  const size_t plane = (size_t)h * w;
  float best = 0.f;
  int best_c = 0;
  if (AA) {
    const AxisAA ax = axis_aa(X, w, scale_x), ay = axis_aa(Y, h, scale_y);
    if (ax.size <= 4 && ay.size <= 4) {   // every up-scaling case: weights in registers
      float wx[4], wy[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        wx[j] = j < ax.size ? axis_w(ax, j) : 0.f;
        wy[j] = j < ay.size ? axis_w(ay, j) : 0.f;
      }
      for (int c = 0; c < obj_n; ++c) {
        const float* p = src + c * plane + (size_t)ay.lo * w + ax.lo;
        float out = 0.f;
#pragma unroll
        for (int jy = 0; jy < 4; ++jy) {
          if (jy < ay.size) {
            float r = __ldg(p + (size_t)jy * w) * wx[0];
#pragma unroll
            for (int jx = 1; jx < 4; ++jx)
              if (jx < ax.size) r += __ldg(p + (size_t)jy * w + jx) * wx[jx];
            out = jy == 0 ? r * wy[0] : out + r * wy[jy];
          }
        }
        if (c == 0 || out > best) { best = out; best_c = c; }
      }
    } else {                               // down-scaling: wide windows, weights recomputed on the fly
      for (int c = 0; c < obj_n; ++c) {
        const float* p = src + c * plane + (size_t)ay.lo * w + ax.lo;
        float out = 0.f;
        for (int jy = 0; jy < ay.size; ++jy) {
          float r = __ldg(p + (size_t)jy * w) * axis_w(ax, 0);
          for (int jx = 1; jx < ax.size; ++jx) r += __ldg(p + (size_t)jy * w + jx) * axis_w(ax, jx);
          const float wyj = axis_w(ay, jy);
          out = jy == 0 ? r * wyj : out + r * wyj;
        }
        if (c == 0 || out > best) { best = out; best_c = c; }
      }
    }
  } else {
    const float ry = __fsub_rn(__fmul_rn(scale_y, Y + 0.5f), 0.5f), rx = __fsub_rn(__fmul_rn(scale_x, X + 0.5f), 0.5f);
    const float fy = floorf(ry), fx = floorf(rx);
    const int iy = (int)fy, ix = (int)fx;
    float cx[4], cy[4];
    cubic_coeffs(rx - fx, cx);
    cubic_coeffs(ry - fy, cy);
    int xs[4], ys[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xs[j] = min(max(ix - 1 + j, 0), w - 1);
      ys[j] = min(max(iy - 1 + j, 0), h - 1);
    }
    for (int c = 0; c < obj_n; ++c) {
      const float* p = src + c * plane;
      float rows[4];
#pragma unroll
      for (int jy = 0; jy < 4; ++jy) {
        const float* q = p + (size_t)ys[jy] * w;
        rows[jy] = __ldg(q + xs[0]) * cx[0] + __ldg(q + xs[1]) * cx[1] + __ldg(q + xs[2]) * cx[2] + __ldg(q + xs[3]) * cx[3];
      }
      const float out = rows[0] * cy[0] + rows[1] * cy[1] + rows[2] * cy[2] + rows[3] * cy[3];
      if (c == 0 || out > best) { best = out; best_c = c; }
    }
  }
  pred[(size_t)Y * W + X] = (uint8_t)best_c;
}

// ------------------------------------------------------------------------------------------------------------------
// largest 8-connected component
// ------------------------------------------------------------------------------------------------------------------
// Pixel ids are block-raster: id = ((y/2) * ceil(W/2) + x/2) * 4 + (y&1)*2 + (x&1).  A component's root is its minimum
// id, i.e. it encodes the first 2x2 block the component touches - the order in which cv2's CCL_GRANA numbers its labels
// (two components never share a 2x2 block under 8-connectivity), so "largest, ties to the lowest label"
// (myutils/data.py:28-37) becomes "largest, ties to the lowest root id".
__device__ __forceinline__ int pix_id(int x, int y, int wb) { return (((y >> 1) * wb + (x >> 1)) << 2) | ((y & 1) << 1) | (x & 1); }

__device__ __forceinline__ int uf_find(const int* lab, int i) {
  int p;
  while ((p = __ldcg(lab + i)) != i) i = p;
  return i;
}
__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
  bool done;
  do {
    a = uf_find(lab, a);
    b = uf_find(lab, b);
    if (a < b) {
      const int old = atomicMin(lab + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(lab + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// parent = first pixel of the horizontal run inside this CTA's 256-pixel span (found from the warps' ballots, so a find
// never walks along a run); a run that continues from the previous span links its first pixel to the pixel on its
// left.  Parents always carry a smaller id.  Sizes zeroed.
__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ pred, int H, int W, int wb,
                                                      int* __restrict__ lab, int* __restrict__ size) {
  __shared__ unsigned fg_bits[8];
  const int x0 = blockIdx.x * 256, x = x0 + threadIdx.x, y = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* row = pred + (size_t)y * W;
  const bool fg = x < W && row[x] != 0;
  const unsigned bits = __ballot_sync(0xffffffffu, fg);
  if (lane == 0) fg_bits[warp] = bits;
  __syncthreads();
  if (x >= W) return;
  const int id = pix_id(x, y, wb);
  int l = -1;
  if (fg) {
    // nearest background pixel to the left inside the span -> the run starts right after it
    int start = x0;
    unsigned zeros = ~bits & ((1u << lane) - 1u);
    int k = warp;
    while (true) {
      if (zeros) { start = x0 + 32 * k + (32 - __clz(zeros)); break; }
      if (--k < 0) break;
      zeros = ~fg_bits[k];
    }
    if (start < x) l = pix_id(start, y, wb);
    else l = (x == x0 && x0 > 0 && row[x0 - 1]) ? pix_id(x0 - 1, y, wb) : id;
  }
  lab[id] = l;
  size[id] = 0;
}

// join every run with the row above.  N is enough when it is set (NW / NE then sit in N's run); a pixel whose W and NW
// are both set is already joined through them.
__global__ void __launch_bounds__(256) cc_merge_kernel(const uint8_t* __restrict__ pred, int H, int W, int wb,
                                                       int* __restrict__ lab) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + 1;
  if (x >= W || y >= H) return;
  const uint8_t* row = pred + (size_t)y * W;
  const uint8_t* up = row - W;
  if (!row[x]) return;
  const bool wl = x > 0 && row[x - 1];
  const bool n = up[x], nw = x > 0 && up[x - 1], ne = x + 1 < W && up[x + 1];
  const int id = pix_id(x, y, wb);
  if (n) {
    if (!(wl && nw)) uf_union(lab, id, pix_id(x, y - 1, wb));
  } else {
    if (nw && !wl) uf_union(lab, id, pix_id(x - 1, y - 1, wb));
    if (ne) uf_union(lab, id, pix_id(x + 1, y - 1, wb));
  }
}

// flatten + component sizes (one atomic per distinct root in a warp)
__global__ void __launch_bounds__(256) cc_count_kernel(int H, int W, int wb, int* __restrict__ lab, int* __restrict__ size) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  int root = -1, id = 0;
  if (x < W) {
    id = pix_id(x, y, wb);
    if (__ldcg(lab + id) >= 0) {
      root = uf_find(lab, id);
      lab[id] = root;
    }
  }
  const unsigned peers = __match_any_sync(0xffffffffu, root);
  if (root >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(size + root, __popc(peers));
}

// stats: [0] foreground pixels, [1] components, [2] size of the kept component, [3] its root id (-1: none)
__global__ void __launch_bounds__(256) cc_select_kernel(int H, int W, int wb, const int* __restrict__ lab,
                                                        const int* __restrict__ size, unsigned long long* best,
                                                        int* __restrict__ stats) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  unsigned long long key = 0;
  bool fg = false, root = false;
  if (x < W) {
    const int id = pix_id(x, y, wb);
    const int l = lab[id];
    fg = l >= 0;
    root = l == id;
    if (root) key = ((unsigned long long)(unsigned)size[id] << 32) | (0xffffffffu - (unsigned)id);
  }
  const unsigned mf = __ballot_sync(0xffffffffu, fg), mr = __ballot_sync(0xffffffffu, root);
  if (mr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    if (mf) atomicAdd(stats + 0, __popc(mf));
    if (mr) {
      atomicAdd(stats + 1, __popc(mr));
      atomicMax(best, key);
    }
  }
}

// mask = (label == kept root); an EMPTY prediction gives an all-ones mask, as the reference's loop does
// (label_cnt == 1: max_label stays 0 and `labels == 0` is true everywhere, myutils/data.py:27-37)
__global__ void __launch_bounds__(256) cc_write_kernel(int H, int W, int wb, const int* __restrict__ lab,
                                                       const unsigned long long* __restrict__ best,
                                                       uint8_t* __restrict__ mask, int* __restrict__ stats) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const unsigned long long b = *best;
  const int keep = b ? (int)(0xffffffffu - (unsigned)(b & 0xffffffffu)) : -1;
  if (x == 0 && y == 0) {
    stats[2] = (int)(b >> 32);
    stats[3] = keep;
  }
  if (x >= W) return;
  mask[(size_t)y * W + x] = b ? (lab[pix_id(x, y, wb)] == keep) : 1;
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
struct TailWs { int* lab; int* size; unsigned long long* best; };
size_t tail_ws_layout(int H, int W, char* base, TailWs* out) {
  const size_t n = (size_t)4 * ((H + 1) / 2) * ((W + 1) / 2);
  size_t off = 0;
  if (out) out->lab = reinterpret_cast<int*>(base + off);
  off += align_up(n * sizeof(int), 256);
  if (out) out->size = reinterpret_cast<int*>(base + off);
  off += align_up(n * sizeof(int), 256);
  if (out) out->best = reinterpret_cast<unsigned long long*>(base + off);
  off += 256;
  return off;
}
}  // namespace

extern "C" {

size_t vfn_tail_workspace_bytes(int32_t H, int32_t W) {
  if (H <= 0 || W <= 0) return 0;
  return tail_ws_layout(H, W, nullptr, nullptr);
}

int vfn_tail_resize_argmax(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                           int32_t antialias, uint8_t* d_pred, void* stream) {
  VFN_CHECK_ARG(d_pred_mask && d_pred, "tail_resize_argmax: NULL argument");
  VFN_CHECK_ARG(obj_n >= 1 && obj_n <= 255 && h > 0 && w > 0 && H > 0 && W > 0 && H <= 65535,
                "tail_resize_argmax: bad shape obj_n=%d (%d,%d)->(%d,%d)", obj_n, h, w, H, W);
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;   // area_pixel_compute_scale, align_corners=False
  dim3 g((unsigned)cdiv(W, 256), (unsigned)H);
  if (antialias)
    resize_argmax_kernel<true><<<g, 256, 0, as_stream(stream)>>>(d_pred_mask, obj_n, h, w, H, W, sy, sx, d_pred);
  else
    resize_argmax_kernel<false><<<g, 256, 0, as_stream(stream)>>>(d_pred_mask, obj_n, h, w, H, W, sy, sx, d_pred);
  VFN_LAUNCH_OK();
  count_launches(1);
  return VFN_OK;
}

int vfn_tail_largest_component(const uint8_t* d_pred, int32_t H, int32_t W, uint8_t* d_mask, int32_t* d_stats,
                               void* d_ws, size_t ws_bytes, void* stream) {
  VFN_CHECK_ARG(d_pred && d_mask && d_stats && d_ws, "tail_largest_component: NULL argument");
  VFN_CHECK_ARG(H > 0 && W > 0 && H <= 65535 && (int64_t)H * W < (1ll << 29), "tail_largest_component: bad shape (%d,%d)", H, W);
  TailWs ws;
  if (ws_bytes < tail_ws_layout(H, W, static_cast<char*>(d_ws), &ws)) {
    set_error("tail_largest_component: workspace %zu < %zu bytes", ws_bytes, tail_ws_layout(H, W, nullptr, nullptr));
    return VFN_E_CAPACITY;
  }
  cudaStream_t st = as_stream(stream);
  const int wb = (W + 1) / 2;
  dim3 g((unsigned)cdiv(W, 256), (unsigned)H);
  VFN_CUDA_OK(cudaMemsetAsync(ws.best, 0, sizeof(unsigned long long), st));
  VFN_CUDA_OK(cudaMemsetAsync(d_stats, 0, 4 * sizeof(int32_t), st));
  cc_init_kernel<<<g, 256, 0, st>>>(d_pred, H, W, wb, ws.lab, ws.size);
  if (H > 1) {
    dim3 gm((unsigned)cdiv(W, 256), (unsigned)(H - 1));
    cc_merge_kernel<<<gm, 256, 0, st>>>(d_pred, H, W, wb, ws.lab);
  }
  cc_count_kernel<<<g, 256, 0, st>>>(H, W, wb, ws.lab, ws.size);
  cc_select_kernel<<<g, 256, 0, st>>>(H, W, wb, ws.lab, ws.size, ws.best, d_stats);
  cc_write_kernel<<<g, 256, 0, st>>>(H, W, wb, ws.lab, ws.best, d_mask, d_stats);
  VFN_LAUNCH_OK();
  count_launches(H > 1 ? 5 : 4);
  return VFN_OK;
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
  int rc = vfn_tail_resize_argmax(d_pred_mask, obj_n, h, w, H, W, antialias, d_pred, stream);
  if (rc) return rc;
  rc = vfn_tail_largest_component(d_pred, H, W, d_mask, d_stats, d_ws, ws_bytes, stream);
  if (rc) return rc;
  return vfn_tail_waterlevel(d_mask, H, W, d_key_pts, n_pts, water_label_id, d_level, stream);
}

}  // extern "C"
