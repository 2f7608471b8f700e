// crop.cu — sliding-crop inference on the device (model.no_cropping=False).
//
// fuvs_crop_grid        <- crop_motion_vector, flow/transform.py:215-261 (called per crop from compute_output,
//                          flow/base.py:204).  The reference moves every grid to the host, slices and re-normalises
//                          it in numpy and resizes it with cv2.resize(INTER_LINEAR), then uploads it again: 2(k-1)
//                          device<->host round trips per crop.  Here the grid never leaves HBM.
// fuvs_crop_accumulate  <- F.softmax(output, dim=1) (flow/base.py:220,233) and the fp64 canvas update
//                          `prediction_crop[:, :, s_h:e_h, s_w:e_w] += ...; count_crop[...] += 1` (flow/base.py:206-207)
// fuvs_crop_finish      <- `prediction_crop /= count_crop` (flow/base.py:208) and output.data.max(1)[1]
//                          (flow/base.py:167,276) on the fp64 canvas
//
// Numerics.  numpy evaluates the re-normalisation op by op in float32 (python scalars are weak); cv2's 32F
// INTER_LINEAR is fl(fl(S0*a0) + fl(S1*a1)) horizontally, then the same vertically, with coordinates
// fx = float((dx+0.5)*scale - 0.5) computed in double (checked bit for bit against cv2 4.13 in
// tests/test_crop_oracle.py).  ATen's spatial soft-max is max -> sum of expf(x - max) in class order -> expf(x - max)
// / sum with an IEEE division; the canvas arithmetic is IEEE double.
#include <cmath>

#include "fuvs_common.cuh"

namespace fuvs {

namespace {

struct CropGeom {
  int Hg, Wg;            // source grid
  int bh_off, bw_off;    // first block row / column of the crop
  int bh, bw;            // blocks covered by the crop
  int oh, ow;            // output grid (crop_h / 16, crop_w / 16)
  float width, height;   // image size as float32 (numpy casts the python int)
  float w_off, h_off;
  float den_x, den_y;    // float32(block_width * pixel_per_block_width), float32(block_height * pixel_per_block_height)
  double scale_x, scale_y;
};

// numpy: ((((m + 1) / 2) * size - offset) / den) * 2 - 1, every operation rounded to float32
__device__ __forceinline__ float renorm(float m, float size, float off, float den) {
  float t = __fadd_rn(m, 1.f);
  t = __fdiv_rn(t, 2.f);
  t = __fmul_rn(t, size);
  t = __fsub_rn(t, off);
  t = __fdiv_rn(t, den);
  t = __fmul_rn(t, 2.f);
  return __fsub_rn(t, 1.f);
}

// cv2 resize(INTER_LINEAR) source coordinate of destination index d: floor index and fraction
__device__ __forceinline__ void cv_coord(int d, double scale, int ssize, int* s0, float* f) {
  float fx = static_cast<float>((d + 0.5) * scale - 0.5);
  int sx = static_cast<int>(floorf(fx));
  fx = __fsub_rn(fx, static_cast<float>(sx));
  if (sx < 0) { fx = 0.f; sx = 0; }
  if (sx >= ssize - 1) { fx = 0.f; sx = ssize - 1; }
  *s0 = sx;
  *f = fx;
}

__global__ void __launch_bounds__(256)
crop_grid_kernel(const float* __restrict__ grid, float* __restrict__ out, CropGeom G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.oh * G.ow) return;
  const int oy = i / G.ow, ox = i - oy * G.ow;
  int sx, sy;
  float fx, fy;
  cv_coord(ox, G.scale_x, G.bw, &sx, &fx);
  cv_coord(oy, G.scale_y, G.bh, &sy, &fy);
  const int sx1 = min(sx + 1, G.bw - 1), sy1 = min(sy + 1, G.bh - 1);
  const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
  auto src = [&](int y, int x, int ch) {
    const float m = __ldg(grid + (static_cast<long long>(G.bh_off + y) * G.Wg + (G.bw_off + x)) * 2 + ch);
    return ch == 0 ? renorm(m, G.width, G.w_off, G.den_x) : renorm(m, G.height, G.h_off, G.den_y);
  };
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const float r0 = __fadd_rn(__fmul_rn(src(sy, sx, ch), a0), __fmul_rn(src(sy, sx1, ch), a1));
    const float r1 = __fadd_rn(__fmul_rn(src(sy1, sx, ch), a0), __fmul_rn(src(sy1, sx1, ch), a1));
    out[static_cast<long long>(i) * 2 + ch] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
  }
}

// canvas[f, c, h_off + y, w_off + x] += softmax_c(logits[f, :, y, x]);  count[h_off + y, w_off + x] += 1
__global__ void __launch_bounds__(256)
crop_accumulate_kernel(const float* __restrict__ logits, double* __restrict__ canvas, double* __restrict__ count,
                       int n, int C, int ch, int cw, int H, int W, int h_off, int w_off) {
  const long long plane = static_cast<long long>(ch) * cw;
  const long long total = plane * n;
  const long long HW = static_cast<long long>(H) * W;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = t / plane;
    const long long p = t - f * plane;
    const int y = static_cast<int>(p / cw), x = static_cast<int>(p - static_cast<long long>(y) * cw);
    const float* in = logits + f * C * plane + p;
    // ATen cunn_SpatialSoftMaxForward: max, then sum of exp(x - max) in class order, then exp(x - max) / sum
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(in + c * plane));
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(in + c * plane), mx)));
    const long long cpix = static_cast<long long>(h_off + y) * W + (w_off + x);
    double* cv = canvas + f * C * HW + cpix;
    for (int c = 0; c < C; ++c) {
      const float pr = __fdiv_rn(expf(__fsub_rn(__ldg(in + c * plane), mx)), sum);
      cv[c * HW] = __dadd_rn(cv[c * HW], static_cast<double>(pr));
    }
    if (f == 0) count[cpix] = __dadd_rn(count[cpix], 1.0);
  }
}

// canvas /= count (in place, like the reference) and labels = argmax_c with torch.max semantics
__global__ void __launch_bounds__(256)
crop_finish_kernel(double* __restrict__ canvas, const double* __restrict__ count, int n, int C, long long HW,
                   uint8_t* __restrict__ labels) {
  const long long total = HW * n;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = t / HW;
    const long long p = t - f * HW;
    const double cnt = count[p];
    double* cv = canvas + f * C * HW + p;
    double best = 0.0;
    int idx = 0;
    for (int c = 0; c < C; ++c) {
      const double v = __ddiv_rn(cv[c * HW], cnt);
      cv[c * HW] = v;
      const bool take = (c == 0) || (v > best) || ((v != v) && (best == best));
      if (take) { best = v; idx = c; }
    }
    if (labels) labels[t] = static_cast<uint8_t>(idx);
  }
}

int grid_for(long long items, int threads) {
  long long g = (items + threads - 1) / threads;
  const long long cap = 16ll * sm_count();
  if (g > cap) g = cap;
  return static_cast<int>(g > 0 ? g : 1);
}

}  // namespace

}  // namespace fuvs

extern "C" int fuvs_crop_grid_shape(int crop_h, int crop_w, int* out_h, int* out_w) {
  if (!out_h || !out_w || crop_h < 16 || crop_w < 16) return fuvs::set_error(FUVS_EINVAL, "crop_grid_shape: crop %dx%d", crop_h, crop_w);
  *out_h = crop_h / 16;   // final_block_height = crop_height // 16 (flow/transform.py:229)
  *out_w = crop_w / 16;
  return FUVS_OK;
}

extern "C" int fuvs_crop_grid(const float* grid, int Hg, int Wg, int H, int W, int crop_h, int crop_w, int h_off,
                              int w_off, float* out, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!grid || !out || Hg < 1 || Wg < 1 || H < 1 || W < 1 || crop_h < 16 || crop_w < 16 || h_off < 0 || w_off < 0)
    return set_error(FUVS_EINVAL, "crop_grid: bad arguments grid=%dx%d image=%dx%d crop=%dx%d@(%d,%d)", Hg, Wg, H, W,
                     crop_h, crop_w, h_off, w_off);
  // flow/transform.py:223-235 in python-double arithmetic; python's round() is round-half-to-even = nearbyint()
  const double ppb_h = static_cast<double>(H) / Hg, ppb_w = static_cast<double>(W) / Wg;
  CropGeom g;
  g.Hg = Hg; g.Wg = Wg;
  g.oh = crop_h / 16; g.ow = crop_w / 16;
  g.bh_off = static_cast<int>(std::nearbyint(h_off / ppb_h));
  g.bw_off = static_cast<int>(std::nearbyint(w_off / ppb_w));
  g.bh = static_cast<int>(std::nearbyint((h_off + crop_h) / ppb_h)) - g.bh_off;
  g.bw = static_cast<int>(std::nearbyint((w_off + crop_w) / ppb_w)) - g.bw_off;
  // the re-normalisation divides by the UNCLIPPED block counts (flow/transform.py:238-239) ...
  const int bh_full = g.bh, bw_full = g.bw;
  // ... while numpy slicing clips at the array end (:237) and cv2.resize scales from the clipped shape; an empty slice
  // makes cv2.resize raise in the reference
  if (g.bh_off + g.bh > Hg) g.bh = Hg - g.bh_off;
  if (g.bw_off + g.bw > Wg) g.bw = Wg - g.bw_off;
  if (g.bh < 1 || g.bw < 1)
    return set_error(FUVS_EINVAL, "crop_grid: the crop covers no grid block (cv2.resize would raise in the reference)");
  g.width = static_cast<float>(W); g.height = static_cast<float>(H);
  g.w_off = static_cast<float>(w_off); g.h_off = static_cast<float>(h_off);
  g.den_x = static_cast<float>(bw_full * ppb_w);
  g.den_y = static_cast<float>(bh_full * ppb_h);
  g.scale_x = static_cast<double>(g.bw) / g.ow;   // cv2: scale = 1 / (dsize / ssize) in double
  g.scale_y = static_cast<double>(g.bh) / g.oh;
  {
    // cv2 computes inv_scale = dsize/ssize, then scale = 1./inv_scale
    const double inv_x = static_cast<double>(g.ow) / g.bw, inv_y = static_cast<double>(g.oh) / g.bh;
    g.scale_x = 1.0 / inv_x;
    g.scale_y = 1.0 / inv_y;
  }
  const int items = g.oh * g.ow;
  crop_grid_kernel<<<(items + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(grid, out, g);
  return check_launch("fuvs_crop_grid");
}

extern "C" int fuvs_crop_accumulate(const float* logits, double* canvas, double* count, int n, int C, int crop_h,
                                    int crop_w, int H, int W, int h_off, int w_off, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!logits || !canvas || !count || n < 1 || C < 1 || crop_h < 1 || crop_w < 1 || h_off < 0 || w_off < 0 ||
      h_off + crop_h > H || w_off + crop_w > W)
    return set_error(FUVS_EINVAL, "crop_accumulate: bad arguments n=%d C=%d crop=%dx%d@(%d,%d) canvas=%dx%d", n, C,
                     crop_h, crop_w, h_off, w_off, H, W);
  if (!aligned8(canvas) || !aligned8(count)) return set_error(FUVS_EALIGN, "crop_accumulate: fp64 buffers must be 8-byte aligned");
  const long long items = static_cast<long long>(n) * crop_h * crop_w;
  crop_accumulate_kernel<<<grid_for(items, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, canvas, count, n, C, crop_h, crop_w, H, W, h_off, w_off);
  return check_launch("fuvs_crop_accumulate");
}

extern "C" int fuvs_crop_finish(double* canvas, const double* count, int n, int C, long long HW, uint8_t* labels,
                                fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!canvas || !count || n < 1 || C < 1 || HW < 1) return set_error(FUVS_EINVAL, "crop_finish: bad arguments n=%d C=%d HW=%lld", n, C, HW);
  if (labels && C > 256) return set_error(FUVS_EINVAL, "crop_finish: uint8 label maps need C <= 256 (C=%d)", C);
  crop_finish_kernel<<<grid_for(HW * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(canvas, count, n, C, HW, labels);
  return check_launch("fuvs_crop_finish");
}
