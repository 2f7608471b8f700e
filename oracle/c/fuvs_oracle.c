/*
 * fuvs_oracle.c — TEST INFRASTRUCTURE (see oracle/__init__.py): a plain-C, single-thread-semantics restatement
 * of the arithmetic the reference executes on this path.
 *
 * The reference (Python) calls torch; the numbers are produced by ATen's CUDA kernels, which are not under
 * /root/reference.  What is restated here is their published algorithm (torch 2.11 headers in this image:
 * ATen/native/cuda/GridSampler.cuh:23-31,55-57 and ATen/native/cuda/UpSample.cuh:96-130; upstream
 * GridSampler.cu / UpSampleBilinear2d.cu / SummaryOps.cu for the kernel bodies), anchored on the reference's call
 * sites:
 *   fo_grid_sample   F.grid_sample(bilinear, border)            flow/model.py:157 (align_corners=True), :248 (False)
 *   fo_upsample_ac   F.interpolate(bilinear, align_corners=True) flow/model.py:193,206,218,228; flow/base.py:275
 *   fo_blend         (n-p)/n * x + p/n * y                       flow/model.py:104,234-236
 *   fo_argmax        output.data.max(1)[1]                       flow/base.py:147,167,276
 *   fo_counts        intersectionAndUnion[GPU]                   util/util.py:36-47, 52-63
 *   fo_interval      FlowModel.predict_segmentation              flow/model.py:184-241
 *   fo_temporal      temporal-consistency loop                   flow/base.py:280-295
 * nvcc contracts a*b+c into FMA where ATen's source allows it; those FMAs are written as fmaf() and the file is
 * compiled with -ffp-contract=off so that nothing else fuses.  Which expressions fuse was read from nvcc's SASS and
 * proven against torch-CUDA on a B200 (tests/test_calibration_gpu.py, profiles/r01_calibration_*.json).
 * The omp pragmas are inert unless built with -fopenmp (not used: this image has no libgomp); every output value is
 * computed by the same sequence of fp32 operations either way.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ---- grid_sampler_2d forward, bilinear, padding_mode=border ------------------------------------------------ */
static float src_index(float coord, int size, int align_corners) {
  float r;
  if (align_corners) r = ((coord + 1.f) / 2) * (float)(size - 1);          /* GridSampler.cuh:25-27 */
  else r = fmaf(coord + 1.f, (float)size, -1.f) / 2;                       /* :29-30, (x+1)*size-1 fused by nvcc */
  r = fmaxf(r, 0.f);                                                        /* clip_coordinates :55-57 */
  return fminf((float)(size - 1), r);
}

API void fo_grid_sample(const float* src, const float* grid, float* dst, int C, int Hin, int Win, int Hg, int Wg,
                        int align_corners) {
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < Hg * Wg; ++idx) {
    const float ix = src_index(grid[2 * idx], Win, align_corners);
    const float iy = src_index(grid[2 * idx + 1], Hin, align_corners);
    const int ix_nw = (int)floorf(ix), iy_nw = (int)floorf(iy);
    const int ix_ne = ix_nw + 1, iy_ne = iy_nw, ix_sw = ix_nw, iy_sw = iy_nw + 1, ix_se = ix_nw + 1, iy_se = iy_nw + 1;
    const float nw = ((float)ix_se - ix) * ((float)iy_se - iy);
    const float ne = (ix - (float)ix_sw) * ((float)iy_sw - iy);
    const float sw = ((float)ix_ne - ix) * (iy - (float)iy_ne);
    const float se = (ix - (float)ix_nw) * (iy - (float)iy_nw);
    for (int c = 0; c < C; ++c) {
      const float* p = src + (size_t)c * Hin * Win;
      float acc = 0.f;                                                      /* out_acc += v * w, each an FMA */
      if (iy_nw >= 0 && iy_nw < Hin && ix_nw >= 0 && ix_nw < Win) acc = fmaf(p[iy_nw * Win + ix_nw], nw, acc);
      if (iy_ne >= 0 && iy_ne < Hin && ix_ne >= 0 && ix_ne < Win) acc = fmaf(p[iy_ne * Win + ix_ne], ne, acc);
      if (iy_sw >= 0 && iy_sw < Hin && ix_sw >= 0 && ix_sw < Win) acc = fmaf(p[iy_sw * Win + ix_sw], sw, acc);
      if (iy_se >= 0 && iy_se < Hin && ix_se >= 0 && ix_se < Win) acc = fmaf(p[iy_se * Win + ix_se], se, acc);
      dst[(size_t)c * Hg * Wg + idx] = acc;
    }
  }
}

/* ---- upsample_bilinear2d forward, align_corners=True ------------------------------------------------------- */
API void fo_upsample_ac(const float* src, float* dst, long long planes, int Hin, int Win, int Hout, int Wout) {
  if (Hin == Hout && Win == Wout) {                                         /* "special case: just copy" */
    memcpy(dst, src, (size_t)planes * Hin * Win * sizeof(float));
    return;
  }
  const float rh = Hout > 1 ? (float)(Hin - 1) / (Hout - 1) : 0.f;          /* area_pixel_compute_scale */
  const float rw = Wout > 1 ? (float)(Win - 1) / (Wout - 1) : 0.f;
#pragma omp parallel for schedule(static)
  for (int h2 = 0; h2 < Hout; ++h2) {
    const float h1r = rh * h2;
    const int h1 = (int)h1r, h1p = (h1 < Hin - 1) ? 1 : 0;
    const float h1l = h1r - h1, h0l = 1.f - h1l;
    for (int w2 = 0; w2 < Wout; ++w2) {
      const float w1r = rw * w2;
      const int w1 = (int)w1r, w1p = (w1 < Win - 1) ? 1 : 0;
      const float w1l = w1r - w1, w0l = 1.f - w1l;
      for (long long pl = 0; pl < planes; ++pl) {
        const float* p = src + (size_t)pl * Hin * Win;
        /* h0l*(w0l*a + w1l*b) + h1l*(w0l*c + w1l*d): nvcc emits fma(first product, mul(second product)) */
        const float r0 = fmaf(w0l, p[h1 * Win + w1], w1l * p[h1 * Win + w1 + w1p]);
        const float r1 = fmaf(w0l, p[(h1 + h1p) * Win + w1], w1l * p[(h1 + h1p) * Win + w1 + w1p]);
        dst[(size_t)pl * Hout * Wout + (size_t)h2 * Wout + w2] = fmaf(h0l, r0, h1l * r1);
      }
    }
  }
}

/* ---- wa*a + wb*b: three separately rounded elementwise launches ------------------------------------------------ */
API void fo_blend(const float* a, const float* b, double wa, double wb, float* out, long long n) {
  const float fa = (float)wa, fb = (float)wb;                               /* mul(Tensor, Scalar): scalar -> float */
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    const float x = a[i] * fa;
    out[i] = b ? x + b[i] * fb : x;
  }
}

/* ---- max(dim=1) indices: lowest index on ties, NaN beats numbers, first NaN stays ---------------------------------- */
API void fo_argmax(const float* logits, int frames, int C, long long HW, uint8_t* labels) {
#pragma omp parallel for schedule(static)
  for (long long t = 0; t < (long long)frames * HW; ++t) {
    const long long f = t / HW, i = t - f * HW;
    const float* p = logits + (size_t)f * C * HW + i;
    float best = p[0];
    int idx = 0;
    for (int c = 1; c < C; ++c) {
      const float v = p[(size_t)c * HW];
      if (v > best || (isnan(v) && !isnan(best))) { best = v; idx = c; }
    }
    labels[t] = (uint8_t)idx;
  }
}

/* ---- (I,U,T) counts ------------------------------------------------------------------------------------------------ */
static int bin_of(long long v, int K, int np_bins) {
  if (np_bins) {                       /* np.histogram(bins=arange(K+1)): last bin closed */
    if (v < 0 || v > K) return -1;
    return v == K ? K - 1 : (int)v;
  }
  if (K == 1) return 0;                /* torch.histc with min == max == 0 uses the data range */
  if (v < 0 || v > K - 1) return -1;   /* torch.histc(bins=K, min=0, max=K-1) drops outliers */
  return (int)v;
}

API void fo_counts(const long long* pred, const long long* target, long long N, int K, long long ignore, int np_bins,
                   long long* counts /* [3,K] accumulate */) {
  long long* I = counts; long long* U = counts + K; long long* T = counts + 2 * K;
  for (long long i = 0; i < N; ++i) {
    long long o = pred[i];
    const long long t = target[i];
    if (t == ignore) o = ignore;                                            /* util/util.py:41 / :57 */
    const int bo = bin_of(o, K, np_bins), bt = bin_of(t, K, np_bins);
    if (bo >= 0) U[bo] += 1;                                                /* area_output */
    if (bt >= 0) { T[bt] += 1; U[bt] += 1; }                                /* + area_target */
    if (bo >= 0 && o == t) { I[bo] += 1; U[bo] -= 1; }                      /* - area_intersection */
  }
}

API void fo_temporal(const uint8_t* labels, int n, long long HW, const uint8_t* tc_prev, int K, long long ignore,
                     long long* counts) {
  long long* a = (long long*)malloc(sizeof(long long) * HW);
  long long* b = (long long*)malloc(sizeof(long long) * HW);
  for (int p = 0; p < n; ++p) {
    const uint8_t* ref = p > 0 ? labels + (size_t)(p - 1) * HW : tc_prev;   /* flow/base.py:282-290 */
    if (!ref) continue;
    for (long long i = 0; i < HW; ++i) { a[i] = labels[(size_t)p * HW + i]; b[i] = ref[i]; }
    fo_counts(a, b, HW, K, ignore, 0, counts);
  }
  free(a); free(b);
}

/* ---- one interval of FlowModel.predict_segmentation with identity encoder/decoder ---------------------------------- */
/* mode 0: no_warp (linear); 1: warp with grids [n-1,Hg,Wg,2].  logits [n,C,H,W], labels [n,H,W] (either may be NULL) */
API void fo_interval(const float* prev, const float* next, const float* gl, const float* gr, int C, int H, int W,
                     int Hg, int Wg, int n, int warp, float* logits_out, uint8_t* labels) {
  const size_t S = (size_t)C * H * W, ls = (size_t)C * Hg * Wg;
  float* logits = logits_out ? logits_out : (float*)malloc(sizeof(float) * S * n);
  memcpy(logits, prev, sizeof(float) * S);                                  /* frame 0: flow/model.py:195-197 */
  if (n > 1) {
    float** F = (float**)calloc(n, sizeof(float*));
    float** B = (float**)calloc(n, sizeof(float*));
    if (warp) {
      const int same = (Hg == H && Wg == W);
      float* curL = (float*)malloc(sizeof(float) * ls);
      float* curR = (float*)malloc(sizeof(float) * ls);
      float* nxt = (float*)malloc(sizeof(float) * ls);
      for (int j = 1; j < n; ++j) {                                         /* flow/model.py:212-229 */
        fo_grid_sample(j == 1 ? prev : curL, gl + (size_t)(j - 1) * Hg * Wg * 2, nxt, C, j == 1 ? H : Hg,
                       j == 1 ? W : Wg, Hg, Wg, 0);
        memcpy(curL, nxt, sizeof(float) * ls);
        fo_grid_sample(j == 1 ? next : curR, gr + (size_t)(j - 1) * Hg * Wg * 2, nxt, C, j == 1 ? H : Hg,
                       j == 1 ? W : Wg, Hg, Wg, 0);
        memcpy(curR, nxt, sizeof(float) * ls);
        F[j] = (float*)malloc(sizeof(float) * S);
        B[j] = (float*)malloc(sizeof(float) * S);
        if (same) { memcpy(F[j], curL, sizeof(float) * S); memcpy(B[j], curR, sizeof(float) * S); }
        else { fo_upsample_ac(curL, F[j], C, Hg, Wg, H, W); fo_upsample_ac(curR, B[j], C, Hg, Wg, H, W); }
      }
      free(curL); free(curR); free(nxt);
    }
    for (int p = 1; p < n; ++p)                                             /* flow/model.py:233-237 */
      fo_blend(warp ? F[p] : prev, warp ? B[n - p] : next, (double)(n - p) / n, (double)p / n, logits + (size_t)p * S,
               (long long)S);
    for (int j = 0; j < n; ++j) { free(F[j]); free(B[j]); }
    free(F); free(B);
  }
  if (labels) fo_argmax(logits, n, C, (long long)H * W, labels);
  if (!logits_out) free(logits);
}
