// block.cu — macro-block-grid interval (the reference's native wire format:
// one H.264 motion vector per 16x16 block, grids [Hg=H/16, Wg=W/16, 2]) and
// the stand-alone bilinear up-sample.
//
// fuvs_block_interval       <- flow/model.py:208-239 + flow/base.py:276-277
// fuvs_upsample_bilinear_ac <- F.interpolate(..., mode="bilinear",
//                              align_corners=True), flow/model.py:193,206,218,228 ...
//
// The warp chains run at grid resolution (C*Hg*Wg*4 B = 161 KB per state at
// C=5, 67x120): step 1 samples the full-resolution key frame, steps >= 2 sample
// the previous low-resolution state (flow/model.py:214-215).  All 2(n-1) states
// stay L2-resident; one streaming kernel then up-samples the two states each
// frame needs (align_corners=True), blends, arg-maxes and writes the labels.
#include "fuvs_common.cuh"

namespace fuvs {

int launch_temporal_counts(const uint8_t* labels, int n, long long HW, const uint8_t* tc_prev, int K,
                           int ignore_index, long long* counts, cudaStream_t st);
int launch_argmax(const float* logits, int frames, int C, long long HW, uint8_t* u8, long long* i64, cudaStream_t st);

// ---------------------------------------------------------------------------
// stand-alone up-sample: one thread per output pixel, planes on grid.y
// ---------------------------------------------------------------------------
template <class NM>
__global__ void __launch_bounds__(256)
upsample_bilinear_ac_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes, int Hin,
                            int Win, int Hout, int Wout, float sh, float sw) {
  const long long opix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long out_plane = static_cast<long long>(Hout) * Wout;
  if (opix >= out_plane) return;
  const int y = static_cast<int>(opix / Wout), x = static_cast<int>(opix - static_cast<long long>(y) * Wout);
  const UpCoord hc = up_coord<NM>(sh, y, Hin), wc = up_coord<NM>(sw, x, Win);
  const long long in_plane = static_cast<long long>(Hin) * Win;
  for (long long pl = blockIdx.y; pl < planes; pl += gridDim.y) {
    dst[pl * out_plane + opix] = up_fetch<NM>(src + pl * in_plane, Win, hc, wc);
  }
}

static inline float ac_scale(int in_size, int out_size) {
  // area_pixel_compute_scale<float>(in, out, align_corners=true): UpSample.cuh
  return out_size > 1 ? static_cast<float>(in_size - 1) / (out_size - 1) : 0.f;
}

template <class NM>
static int launch_upsample(const float* src, float* dst, long long planes, int Hin, int Win, int Hout, int Wout,
                           cudaStream_t st) {
  const long long out_plane = static_cast<long long>(Hout) * Wout;
  const int threads = 256;
  const long long bx = (out_plane + threads - 1) / threads;
  if (bx > 0x7fffffffll) return set_error(FUVS_EINVAL, "upsample: output plane too large");
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(planes < 65535 ? planes : 65535));
  upsample_bilinear_ac_kernel<NM><<<grid, threads, 0, st>>>(src, dst, planes, Hin, Win, Hout, Wout,
                                                             ac_scale(Hin, Hout), ac_scale(Win, Wout));
  return check_launch("fuvs_upsample_bilinear_ac");
}

// ---------------------------------------------------------------------------
// block-grid chain step (both sides, low resolution)
// ---------------------------------------------------------------------------
template <class NM>
__global__ void __launch_bounds__(256)
block_chain_step_kernel(const float* __restrict__ srcL, const float* __restrict__ srcR,
                        const float* __restrict__ gridL, const float* __restrict__ gridR,
                        float* __restrict__ dstL, float* __restrict__ dstR, int C, int Hin, int Win, int Hg, int Wg) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= Wg || y >= Hg) return;
  const bool right = blockIdx.z != 0;
  const float* src = right ? srcR : srcL;
  const float* grid = right ? gridR : gridL;
  float* dst = right ? dstR : dstL;
  const int opix = y * Wg + x;
  const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + opix);
  const GsTap t = gs_setup<NM>(g.x, g.y, Hin, Win, false);
  const long long in_plane = static_cast<long long>(Hin) * Win;
  const int out_plane = Hg * Wg;
#pragma unroll 5
  for (int c = 0; c < C; ++c) dst[c * out_plane + opix] = gs_fetch<NM>(src + c * in_plane, t, Win);
}

// ---------------------------------------------------------------------------
// streaming kernel: frame p = w0 * up(L_p) + w1 * up(R_{n-p}), p = 1..n-1, and
// frame 0 = key frame.  One thread = one output pixel; a warp covers 32
// consecutive pixels of a row, which fall into <= 3 source columns, so the tap
// loads are L1 broadcast hits.
// ---------------------------------------------------------------------------
template <class NM, int CT>
__global__ void __launch_bounds__(256)
block_stream_kernel(const float* __restrict__ key0, const float* __restrict__ Lst, const float* __restrict__ Rst,
                    int Crt, int H, int W, int Hg, int Wg, int n, float sh, float sw,
                    uint8_t* __restrict__ labels, float* __restrict__ logits, const BlendWeights wts) {
  const int C = CT > 0 ? CT : Crt;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  const long long HW = static_cast<long long>(H) * W;
  const long long pix = static_cast<long long>(y) * W + x;
  const int lp = Hg * Wg;                 // low-res plane
  const long long ls = static_cast<long long>(C) * lp;   // low-res state
  // frame 0
  {
    ArgMax am;
    am.init(0.f);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = __ldcs(key0 + c * HW + pix);
      if (c == 0) am.init(v); else am.push(v, c);
      if (logits) __stcs(logits + c * HW + pix, v);
    }
    if (labels) labels[pix] = static_cast<uint8_t>(am.idx);
  }
  const bool same = (Hg == H) && (Wg == W);   // ATen copies when sizes match; the reference skips the call
  const UpCoord hc = up_coord<NM>(sh, y, Hg), wc = up_coord<NM>(sw, x, Wg);
  const int off = hc.i0 * Wg + wc.i0;
  const int o01 = wc.ip, o10 = hc.ip * Wg, o11 = hc.ip * Wg + wc.ip;
  for (int p = 1; p < n; ++p) {
    const float w0 = wts.w0[p], w1 = wts.w1[p];
    const float* Lp = Lst + (p - 1) * ls + off;
    const float* Rp = Rst + (n - p - 1) * ls + off;
    ArgMax am;
    am.init(0.f);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float* l = Lp + c * lp;
      const float* r = Rp + c * lp;
      float f, b;
      if (same) {
        f = __ldg(l);
        b = __ldg(r);
      } else {
        f = up_value<NM>(hc, wc, __ldg(l), __ldg(l + o01), __ldg(l + o10), __ldg(l + o11));
        b = up_value<NM>(hc, wc, __ldg(r), __ldg(r + o01), __ldg(r + o10), __ldg(r + o11));
      }
      const float v = blend2(w0, f, w1, b);
      if (c == 0) am.init(v); else am.push(v, c);
      if (logits) __stcs(logits + (static_cast<long long>(p) * C + c) * HW + pix, v);
    }
    if (labels) labels[p * HW + pix] = static_cast<uint8_t>(am.idx);
  }
}

}  // namespace fuvs

extern "C" int fuvs_upsample_bilinear_ac(const float* src, float* dst, long long planes, int Hin, int Win, int Hout,
                                         int Wout, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!src || !dst || planes < 0 || Hin < 1 || Win < 1 || Hout < 0 || Wout < 0)
    return set_error(FUVS_EINVAL, "upsample: bad arguments planes=%lld in=%dx%d out=%dx%d", planes, Hin, Win, Hout, Wout);
  if (static_cast<long long>(Hin) * Win >= (1ll << 31)) return set_error(FUVS_EINVAL, "upsample: source plane exceeds 2^31 elements");
  if (planes == 0 || Hout == 0 || Wout == 0) return FUVS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Hin == Hout && Win == Wout) {   // ATen's "special case: just copy"
    if (cudaMemcpyAsync(dst, src, planes * Hin * Win * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return set_error(FUVS_ECUDA, "upsample: copy failed");
    return FUVS_OK;
  }
  return launch_upsample<Nm>(src, dst, planes, Hin, Win, Hout, Wout, st);
}

extern "C" long long fuvs_block_scratch_floats(int C, int Hg, int Wg, int n) {
  if (n <= 1) return 0;
  return 2ll * (n - 1) * C * static_cast<long long>(Hg) * Wg;
}

extern "C" int fuvs_block_interval(const float* prev, const float* next, const float* grids_left,
                                   const float* grids_right, int C, int H, int W, int Hg, int Wg, int n,
                                   float* scratch, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                                   long long* counts, int ignore_index, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!prev || C < 1 || H < 1 || W < 1 || n < 1) return set_error(FUVS_EINVAL, "block: bad shape C=%d H=%d W=%d n=%d", C, H, W, n);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "block: n=%d exceeds %d frames per interval", n, FUVS_MAX_FRAMES);
  if (n > 1 && (!next || !grids_left || !grids_right || !scratch || Hg < 1 || Wg < 1))
    return set_error(FUVS_EINVAL, "block: next/grids/scratch are NULL or grid is empty (Hg=%d Wg=%d) but n=%d", Hg, Wg, n);
  if ((labels || counts) && C > 256) return set_error(FUVS_EINVAL, "block: uint8 label maps need C <= 256 (C=%d)", C);
  if (counts && !labels) return set_error(FUVS_EINVAL, "block: counts need the label maps (labels is NULL)");
  const long long HW = static_cast<long long>(H) * W;
  if (HW >= (1ll << 31) || static_cast<long long>(C) * Hg * Wg * 2 * n >= (1ll << 31))
    return set_error(FUVS_EINVAL, "block: problem exceeds 32-bit indexing");
  if (n > 1 && (!aligned8(grids_left) || !aligned8(grids_right))) return set_error(FUVS_EALIGN, "block: grids must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long S = static_cast<long long>(C) * HW;

  if (n == 1) {
    if (labels) {
      if (int e = launch_argmax(prev, 1, C, HW, labels, nullptr, st)) return e;
    }
    if (logits) {
      if (cudaMemcpyAsync(logits, prev, S * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return set_error(FUVS_ECUDA, "block: logits copy failed");
    }
  } else {
    const long long ls = static_cast<long long>(C) * Hg * Wg;
    float* Lst = scratch;                    // L_1 .. L_{n-1}
    float* Rst = scratch + (n - 1) * ls;     // R_1 .. R_{n-1}
    dim3 cgrid((Wg + 31) / 32, (Hg + 7) / 8, 2), cblock(32, 8);
    for (int j = 1; j <= n - 1; ++j) {
      const float* sL = (j == 1) ? prev : Lst + (j - 2) * ls;
      const float* sR = (j == 1) ? next : Rst + (j - 2) * ls;
      const int Hin = (j == 1) ? H : Hg, Win = (j == 1) ? W : Wg;
      block_chain_step_kernel<Nm><<<cgrid, cblock, 0, st>>>(
          sL, sR, grids_left + static_cast<long long>(j - 1) * Hg * Wg * 2,
          grids_right + static_cast<long long>(j - 1) * Hg * Wg * 2, Lst + (j - 1) * ls, Rst + (j - 1) * ls, C, Hin,
          Win, Hg, Wg);
      if (int e = check_launch("fuvs_block_interval(chain)")) return e;
    }
    if (labels || logits) {
      BlendWeights w;
      make_blend_weights(n, &w);
      dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
      if (grid.y > 65535) return set_error(FUVS_EINVAL, "block: H=%d too large", H);
      const float sh = ac_scale(Hg, H), sw = ac_scale(Wg, W);
      switch (C) {
        case 2: block_stream_kernel<Nm, 2><<<grid, block, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, w); break;
        case 5: block_stream_kernel<Nm, 5><<<grid, block, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, w); break;
        default: block_stream_kernel<Nm, 0><<<grid, block, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, w); break;
      }
      if (int e = check_launch("fuvs_block_interval(stream)")) return e;
    }
  }
  if (counts) return launch_temporal_counts(labels, n, HW, tc_prev, C, ignore_index, counts, st);
  return FUVS_OK;
}
