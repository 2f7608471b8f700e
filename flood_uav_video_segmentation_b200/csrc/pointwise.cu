// pointwise.cu — per-pixel kernels shared by the generic (non-fused) routes.
//
// fuvs_blend_argmax <- flow/model.py:104 (x * ((n-index)/n)), :64/:84 (sum of
//                      the two sides), :168,:170,:234-236 (one blended frame)
// fuvs_argmax       <- output.data.max(1)[1], flow/base.py:147,167,276
// fuvs_upsample_argmax <- F.interpolate(output, (1072, 1920), bilinear, align_corners=True) + max(1)[1] + uint8,
//                      flow/base.py:275-277 (predict_step's final resize)
#include "fuvs_common.cuh"

namespace fuvs {

// out[f,c,i] = fl(fl(wa*a) + fl(wb*b)); labels[f,i] = argmax_c out[f,c,i]
template <int VEC>
__global__ void __launch_bounds__(256)
blend_argmax_kernel(const float* __restrict__ a, const float* __restrict__ b, float wa, float wb, int frames, int C,
                    long long HW, float* __restrict__ out, uint8_t* __restrict__ labels) {
  const long long nvec = HW / VEC;
  const long long total = nvec * frames;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long f = t / nvec;
    const long long pix = (t - f * nvec) * VEC;
    const long long base = f * C * HW + pix;
    ArgMax am[VEC];
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      FVec<VEC> va, o;
      va.load_stream(a + base + c * HW);
      if (b) {
        FVec<VEC> vb;
        vb.load_stream(b + base + c * HW);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = blend2(wa, va.v[i], wb, vb.v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = __fmul_rn(va.v[i], wa);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (c == 0) am[i].init(o.v[i]); else am[i].push(o.v[i], c);
      }
      if (out) o.store_stream(out + base + c * HW);
    }
    if (labels) {
      LVec<VEC> lab;
#pragma unroll
      for (int i = 0; i < VEC; ++i) lab.v[i] = am[i].idx;
      lab.store(labels + f * HW + pix);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
argmax_kernel(const float* __restrict__ logits, int frames, int C, long long HW, uint8_t* __restrict__ u8,
              long long* __restrict__ i64) {
  const long long nvec = HW / VEC;
  const long long total = nvec * frames;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long f = t / nvec;
    const long long pix = (t - f * nvec) * VEC;
    const float* p = logits + f * C * HW + pix;
    ArgMax am[VEC];
#pragma unroll 5
    for (int c = 0; c < C; ++c) {
      FVec<VEC> v;
      v.load_stream(p + c * HW);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (c == 0) am[i].init(v.v[i]); else am[i].push(v.v[i], c);
      }
    }
    if (u8) {
      LVec<VEC> lab;
#pragma unroll
      for (int i = 0; i < VEC; ++i) lab.v[i] = am[i].idx;
      lab.store(u8 + f * HW + pix);
    }
    if (i64) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) i64[f * HW + pix + i] = am[i].idx;
    }
  }
}

// labels[f, y, x] = argmax_c up(logits[f, c])(y, x): the resized logits (n*C*Hout*Wout floats, written and read back by
// the reference's two launches) never exist.  One thread = 4 consecutive output pixels of one row; the taps of
// neighbouring threads overlap and come from L1.  Same arithmetic as fuvs_upsample_bilinear_ac (up_coord / up_value).
template <int CT>
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ src, int frames, int Crt, int Hin, int Win, int Hout, int Wout, float sh,
                       float sw, uint8_t* __restrict__ labels, float* __restrict__ out) {
  const int C = CT > 0 ? CT : Crt;
  const int wq = (Wout + 3) >> 2;
  const long long per_frame = static_cast<long long>(Hout) * wq;
  const long long total = per_frame * frames;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long in_plane = static_cast<long long>(Hin) * Win, out_plane = static_cast<long long>(Hout) * Wout;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int f = static_cast<int>(t / per_frame);
    const long long rem = t - f * per_frame;
    const int y = static_cast<int>(rem / wq), x0 = static_cast<int>(rem - static_cast<long long>(y) * wq) * 4;
    const UpCoord hc = up_coord<Nm>(sh, y, Hin);
    UpCoord wc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wc[i] = up_coord<Nm>(sw, min(x0 + i, Wout - 1), Win);
    ArgMax am[4];
    const float* base = src + static_cast<long long>(f) * C * in_plane;
#pragma unroll 5
    for (int c = 0; c < C; ++c) {
      const float* pl = base + c * in_plane;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = up_fetch<Nm>(pl, Win, hc, wc[i]);
        if (c == 0) am[i].init(v); else am[i].push(v, c);
        if (out && x0 + i < Wout) __stcs(out + (static_cast<long long>(f) * C + c) * out_plane + static_cast<long long>(y) * Wout + x0 + i, v);
      }
    }
    if (labels) {
      uint8_t* l = labels + f * out_plane + static_cast<long long>(y) * Wout + x0;
      if (x0 + 3 < Wout && (reinterpret_cast<uintptr_t>(l) & 3u) == 0) {
        *reinterpret_cast<unsigned*>(l) = static_cast<unsigned>(am[0].idx) | (static_cast<unsigned>(am[1].idx) << 8) |
                                          (static_cast<unsigned>(am[2].idx) << 16) | (static_cast<unsigned>(am[3].idx) << 24);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x0 + i < Wout) l[i] = static_cast<uint8_t>(am[i].idx);
      }
    }
  }
}

template <typename K>
static int persistent_grid(K kernel, long long work_items, int threads) {
  const long long need = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * blocks_per_sm(kernel, threads);
  long long g = need < cap ? need : cap;
  return static_cast<int>(g > 0 ? g : 1);
}

int launch_argmax(const float* logits, int frames, int C, long long HW, uint8_t* u8, long long* i64, cudaStream_t st) {
  const bool vec4 = (HW % 4 == 0) && aligned16(logits) && (!u8 || aligned4(u8));
  const int threads = 256;
  if (vec4) {
    const int grid = persistent_grid(argmax_kernel<4>, (HW / 4) * frames, threads);
    argmax_kernel<4><<<grid, threads, 0, st>>>(logits, frames, C, HW, u8, i64);
  } else {
    const int grid = persistent_grid(argmax_kernel<1>, HW * frames, threads);
    argmax_kernel<1><<<grid, threads, 0, st>>>(logits, frames, C, HW, u8, i64);
  }
  return check_launch("fuvs_argmax");
}

}  // namespace fuvs

extern "C" int fuvs_blend_argmax(const float* a, const float* b, double wa, double wb, int frames, int C,
                                 long long HW, float* out, uint8_t* labels, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!a || frames < 0 || C < 1 || HW < 0) return set_error(FUVS_EINVAL, "blend: bad arguments frames=%d C=%d HW=%lld", frames, C, HW);
  if (labels && C > 256) return set_error(FUVS_EINVAL, "blend: uint8 label maps need C <= 256 (C=%d)", C);
  if (frames == 0 || HW == 0 || (!out && !labels)) return FUVS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float fa = static_cast<float>(wa), fb = static_cast<float>(wb);   // ATen: scalar -> opmath_t(float)
  const bool vec4 = (HW % 4 == 0) && aligned16(a) && (!b || aligned16(b)) && (!out || aligned16(out)) &&
                    (!labels || aligned4(labels));
  const int threads = 256;
  if (vec4) {
    const int grid = persistent_grid(blend_argmax_kernel<4>, (HW / 4) * frames, threads);
    blend_argmax_kernel<4><<<grid, threads, 0, st>>>(a, b, fa, fb, frames, C, HW, out, labels);
  } else {
    const int grid = persistent_grid(blend_argmax_kernel<1>, HW * frames, threads);
    blend_argmax_kernel<1><<<grid, threads, 0, st>>>(a, b, fa, fb, frames, C, HW, out, labels);
  }
  return check_launch("fuvs_blend_argmax");
}

extern "C" int fuvs_upsample_argmax(const float* logits, int frames, int C, int Hin, int Win, int Hout, int Wout,
                                    uint8_t* labels, float* resized, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!logits || frames < 0 || C < 1 || Hin < 1 || Win < 1 || Hout < 0 || Wout < 0)
    return set_error(FUVS_EINVAL, "upsample_argmax: bad arguments frames=%d C=%d in=%dx%d out=%dx%d", frames, C, Hin, Win, Hout, Wout);
  if (labels && C > 256) return set_error(FUVS_EINVAL, "upsample_argmax: uint8 label maps need C <= 256 (C=%d)", C);
  if (static_cast<long long>(Hin) * Win >= (1ll << 31)) return set_error(FUVS_EINVAL, "upsample_argmax: source plane exceeds 2^31 elements");
  if (frames == 0 || Hout == 0 || Wout == 0 || (!labels && !resized)) return FUVS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Hin == Hout && Win == Wout) {          // ATen's upsample copies when the sizes match
    const long long HW = static_cast<long long>(Hin) * Win;
    if (resized && cudaMemcpyAsync(resized, logits, static_cast<size_t>(frames) * C * HW * sizeof(float),
                                   cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return set_error(FUVS_ECUDA, "upsample_argmax: copy failed");
    return labels ? launch_argmax(logits, frames, C, HW, labels, nullptr, st) : FUVS_OK;
  }
  const float sh = Hout > 1 ? static_cast<float>(Hin - 1) / (Hout - 1) : 0.f;   // area_pixel_compute_scale, align_corners
  const float sw = Wout > 1 ? static_cast<float>(Win - 1) / (Wout - 1) : 0.f;
  const long long items = static_cast<long long>(frames) * Hout * ((Wout + 3) / 4);
  const int threads = 256;
  switch (C) {
    case 2: upsample_argmax_kernel<2><<<persistent_grid(upsample_argmax_kernel<2>, items, threads), threads, 0, st>>>(
                logits, frames, C, Hin, Win, Hout, Wout, sh, sw, labels, resized); break;
    case 5: upsample_argmax_kernel<5><<<persistent_grid(upsample_argmax_kernel<5>, items, threads), threads, 0, st>>>(
                logits, frames, C, Hin, Win, Hout, Wout, sh, sw, labels, resized); break;
    default: upsample_argmax_kernel<0><<<persistent_grid(upsample_argmax_kernel<0>, items, threads), threads, 0, st>>>(
                 logits, frames, C, Hin, Win, Hout, Wout, sh, sw, labels, resized); break;
  }
  return check_launch("fuvs_upsample_argmax");
}

extern "C" int fuvs_argmax(const float* logits, int frames, int C, long long HW, uint8_t* labels_u8,
                           long long* labels_i64, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!logits || frames < 0 || C < 1 || HW < 0) return set_error(FUVS_EINVAL, "argmax: bad arguments frames=%d C=%d HW=%lld", frames, C, HW);
  if (labels_u8 && C > 256) return set_error(FUVS_EINVAL, "argmax: uint8 label maps need C <= 256 (C=%d)", C);
  if (frames == 0 || HW == 0 || (!labels_u8 && !labels_i64)) return FUVS_OK;
  return launch_argmax(logits, frames, C, HW, labels_u8, labels_i64, static_cast<cudaStream_t>(stream));
}
