// calib.cu — run-time selectable contraction variants of the two ATen kernels
// whose FMA placement is decided by nvcc (see include/fuvs_calib.h).
#include "../../include/fuvs_calib.h"
#include "fuvs_common.cuh"

namespace fuvs {

template <class NM>
__global__ void __launch_bounds__(256)
calib_gs_kernel(const float* __restrict__ src, const float* __restrict__ grid, float* __restrict__ dst, int C, int Hin,
                int Win, int Hg, int Wg, int align_corners) {
  const int opix = blockIdx.x * blockDim.x + threadIdx.x;
  if (opix >= Hg * Wg) return;
  const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + opix);
  const GsTap t = gs_setup<NM>(g.x, g.y, Hin, Win, align_corners != 0);
  for (int c = 0; c < C; ++c)
    dst[static_cast<long long>(c) * Hg * Wg + opix] = gs_fetch<NM>(src + static_cast<long long>(c) * Hin * Win, t, Win);
}

template <class NM>
__global__ void __launch_bounds__(256)
calib_up_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes, int Hin, int Win, int Hout,
                int Wout, float sh, float sw) {
  const int opix = blockIdx.x * blockDim.x + threadIdx.x;
  if (opix >= Hout * Wout) return;
  const int y = opix / Wout, x = opix - y * Wout;
  const UpCoord hc = up_coord<NM>(sh, y, Hin), wc = up_coord<NM>(sw, x, Win);
  for (long long pl = 0; pl < planes; ++pl)
    dst[pl * Hout * Wout + opix] = up_fetch<NM>(src + pl * Hin * Win, Win, hc, wc);
}

template <int U, int T>
static void run_gs(const float* src, const float* grid, float* dst, int C, int Hin, int Win, int Hg, int Wg, int ac,
                   cudaStream_t st) {
  const int n = Hg * Wg;
  calib_gs_kernel<NumericsT<U, T, 0, 0, 0>><<<(n + 255) / 256, 256, 0, st>>>(src, grid, dst, C, Hin, Win, Hg, Wg, ac);
}

template <int L, int I, int O>
static void run_up(const float* src, float* dst, long long planes, int Hin, int Win, int Hout, int Wout, float sh,
                   float sw, cudaStream_t st) {
  const int n = Hout * Wout;
  calib_up_kernel<NumericsT<1, 1, L, I, O>><<<(n + 255) / 256, 256, 0, st>>>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw);
}

template <int L, int I>
static void run_up_o(int outer, const float* src, float* dst, long long planes, int Hin, int Win, int Hout, int Wout,
                     float sh, float sw, cudaStream_t st) {
  if (outer == 0) run_up<L, I, 0>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
  else if (outer == 1) run_up<L, I, 1>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
  else run_up<L, I, 2>(src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
}

template <int L>
static void run_up_io(int inner, int outer, const float* src, float* dst, long long planes, int Hin, int Win, int Hout,
                      int Wout, float sh, float sw, cudaStream_t st) {
  if (inner == 0) run_up_o<L, 0>(outer, src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
  else if (inner == 1) run_up_o<L, 1>(outer, src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
  else run_up_o<L, 2>(outer, src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
}

}  // namespace fuvs

extern "C" int fuvs_calib_grid_sample(const float* src, const float* grid, float* dst, int C, int Hin, int Win, int Hg,
                                      int Wg, int align_corners, int unnorm_fma, int tap_fma, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!src || !grid || !dst || C < 1 || Hin < 1 || Win < 1 || Hg < 1 || Wg < 1)
    return set_error(FUVS_EINVAL, "calib_grid_sample: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (unnorm_fma && tap_fma) run_gs<1, 1>(src, grid, dst, C, Hin, Win, Hg, Wg, align_corners, st);
  else if (unnorm_fma) run_gs<1, 0>(src, grid, dst, C, Hin, Win, Hg, Wg, align_corners, st);
  else if (tap_fma) run_gs<0, 1>(src, grid, dst, C, Hin, Win, Hg, Wg, align_corners, st);
  else run_gs<0, 0>(src, grid, dst, C, Hin, Win, Hg, Wg, align_corners, st);
  return check_launch("fuvs_calib_grid_sample");
}

extern "C" int fuvs_calib_upsample(const float* src, float* dst, long long planes, int Hin, int Win, int Hout, int Wout,
                                   int lambda_fma, int inner, int outer, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!src || !dst || planes < 1 || Hin < 1 || Win < 1 || Hout < 1 || Wout < 1 || inner < 0 || inner > 2 || outer < 0 ||
      outer > 2)
    return set_error(FUVS_EINVAL, "calib_upsample: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sh = Hout > 1 ? static_cast<float>(Hin - 1) / (Hout - 1) : 0.f;
  const float sw = Wout > 1 ? static_cast<float>(Win - 1) / (Wout - 1) : 0.f;
  if (lambda_fma) run_up_io<1>(inner, outer, src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
  else run_up_io<0>(inner, outer, src, dst, planes, Hin, Win, Hout, Wout, sh, sw, st);
  return check_launch("fuvs_calib_upsample");
}

extern "C" int fuvs_calib_default(void) {
  using fuvs::Nm;
  return Nm::kUnnormFma | (Nm::kTapFma << 1) | (Nm::kUpLambdaFma << 2) | (Nm::kUpInner << 3) | (Nm::kUpOuter << 5);
}
