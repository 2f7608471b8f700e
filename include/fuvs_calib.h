/*
 * fuvs_calib.h — calibration entry points of libfuvs.so (test infrastructure
 * inside the product library; not part of the drop-in surface).
 *
 * ATen's grid_sampler_2d / upsample_bilinear2d sources leave floating-point
 * contraction to nvcc.  These two calls run the same kernels as
 * fuvs_warp_step / fuvs_upsample_bilinear_ac with every candidate contraction
 * selectable at run time, so tests/test_calibration_gpu.py can prove on a B200
 * that exactly the compiled-in default (fuvs_common.cuh `Nm`) reproduces the
 * installed torch binary bit for bit.
 */
#ifndef FUVS_CALIB_H_
#define FUVS_CALIB_H_
#include "fuvs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* F.grid_sample(bilinear, border): unnorm_fma, tap_fma in {0,1} */
FUVS_API int fuvs_calib_grid_sample(const float* src, const float* grid, float* dst,
                           int C, int Hin, int Win, int Hg, int Wg, int align_corners,
                           int unnorm_fma, int tap_fma, fuvs_stream_t stream);
/* F.interpolate(bilinear, align_corners=True): lambda_fma in {0,1}; inner, outer in {0,1,2} */
FUVS_API int fuvs_calib_upsample(const float* src, float* dst, long long planes,
                        int Hin, int Win, int Hout, int Wout,
                        int lambda_fma, int inner, int outer, fuvs_stream_t stream);
/* Encodes the compiled-in default as unnorm*1000 + tap*100 + lambda*10... :
 * returns unnorm_fma<<0 | tap_fma<<1 | lambda_fma<<2 | inner<<3 | outer<<5 */
FUVS_API int fuvs_calib_default(void);

#ifdef __cplusplus
}
#endif
#endif
