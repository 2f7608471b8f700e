// dense_common.cuh — launch description shared by the two dense-step kernels
// (warp.cu: direct L1 gather; dense_tma.cu: TMA-staged shared-memory gather).
#pragma once
#include "fuvs_common.cuh"

namespace fuvs {

// One lock-step warp step j of a dense interval (see warp.cu header comment).
struct DenseStep {
  const float* srcL; const float* srcR;      // states j-1, [C,H,W]
  const float* gridL; const float* gridR;    // [H,W,2]
  float* dstL; float* dstR;                  // states j or NULL (last step)
  // frame A = frame j   : wA0 * L_j(reg)      + wA1 * R_{n-j} (pointR, or this step's R if NULL)
  // frame B = frame n-j : wB0 * L_{n-j}(pointL) + wB1 * R_j(reg)
  int emitA, emitB;
  const float* pointR; const float* pointL;
  float wA0, wA1, wB0, wB1;
  uint8_t* labelA; uint8_t* labelB;
  float* logitA; float* logitB;
  // frame 0 on step 1
  const float* key0; uint8_t* label0; float* logit0;
  // 1: the chain states of this interval (dst*, point*, and src* unless key0 is set) are stored "4+1" — channels 0-3
  // interleaved per pixel ([H][W][4]) followed by the plane of channel 4 (C = 5, strip kernel only, dense_strip.cu)
  int il;
  // 1: step 1 of a 4+1 interval whose key frames (srcL, srcR, key0) are 4+1 themselves (fuvs_dense_lowres_interval)
  int key_il;
};

// dense_tma.cu: returns FUVS_OK if it ran the step, 1 if the shape is not eligible (caller uses the direct kernel),
// negative on error.
int launch_dense_step_tma(const DenseStep& a, int C, int H, int W, cudaStream_t st);

// dense_strip.cu (column-strip sliding window, all channels resident in shared memory): same return convention.
int launch_dense_step_strip(const DenseStep& a, int C, int H, int W, cudaStream_t st);
// true if every step of the interval can run on the strip kernel with 4+1 chain states (decided once per interval)
bool dense_strip_il_ok(int C, int H, int W, int n, const float* prev, const float* next, const float* gridsL,
                       const float* gridsR, const float* scratch);

// block.cu: key frame [C,hl,wl] -> [C,H,W] (bilinear, align_corners=True), planar or 4+1 (C = 5)
int launch_upsample_keyframe(const float* src, float* dst, int C, int hl, int wl, int H, int W, bool il, cudaStream_t st);

}  // namespace fuvs
