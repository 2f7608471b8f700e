// warp.cu — backward bilinear warp by a flow grid (F.grid_sample, bilinear,
// padding_mode="border") and the fused dense-flow interval.
//
// fuvs_warp_step      <- FlowModel.warp, flow/model.py:244-249 (align_corners
//                        False) and the default-grid resample flow/model.py:157
//                        (align_corners True)
// fuvs_dense_interval <- flow/model.py:208-239 + flow/base.py:276-277 when the
//                        grids are dense [H,W,2]
//
// Dense schedule: the forward chain L_j = warp(L_{j-1}, gl_j)
// and the backward chain R_j = warp(R_{j-1}, gr_j) advance in lock step, one
// launch per step.  Frame p = w0*L_p + w1*R_{n-p} is finished in the step
// max(p, n-p): the state produced in that step is still in registers, the
// other one is re-read pointwise.  States n-1 are never written; frame 0
// (arg-max of the key frame) rides on step 1.
#include <cstdlib>

#include "dense_common.cuh"

namespace fuvs {

int launch_temporal_counts(const uint8_t* labels, int n, long long HW, const uint8_t* tc_prev, int K,
                           int ignore_index, long long* counts, cudaStream_t st);
int launch_argmax(const float* logits, int frames, int C, long long HW, uint8_t* u8, long long* i64, cudaStream_t st);

// ---------------------------------------------------------------------------
// generic step: dst[c, y, x] = bilinear(src[c], grid[y, x])
// ---------------------------------------------------------------------------
struct WarpProblem {
  const float* src;
  const float* grid;
  float* dst;
};

template <class NM>
__global__ void __launch_bounds__(256)
warp_step_kernel(WarpProblem p0, WarpProblem p1, int C, int Hin, int Win, int Hg, int Wg, int align_corners,
                 int cchunk, int nchunks) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= Wg || y >= Hg) return;
  const int side = blockIdx.z / nchunks;
  const int chunk = blockIdx.z - side * nchunks;
  const WarpProblem P = side ? p1 : p0;
  const long long opix = static_cast<long long>(y) * Wg + x;
  const float2 g = __ldg(reinterpret_cast<const float2*>(P.grid) + opix);
  const GsTap t = gs_setup<NM>(g.x, g.y, Hin, Win, align_corners != 0);
  const long long in_plane = static_cast<long long>(Hin) * Win;
  const long long out_plane = static_cast<long long>(Hg) * Wg;
  const int c0 = chunk * cchunk;
  const int c1 = min(C, c0 + cchunk);
  // Batches of 8 channels: all 32 tap loads first, then the accumulations and stores.  With one channel per
  // iteration the store of channel c (which may alias the source as far as the compiler knows) kept the loads of
  // channel c+1 behind it in program order — 4 loads in flight per thread and a DRAM round trip per channel: 106 us
  // per step at [2048,67,120] (0.32 of the HBM roofline, tools/abi_bench.py).
  const float* __restrict__ srcp = P.src + t.off00;
  float* __restrict__ dstp = P.dst + opix;
  const int o01 = t.dx, o10 = t.dy * Win, o11 = t.dy * Win + t.dx;
  constexpr int U = 8;
  int c = c0;
  for (; c + U <= c1; c += U) {
    float v00[U], v01[U], v10[U], v11[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* q = srcp + (c + u) * in_plane;
      v00[u] = __ldg(q); v01[u] = __ldg(q + o01); v10[u] = __ldg(q + o10); v11[u] = __ldg(q + o11);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc = 0.f;
      acc = tap_acc<NM>(acc, v00[u], t.nw);
      if (t.dx) acc = tap_acc<NM>(acc, v01[u], t.ne);
      if (t.dy) acc = tap_acc<NM>(acc, v10[u], t.sw);
      if (t.dx & t.dy) acc = tap_acc<NM>(acc, v11[u], t.se);
      dstp[(c + u) * out_plane] = acc;
    }
  }
  for (; c < c1; ++c) dstp[c * out_plane] = gs_fetch<NM>(P.src + c * in_plane, t, Win);
}

template <class NM>
static int launch_warp_step(const float* src0, const float* grid0, float* dst0, const float* src1, const float* grid1,
                            float* dst1, int C, int Hin, int Win, int Hg, int Wg, int align_corners, cudaStream_t st) {
  const int sides = src1 ? 2 : 1;
  const int bx = (Wg + 31) / 32, by = (Hg + 7) / 8;
  // split the channel loop so that small grids with many channels (feature
  // maps: C=2048 at 67x120) still fill the 148 SMs a few times over
  const long long spatial_blocks = static_cast<long long>(bx) * by * sides;
  const long long want = 4ll * sm_count();
  int nchunks = 1;
  if (spatial_blocks < want) nchunks = static_cast<int>((want + spatial_blocks - 1) / spatial_blocks);
  if (nchunks > C) nchunks = C;
  int cchunk = (C + nchunks - 1) / nchunks;
  nchunks = (C + cchunk - 1) / cchunk;
  if (static_cast<long long>(sides) * nchunks > 65535 || by > 65535)
    return set_error(FUVS_EINVAL, "warp_step: grid too large (Hg=%d, C=%d)", Hg, C);
  dim3 grid(bx, by, sides * nchunks), block(32, 8);
  WarpProblem p0{src0, grid0, dst0}, p1{src1, grid1, dst1};
  warp_step_kernel<NM><<<grid, block, 0, st>>>(p0, p1, C, Hin, Win, Hg, Wg, align_corners, cchunk, nchunks);
  return check_launch("fuvs_warp_step");
}

// non-template entry for other translation units (feature.cu)
int launch_warp_step_nm(const float* src0, const float* grid0, float* dst0, const float* src1, const float* grid1,
                        float* dst1, int C, int Hin, int Win, int Hg, int Wg, int align_corners, cudaStream_t st) {
  return launch_warp_step<Nm>(src0, grid0, dst0, src1, grid1, dst1, C, Hin, Win, Hg, Wg, align_corners, st);
}

// ---------------------------------------------------------------------------
// dense lock-step kernel
// ---------------------------------------------------------------------------
template <class NM, int CT>
__global__ void __launch_bounds__(256)
dense_step_kernel(const DenseStep A, int Crt, int H, int W) {
  const int C = CT > 0 ? CT : Crt;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  const long long HW = static_cast<long long>(H) * W;
  const long long pix = static_cast<long long>(y) * W + x;
  const float2 gl = __ldg(reinterpret_cast<const float2*>(A.gridL) + pix);
  const float2 gr = __ldg(reinterpret_cast<const float2*>(A.gridR) + pix);
  const GsTap tl = gs_setup<NM>(gl.x, gl.y, H, W, false);
  const GsTap tr = gs_setup<NM>(gr.x, gr.y, H, W, false);
  ArgMax amA, amB, am0;
  amA.init(0.f); amB.init(0.f); am0.init(0.f);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const long long o = c * HW + pix;
    const float Lc = gs_fetch<NM>(A.srcL + c * HW, tl, W);
    const float Rc = gs_fetch<NM>(A.srcR + c * HW, tr, W);
    if (A.dstL) A.dstL[o] = Lc;
    if (A.dstR) A.dstR[o] = Rc;
    if (A.emitA) {
      const float r = A.pointR ? __ldg(A.pointR + o) : Rc;
      const float v = blend2(A.wA0, Lc, A.wA1, r);
      if (c == 0) amA.init(v); else amA.push(v, c);
      if (A.logitA) __stcs(A.logitA + o, v);
    }
    if (A.emitB) {
      const float l = __ldg(A.pointL + o);
      const float v = blend2(A.wB0, l, A.wB1, Rc);
      if (c == 0) amB.init(v); else amB.push(v, c);
      if (A.logitB) __stcs(A.logitB + o, v);
    }
    if (A.key0) {
      const float v = __ldg(A.key0 + o);
      if (c == 0) am0.init(v); else am0.push(v, c);
      if (A.logit0) __stcs(A.logit0 + o, v);
    }
  }
  if (A.emitA && A.labelA) A.labelA[pix] = static_cast<uint8_t>(amA.idx);
  if (A.emitB && A.labelB) A.labelB[pix] = static_cast<uint8_t>(amB.idx);
  if (A.key0 && A.label0) A.label0[pix] = static_cast<uint8_t>(am0.idx);
}

template <class NM>
static int launch_dense_step(const DenseStep& a, int C, int H, int W, cudaStream_t st) {
  dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
  if (grid.y > 65535) return set_error(FUVS_EINVAL, "dense: H=%d too large", H);
  switch (C) {
    case 2: dense_step_kernel<NM, 2><<<grid, block, 0, st>>>(a, C, H, W); break;
    case 5: dense_step_kernel<NM, 5><<<grid, block, 0, st>>>(a, C, H, W); break;
    default: dense_step_kernel<NM, 0><<<grid, block, 0, st>>>(a, C, H, W); break;
  }
  return check_launch("fuvs_dense_interval(step)");
}

}  // namespace fuvs

extern "C" int fuvs_warp_step(const float* src0, const float* grid0, float* dst0, const float* src1,
                              const float* grid1, float* dst1, int C, int Hin, int Win, int Hg, int Wg,
                              int align_corners, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!src0 || !grid0 || !dst0 || C < 1 || Hin < 1 || Win < 1 || Hg < 0 || Wg < 0)
    return set_error(FUVS_EINVAL, "warp_step: bad arguments C=%d in=%dx%d grid=%dx%d", C, Hin, Win, Hg, Wg);
  if ((src1 != nullptr) != (grid1 != nullptr) || (src1 != nullptr) != (dst1 != nullptr))
    return set_error(FUVS_EINVAL, "warp_step: second problem must be all-NULL or all-set");
  if (static_cast<long long>(Hin) * Win >= (1ll << 31))
    return set_error(FUVS_EINVAL, "warp_step: source plane exceeds 2^31 elements");
  if (!aligned8(grid0) || (grid1 && !aligned8(grid1))) return set_error(FUVS_EALIGN, "warp_step: grids must be 8-byte aligned");
  if (Hg == 0 || Wg == 0) return FUVS_OK;
  return launch_warp_step<Nm>(src0, grid0, dst0, src1, grid1, dst1, C, Hin, Win, Hg, Wg, align_corners,
                              static_cast<cudaStream_t>(stream));
}

extern "C" long long fuvs_dense_scratch_floats(int C, int H, int W, int n) {
  if (n <= 2) return 0;
  return 2ll * (n - 2) * C * static_cast<long long>(H) * W;
}

extern "C" int fuvs_dense_interval(const float* prev, const float* next, const float* grids_left,
                                   const float* grids_right, int C, int H, int W, int n, float* scratch,
                                   uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts,
                                   int ignore_index, fuvs_stream_t stream) {
  const float* gl[FUVS_MAX_FRAMES];
  const float* gr[FUVS_MAX_FRAMES];
  if (n > FUVS_MAX_FRAMES) return fuvs::set_error(FUVS_EINVAL, "dense: n=%d exceeds %d frames per interval", n, FUVS_MAX_FRAMES);
  if (n > 1 && (!grids_left || !grids_right)) return fuvs::set_error(FUVS_EINVAL, "dense: next/grids are NULL but n=%d", n);
  const long long g = static_cast<long long>(H) * W * 2;
  for (int j = 0; j < n - 1; ++j) {
    gl[j] = grids_left + j * g;
    gr[j] = grids_right + j * g;
  }
  return fuvs_dense_interval_ptrs(prev, next, gl, gr, C, H, W, n, scratch, labels, logits, tc_prev, counts, ignore_index, stream);
}

namespace fuvs {
// lowres: 0 = prev / next are full-resolution planar key frames (fuvs_dense_interval);
//         1 = they are caller-owned buffers that receive the up-sample of prev_lr / next_lr [C,hl,wl] first
//             (fuvs_dense_lowres_interval; prev_ready: prev already holds it), 4+1 when the interval runs 4+1
static int dense_interval_impl(const float* prev, const float* next, const float* const* grids_left,
                               const float* const* grids_right, int C, int H, int W, int n, float* scratch,
                               uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts,
                               int ignore_index, fuvs_stream_t stream, int lowres, const float* prev_lr,
                               const float* next_lr, int hl, int wl, int prev_ready) {
  if (int e = device_ok()) return e;
  if (!prev || C < 1 || H < 1 || W < 1 || n < 1) return set_error(FUVS_EINVAL, "dense: bad shape C=%d H=%d W=%d n=%d", C, H, W, n);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "dense: n=%d exceeds %d frames per interval", n, FUVS_MAX_FRAMES);
  if (n > 1 && (!next || !grids_left || !grids_right)) return set_error(FUVS_EINVAL, "dense: next/grids are NULL but n=%d", n);
  if (n > 2 && !scratch) return set_error(FUVS_EINVAL, "dense: scratch is NULL (need %lld floats)", fuvs_dense_scratch_floats(C, H, W, n));
  if ((labels || counts) && C > 256) return set_error(FUVS_EINVAL, "dense: uint8 label maps need C <= 256 (C=%d)", C);
  if (counts && !labels) return set_error(FUVS_EINVAL, "dense: counts need the label maps (labels is NULL)");
  const long long HW = static_cast<long long>(H) * W;
  if (HW >= (1ll << 31)) return set_error(FUVS_EINVAL, "dense: plane exceeds 2^31 elements");
  for (int j = 0; n > 1 && j < n - 1; ++j) {
    if (!grids_left[j] || !grids_right[j]) return set_error(FUVS_EINVAL, "dense: grid %d is NULL", j);
    if (!aligned8(grids_left[j]) || !aligned8(grids_right[j])) return set_error(FUVS_EALIGN, "dense: grids must be 8-byte aligned");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long S = static_cast<long long>(C) * HW;
  // C = 5 with odd n keeps its chain states in the strip kernel's 4+1 layout; that is decided here, once, because a
  // state written 4+1 must be read 4+1 (only when frames are emitted: step 1 is then the key-frame variant)
  const bool il = n > 1 && (labels || logits) &&
                  dense_strip_il_ok(C, H, W, n, prev, next, grids_left[0], grids_right[0], scratch);
  if (lowres) {
    // flow/model.py:191-193, 205-206: the decoder output up-sampled to the frame size, straight into the layout step 1 reads
    if (!prev_ready) {
      if (int e = launch_upsample_keyframe(prev_lr, const_cast<float*>(prev), C, hl, wl, H, W, il, st)) return e;
    }
    if (n > 1) {
      if (int e = launch_upsample_keyframe(next_lr, const_cast<float*>(next), C, hl, wl, H, W, il, st)) return e;
    }
  }

  if (n == 1) {
    if (labels) {
      if (int e = launch_argmax(prev, 1, C, HW, labels, nullptr, st)) return e;
    }
    if (logits) {
      if (cudaMemcpyAsync(logits, prev, S * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return set_error(FUVS_ECUDA, "dense: logits copy failed");
    }
  } else {
    BlendWeights w;
    make_blend_weights(n, &w);
    // Step kernels, in order of preference: strip (sliding-window TMA kernel, dense_strip.cu), plane (per-plane TMA
    // kernel, dense_tma.cu: C > 6), direct (L1 gather, this file: W % 4 != 0, even-n middle step).  Shapes a kernel
    // cannot take fall through to the next one (`il`, decided above, excludes that for 4+1 intervals).
    float* Lst = scratch;                       // states 1..n-2
    float* Rst = scratch + (n > 2 ? (n - 2) * S : 0);
    for (int j = 1; j <= n - 1; ++j) {
      DenseStep a{};
      a.srcL = (j == 1) ? prev : Lst + (j - 2) * S;
      a.srcR = (j == 1) ? next : Rst + (j - 2) * S;
      a.gridL = grids_left[j - 1];
      a.gridR = grids_right[j - 1];
      a.dstL = (j <= n - 2) ? Lst + (j - 1) * S : nullptr;
      a.dstR = (j <= n - 2) ? Rst + (j - 1) * S : nullptr;
      const bool want_out = labels || logits;
      if (want_out && 2 * j >= n) {
        const int pA = j, pB = n - j;
        a.emitA = 1;
        a.pointR = (pB == j) ? nullptr : Rst + (pB - 1) * S;   // R_{n-pA}
        a.wA0 = w.w0[pA]; a.wA1 = w.w1[pA];
        a.labelA = labels ? labels + pA * HW : nullptr;
        a.logitA = logits ? logits + pA * S : nullptr;
        if (pB != pA) {
          a.emitB = 1;
          a.pointL = Lst + (pB - 1) * S;                         // L_{pB}
          a.wB0 = w.w0[pB]; a.wB1 = w.w1[pB];
          a.labelB = labels ? labels + pB * HW : nullptr;
          a.logitB = logits ? logits + pB * S : nullptr;
        }
      }
      if (j == 1 && want_out) {
        a.key0 = prev;
        a.label0 = labels;
        a.logit0 = logits;
      }
      a.il = il ? 1 : 0;
      a.key_il = (il && lowres && j == 1) ? 1 : 0;
      int r = launch_dense_step_strip(a, C, H, W, st);
      if (r > 0) r = launch_dense_step_tma(a, C, H, W, st);
      if (r < 0) return r;
      if (r > 0) {   // not eligible for the TMA-staged kernels: direct-gather kernel
        if (int e = launch_dense_step<Nm>(a, C, H, W, st)) return e;
      }
    }
  }
  if (counts) return launch_temporal_counts(labels, n, HW, tc_prev, C, ignore_index, counts, st);
  return FUVS_OK;
}
}  // namespace fuvs

extern "C" int fuvs_dense_interval_ptrs(const float* prev, const float* next, const float* const* grids_left,
                                        const float* const* grids_right, int C, int H, int W, int n, float* scratch,
                                        uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts,
                                        int ignore_index, fuvs_stream_t stream) {
  return fuvs::dense_interval_impl(prev, next, grids_left, grids_right, C, H, W, n, scratch, labels, logits, tc_prev, counts,
                                   ignore_index, stream, 0, nullptr, nullptr, 0, 0, 0);
}

extern "C" int fuvs_dense_lowres_interval_ptrs(const float* prev_lr, const float* next_lr, int hl, int wl, float* prev_up,
                                               int prev_up_ready, float* next_up, const float* const* grids_left,
                                               const float* const* grids_right, int C, int H, int W, int n,
                                               float* scratch, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                                               long long* counts, int ignore_index, fuvs_stream_t stream) {
  using namespace fuvs;
  if (!prev_lr || !prev_up || hl < 1 || wl < 1 || C < 1)
    return set_error(FUVS_EINVAL, "dense_lowres: bad key frame C=%d %dx%d (prev_lr / prev_up NULL?)", C, hl, wl);
  if (n > 1 && (!next_lr || !next_up)) return set_error(FUVS_EINVAL, "dense_lowres: next_lr / next_up are NULL but n=%d", n);
  if (prev_up == next_up) return set_error(FUVS_EINVAL, "dense_lowres: prev_up and next_up must be different buffers");
  if (static_cast<long long>(C) * hl * wl >= (1ll << 31)) return set_error(FUVS_EINVAL, "dense_lowres: key frame exceeds 2^31 elements");
  return dense_interval_impl(prev_up, next_up, grids_left, grids_right, C, H, W, n, scratch, labels, logits, tc_prev, counts,
                             ignore_index, stream, 1, prev_lr, next_lr, hl, wl, prev_up_ready);
}
