// linear.cu — fused linear temporal blend + arg-max + temporal-consistency
// counts for one interval (model.no_warp=True, model.feature_based=False).
//
// Replaces flow/model.py:231-239 with warp()==identity (flow/model.py:244-249),
// flow/base.py:276-277 and flow/base.py:280-295 + util/util.py:52-63.
//
// HBM traffic per interval: read prev+next once (2S), write n
// uint8 label maps, read the previous interval's last label map.  Each thread
// owns VEC consecutive pixels x all C class planes: C coalesced 128-bit loads
// per key frame, all n frames computed from registers.
#include "fuvs_common.cuh"

namespace fuvs {

template <int CT, int VEC, bool COUNTS>
__global__ void __launch_bounds__(256)
linear_blend_argmax_kernel(const float* __restrict__ prev, const float* __restrict__ next,
                           long long HW, int n,
                           uint8_t* __restrict__ labels, float* __restrict__ logits,
                           const uint8_t* __restrict__ tc_prev,
                           unsigned long long* __restrict__ counts, int ignore_index,
                           const BlendWeights wts) {
  __shared__ unsigned sh[24];
  const long long nvec = HW / VEC;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nround = (nvec + 31) & ~31ll;   // whole warps iterate together (REDUX)

  Hist3Packed hist;
  WarpTotals<CT> tot;
  hist.clear();
  tot.clear();

  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nround; v += stride) {
    if (v < nvec) {
      const long long pix = v * VEC;
      FVec<VEC> a[CT], b[CT];
#pragma unroll
      for (int c = 0; c < CT; ++c) a[c].load_stream(prev + c * HW + pix);
      if (n > 1) {
#pragma unroll
        for (int c = 0; c < CT; ++c) b[c].load_stream(next + c * HW + pix);
      }
      LVec<VEC> last;
      bool have_last = false;
      if (COUNTS && tc_prev != nullptr) {
        last.load(tc_prev + pix);
        have_last = true;
      }
      // frame 0: the unblended key frame (flow/model.py:195-197)
      {
        LVec<VEC> lab;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          ArgMax am;
          am.init(a[0].v[i]);
#pragma unroll
          for (int c = 1; c < CT; ++c) am.push(a[c].v[i], c);
          lab.v[i] = am.idx;
        }
        if (labels) lab.store(labels + pix);
        if (logits) {
#pragma unroll
          for (int c = 0; c < CT; ++c) a[c].store_stream(logits + c * HW + pix);
        }
        if (COUNTS) {
          if (have_last) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) hist.add(lab.v[i], last.v[i], ignore_index, CT);
          }
          last = lab;
        }
      }
      for (int p = 1; p < n; ++p) {
        const float w0 = wts.w0[p], w1 = wts.w1[p];
        LVec<VEC> lab;
        ArgMax am[VEC];
        float* lg = logits ? logits + (static_cast<long long>(p) * CT) * HW + pix : nullptr;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          FVec<VEC> o;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            o.v[i] = blend2(w0, a[c].v[i], w1, b[c].v[i]);
            if (c == 0) am[i].init(o.v[i]); else am[i].push(o.v[i], c);
          }
          if (lg) o.store_stream(lg + c * HW);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) lab.v[i] = am[i].idx;
        if (labels) lab.store(labels + static_cast<long long>(p) * HW + pix);
        if (COUNTS) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) hist.add(lab.v[i], last.v[i], ignore_index, CT);
          last = lab;
        }
      }
    }
    if (COUNTS) warp_accumulate<CT>(hist, tot);
  }
  if (COUNTS) block_flush_counts<CT>(tot, sh, counts, CT);
}

// Generic class count (C <= 256 when labels/counts are requested): class loop
// inside the frame loop, key-frame values re-read through L1.
template <int VEC>
__global__ void __launch_bounds__(256)
linear_blend_argmax_generic_kernel(const float* __restrict__ prev, const float* __restrict__ next,
                                   int C, long long HW, int n,
                                   uint8_t* __restrict__ labels, float* __restrict__ logits,
                                   const uint8_t* __restrict__ tc_prev,
                                   unsigned long long* __restrict__ counts, int ignore_index,
                                   const BlendWeights wts) {
  __shared__ unsigned sh[768];
  const bool do_counts = counts != nullptr;
  if (do_counts) smem_hist_clear(sh);
  const long long nvec = HW / VEC;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const long long pix = v * VEC;
    LVec<VEC> last;
    bool have_last = false;
    if (do_counts && tc_prev != nullptr) {
      last.load(tc_prev + pix);
      have_last = true;
    }
    for (int p = 0; p < n; ++p) {
      const float w0 = wts.w0[p], w1 = wts.w1[p];
      ArgMax am[VEC];
      float* lg = logits ? logits + (static_cast<long long>(p) * C) * HW + pix : nullptr;
      for (int c = 0; c < C; ++c) {
        FVec<VEC> a, o;
        a.load(prev + c * HW + pix);
        if (p == 0) {
          o = a;
        } else {
          FVec<VEC> b;
          b.load(next + c * HW + pix);
#pragma unroll
          for (int i = 0; i < VEC; ++i) o.v[i] = blend2(w0, a.v[i], w1, b.v[i]);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (c == 0) am[i].init(o.v[i]); else am[i].push(o.v[i], c);
        }
        if (lg) o.store_stream(lg + c * HW);
      }
      LVec<VEC> lab;
#pragma unroll
      for (int i = 0; i < VEC; ++i) lab.v[i] = am[i].idx;
      if (labels) lab.store(labels + static_cast<long long>(p) * HW + pix);
      if (do_counts) {
        if (p > 0 || have_last) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) smem_hist_add(sh, lab.v[i], last.v[i], ignore_index, C);
        }
        last = lab;
      }
    }
  }
  if (do_counts) smem_hist_flush(sh, counts, C);
}

template <int CT, int VEC>
static int launch_fixed(const float* prev, const float* next, long long HW, int n, uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts, int ignore_index, const BlendWeights& w,
                        cudaStream_t st) {
  const long long nvec = HW / VEC;
  const int threads = 256;
  long long need = (nvec + threads - 1) / threads;
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  if (counts) {
    static int bps = blocks_per_sm(linear_blend_argmax_kernel<CT, VEC, true>, threads);
    const long long cap = static_cast<long long>(sm_count()) * bps;
    const int grid = static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
    linear_blend_argmax_kernel<CT, VEC, true><<<grid, threads, 0, st>>>(prev, next, HW, n, labels, logits, tc_prev, cu,
                                                                        ignore_index, w);
  } else {
    static int bps = blocks_per_sm(linear_blend_argmax_kernel<CT, VEC, false>, threads);
    const long long cap = static_cast<long long>(sm_count()) * bps;
    const int grid = static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
    linear_blend_argmax_kernel<CT, VEC, false><<<grid, threads, 0, st>>>(prev, next, HW, n, labels, logits, nullptr,
                                                                         nullptr, ignore_index, w);
  }
  return check_launch("fuvs_linear_blend_argmax");
}

template <int VEC>
static int launch_by_c(int C, const float* prev, const float* next, long long HW, int n, uint8_t* labels,
                       float* logits, const uint8_t* tc_prev, long long* counts, int ignore_index,
                       const BlendWeights& w, cudaStream_t st) {
  switch (C) {
    case 2: return launch_fixed<2, VEC>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 3: return launch_fixed<3, VEC>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 4: return launch_fixed<4, VEC>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 5: return launch_fixed<5, VEC>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 8: return launch_fixed<8, VEC>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    default: break;
  }
  const long long nvec = HW / VEC;
  const int threads = 256;
  long long need = (nvec + threads - 1) / threads;
  static int bps = blocks_per_sm(linear_blend_argmax_generic_kernel<VEC>, threads);
  const long long cap = static_cast<long long>(sm_count()) * bps;
  const int grid = static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
  linear_blend_argmax_generic_kernel<VEC><<<grid, threads, 0, st>>>(
      prev, next, C, HW, n, labels, logits, tc_prev, reinterpret_cast<unsigned long long*>(counts), ignore_index, w);
  return check_launch("fuvs_linear_blend_argmax(generic C)");
}

}  // namespace fuvs

extern "C" int fuvs_linear_blend_argmax(const float* prev, const float* next, int C, int H, int W, int n,
                                        uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts,
                                        int ignore_index, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!prev || C < 1 || H < 1 || W < 1 || n < 1) return set_error(FUVS_EINVAL, "linear: bad shape C=%d H=%d W=%d n=%d", C, H, W, n);
  if (n > 1 && !next) return set_error(FUVS_EINVAL, "linear: next key frame is NULL but n=%d", n);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "linear: n=%d exceeds %d frames per interval", n, FUVS_MAX_FRAMES);
  if ((labels || counts) && C > 256) return set_error(FUVS_EINVAL, "linear: uint8 label maps need C <= 256 (C=%d)", C);
  if (!labels && !logits && !counts) return FUVS_OK;
  if (H == 0 || W == 0) return FUVS_OK;
  const long long HW = static_cast<long long>(H) * W;
  BlendWeights w;
  make_blend_weights(n, &w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec4 = (HW % 4 == 0) && aligned16(prev) && (n == 1 || aligned16(next)) && (!logits || aligned16(logits)) &&
                    (!labels || aligned4(labels)) && (!tc_prev || aligned4(tc_prev));
  if (vec4) return launch_by_c<4>(C, prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
  return launch_by_c<1>(C, prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
}
