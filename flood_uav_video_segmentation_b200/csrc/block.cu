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
#include <cooperative_groups.h>

#include <cstdlib>

#include "fuvs_common.cuh"

namespace fuvs {

int launch_block_stream_rows(const float* key0, const float* Lst, const float* Rst, int C, int H, int W, int Hg, int Wg,
                             int n, float sh, float sw, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                             long long* counts, int ignore_index, const BlendWeights& w, cudaStream_t st, int hl, int wl,
                             bool* counts_done);
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

// Wout % 4 == 0 and a 16-byte aligned destination: 4 consecutive pixels per thread (one row coordinate, one 128-bit
// store), all planes of a pixel group in one thread so the coordinate arithmetic is paid once.  Same up_fetch<>
// arithmetic as above.  (The key frames of the dense / block routes come through here: 41 MB per 1080p key frame.)
template <class NM>
__global__ void __launch_bounds__(256)
upsample_bilinear_ac_v4_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes, int Hin,
                               int Win, int Hout, int Wout, float sh, float sw) {
  const int groups = Wout >> 2;
  const long long items = static_cast<long long>(Hout) * groups;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= items) return;
  const int y = static_cast<int>(t / groups), x = static_cast<int>(t - static_cast<long long>(y) * groups) << 2;
  const UpCoord hc = up_coord<NM>(sh, y, Hin);
  UpCoord wc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wc[i] = up_coord<NM>(sw, x + i, Win);
  const long long in_plane = static_cast<long long>(Hin) * Win, out_plane = static_cast<long long>(Hout) * Wout;
  float* o = dst + static_cast<long long>(y) * Wout + x;
  for (long long pl = blockIdx.y; pl < planes; pl += gridDim.y) {
    const float* sp = src + pl * in_plane;
    float4 v;
    v.x = up_fetch<NM>(sp, Win, hc, wc[0]);
    v.y = up_fetch<NM>(sp, Win, hc, wc[1]);
    v.z = up_fetch<NM>(sp, Win, hc, wc[2]);
    v.w = up_fetch<NM>(sp, Win, hc, wc[3]);
    *reinterpret_cast<float4*>(o + pl * out_plane) = v;
  }
}

// Staged variant for up-sampling ratios >= 4 (r02).  The kernels above evaluate four taps and three two-terms per output
// value: 19.5 us for one 5 x 1080p key frame (41.5 MB written: 0.32 of the HBM roofline).  upsample_bilinear2d is
// separable in the order ATen evaluates it — val = h0 * (w0*a + w1*b) + h1 * (w0*c + w1*d) — so a CTA that owns one
// source-row interval (the output rows whose floor source row is i0), a chunk of columns and up to UR_PLANES planes
// computes the horizontal two-terms of source rows i0 and i0 + 1 once into shared memory; every output value is then
// one vertical two-term on packed FP32x2 straight from two 128-bit shared-memory loads, like block_rows.cu.
// IL5: 5 planes, written in the dense strip kernel's 4+1 layout (channels 0-3 interleaved per pixel, channel 4 a plane).
constexpr int UR_THREADS = 256;
constexpr int UR_PLANES = 8;
constexpr int UR_XW = 256;

template <class NM, bool IL5>
__global__ void __launch_bounds__(UR_THREADS)
upsample_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes, int Hin, int Win, int Hout,
                     int Wout, float sh, float sw, int nchunks, float one) {
  __shared__ __align__(16) float hs[2 * UR_PLANES * UR_XW];            // [row][plane][column]
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x / nchunks, chunk = blockIdx.x - i0 * nchunks;
  const int x0 = chunk * UR_XW;
  const int xw = min(UR_XW, Wout - x0);
  const long long p0 = static_cast<long long>(blockIdx.y) * UR_PLANES;
  const int pc = static_cast<int>(min(static_cast<long long>(UR_PLANES), planes - p0));
  const long long in_plane = static_cast<long long>(Hin) * Win, out_plane = static_cast<long long>(Hout) * Wout;
  // output rows of this interval: floor(sh * y) == i0 with the float arithmetic of up_coord()
  auto src_row = [&](int y) { return static_cast<int>(__fmul_rn(sh, static_cast<float>(y))); };
  int y_lo = static_cast<int>(static_cast<float>(i0) / sh);
  y_lo = max(0, min(y_lo, Hout - 1));
  while (y_lo > 0 && src_row(y_lo - 1) >= i0) --y_lo;
  while (y_lo < Hout && src_row(y_lo) < i0) ++y_lo;
  int y_hi = y_lo;
  while (y_hi < Hout && src_row(y_hi) == i0) ++y_hi;
  if (y_hi <= y_lo) return;
  const int ip_h = (i0 < Hin - 1) ? 1 : 0;
  // ---- phase 1: horizontal two-terms of source rows i0, i0 + ip_h; item = (plane, column)
  for (int e = tid; e < pc * xw; e += UR_THREADS) {
    const int pl = e / xw, xx = e - pl * xw;
    const UpCoord wc = up_coord<NM>(sw, x0 + xx, Win);
    const float* sp = src + (p0 + pl) * in_plane + wc.i0;
    const int o0 = i0 * Win, o1 = (i0 + ip_h) * Win;
    hs[(0 * UR_PLANES + pl) * UR_XW + xx] = two_term<NM::kUpInner>(wc.l0, __ldg(sp + o0), wc.l1, __ldg(sp + o0 + wc.ip));
    hs[(1 * UR_PLANES + pl) * UR_XW + xx] = two_term<NM::kUpInner>(wc.l0, __ldg(sp + o1), wc.l1, __ldg(sp + o1 + wc.ip));
  }
  __syncthreads();
  // ---- phase 2: thread = (group of 4 columns, row phase)
  const u64 one2 = pack2(one, one);
  const int ngroups = xw >> 2;
  const int grp = tid % ngroups, rphase = tid / ngroups, rsplit = UR_THREADS / ngroups;
  if (rphase >= rsplit) return;
  const int xx = grp * 4;
  for (int y = y_lo + rphase; y < y_hi; y += rsplit) {
    const UpCoord hc = up_coord<NM>(sh, y, Hin);
    const u64 hl0 = pack2(hc.l0, hc.l0), hl1 = pack2(hc.l1, hc.l1);
    const long long pix = static_cast<long long>(y) * Wout + x0 + xx;
    if (IL5) {
      float v[5][4];
#pragma unroll
      for (int pl = 0; pl < 5; ++pl) {
        const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(hs + (0 * UR_PLANES + pl) * UR_XW + xx);
        const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(hs + (1 * UR_PLANES + pl) * UR_XW + xx);
        unpack2(two_term2<NM::kUpOuter>(hl0, f0.x, hl1, f1.x, one2), v[pl][0], v[pl][1]);
        unpack2(two_term2<NM::kUpOuter>(hl0, f0.y, hl1, f1.y, one2), v[pl][2], v[pl][3]);
      }
      float4* q = reinterpret_cast<float4*>(dst) + pix;
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = make_float4(v[0][i], v[1][i], v[2][i], v[3][i]);
      *reinterpret_cast<float4*>(dst + 4 * out_plane + pix) = make_float4(v[4][0], v[4][1], v[4][2], v[4][3]);
    } else {
      for (int pl = 0; pl < pc; ++pl) {
        const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(hs + (0 * UR_PLANES + pl) * UR_XW + xx);
        const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(hs + (1 * UR_PLANES + pl) * UR_XW + xx);
        float4 v;
        unpack2(two_term2<NM::kUpOuter>(hl0, f0.x, hl1, f1.x, one2), v.x, v.y);
        unpack2(two_term2<NM::kUpOuter>(hl0, f0.y, hl1, f1.y, one2), v.z, v.w);
        *reinterpret_cast<float4*>(dst + (p0 + pl) * out_plane + pix) = v;
      }
    }
  }
}

// eligibility of the staged kernel: at least four output rows per source row (measured: 5 x 135x240 -> 1080p 19.7 -> 13.3
// us, but 2048 x 67x120 -> 135x240, two rows per interval, 106 -> 125 us: the staging is not amortised), 128-bit stores
static bool upsample_rows_ok(long long planes, int Hin, int Win, int Hout, int Wout, const float* dst) {
  if ((Wout & 3) != 0 || !aligned16(dst) || Hin < 2 || Hout < 4 * Hin || Win < 1 || planes < 1) return false;
  if (static_cast<long long>(Hout) * Wout >= (1ll << 31) || static_cast<long long>(Hin) * Win >= (1ll << 31)) return false;
  const long long ctas = static_cast<long long>(Hin) * ((Wout + UR_XW - 1) / UR_XW);
  return ctas <= 0x7fffffffll && (planes + UR_PLANES - 1) / UR_PLANES <= 65535;
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
  if (upsample_rows_ok(planes, Hin, Win, Hout, Wout, dst)) {
    const int nchunks = (Wout + UR_XW - 1) / UR_XW;
    dim3 grid(static_cast<unsigned>(Hin * nchunks), static_cast<unsigned>((planes + UR_PLANES - 1) / UR_PLANES));
    upsample_rows_kernel<NM, false><<<grid, UR_THREADS, 0, st>>>(src, dst, planes, Hin, Win, Hout, Wout, ac_scale(Hin, Hout),
                                                                ac_scale(Win, Wout), nchunks, 1.0f);
    return check_launch("fuvs_upsample_bilinear_ac");
  }
  if ((Wout & 3) == 0 && aligned16(dst) && out_plane < (1ll << 31)) {
    const long long items = out_plane >> 2;
    // few planes per thread column (grid.y), so that small plane counts still fill the SMs
    const long long bx4 = (items + threads - 1) / threads;
    long long by = (16ll * sm_count() + bx4 - 1) / bx4;          // planes per thread = planes / by
    by = by < 1 ? 1 : (by > planes ? planes : by);
    if (by > 65535) by = 65535;
    dim3 grid4(static_cast<unsigned>(bx4), static_cast<unsigned>(by));
    upsample_bilinear_ac_v4_kernel<NM><<<grid4, threads, 0, st>>>(src, dst, planes, Hin, Win, Hout, Wout,
                                                                  ac_scale(Hin, Hout), ac_scale(Win, Wout));
    return check_launch("fuvs_upsample_bilinear_ac");
  }
  const long long bx = (out_plane + threads - 1) / threads;
  if (bx > 0x7fffffffll) return set_error(FUVS_EINVAL, "upsample: output plane too large");
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(planes < 65535 ? planes : 65535));
  upsample_bilinear_ac_kernel<NM><<<grid, threads, 0, st>>>(src, dst, planes, Hin, Win, Hout, Wout,
                                                             ac_scale(Hin, Hout), ac_scale(Win, Wout));
  return check_launch("fuvs_upsample_bilinear_ac");
}

// Up-sample of a 5-class key frame into the dense strip kernel's "4+1" state layout (dense_strip.cu `layout`): channels
// 0-3 interleaved per pixel ([H][W][4]) followed by the plane of channel 4.  One thread = 4 consecutive pixels of all
// five classes: four 128-bit stores into the interleaved part, one into the plane.  Same up_fetch<> arithmetic as
// fuvs_upsample_bilinear_ac.
template <class NM>
__global__ void __launch_bounds__(256)
upsample_bilinear_ac_il5_kernel(const float* __restrict__ src, float* __restrict__ dst, int Hin, int Win, int Hout,
                                int Wout, float sh, float sw) {
  const int groups = Wout >> 2;
  const long long items = static_cast<long long>(Hout) * groups;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= items) return;
  const int y = static_cast<int>(t / groups), x = static_cast<int>(t - static_cast<long long>(y) * groups) << 2;
  const UpCoord hc = up_coord<NM>(sh, y, Hin);
  UpCoord wc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wc[i] = up_coord<NM>(sw, x + i, Win);
  const long long in_plane = static_cast<long long>(Hin) * Win, out_plane = static_cast<long long>(Hout) * Wout;
  const long long pix = static_cast<long long>(y) * Wout + x;
  float v[5][4];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[c][i] = up_fetch<NM>(src + c * in_plane, Win, hc, wc[i]);
  }
  float4* q = reinterpret_cast<float4*>(dst) + pix;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float4(v[0][i], v[1][i], v[2][i], v[3][i]);
  *reinterpret_cast<float4*>(dst + 4 * out_plane + pix) = make_float4(v[4][0], v[4][1], v[4][2], v[4][3]);
}

// Key frame [C,hl,wl] -> [C,H,W] (F.interpolate bilinear, align_corners=True; flow/model.py:191-193), planar or 4+1.
int launch_upsample_keyframe(const float* src, float* dst, int C, int hl, int wl, int H, int W, bool il, cudaStream_t st) {
  if (!il) return launch_upsample<Nm>(src, dst, C, hl, wl, H, W, st);
  if (C != 5 || (W & 3) != 0 || !aligned16(dst) || static_cast<long long>(H) * W >= (1ll << 31))
    return set_error(FUVS_EINVAL, "up-sample into the 4+1 layout needs C = 5, W %% 4 == 0 and a 16-byte aligned destination");
  if (upsample_rows_ok(5, hl, wl, H, W, dst)) {
    const int nchunks = (W + UR_XW - 1) / UR_XW;
    upsample_rows_kernel<Nm, true><<<dim3(static_cast<unsigned>(hl * nchunks), 1), UR_THREADS, 0, st>>>(
        src, dst, 5, hl, wl, H, W, ac_scale(hl, H), ac_scale(wl, W), nchunks, 1.0f);
    return check_launch("fuvs_dense_lowres_interval(up-sample)");
  }
  const long long items = (static_cast<long long>(H) * W) >> 2;
  const long long bx = (items + 255) / 256;
  upsample_bilinear_ac_il5_kernel<Nm><<<static_cast<unsigned>(bx), 256, 0, st>>>(src, dst, hl, wl, H, W, ac_scale(hl, H), ac_scale(wl, W));
  return check_launch("fuvs_dense_lowres_interval(up-sample)");
}

// ---------------------------------------------------------------------------
// block-grid chain step (low resolution): one launch advances both chains of up to CHAIN_MAX_IV intervals by one step
// (blockIdx.z = 2 * interval + side).  A clip's intervals are independent until their label maps meet in the
// temporal counts, so fuvs_block_clip pays the n-1 dependent launch latencies once per clip instead of per interval.
// ---------------------------------------------------------------------------
constexpr int CHAIN_MAX_IV = 8;
struct ChainStepBatch {
  const float* src[2 * CHAIN_MAX_IV];
  const float* grid[2 * CHAIN_MAX_IV];
  float* dst[2 * CHAIN_MAX_IV];
};

template <class NM>
__device__ __forceinline__ void gs_fetch_all_cg(const float* src, long long in_plane, const GsTap& t, int Win, int C,
                                                float* dst, int out_plane);

template <class NM>
__global__ void __launch_bounds__(256)
block_chain_step_kernel(const __grid_constant__ ChainStepBatch B, int C, int Hin, int Win, int Hg, int Wg) {
  // programmatic dependent launch (when the launch carries the attribute): nothing is read before the wait — the flow
  // vectors of a later interval may be written by a kernel that is still running
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= Wg || y >= Hg) return;
  const int z = blockIdx.z;
  const int opix = y * Wg + x;
  const float2 g = __ldg(reinterpret_cast<const float2*>(B.grid[z]) + opix);
  const GsTap t = gs_setup<NM>(g.x, g.y, Hin, Win, false);
  gs_fetch_all_cg<NM>(B.src[z], static_cast<long long>(Hin) * Win, t, Win, C, B.dst[z] + opix, Hg * Wg);
}

// Step 1 of the chains with the key frames at DECODER resolution [C,hl,wl]: the reference samples
// up(key) = F.interpolate(key, (H,W), bilinear, align_corners=True) (flow/model.py:191-193, 205-206, 214); here each of
// the four taps of a grid point is that up-sample evaluated at the tap's pixel (up_fetch: the arithmetic of
// fuvs_upsample_bilinear_ac), so the full-resolution key frame is never materialised.  16 loads per point and channel,
// 8 040 points per side.
template <class NM>
__global__ void __launch_bounds__(256)
block_chain_step1_lowres_kernel(const __grid_constant__ ChainStepBatch B, int C, int hl, int wl, int H, int W, int Hg,
                                int Wg, float shk, float swk) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= Wg || y >= Hg) return;
  const int z = blockIdx.z;
  const int opix = y * Wg + x;
  const float2 g = __ldg(reinterpret_cast<const float2*>(B.grid[z]) + opix);
  const GsTap t = gs_setup<NM>(g.x, g.y, H, W, false);
  const UpCoord hn = up_coord<NM>(shk, t.iy, hl), hs = up_coord<NM>(shk, t.iy + t.dy, hl);
  const UpCoord ww = up_coord<NM>(swk, t.ix, wl), we = up_coord<NM>(swk, t.ix + t.dx, wl);
  const int lplane = hl * wl, out_plane = Hg * Wg;
  const float* src = B.src[z];
  float* dst = B.dst[z] + opix;
  for (int c = 0; c < C; ++c) {
    const float* pl = src + c * lplane;
    float acc = tap_acc<NM>(0.f, up_fetch<NM>(pl, wl, hn, ww), t.nw);
    if (t.dx) acc = tap_acc<NM>(acc, up_fetch<NM>(pl, wl, hn, we), t.ne);
    if (t.dy) acc = tap_acc<NM>(acc, up_fetch<NM>(pl, wl, hs, ww), t.sw);
    if (t.dx & t.dy) acc = tap_acc<NM>(acc, up_fetch<NM>(pl, wl, hs, we), t.se);
    dst[c * out_plane] = acc;
  }
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

// ld.global.cg with a 64-byte L2 fetch granularity: the taps of a grid point are scattered 4-byte reads (step 1: from
// the cold full-resolution key frames), a 128-byte fill per tap row would move 52 MB for 1.3 MB of useful data
__device__ __forceinline__ float ld_cg_64(const float* p) {
  float v;
  asm volatile("ld.global.cg.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// All C channels of one grid point: the tap loads of up to 8 channels are issued before the first store.  Written as
// `dst[c] = gs_fetch_cg(src + c)` the store of channel c — which may alias the source for all the compiler knows (both
// live in the chain scratch) — keeps the loads of channel c+1 behind it in program order: C dependent L2 round
// trips per point and step instead of one.
template <class NM>
__device__ __forceinline__ void gs_fetch_all_cg(const float* src, long long in_plane, const GsTap& t, int Win, int C,
                                                float* dst, int out_plane) {
  const float* p = src + t.off00;
  const int o01 = t.dx, o10 = t.dy * Win, o11 = t.dy * Win + t.dx;
  constexpr int U = 8;
  for (int c0 = 0; c0 < C; c0 += U) {
    float v00[U], v01[U], v10[U], v11[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u < C) {
        const float* q = p + (c0 + u) * in_plane;
        v00[u] = ld_cg_64(q); v01[u] = ld_cg_64(q + o01); v10[u] = ld_cg_64(q + o10); v11[u] = ld_cg_64(q + o11);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u < C) {
        float acc = 0.f;
        acc = tap_acc<NM>(acc, v00[u], t.nw);
        if (t.dx) acc = tap_acc<NM>(acc, v01[u], t.ne);
        if (t.dy) acc = tap_acc<NM>(acc, v10[u], t.sw);
        if (t.dx & t.dy) acc = tap_acc<NM>(acc, v11[u], t.se);
        dst[(c0 + u) * out_plane] = acc;
      }
    }
  }
}

// Whole chain of ONE interval as a cooperative launch over the GPU: one thread per (side, grid point), grid.sync()
// between the dependent steps (eager launches: each extra launch costs the host ~3 us).  Spreading the 2 x 8 040
// points over ~63 CTAs keeps the scattered tap loads off a handful of L1s.
struct ChainGrids {
  const float* left[FUVS_MAX_FRAMES];
  const float* right[FUVS_MAX_FRAMES];
};

template <class NM>
__global__ void __launch_bounds__(256)
block_chain_coop_kernel(const float* __restrict__ prev, const float* __restrict__ next,
                        const __grid_constant__ ChainGrids G, float* Lst, float* Rst, int C, int H, int W, int Hg,
                        int Wg, int n) {
  namespace cg = cooperative_groups;
  cg::grid_group gridg = cg::this_grid();
  const int npts = Hg * Wg;
  const long long ls = static_cast<long long>(C) * npts;
  const int total = 2 * npts;
  const int nthreads = gridDim.x * blockDim.x;
  for (int j = 1; j <= n - 1; ++j) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += nthreads) {
      const int side = t >= npts;
      const int pt = t - side * npts;
      float* st = side ? Rst : Lst;
      const float* src = (j == 1) ? (side ? next : prev) : st + (j - 2) * ls;
      const int Hin = (j == 1) ? H : Hg, Win = (j == 1) ? W : Wg;
      const long long in_plane = static_cast<long long>(Hin) * Win;
      const float* grid = side ? G.right[j - 1] : G.left[j - 1];
      float* dst = st + (j - 1) * ls;
      const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + pt);
      const GsTap tp = gs_setup<NM>(g.x, g.y, Hin, Win, false);
      gs_fetch_all_cg<NM>(src, in_plane, tp, Win, C, dst + pt, npts);
    }
    if (j < n - 1) gridg.sync();
  }
}

// ---------------------------------------------------------------------------
// Streaming kernel, column strips.  One thread = one output column x ROWS consecutive rows of a 128 x ROWS tile.
// upsample_bilinear2d is separable in the order ATen evaluates it:
//   val = h0 * (w0*a + w1*b) + h1 * (w0*c + w1*d)
// the two inner row interpolations depend on (x, source row) only, and 16 output rows share a source row at the
// reference's 16x ratio, so they are computed once per (map, class) and reused down the strip: ~20 tap loads per
// pixel instead of 160, same fp32 operations in the same order (bit-identical to the per-pixel evaluation).
// Temporal-consistency counts are fused (CT > 0): the previous frame's labels of the strip stay in registers.
// ---------------------------------------------------------------------------
constexpr int SROWS = 8;
constexpr int STHREADS = 128;

// torch.max semantics in 4 instructions: !(v <= best) is true for v > best or any NaN; a NaN best is never replaced.
__device__ __forceinline__ void argmax_push(float v, int c, float& best, int& idx) {
  const bool take = !(v <= best) && (best == best);
  best = take ? v : best;
  idx = take ? c : idx;
}

// LOGITS: also write the blended logits; FULLROWS: every tile has SROWS valid rows (H % SROWS == 0)
template <class NM, int CT, bool COUNTS, bool LOGITS, bool FULLROWS>
__global__ void __launch_bounds__(STHREADS)
block_stream_cols_kernel(const float* __restrict__ key0, const float* __restrict__ Lst, const float* __restrict__ Rst,
                         int Crt, int H, int W, int Hg, int Wg, int n, float sh, float sw,
                         uint8_t* __restrict__ labels, float* __restrict__ logits,
                         const uint8_t* __restrict__ tc_prev, unsigned long long* __restrict__ counts,
                         int ignore_index, const BlendWeights wts) {
  constexpr int KC = CT > 0 ? CT : 1;
  __shared__ unsigned shc[24];
  const int C = CT > 0 ? CT : Crt;
  const int x = blockIdx.x * STHREADS + threadIdx.x;
  const int y0 = blockIdx.y * SROWS;
  const bool active = x < W;
  const int xc = active ? x : W - 1;                       // inactive lanes compute on a valid column, store nothing
  const long long HW = static_cast<long long>(H) * W;
  const int lp = Hg * Wg;
  const long long ls = static_cast<long long>(C) * lp;
  const int rows = FULLROWS ? SROWS : min(SROWS, H - y0);
  FieldCounts<KC> cnt;
  if (COUNTS) cnt.init();

  const UpCoord wc = up_coord<NM>(sw, xc, Wg);
  int hi0[SROWS];
  float hl0[SROWS], hl1[SROWS];
#pragma unroll
  for (int r = 0; r < SROWS; ++r) {
    const UpCoord hc = up_coord<NM>(sh, min(y0 + r, H - 1), Hg);
    hi0[r] = hc.i0; hl0[r] = hc.l0; hl1[r] = hc.l1;
  }
  int last[SROWS];
  bool have_last = false;

  // ---- frame 0: the key frame itself
  {
    float best[SROWS];
    int idx[SROWS];
#pragma unroll
    for (int r = 0; r < SROWS; ++r) { best[r] = -INFINITY; idx[r] = 0; }
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int r = 0; r < SROWS; ++r) {
        if (FULLROWS || r < rows) {
          const long long o = c * HW + static_cast<long long>(y0 + r) * W + xc;
          const float v = __ldcs(key0 + o);
          if (LOGITS && active) __stcs(logits + o, v);
          argmax_push(v, c, best[r], idx[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SROWS; ++r) {
      if (FULLROWS || r < rows) {
        const long long pix = static_cast<long long>(y0 + r) * W + xc;
        if (labels && active) labels[pix] = static_cast<uint8_t>(idx[r]);
        if (COUNTS) {
          if (tc_prev != nullptr && active) {
            const int t = __ldg(tc_prev + pix);
            const unsigned ft = (t < KC) ? FieldCounts<KC>::field(t) : 0u;
            const unsigned fo = (t == ignore_index) ? 0u : FieldCounts<KC>::field(idx[r]);
            cnt.add(idx[r], fo, t, ft);
          }
          last[r] = idx[r];
        }
      }
    }
    have_last = true;
  }
  (void)have_last;

  // ---- frames 1..n-1
  int since_spill = 1;
  for (int p = 1; p < n; ++p) {
    const float w0 = wts.w0[p], w1 = wts.w1[p];
    const float* Lp = Lst + (p - 1) * ls + wc.i0;
    const float* Rp = Rst + (n - p - 1) * ls + wc.i0;
    float best[SROWS];
    int idx[SROWS];
#pragma unroll
    for (int r = 0; r < SROWS; ++r) { best[r] = -INFINITY; idx[r] = 0; }
    for (int c = 0; c < C; ++c) {
      const float* l = Lp + c * lp;
      const float* rr = Rp + c * lp;
      int cached = -1;
      float f0 = 0.f, f1 = 0.f, b0 = 0.f, b1 = 0.f;        // inner (row) interpolations of the two source rows
#pragma unroll
      for (int r = 0; r < SROWS; ++r) {
        if (FULLROWS || r < rows) {
          if (hi0[r] != cached) {                            // block-uniform: all threads share the rows
            cached = hi0[r];
            const int o0 = cached * Wg, o1 = (cached + ((cached < Hg - 1) ? 1 : 0)) * Wg;
            f0 = two_term<NM::kUpInner>(wc.l0, __ldg(l + o0), wc.l1, __ldg(l + o0 + wc.ip));
            f1 = two_term<NM::kUpInner>(wc.l0, __ldg(l + o1), wc.l1, __ldg(l + o1 + wc.ip));
            b0 = two_term<NM::kUpInner>(wc.l0, __ldg(rr + o0), wc.l1, __ldg(rr + o0 + wc.ip));
            b1 = two_term<NM::kUpInner>(wc.l0, __ldg(rr + o1), wc.l1, __ldg(rr + o1 + wc.ip));
          }
          const float f = two_term<NM::kUpOuter>(hl0[r], f0, hl1[r], f1);
          const float b = two_term<NM::kUpOuter>(hl0[r], b0, hl1[r], b1);
          const float v = blend2(w0, f, w1, b);
          if (LOGITS && active)
            __stcs(logits + (static_cast<long long>(p) * C + c) * HW + static_cast<long long>(y0 + r) * W + xc, v);
          argmax_push(v, c, best[r], idx[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SROWS; ++r) {
      if (FULLROWS || r < rows) {
        if (labels && active) labels[p * HW + static_cast<long long>(y0 + r) * W + xc] = static_cast<uint8_t>(idx[r]);
        if (COUNTS) {
          if (active) cnt.add(idx[r], FieldCounts<KC>::field(idx[r]), last[r], FieldCounts<KC>::field(last[r]));
          last[r] = idx[r];
        }
      }
    }
    if (COUNTS && ++since_spill >= FieldCfg<KC>::CAP / SROWS) {
      cnt.spill();
      since_spill = 0;
    }
  }
  if (COUNTS) cnt.finish(shc, counts, KC);
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

namespace fuvs {

// Frames of one interval from its finished chain states: up-sample + blend + arg-max (+ counts when the row kernel
// takes the shape; otherwise the caller runs fuvs_temporal_counts).  Returns 0 / negative; *counts_done says whether
// the temporal counts were fused.
static int block_frames(const float* prev, const float* Lst, const float* Rst, int C, int H, int W, int Hg, int Wg,
                        int n, uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts,
                        int ignore_index, cudaStream_t st, bool* counts_done, int hl, int wl) {
  *counts_done = false;
  if (!labels && !logits) return FUVS_OK;
  BlendWeights w;
  make_blend_weights(n, &w);
  const float sh = ac_scale(Hg, H), sw = ac_scale(Wg, W);
  if (hl > 0) {      // key frame at decoder resolution: only the row kernel evaluates its up-sample on the fly
    const int r = launch_block_stream_rows(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts,
                                           ignore_index, w, st, hl, wl, counts_done);
    if (r < 0) return r;
    if (r > 0) return set_error(FUVS_EINVAL, "block_lowres: shape not supported (see fuvs_block_lowres_supported)");
    return FUVS_OK;
  }
  if (Hg == H && Wg == W) {
    // sizes already match: the reference skips the interpolate call (flow/model.py:217), per-pixel kernel
    dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
    if (grid.y > 65535) return set_error(FUVS_EINVAL, "block: H=%d too large", H);
    switch (C) {
      case 2: block_stream_kernel<Nm, 2><<<grid, block, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, w); break;
      case 5: block_stream_kernel<Nm, 5><<<grid, block, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, w); break;
      default: block_stream_kernel<Nm, 0><<<grid, block, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, w); break;
    }
    return check_launch("fuvs_block_interval(stream)");
  }
  // source-row intervals (block_rows.cu) for the shapes it takes (2 <= C <= 5, W % 4 == 0), column strips otherwise
  const int r = launch_block_stream_rows(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts,
                                         ignore_index, w, st, 0, 0, counts_done);
  if (r < 0) return r;
  if (r == 0) return FUVS_OK;                   // (the row kernel says whether it took the counts)
  *counts_done = false;
  dim3 grid((W + STHREADS - 1) / STHREADS, (H + SROWS - 1) / SROWS);
  if (grid.y > 65535) return set_error(FUVS_EINVAL, "block: H=%d too large", H);
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  // counts are fused when the class count has a field-packed counter and ignore_index cannot collide with a class
  const bool fuse = counts && labels && C <= 5 && (ignore_index < 0 || ignore_index >= C) && n * SROWS <= 4096;
#define FUVS_COLS2(CT_, CNT_, LG_, FR_)                                                                               \
  block_stream_cols_kernel<Nm, CT_, CNT_, LG_, FR_><<<grid, STHREADS, 0, st>>>(prev, Lst, Rst, C, H, W, Hg, Wg, n, sh, sw, \
                                                                               labels, logits, tc_prev, cu, ignore_index, w)
#define FUVS_COLS(CT_, CNT_)                                            \
  do {                                                                  \
    if (logits) FUVS_COLS2(CT_, CNT_, true, false);                     \
    else if (H % SROWS == 0) FUVS_COLS2(CT_, CNT_, false, true);        \
    else FUVS_COLS2(CT_, CNT_, false, false);                           \
  } while (0)
  if (fuse) {
    switch (C) {
      case 1: FUVS_COLS(1, true); break;
      case 2: FUVS_COLS(2, true); break;
      case 3: FUVS_COLS(3, true); break;
      case 4: FUVS_COLS(4, true); break;
      default: FUVS_COLS(5, true); break;
    }
  } else {
    FUVS_COLS(0, false);
  }
#undef FUVS_COLS
#undef FUVS_COLS2
  if (int e = check_launch("fuvs_block_interval(stream cols)")) return e;
  *counts_done = fuse;
  return FUVS_OK;
}

// m consecutive intervals of one clip: keys[i], keys[i+1] bracket interval i; gl / gr hold m * (n-1) grid pointers.
static int block_clip_impl(int m, const float* const* keys, const float* const* gl, const float* const* gr, int C,
                           int H, int W, int Hg, int Wg, int n, float* scratch, uint8_t* const* labels,
                           float* const* logits, const uint8_t* tc_prev, long long* counts, int ignore_index,
                           cudaStream_t st, const char* who, int hl = 0, int wl = 0) {
  // hl > 0: the key frames are given at decoder resolution [C,hl,wl] (n > 1 only; the caller has checked the shape)
  if (int e = device_ok()) return e;
  if (m < 1 || !keys || C < 1 || H < 1 || W < 1 || n < 1)
    return set_error(FUVS_EINVAL, "%s: bad shape m=%d C=%d H=%d W=%d n=%d", who, m, C, H, W, n);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "%s: n=%d exceeds %d frames per interval", who, n, FUVS_MAX_FRAMES);
  if (n > 1 && (!gl || !gr || !scratch || Hg < 1 || Wg < 1))
    return set_error(FUVS_EINVAL, "%s: grids/scratch are NULL or grid is empty (Hg=%d Wg=%d) but n=%d", who, Hg, Wg, n);
  const bool want_labels = labels != nullptr && labels[0] != nullptr;
  const bool want_logits = logits != nullptr && logits[0] != nullptr;
  if ((want_labels || counts) && C > 256) return set_error(FUVS_EINVAL, "%s: uint8 label maps need C <= 256 (C=%d)", who, C);
  if (counts && !want_labels) return set_error(FUVS_EINVAL, "%s: counts need the label maps (labels is NULL)", who);
  const long long HW = static_cast<long long>(H) * W;
  if (HW >= (1ll << 31) || static_cast<long long>(C) * Hg * Wg * 2 * n >= (1ll << 31))
    return set_error(FUVS_EINVAL, "%s: problem exceeds 32-bit indexing", who);
  for (int i = 0; i <= (n > 1 ? m : m - 1); ++i)
    if (!keys[i]) return set_error(FUVS_EINVAL, "%s: key frame %d is NULL", who, i);
  for (int i = 0; i < m; ++i) {
    if ((want_labels && !labels[i]) || (want_logits && !logits[i]))
      return set_error(FUVS_EINVAL, "%s: output pointer of interval %d is NULL", who, i);
  }
  for (int i = 0; n > 1 && i < m * (n - 1); ++i) {
    if (!gl[i] || !gr[i]) return set_error(FUVS_EINVAL, "%s: grid %d is NULL", who, i);
    if (!aligned8(gl[i]) || !aligned8(gr[i])) return set_error(FUVS_EALIGN, "%s: grids must be 8-byte aligned", who);
  }
  const long long S = static_cast<long long>(C) * HW;
  const long long ls = static_cast<long long>(C) * Hg * Wg;          // one low-resolution state
  const long long per_iv = 2ll * (n - 1) * ls;                       // scratch of one interval: L_1..L_{n-1}, R_1..R_{n-1}

  if (n == 1) {
    for (int i = 0; i < m; ++i) {
      if (want_labels) {
        if (int e = launch_argmax(keys[i], 1, C, HW, labels[i], nullptr, st)) return e;
      }
      if (want_logits) {
        if (cudaMemcpyAsync(logits[i], keys[i], S * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
          return set_error(FUVS_ECUDA, "%s: logits copy failed", who);
      }
      if (counts) {
        const uint8_t* tp = (i == 0) ? tc_prev : labels[i - 1];
        if (int e = launch_temporal_counts(labels[i], 1, HW, tp, C, ignore_index, counts, st)) return e;
      }
    }
    return FUVS_OK;
  }

  // ---- chains.  One interval, eager launch: one cooperative launch with grid.sync between the steps.  While the
  // stream is being captured (kernel-to-kernel latency inside a graph is below a grid-wide barrier: 62.7 vs 65.8 us
  // per 1080p interval) or for several intervals: n-1 launches, each advancing every chain of the batch by one step.
  bool chains_done = false;
  if (m == 1 && hl == 0) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    if (cap != cudaStreamCaptureStatusActive) {
      const int total = 2 * Hg * Wg;
      int cgrid = (total + 255) / 256;
      static int bps = blocks_per_sm(block_chain_coop_kernel<Nm>, 256);
      const int capacity = sm_count() * bps;
      if (cgrid > capacity) cgrid = capacity;
      ChainGrids cgr;
      for (int j = 0; j < n - 1; ++j) { cgr.left[j] = gl[j]; cgr.right[j] = gr[j]; }
      const float* prev = keys[0];
      const float* next = keys[1];
      float* Lst = scratch;
      float* Rst = scratch + (n - 1) * ls;
      void* args[] = {(void*)&prev, (void*)&next, (void*)&cgr, (void*)&Lst, (void*)&Rst,
                      (void*)&C, (void*)&H, (void*)&W, (void*)&Hg, (void*)&Wg, (void*)&n};
      if (cudaLaunchCooperativeKernel((const void*)block_chain_coop_kernel<Nm>, dim3(cgrid), dim3(256), args, 0, st) ==
          cudaSuccess) {
        count_launch();
        chains_done = true;
      } else {
        cudaGetLastError();   // cooperative launch unavailable: per-step launches below
      }
    }
  }
  if (!chains_done) {
    for (int i0 = 0; i0 < m; i0 += CHAIN_MAX_IV) {
      const int mb = (m - i0 < CHAIN_MAX_IV) ? m - i0 : CHAIN_MAX_IV;
      for (int j = 1; j <= n - 1; ++j) {
        ChainStepBatch b;
        for (int i = 0; i < mb; ++i) {
          float* Lst = scratch + (i0 + i) * per_iv;
          float* Rst = Lst + (n - 1) * ls;
          b.src[2 * i] = (j == 1) ? keys[i0 + i] : Lst + (j - 2) * ls;
          b.src[2 * i + 1] = (j == 1) ? keys[i0 + i + 1] : Rst + (j - 2) * ls;
          b.grid[2 * i] = gl[(i0 + i) * (n - 1) + j - 1];
          b.grid[2 * i + 1] = gr[(i0 + i) * (n - 1) + j - 1];
          b.dst[2 * i] = Lst + (j - 1) * ls;
          b.dst[2 * i + 1] = Rst + (j - 1) * ls;
        }
        const int Hin = (j == 1) ? H : Hg, Win = (j == 1) ? W : Wg;
        dim3 cgrid((Wg + 31) / 32, (Hg + 7) / 8, 2 * mb), cblock(32, 8);
        // plain launches on purpose: with programmatic stream serialization the early-launched CTAs of the later steps
        // and of the frame kernel sit on the SMs the running step needs (r01: 97 vs 63 us per 1080p interval)
        if (j == 1 && hl > 0)
          block_chain_step1_lowres_kernel<Nm><<<cgrid, cblock, 0, st>>>(b, C, hl, wl, H, W, Hg, Wg, ac_scale(hl, H), ac_scale(wl, W));
        else
          block_chain_step_kernel<Nm><<<cgrid, cblock, 0, st>>>(b, C, Hin, Win, Hg, Wg);
        if (int e = check_launch("fuvs_block_interval(chain)")) return e;
      }
    }
  }
  // ---- frames of every interval
  for (int i = 0; i < m; ++i) {
    float* Lst = scratch + i * per_iv;
    float* Rst = Lst + (n - 1) * ls;
    uint8_t* lab = want_labels ? labels[i] : nullptr;
    const uint8_t* tp = (i == 0) ? tc_prev : (want_labels ? labels[i - 1] + (n - 1) * HW : nullptr);
    bool counts_done = false;
    if (int e = block_frames(keys[i], Lst, Rst, C, H, W, Hg, Wg, n, lab, want_logits ? logits[i] : nullptr, tp, counts,
                             ignore_index, st, &counts_done, hl, wl))
      return e;
    if (counts && !counts_done) {
      if (int e = launch_temporal_counts(lab, n, HW, tp, C, ignore_index, counts, st)) return e;
    }
  }
  return FUVS_OK;
}

}  // namespace fuvs

extern "C" int fuvs_block_interval_ptrs(const float* prev, const float* next, const float* const* grids_left,
                                        const float* const* grids_right, int C, int H, int W, int Hg, int Wg, int n,
                                        float* scratch, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                                        long long* counts, int ignore_index, fuvs_stream_t stream) {
  const float* keys[2] = {prev, next};
  uint8_t* labs[1] = {labels};
  float* logs[1] = {logits};
  if (n > 1 && !next) return fuvs::set_error(FUVS_EINVAL, "block: next key frame is NULL but n=%d", n);
  return fuvs::block_clip_impl(1, keys, grids_left, grids_right, C, H, W, Hg, Wg, n, scratch, labs, logs, tc_prev, counts,
                               ignore_index, static_cast<cudaStream_t>(stream), "block");
}

extern "C" int fuvs_block_interval(const float* prev, const float* next, const float* grids_left,
                                   const float* grids_right, int C, int H, int W, int Hg, int Wg, int n,
                                   float* scratch, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                                   long long* counts, int ignore_index, fuvs_stream_t stream) {
  const float* gl[FUVS_MAX_FRAMES];
  const float* gr[FUVS_MAX_FRAMES];
  if (n > FUVS_MAX_FRAMES) return fuvs::set_error(FUVS_EINVAL, "block: n=%d exceeds %d frames per interval", n, FUVS_MAX_FRAMES);
  if (n > 1 && (!grids_left || !grids_right)) return fuvs::set_error(FUVS_EINVAL, "block: grids are NULL but n=%d", n);
  const long long g = static_cast<long long>(Hg) * Wg * 2;
  for (int j = 0; j < n - 1; ++j) {
    gl[j] = grids_left + j * g;
    gr[j] = grids_right + j * g;
  }
  return fuvs_block_interval_ptrs(prev, next, gl, gr, C, H, W, Hg, Wg, n, scratch, labels, logits, tc_prev, counts,
                                  ignore_index, stream);
}

extern "C" int fuvs_block_lowres_supported(int C, int hl, int wl, int H, int W, int Hg, int Wg) {
  if (C < 2 || C > 5 || (W & 3) != 0 || hl < 1 || wl < 1 || H < 1 || W < 1 || Hg < 1 || Wg < 1) return 0;
  if (H == Hg && W == Wg) return 0;
  // the row kernel stages at most 4 low-resolution key-frame rows per source-row interval of the grid
  const float sh = fuvs::ac_scale(Hg, H), shk = fuvs::ac_scale(hl, H);
  int y = 0;
  while (y < H) {
    const int i0 = static_cast<int>(sh * static_cast<float>(y));
    int y_hi = y;
    while (y_hi < H && static_cast<int>(sh * static_cast<float>(y_hi)) == i0) ++y_hi;
    const int first = static_cast<int>(shk * static_cast<float>(y)), last = static_cast<int>(shk * static_cast<float>(y_hi - 1));
    if (last + ((last < hl - 1) ? 1 : 0) - first + 1 > 4) return 0;
    y = y_hi;
  }
  return 1;
}

extern "C" int fuvs_block_lowres_interval_ptrs(const float* prev_lr, const float* next_lr, int hl, int wl,
                                               const float* const* grids_left, const float* const* grids_right, int C,
                                               int H, int W, int Hg, int Wg, int n, float* scratch, uint8_t* labels,
                                               float* logits, const uint8_t* tc_prev, long long* counts,
                                               int ignore_index, fuvs_stream_t stream) {
  using namespace fuvs;
  if (hl == H && wl == W)     // the reference skips the interpolate when the sizes match (flow/model.py:191)
    return fuvs_block_interval_ptrs(prev_lr, next_lr, grids_left, grids_right, C, H, W, Hg, Wg, n, scratch, labels, logits,
                                    tc_prev, counts, ignore_index, stream);
  if (n < 2 || !next_lr) return set_error(FUVS_EINVAL, "block_lowres: needs n >= 2 and both key frames (n=%d)", n);
  if (!fuvs_block_lowres_supported(C, hl, wl, H, W, Hg, Wg))
    return set_error(FUVS_EINVAL, "block_lowres: needs 2 <= C <= 5, W %% 4 == 0, a grid coarser than the frame and a key "
                     "frame no finer than ~4 source rows per grid row (C=%d %dx%d -> %dx%d, grid %dx%d); up-sample with "
                     "fuvs_upsample_bilinear_ac and call fuvs_block_interval instead", C, hl, wl, H, W, Hg, Wg);
  if (counts && ignore_index >= 0 && ignore_index < C)
    return set_error(FUVS_EINVAL, "block_lowres: ignore_index=%d collides with a class", ignore_index);
  if ((logits && !aligned16(logits)) || (labels && !aligned4(labels)) || (tc_prev && !aligned4(tc_prev)))
    return set_error(FUVS_EALIGN, "block_lowres: logits must be 16-byte, label maps 4-byte aligned");
  if (counts && !labels) return set_error(FUVS_EINVAL, "block_lowres: counts need the label maps (labels is NULL)");
  const float* keys[2] = {prev_lr, next_lr};
  uint8_t* labs[1] = {labels};
  float* logs[1] = {logits};
  return block_clip_impl(1, keys, grids_left, grids_right, C, H, W, Hg, Wg, n, scratch, labs, logs, tc_prev, counts,
                         ignore_index, static_cast<cudaStream_t>(stream), "block_lowres", hl, wl);
}

extern "C" int fuvs_block_clip(int m, const float* const* keys, const float* const* grids_left,
                               const float* const* grids_right, int C, int H, int W, int Hg, int Wg, int n,
                               float* scratch, uint8_t* const* labels, float* const* logits, const uint8_t* tc_prev,
                               long long* counts, int ignore_index, fuvs_stream_t stream) {
  return fuvs::block_clip_impl(m, keys, grids_left, grids_right, C, H, W, Hg, Wg, n, scratch, labels, logits, tc_prev,
                               counts, ignore_index, static_cast<cudaStream_t>(stream), "block_clip");
}
