// feature.cu — feature-level interpolation of one interval (model.feature_based=True), SURVEY.md §8f rank 2.
//
// fuvs_feature_interval <- FlowModel.predict_feature, flow/model.py:131-173: the two warp chains over encoder
// features, the per-step up-sample back to the feature size (align_corners=True, :138-139, :149-150), the temporal
// blend (:166-171), the key frame's own pass through the default grid with align_corners=True (:154-159, sic) and
// the torch.cat into the decoder batch (:173).  The reference materialises every up-sampled state
// (2(n-1) x [C,fh,fw], 265 MB each for DeepLabv3 features at 1080p) and then reads them back for the blend; here the
// chain states stay at grid resolution and one kernel writes each blended frame straight into its slot of the decoder
// batch: up-sample of both states + blend per output element, 4 consecutive pixels per thread.
#include <cstdlib>

#include "fuvs_common.cuh"

namespace fuvs {

int launch_warp_step_nm(const float* src0, const float* grid0, float* dst0, const float* src1, const float* grid1,
                        float* dst1, int C, int Hin, int Win, int Hg, int Wg, int align_corners, cudaStream_t st);

namespace {

// out[p][c][y][x..x+VEC) = fl(fl(w0p * up(L_p)) + fl(w1p * up(R_{n-p}))) for p = 1..n-1
template <int VEC>
__global__ void __launch_bounds__(256)
feature_up_blend_kernel(const float* __restrict__ Lst, const float* __restrict__ Rst, float* __restrict__ out, int C,
                        int fh, int fw, int Hg, int Wg, int n, float sh, float sw, const BlendWeights wts) {
  const int groups = fw / VEC;
  const long long items = static_cast<long long>(C) * fh * groups;
  const int lplane = Hg * Wg;
  const long long ls = static_cast<long long>(C) * lplane;
  const long long oplane = static_cast<long long>(fh) * fw;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < items;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(t % groups);
    const long long r = t / groups;
    const int y = static_cast<int>(r % fh), c = static_cast<int>(r / fh);
    const UpCoord hc = up_coord<Nm>(sh, y, Hg);
    UpCoord wc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) wc[i] = up_coord<Nm>(sw, g * VEC + i, Wg);
    for (int p = 1; p < n; ++p) {
      const float* L = Lst + (p - 1) * ls + static_cast<long long>(c) * lplane;
      const float* R = Rst + (n - p - 1) * ls + static_cast<long long>(c) * lplane;
      FVec<VEC> o;
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        o.v[i] = blend2(wts.w0[p], up_fetch<Nm>(L, Wg, hc, wc[i]), wts.w1[p], up_fetch<Nm>(R, Wg, hc, wc[i]));
      o.store_stream(out + (static_cast<long long>(p) * C + c) * oplane + static_cast<long long>(y) * fw + g * VEC);
    }
  }
}

// Staged variant (fw % 4 == 0): the kernel above issues 8 gather loads per output element (4 taps of each state) and
// is L1-bound — 1.4 ms for the 1.06 GB of frames 1..n-1 at DeepLabv3 feature size (tools/abi_bench.py).  Here a CTA
// owns (channel, frame, band of source rows): the HORIZONTAL two-terms of the band's source rows of L_p and R_{n-p}
// are computed once into shared memory (the op order of UpSample.cuh is horizontal first), and every output row of the
// band is then one vertical two-term per state on packed FP32x2 straight from 128-bit shared-memory loads, the blend
// and one streaming 128-bit store — like block_rows.cu, without the arg-max.
constexpr int FR_THREADS = 256;

__global__ void __launch_bounds__(FR_THREADS, 4)
feature_rows_kernel(const float* __restrict__ Lst, const float* __restrict__ Rst, float* __restrict__ out, int C, int fh,
                    int fw, int Hg, int Wg, int n, float sh, float sw, int band_rows, int nbands,
                    const BlendWeights wts, float one) {
  extern __shared__ __align__(16) float fr_hs[];            // [state][row of the band + 1][fw]
  const int tid = threadIdx.x;
  int idx = blockIdx.x;
  const int band = idx % nbands;
  idx /= nbands;
  const int p = idx % (n - 1) + 1;
  const int c = idx / (n - 1);
  const int i_lo = band * band_rows, i_hi = min(Hg, i_lo + band_rows);      // floor source rows [i_lo, i_hi)
  const int nrows = i_hi - i_lo + 1;                                        // plus the row below the last one
  const int lplane = Hg * Wg;
  const long long ls = static_cast<long long>(C) * lplane;
  const u64 one2 = pack2(one, one);
  // output rows of the band: floor(sh * y) in [i_lo, i_hi) with the float arithmetic of up_coord()
  auto src_row = [&](int y) { return static_cast<int>(__fmul_rn(sh, static_cast<float>(y))); };
  auto first_row_at = [&](int i) {          // smallest y with src_row(y) >= i
    if (i <= 0) return 0;
    int y = (sh > 0.f) ? static_cast<int>(static_cast<float>(i) / sh) : fh;
    y = max(0, min(y, fh));
    while (y > 0 && src_row(y - 1) >= i) --y;
    while (y < fh && src_row(y) < i) ++y;
    return y;
  };
  const int y_lo = first_row_at(i_lo), y_hi = (i_hi >= Hg) ? fh : first_row_at(i_hi);
  if (y_hi <= y_lo) return;

  // ---- phase 1: horizontal two-terms (UpSample.cuh: w0*a + w1*b) of the band's source rows, both states
  const float* Lp = Lst + (p - 1) * ls + static_cast<long long>(c) * lplane;          // L_p
  const float* Rp = Rst + (n - p - 1) * ls + static_cast<long long>(c) * lplane;      // R_{n-p}
  const int per_state = nrows * fw;
  // per output row of the band: {shared-memory offset of source row i0, offset of row i0 + ip, l0, l1}, computed once
  // per CTA instead of once per 4-pixel item (the coordinate arithmetic with its int<->float conversions was a sixth
  // of the instructions of this instruction-bound kernel)
  float4* rowtab = reinterpret_cast<float4*>(fr_hs + 2 * per_state);
  for (int yy = tid; yy < y_hi - y_lo; yy += FR_THREADS) {
    const UpCoord hc = up_coord<Nm>(sh, y_lo + yy, Hg);
    const int o0 = (hc.i0 - i_lo) * fw;
    rowtab[yy] = make_float4(__int_as_float(o0), __int_as_float(o0 + hc.ip * fw), hc.l0, hc.l1);
  }
  // a thread owns columns x = tid, tid + 256, ...: one up_coord per column, then a walk down the band's rows
  // (the first version recomputed the coordinate and an integer division per element)
  for (int x = tid; x < fw; x += FR_THREADS) {
    const UpCoord wc = up_coord<Nm>(sw, x, Wg);
    const float* ql = Lp + wc.i0;
    const float* qr = Rp + wc.i0;
    float* d = fr_hs + x;
    int row = i_lo;
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
      const int off = row * Wg;
      const float a0 = __ldg(ql + off), a1 = __ldg(ql + off + wc.ip), b0 = __ldg(qr + off), b1 = __ldg(qr + off + wc.ip);
      d[0] = two_term<Nm::kUpInner>(wc.l0, a0, wc.l1, a1);
      d[per_state] = two_term<Nm::kUpInner>(wc.l0, b0, wc.l1, b1);
      d += fw;
      row = min(row + 1, Hg - 1);
    }
  }
  __syncthreads();

  // ---- phase 2: vertical two-term of both states, blend, store; thread = (group of 4 columns, row phase)
  const u64 w0 = pack2(wts.w0[p], wts.w0[p]), w1 = pack2(wts.w1[p], wts.w1[p]);
  const int groups = fw >> 2;
  float* o = out + (static_cast<long long>(p) * C + c) * fh * fw;
  const int rsplit = max(1, FR_THREADS / groups);
  for (int g0 = 0; g0 < groups; g0 += FR_THREADS) {            // one pass unless fw > 1024
    const int grp = (groups >= FR_THREADS) ? g0 + tid : tid % groups;
    const int rphase = (groups >= FR_THREADS) ? 0 : tid / groups;
    if (grp >= groups || rphase >= rsplit) continue;
    const int x = grp << 2;
    const float* hx = fr_hs + x;
    float* ox = o + static_cast<long long>(y_lo) * fw + x;
    for (int yy = rphase; yy < y_hi - y_lo; yy += rsplit) {
      const float4 rt = rowtab[yy];
      const u64 hl0 = pack2(rt.z, rt.z), hl1 = pack2(rt.w, rt.w);
      const float* r0 = hx + __float_as_int(rt.x);
      const float* r1 = hx + __float_as_int(rt.y);
      const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(r0), f1 = *reinterpret_cast<const ulonglong2*>(r1);
      const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(r0 + per_state), b1 = *reinterpret_cast<const ulonglong2*>(r1 + per_state);
      const u64 fa = two_term2<Nm::kUpOuter>(hl0, f0.x, hl1, f1.x, one2), fb = two_term2<Nm::kUpOuter>(hl0, f0.y, hl1, f1.y, one2);
      const u64 ba = two_term2<Nm::kUpOuter>(hl0, b0.x, hl1, b1.x, one2), bb = two_term2<Nm::kUpOuter>(hl0, b0.y, hl1, b1.y, one2);
      float4 v;
      unpack2(blend2x2(w0, fa, w1, ba, one2), v.x, v.y);
      unpack2(blend2x2(w0, fb, w1, bb, one2), v.z, v.w);
      __stcs(reinterpret_cast<float4*>(ox + static_cast<long long>(yy) * fw), v);
    }
  }
}


}  // namespace
}  // namespace fuvs

extern "C" long long fuvs_feature_scratch_floats(int C, int Hg, int Wg, int Hd, int Wd, int n) {
  long long f = 0;
  if (n > 1) f += 2ll * (n - 1) * C * static_cast<long long>(Hg) * Wg;       // chain states of both sides
  f += static_cast<long long>(C) * Hd * Wd;                                  // key frame warped through the default grid
  return f;
}

extern "C" int fuvs_feature_interval(const float* f_prev, const float* f_next, const float* grids_left,
                                     const float* grids_right, const float* default_grid, int Hd, int Wd, int C, int fh,
                                     int fw, int Hg, int Wg, int n, float* scratch, float* out, fuvs_stream_t stream) {
  const float* gl[FUVS_MAX_FRAMES];
  const float* gr[FUVS_MAX_FRAMES];
  if (n > FUVS_MAX_FRAMES) return fuvs::set_error(FUVS_EINVAL, "feature_interval: n=%d exceeds %d frames", n, FUVS_MAX_FRAMES);
  if (n > 1 && (!grids_left || !grids_right)) return fuvs::set_error(FUVS_EINVAL, "feature_interval: grids missing but n=%d", n);
  const long long g = static_cast<long long>(Hg) * Wg * 2;
  for (int j = 0; j < n - 1; ++j) {
    gl[j] = grids_left + j * g;
    gr[j] = grids_right + j * g;
  }
  return fuvs_feature_interval_ptrs(f_prev, f_next, gl, gr, default_grid, Hd, Wd, C, fh, fw, Hg, Wg, n, scratch, out, stream);
}

extern "C" int fuvs_feature_interval_ptrs(const float* f_prev, const float* f_next, const float* const* grids_left,
                                          const float* const* grids_right, const float* default_grid, int Hd, int Wd,
                                          int C, int fh, int fw, int Hg, int Wg, int n, float* scratch, float* out,
                                          fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!f_prev || !out || C < 1 || fh < 1 || fw < 1 || n < 1)
    return set_error(FUVS_EINVAL, "feature_interval: bad shape C=%d %dx%d n=%d", C, fh, fw, n);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "feature_interval: n=%d exceeds %d frames", n, FUVS_MAX_FRAMES);
  if (n > 1 && (!f_next || !grids_left || !grids_right || Hg < 1 || Wg < 1))
    return set_error(FUVS_EINVAL, "feature_interval: next/grids missing but n=%d", n);
  if ((n > 1 || default_grid) && !scratch) return set_error(FUVS_EINVAL, "feature_interval: scratch is NULL");
  if (default_grid && (Hd < 1 || Wd < 1)) return set_error(FUVS_EINVAL, "feature_interval: empty default grid");
  const long long plane = static_cast<long long>(fh) * fw;
  if (plane >= (1ll << 31) || static_cast<long long>(C) * Hg * Wg >= (1ll << 31))
    return set_error(FUVS_EINVAL, "feature_interval: problem exceeds 32-bit indexing");
  for (int j = 0; n > 1 && j < n - 1; ++j) {
    if (!grids_left[j] || !grids_right[j]) return set_error(FUVS_EINVAL, "feature_interval: grid %d is NULL", j);
    if (!aligned8(grids_left[j]) || !aligned8(grids_right[j]))
      return set_error(FUVS_EALIGN, "feature_interval: grids must be 8-byte aligned");
  }
  if (default_grid && !aligned8(default_grid)) return set_error(FUVS_EALIGN, "feature_interval: grids must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long S = static_cast<long long>(C) * plane;
  const long long ls = static_cast<long long>(C) * Hg * Wg;
  float* Lst = scratch;
  float* Rst = scratch + (n > 1 ? (n - 1) * ls : 0);
  float* tmp0 = scratch + (n > 1 ? 2 * (n - 1) * ls : 0);

  if (n > 1) {
    // chains at grid resolution (flow/model.py:133-137, 144-148); channel-chunked launches fill the SMs for C ~ 2048
    for (int j = 1; j <= n - 1; ++j) {
      const float* sL = (j == 1) ? f_prev : Lst + (j - 2) * ls;
      const float* sR = (j == 1) ? f_next : Rst + (j - 2) * ls;
      const int Hin = (j == 1) ? fh : Hg, Win = (j == 1) ? fw : Wg;
      if (int e = launch_warp_step_nm(sL, grids_left[j - 1], Lst + (j - 1) * ls, sR, grids_right[j - 1], Rst + (j - 1) * ls,
                                      C, Hin, Win, Hg, Wg, 0, st))
        return e;
    }
    BlendWeights w;
    make_blend_weights(n, &w);
    if (Hg == fh && Wg == fw) {
      // sizes match: the reference skips the interpolate (flow/model.py:138), plain blends
      for (int p = 1; p < n; ++p) {
        if (int e = fuvs_blend_argmax(Lst + (p - 1) * ls, Rst + (n - p - 1) * ls, static_cast<double>(n - p) / n,
                                      static_cast<double>(p) / n, 1, C, plane, out + p * S, nullptr, stream))
          return e;
      }
    } else {
      const float sh = fh > 1 ? static_cast<float>(Hg - 1) / (fh - 1) : 0.f;
      const float sw = fw > 1 ? static_cast<float>(Wg - 1) / (fw - 1) : 0.f;
      const bool vec4 = (fw % 4 == 0) && aligned16(out);
      // staged kernel: bands of source rows sized for four CTAs per SM (2 states x (rows + 1) x fw floats <= 48 KB)
      int band_rows = static_cast<int>((46 * 1024) / (8ll * fw)) - 1;          // 2 KB of the 48 KB for the row table
      if (band_rows > Hg) band_rows = Hg;
      long long nb = band_rows >= 1 ? (Hg + band_rows - 1) / band_rows : 0;
      if (nb > 0) band_rows = static_cast<int>((Hg + nb - 1) / nb);            // equal bands
      const long long ctas = nb * C * (n - 1);
      if (vec4 && band_rows >= 1 && ctas <= 0x7fffffffll && fh > 1 && fw > 1) {
        // + the row table: a band of band_rows source rows covers at most band_rows (fh-1)/(Hg-1) + 2 output rows
        const long long mo = (fh > 1 && Hg > 1) ? (static_cast<long long>(band_rows) * (fh - 1) + Hg - 2) / (Hg - 1) + 2 : fh;
        const size_t smem = static_cast<size_t>(2) * (band_rows + 1) * fw * sizeof(float) + static_cast<size_t>(mo < fh ? mo : fh) * 16;
        static SmemOptIn optin;
        if (smem <= 48 * 1024 && optin.ensure(feature_rows_kernel, 48 * 1024)) {
          feature_rows_kernel<<<static_cast<int>(ctas), FR_THREADS, smem, st>>>(Lst, Rst, out, C, fh, fw, Hg, Wg, n, sh, sw,
                                                                              band_rows, static_cast<int>(nb), w, 1.0f);
          if (int e = check_launch("fuvs_feature_interval(staged up-sample + blend)")) return e;
          goto frame0;
        }
      }
      {
      const long long items = static_cast<long long>(C) * fh * (vec4 ? fw / 4 : fw);
      long long grid = (items + 255) / 256;
      const long long cap = 32ll * sm_count();
      if (grid > cap) grid = cap;
      if (vec4)
        feature_up_blend_kernel<4><<<static_cast<int>(grid), 256, 0, st>>>(Lst, Rst, out, C, fh, fw, Hg, Wg, n, sh, sw, w);
      else
        feature_up_blend_kernel<1><<<static_cast<int>(grid), 256, 0, st>>>(Lst, Rst, out, C, fh, fw, Hg, Wg, n, sh, sw, w);
      if (int e = check_launch("fuvs_feature_interval(up-sample + blend)")) return e;
      }
    }
  }
frame0:
  // frame 0 (flow/model.py:154-159): the key frame's features through the default grid with align_corners=True,
  // restored to the feature size; without a default grid (no_warp) the features themselves
  if (default_grid) {
    if (Hd == fh && Wd == fw) {
      return launch_warp_step_nm(f_prev, default_grid, out, nullptr, nullptr, nullptr, C, fh, fw, Hd, Wd, 1, st);
    }
    if (int e = launch_warp_step_nm(f_prev, default_grid, tmp0, nullptr, nullptr, nullptr, C, fh, fw, Hd, Wd, 1, st)) return e;
    return fuvs_upsample_bilinear_ac(tmp0, out, C, Hd, Wd, fh, fw, stream);
  }
  if (cudaMemcpyAsync(out, f_prev, S * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return set_error(FUVS_ECUDA, "feature_interval: key-frame copy failed");
  return FUVS_OK;
}
