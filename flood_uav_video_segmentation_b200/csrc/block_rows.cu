// block_rows.cu — streaming half of the macro-block-grid interval, organised by source-row intervals.
//
// frame p = fl(fl(w0p * up(L_p)) + fl(w1p * up(R_{n-p}))), p = 1..n-1, with up() = F.interpolate(bilinear,
// align_corners=True) from the chain states [C,Hg,Wg] to the frame size (flow/model.py:216-219, 226-229, 233-237),
// frame 0 = the key frame, arg-max (flow/base.py:276) and temporal-consistency counts (flow/base.py:280-295).
//
// The column-strip kernel (block.cu: block_stream_cols_kernel) evaluates the vertical AND horizontal two-terms per
// output pixel from L1: 719 instructions per pixel, long-scoreboard bound (profiles/r01_ncu_block_cols.txt).  Here a
// CTA owns one source-row interval i0 (the ~16 output rows whose floor source row is i0) and a chunk of columns:
//   phase 1  the horizontal two-terms of source rows i0 and i0+1 of all 2(n-1) chain states are computed ONCE into
//            shared memory: [n-1 frames][L|R][2 rows][C][XW] floats;
//   phase 2  each thread owns 4 consecutive pixels of some of the rows; per frame it reads 4*C float4 from shared
//            memory, does the vertical two-terms (2 instructions per value), blends with packed FP32x2, arg-maxes and
//            counts exactly like linear.cu.
// Same ATen arithmetic as block.cu (up_coord / two_term policies), so results are bit-identical.
#include "fuvs_common.cuh"
#include "pix4.cuh"

namespace fuvs {

namespace {

constexpr int BR_THREADS = 256;      // two CTAs per SM: one CTA's phase 1 overlaps the other's phase 2

template <int CT, bool COUNTS, bool LOGITS>
__global__ void __launch_bounds__(BR_THREADS, 2)
block_rows_kernel(const float* __restrict__ key0, const float* __restrict__ Lst, const float* __restrict__ Rst, int H,
                  int W, int Hg, int Wg, int n, float sh, float sw, int XW, int nchunks, int rsplit,
                  uint8_t* __restrict__ labels, float* __restrict__ logits, const uint8_t* __restrict__ tc_prev,
                  unsigned long long* __restrict__ counts, int ignore_index, const BlendWeights wts, float one) {
  extern __shared__ __align__(16) float br_hs[];            // [p-1][side][row][c][XW], then the key-frame staging slots
  __shared__ unsigned sh24[24];
  float* key_stage = br_hs + static_cast<size_t>(4) * (n - 1) * CT * XW;     // [BR_THREADS][CT][4]
  using FC = FieldCfg<CT>;
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x / nchunks, chunk = blockIdx.x - i0 * nchunks;
  const int x0 = chunk * XW;
  const int xw = min(XW, W - x0);
  const long long HW = static_cast<long long>(H) * W;
  const int lplane = Hg * Wg;
  const int ls = CT * lplane;
  const u64 one2 = pack2(one, one);
  const float zero = __fsub_rn(one, one);
  const u64 zero2 = pack2(zero, zero);
  FieldCounts<CT> cnt;
  cnt.init();
  // programmatic dependent launch: scheduled while the chain kernel drains; its states are read only after the wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // output rows of this interval: floor(sh * y) == i0 with the float arithmetic of up_coord()
  auto src_row = [&](int y) { return static_cast<int>(__fmul_rn(sh, static_cast<float>(y))); };
  int y_lo = (sh > 0.f) ? static_cast<int>(static_cast<float>(i0) / sh) : 0;
  y_lo = max(0, min(y_lo, H - 1));
  while (y_lo > 0 && src_row(y_lo - 1) >= i0) --y_lo;
  while (y_lo < H && src_row(y_lo) < i0) ++y_lo;
  int y_hi = y_lo;
  while (y_hi < H && src_row(y_hi) == i0) ++y_hi;

  if (y_hi > y_lo) {
    // ---- phase 1: horizontal two-terms (UpSample.cuh: w0*a + w1*b) of source rows i0, i0 + ip, every state
    const int ip_h = (i0 < Hg - 1) ? 1 : 0;
    // work item = (state, column): 2(n-1) * xw items over all threads of the CTA
    const int nstates = 2 * (n - 1);
    for (int e = tid; e < nstates * xw; e += BR_THREADS) {
      const int sidx = e / xw, xx = e - sidx * xw;          // sidx = (p-1)*2 + side
      const int p = (sidx >> 1) + 1, side = sidx & 1;
      const UpCoord wc = up_coord<Nm>(sw, x0 + xx, Wg);
      const int o0 = i0 * Wg + wc.i0, o1 = (i0 + ip_h) * Wg + wc.i0;
      const float* st = side ? Rst + (n - p - 1) * ls : Lst + (p - 1) * ls;         // R_{n-p} / L_p
      float* dst = br_hs + static_cast<size_t>(sidx * 2) * CT * XW + xx;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const float* pl = st + c * lplane;
        dst[(0 * CT + c) * XW] = two_term<Nm::kUpInner>(wc.l0, __ldg(pl + o0), wc.l1, __ldg(pl + o0 + wc.ip));
        dst[(1 * CT + c) * XW] = two_term<Nm::kUpInner>(wc.l0, __ldg(pl + o1), wc.l1, __ldg(pl + o1 + wc.ip));
      }
    }
    __syncthreads();

    // ---- phase 2: thread = (column group of 4 pixels, row phase); rows y_lo + rphase, + rsplit, ...
    const int ngroups = xw >> 2;
    const int grp = tid % ngroups, rphase = tid / ngroups;
    if (rphase < rsplit) {
      const int xx = grp * 4;
      for (int y = y_lo + rphase; y < y_hi; y += rsplit) {
        const UpCoord hc = up_coord<Nm>(sh, y, Hg);
        const u64 hl0 = pack2(hc.l0, hc.l0), hl1 = pack2(hc.l1, hc.l1);
        const long long pix = static_cast<long long>(y) * W + x0 + xx;
        // The key frame (frame 0) is the only operand streamed from HBM: start its copy into this thread's staging
        // slot now (cp.async, no registers held) and consume it AFTER frames 1..n-1 — the counts are sums over
        // (frame p, frame p-1) pairs, so their order is free.  (Consumed first, its load latency was the top stall.)
        float* kslot = key_stage + static_cast<size_t>(tid) * (CT * 4);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const unsigned dsts = static_cast<unsigned>(__cvta_generic_to_shared(kslot + c * 4));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsts), "l"(key0 + c * HW + pix) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");

        int lab1[4] = {0, 0, 0, 0}, last[4] = {0, 0, 0, 0};
        int since_spill = 2;
        for (int p = 1; p < n; ++p) {
          const u64 w0 = pack2(wts.w0[p], wts.w0[p]), w1 = pack2(wts.w1[p], wts.w1[p]);
          const float* hsL = br_hs + static_cast<size_t>(((p - 1) * 2 + 0) * 2) * CT * XW + xx;
          const float* hsR = br_hs + static_cast<size_t>(((p - 1) * 2 + 1) * 2) * CT * XW + xx;
          float x[CT][4];
          u64 probe = zero2;
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            // four consecutive pixels arrive as two register pairs per LDS.128: the vertical two-term of
            // UpSample.cuh (h0*r0 + h1*r1) runs on packed FP32x2 without any repacking
            const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(hsL + (0 * CT + c) * XW);
            const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(hsL + (1 * CT + c) * XW);
            const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(hsR + (0 * CT + c) * XW);
            const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(hsR + (1 * CT + c) * XW);
            const u64 fa = two_term2<Nm::kUpOuter>(hl0, f0.x, hl1, f1.x, one2), fb = two_term2<Nm::kUpOuter>(hl0, f0.y, hl1, f1.y, one2);
            const u64 ba = two_term2<Nm::kUpOuter>(hl0, b0.x, hl1, b1.x, one2), bb = two_term2<Nm::kUpOuter>(hl0, b0.y, hl1, b1.y, one2);
            const u64 va = blend2x2(w0, fa, w1, ba, one2), vb = blend2x2(w0, fb, w1, bb, one2);
            probe = fma2_rn(va, zero2, probe);
            probe = fma2_rn(vb, zero2, probe);
            unpack2(va, x[c][0], x[c][1]);
            unpack2(vb, x[c][2], x[c][3]);
            if (LOGITS) PixIO<2>::store(logits + (static_cast<long long>(p) * CT + c) * HW + pix, x[c]);
          }
          float pr0, pr1;
          unpack2(probe, pr0, pr1);
          int lab[4];
          if ((pr0 == pr0) && (pr1 == pr1)) argmaxN<CT, 4, false>(x, lab);     // x*0 is NaN iff x is Inf/NaN
          else argmaxN<CT, 4, true>(x, lab);
          if (labels) PixIO<2>::store_labels(labels + static_cast<long long>(p) * HW + pix, lab);
          if (COUNTS) {
            if (p == 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) lab1[i] = lab[i];
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) cnt.add(lab[i], FieldCounts<CT>::field(lab[i]), last[i], FieldCounts<CT>::field(last[i]));
              if (++since_spill >= FC::CAP / 4) {
                cnt.spill();
                since_spill = 0;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) last[i] = lab[i];
          }
        }
        // frame 0: the key frame itself (flow/model.py:195-197), then the pairs (0, previous interval) and (1, 0)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        {
          float x[CT][4];
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(kslot + c * 4);
            x[c][0] = v.x; x[c][1] = v.y; x[c][2] = v.z; x[c][3] = v.w;
            if (LOGITS) PixIO<2>::store(logits + c * HW + pix, x[c]);
          }
          int lab[4];
          argmaxN<CT, 4, true>(x, lab);
          if (labels) PixIO<2>::store_labels(labels + pix, lab);
          if (COUNTS) {
            if (tc_prev != nullptr) {
              const unsigned t = PixIO<2>::load_labels(tc_prev + pix);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int tl = (t >> (8 * i)) & 255u;
                const unsigned ft = (tl < CT) ? FieldCounts<CT>::field(tl) : 0u;
                const unsigned fo = (tl == ignore_index) ? 0u : FieldCounts<CT>::field(lab[i]);
                cnt.add(lab[i], fo, tl, ft);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) cnt.add(lab1[i], FieldCounts<CT>::field(lab1[i]), lab[i], FieldCounts<CT>::field(lab[i]));
          }
        }
        if (COUNTS) cnt.spill();
      }
    }
  }
  if (COUNTS) cnt.finish(sh24, counts, CT);
}

template <int CT>
int launch_ct(const float* key0, const float* Lst, const float* Rst, int H, int W, int Hg, int Wg, int n, float sh,
              float sw, uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts, int ignore_index,
              const BlendWeights& w, cudaStream_t st) {
  // shared memory: (n-1) frames x 2 states x 2 rows x CT channels x XW columns
  const long long per_col = 4ll * (n - 1) * CT * 4;
  const int budget = 108 * 1024 - BR_THREADS * CT * 16;   // two CTAs per SM, minus the key-frame staging slots
  // column chunk: a multiple of 128 pixels (32 column groups), so that a warp works on ONE output row and warps
  // never diverge on the row loop; as wide as the budget allows up to 256 (4 rows in flight per CTA)
  int XW = 256;
  while (XW > 128 && per_col * XW > budget) XW -= 128;
  if (per_col * XW > budget) return 1;
  if (W < XW) XW = (W + 3) & ~3;
  const int nchunks = (W + XW - 1) / XW;
  const size_t smem = static_cast<size_t>(per_col) * XW + static_cast<size_t>(BR_THREADS) * CT * 16;
  const int ngroups = XW / 4;
  int rsplit = BR_THREADS / ngroups;              // thread rows per CTA
  if (rsplit < 1) return 1;
  auto cu = reinterpret_cast<unsigned long long*>(counts);
#define FUVS_BR(CNT_, LG_)                                                                                             \
  do {                                                                                                                 \
    auto kern = block_rows_kernel<CT, CNT_, LG_>;                                                                      \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) { \
      cudaGetLastError();                                                                                              \
      return 1;                                                                                                        \
    }                                                                                                                  \
    cudaLaunchConfig_t cfg = {};                                                                                       \
    cfg.gridDim = dim3(Hg * nchunks);                                                                                  \
    cfg.blockDim = dim3(BR_THREADS);                                                                                   \
    cfg.dynamicSmemBytes = smem;                                                                                       \
    cfg.stream = st;                                                                                                   \
    cudaLaunchAttribute attr[1];                                                                                       \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                   \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                                            \
    cfg.attrs = attr;                                                                                                  \
    cfg.numAttrs = 1;                                                                                                  \
    cudaLaunchKernelEx(&cfg, kern, key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, XW, nchunks, rsplit, labels, logits,       \
                       tc_prev, cu, ignore_index, w, 1.0f);                                                            \
  } while (0)
  if (counts) { if (logits) FUVS_BR(true, true); else FUVS_BR(true, false); }
  else        { if (logits) FUVS_BR(false, true); else FUVS_BR(false, false); }
#undef FUVS_BR
  return check_launch("fuvs_block_interval(stream rows)");
}

}  // namespace

// Returns FUVS_OK if it ran (labels, logits and — when counts != NULL — the temporal counts are done), 1 if the
// shape is not eligible (the caller uses block_stream_cols_kernel), negative on error.
int launch_block_stream_rows(const float* key0, const float* Lst, const float* Rst, int C, int H, int W, int Hg, int Wg,
                             int n, float sh, float sw, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                             long long* counts, int ignore_index, const BlendWeights& w, cudaStream_t st) {
  if (C < 2 || C > 5 || (W & 3) != 0 || n < 2) return 1;
  if (!aligned16(key0) || (logits && !aligned16(logits)) || (labels && !aligned4(labels)) || (tc_prev && !aligned4(tc_prev)))
    return 1;
  if (counts && ((ignore_index >= 0 && ignore_index < C) || !labels)) return 1;
  if (H == Hg && W == Wg) return 1;                 // nothing to up-sample: per-pixel kernel
  switch (C) {
    case 2: return launch_ct<2>(key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 3: return launch_ct<3>(key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 4: return launch_ct<4>(key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts, ignore_index, w, st);
    default: return launch_ct<5>(key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts, ignore_index, w, st);
  }
}

}  // namespace fuvs
