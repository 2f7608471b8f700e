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
#include <cstdlib>
#include <type_traits>

#include "fuvs_common.cuh"
#include "pix4.cuh"

namespace fuvs {

namespace {

constexpr int BR_THREADS = 256;      // two CTAs per SM: one CTA's phase 1 overlaps the other's phase 2

constexpr int KROWS = 4;             // KLR: source rows of the decoder-resolution key frame one CTA may touch

// KLR: key0 is the key frame at DECODER resolution [CT,hl,wl]: frame 0 = arg-max of its
// up-sample (F.interpolate(bilinear, align_corners=True), flow/model.py:191-193), evaluated like the chain states — the
// horizontal two-terms of the <= KROWS source rows this CTA's output rows touch are staged once, each output row is one
// vertical two-term.  The 41 MB full-resolution key frame is neither written by an up-sample launch nor read here.
// SPEC: the reference's shape of the route (k = 5 frames per interval, 256-pixel column chunks) with both as compile-time
// constants: the frame loop unrolls and the 20 shared-memory operands of a frame become immediate offsets of one base.
template <int CT, bool COUNTS, bool LOGITS, bool KLR, bool SPEC>
__global__ void __launch_bounds__(BR_THREADS, 2)
block_rows_kernel(const float* __restrict__ key0, const float* __restrict__ Lst, const float* __restrict__ Rst, int H,
                  int W, int Hg, int Wg, int n, float sh, float sw, int XW, int nchunks, int rsplit,
                  uint8_t* __restrict__ labels, float* __restrict__ logits, const uint8_t* __restrict__ tc_prev,
                  unsigned long long* __restrict__ counts, int ignore_index, const BlendWeights wts, float one,
                  int hl, int wl, float shk, float swk) {
  if (SPEC) { XW = 256; n = 5; }
  extern __shared__ __align__(16) float br_hs[];            // [p-1][side][row][c][XW], then the key-frame staging slots
  __shared__ unsigned sh24[24];
  // full-resolution key frame: [BR_THREADS][CT][4] cp.async slots; KLR: [KROWS][CT][XW] horizontal two-terms
  float* key_stage = br_hs + static_cast<size_t>(4) * (n - 1) * CT * XW;
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
    // KLR: first source row of the key frame under this CTA's output rows (same float arithmetic as up_coord)
    const int kfirst = KLR ? static_cast<int>(__fmul_rn(shk, static_cast<float>(y_lo))) : 0;
    // work item = (state, column): 2(n-1) * xw items over all threads of the CTA
    const int nstates = 2 * (n - 1);
    // (four items in flight: the 4*CT gathers of an item are L2 round trips, and eight dependent rounds per thread were
    // 11 % of the warp time, profiles/r02_ncu_block_rows.txt; measured 57.2 -> 53.8 us per interval, unroll 8: 54.8)
#pragma unroll 4
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
    if (KLR) {
      for (int e = tid; e < KROWS * xw; e += BR_THREADS) {
        const int r = e / xw, xx = e - r * xw;
        const UpCoord wc = up_coord<Nm>(swk, x0 + xx, wl);
        const int o = min(kfirst + r, hl - 1) * wl + wc.i0;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const float* pl = key0 + c * (hl * wl);
          key_stage[(r * CT + c) * XW + xx] = two_term<Nm::kUpInner>(wc.l0, __ldg(pl + o), wc.l1, __ldg(pl + o + wc.ip));
        }
      }
    }
    __syncthreads();

    // ---- phase 2: thread = (column group of 4 pixels, row phase); rows y_lo + rphase, + rsplit, ...
    const int ngroups = xw >> 2;
    const int grp = tid % ngroups, rphase = tid / ngroups;
    if (rphase < rsplit) {
      const int xx = grp * 4;
      const bool have_tc = COUNTS && tc_prev != nullptr;
      float* kslot = key_stage + static_cast<size_t>(tid) * (CT * 4);
      // Labels live as packed float indices {pixel 0, pixel 1}, {pixel 2, pixel 3} (pix4.cuh): the arg-max of a
      // frame without NaN runs in the float domain, a frame with a NaN class value takes the exact scan and converts.
      const u64 magic2 = pack2(8388608.f, 8388608.f);
      const u64 fw2 = pack2(static_cast<float>(FC::FW), static_cast<float>(FC::FW));
      auto fields = [&](const u64 (&idx)[2], unsigned (&fld)[4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float ml, mh;
          unpack2(fma2_rn(idx[h], fw2, magic2), ml, mh);
          fld[2 * h] = one_shl_wrap(__float_as_uint(ml));
          fld[2 * h + 1] = one_shl_wrap(__float_as_uint(mh));
        }
      };
      // pair (output frame idx, target frame tgt): O += fields(idx), T += s_tgt, I += fields(idx) where equal
      auto count_pair = [&](const u64 (&idx)[2], const unsigned (&fld)[4], const u64 (&tgt)[2], unsigned s_tgt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float il, ih, tl, th;
          unpack2(idx[h], il, ih);
          unpack2(tgt[h], tl, th);
          cnt.accI += (il == tl) ? fld[2 * h] : 0u;
          cnt.accI += (ih == th) ? fld[2 * h + 1] : 0u;
        }
        cnt.accO += (fld[0] + fld[1]) + (fld[2] + fld[3]);
        cnt.accT += s_tgt;
      };
      auto exact_scan = [&](const u64 (&xp)[2][CT], u64 (&idx)[2]) {
        float x[CT][4];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          unpack2(xp[0][c], x[c][0], x[c][1]);
          unpack2(xp[1][c], x[c][2], x[c][3]);
        }
        int lab[4];
        argmaxN<CT, 4, true>(x, lab);
        idx[0] = pack2(static_cast<float>(lab[0]), static_cast<float>(lab[1]));
        idx[1] = pack2(static_cast<float>(lab[2]), static_cast<float>(lab[3]));
      };
      auto start_key = [&](long long pix) {      // key frame of one row -> this thread's staging slot (no registers held)
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const unsigned dsts = static_cast<unsigned>(__cvta_generic_to_shared(kslot + c * 4));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsts), "l"(key0 + c * HW + pix) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };

      // R output rows per pass (rows y, y + rsplit): every row of this CTA interpolates between the SAME two source
      // rows, so the 4 CT shared-memory operands of a frame are loaded once and serve both rows — half the LDS.128 per
      // pixel and two independent FMA chains per thread (r02: the kernel sat on fixed-latency and short-scoreboard
      // stalls with 4 warps per scheduler, profiles/r02_ncu_block_rows.txt).
      auto rows = [&](auto r_, int y) {
        constexpr int R = decltype(r_)::value;
        u64 hl0[R], hl1[R];
        long long pix[R];
        unsigned tc_word[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const UpCoord hc = up_coord<Nm>(sh, y + r * rsplit, Hg);
          hl0[r] = pack2(hc.l0, hc.l0);
          hl1[r] = pack2(hc.l1, hc.l1);
          pix[r] = static_cast<long long>(y + r * rsplit) * W + x0 + xx;
          // the previous interval's last label map (target of frame 0): loaded here, used at the very end
          tc_word[r] = have_tc ? PixIO<2>::load_labels(tc_prev + pix[r]) : 0u;
        }
        // The key frame (frame 0) is the only operand streamed from HBM: start its copy into this thread's staging
        // slot now (cp.async, no registers held) and consume it AFTER frames 1..n-1 — the counts are sums over
        // (frame p, frame p-1) pairs, so their order is free.  (Consumed first, its load latency was the top stall.)
        if (!KLR) start_key(pix[0]);
        u64 idx1[R][2], last[R][2];
        unsigned s_last[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          idx1[r][0] = idx1[r][1] = last[r][0] = last[r][1] = 0ull;
          s_last[r] = 0u;
        }
        int since_spill = 2 * R;                  // in units of 4 labels
#pragma unroll                                    // (SPEC; without the unroll 50.0 vs 49.3 us per interval)
        for (int p = 1; p < (SPEC ? 5 : n); ++p) {
          const u64 w0 = pack2(wts.w0[p], wts.w0[p]), w1 = pack2(wts.w1[p], wts.w1[p]);
          const float* hsL = br_hs + static_cast<size_t>(((p - 1) * 2 + 0) * 2) * CT * XW + xx;
          const float* hsR = br_hs + static_cast<size_t>(((p - 1) * 2 + 1) * 2) * CT * XW + xx;
          u64 xp[R][2][CT];
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            // four consecutive pixels arrive as two register pairs per LDS.128: the vertical two-term of
            // UpSample.cuh (h0*r0 + h1*r1) runs on packed FP32x2 without any repacking
            const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(hsL + (0 * CT + c) * XW);
            const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(hsL + (1 * CT + c) * XW);
            const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(hsR + (0 * CT + c) * XW);
            const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(hsR + (1 * CT + c) * XW);
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const u64 fa = two_term2<Nm::kUpOuter>(hl0[r], f0.x, hl1[r], f1.x, one2), fb = two_term2<Nm::kUpOuter>(hl0[r], f0.y, hl1[r], f1.y, one2);
              const u64 ba = two_term2<Nm::kUpOuter>(hl0[r], b0.x, hl1[r], b1.x, one2), bb = two_term2<Nm::kUpOuter>(hl0[r], b0.y, hl1[r], b1.y, one2);
              xp[r][0][c] = blend2x2(w0, fa, w1, ba, one2);
              xp[r][1][c] = blend2x2(w0, fb, w1, bb, one2);
              if (LOGITS) {
                float xs[4];
                unpack2(xp[r][0][c], xs[0], xs[1]);
                unpack2(xp[r][1][c], xs[2], xs[3]);
                PixIO<2>::store(logits + (static_cast<long long>(p) * CT + c) * HW + pix[r], xs);
              }
            }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            u64 idx[2];
            bool has_nan = false;                        // the class maxima propagate NaN (max.NaN): no separate probe
            idx[0] = argmax2f_nan<CT>(xp[r][0], &has_nan);
            idx[1] = argmax2f_nan<CT>(xp[r][1], &has_nan);
            if (has_nan) exact_scan(xp[r], idx);
            if (labels) PixIO<2>::store_label_word(labels + static_cast<long long>(p) * HW + pix[r], PixIO<2>::label_word(idx));
            if (COUNTS) {
              unsigned fld[4];
              fields(idx, fld);
              if (p == 1) {
                idx1[r][0] = idx[0];
                idx1[r][1] = idx[1];
              } else {
                count_pair(idx, fld, last[r], s_last[r]);
                if (++since_spill >= FC::CAP / 4) {
                  cnt.spill();
                  since_spill = 0;
                }
              }
              s_last[r] = (fld[0] + fld[1]) + (fld[2] + fld[3]);
              last[r][0] = idx[0];
              last[r][1] = idx[1];
            }
          }
        }
        // frame 0: the key frame itself (flow/model.py:195-197), then the pairs (0, previous interval) and (1, 0)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (!KLR) asm volatile("cp.async.wait_group 0;" ::: "memory");
          if (COUNTS && 4 * since_spill + 8 > FC::CAP) {       // room for the two pairs below
            cnt.spill();
            since_spill = 0;
          }
          since_spill += 2;
          u64 xp[2][CT];
          if (KLR) {
            const UpCoord kh = up_coord<Nm>(shk, y + r * rsplit, hl);
            const u64 kl0 = pack2(kh.l0, kh.l0), kl1 = pack2(kh.l1, kh.l1);
            const float* ka = key_stage + static_cast<size_t>((kh.i0 - kfirst) * CT) * XW + xx;
            const float* kb = ka + static_cast<size_t>(kh.ip * CT) * XW;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
              const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(ka + c * XW);
              const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(kb + c * XW);
              xp[0][c] = two_term2<Nm::kUpOuter>(kl0, f0.x, kl1, f1.x, one2);
              xp[1][c] = two_term2<Nm::kUpOuter>(kl0, f0.y, kl1, f1.y, one2);
              if (LOGITS) {
                float xs[4];
                unpack2(xp[0][c], xs[0], xs[1]);
                unpack2(xp[1][c], xs[2], xs[3]);
                PixIO<2>::store(logits + c * HW + pix[r], xs);
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < CT; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(kslot + c * 4);
              xp[0][c] = pack2(v.x, v.y);
              xp[1][c] = pack2(v.z, v.w);
              if (LOGITS) {
                const float xs[4] = {v.x, v.y, v.z, v.w};
                PixIO<2>::store(logits + c * HW + pix[r], xs);
              }
            }
          }
          u64 idx0[2];
          bool nan0 = false;                           // float-domain arg-max like frames 1..n-1, exact scan only for NaN
          idx0[0] = argmax2f_nan<CT>(xp[0], &nan0);
          idx0[1] = argmax2f_nan<CT>(xp[1], &nan0);
          if (nan0) exact_scan(xp, idx0);
          // the slot's values are consumed (the scan above depends on them): the next row's key frame may land in it
          // (requesting it into registers during the last frame instead hides its latency and changes nothing: 49.1 us)
          if (!KLR && r + 1 < R) start_key(pix[r + 1]);
          const unsigned w = PixIO<2>::label_word(idx0);
          if (labels) PixIO<2>::store_label_word(labels + pix[r], w);
          if (COUNTS) {
            unsigned fld0[4];
            fields(idx0, fld0);
            if (have_tc) {
              const unsigned t = tc_word[r];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int tl = (t >> (8 * i)) & 255u, lab = (w >> (8 * i)) & 255u;
                const unsigned ft = (tl < CT) ? FieldCounts<CT>::field(tl) : 0u;
                const unsigned fo = (tl == ignore_index) ? 0u : fld0[i];
                cnt.add(lab, fo, tl, ft);
              }
            }
            if (n > 1) {
              unsigned fld1[4];
              fields(idx1[r], fld1);
              count_pair(idx1[r], fld1, idx0, (fld0[0] + fld0[1]) + (fld0[2] + fld0[3]));
            }
          }
        }
        if (COUNTS) cnt.spill();
      };
      int y = y_lo + rphase;
      for (; y + rsplit < y_hi; y += 2 * rsplit) rows(std::integral_constant<int, 2>{}, y);
      if (y < y_hi) rows(std::integral_constant<int, 1>{}, y);
    }
  }
  if (COUNTS) cnt.finish(sh24, counts, CT);
}

// KLR eligibility: the output rows of every source-row interval must touch at most KROWS rows of the low-resolution key
// frame (same row arithmetic as the kernel)
static bool key_rows_fit(int H, int Hg, float sh, int hl, float shk) {
  auto src_row = [&](int y) { return static_cast<int>(sh * static_cast<float>(y)); };
  int y = 0;
  while (y < H) {
    const int i0 = src_row(y);
    int y_hi = y;
    while (y_hi < H && src_row(y_hi) == i0) ++y_hi;
    const int first = static_cast<int>(shk * static_cast<float>(y));
    const int last = static_cast<int>(shk * static_cast<float>(y_hi - 1));
    const int need = last + ((last < hl - 1) ? 1 : 0) - first + 1;
    if (need > KROWS) return false;
    y = y_hi;
  }
  (void)Hg;
  return true;
}

template <int CT, bool KLR>
int launch_ct(const float* key0, const float* Lst, const float* Rst, int H, int W, int Hg, int Wg, int n, float sh,
              float sw, uint8_t* labels, float* logits, const uint8_t* tc_prev, long long* counts, int ignore_index,
              const BlendWeights& w, cudaStream_t st, int hl, int wl, float shk, float swk, bool* counts_done) {
  // shared memory: (n-1) frames x 2 states x 2 rows x CT channels x XW columns
  const long long per_col = 4ll * (n - 1) * CT * 4;
  const int budget = 108 * 1024 - BR_THREADS * CT * 16;   // two CTAs per SM, minus the key-frame staging slots
  // column chunk: a multiple of 128 pixels (32 column groups), so that a warp works on ONE output row and warps
  // never diverge on the row loop; as wide as the budget allows up to 256 (4 rows in flight per CTA)
  int XW = 256;
  while (XW > 128 && per_col * XW > budget) XW -= 128;
  // (narrower chunks were measured in r01: 128 px 69.8 us, 64 px 79 us against 68.2 us per interval)
  if (per_col * XW > budget) return 1;
  if (W < XW) XW = (W + 3) & ~3;
  const int nchunks = (W + XW - 1) / XW;
  const size_t smem = static_cast<size_t>(per_col) * XW + static_cast<size_t>(BR_THREADS) * CT * 16;
  const int ngroups = XW / 4;
  int rsplit = BR_THREADS / ngroups;              // thread rows per CTA
  if (rsplit < 1) return 1;
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  const bool spec = CT == 5 && XW == 256 && n == 5;
#define FUVS_BR(CNT_, LG_)                                                                                             \
  do {                                                                                                                 \
    /* (the labels-only low-res variant spills 584 bytes when specialised: it keeps the generic kernel) */              \
    const bool sp = spec && !(KLR && !CNT_ && !LG_);                                                                   \
    auto kern = sp ? block_rows_kernel<CT, CNT_, LG_, KLR, (CT == 5)> : block_rows_kernel<CT, CNT_, LG_, KLR, false>;    \
    static SmemOptIn optin[2];                     /* one per kernel: the opt-in is an attribute of the function */      \
    if (!optin[sp ? 1 : 0].ensure(kern, 110 * 1024)) return 1;                                                         \
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
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, XW, nchunks, rsplit, \
                                              labels, logits, tc_prev, cu, ignore_index, w, 1.0f, hl, wl, shk, swk);   \
    if (le != cudaSuccess) return set_error(FUVS_ECUDA, "fuvs_block_interval(stream rows): %s", cudaGetErrorString(le)); \
  } while (0)
  // The counts stay in the kernel only on the low-resolution route (37.8 us per 1080p interval against ~44 with the
  // separate launch).  With full-resolution key frames they are 12 % of the kernel's instructions and the bit-plane
  // fuvs_temporal_counts (4.3 us) does them cheaper: block 48.1 -> 47.3, clip entry 45.4 -> 42.0 us per interval.
  *counts_done = false;
  bool fused = false;
  if constexpr (KLR) {
    if (counts) {
      if (logits) FUVS_BR(true, true); else FUVS_BR(true, false);
      fused = true;
      *counts_done = true;
    }
  }
  if (!fused) {
    cu = nullptr;
    tc_prev = nullptr;
    if (logits) FUVS_BR(false, true); else FUVS_BR(false, false);
  }
#undef FUVS_BR
  count_launch();
  return FUVS_OK;
}

}  // namespace

// Returns FUVS_OK if it ran (labels and logits are done; *counts_done says whether the temporal counts are too), 1 if the
// shape is not eligible (the caller uses block_stream_cols_kernel), negative on error.
// hl > 0: key0 is the key frame at decoder resolution [C,hl,wl] (frame 0 = arg-max of its up-sample).
int launch_block_stream_rows(const float* key0, const float* Lst, const float* Rst, int C, int H, int W, int Hg, int Wg,
                             int n, float sh, float sw, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                             long long* counts, int ignore_index, const BlendWeights& w, cudaStream_t st, int hl, int wl,
                             bool* counts_done) {
  *counts_done = false;
  if (C < 2 || C > 5 || (W & 3) != 0 || n < 2) return 1;
  const bool klr = hl > 0;
  if ((!klr && !aligned16(key0)) || (logits && !aligned16(logits)) || (labels && !aligned4(labels)) || (tc_prev && !aligned4(tc_prev)))
    return 1;
  if (counts && ((ignore_index >= 0 && ignore_index < C) || !labels)) return 1;
  if (H == Hg && W == Wg) return 1;                 // nothing to up-sample: per-pixel kernel
  float shk = 0.f, swk = 0.f;
  if (klr) {
    shk = H > 1 ? static_cast<float>(hl - 1) / (H - 1) : 0.f;       // area_pixel_compute_scale, align_corners=True
    swk = W > 1 ? static_cast<float>(wl - 1) / (W - 1) : 0.f;
    if (static_cast<long long>(C) * hl * wl >= (1ll << 31) || !key_rows_fit(H, Hg, sh, hl, shk)) return 1;
  }
#define FUVS_BRC(CT_)                                                                                                  \
  return klr ? launch_ct<CT_, true>(key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts, ignore_index, w, \
                                    st, hl, wl, shk, swk, counts_done)                                                  \
             : launch_ct<CT_, false>(key0, Lst, Rst, H, W, Hg, Wg, n, sh, sw, labels, logits, tc_prev, counts, ignore_index, \
                                     w, st, 0, 0, 0.f, 0.f, counts_done)
  switch (C) {
    case 2: FUVS_BRC(2);
    case 3: FUVS_BRC(3);
    case 4: FUVS_BRC(4);
    default: FUVS_BRC(5);
  }
#undef FUVS_BRC
}

}  // namespace fuvs
