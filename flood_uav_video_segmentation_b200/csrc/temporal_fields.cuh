// temporal_fields.cuh — the temporal-consistency counts of one 16-pixel column of a label sequence (flow/base.py:280-295:
// intersectionAndUnion of every frame against the one before it), shared by fuvs_temporal_counts' fast path (metric.cu)
// and the dense strip kernel's last step, which counts the pixels it has just written (dense_strip.cu).
#pragma once
#include "fuvs_common.cuh"

#include <type_traits>
#include <utility>

namespace fuvs {

template <int... I, class F>
__device__ __forceinline__ void static_for_tc_impl(std::integer_sequence<int, I...>, F&& f) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for_tc(F&& f) {
  static_for_tc_impl(std::make_integer_sequence<int, N>{}, f);
}

template <int KT>
__device__ __forceinline__ unsigned tc_invalid(const uint4& w) {
  constexpr unsigned ADD = (0x80u - KT) * 0x01010101u;
  const unsigned a = (((w.x & 0x7f7f7f7fu) + ADD) | w.x), b = (((w.y & 0x7f7f7f7fu) + ADD) | w.y);
  const unsigned c = (((w.z & 0x7f7f7f7fu) + ADD) | w.z), d = (((w.w & 0x7f7f7f7fu) + ADD) | w.w);
  return (a | b | c | d) & 0x80808080u;
}
__device__ __forceinline__ unsigned tc_one_shl_wrap(unsigned c) {
  unsigned d;
  asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(0u), "r"(1u), "r"(c));
  return d;
}
template <int KT>
__device__ __forceinline__ unsigned tc_fields(const uint4& w, unsigned (&fld)[16]) {
  constexpr unsigned FW = FieldCfg<KT>::FW;
  const unsigned ws[4] = {w.x, w.y, w.z, w.w};
  unsigned s = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned m = ws[j] * FW;             // bytes: FW * label <= 24, the low five bits of each are the shift
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      fld[4 * j + i] = tc_one_shl_wrap(m >> (8 * i));
      s += fld[4 * j + i];
    }
  }
  return s;
}

// ---------------------------------------------------------------------------
// Bit-plane counting (r02).  The field-packed counters above cost ~9 ALU instructions per label and frame (a shift, a
// wrap-shift and an add per label for the fields, three more for the I term, plus the spills), which made both
// fuvs_temporal_counts (10 us for 5 x 1080p) and the strip kernel's own counting ALU-bound.  Here the 16 labels of a
// frame (all < 8 once checked) are transposed into three words holding bit 0 / 1 / 2 of every label (12 instructions);
// the indicator word of a class is one LOP3 of the three, its count one POPC, class 0 is 16 minus the others; the
// I term of a pair is the POPC of the indicator under the "labels equal" word (three LOP3 for that word).  About 3
// instructions per label and frame.
// ---------------------------------------------------------------------------
struct LabelPlanes {
  unsigned a, b, c;                            // bit 0 / 1 / 2 of 16 labels; label i of word j sits at bit 4*(i + 4*(j&1)) + (j>>1)
};
__device__ __forceinline__ LabelPlanes tc_planes(const uint4& w) {
  constexpr unsigned M = 0x11111111u, M1 = 0x22222222u;
  const unsigned X = w.x + (w.y << 4), Y = w.z + (w.w << 4);       // labels < 16: one label per nibble
  LabelPlanes p;
  p.a = (X & M) | ((Y << 1) & M1);
  p.b = ((X >> 1) & M) | (Y & M1);
  p.c = ((X >> 2) & M) | ((Y >> 1) & M1);
  return p;
}
// nonzero if any of the 16 labels is >= KT
template <int KT>
__device__ __forceinline__ unsigned tc_planes_invalid(const uint4& w, const LabelPlanes& p) {
  const unsigned big = (w.x | w.y | w.z | w.w) & 0xf8f8f8f8u;       // a label >= 8 (the planes are then meaningless)
  unsigned over;                                                    // a label in [KT, 8)
  if (KT >= 8) over = 0u;
  else if (KT == 5) over = p.c & (p.a | p.b);
  else if (KT == 4) over = p.c;
  else if (KT == 3) over = p.c | (p.a & p.b);
  else if (KT == 2) over = p.c | p.b;
  else over = p.c | p.b | p.a;
  return big | over;
}
// indicator word of class c >= 1 (clean: at least one plane enters un-negated, and the planes are 0 in the unused bits)
template <int c>
__device__ __forceinline__ unsigned tc_indicator(const LabelPlanes& p) {
  const unsigned a = (c & 1) ? p.a : ~p.a, b = (c & 2) ? p.b : ~p.b, cc = (c & 4) ? p.c : ~p.c;
  return a & b & cc;
}

template <int KT>
struct ClassCounts {
  static_assert(KT >= 1 && KT <= 5, "bit-plane counting: at most 5 classes (three planes, classes 5-7 unused)");
  unsigned I[KT], O[KT], T[KT];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int c = 0; c < KT; ++c) { I[c] = 0u; O[c] = 0u; T[c] = 0u; }
  }
  // per-label path: current label o against the previous frame's label t (flow/base.py:283: a pixel whose target is
  // ignore_index is ignored in the output too)
  __device__ __forceinline__ void add_label(int o, int t, int ignore) {
#pragma unroll
    for (int c = 0; c < KT; ++c) {
      const bool oc = (o == c) && (t != ignore);
      O[c] += oc ? 1u : 0u;
      T[c] += (t == c) ? 1u : 0u;
      I[c] += (oc && t == c) ? 1u : 0u;
    }
  }
  // all threads of the block must call; sh: 24 unsigned
  __device__ __forceinline__ void finish(unsigned* sh, unsigned long long* counts, int Kruntime) {
    WarpTotals<KT> wt;
#pragma unroll
    for (int c = 0; c < KT; ++c) {
      wt.I[c] = __reduce_add_sync(0xffffffffu, I[c]);
      wt.O[c] = __reduce_add_sync(0xffffffffu, O[c]);
      wt.T[c] = __reduce_add_sync(0xffffffffu, T[c]);
    }
    block_flush_counts<KT>(wt, sh, counts, Kruntime);
  }
};

// per-class label counts of one frame from its indicator words
template <int KT>
__device__ __forceinline__ void tc_hist(const LabelPlanes& p, unsigned (&ind)[KT], unsigned (&h)[KT]) {
  unsigned rest = 0u;
  static_for_tc<KT - 1>([&](auto c_) {
    constexpr int c = decltype(c_)::value + 1;
    ind[c] = tc_indicator<c>(p);
    h[c] = __popc(ind[c]);
    rest += h[c];
  });
  ind[0] = 0u;
  h[0] = 16u - rest;
}

// One 16-pixel column: `first` (absent when !have_first) followed by nf frames; every frame is counted against the one
// before it.  ld(-1) returns the 16 labels of `first`, ld(p) those of frame p.  K <= 5, ignore outside [0,K) (checked
// on the host).  CHECK_FRAMES = false: the frames are arg-max output of this call (< KT by construction); `first` is
// always checked.  A pair with an out-of-range label (a caller's own label map, ignore_index) takes the per-label path.
// Loads run two frames ahead of the counting.
// NFC > 0: nf is this compile-time constant (the loop unrolls and ld's argument is a constant at every call).
template <int KT, bool CHECK_FRAMES, int NFC = 0, class LD>
__device__ __forceinline__ void tc_chain16_ld(ClassCounts<KT>& cnt, bool have_first, int nf_rt, int ignore, LD&& ld) {
  const int nf = NFC > 0 ? NFC : nf_rt;
  uint4 last = make_uint4(0u, 0u, 0u, 0u);
  LabelPlanes pl = {0u, 0u, 0u};
  unsigned hl[KT];
#pragma unroll
  for (int c = 0; c < KT; ++c) hl[c] = 0u;
  bool have_last = false;
  unsigned bad_last = 0u;
  uint4 nxt = ld(0);
  uint4 nxt2 = nf > 1 ? ld(1) : nxt;
  if (have_first) {
    last = ld(-1);
    have_last = true;
    pl = tc_planes(last);
    bad_last = tc_planes_invalid<KT>(last, pl);
    unsigned ind[KT];
    tc_hist<KT>(pl, ind, hl);                  // only used when bad_last == 0
  }
  auto step = [&](int p) {
    const uint4 cur = nxt;
    nxt = nxt2;
    if (p + 2 < nf) nxt2 = ld(p + 2);
    const LabelPlanes pc = tc_planes(cur);
    const unsigned bad_cur = CHECK_FRAMES ? tc_planes_invalid<KT>(cur, pc) : 0u;
    unsigned ind[KT], hc[KT];
    tc_hist<KT>(pc, ind, hc);
    if (have_last) {
      if ((bad_cur | bad_last) == 0u) {
        const unsigned ne = (pc.a ^ pl.a) | (pc.b ^ pl.b) | (pc.c ^ pl.c);      // labels differ
        unsigned eq_rest = 0u;
#pragma unroll
        for (int c = 1; c < KT; ++c) {
          const unsigned e = __popc(ind[c] & ~ne);
          cnt.I[c] += e;
          eq_rest += e;
        }
        cnt.I[0] += __popc(~ne & 0x33333333u) - eq_rest;
#pragma unroll
        for (int c = 0; c < KT; ++c) {
          cnt.O[c] += hc[c];
          cnt.T[c] += hl[c];
        }
      } else {
        const unsigned cw[4] = {cur.x, cur.y, cur.z, cur.w};
        const unsigned lw[4] = {last.x, last.y, last.z, last.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cnt.add_label((cw[w] >> (8 * i)) & 255u, (lw[w] >> (8 * i)) & 255u, ignore);
        }
      }
    }
    last = cur;
    pl = pc;
    bad_last = bad_cur;
#pragma unroll
    for (int c = 0; c < KT; ++c) hl[c] = hc[c];
    have_last = true;
  };
  if constexpr (NFC > 0) {
    static_for_tc<NFC>([&](auto p_) { step(decltype(p_)::value); });
  } else {
    for (int p = 0; p < nf; ++p) step(p);
  }
}

// The same from global memory: `first` (NULL: none), frames at frames + p * HW.
template <int KT>
__device__ __forceinline__ void tc_chain16(ClassCounts<KT>& cnt, const uint8_t* first, const uint8_t* frames, int nf,
                                           long long HW, int ignore) {
  tc_chain16_ld<KT, true>(cnt, first != nullptr, nf, ignore, [&](int p) {
    return p < 0 ? __ldg(reinterpret_cast<const uint4*>(first))
                 : __ldcs(reinterpret_cast<const uint4*>(frames + static_cast<long long>(p) * HW));
  });
}

}  // namespace fuvs
