// metric.cu — per-class intersection / union / target counts.
//
// fuvs_confusion       <- intersectionAndUnionGPU util/util.py:52-63 (torch.histc
//                         binning) and intersectionAndUnion util/util.py:36-47
//                         (np.histogram binning, closed last bin), as dispatched
//                         by BaseModel.compute_metrics base/foundation.py:333-344
// fuvs_temporal_counts <- the temporal-consistency loop flow/base.py:280-295
//
// Integer-only: per-thread packed 8-bit counters -> REDUX warp sums -> one
// shared-memory merge per block -> 3K 64-bit atomics per block.  No float
// atomics anywhere, so results are run-to-run identical.
#include <cstdlib>

#include "fuvs_common.cuh"
#include "temporal_fields.cuh"

namespace fuvs {

// bin index of a label value, or -1 if the histogram drops it
__device__ __forceinline__ int bin_of(long long v, int K, int np_bins) {
  if (np_bins) {                       // np.histogram(bins=arange(K+1)): [0,1) ... [K-1,K]
    if (v < 0 || v > K) return -1;
    return v == K ? K - 1 : static_cast<int>(v);
  }
  if (K == 1) return 0;   // histc(bins=1, min=0, max=0): "if min and max are both zero, the data's min and max are used" -> every value counts
  if (v < 0 || v > K - 1) return -1;                  // histc(bins=K, min=0, max=K-1) ignores outliers
  return static_cast<int>(v);
}

template <typename PT, typename TT, bool PACKED>
__global__ void __launch_bounds__(256)
confusion_kernel(PT* __restrict__ pred, const TT* __restrict__ target, long long N, int K, long long ignore,
                 int np_bins, int mutate, unsigned long long* __restrict__ counts) {
  __shared__ unsigned sh[PACKED ? 24 : 768];
  Hist3Packed hist;
  WarpTotals<8> tot;
  hist.clear();
  tot.clear();
  if (!PACKED) smem_hist_clear(sh);
  const PT ign_p = static_cast<PT>(ignore);
  const long long per_iter = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  const long long nround = ((N + per_iter - 1) / per_iter) * per_iter;
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x * 4; base < nround; base += per_iter) {
    PT o[4];
    TT t[4];
    bool live[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = base + j * blockDim.x + threadIdx.x;
      live[j] = i < N;
      if (live[j]) {
        o[j] = pred[i];
        t[j] = __ldcs(target + i);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!live[j]) continue;
      const long long tv = static_cast<long long>(t[j]);
      PT ov = o[j];
      if (tv == ignore) {                               // util/util.py:57 / :41
        ov = ign_p;
        if (mutate) pred[base + j * blockDim.x + threadIdx.x] = ign_p;
      }
      const long long ovl = static_cast<long long>(ov);
      const int bo = bin_of(ovl, K, np_bins), bt = bin_of(tv, K, np_bins);
      if (PACKED) {
        const unsigned long long fo = bo >= 0 ? (1ull << (8 * (bo & 7))) : 0ull;
        const unsigned long long ft = bt >= 0 ? (1ull << (8 * (bt & 7))) : 0ull;
        hist.O += fo;
        hist.T += ft;
        hist.I += (ovl == tv) ? fo : 0ull;
      } else {
        if (bo >= 0) atomicAdd(&sh[256 + bo], 1u);
        if (bt >= 0) atomicAdd(&sh[512 + bt], 1u);
        if (bo >= 0 && ovl == tv) atomicAdd(&sh[bo], 1u);
      }
    }
    if (PACKED) warp_accumulate<8>(hist, tot);
  }
  if (PACKED) {
    // block_flush_counts<8> walks 8 bins; only the first K rows exist in counts
    const int tid = threadIdx.x;
    if (tid < 24) sh[tid] = 0u;
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (tot.I[c]) atomicAdd(&sh[c], tot.I[c]);
        if (tot.O[c]) atomicAdd(&sh[8 + c], tot.O[c]);
        if (tot.T[c]) atomicAdd(&sh[16 + c], tot.T[c]);
      }
    }
    __syncthreads();
    if (tid < K) {
      const unsigned long long i = sh[tid], o = sh[8 + tid], t = sh[16 + tid];
      if (i) atomicAdd(counts + tid, i);
      if (o + t - i) atomicAdd(counts + K + tid, o + t - i);
      if (t) atomicAdd(counts + 2 * K + tid, t);
    }
  } else {
    smem_hist_flush(sh, counts, K);
  }
}

template <typename K>
static int persistent_grid(K kernel, long long work_items, int threads) {
  const long long need = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * blocks_per_sm(kernel, threads);
  long long g = need < cap ? need : cap;
  return static_cast<int>(g > 0 ? g : 1);
}

// ---------------------------------------------------------------------------
// Fast path of fuvs_confusion: torch.histc binning, 2 <= K <= 5, ignore_index outside [0,K) (the reference's GPU
// metric at its class counts, util/util.py:52-63).  The general kernel above reduces across the warp after every 4
// labels and loads one label per instruction: 40 us for 5 x 1080p whatever the dtypes (tools/metric_bench.py), i.e.
// issue-bound at 35 % of the HBM roofline for uint8 predictions against int64 targets.  Here a thread takes 16 labels
// per iteration with 128-bit loads, counts in three 32-bit registers with 6/8-bit class fields (FieldCounts) and the
// warp reduction happens once per kernel.
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Byte-domain counting of one group of 16 (prediction, target) labels for fuvs_confusion's fast path.  The per-label
// code below it works on 64-bit values (range checks, ignore test and three field updates: ~20 ALU instructions per
// label for int64 targets, which made the kernel ALU-bound at 0.63 of the HBM roofline).  Here int64 labels are first
// narrowed to bytes (group-wide OR of the high words and one PRMT per two labels); a group whose predictions are all
// < KT and whose targets are all < KT or == ignore_index then gets its counter fields from one multiply per word like
// fuvs_temporal_counts.  Returns false when the group needs the per-label path (out-of-range labels, an
// ignore_index that does not fit a byte).  igff: 0xff in the bytes whose target is ignore_index.
// ---------------------------------------------------------------------------
template <int KT>
__device__ __forceinline__ bool confusion_bytes16(const unsigned (&pw)[4], const unsigned (&tw)[4], bool ig_valid,
                                                  unsigned ig4, bool refuse_ignored, FieldCounts<KT>& cnt,
                                                  unsigned (&igff)[4], bool& any_ignored) {
  constexpr unsigned ADD = (0x80u - KT) * 0x01010101u;
  unsigned pbad = 0u, tbad = 0u, anyig = 0u;
  unsigned ig80[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pbad |= ((pw[j] & 0x7f7f7f7fu) + ADD) | pw[j];
    const unsigned tb = ((tw[j] & 0x7f7f7f7fu) + ADD) | tw[j];
    const unsigned x = tw[j] ^ ig4;
    const unsigned nz = ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x;          // bit 7 of a byte set iff the byte of x is non-zero
    ig80[j] = ig_valid ? (~nz & 0x80808080u) : 0u;
    tbad |= tb & ~ig80[j];
    anyig |= ig80[j];
  }
  if (((pbad | tbad) & 0x80808080u) != 0u) return false;
  any_ignored = anyig != 0u;
  if (refuse_ignored && any_ignored) return false;
  uint4 pq = make_uint4(pw[0], pw[1], pw[2], pw[3]), tq;
#pragma unroll
  for (int j = 0; j < 4; ++j) igff[j] = (ig80[j] >> 7) * 0xffu;
  tq = make_uint4(tw[0] & ~igff[0], tw[1] & ~igff[1], tw[2] & ~igff[2], tw[3] & ~igff[3]);   // ignored bytes -> 0 (masked below)
  unsigned fo[16], ft[16];
  const unsigned s_o = tc_fields<KT>(pq, fo);
  const unsigned s_t = tc_fields<KT>(tq, ft);
  if (!any_ignored) {
    cnt.accO += s_o;
    cnt.accT += s_t;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool keep = (ig80[j] & (0x80u << (8 * i))) == 0u;       // output[target == ignore] = ignore: nothing counts
        cnt.accO += keep ? fo[4 * j + i] : 0u;
        cnt.accT += keep ? ft[4 * j + i] : 0u;
      }
    }
  }
  // a prediction (< KT) never equals an ignored target byte (ignore_index is outside the classes): no mask needed
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned x = pw[j] ^ tw[j];
#pragma unroll
    for (int i = 0; i < 4; ++i) cnt.accI += ((x & (0xffu << (8 * i))) == 0u) ? fo[4 * j + i] : 0u;
  }
  return true;
}

template <typename T> struct Lab16;
template <> struct Lab16<uint8_t> {
  uint4 w;
  __device__ __forceinline__ void load(const uint8_t* p, bool stream) {
    w = stream ? __ldcs(reinterpret_cast<const uint4*>(p)) : *reinterpret_cast<const uint4*>(p);
  }
  __device__ __forceinline__ long long get(int i) const {
    const unsigned word = i < 4 ? w.x : i < 8 ? w.y : i < 12 ? w.z : w.w;
    return (word >> (8 * (i & 3))) & 255u;
  }
  __device__ __forceinline__ void set(int i, long long v) {
    unsigned& word = i < 4 ? w.x : i < 8 ? w.y : i < 12 ? w.z : w.w;
    word = (word & ~(255u << (8 * (i & 3)))) | ((static_cast<unsigned>(v) & 255u) << (8 * (i & 3)));
  }
  __device__ __forceinline__ void store(uint8_t* p) const { *reinterpret_cast<uint4*>(p) = w; }
  // interleaved mapping (see confusion_v16_kernel): label pair k of lane l sits at pair index 32 k + l of the warp's block
  __device__ __forceinline__ void load_il(const uint8_t* blk, int lane, bool stream) {
    const unsigned short* q = reinterpret_cast<const unsigned short*>(blk) + lane;
    unsigned h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = stream ? __ldcs(q + 32 * k) : q[32 * k];
    w = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
  }
  __device__ __forceinline__ void store_il(uint8_t* blk, int lane) const {
    unsigned short* q = reinterpret_cast<unsigned short*>(blk) + lane;
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) q[32 * k] = static_cast<unsigned short>(ws[k >> 1] >> (16 * (k & 1)));
  }
  __device__ __forceinline__ bool narrow(unsigned (&b)[4]) const { b[0] = w.x; b[1] = w.y; b[2] = w.z; b[3] = w.w; return true; }
  // byte-wise: bytes flagged in ff take the bytes of v4
  __device__ __forceinline__ void merge_bytes(const unsigned (&ff)[4], unsigned v4) {
    w.x = (w.x & ~ff[0]) | (v4 & ff[0]); w.y = (w.y & ~ff[1]) | (v4 & ff[1]);
    w.z = (w.z & ~ff[2]) | (v4 & ff[2]); w.w = (w.w & ~ff[3]) | (v4 & ff[3]);
  }
  __device__ __forceinline__ void from_bytes(const unsigned (&b)[4]) { w = make_uint4(b[0], b[1], b[2], b[3]); }
  static constexpr bool kByteMerge = true;
};
template <> struct Lab16<long long> {
  longlong2 w[8];
  __device__ __forceinline__ void load(const long long* p, bool stream) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      w[k] = stream ? __ldcs(reinterpret_cast<const longlong2*>(p) + k) : reinterpret_cast<const longlong2*>(p)[k];
  }
  __device__ __forceinline__ long long get(int i) const { return (i & 1) ? w[i >> 1].y : w[i >> 1].x; }
  __device__ __forceinline__ void set(int i, long long v) { if (i & 1) w[i >> 1].y = v; else w[i >> 1].x = v; }
  __device__ __forceinline__ void store(long long* p) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) reinterpret_cast<longlong2*>(p)[k] = w[k];
  }
  __device__ __forceinline__ void load_il(const long long* blk, int lane, bool stream) {
    const longlong2* q = reinterpret_cast<const longlong2*>(blk) + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = stream ? __ldcs(q + 32 * k) : q[32 * k];
  }
  __device__ __forceinline__ void store_il(long long* blk, int lane) const {
    longlong2* q = reinterpret_cast<longlong2*>(blk) + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) q[32 * k] = w[k];
  }
  // low bytes of the 16 values; false when a value does not fit a byte (negative, >= 256)
  __device__ __forceinline__ bool narrow(unsigned (&b)[4]) const {
    unsigned hi_or = 0u, lo_or = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long v0 = w[2 * j].x, v1 = w[2 * j].y, v2 = w[2 * j + 1].x, v3 = w[2 * j + 1].y;
      const unsigned l0 = static_cast<unsigned>(v0), l1 = static_cast<unsigned>(v1), l2 = static_cast<unsigned>(v2), l3 = static_cast<unsigned>(v3);
      hi_or |= static_cast<unsigned>(v0 >> 32) | static_cast<unsigned>(v1 >> 32) | static_cast<unsigned>(v2 >> 32) | static_cast<unsigned>(v3 >> 32);
      lo_or |= l0 | l1 | l2 | l3;
      b[j] = __byte_perm(__byte_perm(l0, l1, 0x0040), __byte_perm(l2, l3, 0x0040), 0x5410);
    }
    return hi_or == 0u && lo_or < 256u;
  }
  __device__ __forceinline__ void merge_bytes(const unsigned (&)[4], unsigned) {}
  __device__ __forceinline__ void from_bytes(const unsigned (&)[4]) {}
  static constexpr bool kByteMerge = false;
};

// IL (an int64 operand): a thread's 16 labels are the pairs 32 k + lane (k = 0..7) of its warp's 512-label block, so
// that every load instruction of the warp covers one contiguous 512-byte run.  With 128 contiguous bytes per thread
// each 128-bit load touched 32 different lines (8 instructions per line): the L1 tag stage, not HBM, set the pace
// (0.62 of the roofline whatever the ALU work).  Byte operands then come as eight 16-bit loads.
template <typename PT, typename TT, int KT, bool PIPE, bool IL>
__global__ void __launch_bounds__(256, PIPE ? 2 : 1)
confusion_v16_kernel(PT* __restrict__ pred, const TT* __restrict__ target, long long ngroups, long long ignore,
                     int mutate, unsigned long long* __restrict__ counts) {
  using FC = FieldCfg<KT>;
  __shared__ unsigned sh[24];
  FieldCounts<KT> cnt;
  cnt.init();
  int since_spill = 0;
  const int lane = threadIdx.x & 31;
  // unit of work: one 16-label group per thread, or (IL) one 512-label block per warp
  const long long stride = IL ? static_cast<long long>(gridDim.x) * (blockDim.x >> 5) : static_cast<long long>(gridDim.x) * blockDim.x;
  long long g = IL ? static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)
                   : static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto ld = [&](long long gi, Lab16<PT>& oo, Lab16<TT>& tt, bool stream) {
    if (IL) {
      oo.load_il(pred + gi * 512, lane, stream && mutate == 0);
      tt.load_il(target + gi * 512, lane, stream);
    } else {
      oo.load(pred + gi * 16, stream && mutate == 0);
      tt.load(target + gi * 16, stream);
    }
  };
  auto st = [&](long long gi, const Lab16<PT>& oo) {
    if (IL) oo.store_il(pred + gi * 512, lane); else oo.store(pred + gi * 16);
  };
  // software pipeline: the loads of the next group are in flight while this one is counted (a thread's ~250
  // instructions per group used to start only after its 9 loads had come back from DRAM)
  Lab16<PT> o_nxt;
  Lab16<TT> t_nxt;
  if (PIPE && g < ngroups) ld(g, o_nxt, t_nxt, true);
  for (; g < ngroups; g += stride) {
    Lab16<PT> o;
    Lab16<TT> t;
    if (PIPE) {
      o = o_nxt;
      t = t_nxt;
      if (g + stride < ngroups) ld(g + stride, o_nxt, t_nxt, true);
    } else {
      ld(g, o, t, true);
    }
    bool changed = false;
    {
      unsigned pw[4], tw[4], igff[4];
      bool any_ignored = false;
      const bool ig_valid = ignore >= 0 && ignore <= 255;
      const unsigned ig4 = (static_cast<unsigned>(ignore) & 255u) * 0x01010101u;
      // in-place substitution on int64 predictions stays on the per-label path (a rewrite of 128 bytes per group)
      bool counted = false;
      if (o.narrow(pw) & t.narrow(tw))
        counted = confusion_bytes16<KT>(pw, tw, ig_valid, ig4, mutate && !Lab16<PT>::kByteMerge, cnt, igff, any_ignored);
      if (counted) {
        if (mutate && any_ignored) {
          Lab16<PT> m;
          m.from_bytes(pw);
          m.merge_bytes(igff, ig4);
          st(g, m);
        }
        since_spill += 16;
        if (since_spill + 16 > FC::CAP) {
          cnt.spill();
          since_spill = 0;
        }
        continue;
      }
      // rare: out-of-range labels.  The 64-bit values are read again (L1/L2 hits) so that their 64 registers are not
      // held across the byte path above — that cost two thirds of the occupancy.
      if (!PIPE) ld(g, o, t, false);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const long long tv = t.get(i);
      long long ov = o.get(i);
      if (tv == ignore) {                                   // output[target == ignore] = ignore (util/util.py:57)
        ov = static_cast<long long>(static_cast<PT>(ignore));
        if (mutate) { o.set(i, ov); changed = true; }
      }
      const unsigned ft = (tv >= 0 && tv < KT) ? FieldCounts<KT>::field(static_cast<int>(tv)) : 0u;
      const unsigned fo = (ov >= 0 && ov < KT) ? FieldCounts<KT>::field(static_cast<int>(ov)) : 0u;
      cnt.accO += fo;
      cnt.accT += ft;
      cnt.accI += (ov == tv) ? fo : 0u;
    }
    if (mutate && changed) st(g, o);
    since_spill += 16;
    if (since_spill + 16 > FC::CAP) {
      cnt.spill();
      since_spill = 0;
    }
  }
  cnt.finish(sh, counts, KT);
}

template <typename PT, typename TT>
static constexpr bool conf_interleaved() { return sizeof(PT) == 8 || sizeof(TT) == 8; }
// labels one work unit of the fast path covers (a thread's group, or a warp's block with int64 operands)
template <typename PT, typename TT>
static constexpr int conf_unit() { return conf_interleaved<PT, TT>() ? 512 : 16; }

template <typename PT, typename TT>
static int launch_confusion_v16(void* pred, const void* target, long long nunits, int K, long long ignore, int mutate,
                                long long* counts, cudaStream_t st) {
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  const int threads = 256;
  constexpr bool IL = conf_interleaved<PT, TT>();
  const long long work_threads = IL ? nunits * 32 : nunits;
  // software pipelining only for byte labels: with int64 operands two groups in flight do not fit the registers
  constexpr bool pipe = true;
#define FUVS_CF16(KT_)                                                                                         \
  {                                                                                                            \
    bool launched = false;                                                                                     \
    if constexpr (!IL) {                                                                                       \
      if (pipe) {                                                                                              \
        const int grid = persistent_grid(confusion_v16_kernel<PT, TT, KT_, true, IL>, work_threads, threads);  \
        confusion_v16_kernel<PT, TT, KT_, true, IL><<<grid, threads, 0, st>>>(static_cast<PT*>(pred),          \
                                                                             static_cast<const TT*>(target), nunits, ignore, mutate, cu); \
        launched = true;                                                                                       \
      }                                                                                                        \
    }                                                                                                          \
    if (!launched) {                                                                                           \
      const int grid = persistent_grid(confusion_v16_kernel<PT, TT, KT_, false, IL>, work_threads, threads);   \
      confusion_v16_kernel<PT, TT, KT_, false, IL><<<grid, threads, 0, st>>>(static_cast<PT*>(pred),           \
                                                                            static_cast<const TT*>(target), nunits, ignore, mutate, cu); \
    }                                                                                                          \
  }
  switch (K) {
    case 2: FUVS_CF16(2) break;
    case 3: FUVS_CF16(3) break;
    case 4: FUVS_CF16(4) break;
    default: FUVS_CF16(5) break;
  }
#undef FUVS_CF16
  return check_launch("fuvs_confusion(v16)");
}

template <int VEC, int KT>
__global__ void __launch_bounds__(256)
temporal_counts_kernel(const uint8_t* __restrict__ labels, int n, long long HW, const uint8_t* __restrict__ tc_prev,
                       int K, int ignore, unsigned long long* __restrict__ counts) {
  // KT = 8: packed path (K <= 8); KT = 0: shared-memory histogram (K <= 256)
  __shared__ unsigned sh[KT ? 24 : 768];
  Hist3Packed hist;
  WarpTotals<8> tot;
  hist.clear();
  tot.clear();
  if (!KT) smem_hist_clear(sh);
  const long long nvec = HW / VEC;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nround = (nvec + 31) & ~31ll;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nround; v += stride) {
    if (v < nvec) {
      const long long pix = v * VEC;
      LVec<VEC> last;
      bool have_last = false;
      if (tc_prev) {
        last.load(tc_prev + pix);
        have_last = true;
      }
      for (int p = 0; p < n; ++p) {
        LVec<VEC> lab;
        lab.load(labels + static_cast<long long>(p) * HW + pix);
        if (have_last) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if (KT) hist.add(lab.v[i], last.v[i], ignore, K);
            else smem_hist_add(sh, lab.v[i], last.v[i], ignore, K);
          }
        }
        last = lab;
        have_last = true;
      }
    }
    if (KT) warp_accumulate<8>(hist, tot);
  }
  if (KT) block_flush_counts<8>(tot, sh, counts, K);
  else smem_hist_flush(sh, counts, K);
}

// Fast path of the temporal-consistency counts: K <= 5, ignore outside [0,K), HW % 16 == 0.  16 pixels per thread
// (one 128-bit load per frame), bit-plane counting (temporal_fields.cuh: ~3 ALU instructions per label and frame; the
// first version spent ~14, the field-packed second one ~9 and 10 us for 5 x 1080p), one REDUX pass per warp at the end.
// Frames with out-of-range labels (a caller's own label maps, ignore_index) take the per-label path.
template <int KT>
__global__ void __launch_bounds__(256)
temporal_counts_v16_kernel(const uint8_t* __restrict__ labels, int n, long long HW, const uint8_t* __restrict__ tc_prev,
                           int ignore, unsigned long long* __restrict__ counts) {
  __shared__ unsigned sh[24];
  // programmatic dependent launch (see dense_strip.cu): start while the producer of the label maps drains, let the
  // next kernel of the stream do the same, touch memory only after the producer has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  ClassCounts<KT> cnt;
  cnt.init();
  const long long nvec = HW >> 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const long long pix = v << 4;
    if (n == 5) {
      // the reference's interval (k = 5): all six label words are requested before the first is counted
      uint4 q[6];
      q[0] = tc_prev ? __ldg(reinterpret_cast<const uint4*>(tc_prev + pix)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int p = 0; p < 5; ++p) q[p + 1] = __ldcs(reinterpret_cast<const uint4*>(labels + p * HW + pix));
      tc_chain16_ld<KT, true, 5>(cnt, tc_prev != nullptr, 5, ignore, [&](int p) { return q[p + 1]; });
    } else {
      tc_chain16<KT>(cnt, tc_prev ? tc_prev + pix : nullptr, labels + pix, n, HW, ignore);
    }
  }
  cnt.finish(sh, counts, KT);
}


int launch_temporal_counts(const uint8_t* labels, int n, long long HW, const uint8_t* tc_prev, int K,
                           int ignore_index, long long* counts, cudaStream_t st) {
  if (!labels || !counts || n < 1 || HW < 0 || K < 1 || K > 256)
    return set_error(FUVS_EINVAL, "temporal_counts: bad arguments n=%d HW=%lld K=%d", n, HW, K);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "temporal_counts: n=%d exceeds %d", n, FUVS_MAX_FRAMES);
  if (HW == 0 || (n == 1 && !tc_prev)) return FUVS_OK;
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  const bool vec4 = (HW % 4 == 0) && aligned4(labels) && (!tc_prev || aligned4(tc_prev));
  const int threads = 256;
#define FUVS_TC(V, KT_)                                                                                 \
  {                                                                                                     \
    const int grid = persistent_grid(temporal_counts_kernel<V, KT_>, HW / V, threads);                  \
    temporal_counts_kernel<V, KT_><<<grid, threads, 0, st>>>(labels, n, HW, tc_prev, K, ignore_index, cu); \
  }
  const bool v16 = (HW % 16 == 0) && aligned16(labels) && (!tc_prev || aligned16(tc_prev)) &&
                   (ignore_index < 0 || ignore_index >= K);
  if (v16 && K <= 5) {
#define FUVS_TC16(KT_)                                                                                   \
  {                                                                                                      \
    const int grid = persistent_grid(temporal_counts_v16_kernel<KT_>, HW / 16, threads);                 \
    cudaLaunchConfig_t cfg = {};                                                                         \
    cfg.gridDim = dim3(grid);                                                                            \
    cfg.blockDim = dim3(threads);                                                                        \
    cfg.stream = st;                                                                                     \
    cudaLaunchAttribute attr[1];                                                                         \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                     \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                              \
    cfg.attrs = attr;                                                                                    \
    cfg.numAttrs = 1;                                                                                    \
    cudaLaunchKernelEx(&cfg, temporal_counts_v16_kernel<KT_>, labels, n, HW, tc_prev, ignore_index, cu); \
  }
    switch (K) {
      case 1: FUVS_TC16(1) break;
      case 2: FUVS_TC16(2) break;
      case 3: FUVS_TC16(3) break;
      case 4: FUVS_TC16(4) break;
      default: FUVS_TC16(5) break;
    }
#undef FUVS_TC16
  } else if (K <= 8) {
    if (vec4) FUVS_TC(4, 8) else FUVS_TC(1, 8)
  } else {
    if (vec4) FUVS_TC(4, 0) else FUVS_TC(1, 0)
  }
#undef FUVS_TC
  return check_launch("fuvs_temporal_counts");
}

template <typename PT, typename TT>
static int launch_confusion(void* pred, const void* target, long long N, int K, int ignore, int flags,
                            long long* counts, cudaStream_t st) {
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  const int np_bins = flags & 1, mutate = (flags & FUVS_MUTATE_PRED) ? 1 : 0;
  const int threads = 256;
  // fast path: histc binning, 2 <= K <= 5, ignore outside the classes, 16-byte aligned label arrays; the
  // N % 16 tail goes through the general kernel below
  constexpr int UNIT = conf_unit<PT, TT>();
  if (!np_bins && K >= 2 && K <= 5 && (ignore < 0 || ignore >= K) && N >= UNIT && aligned16(pred) && aligned16(target)) {
    const long long nunits = N / UNIT;
    if (int e = launch_confusion_v16<PT, TT>(pred, target, nunits, K, ignore, mutate, counts, st)) return e;
    const long long done = nunits * UNIT;
    if (done == N) return FUVS_OK;
    pred = static_cast<PT*>(pred) + done;
    target = static_cast<const TT*>(target) + done;
    N -= done;
  }
  if (K <= 8) {
    const int grid = persistent_grid(confusion_kernel<PT, TT, true>, (N + 3) / 4, threads);
    confusion_kernel<PT, TT, true><<<grid, threads, 0, st>>>(static_cast<PT*>(pred), static_cast<const TT*>(target), N,
                                                             K, ignore, np_bins, mutate, cu);
  } else {
    const int grid = persistent_grid(confusion_kernel<PT, TT, false>, (N + 3) / 4, threads);
    confusion_kernel<PT, TT, false><<<grid, threads, 0, st>>>(static_cast<PT*>(pred), static_cast<const TT*>(target),
                                                              N, K, ignore, np_bins, mutate, cu);
  }
  return check_launch("fuvs_confusion");
}

}  // namespace fuvs

extern "C" int fuvs_confusion(void* pred, int pred_is_i64, const void* target, int target_is_i64, long long N, int K,
                              int ignore_index, int flags, long long* counts, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (N < 0 || K < 1 || K > 256 || !counts) return set_error(FUVS_EINVAL, "confusion: bad arguments N=%lld K=%d", N, K);
  if (N == 0) return FUVS_OK;
  if (!pred || !target) return set_error(FUVS_EINVAL, "confusion: NULL label pointer");
  if ((pred_is_i64 && !aligned8(pred)) || (target_is_i64 && !aligned8(target)))
    return set_error(FUVS_EALIGN, "confusion: int64 labels must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pred_is_i64) {
    if (target_is_i64) return launch_confusion<long long, long long>(pred, target, N, K, ignore_index, flags, counts, st);
    return launch_confusion<long long, uint8_t>(pred, target, N, K, ignore_index, flags, counts, st);
  }
  if (target_is_i64) return launch_confusion<uint8_t, long long>(pred, target, N, K, ignore_index, flags, counts, st);
  return launch_confusion<uint8_t, uint8_t>(pred, target, N, K, ignore_index, flags, counts, st);
}

extern "C" int fuvs_temporal_counts(const uint8_t* labels, int n, long long HW, const uint8_t* tc_prev, int K,
                                    int ignore_index, long long* counts, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  return launch_temporal_counts(labels, n, HW, tc_prev, K, ignore_index, counts, static_cast<cudaStream_t>(stream));
}
