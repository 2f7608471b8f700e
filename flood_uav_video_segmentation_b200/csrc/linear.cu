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
#include <cstdlib>

#include "fuvs_common.cuh"
#include "pix4.cuh"

// Developer build -DFUVS_STRIP_ASSERT / -DFUVS_STRIP_JITTER (tools/strip_asserts.sh, see dense_strip.cu): the bulk kernel's
// barrier-free stage hand-over checks its counters and indices itself and runs with pseudo-random per-warp stalls.
#ifdef FUVS_STRIP_ASSERT
#include <cassert>
#define LIN_ASSERT(x) assert(x)
#else
#define LIN_ASSERT(x) ((void)0)
#endif
#ifdef FUVS_STRIP_JITTER
#define LIN_JITTER(i_, salt)                                                                            \
  do {                                                                                                  \
    unsigned h_ = (static_cast<unsigned>(i_) * 2654435761u) ^ ((threadIdx.x >> 5) * 40503u) ^           \
                  (blockIdx.x * 2246822519u) ^ (salt);                                                  \
    h_ ^= h_ >> 15; h_ *= 2246822519u; h_ ^= h_ >> 13;                                                  \
    if ((h_ % 3u) == 0u) {                                                                              \
      const long long t_ = clock64() + ((h_ >> 8) & 16383u);                                            \
      while (clock64() < t_) {}                                                                         \
    }                                                                                                   \
  } while (0)
#else
#define LIN_JITTER(i_, salt) ((void)0)
#endif

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

// ---------------------------------------------------------------------------
// Fast path: 4 pixels per thread (HW % 4 == 0), C in {2,3,4,5,8}, ignore_index
// outside [0,C).  Instruction diet (the first version was issue-bound at 1491
// instructions per thread, profiles/r01_ncu_linear_v1.txt):
//   * blends on packed FP32x2 (3 FMUL2/FFMA2 per pixel pair and class);
//   * arg-max without NaN handling when all 2*C*4 inputs are finite (then no
//     blend can be NaN); a thread that sees Inf/NaN takes the exact slow path;
//   * counts in three 32-bit registers with FW-bit fields (one shift + three
//     adds per label), spilled to per-thread 32-bit totals before a field can
//     overflow and REDUX-reduced once at the end of the kernel.
// ---------------------------------------------------------------------------
template <int CT, int NP, bool COUNTS, bool NANSAFE, bool LOGITS = true>
__device__ __forceinline__ void linear_frames(const u64 (&a)[CT][NP], const u64 (&b)[CT][NP], long long HW,
                                              long long pix, int n, uint8_t* __restrict__ labels,
                                              float* __restrict__ logits, bool have_tc, typename PixIO<NP>::LabelWord tc_word,
                                              int ignore_index, const BlendWeights& wts, u64 one2,
                                              FieldCounts<CT>& cnt) {
  using FC = FieldCfg<CT>;
  constexpr int NPX = 2 * NP;
  if constexpr (!NANSAFE) {
    // ---- no value can be NaN: arg-max, counter fields and label bytes in the float domain (pix4.cuh)
    static_assert(CT >= 2 && CT <= 64 && FC::FW * (CT - 1) < 32, "float-domain index packing");
    const u64 magic2 = pack2(8388608.f, 8388608.f);                              // 2^23
    const u64 fw2 = pack2(static_cast<float>(FC::FW), static_cast<float>(FC::FW));
    u64 last[NP];                      // float indices of the previous frame
    unsigned s_prev = 0u;              // sum of its counter fields: the T term of the next frame in one add
    // counter fields of NPX labels: fld[i] = 1 << (FW * idx_i)
    auto fields = [&](const u64 (&idx)[NP], unsigned (&fld)[NPX]) {
#pragma unroll
      for (int h = 0; h < NP; ++h) {
        float ml, mh;
        unpack2(fma2_rn(idx[h], fw2, magic2), ml, mh);
        fld[2 * h] = one_shl_wrap(__float_as_uint(ml));
        fld[2 * h + 1] = one_shl_wrap(__float_as_uint(mh));
      }
    };
    // ---- frame 0: the unblended key frame (flow/model.py:195-197)
    {
      u64 idx[NP];
#pragma unroll
      for (int h = 0; h < NP; ++h) {
        u64 x[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) x[c] = a[c][h];
        idx[h] = argmax2f<CT>(x);
      }
      if (LOGITS && logits) {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          float x[NPX];
#pragma unroll
          for (int h = 0; h < NP; ++h) unpack2(a[c][h], x[2 * h], x[2 * h + 1]);
          PixIO<NP>::store(logits + c * HW + pix, x);
        }
      }
      const typename PixIO<NP>::LabelWord w = PixIO<NP>::label_word(idx);
      if (labels) PixIO<NP>::store_label_word(labels + pix, w);
      if (COUNTS) {
        unsigned fld[NPX];
        fields(idx, fld);
        if (have_tc) {                      // tc_word: the callers load it long before it is needed (a DRAM round trip)
#pragma unroll
          for (int i = 0; i < NPX; ++i) {
            const int tl = static_cast<int>((tc_word >> (8 * i)) & 255u), lab = static_cast<int>((w >> (8 * i)) & 255u);
            const unsigned ft = (tl < CT) ? FieldCounts<CT>::field(tl) : 0u;
            // output[target == ignore] = ignore, and ignore is outside [0,CT) on this path: nothing of this pixel counts
            const unsigned fo = (tl == ignore_index) ? 0u : fld[i];
            cnt.add(lab, fo, tl, ft);
          }
        }
#pragma unroll
        for (int i = 0; i < NPX; ++i) s_prev += fld[i];
#pragma unroll
        for (int h = 0; h < NP; ++h) last[h] = idx[h];
      }
    }
    // ---- frames 1..n-1: fl(fl(w0*a) + fl(w1*b))  (flow/model.py:233-237)
    int since_spill = 1;
    for (int p = 1; p < n; ++p) {
      const u64 w0 = pack2(wts.w0[p], wts.w0[p]), w1 = pack2(wts.w1[p], wts.w1[p]);
      float* lg = (LOGITS && logits) ? logits + (static_cast<long long>(p) * CT) * HW + pix : nullptr;
      u64 idx[NP];
      u64 x[NP][CT];
#pragma unroll
      for (int h = 0; h < NP; ++h) {
#pragma unroll
        for (int c = 0; c < CT; ++c) x[h][c] = blend2x2(w0, a[c][h], w1, b[c][h], one2);
        idx[h] = argmax2f<CT>(x[h]);
      }
      if (LOGITS && lg) {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          float xs[NPX];
#pragma unroll
          for (int h = 0; h < NP; ++h) unpack2(x[h][c], xs[2 * h], xs[2 * h + 1]);
          PixIO<NP>::store(lg + c * HW, xs);
        }
      }
      if (labels) PixIO<NP>::store_label_word(labels + static_cast<long long>(p) * HW + pix, PixIO<NP>::label_word(idx));
      if (COUNTS) {
        unsigned fld[NPX];
        fields(idx, fld);
        unsigned s_cur = 0u;
#pragma unroll
        for (int h = 0; h < NP; ++h) {
          float il, ih, ll, lh;
          unpack2(idx[h], il, ih);
          unpack2(last[h], ll, lh);
          cnt.accI += (il == ll) ? fld[2 * h] : 0u;
          cnt.accI += (ih == lh) ? fld[2 * h + 1] : 0u;
          s_cur += fld[2 * h] + fld[2 * h + 1];
          last[h] = idx[h];
        }
        cnt.accO += s_cur;               // this frame as the output ...
        cnt.accT += s_prev;              // ... against the previous frame as the target (flow/base.py:280-295)
        s_prev = s_cur;
        if (++since_spill >= FC::CAP / NPX) {
          cnt.spill();
          since_spill = 0;
        }
      }
    }
    if (COUNTS) cnt.spill();
    return;
  }
  int last[NPX];
  // ---- frame 0: the unblended key frame (flow/model.py:195-197)
  {
    float x[CT][NPX];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
#pragma unroll
      for (int h = 0; h < NP; ++h) unpack2(a[c][h], x[c][2 * h], x[c][2 * h + 1]);
      if (LOGITS && logits) PixIO<NP>::store(logits + c * HW + pix, x[c]);
    }
    int lab[NPX];
    argmaxN<CT, NPX, NANSAFE>(x, lab);
    if (labels) PixIO<NP>::store_labels(labels + pix, lab);
    if (COUNTS) {
      if (have_tc) {                      // tc_word: the callers load it long before it is needed (a DRAM round trip)
        const typename PixIO<NP>::LabelWord t = tc_word;
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
          const int tl = static_cast<int>((t >> (8 * i)) & 255u);
          const unsigned ft = (tl < CT) ? FieldCounts<CT>::field(tl) : 0u;
          // output[target == ignore] = ignore, and ignore is outside [0,CT) on this path: nothing of this pixel counts
          const unsigned fo = (tl == ignore_index) ? 0u : FieldCounts<CT>::field(lab[i]);
          cnt.add(lab[i], fo, tl, ft);
        }
      }
#pragma unroll
      for (int i = 0; i < NPX; ++i) last[i] = lab[i];
    }
  }
  // ---- frames 1..n-1: fl(fl(w0*a) + fl(w1*b))  (flow/model.py:233-237)
  int since_spill = 1;
  for (int p = 1; p < n; ++p) {
    const u64 w0 = pack2(wts.w0[p], wts.w0[p]), w1 = pack2(wts.w1[p], wts.w1[p]);
    float* lg = (LOGITS && logits) ? logits + (static_cast<long long>(p) * CT) * HW + pix : nullptr;
    float x[CT][NPX];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
#pragma unroll
      for (int h = 0; h < NP; ++h) unpack2(blend2x2(w0, a[c][h], w1, b[c][h], one2), x[c][2 * h], x[c][2 * h + 1]);
      if (LOGITS && lg) PixIO<NP>::store(lg + c * HW, x[c]);
    }
    int lab[NPX];
    argmaxN<CT, NPX, NANSAFE>(x, lab);
    if (labels) PixIO<NP>::store_labels(labels + static_cast<long long>(p) * HW + pix, lab);
    if (COUNTS) {
#pragma unroll
      for (int i = 0; i < NPX; ++i) {
        cnt.add(lab[i], FieldCounts<CT>::field(lab[i]), last[i], FieldCounts<CT>::field(last[i]));
        last[i] = lab[i];
      }
      if (++since_spill >= FC::CAP / NPX) {
        cnt.spill();
        since_spill = 0;
      }
    }
  }
  if (COUNTS) cnt.spill();
}

template <int CT, int NP, bool COUNTS>
__global__ void __launch_bounds__(256, NP == 2 ? 2 : 4)
linear_blend_argmax_v4_kernel(const float* __restrict__ prev, const float* __restrict__ next,
                              long long HW, int n,
                              uint8_t* __restrict__ labels, float* __restrict__ logits,
                              const uint8_t* __restrict__ tc_prev,
                              unsigned long long* __restrict__ counts, int ignore_index,
                              const BlendWeights wts, float one) {
  __shared__ unsigned sh[24];
  constexpr int NPX = 2 * NP;
  const long long nvec = HW / NPX;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const u64 one2 = pack2(one, one);
  const float zero = __fsub_rn(one, one);          // run-time 0 (see the note on ptxas in fuvs_common.cuh)
  const u64 zero2 = pack2(zero, zero);
  FieldCounts<CT> cnt;
  cnt.init();

  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const long long pix = v * NPX;
    u64 a[CT][NP], b[CT][NP];
    const bool have_tc = COUNTS && tc_prev != nullptr;
    const typename PixIO<NP>::LabelWord tc_word = have_tc ? PixIO<NP>::load_labels(tc_prev + pix) : 0u;
#pragma unroll
    for (int c = 0; c < CT; ++c) PixIO<NP>::load(prev + c * HW + pix, a[c]);
    if (n > 1) {
#pragma unroll
      for (int c = 0; c < CT; ++c) PixIO<NP>::load(next + c * HW + pix, b[c]);
    } else {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
#pragma unroll
        for (int h = 0; h < NP; ++h) b[c][h] = zero2;
      }
    }
    // x*0 is 0 for finite x and NaN for Inf/NaN: one FFMA2 per pixel pair detects non-finite inputs
    u64 probe = zero2;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
#pragma unroll
      for (int h = 0; h < NP; ++h) {
        probe = fma2_rn(a[c][h], zero2, probe);
        probe = fma2_rn(b[c][h], zero2, probe);
      }
    }
    float pr0, pr1;
    unpack2(probe, pr0, pr1);
    if ((pr0 == pr0) && (pr1 == pr1))
      linear_frames<CT, NP, COUNTS, false>(a, b, HW, pix, n, labels, logits, have_tc, tc_word, ignore_index, wts, one2, cnt);
    else
      linear_frames<CT, NP, COUNTS, true>(a, b, HW, pix, n, labels, logits, have_tc, tc_word, ignore_index, wts, one2, cnt);
  }
  if (COUNTS) cnt.finish(sh, counts, CT);
}

// ---------------------------------------------------------------------------
// Bulk-copy pipelined variant (HW % 4 == 0, CT <= 5).  The register-load kernel above leaves the SM idle while its
// 16 warps wait ~1.5 us for their 10 global loads (issue slots 56 % used, DRAM 38 %: profiles/r01_ncu_linear_v2.txt).
// Here one thread per CTA streams the 2*CT channel rows of the next 2 048-pixel tile into shared memory with
// cp.async.bulk (UBLKCP, completion on an mbarrier) while all 512 threads compute the current tile; operands are
// copied from shared memory to registers with 128-bit LDS and the buffer is released straight away (the last warp to
// release re-arms it), so a tile's HBM latency is hidden behind one to two tiles of arithmetic.
// The per-pixel arithmetic is linear_frames<> — the same code as the register-load kernel.
// ---------------------------------------------------------------------------
constexpr int BULK_THREADS = 512;
constexpr int BULK_TILE = BULK_THREADS * 4;        // pixels per tile
constexpr int BULK_STAGES = 2;
constexpr int BULK_NQ_MAX = 4;                     // independent slice pipelines per CTA (NQ = 4: 4 warps, 512 pixels each)

__device__ __forceinline__ uint32_t lin_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool lin_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void lin_wait(uint32_t bar, uint32_t parity) {
  if (lin_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!lin_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();          // a fault must not hang the GPU
  }
}

// BULK_NQ slice pipelines per CTA: slice q = warps (16/NQ) q .. owns pixels [2048/NQ q, ...) of every tile, with its own
// mbarrier per stage and its own release counter.  NQ = 1 (one barrier per stage, 8 KB bulk copies) is the default;
// NQ = 4 lets a quarter start when its own 20 KB have landed and refill as soon as its four warps hold their operands,
// but its 2 KB bulk copies cost more than the decoupling gains: 24.0 vs 22.5 us per interval (FUVS_LINEAR_NQ=4 keeps
// the variant reachable for re-measurement).
template <int CT, bool COUNTS, bool LOGITS, int BULK_NQ, int NP>
__global__ void __launch_bounds__(BULK_TILE / (2 * NP), 1)
linear_blend_argmax_bulk_kernel(const float* __restrict__ prev, const float* __restrict__ next,
                                long long HW, int n,
                                uint8_t* __restrict__ labels, float* __restrict__ logits,
                                const uint8_t* __restrict__ tc_prev,
                                unsigned long long* __restrict__ counts, int ignore_index,
                                const BlendWeights wts, float one, int pdl) {
  extern __shared__ __align__(128) unsigned char lin_smem[];
  __shared__ unsigned sh[24];
  constexpr int NPX = 2 * NP;                              // pixels per thread: 4 (512 threads) or 2 (1024 threads)
  constexpr int THREADS = BULK_TILE / NPX;
  constexpr int BULK_QPX = BULK_TILE / BULK_NQ;
  constexpr int BULK_QWARPS = THREADS / 32 / BULK_NQ;
  const int nplanes = (n > 1) ? 2 * CT : CT;
  float* stage_base = reinterpret_cast<float*>(lin_smem);                               // [STAGES][2*CT][BULK_TILE]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(lin_smem + static_cast<size_t>(BULK_STAGES) * 2 * CT * BULK_TILE * 4);
  unsigned* done = reinterpret_cast<unsigned*>(bars + BULK_STAGES * BULK_NQ);           // [STAGES][NQ]
  const int tid = threadIdx.x, lane = tid & 31;
  const int q = tid / (BULK_QPX / NPX);                                                   // this thread's quarter
  const long long ntiles = (HW + BULK_TILE - 1) / BULK_TILE;
  const u64 one2 = pack2(one, one);
  const float zero = __fsub_rn(one, one);
  const u64 zero2 = pack2(zero, zero);

  if (tid == 0) {
    for (int s = 0; s < BULK_STAGES * BULK_NQ; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lin_smem_u32(&bars[s])), "r"(1) : "memory");
      done[s] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // programmatic dependent launch: this grid may have been scheduled while the previous kernel of the stream drains
  // (its CTAs take an SM as soon as one is free); nothing in global memory is touched before the wait
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

  // one thread of quarter q: stream the channel rows of the quarter's slice of this CTA's i-th tile into stage i % STAGES
  auto issue = [&](long long i) {
    const long long t = blockIdx.x + i * gridDim.x;
    if (t >= ntiles) return;
    const int s = static_cast<int>(i % BULK_STAGES);
    const long long pix0 = t * BULK_TILE + q * BULK_QPX;
    const long long rem = HW - pix0;
    if (rem <= 0) return;                                                                      // nobody waits for it
    const unsigned bytes = static_cast<unsigned>((rem < BULK_QPX ? rem : BULK_QPX) * 4);     // multiple of 16: HW % 4 == 0
    const uint32_t bar = lin_smem_u32(&bars[s * BULK_NQ + q]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * nplanes) : "memory");
    float* sb = stage_base + static_cast<size_t>(s) * 2 * CT * BULK_TILE + q * BULK_QPX;
    for (int p = 0; p < nplanes; ++p) {
      const float* g = (p < CT ? prev + p * HW : next + (p - CT) * HW) + pix0;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(lin_smem_u32(sb + static_cast<size_t>(p) * BULK_TILE)), "l"(g), "r"(bytes), "r"(bar) : "memory");
    }
  };
  if ((tid & (BULK_QPX / NPX - 1)) == 0) {
    for (int i = 0; i < BULK_STAGES; ++i) issue(i);
  }

  FieldCounts<CT> cnt;
  cnt.init();
  // previous interval's last label map (temporal-consistency target of frame 0): 4 labels per thread and tile, loaded
  // one tile ahead — consumed straight after a load it stalled every warp for a DRAM round trip per tile
  // (long-scoreboard 2.2 warps per issue, profiles/r01_ncu_linear_final.txt)
  const bool have_tc = COUNTS && tc_prev != nullptr;
  using LabelWord = typename PixIO<NP>::LabelWord;
  auto tc_load = [&](long long tile) -> LabelWord {
    const long long px = tile * BULK_TILE + tid * NPX;
    return (have_tc && px < HW) ? PixIO<NP>::load_labels(tc_prev + px) : LabelWord(0);
  };
  LabelWord tc_word = tc_load(blockIdx.x);
  for (long long i = 0;; ++i) {
    const long long t = blockIdx.x + i * gridDim.x;
    if (t >= ntiles) break;
    if (t * BULK_TILE + q * BULK_QPX >= HW) break;              // this quarter of the last tile is past the end: no load was issued
    const LabelWord tc_next = tc_load(t + gridDim.x);
    const int s = static_cast<int>(i % BULK_STAGES);
    LIN_JITTER(i, 0x9e3779b9u);
    LIN_ASSERT(s >= 0 && s < BULK_STAGES && q >= 0 && q < BULK_NQ);
    lin_wait(lin_smem_u32(&bars[s * BULK_NQ + q]), static_cast<uint32_t>((i / BULK_STAGES) & 1));
    const float* sb = stage_base + static_cast<size_t>(s) * 2 * CT * BULK_TILE + tid * NPX;
    const long long pix = t * BULK_TILE + tid * NPX;
    const bool live = pix < HW;                                  // HW % 4 == 0: a thread's 4 pixels are all in or all out
    u64 a[CT][NP], b[CT][NP];
    auto lds = [&](const float* p, u64 (&d)[NP]) {
      if constexpr (NP == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p), u = *(reinterpret_cast<const float4*>(p) + 1);
        d[0] = pack2(v.x, v.y);
        d[1] = pack2(v.z, v.w);
        d[2] = pack2(u.x, u.y);
        d[3] = pack2(u.z, u.w);
      } else if constexpr (NP == 2) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        d[0] = pack2(v.x, v.y);
        d[1] = pack2(v.z, v.w);
      } else {
        const float2 v = *reinterpret_cast<const float2*>(p);
        d[0] = pack2(v.x, v.y);
      }
    };
#pragma unroll
    for (int c = 0; c < CT; ++c) lds(sb + static_cast<size_t>(c) * BULK_TILE, a[c]);
    if (n > 1) {
#pragma unroll
      for (int c = 0; c < CT; ++c) lds(sb + static_cast<size_t>(CT + c) * BULK_TILE, b[c]);
    } else {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
#pragma unroll
        for (int h = 0; h < NP; ++h) b[c][h] = zero2;
      }
    }
    // operands are in registers: hand the slice back; the last warp of the quarter refills it with tile i + STAGES
    LIN_JITTER(i, 0x85ebca6bu);
    LIN_ASSERT(!live || (pix >= 0 && pix + NPX <= HW));
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      // the counter only grows (no reset to race with): visit v = i / STAGES of a stage is complete at (v + 1) * QWARPS
      const unsigned last = (static_cast<unsigned>(i / BULK_STAGES) + 1u) * BULK_QWARPS - 1u;
      const unsigned old = atomicAdd(&done[s * BULK_NQ + q], 1u);
      LIN_ASSERT(old + BULK_QWARPS > last && old <= last);        // visits of a stage's counter never mix
      if (old == last) issue(i + BULK_STAGES);
    }
    if (live) {
      u64 probe = zero2;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
#pragma unroll
        for (int h = 0; h < NP; ++h) {
          probe = fma2_rn(a[c][h], zero2, probe);
          probe = fma2_rn(b[c][h], zero2, probe);
        }
      }
      float pr0, pr1;
      unpack2(probe, pr0, pr1);
      if ((pr0 == pr0) && (pr1 == pr1))
        linear_frames<CT, NP, COUNTS, false, LOGITS>(a, b, HW, pix, n, labels, logits, have_tc, tc_word, ignore_index, wts, one2, cnt);
      else
        linear_frames<CT, NP, COUNTS, true, LOGITS>(a, b, HW, pix, n, labels, logits, have_tc, tc_word, ignore_index, wts, one2, cnt);
    }
    tc_word = tc_next;
  }
  if (COUNTS) cnt.finish(sh, counts, CT);
}

template <int CT, bool COUNTS, bool LOGITS, int NQ, int NP>
static int launch_bulk_nq(const float* prev, const float* next, long long HW, int n, uint8_t* labels, float* logits,
                          const uint8_t* tc_prev, long long* counts, int ignore_index, const BlendWeights& w,
                          cudaStream_t st) {
  const size_t smem = static_cast<size_t>(BULK_STAGES) * 2 * CT * BULK_TILE * 4 + BULK_STAGES * BULK_NQ_MAX * 12 + 32;
  auto kern = linear_blend_argmax_bulk_kernel<CT, COUNTS, LOGITS, NQ, NP>;
  static SmemOptIn optin;
  if (!optin.ensure(kern, static_cast<int>(smem))) return 1;   // caller uses the register-load kernel
  constexpr int pdl = 1;                       // programmatic dependent launch (r01: 23.5 -> 22.5 us per interval)
  const long long ntiles = (HW + BULK_TILE - 1) / BULK_TILE;
  const long long cap = sm_count();
  const int grid = static_cast<int>(ntiles < cap ? ntiles : cap);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(BULK_TILE / (2 * NP));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, prev, next, HW, n, labels, logits, tc_prev,
                                           reinterpret_cast<unsigned long long*>(counts), ignore_index, w, 1.0f, pdl);
  if (e != cudaSuccess) return set_error(FUVS_ECUDA, "fuvs_linear_blend_argmax(bulk): %s", cudaGetErrorString(e));
  count_launch();
  return FUVS_OK;
}

template <int CT, bool COUNTS, bool LOGITS>
static int launch_bulk(const float* prev, const float* next, long long HW, int n, uint8_t* labels, float* logits,
                       const uint8_t* tc_prev, long long* counts, int ignore_index, const BlendWeights& w,
                       cudaStream_t st) {
  // one pipeline per CTA, 4 pixels x 512 threads: the optimum of the r01 sweep (2 px x 1024 threads 25.9 us, 8 px x 256
  // threads 21.7 us, four 512-pixel slice pipelines 24.0 us, against 20.8 us)
  return launch_bulk_nq<CT, COUNTS, LOGITS, 1, 2>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
}

// ---------------------------------------------------------------------------
// Key frames given at DECODER resolution.  The reference up-samples the decoder output to
// the frame size (F.interpolate(bilinear, align_corners=True), flow/model.py:191-193,205-206) before anything else;
// that writes 2*S to HBM per interval and the blend reads it back.  Here the up-sample is evaluated on the fly with
// the op order kept (up-sample, THEN blend): a CTA owns one source-row interval i0 (the output rows whose floor source
// row is i0 — 8 rows at the usual stride-8 decoders), computes the horizontal interpolations of source rows i0 and
// i0+1 once into shared memory ([2 key frames][2 rows][C][W] floats), and every output row is then one vertical
// two-term per value followed by linear_frames<> — the same per-pixel code as the full-resolution kernels.
// ---------------------------------------------------------------------------
constexpr int LR_THREADS = 256;      // two CTAs per SM: one CTA's phase 1 overlaps the other's phase 2

template <int CT, bool COUNTS, bool LOGITS>
__global__ void __launch_bounds__(LR_THREADS, 2)
linear_lowres_kernel(const float* __restrict__ prev_lr, const float* __restrict__ next_lr, int hl, int wl, int H, int W,
                     int n, float sh, float sw, int XW, int nchunks,
                     uint8_t* __restrict__ labels, float* __restrict__ logits, const uint8_t* __restrict__ tc_prev,
                     unsigned long long* __restrict__ counts, int ignore_index, const BlendWeights wts, float one) {
  extern __shared__ __align__(16) float lr_hs[];            // [kf][row][c][XW]
  __shared__ unsigned sh24[24];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x / nchunks, chunk = blockIdx.x - i0 * nchunks;
  const int x0 = chunk * XW;
  const int xw = min(XW, W - x0);
  const long long HW = static_cast<long long>(H) * W;
  const int lplane = hl * wl;
  const u64 one2 = pack2(one, one);
  const float zero = __fsub_rn(one, one);
  const u64 zero2 = pack2(zero, zero);
  FieldCounts<CT> cnt;
  cnt.init();

  // output rows of this interval: floor(sh * y) == i0, with the float arithmetic of up_coord()
  auto src_row = [&](int y) { return static_cast<int>(__fmul_rn(sh, static_cast<float>(y))); };
  int y_lo = (sh > 0.f) ? static_cast<int>(static_cast<float>(i0) / sh) : 0;
  y_lo = max(0, min(y_lo, H - 1));
  while (y_lo > 0 && src_row(y_lo - 1) >= i0) --y_lo;
  while (y_lo < H && src_row(y_lo) < i0) ++y_lo;
  int y_hi = y_lo;
  while (y_hi < H && src_row(y_hi) == i0) ++y_hi;

  if (y_hi > y_lo) {
    // ---- phase 1: horizontal two-terms of source rows i0 and i0 + ip (UpSample.cuh: w0*a + w1*b)
    const int ip_h = (i0 < hl - 1) ? 1 : 0;
    const int nkf = (n > 1) ? 2 : 1;
    for (int xx = tid; xx < xw; xx += LR_THREADS) {
      const UpCoord wc = up_coord<Nm>(sw, x0 + xx, wl);
      const int o0 = i0 * wl + wc.i0, o1 = (i0 + ip_h) * wl + wc.i0;
#pragma unroll
      for (int kf = 0; kf < 2; ++kf) {
        if (kf < nkf) {
          const float* src = kf ? next_lr : prev_lr;
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            const float* pl = src + c * lplane;
            lr_hs[((kf * 2 + 0) * CT + c) * XW + xx] = two_term<Nm::kUpInner>(wc.l0, __ldg(pl + o0), wc.l1, __ldg(pl + o0 + wc.ip));
            lr_hs[((kf * 2 + 1) * CT + c) * XW + xx] = two_term<Nm::kUpInner>(wc.l0, __ldg(pl + o1), wc.l1, __ldg(pl + o1 + wc.ip));
          }
        }
      }
    }
    __syncthreads();
    // ---- phase 2: vertical two-term, blends, arg-max, counts
    for (int y = y_lo; y < y_hi; ++y) {
      const UpCoord hc = up_coord<Nm>(sh, y, hl);
      const u64 hl0 = pack2(hc.l0, hc.l0), hl1 = pack2(hc.l1, hc.l1);
      for (int xx = tid * 4; xx < xw; xx += LR_THREADS * 4) {
        u64 a[CT][2], b[CT][2];
        const long long pix = static_cast<long long>(y) * W + x0 + xx;
        const bool have_tc = COUNTS && tc_prev != nullptr;
        const unsigned tc_word = have_tc ? __ldg(reinterpret_cast<const unsigned*>(tc_prev + pix)) : 0u;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const ulonglong2 r0 = *reinterpret_cast<const ulonglong2*>(lr_hs + ((0 * 2 + 0) * CT + c) * XW + xx);
          const ulonglong2 r1 = *reinterpret_cast<const ulonglong2*>(lr_hs + ((0 * 2 + 1) * CT + c) * XW + xx);
          a[c][0] = two_term2<Nm::kUpOuter>(hl0, r0.x, hl1, r1.x, one2);
          a[c][1] = two_term2<Nm::kUpOuter>(hl0, r0.y, hl1, r1.y, one2);
          if (n > 1) {
            const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(lr_hs + ((1 * 2 + 0) * CT + c) * XW + xx);
            const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(lr_hs + ((1 * 2 + 1) * CT + c) * XW + xx);
            b[c][0] = two_term2<Nm::kUpOuter>(hl0, s0.x, hl1, s1.x, one2);
            b[c][1] = two_term2<Nm::kUpOuter>(hl0, s0.y, hl1, s1.y, one2);
          } else {
            b[c][0] = zero2;
            b[c][1] = zero2;
          }
        }
        u64 probe = zero2;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          probe = fma2_rn(a[c][0], zero2, probe);
          probe = fma2_rn(a[c][1], zero2, probe);
          probe = fma2_rn(b[c][0], zero2, probe);
          probe = fma2_rn(b[c][1], zero2, probe);
        }
        float pr0, pr1;
        unpack2(probe, pr0, pr1);
        if ((pr0 == pr0) && (pr1 == pr1))
          linear_frames<CT, 2, COUNTS, false, LOGITS>(a, b, HW, pix, n, labels, logits, have_tc, tc_word, ignore_index, wts, one2, cnt);
        else
          linear_frames<CT, 2, COUNTS, true, LOGITS>(a, b, HW, pix, n, labels, logits, have_tc, tc_word, ignore_index, wts, one2, cnt);
      }
    }
  }
  if (COUNTS) cnt.finish(sh24, counts, CT);
}

static inline float lr_scale(int in_size, int out_size) {
  return out_size > 1 ? static_cast<float>(in_size - 1) / (out_size - 1) : 0.f;   // area_pixel_compute_scale, align_corners
}

template <int CT>
static int launch_lowres(const float* prev_lr, const float* next_lr, int hl, int wl, int H, int W, int n, uint8_t* labels,
                         float* logits, const uint8_t* tc_prev, long long* counts, int ignore_index,
                         const BlendWeights& w, cudaStream_t st) {
  const int budget_floats = (108 * 1024) / 4;      // two CTAs per SM
  int nchunks = 1;
  while (4ll * CT * (((W + nchunks - 1) / nchunks + 3) & ~3) > budget_floats) ++nchunks;
  const int XW = ((W + nchunks - 1) / nchunks + 3) & ~3;
  const size_t smem = static_cast<size_t>(4) * CT * XW * sizeof(float);
  const int grid = hl * nchunks;
  auto cu = reinterpret_cast<unsigned long long*>(counts);
  const float sh = lr_scale(hl, H), sw = lr_scale(wl, W);
#define FUVS_LR(CNT_, LG_)                                                                                              \
  do {                                                                                                                  \
    auto kern = linear_lowres_kernel<CT, CNT_, LG_>;                                                                    \
    static SmemOptIn optin;                 /* once per device, for the whole two-CTAs-per-SM budget */                 \
    if (!optin.ensure(kern, 108 * 1024))                                                                                \
      return set_error(FUVS_ECUDA, "linear_lowres: cannot reserve %zu bytes of shared memory", smem);                   \
    kern<<<grid, LR_THREADS, smem, st>>>(prev_lr, next_lr, hl, wl, H, W, n, sh, sw, XW, nchunks, labels, logits,        \
                                         tc_prev, cu, ignore_index, w, 1.0f);                                           \
  } while (0)
  if (counts) { if (logits) FUVS_LR(true, true); else FUVS_LR(true, false); }
  else        { if (logits) FUVS_LR(false, true); else FUVS_LR(false, false); }
#undef FUVS_LR
  return check_launch("fuvs_linear_lowres_blend_argmax");
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

template <int CT, int NP, bool COUNTS>
static int launch_v4(const float* prev, const float* next, long long HW, int n, uint8_t* labels, float* logits,
                     const uint8_t* tc_prev, long long* counts, int ignore_index, const BlendWeights& w,
                     cudaStream_t st) {
  const int threads = 256;
  const long long need = ((HW / (2 * NP)) + threads - 1) / threads;
  static int bps = blocks_per_sm(linear_blend_argmax_v4_kernel<CT, NP, COUNTS>, threads);
  const long long cap = static_cast<long long>(sm_count()) * bps;
  const int grid = static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
  linear_blend_argmax_v4_kernel<CT, NP, COUNTS><<<grid, threads, 0, st>>>(
      prev, next, HW, n, labels, logits, tc_prev, reinterpret_cast<unsigned long long*>(counts), ignore_index, w, 1.0f);
  return check_launch("fuvs_linear_blend_argmax");
}

template <int CT, int VEC>
static int launch_fixed(const float* prev, const float* next, long long HW, int n, uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts, int ignore_index, const BlendWeights& w,
                        cudaStream_t st) {
  if (VEC == 4 && (ignore_index < 0 || ignore_index >= CT)) {
    if (CT <= 5 && HW >= 4 * BULK_TILE) {
      constexpr int CB = CT <= 5 ? CT : 2;
      int r;
      if (logits)
        r = counts ? launch_bulk<CB, true, true>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st)
                   : launch_bulk<CB, false, true>(prev, next, HW, n, labels, logits, nullptr, nullptr, ignore_index, w, st);
      else
        r = counts ? launch_bulk<CB, true, false>(prev, next, HW, n, labels, nullptr, tc_prev, counts, ignore_index, w, st)
                   : launch_bulk<CB, false, false>(prev, next, HW, n, labels, nullptr, nullptr, nullptr, ignore_index, w, st);
      if (r <= 0) return r;
    }
    // small frames, C > 5, or no shared-memory opt-in: register-load kernel, 4 pixels per thread
    if (counts) return launch_v4<CT, 2, true>(prev, next, HW, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    return launch_v4<CT, 2, false>(prev, next, HW, n, labels, logits, nullptr, nullptr, ignore_index, w, st);
  }
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

extern "C" int fuvs_linear_lowres_supported(int C, int H, int W) {
  return (C >= 2 && C <= 5 && (W % 4) == 0) ? 1 : 0;
}

extern "C" int fuvs_linear_lowres_blend_argmax(const float* prev_lr, const float* next_lr, int C, int hl, int wl, int H,
                                               int W, int n, uint8_t* labels, float* logits, const uint8_t* tc_prev,
                                               long long* counts, int ignore_index, fuvs_stream_t stream) {
  using namespace fuvs;
  if (int e = device_ok()) return e;
  if (!prev_lr || C < 1 || hl < 1 || wl < 1 || H < 1 || W < 1 || n < 1)
    return set_error(FUVS_EINVAL, "linear_lowres: bad shape C=%d %dx%d -> %dx%d n=%d", C, hl, wl, H, W, n);
  if (hl == H && wl == W)     // the reference skips the interpolate when the sizes match (flow/model.py:191)
    return fuvs_linear_blend_argmax(prev_lr, next_lr, C, H, W, n, labels, logits, tc_prev, counts, ignore_index, stream);
  if (n > 1 && !next_lr) return set_error(FUVS_EINVAL, "linear_lowres: next key frame is NULL but n=%d", n);
  if (n > FUVS_MAX_FRAMES) return set_error(FUVS_EINVAL, "linear_lowres: n=%d exceeds %d frames per interval", n, FUVS_MAX_FRAMES);
  if (!fuvs_linear_lowres_supported(C, H, W))
    return set_error(FUVS_EINVAL, "linear_lowres: needs 2 <= C <= 5 and W %% 4 == 0 (C=%d W=%d); up-sample with "
                     "fuvs_upsample_bilinear_ac and call fuvs_linear_blend_argmax instead", C, W);
  if (counts && ignore_index >= 0 && ignore_index < C)
    return set_error(FUVS_EINVAL, "linear_lowres: ignore_index=%d collides with a class", ignore_index);
  if (static_cast<long long>(hl) * wl * C >= (1ll << 31)) return set_error(FUVS_EINVAL, "linear_lowres: source too large");
  if ((logits && !aligned16(logits)) || (labels && !aligned4(labels)) || (tc_prev && !aligned4(tc_prev)))
    return set_error(FUVS_EALIGN, "linear_lowres: logits must be 16-byte, label maps 4-byte aligned");
  if (!labels && !logits && !counts) return FUVS_OK;
  BlendWeights w;
  make_blend_weights(n, &w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 2: return launch_lowres<2>(prev_lr, next_lr, hl, wl, H, W, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 3: return launch_lowres<3>(prev_lr, next_lr, hl, wl, H, W, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    case 4: return launch_lowres<4>(prev_lr, next_lr, hl, wl, H, W, n, labels, logits, tc_prev, counts, ignore_index, w, st);
    default: return launch_lowres<5>(prev_lr, next_lr, hl, wl, H, W, n, labels, logits, tc_prev, counts, ignore_index, w, st);
  }
}
