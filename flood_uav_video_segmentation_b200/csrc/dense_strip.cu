// dense_strip.cu — dense-flow warp step as a column-strip sliding window (sm_100a).
//
// Why: the per-plane TMA kernel (dense_tma.cu) stages, for every 128x16 output tile, a 192x48 source box per channel:
// every source byte crosses the L2->SM fabric 4.5 times (409 MB per step at 1080p, profiles/r01_ncu_dense_tma_v5.txt),
// which is what bounds it.  Here a CTA owns a run of consecutive 8-row blocks of one 128-pixel column strip and keeps
// the source window of ALL channels resident in shared memory as a ring of rows.  Moving one block down the strip
// costs 8 new rows per channel, so a source byte crosses the fabric ~1.5 times (x-halo only) plus a 32-row warm-up per
// run.
//
//   work     : 2 sides x ceil(W/128) strips x ceil(H/8) blocks, numbered side-major/strip/row; CTA b of a persistent
//              grid (one CTA per SM) takes a contiguous cost-weighted range.  A range may cross a strip boundary: it
//              is then processed as several "segments", each with its own window warm-up.
//   ring     : per channel NS*8 rows of 192 floats (NS = 7 slots of 8 rows for C=5: 215 KB), channel-major
//              [c][ring row][x], so that a tap address is ring_row*192 + x with the channel as an immediate offset
//              and the row below is +192 except at the wrap.  A block needs the 5 slots covering rows
//              [y0-16, y0+24); the others are prefetched slots of the following blocks.  Slot loads are numbered
//              s = 0,1,2.. in the order of use; load s lives in slot s % NS and completes phase (s / NS) & 1 of
//              mbarrier full[s % NS] (C cp.async.bulk.tensor copies of 192 x 8 floats, one per channel).
//   protocol : no block-wide barrier.  A warp that has finished block k bumps done[k & 3] (counters only grow, so there
//              is no reset to race with); the LAST warp to do so knows which slot loads are now dead and issues the
//              loads that reuse their rows.  The counter's old value returns through the MIO queue behind the gathers
//              of the other warps, so it is looked at only after the coordinates of the next block are computed
//              (measured r02: the warp used to sit 8 % of its time on that round trip and the fence in front of it).
//   registers: flow vectors are prefetched two blocks ahead and the pointwise operand of emitting steps one block
//              ahead into two alternating register sets (the block loop is unrolled by two), so no register move
//              waits for a load in the iteration that issued it.
//   body     : (r02) per block and thread: coordinates and weights of its two pixels from registers, then ALL tap
//              loads of both pixels in one burst of volatile ld.shared (one round trip through the shared-memory
//              pipe per block instead of the four dependent ones ptxas produced when it sank loads under the store
//              predicates), then the FMA chains, blends and stores.  Full-size shapes (W % 128 == 0, H % 8 == 0)
//              compile without per-pixel validity predicates.
//   layout   : (r02) C = 5, odd n: the chain states in the scratch (never seen by the caller) are stored "4+1":
//              channels 0-3 interleaved per pixel ([H][W][4]) followed by channel 4 as a plane.  A tap is then one
//              128-bit + one 32-bit shared-memory load instead of five 32-bit ones (16 instead of 40 loads per block
//              and thread, ~13 % fewer shared-memory cycles with per-pixel random flow: tools/probes/lds_gather_probe.cu),
//              a state write one 128-bit + one 32-bit store, the pointwise operand likewise.  Step 1 reads the caller's
//              planar key frames and writes 4+1; measured 239 -> 22x us per interval.
//   taps     : a warp whose taps leave the window (large motion) gathers its pixels of that block from global memory
//              instead, so any flow field stays correct.
//   counts   : stay in fuvs_temporal_counts (metric.cu), 8 us of stream time behind the last step (nothing can share an
//              SM with a strip CTA, which holds all registers).  Measured in r02 and rejected: counting inside the block
//              loop (two more loads per pixel on the gather's path, 242 vs 230 us); the last step's CTAs counting
//              their own pixels after a block-wide barrier, labels staged through the dead ring with cp.async
//              (227.2 / 222.6 us on 1 / 2 streams against 228.0 / 221.0 for the separate launch, both with the
//              bit-plane counters of temporal_fields.cuh); every warp counting its own pixels as soon as it is done,
//              without a barrier (236.7 / 230.1: the unrolled counting code is fetched cold by one warp after the other).
// Arithmetic is gs_setup/tap_acc/blend2 from fuvs_common.cuh: bit-identical to the direct kernel (warp.cu) and to
// ATen's grid_sampler_2d (flow/model.py:244-249).
#include <type_traits>
#include <utility>

#include "dense_common.cuh"
#include "pix4.cuh"
#include "tma_ptx.cuh"

// Developer build -DFUVS_STRIP_ASSERT (tools/strip_asserts.sh): bounds checks of every ring address, global index,
// slot number and completion-counter value of the barrier-free protocol — compute-sanitizer is closed on the GPU pool
// this was developed on (profiles/r02_sanitizer_unavailable.txt), so the kernel checks itself under the parity tests.
#ifdef FUVS_STRIP_ASSERT
#include <cassert>
#define STRIP_ASSERT(x) assert(x)
#else
#define STRIP_ASSERT(x) ((void)0)
#endif
// -DFUVS_STRIP_JITTER: every warp stalls for a pseudo-random 0 .. 16 k cycles at a pseudo-random third of its blocks
// (before it reads the ring and before it hands a block back), so that the warps of a CTA drift apart as far as the
// protocol lets them and the hand-over runs under every interleaving the parity tests can provoke.
#ifdef FUVS_STRIP_JITTER
#define STRIP_JITTER(salt)                                                                              \
  do {                                                                                                  \
    unsigned h_ = (static_cast<unsigned>(kglob) * 2654435761u) ^ ((threadIdx.x >> 5) * 40503u) ^        \
                  (blockIdx.x * 2246822519u) ^ (salt);                                                  \
    h_ ^= h_ >> 15; h_ *= 2246822519u; h_ ^= h_ >> 13;                                                  \
    if ((h_ % 3u) == 0u) {                                                                              \
      const long long t_ = clock64() + ((h_ >> 8) & 16383u);                                            \
      while (clock64() < t_) {}                                                                         \
    }                                                                                                   \
  } while (0)
#else
#define STRIP_JITTER(salt) ((void)0)
#endif

namespace fuvs {

namespace {

using namespace tma;

constexpr int TW = 128;                      // strip width = threads per thread-row
constexpr int TROWS = 4;                     // thread rows; a thread owns rows ty and ty + 4 of a block
constexpr int PX = 2;                        // (r02: 8 thread rows x 1 pixel, 1024 threads at 64 registers: 324 vs 230 us)
constexpr int THREADS = TW * TROWS;
constexpr int NWARPS = THREADS / 32;
constexpr int HALO_X = 32, HALO_Y = 16;
// Row stride 192 = 0 mod 32 banks on purpose: with coherent flow a warp's taps straddle at most two rows and stay
// conflict-free whatever the rows are (a skewed stride was measured in r01: iid flow -1 %, coherent flow +5 %).
constexpr int BOXW = TW + 2 * HALO_X;        // 192
constexpr int RB = 8;                        // rows per block = rows per ring slot
constexpr int WIN = (RB + 2 * HALO_Y) / RB;  // 5 slots cover the window of one block
constexpr int WIN_ROWS = WIN * RB;           // 40
constexpr int MAX_SLOTS = 8;
constexpr int SMEM_LIMIT = 232448;           // 227 KB opt-in maximum per CTA
constexpr int PLANE = RB * BOXW;             // floats of one channel inside a slot
constexpr int nslot_for(int C) {
  return (SMEM_LIMIT - 256) / (C * PLANE * 4) > MAX_SLOTS ? MAX_SLOTS : (SMEM_LIMIT - 256) / (C * PLANE * 4);
}
constexpr int NSLOT_C2 = nslot_for(2), NSLOT_C5 = nslot_for(5);   // 8, 7

struct StripGeom {
  int nsx, nby, total, nslot;
#ifdef FUVS_STRIP_TIMING
  int seq;                                   // launch number (developer build)
#endif
};
constexpr int MAX_GRID = 255;
struct StripPartition {
  int b[MAX_GRID + 1];                       // CTA i owns blocks [b[i], b[i+1])
};

// CTA i of `grid` takes the blocks whose cumulative cost lies in [i, i+1) * total_cost / grid.  A block's cost is
// side weight (step 1: the forward side also emits frame 0) x strip weight (the two strips at the image edge cost
// ~12 % more: border clipping sends many lanes to the same column, i.e. the same shared-memory bank).  A CTA whose
// range crosses into a new strip has to refill its whole window there: the first block of every strip carries an
// extra cost of wstart / 8 blocks, which hands the CTA that owns it correspondingly fewer blocks.
//   wl, wr : relative cost of a forward-side / backward-side block;  wedge : cost of an edge-strip block in sixteenths
void strip_partition(const StripGeom& G, int grid, int wl, int wr, int wedge, int wstart, StripPartition* P) {
  auto unit_w = [&](int u) {
    const bool side = u >= G.nsx;
    const int strip = side ? u - G.nsx : u;
    const bool edge = (G.nsx > 2) && (strip == 0 || strip == G.nsx - 1);
    return static_cast<long long>(side ? wr : wl) * (edge ? wedge : 16);
  };
  auto unit_start = [&](int u) { return (static_cast<long long>(wstart) * unit_w(u)) >> 3; };
  long long cost_all = 0;
  for (int u = 0; u < 2 * G.nsx; ++u) cost_all += static_cast<long long>(G.nby) * unit_w(u) + unit_start(u);
  auto block_at = [&](long long num) {        // first block whose start cost is >= num * cost_all / grid
    long long pos = (num * cost_all + grid - 1) / grid;
    for (int u = 0; u < 2 * G.nsx; ++u) {
      const long long w = unit_w(u), st = unit_start(u), cu = static_cast<long long>(G.nby) * w + st;
      if (pos <= cu) {
        if (pos <= 0) return u * G.nby;
        const long long k = (pos - st + w - 1) / w;          // block k of the strip starts at cost st + k w (k >= 1)
        return u * G.nby + static_cast<int>(k < 1 ? 1 : k);
      }
      pos -= cu;
    }
    return G.total;
  };
  for (int i = 0; i < grid; ++i) P->b[i] = block_at(i);
  P->b[grid] = G.total;
}
struct StripMaps {
  CUtensorMap srcL, srcR;                    // planar [C][H][W] fp32, box BOXW x RB x 1 (4+1 sources: the plane of channel 4;
};                                           // their interleaved part comes in row-wise bulk copies, see issue())

#ifdef FUVS_STRIP_TIMING
__device__ long long* g_strip_timing = nullptr;      // developer build only (tools/strip_timing.py)
int g_strip_seq = 0;
#endif

template <int... I, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F&& f) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, f);
}

// volatile: ptxas keeps these in program order and cannot sink them under a later predicate
template <int OFF>
__device__ __forceinline__ float lds_imm(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ float4 lds4_imm(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ float lds_rt(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// Relaxed on purpose: every ring read of the calling warp has been consumed by an FMA before this instruction issues
// (an SM issues a warp's instructions in order), so there is nothing a release fence would still have to wait for — a
// fence here would wait for the warp's global stores and prefetches instead.  The winner fences before it refills.
__device__ __forceinline__ unsigned atom_add_relaxed(uint32_t addr, unsigned v) {
  unsigned old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}

// First maximum over CT class values with torch.max's rules (lowest index wins ties, the first NaN wins outright): the
// class maxima propagate NaN (max.NaN), the index is the number of leading classes that differ from the maximum (pix4.cuh);
// a pixel with a NaN class value takes the compare-select scan.  10 instead of 20 instructions per pixel.
template <int CT>
__device__ __forceinline__ int argmax_classes(const float (&v)[CT]) {
  if constexpr (CT < 2) {
    return 0;
  } else {
    const float m = max_classes_nan<CT>(v);
    if (m != m) {
      ArgMax am;
      am.init(-INFINITY);
#pragma unroll
      for (int c = 0; c < CT; ++c) am.push(v[c], c);
      return am.idx;
    }
    float t = differs(v[CT - 2], m);
#pragma unroll
    for (int c = CT - 3; c >= 0; --c) {
      const float nc = differs(v[c], m);
      t = __fmaf_rn(nc, t, nc);
    }
    return __float2int_rz(t);
  }
}

// Element (channel c, pixel off) of a state: planar [C][HW], or 4+1 (channels 0-3 interleaved, channel 4 a plane).
__device__ __forceinline__ long long state_index(bool il, int c, long long off, long long HW) {
  return il ? (c < 4 ? off * 4 + c : 4 * HW + off) : c * HW + off;
}

// Rare path: the warp computes its pixels of this block straight from global memory (same math as warp.cu).
__device__ __noinline__ void block_from_global(const DenseStep* __restrict__ Ap, bool side, bool emit, bool key0, int C,
                                               int H, int W, int x, int ytop, unsigned valid, bool il, bool il_src) {
  const DenseStep& A = *Ap;
  const long long HW = static_cast<long long>(H) * W;
  const float* grid = side ? A.gridR : A.gridL;
  const float* src = side ? A.srcR : A.srcL;
  float* dst = side ? A.dstR : A.dstL;
  const float* point = side ? A.pointL : A.pointR;
  const float w_this = side ? A.wB1 : A.wA0, w_point = side ? A.wB0 : A.wA1;
  uint8_t* lab_out = side ? A.labelB : A.labelA;
  float* logit_out = side ? A.logitB : A.logitA;
  for (int r = 0; r < PX; ++r) {
    if (!((valid >> r) & 1u)) continue;
    const long long pix = static_cast<long long>(ytop + r * TROWS) * W + x;
    const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + pix);
    const GsTap t = gs_setup<Nm>(g.x, g.y, H, W, false);
    ArgMax am, am0;
    am.init(-INFINITY);
    am0.init(-INFINITY);
    for (int c = 0; c < C; ++c) {
      const long long o = c * HW + pix;
      const long long n0 = t.off00, n1 = n0 + t.dx, s0 = n0 + static_cast<long long>(t.dy) * W, s1 = s0 + t.dx;
      float acc = tap_acc<Nm>(0.f, __ldg(src + state_index(il_src, c, n0, HW)), t.nw);
      if (t.dx) acc = tap_acc<Nm>(acc, __ldg(src + state_index(il_src, c, n1, HW)), t.ne);
      if (t.dy) acc = tap_acc<Nm>(acc, __ldg(src + state_index(il_src, c, s0, HW)), t.sw);
      if (t.dx & t.dy) acc = tap_acc<Nm>(acc, __ldg(src + state_index(il_src, c, s1, HW)), t.se);
      if (dst) dst[state_index(il, c, pix, HW)] = acc;
      if (emit) {
        const float other = __ldg(point + state_index(il, c, pix, HW));
        const float v = blend2(w_this, acc, w_point, other);
        am.push(v, c);
        if (logit_out) __stcs(logit_out + o, v);
      }
      if (key0) {
        const float v = __ldg(A.key0 + state_index(il_src, c, pix, HW));
        am0.push(v, c);
        if (A.logit0) __stcs(A.logit0 + o, v);
      }
    }
    if (emit && lab_out) lab_out[pix] = static_cast<uint8_t>(am.idx);
    if (key0 && A.label0) A.label0[pix] = static_cast<uint8_t>(am0.idx);
  }
}

//   EMIT   : both sides complete a frame this step (blend with the other chain's state read pointwise, arg-max)
//   KEY0   : step 1: the forward side also emits frame 0 = arg-max of the key frame (read from the ring itself)
//   WDST   : the step writes its states (every step but the last)
//   FULLV  : W % 128 == 0 and H % 8 == 0: every pixel of every block exists
//   NSC    : ring slots when known at compile time (0: G.nslot)
//   IL     : C = 5 only: the states of this interval (dst, pointwise operand, and the source unless KEY0) are 4+1
//   KIL    : KEY0 and IL: the key frames themselves are 4+1 (written so by fuvs_dense_lowres_interval's up-sample)
template <int CT, int NSC, bool EMIT, bool KEY0, bool WDST, bool FULLV, bool IL, bool KIL>
__global__ void __launch_bounds__(THREADS, 1)
dense_strip_kernel(const __grid_constant__ StripMaps M, const __grid_constant__ DenseStep A, int Crt, int H, int W,
                   StripGeom G, const __grid_constant__ StripPartition P) {
  static_assert(!(EMIT && KEY0), "frame 0 and a completed frame never share a step here (the host routes n<=2 elsewhere)");
  static_assert(!IL || CT == 5, "the 4+1 state layout exists for C = 5");
  static_assert(!KIL || (KEY0 && IL), "4+1 key frames belong to step 1 of a 4+1 interval");
  constexpr bool SIL = IL && (!KEY0 || KIL);     // the source window is 4+1: [ring row][x][4] then the plane of channel 4
  constexpr int CR = CT > 0 ? CT : 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int C = CT > 0 ? CT : Crt;
  const int NS = NSC > 0 ? NSC : G.nslot;
  const int NSR = NS * RB;                                             // ring rows per channel
  const int chan_floats = NSR * BOXW;
  const uint32_t slot_bytes = static_cast<uint32_t>(C) * PLANE * 4u;
  float* bufs = reinterpret_cast<float*>(smem_raw);
  // 64 bytes of pad behind the ring: the (never accumulated) east tap of a pixel clipped to the last window column of
  // the last ring row must not alias the barrier words
  unsigned long long* bars =
      reinterpret_cast<unsigned long long*>(smem_raw + static_cast<size_t>(C) * chan_floats * sizeof(float) + 64);
  const uint32_t bar0 = smem_u32(bars);                               // full[MAX_SLOTS]
  const uint32_t done0 = bar0 + 8u * MAX_SLOTS;                       // done[4] block-completion counters
  const uint32_t ring0 = smem_u32(bufs);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int tx = tid & (TW - 1), ty = tid >> 7;                        // TW == 128
  const int HWi = H * W;                                               // the host guarantees C*H*W < 2^31
  // This CTA's contiguous, cost-weighted range of blocks: computed on the host (strip_partition) — on the device the
  // two searches were ~1 500 instructions per thread in front of every step's first load.
  const int B0 = P.b[blockIdx.x], B1 = P.b[blockIdx.x + 1];

  // Programmatic dependent launch: this CTA may have been scheduled while the previous kernel of the stream (the
  // step that produced our source states) is still draining.  Let our own successor do the same, set up the
  // barriers, then wait for the predecessor's memory to be complete before the first global access.
  // (r02: triggering only 2 / 4 / 8 blocks before the CTA's end, or never, changes nothing measurable on one or two streams)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    for (int b = 0; b < MAX_SLOTS; ++b) mbar_init(bar0 + 8u * b, 1);
    for (int b = 0; b < 4; ++b) reinterpret_cast<unsigned*>(bars + MAX_SLOTS)[b] = 0u;
    fence_barrier_init();
  }
  __syncthreads();
#ifdef FUVS_STRIP_TIMING
  const long long tm_start = clock64();
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef FUVS_STRIP_TIMING
  const long long tm_wait = clock64();
  long long tm_first = 0;
#endif

  // Left edge of a strip's window.  Kept inside the image where the image is wide enough: a box that hangs over
  // the image edge is zero-filled by TMA but loads far slower, and border clipping never reads beyond the edge.
  auto box_x = [&](int strip) { return max(0, min(strip * TW - HALO_X, W - BOXW)); };
  // Segments of this CTA's range [B0, B1): the first starts at block J00 of unit U0, the following ones at block 0 of
  // the next units.  visit(u, j0, nb, sb, k0) gets the unit, its first block, the number of blocks, the index of the
  // segment's first slot load and the CTA-local index of its first block; returning true ends the walk.
  const int U0 = B0 / G.nby, J00 = B0 - U0 * G.nby;
  auto walk = [&](auto&& visit) {
    int u = U0, j0 = J00, left = B1 - B0, sb = 0, k0 = 0;
    while (left > 0) {
      const int nb = min(G.nby - j0, left);
      if (visit(u, j0, nb, sb, k0)) return;
      sb += nb + WIN - 1;
      k0 += nb;
      left -= nb;
      ++u;
      j0 = 0;
    }
  };
  // one thread: start slot load s of this CTA's sequence (its rows are known to be dead)
  auto issue = [&](int s) {
    walk([&](int u, int j0, int nb, int sb, int) {
      if (s >= sb + nb + WIN - 1) return false;
      const int ytop = (j0 + (s - sb)) * RB - HALO_Y;
      const bool side = u >= G.nsx;
      const int strip = side ? u - G.nsx : u;
      const int b = s % NS;
      STRIP_ASSERT(s >= 0 && b >= 0 && b < NS && NS <= MAX_SLOTS && u < 2 * G.nsx && strip < G.nsx);
      const uint32_t bar = bar0 + 8u * b;
      if (ytop + RB <= 0 || ytop >= H) {
        mbar_arrive(bar);                    // slot entirely outside the image: never read (border clipping)
      } else {
        const int bx = box_x(strip);
        if (SIL) {
          // channels 0-3: one window row is 192 x 16 contiguous bytes of the [H][W][4] part -> one bulk copy per row
          // that exists (a tensor box with a 16-byte inner extent would be split into 16-byte requests); channel 4:
          // one box of its plane (rows outside the image are zero-filled and count as transferred)
          const int r0 = max(0, -ytop), r1 = min(RB, H - ytop);
          const uint32_t rowb = static_cast<uint32_t>(min(BOXW, W - bx)) * 16u;
          mbar_expect_tx(bar, static_cast<uint32_t>(r1 - r0) * rowb + PLANE * 4u);
          const float* q = side ? A.srcR : A.srcL;
          for (int r = r0; r < r1; ++r)
            load_bulk(ring0 + static_cast<uint32_t>(b * PLANE + r * BOXW) * 16u,
                      q + (static_cast<long long>(ytop + r) * W + bx) * 4, rowb, bar);
          load_3d(ring0 + static_cast<uint32_t>(4 * chan_floats + b * PLANE) * 4u, side ? &M.srcR : &M.srcL, bx, ytop, 0, bar);
        } else {
          mbar_expect_tx(bar, slot_bytes);
          const uint32_t dst = ring0 + static_cast<uint32_t>(b * PLANE) * 4u;
          for (int c = 0; c < C; ++c)
            load_3d(dst + static_cast<uint32_t>(c * chan_floats) * 4u, side ? &M.srcR : &M.srcL, bx, ytop, c, bar);
        }
      }
      return true;
    });
  };
  // initial fill: slot s by lane 0 of warp s (one thread issuing all 7 x 9 copies back to back took ~1 us, which every
  // CTA of every step spent before its first block)
  // (Issuing only the WIN slots of the first window here and the prefetch slots once it is complete shortens the warm-up
  // of step 1 from 8.2 k to 7.0 k cycles and changes nothing per interval: tools/strip_timing.py.)
  if (lane == 0 && (tid >> 5) < NS) issue(tid >> 5);
  const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
  const float Wm1 = static_cast<float>(W - 1), Hm1 = static_cast<float>(H - 1);
  int u = U0, j0 = J00, left = B1 - B0, sb = 0, kglob = 0, released = 0;
  // lane 0 of every warp: old value of the completion counter of the block it finished last, and what to refill if
  // that value says it was the last warp (checked after the next block's coordinates, see `protocol` above)
  unsigned tok = 0u, tok_want = 1u;
  int tok_from = 0, tok_to = 0;
  auto refill_if_last = [&]() {
    if (lane == 0 && tok == tok_want) {
      __threadfence_block();                       // winner only: acquire side of the hand-over
      for (int s = tok_from + NS; s < tok_to + NS; ++s) issue(s);
    }
    tok_want = tok + 1u;                           // checked once
  };
  while (left > 0) {                             // ---- one segment: nb consecutive blocks of one (side, strip)
    const int nb = min(G.nby - j0, left);
    const bool side = u >= G.nsx;
    const int strip = side ? u - G.nsx : u;
    const int x0 = strip * TW, xbase = box_x(strip);
    const int x = x0 + tx;
    const float2* grid = reinterpret_cast<const float2*>(side ? A.gridR : A.gridL);
    float* dst = side ? A.dstR : A.dstL;
    const float* point = side ? A.pointL : A.pointR;
    const float w_this = side ? A.wB1 : A.wA0;     // weight of the state computed here
    const float w_point = side ? A.wB0 : A.wA1;    // weight of the state read pointwise
    uint8_t* lab_out = side ? A.labelB : A.labelA;
    float* logit_out = side ? A.logitB : A.logitA;
    const bool do_key0 = KEY0 && !side;
    const int rstride = TROWS * W;
    const bool vx = FULLV || x < W;
    const int kx = -0x4B000000 - xbase;            // window column of a tap: mantissa bits of (ix + 2^23) + kx
    const uint32_t colbase = ring0 + static_cast<uint32_t>(min(x, W - 1) - xbase) * 4u;   // this thread's own column

    auto load_grid = [&](int yb, bool on, float2 (&dstg)[PX]) {
#pragma unroll
      for (int r = 0; r < PX; ++r) {
        const int y = yb + ty + r * TROWS;
        dstg[r] = (on && vx && (FULLV || y < H)) ? __ldcs(grid + y * W + x) : make_float2(0.f, 0.f);   // read once: evict first
      }
    };
    auto load_other = [&](int yb, bool on, float (&dsto)[PX][CR]) {
      if (!(EMIT && CT > 0)) return;
#pragma unroll
      for (int r = 0; r < PX; ++r) {
        const int y = yb + ty + r * TROWS;
        const bool live = on && vx && (FULLV || y < H);
        if constexpr (IL) {
          const float4 q = live ? __ldcs(reinterpret_cast<const float4*>(point) + (y * W + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
          dsto[r][0] = q.x; dsto[r][1] = q.y; dsto[r][2] = q.z; dsto[r][3] = q.w;
          dsto[r][CR - 1] = live ? __ldcs(point + (4 * HWi + y * W + x)) : 0.f;
        } else {
#pragma unroll
          for (int c = 0; c < CR; ++c) dsto[r][c] = live ? __ldcs(point + (c * HWi + y * W + x)) : 0.f;   // last use of that state
        }
      }
    };
    // two alternating register sets: block jj uses set jj & 1.  Flow vectors of block jj+2 are loaded into the set of
    // block jj as soon as its coordinates are computed; the pointwise operand of block jj+1 goes into the other set at
    // the top of block jj (that set was consumed at the end of block jj-1).
    float2 gbuf[2][PX];
    float obuf[2][PX][CR];
    load_grid(j0 * RB, true, gbuf[0]);
    load_grid(j0 * RB + RB, nb > 1, gbuf[1]);
    load_other(j0 * RB, true, obuf[0]);
    // ring position of the window's top slot (slot load sb + jj): slot b0, phase parity q0
    int b0 = sb % NS;
    unsigned q0 = static_cast<unsigned>(sb / NS) & 1u;

    for (int jj0 = 0; jj0 < nb; jj0 += 2) {
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int jj = jj0 + hb;
        if (jj >= nb) break;
        float2 (&g)[PX] = gbuf[hb];
        float (&other)[PX][CR] = obuf[hb];
        const int y0 = (j0 + jj) * RB;
        const int wy0 = y0 - HALO_Y;
        const int pix0 = (y0 + ty) * W + x;
        STRIP_ASSERT(y0 >= 0 && y0 < H + RB && (j0 + jj) < G.nby);
        const int ky = -0x4B000000 - wy0;            // window row of a tap: mantissa bits of (iy + 2^23) + ky
        const int rb0 = b0 * RB;                     // ring row of the window's top row

        load_other(y0 + RB, jj + 1 < nb, obuf[hb ^ 1]);

        // ---- coordinates, weights and ring addresses of this thread's two pixels (registers only, branch-free)
        uint32_t aN[PX], aS[PX];
        float wnw[PX], wne[PX], wsw[PX], wse[PX];
        bool pdx[PX], pdy[PX], live[PX];
        bool outside = false;
#pragma unroll
        for (int r = 0; r < PX; ++r) {
          live[r] = FULLV || (vx && (y0 + ty + r * TROWS < H));
          // gs_source_index<Nm>(.., align_corners=False) + clip_coordinates, as in fuvs_common.cuh
          const float ix = fminf(Wm1, fmaxf(__fmul_rn(__fmaf_rn(__fadd_rn(g[r].x, 1.f), Wf, -1.f), 0.5f), 0.f));
          const float iy = fminf(Hm1, fmaxf(__fmul_rn(__fmaf_rn(__fadd_rn(g[r].y, 1.f), Hf, -1.f), 0.5f), 0.f));
          // floor and float->int without the conversion pipe: t = i + 2^23 rounded down carries floor(i) in its mantissa
          const float tfx = __fadd_rd(ix, 8388608.f), tfy = __fadd_rd(iy, 8388608.f);
          const float fx = __fsub_rn(tfx, 8388608.f), fy = __fsub_rn(tfy, 8388608.f);
          const float ax = __fsub_rn(__fadd_rn(fx, 1.f), ix), bx = __fsub_rn(ix, fx);
          const float ay = __fsub_rn(__fadd_rn(fy, 1.f), iy), by = __fsub_rn(iy, fy);
          wnw[r] = __fmul_rn(ax, ay); wne[r] = __fmul_rn(bx, ay); wsw[r] = __fmul_rn(ax, by); wse[r] = __fmul_rn(bx, by);
          pdx[r] = ix < Wm1;                         // the east neighbour exists  (ix_nw + 1 < W)
          pdy[r] = iy < Hm1;                         // the south neighbour exists (iy_nw + 1 < H)
          const unsigned lx = static_cast<unsigned>(__float_as_int(tfx) + kx);
          const unsigned d = static_cast<unsigned>(__float_as_int(tfy) + ky);
          // the 2x2 footprint inside the window: lx in [0, BOXW-2], d in [0, WIN_ROWS-2].  A tap clipped to the right
          // image edge (no east neighbour) may sit in the window's last column: the east address then points at the
          // next row / the pad behind the ring and its value is never accumulated.
          const bool in_box = (lx <= static_cast<unsigned>(pdx[r] ? BOXW - 2 : BOXW - 1)) &&
                              (d <= static_cast<unsigned>(WIN_ROWS - 2));
          if (live[r] && !in_box) outside = true;
          unsigned rr = d + static_cast<unsigned>(rb0);
          rr = min(rr, rr - static_cast<unsigned>(NSR));            // wrap: rr - NSR is huge when rr < NSR
          // ring element (row, column) of the north-west tap and of the one below it; byte addresses are formed at
          // the loads (x4 planar, x16 and x4 for the two parts of a 4+1 window)
          const unsigned e = rr * BOXW + lx;
          STRIP_ASSERT(!in_box || (rr < static_cast<unsigned>(NSR) && lx < static_cast<unsigned>(BOXW)));
          aN[r] = in_box ? e : 0u;
          aS[r] = in_box ? ((rr == static_cast<unsigned>(NSR - 1)) ? e - static_cast<unsigned>((NSR - 1) * BOXW) : e + BOXW)
                         : 0u;
          STRIP_ASSERT(aN[r] < static_cast<unsigned>(NSR * BOXW) && aS[r] < static_cast<unsigned>(NSR * BOXW));
          // the east tap may sit one element further (pad / next region), never beyond the 64 bytes behind the ring
          STRIP_ASSERT(pdx[r] ? (aN[r] % BOXW) + 1u < static_cast<unsigned>(BOXW) || !in_box : true);
        }
        // ---- this set's flow vectors are consumed: refill it for block jj + 2 (nothing here depends on the ring)
        load_grid(y0 + 2 * RB, jj + 2 < nb, g);
        const bool use_global = __any_sync(0xffffffffu, outside);
        refill_if_last();

        STRIP_JITTER(0x9e3779b9u);
        // ---- the ring slots of this block's window: the newest one, and all of them at the start of a segment
        if (jj == 0) {
#pragma unroll
          for (int i = 0; i < WIN - 1; ++i) {
            const int b = b0 + i;
            const bool wrap = b >= NS;
            mbar_wait(bar0 + 8u * (wrap ? b - NS : b), q0 ^ (wrap ? 1u : 0u));
          }
        }
        {
          const int b = b0 + WIN - 1;
          const bool wrap = b >= NS;
          mbar_wait(bar0 + 8u * (wrap ? b - NS : b), q0 ^ (wrap ? 1u : 0u));
        }

#ifdef FUVS_STRIP_TIMING
        if (kglob == 0) tm_first = clock64();
#endif
        if (use_global) {
          unsigned valid = 0u;
#pragma unroll
          for (int r = 0; r < PX; ++r) valid |= (live[r] ? 1u : 0u) << r;
          block_from_global(&A, side, EMIT, do_key0, C, H, W, x, y0 + ty, valid, IL, SIL);
        } else {
          const bool w_dst = WDST && dst != nullptr;
          const bool w_lp = EMIT && logit_out != nullptr;
          const bool w_l0 = KEY0 && A.logit0 != nullptr;
          // own pixels inside the ring (frame 0 of step 1): window rows HALO_Y + ty + r*TROWS
          uint32_t akey[PX];
          if (KEY0) {
#pragma unroll
            for (int r = 0; r < PX; ++r) {
              unsigned rr = static_cast<unsigned>(rb0 + HALO_Y + ty + r * TROWS);
              rr = min(rr, rr - static_cast<unsigned>(NSR));
              akey[r] = SIL ? rr * BOXW + (colbase - ring0) / 4u : colbase + rr * (BOXW * 4u);   // 4+1: element index
            }
          }
          ArgMax am[PX];
#pragma unroll
          for (int r = 0; r < PX; ++r) am[r].init(-INFINITY);   // push(v, 0) then always selects class 0 first
          if constexpr (CT > 0) {
            constexpr int CHB = NSC * RB * BOXW * 4;             // bytes between channels of the ring
            // ---- one burst: every tap of both pixels (4 CT loads each), then the key-frame values
            float v[PX][CT][4];
            float vk[PX][CT];
            if constexpr (SIL) {
              static_for<PX>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const uint32_t qn = ring0 + aN[r] * 16u, qs = ring0 + aS[r] * 16u;
                const uint32_t pn = ring0 + 4u * CHB + aN[r] * 4u, ps = ring0 + 4u * CHB + aS[r] * 4u;
                const float4 t0 = lds4_imm<0>(qn), t1 = lds4_imm<16>(qn), t2 = lds4_imm<0>(qs), t3 = lds4_imm<16>(qs);
                v[r][0][0] = t0.x; v[r][1][0] = t0.y; v[r][2][0] = t0.z; v[r][3][0] = t0.w;
                v[r][0][1] = t1.x; v[r][1][1] = t1.y; v[r][2][1] = t1.z; v[r][3][1] = t1.w;
                v[r][0][2] = t2.x; v[r][1][2] = t2.y; v[r][2][2] = t2.z; v[r][3][2] = t2.w;
                v[r][0][3] = t3.x; v[r][1][3] = t3.y; v[r][2][3] = t3.z; v[r][3][3] = t3.w;
                v[r][4][0] = lds_imm<0>(pn);
                v[r][4][1] = lds_imm<4>(pn);
                v[r][4][2] = lds_imm<0>(ps);
                v[r][4][3] = lds_imm<4>(ps);
              });
            } else {
              static_for<PX>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const uint32_t an = ring0 + aN[r] * 4u, as = ring0 + aS[r] * 4u;
                static_for<CT>([&](auto c_) {
                  constexpr int c = decltype(c_)::value;
                  v[r][c][0] = lds_imm<c * CHB>(an);
                  v[r][c][1] = lds_imm<c * CHB + 4>(an);
                  v[r][c][2] = lds_imm<c * CHB>(as);
                  v[r][c][3] = lds_imm<c * CHB + 4>(as);
                });
              });
            }
            if (KEY0 && do_key0) {
              static_for<PX>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                if constexpr (SIL) {
                  const float4 t = lds4_imm<0>(ring0 + akey[r] * 16u);
                  vk[r][0] = t.x; vk[r][1] = t.y; vk[r][2] = t.z; vk[r][3] = t.w;
                  vk[r][CT - 1] = lds_imm<0>(ring0 + 4u * CHB + akey[r] * 4u);
                } else {
                  static_for<CT>([&](auto c_) {
                    constexpr int c = decltype(c_)::value;
                    vk[r][c] = lds_imm<c * CHB>(akey[r]);
                  });
                }
              });
            }
            // ---- FMA chains in ATen's order (nw, ne, sw, se; neighbours outside the image are skipped), stores, blends
#pragma unroll
            for (int r = 0; r < PX; ++r) {
              const bool dxy = pdx[r] && pdy[r];
              float accs[CT];
              float vals[CT];                           // the frame's class values of this pixel (EMIT / frame 0)
#pragma unroll
              for (int c = 0; c < CT; ++c) {
                const int gi = c * HWi + pix0 + r * rstride;
                float acc = tap_acc<Nm>(0.f, v[r][c][0], wnw[r]);
                if (pdx[r]) acc = tap_acc<Nm>(acc, v[r][c][1], wne[r]);
                if (pdy[r]) acc = tap_acc<Nm>(acc, v[r][c][2], wsw[r]);
                if (dxy) acc = tap_acc<Nm>(acc, v[r][c][3], wse[r]);
                accs[c] = acc;
                if (!IL && w_dst && live[r]) dst[gi] = acc;
                if (EMIT) {
                  // fl(fl(w_this*acc) + fl(w_point*o)): IEEE addition is commutative, so one operand order serves both
                  // sides (the reference adds the forward term first on both) and the code is not duplicated per side
                  const float val = blend2(w_this, acc, w_point, other[r][c]);
                  vals[c] = val;
                  if (w_lp && live[r]) __stcs(logit_out + gi, val);
                }
                if (KEY0 && do_key0) {
                  vals[c] = vk[r][c];
                  if (w_l0 && live[r]) __stcs(A.logit0 + gi, vk[r][c]);
                }
              }
              if (EMIT || (KEY0 && do_key0)) am[r].idx = argmax_classes<CT>(vals);
              if constexpr (IL) {
                if (w_dst && live[r]) {
                  const int pix = pix0 + r * rstride;
                  reinterpret_cast<float4*>(dst)[pix] = make_float4(accs[0], accs[1], accs[2], accs[3]);
                  dst[4 * HWi + pix] = accs[CT - 1];
                }
              }
            }
          } else {
            const uint32_t chb = static_cast<uint32_t>(chan_floats) * 4u;
#pragma unroll
            for (int r = 0; r < PX; ++r) {
              const bool dxy = pdx[r] && pdy[r];
              for (int c = 0; c < C; ++c) {
                const uint32_t co = static_cast<uint32_t>(c) * chb;
                const uint32_t an = ring0 + aN[r] * 4u + co, as = ring0 + aS[r] * 4u + co;
                const float v00 = lds_rt(an), v01 = lds_rt(an + 4u);
                const float v10 = lds_rt(as), v11 = lds_rt(as + 4u);
                const int gi = c * HWi + pix0 + r * rstride;
                float acc = tap_acc<Nm>(0.f, v00, wnw[r]);
                if (pdx[r]) acc = tap_acc<Nm>(acc, v01, wne[r]);
                if (pdy[r]) acc = tap_acc<Nm>(acc, v10, wsw[r]);
                if (dxy) acc = tap_acc<Nm>(acc, v11, wse[r]);
                if (w_dst && live[r]) dst[gi] = acc;
                if (EMIT) {
                  const float o = live[r] ? __ldg(point + gi) : 0.f;
                  const float val = blend2(w_this, acc, w_point, o);
                  am[r].push(val, c);
                  if (w_lp && live[r]) __stcs(logit_out + gi, val);
                }
                if (KEY0 && do_key0) {
                  const float val = lds_rt(akey[r] + co);
                  am[r].push(val, c);
                  if (w_l0 && live[r]) __stcs(A.logit0 + gi, val);
                }
              }
            }
          }
#pragma unroll
          for (int r = 0; r < PX; ++r) {
            if (!live[r]) continue;
            const int pix = pix0 + r * rstride;
            STRIP_ASSERT(pix >= 0 && pix < HWi && am[r].idx >= 0 && am[r].idx < C);
            if (EMIT && lab_out) lab_out[pix] = static_cast<uint8_t>(am[r].idx);
            if (KEY0 && do_key0 && A.label0) A.label0[pix] = static_cast<uint8_t>(am[r].idx);
          }
        }

        // ---- this warp is done with the ring for block kglob (its gathers have returned: the FMAs consumed them).
        // done[kglob & 3] only grows: visit v = kglob >> 2 of a counter is complete at (v + 1) * NWARPS; the skew
        // between warps is bounded by the prefetch depth (< 4 blocks), so visits never mix.
        STRIP_JITTER(0x85ebca6bu);
        __syncwarp();
        tok_from = released;
        tok_to = (jj < nb - 1) ? sb + jj + 1 : sb + nb + WIN - 1;
        released = tok_to;
        tok_want = (static_cast<unsigned>(kglob >> 2) + 1u) * NWARPS - 1u;
        if (lane == 0) {
          tok = atom_add_relaxed(done0 + 4u * (kglob & 3), 1u);
          // visits of a counter never mix: the old value lies inside this visit's window of NWARPS arrivals
          STRIP_ASSERT(tok >= static_cast<unsigned>(kglob >> 2) * NWARPS && tok <= tok_want);
        }
        ++kglob;
        if (++b0 == NS) { b0 = 0; q0 ^= 1u; }
      }
    }
    sb += nb + WIN - 1;
    left -= nb;
    ++u;
    j0 = 0;
  }
  refill_if_last();                                // nothing left to load for this CTA: issue() ignores indices past the end
#ifdef FUVS_STRIP_TIMING
  if (g_strip_timing && lane == 0) {   // per warp: start, wait passed, first block's window complete, end, blocks
    long long* o = g_strip_timing + ((static_cast<long long>(G.seq & 63) * MAX_GRID + blockIdx.x) * NWARPS + (tid >> 5)) * 5;
    o[0] = tm_start; o[1] = tm_wait; o[2] = tm_first; o[3] = clock64(); o[4] = B1 - B0;
  }
#endif
}

template <int CT, int NSC, bool EMIT, bool KEY0, bool WDST, bool FULLV, bool IL, bool KIL = false>
int launch_variant(const StripMaps& maps, const DenseStep& a, int C, int H, int W, int nslot, cudaStream_t st) {
  static SmemOptIn optin;
  auto kern = dense_strip_kernel<CT, NSC, EMIT, KEY0, WDST, FULLV, IL, KIL>;
  const size_t smem = static_cast<size_t>(nslot) * C * PLANE * 4 + 256;
  if (!optin.ensure(kern, SMEM_LIMIT)) return set_error(FUVS_ECUDA, "fuvs_dense_interval(strip step): shared-memory opt-in failed");
  StripGeom g;
  g.nsx = (W + TW - 1) / TW;
  g.nby = (H + RB - 1) / RB;
  g.total = 2 * g.nsx * g.nby;
  g.nslot = nslot;
#ifdef FUVS_STRIP_TIMING
  g.seq = g_strip_seq++;
#endif
  int grid = sm_count();                       // persistent: one CTA per SM
  if (grid > g.total) grid = g.total;
  if (grid > MAX_GRID) grid = MAX_GRID;
  // weights measured in r01: forward side of step 1 (eighths) 8 -> 244.6, 9 -> 242.7, 10 -> 245.3 us per interval; edge
  // strips 18 sixteenths; strip start 0 -> 249.1, 16 -> 245.1, 32 -> 248.6 us
  StripPartition part;
  strip_partition(g, grid, KEY0 ? 9 : 8, 8, 18, 16, &part);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, a, C, H, W, g, part);
  if (e != cudaSuccess) return set_error(FUVS_ECUDA, "fuvs_dense_interval(strip step): %s", cudaGetErrorString(e));
  count_launch();
  return FUVS_OK;
}

template <int CT, int NSC, bool FULLV, bool IL>
int launch_ct(const StripMaps& maps, const DenseStep& a, int C, int H, int W, int nslot, cudaStream_t st) {
  const bool wdst = a.dstL != nullptr;
  if (a.emitA) {
    return wdst ? launch_variant<CT, NSC, true, false, true, FULLV, IL>(maps, a, C, H, W, nslot, st)
                : launch_variant<CT, NSC, true, false, false, FULLV, IL>(maps, a, C, H, W, nslot, st);
  }
  if constexpr (IL) {
    if (a.key0 && a.key_il) return launch_variant<CT, NSC, false, true, true, FULLV, IL, true>(maps, a, C, H, W, nslot, st);
  }
  if (a.key0) return launch_variant<CT, NSC, false, true, true, FULLV, IL>(maps, a, C, H, W, nslot, st);
  return launch_variant<CT, NSC, false, false, true, FULLV, IL>(maps, a, C, H, W, nslot, st);
}

template <bool FULLV>
int launch_full(const StripMaps& maps, const DenseStep& a, int C, int H, int W, int nslot, cudaStream_t st) {
  switch (C) {
    case 2: return launch_ct<2, NSLOT_C2, FULLV, false>(maps, a, C, H, W, nslot, st);
    case 5:
      return a.il ? launch_ct<5, NSLOT_C5, FULLV, true>(maps, a, C, H, W, nslot, st)
                  : launch_ct<5, NSLOT_C5, FULLV, false>(maps, a, C, H, W, nslot, st);
    default: return launch_ct<0, 0, FULLV, false>(maps, a, C, H, W, nslot, st);
  }
}

bool shape_ok(int C, int H, int W) {
  return (W & 3) == 0 && W >= 4 && H < 32768 && W < 32768 && C <= 16 && static_cast<long long>(C) * H * W < (1ll << 31) &&
         nslot_for(C) >= WIN + 1;
}

}  // namespace

// Can every step of an n-frame interval run on the strip kernel with the chain states stored 4+1 (see `layout` above)?
// The caller (fuvs_dense_interval) decides once per interval: a state written 4+1 must be read 4+1.
bool dense_strip_il_ok(int C, int H, int W, int n, const float* prev, const float* next, const float* gridsL,
                       const float* gridsR, const float* scratch) {
  return C == 5 && n >= 3 && (n & 1) == 1 && shape_ok(C, H, W) && aligned16(prev) && aligned16(next) &&
         aligned16(scratch) && aligned8(gridsL) && aligned8(gridsR) && get_encode() != nullptr;
}

// Returns FUVS_OK if it ran the step, 1 if the shape is not eligible (the caller falls back; never for a.il steps, whose
// eligibility dense_strip_il_ok has established), negative on error.
int launch_dense_step_strip(const DenseStep& a, int C, int H, int W, cudaStream_t st) {
  // not eligible: the caller falls back to another kernel, which is not possible for a step whose states are 4+1
  auto ineligible = [&]() {
    return a.il ? set_error(FUVS_EINVAL, "fuvs_dense_interval: a 4+1 step is not eligible for the strip kernel") : 1;
  };
  if (!shape_ok(C, H, W) || !aligned16(a.srcL) || !aligned16(a.srcR)) return ineligible();
  if ((a.emitA && !a.pointR) || (a.emitB && !a.pointL) || (a.emitA != a.emitB)) return ineligible();
  if (a.key0 && (a.key0 != a.srcL || a.emitA)) return ineligible();
  if ((a.dstL != nullptr) != (a.dstR != nullptr)) return ineligible();
  if (!a.emitA && !a.dstL) return ineligible();           // a step that neither writes states nor emits frames does not exist
  if (!aligned8(a.gridL) || !aligned8(a.gridR)) return ineligible();
  if (a.il && C != 5) return ineligible();
  const int nslot = nslot_for(C);
  StripMaps maps;
  if (a.key_il && !(a.il && a.key0)) return ineligible();
  const bool sil = a.il && (!a.key0 || a.key_il);
  const long long HW = static_cast<long long>(H) * W;
  if (sil) {
    // 4+1 source: channels 0-3 interleaved, channel 4 a plane behind them
    if (!make_map_chw(&maps.srcL, a.srcL + 4 * HW, 1, H, W, BOXW, RB, 1) ||
        !make_map_chw(&maps.srcR, a.srcR + 4 * HW, 1, H, W, BOXW, RB, 1))
      return ineligible();
  } else {
    if (!make_map_chw(&maps.srcL, a.srcL, C, H, W, BOXW, RB, 1) || !make_map_chw(&maps.srcR, a.srcR, C, H, W, BOXW, RB, 1))
      return ineligible();
  }
  if ((W % TW) == 0 && (H % RB) == 0) return launch_full<true>(maps, a, C, H, W, nslot, st);
  return launch_full<false>(maps, a, C, H, W, nslot, st);
}

}  // namespace fuvs

#ifdef FUVS_STRIP_TIMING
extern "C" __attribute__((visibility("default"))) int fuvs_dev_set_strip_timing(long long* buf) {
  return cudaMemcpyToSymbol(fuvs::g_strip_timing, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#endif
