// dense_strip.cu — dense-flow warp step as a column-strip sliding window (sm_100a).
//
// Why: the per-plane TMA kernel (dense_tma.cu) stages, for every 128x16 output tile, a 192x48 source box per channel:
// every source byte crosses the L2->SM fabric 4.5 times (409 MB per step at 1080p, profiles/r01_ncu_dense_tma_v5.txt),
// which is what bounds it.  Here a CTA owns a run of consecutive 8-row blocks of one 128-pixel column strip and keeps
// the source window of ALL channels resident in shared memory as a ring of 8-row slots.  Moving one block down the
// strip costs one new slot (192 x 8 x C floats, one cp.async.bulk.tensor.3d), so a source byte crosses the fabric
// ~1.5 times (x-halo only) plus a 32-row warm-up per run.
//
//   work     : 2 sides x ceil(W/128) strips x ceil(H/8) blocks, numbered side-major/strip/row; CTA b of a persistent
//              grid (one CTA per SM) takes the contiguous range [b*T/G, (b+1)*T/G).  A range may cross a strip
//              boundary: it is then processed as several "segments", each with its own window warm-up.
//   ring     : NS slots (7 for C=5: 215 KB).  A block needs the 5 slots covering rows [y0-16, y0+24); the others are
//              prefetched slots of the following blocks.  Slot loads are numbered s = 0,1,2.. in the order of use;
//              load s lives in buffer s % NS and completes phase (s / NS) & 1 of mbarrier full[s % NS].
//   protocol : no block-wide barrier.  A warp that has finished block k bumps done[k & 3]; the LAST warp to do so
//              knows which slot loads are now dead and issues the loads that reuse their buffers.
//   taps     : shared-memory loads with immediate channel offsets; a warp whose taps leave the window (large motion)
//              gathers its pixels of that block from global memory instead, so any flow field stays correct.
//   counts   : stay in fuvs_temporal_counts (metric.cu).  Fusing them into the last step was measured: the byte-wide
//              label re-reads and ~40 extra instructions per pixel cost 38 us against 15 us for the separate kernel.
// Arithmetic is gs_setup/tap_acc/blend2 from fuvs_common.cuh: bit-identical to the direct kernel (warp.cu) and to
// ATen's grid_sampler_2d (flow/model.py:244-249).
#include <cstdlib>

#include "dense_common.cuh"
#include "tma_ptx.cuh"

namespace fuvs {

namespace {

using namespace tma;

constexpr int TW = 128;                      // strip width = threads per thread-row
constexpr int HALO_X = 32, HALO_Y = 16;
// Row stride 192 = 0 mod 32 banks on purpose: with coherent flow a warp's taps straddle at most two rows and stay
// conflict-free whatever the rows are.  A skewed stride (196) was measured: it spreads the taps that border clipping
// sends to one column (edge strips, iid flow: -1 %) but costs coherent flow 5 % (2-way conflicts at row changes).
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

#ifdef FUVS_STRIP_PROF
// developer instrumentation (not in the product build): per CTA, cycles thread 0 spent waiting for the ring vs in total
__device__ unsigned long long g_strip_prof[16 * 148 * 4];
__device__ unsigned g_strip_launch;
__device__ int g_strip_dbg;   // bit 0: no tap loads, 1: no state stores, 2: fixed taps, 3: no TMA / ring waits, 4: no grid loads
#endif

struct StripGeom {
  int nsx, nby, total, nslot;
  int wl, wr;                                // relative cost of a forward-side / backward-side block (CTA partition)
  int wedge;                                 // cost of a block of the two edge strips, in sixteenths
  int wstart;                                // extra cost of the first block of a strip, in eighths of a block
};
struct StripMaps {
  CUtensorMap srcL, srcR;                    // [C][H][W] fp32, box BOXW x RB x C
};

template <int PX>
struct BlockTaps {
  int aN[PX], aS[PX];                        // float offsets of the nw / sw taps inside the ring
  float wnw[PX], wne[PX], wsw[PX], wse[PX];
  unsigned dxm, dym, valid;
};

// Rare path: the warp computes its pixels of this block straight from global memory (same math as warp.cu).
// Returns the labels of the PX pixels packed 8 bits each (emitted frame; frame 0 for KEY0 launches).
template <class NM, int PX, int TROWS>
__device__ __noinline__ unsigned block_from_global(const DenseStep* __restrict__ Ap, bool side, bool emit, bool key0,
                                                   int C, int H, int W, int x, int ytop, unsigned valid) {
  const DenseStep& A = *Ap;
  const long long HW = static_cast<long long>(H) * W;
  const float* grid = side ? A.gridR : A.gridL;
  const float* src = side ? A.srcR : A.srcL;
  float* dst = side ? A.dstR : A.dstL;
  const float* point = side ? A.pointL : A.pointR;
  const float w_this = side ? A.wB1 : A.wA0, w_point = side ? A.wB0 : A.wA1;
  uint8_t* lab_out = side ? A.labelB : A.labelA;
  float* logit_out = side ? A.logitB : A.logitA;
  unsigned packed = 0u;
  for (int r = 0; r < PX; ++r) {
    if (!((valid >> r) & 1u)) continue;
    const long long pix = static_cast<long long>(ytop + r * TROWS) * W + x;
    const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + pix);
    const GsTap t = gs_setup<NM>(g.x, g.y, H, W, false);
    ArgMax am, am0;
    am.init(-INFINITY);
    am0.init(-INFINITY);
    for (int c = 0; c < C; ++c) {
      const long long o = c * HW + pix;
      const float acc = gs_fetch<NM>(src + c * HW, t, W);
      if (dst) dst[o] = acc;
      if (emit) {
        const float other = __ldg(point + o);
        const float v = side ? blend2(w_point, other, w_this, acc) : blend2(w_this, acc, w_point, other);
        am.push(v, c);
        if (logit_out) __stcs(logit_out + o, v);
      }
      if (key0) {
        const float v = __ldg(A.key0 + o);
        am0.push(v, c);
        if (A.logit0) __stcs(A.logit0 + o, v);
      }
    }
    if (emit && lab_out) lab_out[pix] = static_cast<uint8_t>(am.idx);
    if (key0 && A.label0) A.label0[pix] = static_cast<uint8_t>(am0.idx);
    packed |= static_cast<unsigned>(emit ? am.idx : am0.idx) << (8 * r);
  }
  return packed;
}

//   EMIT   : both sides complete a frame this step (blend with the other chain's state read pointwise, arg-max)
//   KEY0   : step 1: the forward side also emits frame 0 = arg-max of the key frame (read from the ring itself)
//   NSC    : ring slots when known at compile time (0: G.nslot)
template <class NM, int CT, int NSC, int TROWS, bool EMIT, bool KEY0>
__global__ void __launch_bounds__(TW * TROWS, 1)
dense_strip_kernel(const __grid_constant__ StripMaps M, const __grid_constant__ DenseStep A, int Crt, int H, int W,
                   StripGeom G) {
  static_assert(!(EMIT && KEY0), "frame 0 and a completed frame never share a step here (the host routes n<=2 elsewhere)");
  constexpr int THREADS = TW * TROWS;
  constexpr int PX = RB / TROWS;
  constexpr int NWARPS = THREADS / 32;
  constexpr int CR = CT > 0 ? CT : 1;
  constexpr bool FULLV = false;   // pixel validity is tested per store (partial strips / blocks at the image edge)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int C = CT > 0 ? CT : Crt;
  const int slot_floats = C * PLANE;
  const uint32_t slot_bytes = static_cast<uint32_t>(slot_floats) * 4u;
  const int NS = NSC > 0 ? NSC : G.nslot;
  float* bufs = reinterpret_cast<float*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + static_cast<size_t>(NS) * slot_bytes);
  unsigned* done = reinterpret_cast<unsigned*>(bars + MAX_SLOTS);     // 4 counters
  const uint32_t bar0 = smem_u32(bars);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int tx = tid & (TW - 1), ty = tid >> 7;                        // TW == 128
  const int HWi = H * W;                                               // the host guarantees C*H*W < 2^31
  // CTA b takes the blocks whose cumulative cost lies in [b, b+1) * total_cost / grid.  A block's cost is
  // side weight (step 1: the forward side also emits frame 0) x strip weight (the two strips at the image edge cost
  // ~12 % more: border clipping sends many lanes to the same column, i.e. the same shared-memory bank).
  auto unit_w = [&](int u) {
    const bool side = u >= G.nsx;
    const int strip = side ? u - G.nsx : u;
    const bool edge = (G.nsx > 2) && (strip == 0 || strip == G.nsx - 1);
    return (side ? G.wr : G.wl) * (edge ? G.wedge : 16);
  };
  // A CTA whose range crosses into a new strip has to refill its whole window there (WIN slots = 192 KB instead of
  // one slot per block): measured, those CTAs ran 11 % longer than the mean and set the launch time
  // (tools/strip_prof.py).  The first block of every strip therefore carries an extra cost of G.wstart / 8 blocks,
  // which hands the CTA that owns it correspondingly fewer blocks.
  auto unit_start = [&](int u) { return (static_cast<long long>(G.wstart) * unit_w(u)) >> 3; };
  long long cost_all = 0;
  for (int u = 0; u < 2 * G.nsx; ++u) cost_all += static_cast<long long>(G.nby) * unit_w(u) + unit_start(u);
  auto block_at = [&](long long num) {        // first block whose start cost is >= num * cost_all / grid
    long long pos = (num * cost_all + gridDim.x - 1) / gridDim.x;
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
  const int B0 = block_at(blockIdx.x);
  const int B1 = (blockIdx.x + 1 == gridDim.x) ? G.total : block_at(blockIdx.x + 1);

#ifdef FUVS_STRIP_PROF
  const int dbg0 = g_strip_dbg;
#else
  constexpr int dbg0 = 0;
#endif
  // Programmatic dependent launch: this CTA may have been scheduled while the previous kernel of the stream (the
  // step that produced our source states) is still draining.  Let our own successor do the same, set up the
  // barriers, then wait for the predecessor's memory to be complete before the first global access.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    for (int b = 0; b < MAX_SLOTS; ++b) mbar_init(bar0 + 8u * b, 1);
    for (int b = 0; b < 4; ++b) done[b] = 0u;
    fence_barrier_init();
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // Left edge of a strip's window.  Kept inside the image where the image is wide enough: a box that hangs over
  // the image edge is zero-filled by TMA but loads far slower (measured: the CTAs of the last strip waited 50k
  // cycles per launch for their ring), and border clipping never reads beyond the edge anyway.
  auto box_x = [&](int strip) { return max(0, min(strip * TW - HALO_X, W - BOXW)); };
  // one thread: start slot load s of this CTA's sequence (its buffer is known to be dead)
  auto issue = [&](int s) {
    int kb = B0, sb = 0;
    while (kb < B1) {
      const int u = kb / G.nby, j0 = kb - u * G.nby;
      const int nb = min(G.nby - j0, B1 - kb);
      if (s < sb + nb + WIN - 1) {
        const int ytop = (j0 + (s - sb)) * RB - HALO_Y;
        const bool side = u >= G.nsx;
        const int strip = side ? u - G.nsx : u;
        const int b = s % NS;
        const uint32_t bar = bar0 + 8u * b;
        if (ytop + RB <= 0 || ytop >= H || (dbg0 & 8)) {
          mbar_arrive(bar);                    // slot entirely outside the image: never read (border clipping)
        } else {
          mbar_expect_tx(bar, slot_bytes);
          load_3d(smem_u32(bufs + static_cast<size_t>(b) * slot_floats), side ? &M.srcR : &M.srcL, box_x(strip), ytop, 0,
                  bar);
        }
        return;
      }
      sb += nb + WIN - 1;
      kb += nb;
    }
  };
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) issue(s);
  }

#ifdef FUVS_STRIP_PROF
  long long prof_wait = 0, prof_first = 0;
  const unsigned prof_launch = *reinterpret_cast<volatile unsigned*>(&g_strip_launch) & 15u;
  const long long prof_t0 = clock64();
#endif

  int kb = B0, sb = 0, kglob = 0, released = 0;
  while (kb < B1) {                              // ---- one segment: nb consecutive blocks of one (side, strip)
    const int u = kb / G.nby, j0 = kb - u * G.nby;
    const int nb = min(G.nby - j0, B1 - kb);
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
    const bool vx = x < W;

    // Flow vectors are prefetched TWO blocks ahead and the pointwise operand of emitting steps ONE block ahead: a
    // block takes ~2.5k cycles, less than a loaded HBM round trip (with one block of prefetch the first use of the
    // vector was the top stall of the kernel, profiles/r01_ncu_dense_strip_smooth.txt).
    auto load_grid = [&](int yb, bool on, float2 (&dstg)[PX]) {
#pragma unroll
      for (int r = 0; r < PX; ++r) {
        const int y = yb + ty + r * TROWS;
        dstg[r] = (on && vx && y < H && !(dbg0 & 16)) ? __ldg(grid + y * W + x) : make_float2(0.f, 0.f);
      }
    };
    auto load_other = [&](int yb, bool on, float (&dsto)[PX][CR]) {
      if (!(EMIT && CT > 0)) return;
#pragma unroll
      for (int r = 0; r < PX; ++r) {
        const int y = yb + ty + r * TROWS;
        const bool live = on && vx && y < H;
#pragma unroll
        for (int c = 0; c < CR; ++c) dsto[r][c] = live ? __ldg(point + (c * HWi + y * W + x)) : 0.f;
      }
    };
    float2 g[PX], g1[PX];
    float other[PX][CR];
    load_grid(j0 * RB, true, g);
    load_grid(j0 * RB + RB, nb > 1, g1);
    load_other(j0 * RB, true, other);
    // ring position of the window's top slot (slot load sb + jj): buffer b0, phase parity q0
    int b0 = sb % NS;
    unsigned q0 = static_cast<unsigned>(sb / NS) & 1u;

    for (int jj = 0; jj < nb; ++jj, ++kglob) {
      const int y0 = (j0 + jj) * RB;
      const int wy0 = y0 - HALO_Y;
      const int pix0 = (y0 + ty) * W + x;

      // ---- taps of this thread's PX pixels (registers only, branch-free)
      BlockTaps<PX> T;
      T.dxm = 0u; T.dym = 0u; T.valid = 0u;
      unsigned outm = 0u;
#pragma unroll
      for (int r = 0; r < PX; ++r) {
        const int y = y0 + ty + r * TROWS;
        const bool live = vx && (y < H);
        const GsTap t = gs_setup<NM>(g[r].x, g[r].y, H, W, false);
        T.wnw[r] = t.nw; T.wne[r] = t.ne; T.wsw[r] = t.sw; T.wse[r] = t.se;
        const int lx = t.ix - xbase, d = t.iy - wy0;
        // the taps that are accumulated must lie in the window: lx in [0, BOXW-1-dx], d in [0, WIN_ROWS-1-dy]
        const bool in_box = (static_cast<unsigned>(lx) <= static_cast<unsigned>(BOXW - 1 - t.dx)) &&
                            (static_cast<unsigned>(d) <= static_cast<unsigned>(WIN_ROWS - 1 - t.dy));
        int sN = b0 + (d >> 3);
        sN -= (sN >= NS) ? NS : 0;
        int sS = b0 + ((d + 1) >> 3);
        sS -= (sS >= NS) ? NS : 0;
        const int aN = sN * slot_floats + (d & 7) * BOXW + lx;
        const int aS = sS * slot_floats + ((d + 1) & 7) * BOXW + lx;
        T.aN[r] = in_box ? aN : 0;
        T.aS[r] = in_box ? aS : 0;
        if (dbg0 & 4) { T.aN[r] = tid + r * BOXW; T.aS[r] = tid + (r + 1) * BOXW; T.wnw[r] = T.wne[r] = T.wsw[r] = T.wse[r] = 0.25f; }
        if (live && !in_box) outm |= 1u << r;
        if (live) T.valid |= 1u << r;
        if (t.dx) T.dxm |= 1u << r;
        if (t.dy) T.dym |= 1u << r;
      }
      // ---- prefetches for the following blocks (registers only; nothing here depends on the ring)
      float2 g2[PX];
      float other_next[PX][CR];
      load_grid(y0 + 2 * RB, jj + 2 < nb, g2);
      load_other(y0 + RB, jj + 1 < nb, other_next);
      const bool use_global = __any_sync(0xffffffffu, outm != 0u);

      // ---- the ring slots of this block's window: the newest one, and all of them at the start of a segment
#ifdef FUVS_STRIP_PROF
      const long long prof_w0 = clock64();
#endif
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

#ifdef FUVS_STRIP_PROF
      if (kglob == 0) prof_first = clock64() - prof_w0; else prof_wait += clock64() - prof_w0;
#endif
      unsigned labs = 0u;                        // labels of this thread's pixels, 8 bits each
      if (use_global) {
        labs = block_from_global<NM, PX, TROWS>(&A, side, EMIT, do_key0, C, H, W, x, y0 + ty, T.valid);
      } else {
        ArgMax am[PX];
#pragma unroll
        for (int r = 0; r < PX; ++r) am[r].init(-INFINITY);   // push(v, 0) then always selects class 0 first
        int sK = b0 + HALO_Y / RB;
        sK -= (sK >= NS) ? NS : 0;
        const int key_loc = sK * slot_floats + ty * BOXW + (min(x, W - 1) - xbase);
        const bool w_dst = dst != nullptr && !(dbg0 & 2);
        const bool w_lp = EMIT && logit_out != nullptr;
        const bool w_l0 = KEY0 && A.logit0 != nullptr;
        // One code path for interior and border pixels: ATen skips the taps whose neighbour is outside the image
        // (within_bounds_2d), which is a predicated FFMA here — the tap value is loaded regardless (its address
        // stays inside the ring) and simply not accumulated.  A separate predicated slow path made the CTAs of the
        // edge strips, where border clipping is common, pace the whole launch.
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float* pl = bufs + c * PLANE;
          float v00[PX], v01[PX], v10[PX], v11[PX];
#pragma unroll
          for (int r = 0; r < PX; ++r) {
            if (dbg0 & 1) { v00[r] = v01[r] = v10[r] = v11[r] = __int_as_float(T.aN[r] + c); continue; }
            v00[r] = pl[T.aN[r]]; v01[r] = pl[T.aN[r] + 1];
            v10[r] = pl[T.aS[r]]; v11[r] = pl[T.aS[r] + 1];
          }
#pragma unroll
          for (int r = 0; r < PX; ++r) {
            const bool live = FULLV || ((T.valid >> r) & 1u);
            const bool dx = (T.dxm >> r) & 1u, dy = (T.dym >> r) & 1u;
            const int gi = c * HWi + pix0 + r * rstride;
            float acc = 0.f;
            acc = tap_acc<NM>(acc, v00[r], T.wnw[r]);
            if (dx) acc = tap_acc<NM>(acc, v01[r], T.wne[r]);
            if (dy) acc = tap_acc<NM>(acc, v10[r], T.wsw[r]);
            if (dx && dy) acc = tap_acc<NM>(acc, v11[r], T.wse[r]);
            if (w_dst && live) dst[gi] = acc;
            if (EMIT) {
              const float o = CT > 0 ? other[r][CT > 0 ? c : 0] : (live ? __ldg(point + gi) : 0.f);
              // fl(fl(w_this*acc) + fl(w_point*o)): IEEE addition is commutative, so one operand order serves both
              // sides (the reference adds the forward term first on both) and the code is not duplicated per side
              const float v = blend2(w_this, acc, w_point, o);
              am[r].push(v, c);
              if (w_lp && live) __stcs(logit_out + gi, v);
            }
            if (KEY0 && do_key0) {
              const float v = pl[key_loc + r * (TROWS * BOXW)];
              am[r].push(v, c);
              if (w_l0 && live) __stcs(A.logit0 + gi, v);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < PX; ++r) {
          labs |= static_cast<unsigned>(am[r].idx) << (8 * r);
          if (!((T.valid >> r) & 1u)) continue;
          const int pix = pix0 + r * rstride;
          if (EMIT && lab_out) lab_out[pix] = static_cast<uint8_t>(am[r].idx);
          if (KEY0 && do_key0 && A.label0) A.label0[pix] = static_cast<uint8_t>(am[r].idx);
        }
      }

      // ---- this warp is done with the ring for block kglob; the last warp to say so refills the dead buffers
      const int rel_after = (jj < nb - 1) ? sb + jj + 1 : sb + nb + WIN - 1;
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&done[kglob & 3], 1u) == NWARPS - 1) {
          done[kglob & 3] = 0u;          // published to the other warps by the release-arrive inside issue()
          for (int s = released + NS; s < rel_after + NS; ++s) issue(s);
        }
      }
      released = rel_after;

#pragma unroll
      for (int r = 0; r < PX; ++r) {
        g[r] = g1[r];
        g1[r] = g2[r];
#pragma unroll
        for (int c = 0; c < CR; ++c) other[r][c] = other_next[r][c];
      }
      if (++b0 == NS) { b0 = 0; q0 ^= 1u; }
    }
    sb += nb + WIN - 1;
    kb += nb;
  }
#ifdef FUVS_STRIP_PROF
  if (tid == 0 && blockIdx.x < 148) {
    unsigned long long* o = g_strip_prof + (prof_launch * 148 + blockIdx.x) * 4;
    o[0] = clock64() - prof_t0;
    o[1] = prof_wait;
    o[2] = prof_first;
    o[3] = B1 - B0;
    if (blockIdx.x == 0) atomicAdd(&g_strip_launch, 1u);
  }
#endif
}

template <int CT, int NSC, int TROWS, bool EMIT, bool KEY0>
int launch_variant(const StripMaps& maps, const DenseStep& a, int C, int H, int W, int nslot, cudaStream_t st) {
  static SmemOptIn optin;
  auto kern = dense_strip_kernel<Nm, CT, NSC, TROWS, EMIT, KEY0>;
  const size_t smem = static_cast<size_t>(nslot) * C * PLANE * 4 + 256;
  if (!optin.ensure(kern, SMEM_LIMIT)) return 1;
  StripGeom g;
  g.nsx = (W + TW - 1) / TW;
  g.nby = (H + RB - 1) / RB;
  g.total = 2 * g.nsx * g.nby;
  g.nslot = nslot;
  static const int wkey0 = []() { const char* e = getenv("FUVS_STRIP_WKEY0"); return e ? atoi(e) : 9; }();
  g.wl = KEY0 ? wkey0 : 8;                     // measured (eighths): 8 -> 244.6, 9 -> 242.7, 10 -> 245.3, 11 -> 247.6 us per interval
  g.wr = 8;
  static const int wedge = []() { const char* e = getenv("FUVS_STRIP_WEDGE"); return e ? atoi(e) : 18; }();   // sixteenths
  g.wedge = wedge;
  static const int wstart = []() { const char* e = getenv("FUVS_STRIP_WSTART"); return e ? atoi(e) : 16; }();   // measured: 0 -> 249.1, 16 -> 245.1, 32 -> 248.6 us
  g.wstart = wstart < 0 ? 0 : wstart;
  int grid = sm_count();                       // persistent: one CTA per SM
  if (grid > g.total) grid = g.total;
  static const bool pdl = []() { const char* e = getenv("FUVS_STRIP_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TW * TROWS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, kern, maps, a, C, H, W, g) != cudaSuccess) return check_launch("fuvs_dense_interval(strip step)");
  return check_launch("fuvs_dense_interval(strip step)");
}

template <int CT, int NSC, int TROWS>
int launch_ct(const StripMaps& maps, const DenseStep& a, int C, int H, int W, int nslot, cudaStream_t st) {
  if (a.emitA) return launch_variant<CT, NSC, TROWS, true, false>(maps, a, C, H, W, nslot, st);
  if (a.key0) return launch_variant<CT, NSC, TROWS, false, true>(maps, a, C, H, W, nslot, st);
  return launch_variant<CT, NSC, TROWS, false, false>(maps, a, C, H, W, nslot, st);
}

template <int TROWS>
int launch_rows(const StripMaps& maps, const DenseStep& a, int C, int H, int W, int nslot, cudaStream_t st) {
  switch (C) {
    case 2: return launch_ct<2, NSLOT_C2, TROWS>(maps, a, C, H, W, nslot, st);
    case 5: return launch_ct<5, NSLOT_C5, TROWS>(maps, a, C, H, W, nslot, st);
    default: return launch_ct<0, 0, TROWS>(maps, a, C, H, W, nslot, st);
  }
}

}  // namespace

// Returns FUVS_OK if it ran the step, 1 if the shape is not eligible (the caller falls back), negative on error.
int launch_dense_step_strip(const DenseStep& a, int C, int H, int W, cudaStream_t st) {
  if ((W & 3) != 0 || W < 4 || !aligned16(a.srcL) || !aligned16(a.srcR) || H >= 32768 || W >= 32768) return 1;
  if ((a.emitA && !a.pointR) || (a.emitB && !a.pointL) || (a.emitA != a.emitB)) return 1;
  if (a.key0 && (a.key0 != a.srcL || a.emitA)) return 1;
  if (!aligned8(a.gridL) || !aligned8(a.gridR)) return 1;
  if (C > 16 || static_cast<long long>(C) * H * W >= (1ll << 31)) return 1;
  const int nslot = nslot_for(C);
  if (nslot < WIN + 1) return 1;                 // the window of all channels does not fit: per-plane kernel
  StripMaps maps;
  if (!make_map_chw(&maps.srcL, a.srcL, C, H, W, BOXW, RB, C) || !make_map_chw(&maps.srcR, a.srcR, C, H, W, BOXW, RB, C))
    return 1;
#ifdef FUVS_STRIP_PROF
  {
    const char* e = getenv("FUVS_STRIP_DBG");
    const int v = e ? atoi(e) : 0;
    cudaMemcpyToSymbolAsync(g_strip_dbg, &v, sizeof(int), 0, cudaMemcpyHostToDevice, st);
  }
#endif
  static const int trows = []() {
    const char* e = getenv("FUVS_STRIP_TROWS");
    return (e && e[0] == '8') ? 8 : (e && e[0] == '2') ? 2 : 4;
  }();
  if (trows == 2) return launch_rows<2>(maps, a, C, H, W, nslot, st);
  if (trows == 8) return launch_rows<8>(maps, a, C, H, W, nslot, st);
  return launch_rows<4>(maps, a, C, H, W, nslot, st);
}

#ifdef FUVS_STRIP_PROF
extern "C" __attribute__((visibility("default"))) int fuvs_debug_strip_prof(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_strip_prof, sizeof(g_strip_prof)) == cudaSuccess ? 0 : -1;
}
#endif

}  // namespace fuvs
