// dense_tma.cu — dense-flow warp step with TMA-staged source tiles (sm_100a).
//
// Why: with per-pixel flow the four bilinear taps of neighbouring pixels land in
// unrelated cache lines; the direct kernel (warp.cu) then pays one L1 tag lookup
// per lane and tap (26 sectors per request, l1tex at 84 % while DRAM idles at
// 18 %: profiles/r01_ncu_dense_step_v1.txt).  Here every CTA stages the source
// window its output tile can touch in shared memory with one TMA box copy per
// channel plane (cp.async.bulk.tensor.3d -> UTMALDG, completion on an mbarrier),
// three planes in flight, and the taps become shared-memory loads.
//
// Work item = (128x16 output tile, side).  The forward and the backward chain of a step are independent, so
// 2 * tiles items are spread round-robin over a persistent grid of two 512-thread CTAs per SM (2040 items at 1080p).
// The window is a fixed 192x48 box around the tile (halo 32 px in x, 16 px in y), one channel plane at a time through
// three buffers; a warp whose taps leave the box (large motion) gathers from global memory instead, so any flow
// field stays correct.  Since dense_strip.cu (all channels resident, 1.5x instead of 4.5x fabric amplification) this
// kernel is the fall-back for class counts whose window does not fit shared memory (C > 6).
// Arithmetic is gs_setup/tap_acc from fuvs_common.cuh: bit-identical to the
// direct kernel and to ATen's grid_sampler_2d.
#include <cuda.h>

#include "dense_common.cuh"

namespace fuvs {

namespace {

constexpr int TW = 128, TH = 16;            // output tile
constexpr int TROWS = 4;                    // thread rows per CTA
constexpr int PX = TH / TROWS;              // pixels per thread (one column, stride TROWS rows)
constexpr int HALO_X = 32, HALO_Y = 16;
constexpr int BOXW = TW + 2 * HALO_X;       // 192
constexpr int BOXH = TH + 2 * HALO_Y;       // 48
constexpr int NBUF = 3;
constexpr int BOX_BYTES = BOXW * BOXH * 4;  // 36864
constexpr int THREADS = TW * TROWS;         // 512
constexpr size_t SMEM_BYTES = static_cast<size_t>(NBUF) * BOX_BYTES + 128;   // + 2*NBUF mbarriers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken tensor map must surface as a launch error, not as a hung GPU.  try_wait itself suspends the
// warp for a hardware-defined interval, so the spin count only bounds the pathological case.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct TileGeom {
  int tiles_x, tiles_y, n_items;
};

// Per-thread state of one work item: PX pixels in one column (rows ty, ty+TROWS, ...).
struct ItemTaps {
  int loc[PX];                               // float offset of the nw tap inside the staged box
  float wnw[PX], wne[PX], wsw[PX], wse[PX];
  unsigned dxm, dym, valid;                  // per-pixel bits: east / south neighbour in bounds, pixel inside the image
};

// One channel plane of one item.  Template flags are uniform per launch/side, so the pixel loop carries no
// run-time mode tests (the first version did, and was instruction-bound: profiles/r01_ncu_dense_tma_v1.txt).
//   EMIT : this side completes a frame: blend with the other chain's state read pointwise, arg-max, logits
//   KEY0 : step 1, forward side: frame 0 = arg-max of the key frame, read from the staged box itself
//   FULL : every pixel of the thread is inside the image and has both neighbours (no predication needed)
template <class NM, bool EMIT, bool KEY0, bool FULL>
__device__ __forceinline__ void plane_from_smem(const float* __restrict__ plane_s, const ItemTaps& T, int c,
                                                long long plane_off, long long pix0, int W, int key_loc,
                                                float* __restrict__ dst, const float* __restrict__ point,
                                                float w_this, float w_point, bool side,
                                                float* __restrict__ logit_out, float* __restrict__ logit0,
                                                ArgMax (&am)[PX]) {
  static_assert(!(EMIT && KEY0), "a side either completes a frame or carries frame 0 (the host routes n<=2 elsewhere)");
  const long long rstride = static_cast<long long>(TROWS) * W;
  float other[PX];
  if (EMIT) {
    const float* pp = point + plane_off + pix0;
#pragma unroll
    for (int r = 0; r < PX; ++r) other[r] = (FULL || ((T.valid >> r) & 1u)) ? __ldg(pp + r * rstride) : 0.f;
  }
  float* dp = dst ? dst + plane_off + pix0 : nullptr;
  float* lp = (EMIT && logit_out) ? logit_out + plane_off + pix0 : nullptr;
  float* l0p = (KEY0 && logit0) ? logit0 + plane_off + pix0 : nullptr;
  if (!EMIT && FULL) {
    // No pointwise operand to wait for: issue all PX*4 shared-memory loads before the first FMA, so 32 independent
    // LDS are in flight per thread (4 warps per scheduler cannot hide the conflict-stretched LDS latency otherwise).
    float v00[PX], v01[PX], v10[PX], v11[PX];
#pragma unroll
    for (int r = 0; r < PX; ++r) {
      const float* p = plane_s + T.loc[r];
      v00[r] = p[0]; v01[r] = p[1]; v10[r] = p[BOXW]; v11[r] = p[BOXW + 1];
    }
    float keyv[PX];
    if (KEY0) {
#pragma unroll
      for (int r = 0; r < PX; ++r) keyv[r] = plane_s[key_loc + r * (TROWS * BOXW)];
    }
#pragma unroll
    for (int r = 0; r < PX; ++r) {
      float acc = 0.f;
      acc = tap_acc<NM>(acc, v00[r], T.wnw[r]);
      acc = tap_acc<NM>(acc, v01[r], T.wne[r]);
      acc = tap_acc<NM>(acc, v10[r], T.wsw[r]);
      acc = tap_acc<NM>(acc, v11[r], T.wse[r]);
      if (dp) dp[r * rstride] = acc;
      if (KEY0) {
        am[r].push(keyv[r], c);
        if (l0p) __stcs(l0p + r * rstride, keyv[r]);
      }
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < PX; ++r) {
    const float* p = plane_s + T.loc[r];
    float acc = 0.f;
    if (FULL) {
      acc = tap_acc<NM>(acc, p[0], T.wnw[r]);
      acc = tap_acc<NM>(acc, p[1], T.wne[r]);
      acc = tap_acc<NM>(acc, p[BOXW], T.wsw[r]);
      acc = tap_acc<NM>(acc, p[BOXW + 1], T.wse[r]);
    } else {
      const int dx = (T.dxm >> r) & 1u, dy = (T.dym >> r) & 1u;
      const float v00 = p[0], v01 = p[dx], v10 = p[dy * BOXW], v11 = p[dy * BOXW + dx];
      acc = tap_acc<NM>(acc, v00, T.wnw[r]);
      if (dx) acc = tap_acc<NM>(acc, v01, T.wne[r]);
      if (dy) acc = tap_acc<NM>(acc, v10, T.wsw[r]);
      if (dx & dy) acc = tap_acc<NM>(acc, v11, T.wse[r]);
    }
    const bool live = FULL || ((T.valid >> r) & 1u);
    if (dp && live) dp[r * rstride] = acc;
    if (EMIT) {
      const float v = side ? blend2(w_point, other[r], w_this, acc) : blend2(w_this, acc, w_point, other[r]);
      am[r].push(v, c);
      if (lp && live) __stcs(lp + r * rstride, v);
    }
    if (KEY0) {
      const float v = plane_s[key_loc + r * (TROWS * BOXW)];
      am[r].push(v, c);
      if (l0p && live) __stcs(l0p + r * rstride, v);
    }
  }
}

// Rare path: a warp whose taps leave the staged box (large motion) computes its pixels of the whole item straight
// from global memory, exactly like the direct kernel.  Deliberately NOT inlined and fed with scalars only, so that
// its registers do not burden the shared-memory fast path.
template <class NM>
__device__ __forceinline__ void item_from_global(const DenseStep* __restrict__ Ap, bool side, int C, int H, int W, int x,
                                              int ytop, unsigned valid) {
  const DenseStep& A = *Ap;
  const long long HW = static_cast<long long>(H) * W;
  const float* grid = side ? A.gridR : A.gridL;
  const float* src = side ? A.srcR : A.srcL;
  float* dst = side ? A.dstR : A.dstL;
  const bool emit = side ? (A.emitB != 0) : (A.emitA != 0);
  const float* point = side ? A.pointL : A.pointR;
  const float w_this = side ? A.wB1 : A.wA0, w_point = side ? A.wB0 : A.wA1;
  uint8_t* lab_out = side ? A.labelB : A.labelA;
  float* logit_out = side ? A.logitB : A.logitA;
  const bool key0 = !side && (A.key0 != nullptr);
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
  }
}

constexpr int GRID_BOXW = 2 * TW;                     // flow grid tile as floats: (x,y) pairs of TW pixels
constexpr int GRID_BYTES = GRID_BOXW * TH * 4;        // 16384
constexpr int NWARPS = THREADS / 32;

struct TmaMaps {
  CUtensorMap srcL, srcR;     // [C][H][W] fp32, box BOXW x BOXH x 1
  CUtensorMap gridL, gridR;   // [H][2W]   fp32, box 2*TW x TH
};

// Pipeline.  Every work item contributes C+1 "planes" to its CTA's sequence: the flow-grid tile first, then the C
// channel windows.  NBUF shared-memory buffers cycle through the sequence; full[b] (mbarrier) completes when the TMA
// bytes of the plane in buffer b have landed.  A warp that has finished reading buffer b bumps done[b]; the warp
// that arrives LAST resets the counter, re-arms full[b] and issues the TMA for the plane NBUF positions ahead.
// Nobody ever waits for a buffer to drain (a designated producer thread that did became a convoy: every warp
// ended up synchronised to it once per plane), and warps never meet at a block-wide barrier.
// EMIT / KEY0 are launch-uniform (the host picks the instantiation), so each variant gets its own register budget.
template <class NM, int CT, bool EMIT, bool KEY0>
__global__ void __launch_bounds__(THREADS, 2)
dense_step_tma_kernel(const __grid_constant__ TmaMaps M, const __grid_constant__ DenseStep A, int Crt, int H, int W,
                      TileGeom G) {
  // staging buffers (TMA destinations need 128-byte alignment), barriers behind them.  No integer round-trip on
  // the pointer: the compiler must keep seeing shared-memory addresses to emit LDS instead of generic loads.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int C = CT > 0 ? CT : Crt;
  const int PPI = C + 1;                                // planes per item
  float* bufs = reinterpret_cast<float*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + static_cast<size_t>(NBUF) * BOX_BYTES);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int tx = tid & (TW - 1), ty = tid >> 7;           // TW == 128
  const long long HW = static_cast<long long>(H) * W;
  // bars[0..NBUF) = full mbarriers; done[0..NBUF) = warps finished with the buffer
  unsigned* done = reinterpret_cast<unsigned*>(bars + NBUF);

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(smem_u32(&bars[b]), 1);
      done[b] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Window origin of a tile, kept inside the image where the image is large enough: boxes hanging over the image
  // edge are zero-filled by TMA but load far slower (measured in dense_strip.cu), and border clipping never reads there.
  auto box_x = [&](int txi) { return max(0, min(txi * TW - HALO_X, W - BOXW)); };
  auto box_y = [&](int tyi) { return tyi * TH - HALO_Y; };   // rows above / below the image: TMA zero fill, never read
  // one thread: load plane q of this CTA's sequence into buffer q % NBUF (the buffer is known to be drained)
  auto issue = [&](int q) {
    const int k = q / PPI, j = q - k * PPI;
    const int it = blockIdx.x + k * gridDim.x;
    if (it >= G.n_items) return;
    const int b = q % NBUF;
    const bool side = (it & 1) != 0;
    const int tile = it >> 1;
    const int tyi = tile / G.tiles_x, txi = tile - tyi * G.tiles_x;
    const uint32_t bar = smem_u32(&bars[b]);
    const uint32_t dst = smem_u32(bufs + static_cast<size_t>(b) * BOXW * BOXH);
    if (j == 0) {
      mbar_expect_tx(bar, GRID_BYTES);
      tma_load_2d(dst, side ? &M.gridR : &M.gridL, 2 * txi * TW, tyi * TH, bar);
    } else {
      mbar_expect_tx(bar, BOX_BYTES);
      tma_load_3d(dst, side ? &M.srcR : &M.srcL, box_x(txi), box_y(tyi), j - 1, bar);
    }
  };
  // all threads: this warp is done reading the buffer of plane q; the last warp to say so refills it
  auto release = [&](int q) {
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int b = q % NBUF;
      if (atomicAdd(&done[b], 1u) == NWARPS - 1) {
        done[b] = 0u;               // published to the other warps by the release-arrive on full[b] inside issue()
        issue(q + NBUF);
      }
    }
  };
  if (tid == 0) {
    for (int q = 0; q < NBUF; ++q) issue(q);
  }

  int q = 0;
  for (int k = 0;; ++k) {
    const int it = blockIdx.x + k * gridDim.x;
    if (it >= G.n_items) break;
    const bool side = (it & 1) != 0;
    const int tile = it >> 1;
    const int tyi = tile / G.tiles_x, txi = tile - tyi * G.tiles_x;
    const int x0 = txi * TW, y0 = tyi * TH;
    const int xbase = box_x(txi), ybase = box_y(tyi);
    const int x = x0 + tx;
    float* dst = side ? A.dstR : A.dstL;
    // the frame this side completes: side L -> frame A (other operand R_{n-j} from memory),
    //                                side R -> frame B (other operand L_{n-j} from memory)
    constexpr bool emit = EMIT;
    const float* point = side ? A.pointL : A.pointR;
    const float w_this = side ? A.wB1 : A.wA0;     // weight of the state computed here
    const float w_point = side ? A.wB0 : A.wA1;    // weight of the state read pointwise
    uint8_t* lab_out = side ? A.labelB : A.labelA;
    float* logit_out = side ? A.logitB : A.logitA;
    const bool do_key0 = KEY0 && !side;

    // ---- plane 0 of the item: the flow-grid tile -> taps of this thread's PX pixels
    ItemTaps T;
    T.dxm = 0u; T.dym = 0u; T.valid = 0u;
    bool outside = false;
    {
      const int b = q % NBUF;
      mbar_wait(smem_u32(&bars[b]), static_cast<uint32_t>((q / NBUF) & 1));
      const float* gs = bufs + static_cast<size_t>(b) * BOXW * BOXH;
#pragma unroll
      for (int r = 0; r < PX; ++r) {
        const int y = y0 + ty + r * TROWS;
        T.loc[r] = 0;
        T.wnw[r] = T.wne[r] = T.wsw[r] = T.wse[r] = 0.f;
        if (x < W && y < H) {
          const float2 g = *reinterpret_cast<const float2*>(gs + (ty + r * TROWS) * GRID_BOXW + 2 * tx);
          const GsTap t = gs_setup<NM>(g.x, g.y, H, W, false);
          T.wnw[r] = t.nw; T.wne[r] = t.ne; T.wsw[r] = t.sw; T.wse[r] = t.se;
          const int lx = t.ix - xbase, ly = t.iy - ybase;
          const bool in_box = (lx >= 0) && (lx + t.dx <= BOXW - 1) && (ly >= 0) && (ly + t.dy <= BOXH - 1);
          T.loc[r] = in_box ? ly * BOXW + lx : 0;
          outside |= !in_box;
          T.valid |= 1u << r;
          if (t.dx) T.dxm |= 1u << r;
          if (t.dy) T.dym |= 1u << r;
        }
      }
      release(q);
      ++q;
    }
    const unsigned all = (1u << PX) - 1u;
    const bool full = (T.valid == all) && (T.dxm == all) && (T.dym == all);
    // warp-uniform decisions: some tap of this warp leaves the staged box -> the warp gathers from global memory;
    // some lane needs predication (image border) -> predicated shared-memory path
    const bool use_global = __any_sync(0xffffffffu, outside);
    const bool all_full = __all_sync(0xffffffffu, full);

    if (use_global) {
      // whole warp leaves the fast path for this item; it still follows the buffer protocol plane by plane
      item_from_global<NM>(&A, side, C, H, W, x, y0 + ty, T.valid);
      for (int c = 0; c < C; ++c, ++q) {
        mbar_wait(smem_u32(&bars[q % NBUF]), static_cast<uint32_t>((q / NBUF) & 1));
        release(q);
      }
      continue;
    }

    ArgMax am[PX];
#pragma unroll
    for (int r = 0; r < PX; ++r) am[r].init(-INFINITY);   // push(v, 0) then always selects class 0 first (also for -inf / NaN)
    const long long pix0 = static_cast<long long>(y0 + ty) * W + x;
    const int key_loc = (HALO_Y + ty) * BOXW + (min(x, W - 1) - xbase);

    for (int c = 0; c < C; ++c, ++q) {
      const int b = q % NBUF;
      mbar_wait(smem_u32(&bars[b]), static_cast<uint32_t>((q / NBUF) & 1));
      const float* plane_s = bufs + static_cast<size_t>(b) * BOXW * BOXH;
      const long long plane_off = c * HW;
#define FUVS_PLANE(EMIT_, KEY0_)                                                                                        \
  do {                                                                                                                  \
    if (all_full)                                                                                                       \
      plane_from_smem<NM, EMIT_, KEY0_, true>(plane_s, T, c, plane_off, pix0, W, key_loc, dst, point, w_this, w_point,  \
                                              side, logit_out, A.logit0, am);                                           \
    else                                                                                                                \
      plane_from_smem<NM, EMIT_, KEY0_, false>(plane_s, T, c, plane_off, pix0, W, key_loc, dst, point, w_this, w_point, \
                                               side, logit_out, A.logit0, am);                                          \
  } while (0)
      if (EMIT) FUVS_PLANE(true, false);
      else if (KEY0 && do_key0) FUVS_PLANE(false, true);
      else FUVS_PLANE(false, false);
#undef FUVS_PLANE
      release(q);
    }
#pragma unroll
    for (int r = 0; r < PX; ++r) {
      if (!((T.valid >> r) & 1u)) continue;
      const long long pix = pix0 + static_cast<long long>(r) * TROWS * W;
      if (emit && lab_out) lab_out[pix] = static_cast<uint8_t>(am[r].idx);
      if (do_key0 && A.label0) A.label0[pix] = static_cast<uint8_t>(am[r].idx);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool make_map(CUtensorMap* m, const float* ptr, int C, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(C)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(W) * 4ull, static_cast<cuuint64_t>(W) * H * 4ull};
  const cuuint32_t box[3] = {BOXW, BOXH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// flow grid [H][W][2] viewed as a 2-D fp32 tensor [H][2W]
bool make_grid_map(CUtensorMap* m, const float* ptr, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(W) * 2ull, static_cast<cuuint64_t>(H)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(W) * 8ull};
  const cuuint32_t box[2] = {GRID_BOXW, TH};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CT, bool EMIT, bool KEY0>
int launch_variant(const TmaMaps& maps, const DenseStep& a, int C, int H, int W, cudaStream_t st) {
  static SmemOptIn optin;
  auto kern = dense_step_tma_kernel<Nm, CT, EMIT, KEY0>;
  if (!optin.ensure(kern, static_cast<int>(SMEM_BYTES))) return 1;
  TileGeom g;
  g.tiles_x = (W + TW - 1) / TW;
  g.tiles_y = (H + TH - 1) / TH;
  g.n_items = 2 * g.tiles_x * g.tiles_y;
  int grid = 2 * sm_count();                 // two CTAs per SM
  if (grid > g.n_items) grid = g.n_items;
  kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, a, C, H, W, g);
  return check_launch("fuvs_dense_interval(tma step)");
}

template <int CT>
int launch_ct(const TmaMaps& maps, const DenseStep& a, int C, int H, int W, cudaStream_t st) {
  if (a.emitA) return launch_variant<CT, true, false>(maps, a, C, H, W, st);
  if (a.key0) return launch_variant<CT, false, true>(maps, a, C, H, W, st);
  return launch_variant<CT, false, false>(maps, a, C, H, W, st);
}

}  // namespace

int launch_dense_step_tma(const DenseStep& a, int C, int H, int W, cudaStream_t st) {
  // eligibility: TMA needs 16-byte aligned bases and row pitch; the even-n middle step (both operands fresh)
  // is emitted by the caller with a pointwise kernel instead
  if ((W & 3) != 0 || W < 4 || !aligned16(a.srcL) || !aligned16(a.srcR) || H >= 32768 || W >= 32768) return 1;
  if ((a.emitA && !a.pointR) || (a.emitB && !a.pointL) || (a.emitA != a.emitB)) return 1;
  if (a.key0 && (a.key0 != a.srcL || a.emitA)) return 1;   // frame 0 and a completed frame never share a side here
  if (!aligned16(a.gridL) || !aligned16(a.gridR)) return 1;
  TmaMaps maps;
  if (!make_map(&maps.srcL, a.srcL, C, H, W) || !make_map(&maps.srcR, a.srcR, C, H, W) ||
      !make_grid_map(&maps.gridL, a.gridL, H, W) || !make_grid_map(&maps.gridR, a.gridR, H, W))
    return 1;
  switch (C) {
    case 2: return launch_ct<2>(maps, a, C, H, W, st);
    case 5: return launch_ct<5>(maps, a, C, H, W, st);
    default: return launch_ct<0>(maps, a, C, H, W, st);
  }
}

}  // namespace fuvs
