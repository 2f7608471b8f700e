// fuvs_common.cuh — shared device helpers of libfuvs (sm_100a).
//
// Numerics contract.  The reference computes this path with eager ATen CUDA
// kernels: grid_sampler_2d, upsample_bilinear2d, mul(Tensor,
// Scalar), add, max(dim=1).  Label maps flip on 1-ulp logit differences, so
// every fp32 rounding step of those kernels is spelled out here with explicit
// intrinsics and the translation units are compiled with -fmad=false: the
// only fused multiply-adds are the ones written as __fmaf_rn.  Where ATen's
// source leaves the contraction to nvcc (a*b+c*d ...), the choice nvcc makes
// is a template policy (NumericsT) so that tests/test_calibration_gpu.py can
// instantiate every candidate and prove on the GPU which one torch's binary
// uses.  Formula sources: torch ATen/native/cuda/GridSampler.cuh:23-31,55-57
// and ATen/native/cuda/UpSample.cuh:96-130 (headers shipped with torch).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fuvs.h"

namespace fuvs {

// ---------------------------------------------------------------------------
// host-side plumbing (abi.cu)
// ---------------------------------------------------------------------------
int  set_error(int code, const char* fmt, ...);
int  check_launch(const char* what);           // cudaPeekAtLastError -> code
void count_launch();
int  device_ok();                              // 0 / FUVS_ENODEV
int  sm_count();
template <typename K>
int blocks_per_sm(K kernel, int threads, size_t smem = 0) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) {
    cudaGetLastError();
    nb = 1;
  }
  return nb;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember per device whether a kernel has it.
struct SmemOptIn {
  bool done[64] = {};
  template <typename K>
  bool ensure(K kernel, int bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (done[dev]) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    done[dev] = true;
    return true;
  }
};

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }
static inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

#define FUVS_MAX_FRAMES 60  /* 4 px x n frames must fit an 8-bit packed counter */

// Blend weights exactly as ATen's mul(Tensor, Scalar) sees them: the python
// double (n-p)/n rounded once to fp32 (flow/model.py:234-235).
struct BlendWeights {
  float w0[FUVS_MAX_FRAMES + 4];
  float w1[FUVS_MAX_FRAMES + 4];
};
static inline void make_blend_weights(int n, BlendWeights* w) {
  for (int p = 0; p < n && p < FUVS_MAX_FRAMES + 4; ++p) {
    w->w0[p] = static_cast<float>(static_cast<double>(n - p) / static_cast<double>(n));
    w->w1[p] = static_cast<float>(static_cast<double>(p) / static_cast<double>(n));
  }
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------
// numerics policy
// ---------------------------------------------------------------------------
//  UNNORM_FMA    grid_sampler_unnormalize, align_corners=False:
//                1: fma(x+1, size, -1)/2      0: ((x+1)*size - 1)/2
//  TAP_FMA       out_acc += v*w               1: fma chain from 0   0: mul, add
//  UP_LAMBDA_FMA h1lambda = scale*i - floor   1: fma(scale,i,-h1)   0: mul, sub
//  UP_INNER      w0*a + w1*b                  0: fma(w0,a,w1*b)  1: fma(w1,b,w0*a)  2: unfused
//  UP_OUTER      h0*r0 + h1*r1                same encoding
template <int UNNORM_FMA, int TAP_FMA, int UP_LAMBDA_FMA, int UP_INNER, int UP_OUTER>
struct NumericsT {
  static constexpr int kUnnormFma = UNNORM_FMA;
  static constexpr int kTapFma = TAP_FMA;
  static constexpr int kUpLambdaFma = UP_LAMBDA_FMA;
  static constexpr int kUpInner = UP_INNER;
  static constexpr int kUpOuter = UP_OUTER;
};
// What nvcc 12.x emits for ATen's sources (checked in SASS, and against
// torch 2.11+cu128 on a B200 by the calibration test).
using Nm = NumericsT<1, 1, 0, 0, 0>;

// ---------------------------------------------------------------------------
// cache-hinted memory access (streaming data is touched once)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_stream1(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_stream1(float* p, float v) { __stcs(p, v); }

template <int VEC> struct FVec;
template <> struct FVec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void load_stream(const float* p) { v[0] = __ldcs(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
  __device__ __forceinline__ void store_stream(float* p) const { __stcs(p, v[0]); }
};
template <> struct FVec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void load_stream(const float* p) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ void store_stream(float* p) const {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};

template <int VEC> struct LVec;   // VEC uint8 labels
template <> struct LVec<1> {
  int v[1];
  __device__ __forceinline__ void load(const uint8_t* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void store(uint8_t* p) const { *p = static_cast<uint8_t>(v[0]); }
};
template <> struct LVec<4> {
  int v[4];
  __device__ __forceinline__ void load(const uint8_t* p) {
    unsigned t = __ldg(reinterpret_cast<const unsigned*>(p));
    v[0] = t & 255u; v[1] = (t >> 8) & 255u; v[2] = (t >> 16) & 255u; v[3] = t >> 24;
  }
  __device__ __forceinline__ void store(uint8_t* p) const {
    unsigned t = (unsigned)v[0] | ((unsigned)v[1] << 8) | ((unsigned)v[2] << 16) | ((unsigned)v[3] << 24);
    *reinterpret_cast<unsigned*>(p) = t;
  }
};

// ---------------------------------------------------------------------------
// blend: fl(fl(wa*a) + fl(wb*b))  — three separately rounded ATen launches
// (flow/model.py:234-236; SURVEY.md §2c: an FMA here mismatches ~25 % of bits)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float blend2(float wa, float a, float wb, float b) {
  return __fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb));
}

// ---------------------------------------------------------------------------
// packed FP32x2 arithmetic (Blackwell FMUL2 / FFMA2): two IEEE-rn fp32 results
// per instruction, which halves the issue slots of the blend.
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though the
// .rn forms must not fuse (checked in SASS), which would break bit-exactness.
// The add is therefore written as fma(p, one, q) with `one` a RUN-TIME 1.0f
// (a kernel argument): p*1 is exact, so the result is rn(p+q), and ptxas
// cannot fold a multiplier it does not know.
// ---------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 mul2_rn(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2_rn(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// {fl(fl(a.lo*wa)+fl(b.lo*wb)), fl(fl(a.hi*wa)+fl(b.hi*wb))}; wa2/wb2/one2 hold the scalar in both halves
__device__ __forceinline__ u64 blend2x2(u64 wa2, u64 a, u64 wb2, u64 b, u64 one2) {
  return fma2_rn(mul2_rn(a, wa2), one2, mul2_rn(b, wb2));
}

// ---------------------------------------------------------------------------
// arg-max over classes with torch.max(dim) semantics (flow/base.py:147,167,276):
// lowest index wins ties, a NaN beats every number, the first NaN stays.
// ---------------------------------------------------------------------------
struct ArgMax {
  float best;
  int idx;
  __device__ __forceinline__ void init(float v) { best = v; idx = 0; }
  __device__ __forceinline__ void push(float v, int c) {
    // v > best, or v is NaN and best is not: !(v <= best) is true for both, a NaN best is never replaced
    const bool take = !(v <= best) && (best == best);
    if (take) { best = v; idx = c; }
  }
};

// ---------------------------------------------------------------------------
// grid_sample(bilinear, padding_mode="border") source coordinates and taps
// ---------------------------------------------------------------------------
struct GsTap {
  int off00;        // iy_nw * Win + ix_nw (always in bounds in border mode)
  int ix, iy;       // the north-west corner itself
  int dx, dy;       // 1 if the east / south neighbour is in bounds, else 0
  float nw, ne, sw, se;
};

template <class NM>
__device__ __forceinline__ float gs_source_index(float coord, int size, bool align_corners) {
  const float t = __fadd_rn(coord, 1.f);
  float r;
  if (align_corners) {
    r = __fmul_rn(__fmul_rn(t, 0.5f), static_cast<float>(size - 1));
  } else {
    const float u = NM::kUnnormFma ? __fmaf_rn(t, static_cast<float>(size), -1.f)
                                   : __fsub_rn(__fmul_rn(t, static_cast<float>(size)), 1.f);
    r = __fmul_rn(u, 0.5f);
  }
  // clip_coordinates: min(size-1, max(r, 0)); fmaxf(NaN, 0) == 0 like ::max
  return fminf(static_cast<float>(size - 1), fmaxf(r, 0.f));
}

template <class NM>
__device__ __forceinline__ GsTap gs_setup(float gx, float gy, int Hin, int Win, bool align_corners) {
  const float ix = gs_source_index<NM>(gx, Win, align_corners);
  const float iy = gs_source_index<NM>(gy, Hin, align_corners);
  // floor() and the float->int conversion without the XU pipe (FRND / F2I sit on the critical path of every pixel):
  // after clip_coordinates 0 <= ix <= size-1 < 2^22, so t = ix + 2^23 rounded DOWN is exactly 2^23 + floor(ix)
  // (ulp(t) == 1), t - 2^23 is exact, and the low mantissa bits of t are the integer itself.
  const float tx = __fadd_rd(ix, 8388608.f), ty = __fadd_rd(iy, 8388608.f);
  const float fx = __fsub_rn(tx, 8388608.f), fy = __fsub_rn(ty, 8388608.f);
  const int ix_nw = __float_as_int(tx) - 0x4B000000, iy_nw = __float_as_int(ty) - 0x4B000000;
  // ATen converts the integer corners back to float before subtracting: float(ix_nw) == fx and
  // float(ix_nw + 1) == fx + 1 exactly (integers far below 2^24), which saves four I2F conversions per pixel
  const float x_w = fx, x_e = __fadd_rn(fx, 1.f);
  const float y_n = fy, y_s = __fadd_rn(fy, 1.f);
  GsTap t;
  t.nw = __fmul_rn(__fsub_rn(x_e, ix), __fsub_rn(y_s, iy));
  t.ne = __fmul_rn(__fsub_rn(ix, x_w), __fsub_rn(y_s, iy));
  t.sw = __fmul_rn(__fsub_rn(x_e, ix), __fsub_rn(iy, y_n));
  t.se = __fmul_rn(__fsub_rn(ix, x_w), __fsub_rn(iy, y_n));
  t.dx = (ix_nw + 1 < Win) ? 1 : 0;
  t.dy = (iy_nw + 1 < Hin) ? 1 : 0;
  t.off00 = iy_nw * Win + ix_nw;
  t.ix = ix_nw;
  t.iy = iy_nw;
  return t;
}

template <class NM>
__device__ __forceinline__ float tap_acc(float acc, float v, float w) {
  return NM::kTapFma ? __fmaf_rn(v, w, acc) : __fadd_rn(acc, __fmul_rn(v, w));
}

// One channel plane; out-of-bounds neighbours are skipped like ATen's
// within_bounds_2d (their weight is 0 in border mode but the value may be
// Inf/NaN, so it must not enter the sum).
template <class NM>
__device__ __forceinline__ float gs_fetch(const float* __restrict__ plane, const GsTap& t, int Win) {
  const float* p = plane + t.off00;
  const float v00 = __ldg(p);
  const float v01 = __ldg(p + t.dx);
  const float v10 = __ldg(p + t.dy * Win);
  const float v11 = __ldg(p + t.dy * Win + t.dx);
  float acc = 0.f;
  acc = tap_acc<NM>(acc, v00, t.nw);
  if (t.dx) acc = tap_acc<NM>(acc, v01, t.ne);
  if (t.dy) acc = tap_acc<NM>(acc, v10, t.sw);
  if (t.dx & t.dy) acc = tap_acc<NM>(acc, v11, t.se);
  return acc;
}

// ---------------------------------------------------------------------------
// upsample_bilinear2d(align_corners=True) source coordinates and formula
// ---------------------------------------------------------------------------
struct UpCoord {
  int i0;       // floor source index
  int ip;       // 1 if i0 + 1 is in range
  float l0, l1; // lambdas
};

template <class NM>
__device__ __forceinline__ UpCoord up_coord(float scale, int dst, int in_size) {
  const float fd = static_cast<float>(dst);
  const float r = __fmul_rn(scale, fd);
  UpCoord u;
  u.i0 = static_cast<int>(r);
  u.ip = (u.i0 < in_size - 1) ? 1 : 0;
  u.l1 = NM::kUpLambdaFma ? __fmaf_rn(scale, fd, -static_cast<float>(u.i0))
                          : __fsub_rn(r, static_cast<float>(u.i0));
  u.l0 = __fsub_rn(1.f, u.l1);
  return u;
}

template <int MODE>
__device__ __forceinline__ float two_term(float wa, float a, float wb, float b) {
  if (MODE == 0) return __fmaf_rn(wa, a, __fmul_rn(wb, b));
  if (MODE == 1) return __fmaf_rn(wb, b, __fmul_rn(wa, a));
  return __fadd_rn(__fmul_rn(wa, a), __fmul_rn(wb, b));
}

// packed FP32x2 form: two IEEE-rn results per instruction, same roundings as two_term<MODE> on each half
template <int MODE>
__device__ __forceinline__ u64 two_term2(u64 wa2, u64 a2, u64 wb2, u64 b2, u64 one2) {
  if (MODE == 0) return fma2_rn(wa2, a2, mul2_rn(wb2, b2));
  if (MODE == 1) return fma2_rn(wb2, b2, mul2_rn(wa2, a2));
  return fma2_rn(mul2_rn(wa2, a2), one2, mul2_rn(wb2, b2));     // unfused: see blend2x2 for the run-time one
}

template <class NM>
__device__ __forceinline__ float up_value(const UpCoord& h, const UpCoord& w,
                                          float v00, float v01, float v10, float v11) {
  const float r0 = two_term<NM::kUpInner>(w.l0, v00, w.l1, v01);
  const float r1 = two_term<NM::kUpInner>(w.l0, v10, w.l1, v11);
  return two_term<NM::kUpOuter>(h.l0, r0, h.l1, r1);
}

template <class NM>
__device__ __forceinline__ float up_fetch(const float* __restrict__ plane, int Win,
                                          const UpCoord& h, const UpCoord& w) {
  const float* p = plane + h.i0 * Win + w.i0;
  return up_value<NM>(h, w, __ldg(p), __ldg(p + w.ip), __ldg(p + h.ip * Win), __ldg(p + h.ip * Win + w.ip));
}

// ---------------------------------------------------------------------------
// (I, O, T) class histograms of util/util.py:52-63.
// Fast path K <= 8: 8-bit fields packed in 64-bit registers per thread, widened
// to 16-bit pairs and summed across the warp with REDUX, kept as per-warp
// 32-bit totals in registers, merged once per block in shared memory and
// flushed with 3K 64-bit atomics per block.  Integer only => deterministic.
// ---------------------------------------------------------------------------
struct Hist3Packed {
  unsigned long long I, O, T;
  __device__ __forceinline__ void clear() { I = 0ull; O = 0ull; T = 0ull; }
  // temporal / histc convention: a value counts iff 0 <= v < K
  __device__ __forceinline__ void add(int o, int t, int ignore, int K) {
    if (t == ignore) o = ignore;
    const unsigned long long fo = (static_cast<unsigned>(o) < static_cast<unsigned>(K)) ? (1ull << (8 * (o & 7))) : 0ull;
    const unsigned long long ft = (static_cast<unsigned>(t) < static_cast<unsigned>(K)) ? (1ull << (8 * (t & 7))) : 0ull;
    O += fo;
    T += ft;
    I += (o == t) ? fo : 0ull;
  }
};

template <int K>
struct WarpTotals {
  unsigned I[K], O[K], T[K];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int c = 0; c < K; ++c) { I[c] = 0u; O[c] = 0u; T[c] = 0u; }
  }
};

template <int K>
__device__ __forceinline__ void warp_sum_packed(unsigned long long packed, unsigned* tot) {
  const unsigned lo = static_cast<unsigned>(packed), hi = static_cast<unsigned>(packed >> 32);
#pragma unroll
  for (int r = 0; r < (K + 1) / 2; ++r) {
    const unsigned src = (r < 2) ? lo : hi;
    const unsigned pair = __byte_perm(src, 0u, (r & 1) ? 0x4342u : 0x4140u);  // {b0|b1<<16} or {b2|b3<<16}
    const unsigned s = __reduce_add_sync(0xffffffffu, pair);
    tot[2 * r] += s & 0xffffu;
    if (2 * r + 1 < K) tot[2 * r + 1] += s >> 16;
  }
}

// All 32 lanes must call.  Resets the per-thread packed counters.
template <int K>
__device__ __forceinline__ void warp_accumulate(Hist3Packed& h, WarpTotals<K>& tot) {
  warp_sum_packed<K>(h.I, tot.I);
  warp_sum_packed<K>(h.O, tot.O);
  warp_sum_packed<K>(h.T, tot.T);
  h.clear();
}

// All threads of the block must call (uses __syncthreads).  sh: 3*8 unsigned.
template <int K>
__device__ __forceinline__ void block_flush_counts(const WarpTotals<K>& tot, unsigned* sh,
                                                   unsigned long long* counts, int Kruntime) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  if (tid < 24) sh[tid] = 0u;
  __syncthreads();
  if ((tid & 31) == 0) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      if (tot.I[c]) atomicAdd(&sh[c], tot.I[c]);
      if (tot.O[c]) atomicAdd(&sh[8 + c], tot.O[c]);
      if (tot.T[c]) atomicAdd(&sh[16 + c], tot.T[c]);
    }
  }
  __syncthreads();
  if (tid < K && tid < Kruntime) {
    const unsigned long long i = sh[tid], o = sh[8 + tid], t = sh[16 + tid];
    if (i) atomicAdd(counts + tid, i);
    if (o + t - i) atomicAdd(counts + Kruntime + tid, o + t - i);   // area_union = O + T - I
    if (t) atomicAdd(counts + 2 * Kruntime + tid, t);
  }
}

// ---------------------------------------------------------------------------
// Cheaper per-thread counters for label maps whose values are known to be < K
// (arg-max outputs): three 32-bit registers with FW-bit class fields, one shift
// and three adds per label; spill() widens into 32-bit per-thread totals before
// a field can overflow; finish() REDUX-reduces per warp once per kernel.
// Requires ignore_index outside [0,K) (checked on the host).
// ---------------------------------------------------------------------------
template <int K> struct FieldCfg {
  static constexpr int FW = (K <= 4) ? 8 : (K <= 5 ? 6 : 4);
  static constexpr unsigned MASK = (1u << FW) - 1u;
  static constexpr int CAP = static_cast<int>(MASK);     // labels a field can absorb between spills
};

template <int K>
struct FieldCounts {
  using FC = FieldCfg<K>;
  unsigned accI, accO, accT;
  unsigned totI[K], totO[K], totT[K];
  __device__ __forceinline__ void init() {
    accI = accO = accT = 0u;
#pragma unroll
    for (int c = 0; c < K; ++c) { totI[c] = 0u; totO[c] = 0u; totT[c] = 0u; }
  }
  __device__ __forceinline__ static unsigned field(int lab) { return 1u << (FC::FW * lab); }
  // current label `lab` (field fo, 0 if the pixel is ignored) against the previous frame's label `last` (field flast)
  __device__ __forceinline__ void add(int lab, unsigned fo, int last, unsigned flast) {
    accO += fo;
    accT += flast;
    accI += (lab == last) ? fo : 0u;
  }
  __device__ __forceinline__ void spill() {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      totI[c] += (accI >> (FC::FW * c)) & FC::MASK;
      totO[c] += (accO >> (FC::FW * c)) & FC::MASK;
      totT[c] += (accT >> (FC::FW * c)) & FC::MASK;
    }
    accI = accO = accT = 0u;
  }
  // all threads of the block must call; sh: 24 unsigned
  __device__ __forceinline__ void finish(unsigned* sh, unsigned long long* counts, int Kruntime) {
    spill();
    WarpTotals<K> wt;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      wt.I[c] = __reduce_add_sync(0xffffffffu, totI[c]);
      wt.O[c] = __reduce_add_sync(0xffffffffu, totO[c]);
      wt.T[c] = __reduce_add_sync(0xffffffffu, totT[c]);
    }
    block_flush_counts<K>(wt, sh, counts, Kruntime);
  }
};

// Generic K <= 256: block histogram in shared memory (3*256 unsigned).
__device__ __forceinline__ void smem_hist_add(unsigned* sh, int o, int t, int ignore, int K) {
  if (t == ignore) o = ignore;
  const bool vo = static_cast<unsigned>(o) < static_cast<unsigned>(K);
  const bool vt = static_cast<unsigned>(t) < static_cast<unsigned>(K);
  if (vo) atomicAdd(&sh[256 + o], 1u);
  if (vt) atomicAdd(&sh[512 + t], 1u);
  if (vo && o == t) atomicAdd(&sh[o], 1u);
}
__device__ __forceinline__ void smem_hist_flush(unsigned* sh, unsigned long long* counts, int K) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  const int nt = blockDim.x * blockDim.y;
  __syncthreads();
  for (int c = tid; c < K; c += nt) {
    const unsigned long long i = sh[c], o = sh[256 + c], t = sh[512 + c];
    if (i) atomicAdd(counts + c, i);
    if (o + t - i) atomicAdd(counts + K + c, o + t - i);
    if (t) atomicAdd(counts + 2 * K + c, t);
  }
}
__device__ __forceinline__ void smem_hist_clear(unsigned* sh) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  const int nt = blockDim.x * blockDim.y;
  for (int c = tid; c < 768; c += nt) sh[c] = 0u;
  __syncthreads();
}

#endif  // __CUDACC__

}  // namespace fuvs
