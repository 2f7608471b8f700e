// pix4.cuh — per-thread pixel-group helpers shared by the streaming kernels (linear.cu, block_rows.cu): arg-max
// over classes for a group of pixels and 128-bit / 64-bit pixel-group I/O.
#pragma once
#include "fuvs_common.cuh"

namespace fuvs {

// arg-max of NPX pixels over CT classes (torch.max semantics: lowest index wins ties).  NANSAFE=false is only used
// when no value can be NaN.  The scan is 1 FSETP + 1 FSEL + 1 SEL per class and pixel, all on the half-rate ALU pipe
// — the bound of the linear kernels; PTX predicated moves (which could issue on the FMA pipe) are folded back into
// SEL by ptxas 12.9, so the plain form is kept.
template <int CT, int NPX, bool NANSAFE>
__device__ __forceinline__ void argmaxN(const float (&x)[CT][NPX], int (&lab)[NPX]) {
#pragma unroll
  for (int i = 0; i < NPX; ++i) {
    float best = x[0][i];
    int idx = 0;
#pragma unroll
    for (int c = 1; c < CT; ++c) {
      const float v = x[c][i];
      if (NANSAFE) {
        const bool take = (v > best) || ((v != v) && (best == best));
        best = take ? v : best;
        idx = take ? c : idx;
      } else {
        const bool take = v > best;
        best = take ? v : best;
        idx = take ? c : idx;
      }
    }
    lab[i] = idx;
  }
}

// ---------------------------------------------------------------------------
// Arg-max in the float domain for values that cannot be NaN (the fast path of the streaming kernels).
// The compare-select scan above costs 3 instructions per class and pixel on the ALU pipe (FSETP, FSEL, SEL), 12 per
// pixel at C = 5, in one dependent chain — the bound of the linear kernel (profiles/r01_ncu_linear_final.txt).  Here
//   m   = max over classes                      2 FMNMX3 at C = 5 (Blackwell three-input min/max; +0 > -0)
//   n_c = (x_c != m) ? 1.0f : 0.0f              1 FSET per class, independent of each other
//   idx = n_0 (1 + n_1 (1 + n_2 (1 + ...)))     = number of leading classes that differ from the maximum
//                                               = lowest index that attains it: torch.max's tie rule, -0 == +0
// and the Horner chain t <- fma(n_c, t, n_c) runs on packed FP32x2 for a pixel pair (FMA pipe, exact small
// integers).  6 ALU + 1.5 FMA-pipe instructions per pixel instead of 12 ALU.  The index stays a float: it is
// compared as a float, turned into a counter field with one FFMA2 (FW * idx + 2^23: the low mantissa bits are the
// shift amount) and packed into label bytes the same way.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int CT>
__device__ __forceinline__ float max_classes(const float (&v)[CT]) {
  float m = v[0];
  int c = 1;
#pragma unroll
  for (; c + 1 < CT; c += 2) m = fmax3(m, v[c], v[c + 1]);
  if (c < CT) m = fmaxf(m, v[c]);
  return m;
}
__device__ __forceinline__ float differs(float a, float b) { return a != b ? 1.0f : 0.0f; }
// the same with NaN propagation: NaN if any class value is NaN (torch.max then returns the first NaN's index: exact scan)
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) {
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmax2_nan(float a, float b) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
template <int CT>
__device__ __forceinline__ float max_classes_nan(const float (&v)[CT]) {
  float m = v[0];
  int c = 1;
#pragma unroll
  for (; c + 1 < CT; c += 2) m = fmax3_nan(m, v[c], v[c + 1]);
  if (c < CT) m = fmax2_nan(m, v[c]);
  return m;
}

// {float index of pixel 0, float index of pixel 1} of the first maximum over CT packed class values (CT >= 2)
template <int CT>
__device__ __forceinline__ u64 argmax2f(const u64 (&x)[CT]) {
  float lo[CT], hi[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) unpack2(x[c], lo[c], hi[c]);
  const float mlo = max_classes<CT>(lo), mhi = max_classes<CT>(hi);
  u64 t = pack2(differs(lo[CT - 2], mlo), differs(hi[CT - 2], mhi));
#pragma unroll
  for (int c = CT - 3; c >= 0; --c) {
    const u64 nc = pack2(differs(lo[c], mlo), differs(hi[c], mhi));
    t = fma2_rn(nc, t, nc);
  }
  return t;
}
// The same for values that may be NaN (Inf is fine: it compares like any other value): *has_nan is set when either
// pixel holds a NaN class value — the caller then takes the exact scan — and the returned indices are meaningless.
template <int CT>
__device__ __forceinline__ u64 argmax2f_nan(const u64 (&x)[CT], bool* has_nan) {
  float lo[CT], hi[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) unpack2(x[c], lo[c], hi[c]);
  const float mlo = max_classes_nan<CT>(lo), mhi = max_classes_nan<CT>(hi);
  *has_nan = *has_nan || (mlo != mlo) || (mhi != mhi);
  u64 t = pack2(differs(lo[CT - 2], mlo), differs(hi[CT - 2], mhi));
#pragma unroll
  for (int c = CT - 3; c >= 0; --c) {
    const u64 nc = pack2(differs(lo[c], mlo), differs(hi[c], mhi));
    t = fma2_rn(nc, t, nc);
  }
  return t;
}
// 1 << (low 5 bits of c): a float index biased by 2^23 carries its integer in the low mantissa bits
__device__ __forceinline__ unsigned one_shl_wrap(unsigned c) {
  unsigned d;
  asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(0u), "r"(1u), "r"(c));
  return d;
}

// NP pixel pairs per thread: NP = 2 -> 4 pixels (128-bit loads, 32-bit label stores),
//                            NP = 1 -> 2 pixels (64-bit loads, 16-bit label stores; half the registers, twice the warps)
template <int NP> struct PixIO;
template <> struct PixIO<2> {
  using LabelWord = unsigned;
  static __device__ __forceinline__ void load(const float* p, u64 (&d)[2]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    d[0] = pack2(t.x, t.y);
    d[1] = pack2(t.z, t.w);
  }
  static __device__ __forceinline__ void store(float* p, const float (&x)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
  }
  static __device__ __forceinline__ void store_labels(uint8_t* p, const int (&l)[4]) {
    *reinterpret_cast<unsigned*>(p) = (unsigned)l[0] | ((unsigned)l[1] << 8) | ((unsigned)l[2] << 16) | ((unsigned)l[3] << 24);
  }
  static __device__ __forceinline__ unsigned load_labels(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned*>(p)); }
  // label bytes of 4 pixels from the packed float indices {i0,i1}, {i2,i3}: {i0 + 65536 i2, i1 + 65536 i3} + 2^23, then
  // one byte permute of the two mantissas
  static __device__ __forceinline__ unsigned label_word(const u64 (&idx)[2]) {
    const u64 r = fma2_rn(idx[1], pack2(65536.f, 65536.f), fma2_rn(idx[0], pack2(1.f, 1.f), pack2(8388608.f, 8388608.f)));
    float rl, rh;
    unpack2(r, rl, rh);
    return __byte_perm(__float_as_uint(rl), __float_as_uint(rh), 0x6240);
  }
  static __device__ __forceinline__ void store_label_word(uint8_t* p, unsigned w) { *reinterpret_cast<unsigned*>(p) = w; }
};
template <> struct PixIO<1> {
  using LabelWord = unsigned;
  static __device__ __forceinline__ void load(const float* p, u64 (&d)[1]) {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    d[0] = pack2(t.x, t.y);
  }
  static __device__ __forceinline__ void store(float* p, const float (&x)[2]) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(x[0], x[1]));
  }
  static __device__ __forceinline__ void store_labels(uint8_t* p, const int (&l)[2]) {
    *reinterpret_cast<unsigned short*>(p) = static_cast<unsigned short>((unsigned)l[0] | ((unsigned)l[1] << 8));
  }
  static __device__ __forceinline__ unsigned load_labels(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned short*>(p)); }
  static __device__ __forceinline__ unsigned label_word(const u64 (&idx)[1]) {
    const u64 r = fma2_rn(idx[0], pack2(1.f, 1.f), pack2(8388608.f, 8388608.f));
    float rl, rh;
    unpack2(r, rl, rh);
    return __byte_perm(__float_as_uint(rl), __float_as_uint(rh), 0x0040) & 0xffffu;
  }
  static __device__ __forceinline__ void store_label_word(uint8_t* p, unsigned w) {
    *reinterpret_cast<unsigned short*>(p) = static_cast<unsigned short>(w);
  }
};

// 8 pixels per thread (two 128-bit loads per plane, one 64-bit label store): measured variant of the bulk linear kernel
template <> struct PixIO<4> {
  using LabelWord = unsigned long long;
  static __device__ __forceinline__ void load(const float* p, u64 (&d)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p)), u = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    d[0] = pack2(t.x, t.y); d[1] = pack2(t.z, t.w); d[2] = pack2(u.x, u.y); d[3] = pack2(u.z, u.w);
  }
  static __device__ __forceinline__ void store(float* p, const float (&x)[8]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
    __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(x[4], x[5], x[6], x[7]));
  }
  static __device__ __forceinline__ void store_labels(uint8_t* p, const int (&l)[8]) {
    unsigned long long w = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) w |= static_cast<unsigned long long>(static_cast<unsigned>(l[i]) & 255u) << (8 * i);
    *reinterpret_cast<unsigned long long*>(p) = w;
  }
  static __device__ __forceinline__ LabelWord load_labels(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned long long*>(p)); }
  static __device__ __forceinline__ LabelWord label_word(const u64 (&idx)[4]) {
    const u64 lo[2] = {idx[0], idx[1]}, hi[2] = {idx[2], idx[3]};
    return static_cast<unsigned long long>(PixIO<2>::label_word(lo)) | (static_cast<unsigned long long>(PixIO<2>::label_word(hi)) << 32);
  }
  static __device__ __forceinline__ void store_label_word(uint8_t* p, LabelWord w) { *reinterpret_cast<unsigned long long*>(p) = w; }
};

}  // namespace fuvs
