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

// NP pixel pairs per thread: NP = 2 -> 4 pixels (128-bit loads, 32-bit label stores),
//                            NP = 1 -> 2 pixels (64-bit loads, 16-bit label stores; half the registers, twice the warps)
template <int NP> struct PixIO;
template <> struct PixIO<2> {
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
};
template <> struct PixIO<1> {
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
};

}  // namespace fuvs
