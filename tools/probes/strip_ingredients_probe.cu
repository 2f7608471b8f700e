// Probe (not part of the library): which ingredient of the dense strip kernel's block loop costs what?
// Starts from the pure shared-memory gather of lds_gather_probe.cu (planar layout, iid taps, 512 threads, one CTA per SM,
// 2 pixels per thread and iteration = one 1024-pixel block per CTA and iteration) and adds, cumulatively:
//   A  200 dependent FMAs per iteration (stand-in for the coordinate arithmetic)
//   B  10 coalesced 4-byte global stores per thread and iteration (the state writes)
//   C  2 prefetched 8-byte global loads per thread and iteration (the flow vectors, two iterations ahead)
//   B' / C'  the same with the stores of the 4+1 layout (one 128-bit + one 32-bit store per pixel)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o strip_ingredients_probe strip_ingredients_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int BOXW = 192, ROWS = 56, NPIX = BOXW * ROWS, C = 5;
constexpr int SLOT_BYTES = 8 * BOXW * C * 4;            // 30 720

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

template <int LEVEL>
__global__ void __launch_bounds__(512, 1) probe(float* out, const float* src, const float2* grid, long long* cyc, int iters,
                                                long long plane) {
  extern __shared__ __align__(128) float sm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + NPIX * C);
  unsigned* done = reinterpret_cast<unsigned*>(bar + 8);
  for (int i = threadIdx.x; i < NPIX * C; i += blockDim.x) sm[i] = static_cast<float>(i & 1023) * 0.001f;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 8; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + b)));
    done[0] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned seed = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  float acc[2][C];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[r][c] = 0.f;
  float dummy = static_cast<float>(threadIdx.x);
  const long long base = static_cast<long long>(blockIdx.x) * 1024 * 64;     // this CTA's region of the planes
  float2 g0 = make_float2(0.f, 0.f), g1 = g0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const long long pix = base + static_cast<long long>(it % 64) * 1024 + threadIdx.x;
    float2 g2 = make_float2(0.f, 0.f);
    if (LEVEL == 3 || LEVEL == 7) {
      const float2 ga = __ldg(grid + pix), gb = __ldg(grid + pix + 512);
      g2 = make_float2(ga.x + gb.x, ga.y + gb.y);
    }
    int p[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const unsigned rnd = lcg(seed);
      p[r] = ((rnd >> 10) % (ROWS - 1)) * BOXW + rnd % (BOXW - 1);
    }
    if (LEVEL >= 1) {
#pragma unroll
      for (int a = 0; a < 200; ++a) dummy = fmaf(dummy, 1.0001f, 0.5f + g0.x);
    }
    const float w0 = 0.25f + dummy * 1e-30f;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* pl = sm + c * NPIX + p[r];
        float v0, v1, v2, v3;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(smem_u32(pl)));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v1) : "r"(smem_u32(pl)));
        asm volatile("ld.shared.f32 %0, [%1+768];" : "=f"(v2) : "r"(smem_u32(pl)));
        asm volatile("ld.shared.f32 %0, [%1+772];" : "=f"(v3) : "r"(smem_u32(pl)));
        acc[r][c] = fmaf(v0, w0, fmaf(v1, w0, fmaf(v2, w0, fmaf(v3, w0, acc[r][c] * 0.5f))));
      }
    }
    if (LEVEL >= 2 && LEVEL < 6) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) out[c * plane + pix + r * 512] = acc[r][c];
    }
    if (LEVEL >= 6) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        reinterpret_cast<float4*>(out)[pix + r * 512] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        out[4 * plane + pix + r * 512] = acc[r][4];
      }
    }
    g0 = g1;
    g1 = g2;
  }
  const long long t1 = clock64();
  float s = dummy;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[r][c];
  if (s == 1.2345f) out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int LEVEL>
void run(const char* name, float* out, const float* src, const float2* grid, long long* cyc, int nsm, long long plane) {
  const int iters = 1500;
  const size_t smem = sizeof(float) * NPIX * C + 256;
  CK(cudaFuncSetAttribute(probe<LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  for (int rep = 0; rep < 2; ++rep) {
    probe<LEVEL><<<nsm, 512, smem>>>(out, src, grid, cyc, iters, plane);
    CK(cudaDeviceSynchronize());
  }
  long long h[256];
  CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  double m = 0, mx = 0;
  for (int i = 0; i < nsm; ++i) { m += h[i]; mx = h[i] > mx ? h[i] : mx; }
  m /= nsm;
  printf("%-70s %8.1f cycles / block (mean over SMs), slowest SM %8.1f\n", name, m / iters, mx / iters);
  fflush(stdout);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  const long long plane = static_cast<long long>(nsm) * 1024 * 64 + 2048;
  float *out, *src;
  float2* grid;
  long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * plane * C));
  CK(cudaMalloc(&src, sizeof(float) * plane * 4 + SLOT_BYTES));
  CK(cudaMalloc(&grid, sizeof(float2) * plane));
  CK(cudaMalloc(&cyc, sizeof(long long) * 256));
  CK(cudaMemset(src, 0, sizeof(float) * plane * 4 + SLOT_BYTES));
  CK(cudaMemset(grid, 0, sizeof(float2) * plane));
  printf("%s, %d SMs\n", prop.name, nsm);
  fflush(stdout);
  run<0>("0  gather only (20 LDS.32 per pixel, iid)", out, src, grid, cyc, nsm, plane);
  run<1>("A  + 200 dependent FMAs", out, src, grid, cyc, nsm, plane);
  run<2>("B  + 10 global stores per thread", out, src, grid, cyc, nsm, plane);
  run<3>("C  + 2 prefetched 8-byte global loads per thread", out, src, grid, cyc, nsm, plane);
  run<6>("B' as B, but the 10 stores as 2 x (128-bit + 32-bit) per thread (4+1 layout)", out, src, grid, cyc, nsm, plane);
  run<7>("C' as C with the 4+1 stores", out, src, grid, cyc, nsm, plane);
  return 0;
}
