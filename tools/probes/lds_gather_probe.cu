// Probe (not part of the library): what does a random 2x2-footprint gather of C=5 channels cost on the shared-memory
// pipe, as a function of the layout of the staged window?
//   V0  planar [c][row][x]            : 20 LDS.32 per pixel (the round-1 strip kernel)
//   V1  [row][x][4] + planar c4       : 4 LDS.128 + 4 LDS.32 per pixel
//   V2  [row][x][2] x2 + planar c4    : 8 LDS.64 + 4 LDS.32 per pixel
// ADDR 0: iid random tap position per lane (SURVEY.md 8d's per-pixel jitter); 1: coherent (lanes along x).
// 512 threads, one CTA per SM, 2 pixels per thread and iteration, like dense_strip_kernel.  ALU: dependent dummy FMAs per
// iteration (stands for the coordinate set-up), so that the LDS pipe is not the only consumer of issue slots.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_gather_probe lds_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int BOXW = 192, ROWS = 56, NPIX = BOXW * ROWS;   // 7 slots x 8 rows
constexpr int C = 5;

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int V, int ADDR, int ALU>
__global__ void __launch_bounds__(512, 1) probe(float* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < NPIX * C; i += blockDim.x) sm[i] = static_cast<float>(i & 1023) * 0.001f;
  __syncthreads();
  unsigned seed = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  const int lane_x = threadIdx.x & 127;
  float acc[2][C];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[r][c] = 0.f;
  float dummy = static_cast<float>(threadIdx.x);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    int p[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const unsigned rnd = lcg(seed);
      if (ADDR == 0) {
        const int x = rnd % (BOXW - 1), y = (rnd >> 10) % (ROWS - 1);
        p[r] = y * BOXW + x;
      } else {
        const int y = (rnd >> 10) % (ROWS - 1);      // per-thread row differs little inside a warp: use the warp's
        const int yw = __shfl_sync(0xffffffffu, y, 0);
        p[r] = yw * BOXW + 20 + lane_x + (rnd & 1);  // sub-pixel jitter across lanes
      }
    }
#pragma unroll
    for (int a = 0; a < ALU; ++a) dummy = fmaf(dummy, 1.0001f, 0.5f);
    float w0 = 0.25f + dummy * 1e-30f;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (V == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float* pl = sm + c * NPIX + p[r];
          acc[r][c] = fmaf(pl[0], w0, fmaf(pl[1], w0, fmaf(pl[BOXW], w0, fmaf(pl[BOXW + 1], w0, acc[r][c]))));
        }
      } else if (V == 1) {
        const float4* q = reinterpret_cast<const float4*>(sm) + p[r];
        const float* pl = sm + 4 * NPIX + p[r];
        const float4 a = q[0], b = q[1], c_ = q[BOXW], d = q[BOXW + 1];
        acc[r][0] = fmaf(a.x, w0, fmaf(b.x, w0, fmaf(c_.x, w0, fmaf(d.x, w0, acc[r][0]))));
        acc[r][1] = fmaf(a.y, w0, fmaf(b.y, w0, fmaf(c_.y, w0, fmaf(d.y, w0, acc[r][1]))));
        acc[r][2] = fmaf(a.z, w0, fmaf(b.z, w0, fmaf(c_.z, w0, fmaf(d.z, w0, acc[r][2]))));
        acc[r][3] = fmaf(a.w, w0, fmaf(b.w, w0, fmaf(c_.w, w0, fmaf(d.w, w0, acc[r][3]))));
        acc[r][4] = fmaf(pl[0], w0, fmaf(pl[1], w0, fmaf(pl[BOXW], w0, fmaf(pl[BOXW + 1], w0, acc[r][4]))));
      } else {
        const float2* q0 = reinterpret_cast<const float2*>(sm) + p[r];
        const float2* q1 = reinterpret_cast<const float2*>(sm + 2 * NPIX) + p[r];
        const float* pl = sm + 4 * NPIX + p[r];
        const float2 a = q0[0], b = q0[1], c_ = q0[BOXW], d = q0[BOXW + 1];
        const float2 e = q1[0], f = q1[1], g = q1[BOXW], h = q1[BOXW + 1];
        acc[r][0] = fmaf(a.x, w0, fmaf(b.x, w0, fmaf(c_.x, w0, fmaf(d.x, w0, acc[r][0]))));
        acc[r][1] = fmaf(a.y, w0, fmaf(b.y, w0, fmaf(c_.y, w0, fmaf(d.y, w0, acc[r][1]))));
        acc[r][2] = fmaf(e.x, w0, fmaf(f.x, w0, fmaf(g.x, w0, fmaf(h.x, w0, acc[r][2]))));
        acc[r][3] = fmaf(e.y, w0, fmaf(f.y, w0, fmaf(g.y, w0, fmaf(h.y, w0, acc[r][3]))));
        acc[r][4] = fmaf(pl[0], w0, fmaf(pl[1], w0, fmaf(pl[BOXW], w0, fmaf(pl[BOXW + 1], w0, acc[r][4]))));
      }
    }
  }
  const long long t1 = clock64();
  float s = dummy;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V, int ADDR, int ALU>
void run(const char* name, float* out, long long* cyc, int nsm) {
  const int iters = 2000;
  const size_t smem = sizeof(float) * NPIX * C;
  CK(cudaFuncSetAttribute(probe<V, ADDR, ALU>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  probe<V, ADDR, ALU><<<nsm, 512, smem>>>(out, cyc, iters);
  CK(cudaDeviceSynchronize());
  probe<V, ADDR, ALU><<<nsm, 512, smem>>>(out, cyc, iters);
  CK(cudaDeviceSynchronize());
  long long h[256];
  CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  double m = 0;
  for (int i = 0; i < nsm; ++i) m += h[i];
  m /= nsm;
  // one iteration = 16 warps x 2 pixels x 32 lanes = 1024 pixels per SM
  printf("%-44s %8.1f cycles / iteration (1024 px per SM)  -> %.3f cycles per warp-pixel-row (32 px)\n", name, m / iters,
         m / iters / 32.0);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * nsm * 512));
  CK(cudaMalloc(&cyc, sizeof(long long) * 256));
  printf("%s, %d SMs\n", prop.name, nsm);
  run<0, 0, 0>("V0 planar 20xLDS.32, iid, no ALU", out, cyc, nsm);
  run<1, 0, 0>("V1 4xLDS.128+4xLDS.32, iid, no ALU", out, cyc, nsm);
  run<2, 0, 0>("V2 8xLDS.64+4xLDS.32, iid, no ALU", out, cyc, nsm);
  run<0, 1, 0>("V0 planar, coherent, no ALU", out, cyc, nsm);
  run<1, 1, 0>("V1 4+1, coherent, no ALU", out, cyc, nsm);
  run<2, 1, 0>("V2 2+2+1, coherent, no ALU", out, cyc, nsm);
  run<0, 0, 200>("V0 planar, iid, 200 dep FMAs", out, cyc, nsm);
  run<1, 0, 200>("V1 4+1, iid, 200 dep FMAs", out, cyc, nsm);
  run<2, 0, 200>("V2 2+2+1, iid, 200 dep FMAs", out, cyc, nsm);
  run<0, 1, 200>("V0 planar, coherent, 200 dep FMAs", out, cyc, nsm);
  run<1, 1, 200>("V1 4+1, coherent, 200 dep FMAs", out, cyc, nsm);
  return 0;
}
