// Probe (not part of the library): can the texture unit replace the shared-memory ring of the dense warp step?
// One dense step = both chains x C=5 planes of 1080x1920 floats sampled at per-pixel flow positions (4 taps each).
// Variant T: state planes live in CUDA arrays, the 4 taps of one plane come from ONE tex2Dgather (point addressing at
// the shared corner of the 2x2 footprint, so the selection is exact), results go to linear memory or (S) to surfaces.
// Variant G: the same arithmetic with four plain global loads per plane (the library's direct kernel).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tex_gather_probe tex_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int C = 5, H = 1080, W = 1920;

struct Maps {
  cudaTextureObject_t tex[2][C];
  cudaSurfaceObject_t surf[2][C];
  const float* src[2];      // [C][H][W] linear copies of the same planes
  const float2* grid[2];    // [H][W] normalised (x, y)
  float* dst[2];            // [C][H][W]
};

__device__ __forceinline__ float src_index(float coord, int size) {
  const float t = __fadd_rn(coord, 1.f);
  const float u = __fsub_rn(__fmul_rn(t, static_cast<float>(size)), 1.f);
  const float r = __fmul_rn(u, 0.5f);
  return fminf(static_cast<float>(size - 1), fmaxf(r, 0.f));
}

template <int MODE>   // 0: gather -> linear stores, 1: gather -> surface stores, 2: global loads -> linear stores
__global__ void __launch_bounds__(256) step_kernel(Maps m) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), side = blockIdx.z;
  if (x >= W || y >= H) return;
  const float2 g = __ldg(m.grid[side] + y * W + x);
  const float ix = src_index(g.x, W), iy = src_index(g.y, H);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float tx = __fsub_rn(ix, fx), ty = __fsub_rn(iy, fy);
  const float wnw = __fmul_rn(__fsub_rn(1.f, tx), __fsub_rn(1.f, ty)), wne = __fmul_rn(tx, __fsub_rn(1.f, ty));
  const float wsw = __fmul_rn(__fsub_rn(1.f, tx), ty), wse = __fmul_rn(tx, ty);
  const bool dx = x0 + 1 < W, dy = y0 + 1 < H;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v00, v01, v10, v11;
    if (MODE == 2) {
      const float* p = m.src[side] + (c * H + y0) * W + x0;
      v00 = __ldg(p);
      v01 = __ldg(p + (dx ? 1 : 0));
      v10 = __ldg(p + (dy ? W : 0));
      v11 = __ldg(p + (dy ? W : 0) + (dx ? 1 : 0));
    } else {
      // footprint of a gather at (x0 + 1, y0 + 1): texels x0, x0+1 x y0, y0+1; components: w = (i, j), z = (i+1, j), x = (i, j+1), y = (i+1, j+1)
      const float4 q = tex2Dgather<float4>(m.tex[side][c], fx + 1.0f, fy + 1.0f, 0);
      v00 = q.w; v01 = q.z; v10 = q.x; v11 = q.y;
    }
    float acc = __fmul_rn(v00, wnw);
    if (dx) acc = __fmaf_rn(v01, wne, acc);
    if (dy) acc = __fmaf_rn(v10, wsw, acc);
    if (dx && dy) acc = __fmaf_rn(v11, wse, acc);
    if (MODE == 1) surf2Dwrite(acc, m.surf[side][c], x * 4, y);
    else m.dst[side][(c * H + y) * W + x] = acc;
  }
}

int main(int argc, char** argv) {
  const float jitter = argc > 1 ? atof(argv[1]) : 0.05f;
  const size_t plane = static_cast<size_t>(H) * W;
  std::vector<float> hsrc(2 * C * plane);
  std::vector<float2> hgrid(2 * plane);
  srand(1234);
  for (auto& v : hsrc) v = (rand() / (float)RAND_MAX - 0.5f) * 8.f;
  for (int s = 0; s < 2; ++s)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float bx = ((x + 0.5f) / W) * 2.f - 1.f, by = ((y + 0.5f) / H) * 2.f - 1.f;
        hgrid[s * plane + y * W + x] = make_float2(bx + (rand() / (float)RAND_MAX - 0.5f) * jitter, by + (rand() / (float)RAND_MAX - 0.5f) * jitter);
      }
  Maps m;
  float *dsrc, *ddst, *dref;
  float2* dgrid;
  CK(cudaMalloc(&dsrc, hsrc.size() * 4));
  CK(cudaMalloc(&ddst, hsrc.size() * 4));
  CK(cudaMalloc(&dref, hsrc.size() * 4));
  CK(cudaMalloc(&dgrid, hgrid.size() * 8));
  CK(cudaMemcpy(dsrc, hsrc.data(), hsrc.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dgrid, hgrid.data(), hgrid.size() * 8, cudaMemcpyHostToDevice));
  cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
  cudaArray_t arr[2][C], out[2][C];
  for (int s = 0; s < 2; ++s) {
    m.src[s] = dsrc + s * C * plane;
    m.dst[s] = ddst + s * C * plane;
    m.grid[s] = dgrid + s * plane;
    for (int c = 0; c < C; ++c) {
      CK(cudaMallocArray(&arr[s][c], &fd, W, H, cudaArrayTextureGather | cudaArraySurfaceLoadStore));
      CK(cudaMallocArray(&out[s][c], &fd, W, H, cudaArrayTextureGather | cudaArraySurfaceLoadStore));
      CK(cudaMemcpy2DToArray(arr[s][c], 0, 0, dsrc + (s * C + c) * plane, W * 4, W * 4, H, cudaMemcpyDeviceToDevice));
      cudaResourceDesc rd;
      memset(&rd, 0, sizeof(rd));
      rd.resType = cudaResourceTypeArray;
      rd.res.array.array = arr[s][c];
      cudaTextureDesc td;
      memset(&td, 0, sizeof(td));
      td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
      td.filterMode = cudaFilterModePoint;
      td.readMode = cudaReadModeElementType;
      td.normalizedCoords = 0;
      CK(cudaCreateTextureObject(&m.tex[s][c], &rd, &td, nullptr));
      rd.res.array.array = out[s][c];
      CK(cudaCreateSurfaceObject(&m.surf[s][c], &rd));
    }
  }
  dim3 grid((W + 31) / 32, (H + 7) / 8, 2), block(256);
  // reference (global loads) and gather result
  m.dst[0] = dref; m.dst[1] = dref + C * plane;
  step_kernel<2><<<grid, block>>>(m);
  m.dst[0] = ddst; m.dst[1] = ddst + C * plane;
  step_kernel<0><<<grid, block>>>(m);
  CK(cudaDeviceSynchronize());
  std::vector<float> a(hsrc.size()), b(hsrc.size());
  CK(cudaMemcpy(a.data(), dref, a.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), ddst, b.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < a.size(); ++i) bad += memcmp(&a[i], &b[i], 4) != 0;
  printf("jitter %.3f: gather vs global loads: %zu of %zu values differ\n", jitter, bad, a.size());
  // surface variant: read back one array
  step_kernel<1><<<grid, block>>>(m);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy2DFromArray(b.data(), W * 4, out[1][3], 0, 0, W * 4, H, cudaMemcpyDeviceToHost));
  bad = 0;
  for (size_t i = 0; i < plane; ++i) bad += memcmp(&a[(C + 3) * plane + i], &b[i], 4) != 0;
  printf("surface output plane (1,3): %zu of %zu values differ\n", bad, plane);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      for (int i = 0; i < 20; ++i) {
        if (mode == 0) step_kernel<0><<<grid, block>>>(m);
        else if (mode == 1) step_kernel<1><<<grid, block>>>(m);
        else step_kernel<2><<<grid, block>>>(m);
      }
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("mode %d (%s): %.1f us per step (both sides, C=%d)\n", mode, mode == 0 ? "tex2Dgather -> linear" : mode == 1 ? "tex2Dgather -> surface" : "global loads -> linear", best * 1e3f / 20, C);
  }
  return 0;
}
