// scratch microbenchmark (not product): throughput of coalesced fp32 reductions to global memory,
// the pattern a bilinear scatter of grad_x (NHWC fp32 rows of 256 channels) would generate.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
// each warp: `iters` row-adds; row chosen near (local) or random; 256 floats per row
template <int MODE>
__global__ void k(float* buf, int rows, int iters, int local) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = 0; i < iters; ++i) {
    unsigned r = local ? ((unsigned)((long long)warp * rows / nwarps) + (hash(i * 977 + warp) % 512)) % rows
                       : hash(i * 7919u + warp * 104729u) % rows;
    float* row = buf + (size_t)r * 256;
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(row + j * 32 + lane, 1.0f);
    } else if (MODE == 1) {
      red4(row + lane * 4, 1.f, 1.f, 1.f, 1.f);
      red4(row + 128 + lane * 4, 1.f, 1.f, 1.f, 1.f);
    } else {  // plain stores, for reference
      reinterpret_cast<float4*>(row)[lane] = make_float4(1, 1, 1, 1);
      reinterpret_cast<float4*>(row)[32 + lane] = make_float4(1, 1, 1, 1);
    }
  }
}
int main() {
  const int rows = 44800;
  float* buf; cudaMalloc(&buf, (size_t)rows * 256 * 4); cudaMemset(buf, 0, (size_t)rows * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 4, threads = 256, iters = 341;  // 4736 warps * 341 * 1 KB = 1.65 GB
  for (int local = 0; local < 2; ++local)
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<blocks, threads>>>(buf, rows, iters, local);
        if (mode == 1) k<1><<<blocks, threads>>>(buf, rows, iters, local);
        if (mode == 2) k<2><<<blocks, threads>>>(buf, rows, iters, local);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double gb = (double)blocks * threads / 32 * iters * 1024 / 1e9;
      printf("local=%d mode=%s : %.3f ms  %.1f GB/s of fp32 adds (%.2f GB)\n", local,
             mode == 0 ? "red.f32 x8" : mode == 1 ? "red.v4.f32 x2" : "st.v4 x2", best, gb / (best * 1e-3), gb);
    }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
