// scratch microbenchmark (not product): do tcgen05.mma operand reads, LDS/STS traffic and bulk-copy (TMA
// engine) writes share the shared-memory port of an SM?  One persistent CTA per SM:
//   warp 0 : optionally streams 32 KB bulk copies global -> smem (mode bit 2)
//   warp 1 : optionally issues back-to-back 128x256x16 bf16 MMAs from fixed smem operands (mode bit 0)
//   warps 2..9 : optionally run conflict-free LDS.128 x4 + STS.128 loops (mode bit 1)
// and the kernel time for each combination tells whether the three add or overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I slenderobjdet_b200/csrc tools/smem_port_bench.cu -o tools/smem_port_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace sdb::tc;

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(320, 1) bench(int mode, int mma_iters, int lds_iters, int tma_iters, const uint8_t* src,
                                                unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_mma, bar_tma[2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // layout: A 16 KB | B 32 KB | TMA landing 2 x 32 KB | LDS area 64 KB
  const uint32_t sA = base, sB = base + 16384, sT = base + 49152, sL = base + 114688;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma[0], 1);
    mbar_init(&bar_tma[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const unsigned long long t0 = clock64();
  if (warp == 0) {
    if ((mode & 4) && elect_one()) {
      uint32_t ph[2] = {0, 0};
      for (int i = 0; i < tma_iters; ++i) {
        const int b = i & 1;
        if (i >= 2) { mbar_wait(&bar_tma[b], ph[b]); ph[b] ^= 1; }
        mbar_arrive_expect_tx(&bar_tma[b], 32768);
        bulk_g2s(smem_raw + (sT - smem_u32(smem_raw)) + b * 32768, src + (size_t)((i * 148 + blockIdx.x) % 512) * 32768, 32768, &bar_tma[b]);
      }
      for (int b = 0; b < 2; ++b) if (tma_iters > b) mbar_wait(&bar_tma[(tma_iters - 1 - b) & 1], ph[(tma_iters - 1 - b) & 1]);
    }
  } else if (warp == 1) {
    if (mode & 1) {
      const uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
      const uint64_t ad = make_smem_desc_sw128(sA, 16, 1024), bd = make_smem_desc_sw128(sB, 16, 1024);
      if (elect_one()) {
        for (int i = 0; i < mma_iters; ++i) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem_base, ad + 2 * k4, bd + 2 * k4, idesc, 1u);
        }
        umma_commit(&bar_mma);
      }
      __syncwarp();
      mbar_wait(&bar_mma, 0);
    }
  } else if (mode & 2) {
    const int w = warp - 2;
    // each 8-lane group reads one 128-byte row; rows differ per group and per iteration
    uint32_t a = sL + ((w * 4 + (lane >> 3)) * 2048u) % 65536u + (lane & 7) * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < lds_iters; ++i) {
      const uint4 v0 = lds128(a), v1 = lds128(a + 128), v2 = lds128(a + 256), v3 = lds128(a + 384);
      acc.x ^= v0.x ^ v1.y ^ v2.z ^ v3.w;
      sts128(a + 512, acc);
      a = sL + ((a - sL) + 1024u) % 65536u;
    }
    if (acc.x == 0x12345678u) out[1] = acc.x;
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

int main() {
  uint8_t* src;
  cudaMalloc(&src, 512 * 32768);
  cudaMemset(src, 0, 512 * 32768);
  unsigned long long* out;
  cudaMallocManaged(&out, 16);
  const int smem = 1024 + 114688 + 65536 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int mma_iters = 2000;   // x4 MMAs of 128 clk = 1.02 M clk ideal
  const int lds_iters = 6250;   // per warp: 5 x 4 wavefronts = 20 -> 8 warps x 20 x 12500 = 2.0 M wavefront-clk
  const int tma_iters = 2000;   // 2000 x 32 KB = 64 MB per SM -> 512 K clk at 128 B/clk
  for (int mode = 1; mode < 8; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      bench<<<148, 320, smem>>>(mode, mma_iters, lds_iters, tma_iters, src, out);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("mode %d (mma=%d lds=%d tma=%d): %8.1f us  %llu clk   %s\n", mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1,
                      ms * 1e3, out[0], cudaGetErrorString(err));
    }
  }
  printf("ideal: mma %d x 4 x 128 clk = %d clk; lds full %d / half %d wavefront-clk; tma %d x 256 clk = %d clk\n", mma_iters,
         mma_iters * 512, 8 * 20 * lds_iters, 8 * 20 * lds_iters / 2, tma_iters, tma_iters * 256);
  return 0;
}
