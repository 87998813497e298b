// debug_umma.cu -- single-CTA tcgen05 probe: D[128,N] = A[128,K] * B[N,K]^T with every operand
// major-ness the DCN kernels use (K-major and MN-major, 128B swizzle), B optionally brought in by a
// bulk async copy from a pre-swizzled global image.  Not on the product path: it pins the
// descriptor / swizzle / TMEM-lane conventions of tc_common.cuh against a plain matmul
// (tests/test_gpu_umma_probe.py).
// TEST INFRASTRUCTURE: compiled by tests/test_gpu_umma_probe.py into tests/probe/libsdb_probe.so, never linked into
// libslender_b200.so.
#include "../../slenderobjdet_b200/csrc/common.cuh"
#include "../../slenderobjdet_b200/csrc/tc_common.cuh"

namespace sdb {
// the two symbols common.cuh's macros expect from api.cu (this file is built on its own)
long long g_launches = 0;
void set_error(const char* fmt, ...) { fprintf(stderr, "sdb probe: %s\n", fmt); }
namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                  const uint8_t* __restrict__ B_img, float* __restrict__ D, int N, int K, int a_mn,
                  int b_mn, int use_bulk) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
  uint8_t* sA = sm;
  uint8_t* sB = sm + a_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  // operand fill by plain stores (what the gather producers do)
  for (int e = tid; e < 128 * K; e += blockDim.x) {
    const int m = e / K, k = e - m * K;
    uint32_t off;
    if (!a_mn) off = (k >> 6) * (128u * 128u) + sw128_offset(m, (k & 63) >> 3) + (k & 7) * 2;
    else       off = (m >> 6) * ((uint32_t)K * 128u) + sw128_offset(k, (m & 63) >> 3) + (m & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[e];
  }
  if (!use_bulk) {
    for (int e = tid; e < N * K; e += blockDim.x) {
      const int n = e / K, k = e - n * K;
      uint32_t off;
      if (!b_mn) off = (k >> 6) * ((uint32_t)N * 128u) + sw128_offset(n, (k & 63) >> 3) + (k & 7) * 2;
      else       off = (n >> 6) * ((uint32_t)K * 128u) + sw128_offset(k, (n & 63) >> 3) + (n & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[e];
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      if (use_bulk) {
        mbar_arrive_expect_tx(&bar_b, b_bytes);
        bulk_g2s(sB, B_img, b_bytes, &bar_b);
        mbar_wait(&bar_b, 0);
      }
      tc_fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (int s = 0; s < K / 16; ++s) {
        const int k0 = s * 16;
        uint64_t ad, bd;
        if (!a_mn) ad = make_smem_desc_sw128(a0 + (k0 >> 6) * (128u * 128u) + (k0 & 63) * 2, 16, 1024);
        else       ad = make_smem_desc_sw128(a0 + k0 * 128u, (uint32_t)K * 128u, 1024);
        if (!b_mn) bd = make_smem_desc_sw128(b0 + (k0 >> 6) * ((uint32_t)N * 128u) + (k0 & 63) * 2, 16, 1024);
        else       bd = make_smem_desc_sw128(b0 + k0 * 128u, (uint32_t)K * 128u, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, s > 0);
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    const int m = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < N) D[(size_t)m * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace
}  // namespace sdb

extern "C" int sdb_debug_umma_gemm(const void* A, const void* B, const void* B_img, float* D, int N,
                                   int K, int a_mn, int b_mn, int use_bulk, void* stream) {
  using namespace sdb;
  SDB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K % 64 == 0 && K <= 256,
              SDB_ERR_INVALID, "probe needs 16<=N<=256 (mult of 16), K in {64,128,192,256}");
  SDB_REQUIRE(!(a_mn || b_mn) || (N % 64 == 0), SDB_ERR_INVALID, "MN-major probe needs N % 64 == 0");
  const size_t smem = (size_t)(128 + N) * K * 2 + 1024;
  SDB_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (const uint8_t*)B_img, D, N, K, a_mn, b_mn, use_bulk); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}
