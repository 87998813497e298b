// dcn_tc_shared.cuh -- helpers shared by the tensor-core DCN translation units (dcn_tc.cu: forward,
// dcn_tc_bwd.cu: backward data / backward weight).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace sdb {

size_t tc_bwd_workspace_bytes(int op, const Geo& g);
// grad_input as a gathered implicit GEMM over the transposed sampling index (dcn_tc.cu, MODE_DX)
int tc_dx(const void* w, const void* gy_nhwc, const void* desc, const int* start, const void* entries,
          uint8_t* wimg, void* gx, const Geo& g, int okb, int io_dtype, cudaStream_t st);

// forward with the sampling window of each tile staged in shared memory by TMA (dcn_tc_win.cu)
bool tc_win_supported(const Geo& g);
int tc_forward_win(const __nv_bfloat16* xp, const float* off, const float* mask, const void* w, const void* bias,
                   uint8_t* wimg, float* bias32, void* out, const Geo& g, int io_dtype, cudaStream_t st);

namespace tcshared {
using namespace tc;

constexpr int TILE_M = 128;

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// Pixel walk order of the tensor-core kernels.  Output pixels of one image are enumerated band by
// band (g.th rows), inside a band column-group by column-group (g.tw columns), row-major inside a
// group; a 128-pixel tile is 128 consecutive indices q of that order, i.e. (for th*tw == 128) a
// compact th x tw patch.  Compact patches are what lets the bilinear gather hit in L1: the 3x3
// neighbourhood of an 8x16 patch is ~180 input pixels instead of 3 rows x 130.  Partial bands /
// groups at the image border are shorter, never padded, so ceil(P/128) tiles cover P pixels.
// th = 1, tw >= Wo degenerates to plain row-major order.
// generic form: grid of Hd x Wd pixels per image walked in th x tw blocks
__host__ __device__ __forceinline__ void decode_pos(int Hd, int Wd, int th, int tw, long long q, int& n, int& y, int& x) {
  const int hw = Hd * Wd;
  n = (int)(q / hw);
  const int r = (int)(q - (long long)n * hw);
  const int band_px = th * Wd;
  const int band = r / band_px, rb = r - band * band_px;
  const int rows_b = min(th, Hd - band * th);
  const int grp_px = rows_b * tw;
  const int cg = rb / grp_px, rg = rb - cg * grp_px;
  const int cols_g = min(tw, Wd - cg * tw);
  const int dy = rg / cols_g;
  y = band * th + dy;
  x = cg * tw + (rg - dy * cols_g);
}
// inverse of decode_pos
__host__ __device__ __forceinline__ long long encode_pos(int Hd, int Wd, int th, int tw, int n, int y, int x) {
  const int band = y / th, dy = y - band * th;
  const int rows_b = min(th, Hd - band * th);
  const int cg = x / tw, dx = x - cg * tw;
  const int cols_g = min(tw, Wd - cg * tw);
  return (long long)n * Hd * Wd + (long long)band * th * Wd + cg * (rows_b * tw) + dy * cols_g + dx;
}
__device__ __forceinline__ void decode_q(const Geo& g, long long q, int& n, int& ho, int& wo) {
  decode_pos(g.Ho, g.Wo, g.th, g.tw, q, n, ho, wo);
}

// ------------------------------------------------------------------------------------------------
// layout / dtype conversion kernels
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// [N][C][HW] (T) -> [N][HW][Cd] bf16, Cd >= C (channels C..Cd-1 zero).  Tile 64 channels x 32 pixels.
template <typename T>
__global__ void __launch_bounds__(256) pack_nhwc_kernel(const T* __restrict__ src,
                                                        __nv_bfloat16* __restrict__ dst, int C, int HW, int Cd) {
  __shared__ float s[64][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* sp = src + ((size_t)n * C + c0) * HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ty + 8 * j;
    s[c][tx] = (c0 + c < C && p0 + tx < HW) ? to_f32(sp[(size_t)c * HW + p0 + tx]) : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* dp = dst + ((size_t)n * HW + p0) * Cd + c0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = ty + 8 * j;
    if (p0 + p < HW && c0 + 2 * tx < Cd)
      *reinterpret_cast<__nv_bfloat162*>(dp + (size_t)p * Cd + 2 * tx) =
          __floats2bfloat162_rn(s[2 * tx][p], s[2 * tx + 1][p]);
  }
}

// bf16 source with an even pixel count per plane: 64 channels x 64 pixels per block, bf16x2 loads (128 B per
// warp instruction), one 16-byte store (8 channels of one pixel) per thread and round -- half the load and a
// quarter of the store instructions of the generic kernel.  Requires HW % 2 == 0 and Cd % 8 == 0.
static __global__ void __launch_bounds__(256) pack_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ src,
                                                             __nv_bfloat16* __restrict__ dst, int C, int HW, int Cd) {
  __shared__ uint32_t s[64][33];   // [channel][pixel pair]
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const __nv_bfloat16* sp = src + ((size_t)n * C + c0) * HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ty + 8 * j, px = p0 + 2 * tx;
    s[c][tx ^ ((c >> 5) << 4)] = (c0 + c < C && px < HW) ? *reinterpret_cast<const uint32_t*>(sp + (size_t)c * HW + px) : 0u;
  }
  __syncthreads();
  // thread -> (pixel, 8-channel chunk): 64 pixels x 8 chunks = 512 stores, two rounds
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = threadIdx.x + 256 * j, px = i >> 3, ch = i & 7;
    if (p0 + px < HW && c0 + ch * 8 < Cd) {
      const int pp = px >> 1, hi = px & 1;
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = s[ch * 8 + k][pp ^ ((ch >> 2) << 4)];   // column swizzle: chunks ch and ch + 4 on different banks
      uint4 pk;
      pk.x = hi ? __byte_perm(v[0], v[1], 0x7632) : __byte_perm(v[0], v[1], 0x5410);
      pk.y = hi ? __byte_perm(v[2], v[3], 0x7632) : __byte_perm(v[2], v[3], 0x5410);
      pk.z = hi ? __byte_perm(v[4], v[5], 0x7632) : __byte_perm(v[4], v[5], 0x5410);
      pk.w = hi ? __byte_perm(v[6], v[7], 0x7632) : __byte_perm(v[6], v[7], 0x5410);
      *reinterpret_cast<uint4*>(dst + ((size_t)n * HW + p0 + px) * Cd + c0 + ch * 8) = pk;
    }
  }
}

__device__ __forceinline__ void fma8(float (&acc)[8], const uint4 v, float w) {
  acc[0] = fmaf(w, __uint_as_float(v.x << 16), acc[0]);
  acc[1] = fmaf(w, __uint_as_float(v.x & 0xffff0000u), acc[1]);
  acc[2] = fmaf(w, __uint_as_float(v.y << 16), acc[2]);
  acc[3] = fmaf(w, __uint_as_float(v.y & 0xffff0000u), acc[3]);
  acc[4] = fmaf(w, __uint_as_float(v.z << 16), acc[4]);
  acc[5] = fmaf(w, __uint_as_float(v.z & 0xffff0000u), acc[5]);
  acc[6] = fmaf(w, __uint_as_float(v.w << 16), acc[6]);
  acc[7] = fmaf(w, __uint_as_float(v.w & 0xffff0000u), acc[7]);
}

// ------------------------------------------------------------------------------------------------
// Gather of one (tile, tap, channel chunk) into a 128B-swizzled operand stage, shared by the
// forward kernel (A operand, K-major) and the weight-gradient kernel (B operand, MN-major): both
// want rows = pixels, 64 channels (128 B) per row, k-blocks TILE_M*128 bytes apart.
//
// Per (pixel, tap) a 32-byte descriptor sits in shared memory: the four corner rows as offsets in
// 16-byte units into the NHWC bf16 input, and the four bilinear weights (x mask) as packed
// bf16x2 (w, w).  A group of LPP lanes handles one pixel, 8 channels per lane: four 16-byte loads,
// 16 packed bf16x2 FMAs (HFMA2.BF16), one 16-byte swizzled store -- ~35 instructions per 2 pixels
// where fp32 interpolation with unpack/pack needed ~110.  The sampled value is rounded to bf16
// anyway (it is a tensor-core operand); interpolating in bf16 adds ~3 more roundings per sample.
struct __align__(16) GDesc {
  uint32_t off[4];  // corner row offset, units of 16 bytes (0 for a corner that does not contribute)
  uint32_t w2[4];   // bf16x2 (w, w); 0 for a corner that does not contribute
};

// overflow descriptor of the transposed index (MODE_DX): four more entries of the list of one row
struct __align__(16) ODesc {
  uint4 o;   // row offsets of dY, units of 16 bytes
  uint4 m;   // x = w0 | w1 << 16, y = w2 | w3 << 16 (bf16 weights, 0 = unused slot), z = row in tile, w = 0
};

__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// x16: input viewed as 16-byte vectors, already advanced to (chunk base + this lane's 8 channels).
// sdesc: this warp's PIXW descriptors.  stage: operand stage base.  row0: first tile row of the warp.
template <int LPP, int PIXW>
__device__ __forceinline__ void gather_stage_bf16(const uint4* __restrict__ x16, const GDesc* __restrict__ sdesc,
                                                  uint8_t* __restrict__ stage, int row0, int lane) {
  constexpr int PPI = 32 / LPP;
  constexpr int ITERS = PIXW / PPI;
  constexpr int U = ITERS < 4 ? ITERS : 4;   // iterations whose 4 loads each are in flight together
  const int grp = lane / LPP, lig = lane % LPP;
  uint8_t* dst = stage + (lig >> 3) * (TILE_M * 128);
#pragma unroll 1
  for (int it0 = 0; it0 < ITERS; it0 += U) {
    uint4 o[U], w[U], v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int src = (it0 + u) * PPI + grp;
      o[u] = *reinterpret_cast<const uint4*>(sdesc[src].off);
      w[u] = *reinterpret_cast<const uint4*>(sdesc[src].w2);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u][0] = __ldg(x16 + o[u].x);
      v[u][1] = __ldg(x16 + o[u].y);
      v[u][2] = __ldg(x16 + o[u].z);
      v[u][3] = __ldg(x16 + o[u].w);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint4 a;
      a.x = bf2_fma(w[u].w, v[u][3].x, bf2_fma(w[u].z, v[u][2].x, bf2_fma(w[u].y, v[u][1].x, bf2_mul(w[u].x, v[u][0].x))));
      a.y = bf2_fma(w[u].w, v[u][3].y, bf2_fma(w[u].z, v[u][2].y, bf2_fma(w[u].y, v[u][1].y, bf2_mul(w[u].x, v[u][0].y))));
      a.z = bf2_fma(w[u].w, v[u][3].z, bf2_fma(w[u].z, v[u][2].z, bf2_fma(w[u].y, v[u][1].z, bf2_mul(w[u].x, v[u][0].z))));
      a.w = bf2_fma(w[u].w, v[u][3].w, bf2_fma(w[u].z, v[u][2].w, bf2_fma(w[u].y, v[u][1].w, bf2_mul(w[u].x, v[u][0].w))));
      *reinterpret_cast<uint4*>(dst + sw128_offset(row0 + (it0 + u) * PPI + grp, lig & 7)) = a;
    }
  }
}

template <typename T>
inline int pack_input(const void* x, __nv_bfloat16* xp, const Geo& g, cudaStream_t st) {
  const int HW = g.H * g.W;
  if (sizeof(T) == 2 && HW % 2 == 0 && g.C % 8 == 0) {
    dim3 grid2(cdiv(HW, 64), cdiv(g.C, 64), g.N);
    pack_nhwc_bf16_kernel<<<grid2, 256, 0, st>>>((const __nv_bfloat16*)x, xp, g.C, HW, g.C); SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
    return SDB_OK;
  }
  dim3 grid(cdiv(HW, 32), cdiv(g.C, 64), g.N);
  pack_nhwc_kernel<T><<<grid, 256, 0, st>>>((const T*)x, xp, g.C, HW, g.C); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}
// grad_out [N][O][HWo] -> [N][HWo][Od] bf16 (Od = okb*64 >= O, zero padded)
template <typename T>
inline int pack_grad_nhwc(const void* gy, __nv_bfloat16* gyp, const Geo& g, int Od, cudaStream_t st) {
  const int HW = g.Ho * g.Wo;
  if (sizeof(T) == 2 && HW % 2 == 0 && Od % 8 == 0) {
    dim3 grid2(cdiv(HW, 64), cdiv(Od, 64), g.N);
    pack_nhwc_bf16_kernel<<<grid2, 256, 0, st>>>((const __nv_bfloat16*)gy, gyp, g.O, HW, Od); SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
    return SDB_OK;
  }
  dim3 grid(cdiv(HW, 32), cdiv(Od, 64), g.N);
  pack_nhwc_kernel<T><<<grid, 256, 0, st>>>((const T*)gy, gyp, g.O, HW, Od); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}


}  // namespace tcshared
}  // namespace sdb
