// dcn_tc_bwd.cu -- backward of the deformable convolution on tcgen05 tensor cores (SDB_MATH_BF16).
//
// backward_data has two halves, neither of which uses atomics on tensors:
//  (1) grad_offset / grad_mask -- one fused persistent kernel:
//   dcol[p, (tap,c)] = sum_o dY[p,o] W[o,c,tap]        tcgen05 GEMM, M=128 pixels, N=128 channels,
//                                                      K = C_out, accumulator in TMEM (never in HBM)
//   4 drain warps move each 128x128 fp32 accumulator to a swizzled bf16 staging tile in smem;
//   reduce warps (16 lanes x 8 channels per pixel) re-read the 4 input corners of (pixel, tap) and
//   reduce  sum_c dcol * d(bilinear)/d(y|x)  and  sum_c dcol * bilinear  over channels with warp
//   shuffles -> grad_offset / grad_mask, plain stores (one owner per (pixel, tap)).
//  (2) grad_input -- the reference scatters w * dcol with fp32 atomics (K3/K6); here the scatter is
//   inverted once per call into a CSR index "which (output pixel, tap, weight) touch input pixel q"
//   (count -> scan -> fill, ~36 eight-byte entries per output pixel) and grad_input becomes a second
//   gathered implicit GEMM  dX[q,c] = sum_{tap,o} (sum_e w_e dY[p_e,o]) W[o,c,tap]  run by the
//   forward kernel in MODE_DX (dcn_tc.cu).  Tensor work goes up by one GEMM, HBM/L2 atomic traffic
//   (36 KB per output pixel at C=256) goes to zero, and the fp32 NHWC accumulation buffer, its
//   memset and the NHWC->NCHW conversion pass disappear.
//   Replaces G2 + K2/K5 + K3/K6 of the reference (deform_conv_cuda.cu:553-559,
//   deform_conv_cuda_kernel.cu:291-452, :870-1066); `columns` is never written to HBM.
//
// backward_weight: dW[o, c, tap] = sum_p dY[p,o] col[p,(tap,c)]  -- tcgen05 GEMM with BOTH operands
//   MN-major: A = dY^T tiles (bulk-copied), B = re-gathered col tile (same producer as forward),
//   K = pixels.  One CTA per (tap, 128-channel chunk, pixel split); both 128-row halves of C_out
//   accumulate in TMEM for the CTA's whole pixel range; split partials are reduced (and permuted to
//   [O][C][kH][kW]) by a second tiny kernel -- deterministic, no atomics.
//   Replaces K1' + G3 of the reference (deform_conv_cuda.cu:738-778).
#include "dcn_tc_shared.cuh"

namespace sdb {
namespace {
using namespace tc;
using namespace tcshared;

// NCH (template parameter, 128 or 64) = channels per dgrad accumulator / wgrad N tile;
// LPB = NCH/8 lanes per pixel in the scatter / gather (8 channels per lane), PPI = 32/LPB pixels
// per warp instruction.
constexpr int NSW = 8;                   // scatter (dgrad) / gather (wgrad) warps
constexpr int FIRST_SW = 6;              // warps: 0 bulk producer, 1 mma, 2-5 drain/epilogue, 6.. SIMT
constexpr int BWD_THREADS = (FIRST_SW + NSW) * 32;
constexpr int PIX_PER_WARP = TILE_M / NSW;   // 16
constexpr int MAX_B_STAGES = 8;

inline int okb_of(const Geo& g) { return 2 * ((g.O + 127) / 128); }     // 64-wide o blocks, even count
inline int nch_of(const Geo& g) { return g.C % 128 == 0 ? 128 : 64; }
inline int nch_chunks(const Geo& g) { return g.C / nch_of(g); }

// ------------------------------------------------------------------------------------------------
// packing kernels
// ------------------------------------------------------------------------------------------------
// dY [N][O][HWo] (T) -> image of 128-pixel tiles: tile t, o-block kb -> [128 rows][64 o] bf16,
// 128B-swizzled (K-major for dgrad's A operand, MN-major for wgrad's A operand).  Zero padded.
template <typename T>
__global__ void __launch_bounds__(256) pack_gy_kernel(const T* __restrict__ gy, uint8_t* __restrict__ img,
                                                      const Geo g, int okb) {
  const int O = g.O, hw = g.Ho * g.Wo;
  __shared__ float s[64][129];
  const int tile = blockIdx.x, kb = blockIdx.y;
  const int tid = threadIdx.x;
  {
    const int px = tid & 127;
    const long long p = (long long)tile * TILE_M + px;
    const bool valid = p < g.P();
    int n = 0, ho = 0, wo = 0;
    if (valid) decode_q(g, p, n, ho, wo);
    const int rem = ho * g.Wo + wo;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int ol = (tid >> 7) + 2 * j, o = kb * 64 + ol;
      s[ol][px] = (valid && o < O) ? to_f32(gy[((size_t)n * O + o) * hw + rem]) : 0.f;
    }
  }
  __syncthreads();
  uint8_t* dst = img + ((size_t)tile * okb + kb) * (TILE_M * 128);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + 256 * j, row = c >> 3, ch = c & 7;
    uint4 pk;
    pk.x = pack_bf16x2(s[ch * 8 + 0][row], s[ch * 8 + 1][row]);
    pk.y = pack_bf16x2(s[ch * 8 + 2][row], s[ch * 8 + 3][row]);
    pk.z = pack_bf16x2(s[ch * 8 + 4][row], s[ch * 8 + 5][row]);
    pk.w = pack_bf16x2(s[ch * 8 + 6][row], s[ch * 8 + 7][row]);
    *reinterpret_cast<uint4*>(dst + sw128_offset(row, ch)) = pk;
  }
}

// bf16 source whose row segments are 8-byte aligned (Wo, tw, Ho*Wo multiples of 4): a lane loads FOUR pixels of
// one channel with one 8-byte load -- a whole 128-pixel tile row per warp instruction, 8 per lane instead of 32
// two-byte loads -- and one pixel decode per thread.
__global__ void __launch_bounds__(256) pack_gy_bf16v_kernel(const __nv_bfloat16* __restrict__ gy, uint8_t* __restrict__ img,
                                                            const Geo g, int okb) {
  const int O = g.O, hw = g.Ho * g.Wo;
  __shared__ __align__(8) __nv_bfloat16 s[64][136];   // [o][pixel (column ^ 32 for o >= 32)], 8 spare columns
  const int tile = blockIdx.x, kb = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const long long p0 = (long long)tile * TILE_M + 4 * lane;
    const bool valid = p0 < g.P();   // P % 4 == 0: all four pixels or none
    int n = 0, ho = 0, wo = 0;
    if (valid) decode_q(g, p0, n, ho, wo);
    const size_t r0 = (size_t)n * O * hw + (size_t)ho * g.Wo + wo;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ol = warp + 8 * j, o = kb * 64 + ol;
      uint2 v = make_uint2(0u, 0u);
      if (valid && o < O) v = *reinterpret_cast<const uint2*>(gy + (size_t)o * hw + r0);
      *reinterpret_cast<uint2*>(&s[ol][(4 * lane) ^ ((ol >> 5) << 5)]) = v;
    }
  }
  __syncthreads();
  uint8_t* dst = img + ((size_t)tile * okb + kb) * (TILE_M * 128);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + 256 * j, row = c >> 3, ch = c & 7;
    const int col = row ^ ((ch >> 2) << 5);
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (uint32_t)__bfloat16_as_ushort(s[ch * 8 + k][col]);
    uint4 pk;
    pk.x = v[0] | (v[1] << 16);
    pk.y = v[2] | (v[3] << 16);
    pk.z = v[4] | (v[5] << 16);
    pk.w = v[6] | (v[7] << 16);
    *reinterpret_cast<uint4*>(dst + sw128_offset(row, ch)) = pk;
  }
}

// W [O][C][taps] -> tiles ordered (tap, cchunk, okb): [NCH rows (c)][64 o] bf16 K-major swizzled.
template <typename T, int NCH>
__global__ void __launch_bounds__(256) prep_weight_dgrad_kernel(const T* __restrict__ w, uint8_t* __restrict__ img,
                                                                int O, int C, int taps, int okb) {
  const int o8n = okb * 8;  // 8-wide o chunks incl. zero padding
  const long long total = (long long)taps * C * o8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o8 = (int)(i % o8n);
    const int c = (int)((i / o8n) % C);
    const int tap = (int)(i / ((long long)o8n * C));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = o8 * 8 + j;
      v[j] = o < O ? to_f32(w[((size_t)o * C + c) * taps + tap]) : 0.f;
    }
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]);
    pk.y = pack_bf16x2(v[2], v[3]);
    pk.z = pack_bf16x2(v[4], v[5]);
    pk.w = pack_bf16x2(v[6], v[7]);
    const size_t tile = ((size_t)tap * (C / NCH) + c / NCH) * okb + (o8 >> 3);
    *reinterpret_cast<uint4*>(img + tile * ((size_t)NCH * 128) + sw128_offset(c % NCH, o8 & 7)) = pk;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const T* __restrict__ gy, float* __restrict__ gb,
                                                        float scale, int N, int O, int hw) {
  const int o = blockIdx.x;
  float sum = 0.f;
  for (int n = 0; n < N; ++n)
    for (int i = threadIdx.x; i < hw; i += blockDim.x) sum += to_f32(gy[((size_t)n * O + o) * hw + i]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(gb + o, scale * t);
  }
}

// ------------------------------------------------------------------------------------------------
// sampling descriptor for the backward pass
// ------------------------------------------------------------------------------------------------
struct BSample {
  int idx[4];   // corner pixel index (n*H + y)*W + x, or -1 when that corner is outside the image
  float lh, lw, m;
};

// raw (dy, dx, mask) of one (pixel, tap): fetched ahead of use so the loads overlap other work
struct RawB {
  float dy, dx, m;
};
__device__ __forceinline__ RawB fetch_rawb(const Geo& g, const float* __restrict__ off,
                                           const float* __restrict__ mask, bool valid, int n, int ho, int wo,
                                           int tap) {
  RawB r = {0.f, 0.f, 1.f};
  if (!valid) return r;
  const int hw = g.Ho * g.Wo, k2 = g.KH * g.KW;
  const float* o = off + ((size_t)n * 2 * k2 + 2 * tap) * hw + ho * g.Wo + wo;
  r.dy = __ldg(o);
  r.dx = __ldg(o + hw);
  if (mask) r.m = __ldg(mask + ((size_t)n * k2 + tap) * hw + ho * g.Wo + wo);
  return r;
}

__device__ __forceinline__ BSample make_bsample_raw(const Geo& g, const RawB raw, bool valid, int n, int ho,
                                                    int wo, int tap) {
  BSample s;
  s.idx[0] = s.idx[1] = s.idx[2] = s.idx[3] = -1;
  s.lh = s.lw = 0.f;
  s.m = 0.f;
  if (!valid) return s;
  const int i = tap / g.KW, j = tap - i * g.KW;
  const float h = (float)(ho * g.sh - g.ph + i * g.dh) + raw.dy;
  const float w = (float)(wo * g.sw - g.pw + j * g.dw) + raw.dx;
  // gradient-side validity test is the non-strict one (deform_conv_cuda_kernel.cu:140-144, :435-437)
  if (h <= -1.f || w <= -1.f || h >= (float)g.H || w >= (float)g.W) return s;
  s.m = raw.m;
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  s.lh = h - h_low;
  s.lw = w - w_low;
  const bool t = h_low >= 0, b = h_high <= g.H - 1, l = w_low >= 0, r = w_high <= g.W - 1;
  const int base = n * g.H;
  if (t && l) s.idx[0] = (base + h_low) * g.W + w_low;
  if (t && r) s.idx[1] = (base + h_low) * g.W + w_high;
  if (b && l) s.idx[2] = (base + h_high) * g.W + w_low;
  if (b && r) s.idx[3] = (base + h_high) * g.W + w_high;
  return s;
}

__device__ __forceinline__ BSample make_bsample(const Geo& g, const float* __restrict__ off,
                                                const float* __restrict__ mask, bool valid, int n, int ho,
                                                int wo, int tap) {
  return make_bsample_raw(g, fetch_rawb(g, off, mask, valid, n, ho, wo, tap), valid, n, ho, wo, tap);
}

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// ------------------------------------------------------------------------------------------------
// index of the transposed sampling pattern (grad_input as a gather, see file header)
// ------------------------------------------------------------------------------------------------
// key(q, tap) = (tile(q) * taps + tap) * 128 + row(q), q = band-order position of the INPUT pixel.
// The list of (output pixel p, weight w = bilinear x mask) that reach q through `tap` is stored as
//   desc[key]  : its first four entries in the forward gather's descriptor format (row offset of p in
//                the NHWC bf16 dY in 16 B units + bf16x2 weight; unused slots are zero), and
//   overflow   : entries five and up, four per ODesc (same fields + the row), descriptors of a key
//                contiguous (start[key] .. start[key+1]) and keys in (tile, tap, row) order, so the
//                descriptors of one (tile, tap, 16-row warp slice) are one contiguous run.
// With stride 1 a list holds four entries on average, so most of the work takes the fixed-width path.
constexpr int DESC_W = 4;
template <typename F>
__device__ __forceinline__ void for_each_hit(const Geo& g, const float* __restrict__ off,
                                             const float* __restrict__ mask, int n, int ho, int wo, int tap,
                                             F f) {
  const BSample s = make_bsample(g, off, mask, true, n, ho, wo, tap);
  const float wk[4] = {(1.f - s.lh) * (1.f - s.lw), (1.f - s.lh) * s.lw, s.lh * (1.f - s.lw), s.lh * s.lw};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (s.idx[k] < 0) continue;
    const float wv = wk[k] * s.m;
    const uint32_t wb = pack_bf16x2(wv, wv) >> 16;
    if (wb == 0u || wb == 0x8000u) continue;       // rounds to zero as a bf16 operand: contributes nothing
    const int pixel = s.idx[k] - n * g.H * g.W;    // y * W + x
    const int y = pixel / g.W, x = pixel - y * g.W;
    const long long q = encode_pos(g.H, g.W, g.th, g.tw, n, y, x);
    const long long key = ((q >> 7) * g.taps() + tap) * TILE_M + (q & 127);
    f(key, wb, (uint32_t)(q & 127));
  }
}

__global__ void __launch_bounds__(256) csr_count_kernel(const float* __restrict__ off, const float* __restrict__ mask,
                                                        int* __restrict__ cnt, const Geo g) {
  const int tap = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P()) return;
  const int hw = g.Ho * g.Wo;
  const int n = (int)(p / hw), r = (int)(p - (long long)n * hw);
  for_each_hit(g, off, mask, n, r / g.Wo, r % g.Wo, tap,
               [&](long long key, uint32_t, uint32_t) { atomicAdd(cnt + key, 1); });
}

constexpr int SCAN_PER_BLOCK = 2048;   // keys per 256-thread block
__device__ __forceinline__ int overflow_of(int c) { return c > DESC_W ? (c - DESC_W + 3) >> 2 : 0; }   // descriptors
__global__ void __launch_bounds__(256) csr_block_sums_kernel(const int* __restrict__ cnt, int* __restrict__ bsum,
                                                             int nkeys) {
  const int base = blockIdx.x * SCAN_PER_BLOCK;
  int s = 0;
  for (int i = threadIdx.x; i < SCAN_PER_BLOCK; i += 256) {
    const int k = base + i;
    if (k < nkeys) s += overflow_of(cnt[k]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  __shared__ int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    bsum[blockIdx.x] = t;
  }
}
// exclusive scan of the block sums in place (single block)
__global__ void __launch_bounds__(1024) csr_scan_top_kernel(int* __restrict__ bsum, int nblocks) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < nblocks ? bsum[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, d);
        if (threadIdx.x >= d) w += t;
      }
      wsum[threadIdx.x] = w;   // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int wbase = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    if (i < nblocks) bsum[i] = carry + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wbase + incl;
    __syncthreads();
  }
}
// start[k] = exclusive scan of the overflow descriptor counts; start[nkeys] = total; the descriptors of
// key k are cleared and tagged with their row here
__global__ void __launch_bounds__(256) csr_scan_final_kernel(const int* __restrict__ cnt, const int* __restrict__ bsum,
                                                             int* __restrict__ start, ODesc* __restrict__ odesc,
                                                             int nkeys) {
  __shared__ int wsum[8];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = bsum[blockIdx.x];
  __syncthreads();
  const int base = blockIdx.x * SCAN_PER_BLOCK;
  for (int i0 = 0; i0 < SCAN_PER_BLOCK; i0 += 256) {
    const int k = base + i0 + threadIdx.x;
    const int v = k < nkeys ? overflow_of(cnt[k]) : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
    const int carry = carry_s;
    const int excl = carry + wbase + incl - v;
    if (k < nkeys) {
      start[k] = excl;
      if (k == nkeys - 1) start[nkeys] = excl + v;
      for (int d = 0; d < v; ++d) {
        odesc[excl + d].o = make_uint4(0u, 0u, 0u, 0u);
        odesc[excl + d].m = make_uint4(0u, 0u, (uint32_t)(k & (TILE_M - 1)), 0u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + wbase + incl;
    __syncthreads();
  }
}

// row_units = 16-byte units per pixel row of the NHWC dY (okb * 8)
__global__ void __launch_bounds__(256) csr_fill_kernel(const float* __restrict__ off, const float* __restrict__ mask,
                                                       int* __restrict__ cnt, const int* __restrict__ start,
                                                       GDesc* __restrict__ desc, ODesc* __restrict__ odesc,
                                                       const Geo g, int row_units) {
  const int tap = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.P()) return;
  const int hw = g.Ho * g.Wo;
  const int n = (int)(p / hw), r = (int)(p - (long long)n * hw);
  const uint32_t poff = (uint32_t)p * (uint32_t)row_units;
  for_each_hit(g, off, mask, n, r / g.Wo, r % g.Wo, tap, [&](long long key, uint32_t wb, uint32_t row) {
    const int pos = atomicSub(cnt + key, 1) - 1;   // slots are handed out from the back
    if (pos < DESC_W) {
      desc[key].off[pos] = poff;
      desc[key].w2[pos] = (wb << 16) | wb;
    } else {
      ODesc* od = odesc + start[key] + ((pos - DESC_W) >> 2);
      const int sl = (pos - DESC_W) & 3;
      reinterpret_cast<uint32_t*>(&od->o)[sl] = poff;
      reinterpret_cast<uint16_t*>(&od->m)[sl] = (uint16_t)wb;
    }
  });
}

// byte offset of 16-byte chunk `chunk` of row `row` in the bf16 staging tile [128][NCH]
template <int NCH>
__device__ __forceinline__ uint32_t stg_offset(uint32_t row, uint32_t chunk) {
  return row * (NCH * 2) + (((chunk & ~7u) | ((chunk & 7u) ^ (row & 7u))) << 4);
}

// ------------------------------------------------------------------------------------------------
// backward data kernel
// ------------------------------------------------------------------------------------------------
struct DgradParams {
  const __nv_bfloat16* xp;   // NHWC bf16 input
  const float* off;
  const float* mask;
  const uint8_t* gy_img;     // dY tiles
  const uint8_t* wt_img;     // W^T tiles
  float* goff;               // [N][2*taps][HWo] fp32, pre-zeroed, or nullptr
  float* gmask;              // [N][taps][HWo] fp32, pre-zeroed, or nullptr
  Geo g;
  int num_tiles, nsb, okb;
};

template <int NCH>
__global__ void __launch_bounds__(BWD_THREADS, 1) dcn_bwd_data_tc_kernel(const DgradParams p) {
  constexpr int LPB = NCH / 8, PPI = 32 / LPB;
  constexpr uint32_t B_BYTES = NCH * 128;            // one [NCH c][64 o] weight tile
  constexpr uint32_t STG_BYTES = TILE_M * NCH * 2;   // bf16 staging tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, a_empty;
  __shared__ __align__(8) uint64_t b_full[MAX_B_STAGES], b_empty[MAX_B_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], stg_full[2], stg_empty[2];
  __shared__ uint32_t tmem_base_s;

  const Geo& g = p.g;
  const int C = g.C, taps = g.KH * g.KW, nch = C / NCH, units = taps * nch, okb = p.okb;
  const uint32_t A_BYTES = (uint32_t)okb * (TILE_M * 128);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sA = sm;
  uint8_t* sB = sA + A_BYTES;
  uint8_t* sS = sB + (size_t)p.nsb * B_BYTES;
  // warp index via shfl = provably warp-uniform: role branches and loop counters stay in uniform registers,
  // so tcgen05.mma takes its descriptors from the uniform datapath without an ELECT/R2UR waterfall per instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t NCOLS = 2 * NCH;

  if (threadIdx.x == 0) {
    mbar_init(&a_full, 1);
    mbar_init(&a_empty, 1);
    for (int s = 0; s < p.nsb; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);        // one arrival per drain warp
      mbar_init(&stg_full[s], 4);
      mbar_init(&stg_empty[s], NSW);      // one arrival per reduce warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, NCOLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // ===== bulk producer: dY tile once per tile, W^T tiles per (tap, chunk, o-block) =====
    if (lane == 0) {
      uint32_t bs = 0, bp = 0, ap = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&a_empty, ap ^ 1);
        mbar_arrive_expect_tx(&a_full, A_BYTES);
        bulk_g2s(sA, p.gy_img + (size_t)tile * A_BYTES, A_BYTES, &a_full);
        ap ^= 1;
        for (int i = 0; i < units * okb; ++i) {
          mbar_wait(&b_empty[bs], bp ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
          bulk_g2s(sB + (size_t)bs * B_BYTES, p.wt_img + (size_t)i * B_BYTES, B_BYTES, &b_full[bs]);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: dcol[128 x NCH] = dY_tile[128 x O] * W^T =====
    const uint32_t idesc = make_idesc_bf16(TILE_M, NCH, 0, 0);
    uint32_t bs = 0, bp = 0, acc = 0, accp = 0, ap = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&a_full, ap);
      ap ^= 1;
      for (int u = 0; u < units; ++u) {
        mbar_wait(&acc_empty[acc], accp ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * NCH;
        for (int kb = 0; kb < okb; ++kb) {
          mbar_wait(&b_full[bs], bp);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + kb * (TILE_M * 128);
            const uint32_t b_addr = smem_base + A_BYTES + bs * B_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32, 16, 1024),
                        make_smem_desc_sw128(b_addr + k4 * 32, 16, 1024), idesc, (kb | k4) != 0);
            umma_commit(&b_empty[bs]);
          }
          __syncwarp();
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
        if (elect_one()) umma_commit(&acc_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; accp ^= 1; }
      }
      if (elect_one()) umma_commit(&a_empty);
      __syncwarp();
    }
  } else if (warp < FIRST_SW) {
    // ===== drain: TMEM accumulator -> bf16 staging tile (row = pixel, swizzled 16-byte chunks) =====
    const int q = warp & 3;
    const uint32_t row = q * 32 + lane;
    uint32_t acc = 0, accp = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int u = 0; u < units; ++u) {
        mbar_wait(&acc_full[acc], accp);
        tc_fence_after_sync();
        mbar_wait(&stg_empty[acc], accp ^ 1);
        uint8_t* stg = sS + (size_t)acc * STG_BYTES;
#pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + acc * NCH + ((uint32_t)(q * 32) << 16) + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
            pk.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
            pk.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
            pk.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
            *reinterpret_cast<uint4*>(stg + stg_offset<NCH>(row, c0 / 8 + j)) = pk;
          }
        }
        tc_fence_before_sync();
        mbar_arrive_warp(&acc_empty[acc]);   // TMEM buffer may be overwritten
        mbar_arrive_warp(&stg_full[acc]);    // staging tile is ready (release)
        if (++acc == 2) { acc = 0; accp ^= 1; }
      }
    }
  } else {
    // ===== reduce warps: grad_offset / grad_mask = channel reductions of dcol against the corners =====
    // Same load pipeline as the forward gather: per (pixel, tap) a descriptor in shared memory (corner
    // offsets + lh, lw, mask, corner-valid flags, built one tap ahead by lanes 0..15), and a 4-slot
    // register ring that keeps 16 sixteen-byte corner loads per lane in flight across unit boundaries.
    // Per (pixel, tap, 8 channels): four packed-bf16 dot products S_k = <dcol, corner_k> (16 HFMA2),
    // combined in fp32 into this lane's share of d/dy, d/dx and d/dmask, then a 5-shuffle
    // reduce-scatter over the pixel's lane group.  The sums over channel chunks stay in registers;
    // one plain store per (pixel, tap, quantity) at the end of the tap.
    constexpr int ITERS = PIX_PER_WARP / PPI;
    constexpr int RING = 4;
    static_assert(ITERS % RING == 0, "ring must divide the per-unit iteration count");
    __shared__ uint4 s_od[NSW][2][PIX_PER_WARP][2];   // {off[4]}, {lh, lw, m, flags}
    __shared__ int2 s_px[NSW][PIX_PER_WARP];          // (n, ho*Wo+wo) of the warp's pixels, n = -1 when padded
    const int sw = warp - FIRST_SW, r0 = sw * PIX_PER_WARP;
    const int grp = lane / LPB, lig = lane % LPB;
    const int hw = g.Ho * g.Wo;
    const uint4* xbase = reinterpret_cast<const uint4*>(p.xp) + lig;
    uint32_t sb = 0, sp = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long pix = (long long)tile * TILE_M + r0 + lane;
      const bool valid = lane < PIX_PER_WARP && pix < g.P();
      int n = 0, ho = 0, wo = 0;
      if (valid) decode_q(g, pix, n, ho, wo);
      __syncwarp();   // previous tile's descriptors / pixel table no longer read
      if (lane < PIX_PER_WARP) s_px[sw][lane] = make_int2(valid ? n : -1, ho * g.Wo + wo);
      // descriptor of this lane's pixel for tap `tap_` -> buffer tap_ & 1
      auto build_desc = [&](int tap_, const RawB raw_) {
        if (lane < PIX_PER_WARP) {
          const BSample bs = make_bsample_raw(g, raw_, valid, n, ho, wo, tap_);
          uint4 o, f;
          uint32_t flags = 0;
          o.x = bs.idx[0] >= 0 ? (flags |= 1u, (uint32_t)bs.idx[0] * (uint32_t)(C / 8)) : 0u;
          o.y = bs.idx[1] >= 0 ? (flags |= 2u, (uint32_t)bs.idx[1] * (uint32_t)(C / 8)) : 0u;
          o.z = bs.idx[2] >= 0 ? (flags |= 4u, (uint32_t)bs.idx[2] * (uint32_t)(C / 8)) : 0u;
          o.w = bs.idx[3] >= 0 ? (flags |= 8u, (uint32_t)bs.idx[3] * (uint32_t)(C / 8)) : 0u;
          f.x = __float_as_uint(bs.lh); f.y = __float_as_uint(bs.lw); f.z = __float_as_uint(bs.m); f.w = flags;
          s_od[sw][tap_ & 1][lane][0] = o;
          s_od[sw][tap_ & 1][lane][1] = f;
        }
      };
      build_desc(0, fetch_rawb(g, p.off, p.mask, valid, n, ho, wo, 0));
      RawB raw_next = fetch_rawb(g, p.off, p.mask, valid && taps > 1, n, ho, wo, 1);   // raw offsets run two taps ahead
      __syncwarp();
      uint4 v[RING][4];
#define SDB_ISSUE(tap_, ch_, it_, slot_)                                                         \
      {                                                                                          \
        const uint4 o_ = s_od[sw][(tap_) & 1][(it_) * PPI + grp][0];                             \
        const uint4* xb_ = xbase + (ch_) * (NCH / 8);                                            \
        v[slot_][0] = __ldg(xb_ + o_.x);                                                         \
        v[slot_][1] = __ldg(xb_ + o_.y);                                                         \
        v[slot_][2] = __ldg(xb_ + o_.z);                                                         \
        v[slot_][3] = __ldg(xb_ + o_.w);                                                         \
      }
#pragma unroll
      for (int u = 0; u < RING; ++u) SDB_ISSUE(0, 0, u, u)
      for (int tap = 0; tap < taps; ++tap) {
        __syncwarp();
        if (tap + 1 < taps) build_desc(tap + 1, raw_next);   // buffer (tap+1)&1 was last read during tap-1
        raw_next = fetch_rawb(g, p.off, p.mask, valid && tap + 2 < taps, n, ho, wo, tap + 2);
        __syncwarp();
        float racc[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) racc[it] = 0.f;
        for (int ch = 0; ch < nch; ++ch) {
          int ntap = tap, nchk = ch + 1;
          if (nchk == nch) { nchk = 0; ++ntap; }
          const bool has_next = ntap < taps;
          mbar_wait(&stg_full[sb], sp);
          const uint8_t* stg = sS + (size_t)sb * STG_BYTES;
#pragma unroll
          for (int it = 0; it < ITERS; ++it) {
            const int slot = it % RING;
            const int px = it * PPI + grp;
            const uint4 f = s_od[sw][tap & 1][px][1];
            const uint4 d = *reinterpret_cast<const uint4*>(stg + stg_offset<NCH>(r0 + px, lig));
            float S[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t a2 = bf2_fma(d.w, v[slot][k].w, bf2_fma(d.z, v[slot][k].z,
                                  bf2_fma(d.y, v[slot][k].y, bf2_mul(d.x, v[slot][k].x))));
              const float sk = __uint_as_float(a2 << 16) + __uint_as_float(a2 & 0xffff0000u);
              S[k] = (f.w >> k) & 1u ? sk : 0.f;
            }
            if (it + RING < ITERS) {
              SDB_ISSUE(tap, ch, it + RING, slot)
            } else if (has_next) {
              SDB_ISSUE(ntap, nchk, it + RING - ITERS, slot)
            }
            const float lh = __uint_as_float(f.x), lw = __uint_as_float(f.y), m = __uint_as_float(f.z);
            // d(bilinear)/dh, /dw (get_coordinate_weight :185-211) and the unmasked sample value
            float qa = m * ((1.f - lw) * (S[2] - S[0]) + lw * (S[3] - S[1]));
            float qb = m * ((1.f - lh) * (S[1] - S[0]) + lh * (S[3] - S[2]));
            float qc = (1.f - lh) * ((1.f - lw) * S[0] + lw * S[1]) + lh * ((1.f - lw) * S[2] + lw * S[3]);
            // reduce-scatter over the LPB lanes of this pixel: lanes [0,Q) end with sum(qa), [Q,2Q) sum(qb),
            // [2Q,3Q) sum(qc)
            constexpr int H = LPB / 2, Q = LPB / 4;
            const bool up = (lig & H) != 0;
            const float r0_ = __shfl_xor_sync(0xffffffffu, up ? qa : qc, H);
            const float r1_ = __shfl_xor_sync(0xffffffffu, up ? qb : 0.f, H);
            const float k0 = (up ? qc : qa) + r0_;
            const float k1 = (up ? 0.f : qb) + r1_;
            const bool uq = (lig & Q) != 0;
            float kk = (uq ? k1 : k0) + __shfl_xor_sync(0xffffffffu, uq ? k0 : k1, Q);
#pragma unroll
            for (int dlt = Q / 2; dlt > 0; dlt >>= 1) kk += __shfl_xor_sync(0xffffffffu, kk, dlt);
            racc[it] += kk;
          }
          mbar_arrive_warp(&stg_empty[sb]);
          if (++sb == 2) { sb = 0; sp ^= 1; }
        }
        // lanes lig == 0, Q, 2Q hold d/dy, d/dx, d/dmask of pixel it*PPI+grp
        {
          constexpr int Q = LPB / 4;
          const int quant = lig / Q;
          if ((lig % Q) == 0 && quant < 3) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
              const int2 pi = s_px[sw][it * PPI + grp];
              if (pi.x >= 0) {
                if (quant < 2) {
                  if (p.goff) p.goff[((size_t)pi.x * 2 * taps + 2 * tap + quant) * hw + pi.y] = racc[it];
                } else if (p.gmask) {
                  p.gmask[((size_t)pi.x * taps + tap) * hw + pi.y] = racc[it];
                }
              }
            }
          }
        }
      }
#undef SDB_ISSUE
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, NCOLS);
}

// ------------------------------------------------------------------------------------------------
// backward weight kernel
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  const __nv_bfloat16* xp;
  const float* off;
  const float* mask;
  const uint8_t* gy_img;
  float* part;               // [splits][taps][O][C] fp32 partial sums
  Geo g;
  int num_tiles, tiles_per_split, splits, okb, nsg, nsy;
};

template <int NCH>
__global__ void __launch_bounds__(BWD_THREADS, 1) dcn_bwd_weight_tc_kernel(const WgradParams p) {
  constexpr int LPB = NCH / 8;
  constexpr uint32_t G_BYTES = TILE_M * NCH * 2;   // gathered col tile [NCH/64 blocks][128 px][64 c]
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t g_full[4], g_empty[4], y_full[4], y_empty[4], acc_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ GDesc sdesc[NSW][2][PIX_PER_WARP];   // per gather warp, double buffered over tiles

  const Geo& g = p.g;
  const int C = g.C, O = g.O, taps = g.KH * g.KW, nch = C / NCH, okb = p.okb, mh_n = okb / 2;
  const uint32_t Y_BYTES = (uint32_t)okb * (TILE_M * 128);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sG = sm;
  uint8_t* sY = sG + (size_t)p.nsg * G_BYTES;
  // warp index via shfl = provably warp-uniform: role branches and loop counters stay in uniform registers,
  // so tcgen05.mma takes its descriptors from the uniform datapath without an ELECT/R2UR waterfall per instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // work item of this CTA
  const int split = blockIdx.x % p.splits;
  const int ch = (blockIdx.x / p.splits) % nch;
  const int tap = blockIdx.x / (p.splits * nch);
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(p.num_tiles, t0 + p.tiles_per_split);
  uint32_t ncols = 32;
  while ((int)ncols < mh_n * NCH) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nsg; ++s) {
      mbar_init(&g_full[s], NSW);   // one arrival per gather warp
      mbar_init(&g_empty[s], 1);
    }
    for (int s = 0; s < p.nsy; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ys = 0, yp = 0;
      for (int tile = t0; tile < t1; ++tile) {
        mbar_wait(&y_empty[ys], yp ^ 1);
        mbar_arrive_expect_tx(&y_full[ys], Y_BYTES);
        bulk_g2s(sY + (size_t)ys * Y_BYTES, p.gy_img + (size_t)tile * Y_BYTES, Y_BYTES, &y_full[ys]);
        if (++ys == (uint32_t)p.nsy) { ys = 0; yp ^= 1; }
      }
    }
  } else if (warp == 1) {
    // dW^T-free formulation: D[o, c] += sum_pixels dY[pix, o] * col[pix, c]; both operands MN-major
    const uint32_t idesc = make_idesc_bf16(128, NCH, 1, 1);
    uint32_t gs = 0, gp = 0, ys = 0, yp = 0;
    uint32_t accumulate = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(&y_full[ys], yp);
      mbar_wait(&g_full[gs], gp);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t y_addr = smem_base + p.nsg * G_BYTES + ys * Y_BYTES;
        const uint32_t g_addr = smem_base + gs * G_BYTES;
        for (int s = 0; s < TILE_M / 16; ++s) {
          for (int mh = 0; mh < mh_n; ++mh)
            umma_bf16(tmem_base + mh * NCH,
                      make_smem_desc_sw128(y_addr + (2 * mh) * (TILE_M * 128) + s * 2048, TILE_M * 128, 1024),
                      make_smem_desc_sw128(g_addr + s * 2048, TILE_M * 128, 1024), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(&y_empty[ys]);
        umma_commit(&g_empty[gs]);
      }
      __syncwarp();
      if (++ys == (uint32_t)p.nsy) { ys = 0; yp ^= 1; }
      if (++gs == (uint32_t)p.nsg) { gs = 0; gp ^= 1; }
    }
    if (elect_one()) umma_commit(&acc_full);
    __syncwarp();
  } else if (warp < FIRST_SW) {
    // epilogue: partial dW tile -> workspace [split][tap][O][C]
    const int q = warp & 3;
    mbar_wait(&acc_full, 0);
    tc_fence_after_sync();
    for (int mh = 0; mh < mh_n; ++mh) {
      const int o = mh * 128 + q * 32 + lane;
#pragma unroll
      for (int c0 = 0; c0 < NCH; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + mh * NCH + ((uint32_t)(q * 32) << 16) + c0, r);
        tmem_ld_wait();
        if (o < O && t1 > t0) {
          float4* dst = reinterpret_cast<float4*>(p.part + (((size_t)split * taps + tap) * O + o) * C + ch * NCH + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
      }
    }
  } else {
    // gather producers (same sampling and the same continuous 4-slot load ring as the forward pass,
    // running across TILE boundaries here: a CTA handles one (tap, channel chunk) for many tiles).
    // Descriptors are double buffered per warp and built one tile ahead; the raw offsets they are built
    // from are fetched two tiles ahead, so neither latency is exposed.
    constexpr int PPI = 32 / LPB, ITERS = PIX_PER_WARP / PPI, RING = 4;
    static_assert(ITERS % RING == 0, "ring must divide the per-tile iteration count");
    const int sw = warp - FIRST_SW, r0 = sw * PIX_PER_WARP;
    const int grp = lane / LPB, lig = lane % LPB;
    uint32_t gs = 0, gp = 0;
    const uint4* x16 = reinterpret_cast<const uint4*>(p.xp + ch * NCH + lig * 8);
    auto locate = [&](int tile_, bool& valid_, int& n_, int& ho_, int& wo_) {
      const long long pix = (long long)tile_ * TILE_M + r0 + lane;
      valid_ = lane < PIX_PER_WARP && tile_ < t1 && pix < g.P();
      n_ = ho_ = wo_ = 0;
      if (valid_) decode_q(g, pix, n_, ho_, wo_);
    };
    // forward-style descriptor (weights already x mask, zero outside) of this lane's pixel -> sdesc[buf_]
    auto build_desc = [&](const RawB raw_, bool valid_, int n_, int ho_, int wo_, int buf_) {
      if (lane < PIX_PER_WARP) {
        const BSample bs = make_bsample_raw(g, raw_, valid_, n_, ho_, wo_, tap);
        const float wk[4] = {(1.f - bs.lh) * (1.f - bs.lw), (1.f - bs.lh) * bs.lw, bs.lh * (1.f - bs.lw), bs.lh * bs.lw};
        uint4 o, w;
        uint32_t* op = &o.x;
        uint32_t* wp = &w.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool on = bs.idx[k] >= 0;
          op[k] = on ? (uint32_t)bs.idx[k] * (uint32_t)(C / 8) : 0u;
          const float wv = on ? wk[k] * bs.m : 0.f;
          wp[k] = pack_bf16x2(wv, wv);
        }
        *reinterpret_cast<uint4*>(sdesc[sw][buf_][lane].off) = o;
        *reinterpret_cast<uint4*>(sdesc[sw][buf_][lane].w2) = w;
      }
    };
    bool valid;
    int n, ho, wo;
    locate(t0, valid, n, ho, wo);
    build_desc(fetch_rawb(g, p.off, p.mask, valid, n, ho, wo, tap), valid, n, ho, wo, t0 & 1);
    locate(t0 + 1, valid, n, ho, wo);
    RawB raw = fetch_rawb(g, p.off, p.mask, valid, n, ho, wo, tap);   // of tile t0 + 1
    __syncwarp();
    uint4 v[RING][4], wq[RING];
#define SDB_WISSUE(tile_, it_, slot_)                                                            \
    {                                                                                            \
      const GDesc* d_ = &sdesc[sw][(tile_) & 1][(it_) * PPI + grp];                              \
      const uint4 o_ = *reinterpret_cast<const uint4*>(d_->off);                                 \
      wq[slot_] = *reinterpret_cast<const uint4*>(d_->w2);                                       \
      v[slot_][0] = __ldg(x16 + o_.x);                                                           \
      v[slot_][1] = __ldg(x16 + o_.y);                                                           \
      v[slot_][2] = __ldg(x16 + o_.z);                                                           \
      v[slot_][3] = __ldg(x16 + o_.w);                                                           \
    }
    if (t0 < t1) {
#pragma unroll
      for (int u = 0; u < RING; ++u) SDB_WISSUE(t0, u, u)
    }
    for (int tile = t0; tile < t1; ++tile) {
      const bool has_next = tile + 1 < t1;
      // descriptors of tile+1 (its raw offsets arrived during the previous tile), then raw offsets of tile+2
      __syncwarp();   // buffer (tile+1)&1 was read while tile-1 was gathered
      build_desc(raw, valid, n, ho, wo, (tile + 1) & 1);
      locate(tile + 2, valid, n, ho, wo);
      raw = fetch_rawb(g, p.off, p.mask, valid, n, ho, wo, tap);
      __syncwarp();
      mbar_wait(&g_empty[gs], gp ^ 1);
      uint8_t* dst = sG + (size_t)gs * G_BYTES + (lig >> 3) * (TILE_M * 128);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int slot = it % RING;
        uint4 a;
        a.x = bf2_fma(wq[slot].w, v[slot][3].x, bf2_fma(wq[slot].z, v[slot][2].x, bf2_fma(wq[slot].y, v[slot][1].x, bf2_mul(wq[slot].x, v[slot][0].x))));
        a.y = bf2_fma(wq[slot].w, v[slot][3].y, bf2_fma(wq[slot].z, v[slot][2].y, bf2_fma(wq[slot].y, v[slot][1].y, bf2_mul(wq[slot].x, v[slot][0].y))));
        a.z = bf2_fma(wq[slot].w, v[slot][3].z, bf2_fma(wq[slot].z, v[slot][2].z, bf2_fma(wq[slot].y, v[slot][1].z, bf2_mul(wq[slot].x, v[slot][0].z))));
        a.w = bf2_fma(wq[slot].w, v[slot][3].w, bf2_fma(wq[slot].z, v[slot][2].w, bf2_fma(wq[slot].y, v[slot][1].w, bf2_mul(wq[slot].x, v[slot][0].w))));
        *reinterpret_cast<uint4*>(dst + sw128_offset(r0 + it * PPI + grp, lig & 7)) = a;
        if (it + RING < ITERS) {
          SDB_WISSUE(tile, it + RING, slot)
        } else if (has_next) {
          SDB_WISSUE(tile + 1, it + RING - ITERS, slot)
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(&g_full[gs]);
      if (++gs == (uint32_t)p.nsg) { gs = 0; gp ^= 1; }
    }
#undef SDB_WISSUE
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// grad_weight[o][c][tap] += scale * sum_split part[split][tap][o][c]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw,
                                                           float scale, int splits, int taps, int O, int C) {
  const long long total = (long long)O * C * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int o = (int)((i / C) % O);
    const int tap = (int)(i / ((long long)C * O));
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += part[(((size_t)sp * taps + tap) * O + o) * C + c];
    gw[((size_t)o * C + c) * taps + tap] += scale * s;
  }
}

// ---- workspace layouts ----------------------------------------------------------------------
struct BwdWs {
  size_t xp_off, gy_off, wt_off, part_off, cnt_off, start_off, bsum_off, ent_off, wdx_off, desc_off, gyn_off, total;
  long long nkeys, max_entries;
  int scan_blocks;
};
int wgrad_splits(const Geo& g, int num_tiles) {
  int s = num_sms() / (g.taps() * nch_chunks(g));
  if (s < 1) s = 1;
  if (s > num_tiles) s = num_tiles;
  if (s > 64) s = 64;
  return s;
}
BwdWs bwd_ws(int op, const Geo& g) {
  BwdWs w{};
  const int okb = okb_of(g);
  const size_t tiles = (size_t)cdiv(g.P(), TILE_M);
  size_t o = 0;
  w.xp_off = o; o = align_up(o + (size_t)g.N * g.H * g.W * g.C * 2, 1024);
  w.gy_off = o; o = align_up(o + tiles * okb * (TILE_M * 128), 1024);
  if (op == SDB_OP_BACKWARD_DATA) {
    w.wt_off = o; o = align_up(o + (size_t)g.taps() * g.C * okb * 64 * 2, 1024);
    const long long tiles_in = cdiv((long long)g.N * g.H * g.W, TILE_M);
    w.nkeys = tiles_in * g.taps() * TILE_M;
    w.max_entries = g.P() * g.taps() + 1;   // overflow descriptors: a key with c > 4 entries needs ceil((c-4)/4) <= c/4
    w.scan_blocks = cdiv(w.nkeys, SCAN_PER_BLOCK);
    w.cnt_off = o;   o = align_up(o + (size_t)w.nkeys * 4, 1024);
    w.start_off = o; o = align_up(o + (size_t)(w.nkeys + 1) * 4, 1024);
    w.bsum_off = o;  o = align_up(o + (size_t)w.scan_blocks * 4, 1024);
    w.ent_off = o;   o = align_up(o + (size_t)w.max_entries * sizeof(ODesc), 1024);
    w.wdx_off = o;   o = align_up(o + (size_t)g.taps() * g.C * okb * 64 * 2, 1024);
    w.desc_off = o;  o = align_up(o + (size_t)w.nkeys * sizeof(GDesc), 1024);
    w.gyn_off = o;   o = align_up(o + (size_t)g.P() * okb * 64 * 2, 1024);
  } else {
    w.part_off = o; o = align_up(o + (size_t)wgrad_splits(g, (int)tiles) * g.taps() * g.O * g.C * 4, 1024);
  }
  w.total = o;
  return w;
}

template <typename T>
int pack_gy(const void* gy, uint8_t* img, const Geo& g, cudaStream_t st) {
  dim3 grid(cdiv(g.P(), TILE_M), okb_of(g));
  const int hw = g.Ho * g.Wo;
  if (sizeof(T) == 2 && g.Wo % 4 == 0 && g.tw % 4 == 0 && hw % 4 == 0 && ((size_t)gy & 7) == 0)
    pack_gy_bf16v_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)gy, img, g, okb_of(g));
  else
    pack_gy_kernel<T><<<grid, 256, 0, st>>>((const T*)gy, img, g, okb_of(g));
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // namespace

size_t tc_bwd_workspace_bytes(int op, const Geo& g) { return bwd_ws(op, g).total; }

int tc_backward_data(const void* x, const float* off, const float* mask, const void* w, const void* gy,
                     void* gx, float* goff, float* gmask, const Geo& g, int io_dtype, void* ws,
                     size_t ws_bytes, const void* x_packed, cudaStream_t st) {
  const int NCH = nch_of(g);
  const BwdWs L = bwd_ws(SDB_OP_BACKWARD_DATA, g);
  SDB_REQUIRE(ws && ws_bytes >= L.total, SDB_ERR_WORKSPACE, "backward_data workspace too small: %zu < %zu", ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  const bool f32 = io_dtype == SDB_F32;
  if (!mask) gmask = nullptr;
  int rc;
  uint8_t* gy_img = base + L.gy_off;
  rc = f32 ? pack_gy<float>(gy, gy_img, g, st) : pack_gy<__nv_bfloat16>(gy, gy_img, g, st);
  if (rc) return rc;
  const int okb = okb_of(g);

  if (goff || gmask) {
    // ---- (1) grad_offset / grad_mask: dcol GEMM + channel reduction ----
    const __nv_bfloat16* xp = (const __nv_bfloat16*)x_packed;
    if (!xp) {
      rc = f32 ? pack_input<float>(x, (__nv_bfloat16*)(base + L.xp_off), g, st)
               : pack_input<__nv_bfloat16>(x, (__nv_bfloat16*)(base + L.xp_off), g, st);
      if (rc) return rc;
      xp = (const __nv_bfloat16*)(base + L.xp_off);
    }
    uint8_t* wt_img = base + L.wt_off;
    const long long total = (long long)g.taps() * g.C * okb * 8;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    if (NCH == 128) {
      if (f32) prep_weight_dgrad_kernel<float, 128><<<blocks, 256, 0, st>>>((const float*)w, wt_img, g.O, g.C, g.taps(), okb);
      else prep_weight_dgrad_kernel<__nv_bfloat16, 128><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)w, wt_img, g.O, g.C, g.taps(), okb);
    } else {
      if (f32) prep_weight_dgrad_kernel<float, 64><<<blocks, 256, 0, st>>>((const float*)w, wt_img, g.O, g.C, g.taps(), okb);
      else prep_weight_dgrad_kernel<__nv_bfloat16, 64><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)w, wt_img, g.O, g.C, g.taps(), okb);
    }
    SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());

    DgradParams p;
    p.xp = xp; p.off = off; p.mask = mask; p.gy_img = gy_img; p.wt_img = wt_img;
    p.goff = goff; p.gmask = gmask; p.g = g;
    p.num_tiles = cdiv(g.P(), TILE_M);
    p.okb = okb;
    const size_t a_bytes = (size_t)okb * TILE_M * 128, b_bytes = NCH * 128, stg = 2 * (size_t)TILE_M * NCH * 2;
    long long nsb = ((long long)(208 * 1024) - 1024 - (long long)a_bytes - (long long)stg) / (long long)b_bytes;
    if (nsb > MAX_B_STAGES) nsb = MAX_B_STAGES;
    SDB_REQUIRE(nsb >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for backward_data");
    p.nsb = (int)nsb;
    const size_t smem = a_bytes + p.nsb * b_bytes + stg + 1024;
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    if (NCH == 128) SDB_ENSURE_SMEM(dcn_bwd_data_tc_kernel<128>, smem);
    else SDB_ENSURE_SMEM(dcn_bwd_data_tc_kernel<64>, smem);
    {
      ProfScope prof(SDB_OP_BACKWARD_DATA, st);
      if (NCH == 128) dcn_bwd_data_tc_kernel<128><<<grid, BWD_THREADS, smem, st>>>(p);
      else dcn_bwd_data_tc_kernel<64><<<grid, BWD_THREADS, smem, st>>>(p);
      SDB_LAUNCHED(1);
    }
    SDB_CHECK_CUDA(cudaGetLastError());
  }

  if (gx) {
    // ---- (2) grad_input: transposed sampling index + gathered implicit GEMM ----
    int* cnt = (int*)(base + L.cnt_off);
    int* start = (int*)(base + L.start_off);
    int* bsum = (int*)(base + L.bsum_off);
    ODesc* odesc = (ODesc*)(base + L.ent_off);
    GDesc* desc = (GDesc*)(base + L.desc_off);
    __nv_bfloat16* gyn = (__nv_bfloat16*)(base + L.gyn_off);
    const int nkeys = (int)L.nkeys;
    rc = f32 ? pack_grad_nhwc<float>(gy, gyn, g, okb * 64, st) : pack_grad_nhwc<__nv_bfloat16>(gy, gyn, g, okb * 64, st);
    if (rc) return rc;
    SDB_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nkeys * 4, st));
    SDB_CHECK_CUDA(cudaMemsetAsync(desc, 0, (size_t)nkeys * sizeof(GDesc), st));
    dim3 hgrid(cdiv(g.P(), 256), g.taps());
    csr_count_kernel<<<hgrid, 256, 0, st>>>(off, mask, cnt, g);
    csr_block_sums_kernel<<<L.scan_blocks, 256, 0, st>>>(cnt, bsum, nkeys);
    csr_scan_top_kernel<<<1, 1024, 0, st>>>(bsum, L.scan_blocks);
    csr_scan_final_kernel<<<L.scan_blocks, 256, 0, st>>>(cnt, bsum, start, odesc, nkeys);
    csr_fill_kernel<<<hgrid, 256, 0, st>>>(off, mask, cnt, start, desc, odesc, g, okb * 8);
    SDB_LAUNCHED(5);
    SDB_CHECK_CUDA(cudaGetLastError());
    rc = tc_dx(w, gyn, desc, start, odesc, base + L.wdx_off, gx, g, okb, io_dtype, st);
    if (rc) return rc;
  }
  return SDB_OK;
}

int tc_backward_weight(const void* x, const float* off, const float* mask, const void* gy, float* gw,
                       float* gb, float scale, const Geo& g, int io_dtype, void* ws, size_t ws_bytes,
                       const void* x_packed, cudaStream_t st) {
  const int NCH = nch_of(g);
  const BwdWs L = bwd_ws(SDB_OP_BACKWARD_WEIGHT, g);
  SDB_REQUIRE(ws && ws_bytes >= L.total, SDB_ERR_WORKSPACE, "backward_weight workspace too small: %zu < %zu", ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  const bool f32 = io_dtype == SDB_F32;
  int rc;
  if (gw) {
    const __nv_bfloat16* xp = (const __nv_bfloat16*)x_packed;
    if (!xp) {
      rc = f32 ? pack_input<float>(x, (__nv_bfloat16*)(base + L.xp_off), g, st)
               : pack_input<__nv_bfloat16>(x, (__nv_bfloat16*)(base + L.xp_off), g, st);
      if (rc) return rc;
      xp = (const __nv_bfloat16*)(base + L.xp_off);
    }
    uint8_t* gy_img = base + L.gy_off;
    rc = f32 ? pack_gy<float>(gy, gy_img, g, st) : pack_gy<__nv_bfloat16>(gy, gy_img, g, st);
    if (rc) return rc;
    WgradParams p;
    p.xp = xp; p.off = off; p.mask = mask; p.gy_img = gy_img; p.part = (float*)(base + L.part_off); p.g = g;
    p.num_tiles = cdiv(g.P(), TILE_M);
    p.tiles_per_split = cdiv(p.num_tiles, wgrad_splits(g, p.num_tiles));
    p.splits = cdiv(p.num_tiles, p.tiles_per_split);   // no empty split
    p.okb = okb_of(g);
    const size_t g_bytes = (size_t)TILE_M * NCH * 2, y_bytes = (size_t)p.okb * TILE_M * 128;
    p.nsy = 2;
    long long nsg = ((long long)(208 * 1024) - 1024 - (long long)(p.nsy * y_bytes)) / (long long)g_bytes;
    if (nsg > 4) nsg = 4;
    SDB_REQUIRE(nsg >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for backward_weight");
    p.nsg = (int)nsg;
    const size_t smem = p.nsg * g_bytes + p.nsy * y_bytes + 1024;
    const int grid = g.taps() * nch_chunks(g) * p.splits;
    if (NCH == 128) SDB_ENSURE_SMEM(dcn_bwd_weight_tc_kernel<128>, smem);
    else SDB_ENSURE_SMEM(dcn_bwd_weight_tc_kernel<64>, smem);
    {
      ProfScope prof(SDB_OP_BACKWARD_WEIGHT, st);
      if (NCH == 128) dcn_bwd_weight_tc_kernel<128><<<grid, BWD_THREADS, smem, st>>>(p);
      else dcn_bwd_weight_tc_kernel<64><<<grid, BWD_THREADS, smem, st>>>(p);
      SDB_LAUNCHED(1);
    }
    SDB_CHECK_CUDA(cudaGetLastError());
    const long long total = (long long)g.O * g.C * g.taps();
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(p.part, gw, scale, p.splits, g.taps(), g.O, g.C); SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
  }
  if (gb) {
    if (f32) bias_grad_kernel<float><<<g.O, 256, 0, st>>>((const float*)gy, gb, scale, g.N, g.O, g.HWo());
    else bias_grad_kernel<__nv_bfloat16><<<g.O, 256, 0, st>>>((const __nv_bfloat16*)gy, gb, scale, g.N, g.O, g.HWo());
    SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
  }
  return SDB_OK;
}

}  // namespace sdb
