// dcn_tc_shared.cuh -- helpers shared by the tensor-core DCN translation units (dcn_tc.cu: forward,
// dcn_tc_bwd.cu: backward data / backward weight).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace sdb {

// operand images of one prepared weight tensor (dcn_tc.cu: prep_weights_kernel)
struct TcWeightImages {
  const uint8_t* fwd;     // forward B operand
  const uint8_t* dgrad;   // W^T tiles of the dcol GEMM (grad_offset / grad_mask / grad_input)
  const uint8_t* convt;   // forward image of the transposed, tap-reversed weights (plain-convolution grad_input)
  const float* bias;      // fp32 [O] or nullptr
};
// host-side description of one problem of a multi-problem call, pointers resolved into the workspace
// spatial geometry of one problem (the rest of Geo is common to the launch)
struct Dims {
  int N, H, W, Ho, Wo;
};
__host__ __device__ __forceinline__ Geo with_dims(Geo g, const Dims& d) {
  g.N = d.N; g.H = d.H; g.W = d.W; g.Ho = d.Ho; g.Wo = d.Wo;
  return g;
}
struct TcProblem {
  Dims d;
  int weight_id, group;         // convolution this problem belongs to; offset group (shared transposed index)
  const void* x;                // NCHW input (io dtype)
  const float* off;
  const float* mask;
  void* out;
  const void* gy;               // NCHW grad_out (io dtype)
  void* gx;                     // NCHW grad_input or nullptr
  float* goff;
  float* gmask;
  TcWeightImages w;
  void* xp;                     // NHWC bf16 input
  uint8_t* gy_img;              // dY as swizzled 128-pixel tiles
  uint8_t* col;                 // sampled columns saved by the forward (A stages [tile][chunk][tap]), or nullptr
  uint8_t* dcol;                // dcol = dY W^T as bf16 staging tiles [tile][tap][chunk] (written by the grad_offset kernel)
  const int* start;             // transposed index of this problem's group: CSR offsets (per input pixel and tap)
  const void* ent;              // and the entry pool
};
// workspace layout of a multi-problem call (dcn_tc_bwd.cu: tc_plan)
struct TcPlan {
  size_t xp_off[16], gy_off[16], dcol_off[16];
  size_t prep_off[4], part_off[4];
  int conv;                      // plain-convolution call (every problem has offset == nullptr)
  size_t gyp_off[16];            // conv backward: NHWC bf16 copy of grad_out (input of the transposed convolution)
  size_t convw_off[4];           // conv backward: weight image 2 of every weight tensor
  int group_of[16], group_rep[16], ngroups;
  long long key_base[16], nkeys;
  int scan_blocks;
  size_t cnt_off, ent_off, clear_bytes, start_off, bsum_off, blk_off;
  long long ent_cap;             // entries the pool holds (multiple of LIST_ALIGN)
  int tiles_per_split[4], splits[4];           // weight-gradient split-K of the re-sampling kernel
  int tiles_per_split_col[4], splits_col[4];   // ... of the kernel over saved columns
  size_t total;
};
TcPlan tc_plan(const TcProblem* pb, int n, int nweights, const bool* have_prepared, const Geo& g, bool backward);
int tc_forward_multi(const TcProblem* pb, int n, const Geo& g, int io_dtype, cudaStream_t st);
int tc_dx_multi(const TcProblem* pb, int n, const Geo& g, int io_dtype, int accumulate, cudaStream_t st);
int tc_build_transposed_index(TcProblem* pb, int n, const TcPlan& P, const Geo& g, uint8_t* base, cudaStream_t st,
                              bool pointers_only = false);
int tc_forward_all(TcProblem* pb, int n, const Geo& g, int io_dtype, cudaStream_t st);
// gather_phase: 0 = everything; 1 = everything but the grad_input gather (dcol tiles and the transposed index stay in the
// workspace); 2 = only the gather, from the workspace a phase-1 call on the same table left behind
int tc_backward_all(TcProblem* pb, int n, float* const* gw, float* const* gb, int nweights, const TcPlan& P, const Geo& g,
                    int io_dtype, float scale, bool pack_x, int accumulate_gx, bool grad_packed, uint8_t* base,
                    cudaStream_t st, int gather_phase = 0);
// plain convolution (offset == nullptr): grad_weight over the saved columns, grad_input = conv(dY, W') by the forward kernel
int tc_conv_backward_all(TcProblem* pb, int n, const void* const* weights, float* const* gw, float* const* gb, int nweights,
                         const TcPlan& P, const Geo& g, int io_dtype, float scale, bool pack_x, uint8_t* base,
                         cudaStream_t st);
bool tc_conv_supported(const Geo& g, bool need_grad_input, const char** why);
size_t tc_prepared_weight_bytes(const Geo& g);
TcWeightImages tc_weight_images(const Geo& g, const void* prepared, bool has_bias);
int tc_prepare_weights(const void* w, const void* bias, const Geo& g, int io_dtype, void* prepared, int which,
                       cudaStream_t st);   // which: 1 = forward image, 2 = dcol image (3 = both)
int tc_lanes_per_pixel(const Geo& g);
// in-call weight preparation on the side stream (dcn_tc_bwd.cu)
cudaStream_t tc_prep_begin(cudaStream_t st);
void tc_prep_end(cudaStream_t st, cudaStream_t used);
int tc_prep_wait(cudaStream_t st);

// tcgen05 kind::tf32 forward with float32 tensors (dcn_tf32.cu); passes = 1 (tf32) or 3 (error-compensated, fp32-accurate)
bool tf32_supported(const Geo& g, const char** why);
size_t tf32_prepared_weight_bytes(const Geo& g, int passes);
int tf32_prepare_weights(const float* w, const float* bias, const Geo& g, int passes, void* prepared, cudaStream_t st);
size_t tf32_forward_workspace_bytes(const TcProblem* pb, int n, int nweights, const bool* have_prepared, const Geo& g,
                                    int passes);
int tf32_forward_all(const TcProblem* pb, int n, const void* const* weights, const void* const* biases,
                     const void* const* prepared, int nweights, const Geo& g, int passes, uint8_t* ws, cudaStream_t st);

namespace tcshared {
using namespace tc;

constexpr int TILE_M = 128;

// ------------------------------------------------------------------------------------------------
// Multi-problem launches.  One launch of every tensor-core kernel covers ALL problems of a call
// (the FPN levels x the convolutions of a head: RepPoints = 5 x 2), which share the channel counts
// and the kernel geometry and differ in N, H, W and their tensors.  The problem table travels as a
// kernel parameter (no device allocation, safe under CUDA-graph capture); a work item (128-pixel
// tile) is mapped to (problem, tile) through the prefix array below -- at most MAX_PROBS entries.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_PROBS = 16;     // problems per launch
constexpr int MAX_WEIGHTS = 4;    // distinct weight tensors (convolutions) per launch
struct TileMap {
  int n;                      // ranges in use
  int start[MAX_PROBS + 1];   // first work item of range i; start[n] = total
};
__device__ __forceinline__ int find_range(const TileMap& m, int work) {
  int i = 0;
  while (i + 1 < m.n && work >= m.start[i + 1]) ++i;
  return i;
}

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

// NCHW -> NHWC bf16 for every problem of a call in one launch: grid.x walks (entry, image, pixel block) through
// the TileMap (entry i owns N_i * nblk_i blocks), grid.y the 64-channel blocks.
struct PackTable {
  TileMap map;
  struct E { const void* src; __nv_bfloat16* dst; int HW, nblk, fast; } e[MAX_PROBS];
  int C, Cd;
};
// [N][C][HW] (T) -> [N][HW][Cd] bf16, Cd >= C (channels C..Cd-1 zero).  Tile 64 channels x 32 pixels.
template <typename T>
__device__ __forceinline__ void pack_nhwc_body(const PackTable& t, int ei, int local, float (*s)[33]) {
  const int HW = t.e[ei].HW, C = t.C, Cd = t.Cd;
  const int n = local / t.e[ei].nblk, c0 = blockIdx.y * 64, p0 = (local % t.e[ei].nblk) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* sp = (const T*)t.e[ei].src + ((size_t)n * C + c0) * HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ty + 8 * j;
    s[c][tx] = (c0 + c < C && p0 + tx < HW) ? to_f32(sp[(size_t)c * HW + p0 + tx]) : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* dp = t.e[ei].dst + ((size_t)n * HW + p0) * Cd + c0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = ty + 8 * j;
    if (p0 + p < HW && c0 + 2 * tx < Cd)
      *reinterpret_cast<__nv_bfloat162*>(dp + (size_t)p * Cd + 2 * tx) =
          __floats2bfloat162_rn(s[2 * tx][p], s[2 * tx + 1][p]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) pack_nhwc_kernel(const __grid_constant__ PackTable t) {
  __shared__ float s[64][33];
  const int ei = find_range(t.map, blockIdx.x);
  pack_nhwc_body<T>(t, ei, blockIdx.x - t.map.start[ei], s);
}

// bf16 source with an even pixel count per plane: 64 channels x 64 pixels per block, bf16x2 loads (128 B per
// warp instruction), one 16-byte store (8 channels of one pixel) per thread and round -- half the load and a
// quarter of the store instructions of the generic kernel.  Requires HW % 2 == 0 and Cd % 8 == 0; entries that do not
// qualify (`fast` == 0: odd plane sizes, the small pyramid levels) take the generic body in the same launch.
static __global__ void __launch_bounds__(256) pack_nhwc_bf16_kernel(const __grid_constant__ PackTable t) {
  __shared__ uint32_t s[64][33];   // [channel][pixel pair]
  const int ei = find_range(t.map, blockIdx.x);
  const int local = blockIdx.x - t.map.start[ei];
  if (!t.e[ei].fast) {
    pack_nhwc_body<__nv_bfloat16>(t, ei, local, reinterpret_cast<float(*)[33]>(s));
    return;
  }
  const int HW = t.e[ei].HW, C = t.C, Cd = t.Cd;
  const int n = local / t.e[ei].nblk, c0 = blockIdx.y * 64, p0 = (local % t.e[ei].nblk) * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const __nv_bfloat16* sp = (const __nv_bfloat16*)t.e[ei].src + ((size_t)n * C + c0) * HW;
  __nv_bfloat16* dst = t.e[ei].dst;
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

// dcol tiles: bf16 [128 pixels][NCH channels], every row followed by 16 bytes of padding.  Written by the grad_offset
// kernel's drain warps (one row per lane), read by its reduce warps (one row per lane group) and -- through HBM,
// exported verbatim by one bulk copy per tile -- by the grad_input gather.  The padding shifts consecutive rows by one
// 16-byte bank group, which makes the row-per-lane writes conflict-free without an XOR swizzle, so a gather lane's
// address is simply (row base + lane): one multiply-add per list entry.
__host__ __device__ constexpr uint32_t stg_row_units(int nch) { return (uint32_t)nch / 8u + 1u; }   // 16-byte units per row
__host__ __device__ constexpr uint32_t stg_tile_bytes(int nch) { return 128u * 16u * stg_row_units(nch); }
template <int NCH>
__host__ __device__ __forceinline__ uint32_t stg_offset(uint32_t row, uint32_t chunk) {
  return (row * stg_row_units(NCH) + chunk) * 16u;
}

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

// NCHW (io dtype) -> NHWC bf16 [N][HW][Cd] for `n` tensors in one launch
struct PackJob { const void* src; void* dst; int N, HW; };
inline int pack_nhwc_multi(const PackJob* jobs, int n, int C, int Cd, bool src_bf16, cudaStream_t st) {
  PackTable t{};
  t.C = C; t.Cd = Cd;
  int m = 0, total = 0;
  for (int i = 0; i < n; ++i) {
    if (!jobs[i].src || jobs[i].N * jobs[i].HW == 0) continue;
    const bool fast = src_bf16 && jobs[i].HW % 2 == 0 && Cd % 8 == 0;
    const int nblk = (jobs[i].HW + (fast ? 63 : 31)) / (fast ? 64 : 32);
    t.e[m].src = jobs[i].src; t.e[m].dst = (__nv_bfloat16*)jobs[i].dst; t.e[m].HW = jobs[i].HW; t.e[m].nblk = nblk;
    t.e[m].fast = fast;
    t.map.start[m] = total;
    total += jobs[i].N * nblk;
    ++m;
  }
  t.map.n = m; t.map.start[m] = total;
  if (total == 0) return SDB_OK;
  dim3 grid(total, (Cd + 63) / 64);
  if (src_bf16) pack_nhwc_bf16_kernel<<<grid, 256, 0, st>>>(t);
  else pack_nhwc_kernel<float><<<grid, 256, 0, st>>>(t);
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

// channel chunking of the backward kernels: NCH = channels per dcol accumulator / staging tile (128 or 64)
inline int okb_of(const Geo& g) { return 2 * ((g.O + 127) / 128); }     // 64-wide o blocks, even count
__host__ __device__ inline int nch_of(const Geo& g) { return g.C % 128 == 0 ? 128 : 64; }
inline int nch_chunks(const Geo& g) { return g.C / nch_of(g); }

// entry of the transposed sampling index while it is built (dcn_tc_dx.cu)
struct __align__(8) CEntry {
  uint32_t row16;   // dcol row in 16-byte units into the problem's dcol tiles, channel chunk 0:
                    // ((pos >> 7) * taps * nch + tap * nch) * (tile bytes / 16) + (pos & 127) * (row bytes / 16), pos = band-order position of p
  uint32_t tw;      // tap << 16 | weight (bf16 bits; zero only in padding)
};
constexpr int LIST_ALIGN = 8;            // every input pixel's list is padded to a multiple of eight entries
constexpr int ENT_BLOCK_BYTES = 48;      // final form: blocks of eight entries = 8 x u32 row + 8 x bf16 weight (csr_pack_kernel)
constexpr int SCAN_PER_BLOCK = 2048;   // keys per 256-thread block of the index scan

// ------------------------------------------------------------------------------------------------
// sampling descriptor of the backward kernels
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
  if (!valid || !off) return r;   // off == nullptr: plain convolution (zero offsets)
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


// SMs the persistent kernels may occupy: all of them minus the caller's reserve (sdb_set_sm_reserve) -- a statically
// scheduled persistent CTA that has to wait for an SM held by a co-running collective delays the whole kernel by that wait
extern int g_sm_reserve;
inline int num_sms();
inline int grid_sms() {
  const int n = num_sms() - g_sm_reserve;
  return n > 1 ? n : 1;
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
