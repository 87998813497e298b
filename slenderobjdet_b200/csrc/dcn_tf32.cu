// dcn_tf32.cu -- deformable-convolution FORWARD on tcgen05 tensor cores with float32 tensors
// (SDB_MATH_TF32 / SDB_MATH_TF32X3), sm_100a only.
//
// Same implicit GEMM as dcn_tc.cu (out[p, o] = sum_{tap, c} col[p, (tap, c)] * W[o, (tap, c)], M = 128 output pixels
// per tile, persistent grid, one launch over every problem of a call), but nothing is rounded to bf16:
//   * the input is NHWC float32; a gather lane loads 16 bytes = 4 channels per corner and interpolates in fp32 with
//     the reference's expression (deform_conv_cuda_kernel.cu:96-130), so the sampled column is the reference's fp32
//     `columns` value up to fp32 rounding;
//   * tcgen05.mma.kind::tf32 (K = 8 per instruction, fp32 accumulation in TMEM).  A 128-byte swizzled operand row is
//     32 channels, so a stage is (one tap, 32 channels);
//   * PASSES = 1 (SDB_MATH_TF32): column and weight are rounded to tf32 (round-to-nearest, 10 mantissa bits):
//     rel ~4e-4 on the output;
//   * PASSES = 3 (SDB_MATH_TF32X3): error-compensated split a = a_hi + a_lo, w = w_hi + w_lo (hi = tf32(a), lo = a - hi
//     exactly) and out += a_lo*w_hi + a_hi*w_lo + a_hi*w_hi: the dropped term a_lo*w_lo is ~2^-22 relative, i.e. the
//     operands are fp32-accurate; measured rel 1.5e-5 against the fp32 oracle at K = 2304 (what is left is the
//     accumulator's fp32 additions) at three tensor-core passes -- what makes the north star's "<= 1e-4 for tf32/fp32"
//     reachable on tensor cores.
//   * C_out is covered in halves of <= 128 accumulator columns so a weight tile is at most 16 KB.
// The backward of these math modes runs the exact fp32 kernels (dcn_simt.cu): gradients keep fp32 accuracy.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "dcn_tc_shared.cuh"

namespace sdb {
extern int g_fwd_pair;   // dcn_tc.cu: CTA-pair forward switch (sdb_set_forward_pair)
namespace {
using namespace tc;
using namespace tcshared;

constexpr int NPW = 8;                       // gather producer warps
constexpr int FIRST_PW = 6;                  // warps: 0 weights, 1 mma, 2-5 epilogue, 6.. gather
constexpr int NTHREADS = (FIRST_PW + NPW) * 32;
constexpr int MAX_A = 4, MAX_B = 8;
constexpr int CPS = 32;                      // channels per stage = one 128-byte operand row of fp32
constexpr int LPP = CPS / 4;                 // gather lanes per pixel (4 channels per lane)
constexpr int PART_BYTES = TILE_M * 128;     // one [128 x 32 fp32] operand tile
constexpr int B_SLOT = 128 * 128;            // weight tile slot: <= 128 output channels x 32 input channels

__device__ __forceinline__ uint32_t rna_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return u;
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4)                      // D format: f32
         | (2u << 7) | (2u << 10)       // A, B format: tf32, both K-major
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- layout kernels ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC fp32 for every problem of a call in one launch (32 channels x 32 pixels per block)
struct PackF32Table {
  TileMap map;
  struct E { const float* src; float* dst; int HW, nblk; } e[MAX_PROBS];
  int C;
};
__global__ void __launch_bounds__(256) pack_nhwc_f32_kernel(const __grid_constant__ PackF32Table t) {
  __shared__ float s[32][33];
  const int ei = find_range(t.map, blockIdx.x);
  const int local = blockIdx.x - t.map.start[ei];
  const int HW = t.e[ei].HW, C = t.C;
  const int n = local / t.e[ei].nblk, p0 = (local % t.e[ei].nblk) * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* sp = t.e[ei].src + ((size_t)n * C + c0) * HW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = ty + 8 * j;
    s[c][tx] = (p0 + tx < HW) ? __ldg(sp + (size_t)c * HW + p0 + tx) : 0.f;
  }
  __syncthreads();
  float* dp = t.e[ei].dst + ((size_t)n * HW + p0) * C + c0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int px = ty + 8 * j;
    if (p0 + px < HW) dp[(size_t)px * C + tx] = s[tx][px];
  }
}

// Weight image of the tf32 forward: per stage (32-channel chunk outermost, tap inside), per half of <= 128 output
// channels, per part (hi [, lo]) one 128B-swizzled K-major tile [rows = output channels][32 input channels], in the
// order the kernel streams them.  fp32 bias behind the image.
struct PrepF32 {
  size_t bias_off, total;
  int parts;
};
__host__ __device__ inline PrepF32 prep_f32_layout(int O, int C, int taps, int passes) {
  PrepF32 L{};
  L.parts = passes == 3 ? 2 : 1;
  L.bias_off = align_up((size_t)taps * C * O * 4 * L.parts, 1024);
  L.total = align_up(L.bias_off + (size_t)O * 4, 1024);
  return L;
}
__global__ void __launch_bounds__(256) prep_weights_tf32_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                                uint8_t* __restrict__ img, int O, int C, int taps, int parts,
                                                                size_t bias_off) {
  const long long total = (long long)O * taps * (C / 4);
  const size_t stage_bytes = (size_t)parts * O * 128;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (C / 4));
    const int tap = (int)((i / (C / 4)) % taps);
    const int o = (int)(i / ((long long)(C / 4) * taps));
    const int c = c4 * 4;
    uint4 hi, lo;
    uint32_t* hp = &hi.x;
    uint32_t* lp = &lo.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = w[((size_t)o * C + c + j) * taps + tap];
      hp[j] = rna_tf32(v);
      lp[j] = rna_tf32(v - __uint_as_float(hp[j]));
    }
    const int h = o >> 7, ol = o & 127;
    const int nh = min(128, O - h * 128);
    // halves before h hold 128 rows each; inside a half: hi tile, then lo tile
    uint8_t* tile = img + ((size_t)(c / CPS) * taps + tap) * stage_bytes + (size_t)h * 128 * 128 * parts;
    const uint32_t so = sw128_offset(ol, (c % CPS) >> 2);
    *reinterpret_cast<uint4*>(tile + so) = hi;
    if (parts == 2) *reinterpret_cast<uint4*>(tile + (size_t)nh * 128 + so) = lo;
  }
  if (blockIdx.x == 0)
    for (int o = threadIdx.x; o < O; o += blockDim.x)
      reinterpret_cast<float*>(img + bias_off)[o] = bias ? bias[o] : 0.f;
}

// ---- forward kernel ---------------------------------------------------------------------------------------------
struct F32Prob {
  const float* xp;          // NHWC fp32 input
  const float* off;
  const float* mask;
  const uint8_t* wimg;
  const float* bias;        // fp32 [O] or nullptr
  float* out;               // NCHW fp32
  Dims d;
  long long mP;             // N * Ho * Wo
};
struct F32Params {
  TileMap map;
  F32Prob pr[MAX_PROBS];
  Geo g;
  int nsa, nsb;
};

// descriptor of one (pixel, tap): corner rows in 16-byte units into the NHWC fp32 input + fp32 weights (x mask)
struct __align__(16) FDesc {
  uint32_t off[4];
  float w[4];
};

__device__ __forceinline__ void make_fdesc(const Geo& g, float dy, float dx, float m, bool valid, int n, int ho, int wo,
                                           int tap, uint32_t row_units, FDesc& d) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { d.off[k] = 0; d.w[k] = 0.f; }
  if (!valid) return;
  const int i = tap / g.KW, j = tap - i * g.KW;
  const float h = (float)(ho * g.sh - g.ph + i * g.dh) + dy;
  const float w = (float)(wo * g.sw - g.pw + j * g.dw) + dx;
  if (!(h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W)) return;   // deform_conv_cuda_kernel.cu:273, :852
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low, hh = 1.f - lh, hw_ = 1.f - lw;
  const bool t = h_low >= 0, b = h_high <= g.H - 1, l = w_low >= 0, r = w_high <= g.W - 1;
  const int base = n * g.H;
  if (t && l) { d.off[0] = (uint32_t)((base + h_low) * g.W + w_low) * row_units;   d.w[0] = hh * hw_ * m; }
  if (t && r) { d.off[1] = (uint32_t)((base + h_low) * g.W + w_high) * row_units;  d.w[1] = hh * lw * m; }
  if (b && l) { d.off[2] = (uint32_t)((base + h_high) * g.W + w_low) * row_units;  d.w[2] = lh * hw_ * m; }
  if (b && r) { d.off[3] = (uint32_t)((base + h_high) * g.W + w_high) * row_units; d.w[3] = lh * lw * m; }
}

// PAIR: CTA pairs (cta_group::2, M = 256) exactly as in dcn_tc.cu: a work item is a pair of tiles of one problem, each CTA
// streams half the rows of every weight tile, the leader issues, commits are multicast, the peer relays its weight barriers.
template <int PASSES, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1) dcn_fwd_tf32_kernel(const __grid_constant__ F32Params p) {
  constexpr int PARTS = PASSES == 3 ? 2 : 1;
  constexpr int SLOT = PAIR ? B_SLOT / 2 : B_SLOT;   // weight slot: this CTA's rows of a tile
  constexpr int A_BYTES = PARTS * PART_BYTES;
  constexpr int PPI = 32 / LPP;                // pixels per warp instruction (4)
  constexpr int PIX_PER_WARP = TILE_M / NPW;   // 16
  constexpr int ITERS = PIX_PER_WARP / PPI;    // 4

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[MAX_A], a_empty[MAX_A];
  __shared__ __align__(8) uint64_t b_full[MAX_B], b_empty[MAX_B];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t peer_b[MAX_B];   // PAIR, leader: the peer's weight slot is full
  __shared__ uint32_t tmem_base_s;

  const int O = p.g.O, C = p.g.C, taps = p.g.KH * p.g.KW, nchunks = C / CPS;
  const int nstages = taps * nchunks;
  const int nhalves = (O + 127) >> 7;
  const int num_work = p.map.start[p.map.n];   // PAIR: tile pairs
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int work0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sA = sm;
  uint8_t* sB = sm + (size_t)p.nsa * A_BYTES;
  const uint32_t sB_u32 = smem_base + (uint32_t)p.nsa * A_BYTES;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint32_t acc_stride = 32;
  while ((int)acc_stride < O) acc_stride <<= 1;
  const uint32_t ncols = 2 * acc_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nsa; ++s) { mbar_init(&a_full[s], PAIR ? 2 * NPW : NPW); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.nsb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PAIR ? 8 : 4); }
    if (PAIR)
      for (int s = 0; s < p.nsb; ++s) mbar_init(&peer_b[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc2(&tmem_base_s, ncols);
    else tmem_alloc(&tmem_base_s, ncols);
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // ===== weight producer: one bulk copy per (stage, half, part) tile, in image order =====
    if (lane == 0) {
      uint32_t bs = 0, bp = 0;
      for (int work = work0; work < num_work; work += wstep) {
        const uint8_t* src = p.pr[find_range(p.map, work)].wimg;
        for (int st = 0; st < nstages; ++st)
          for (int h = 0; h < nhalves; ++h) {
            const uint32_t tile_bytes = (uint32_t)min(128, O - h * 128) * 128u;
            const uint32_t bytes = PAIR ? tile_bytes / 2 : tile_bytes;   // PAIR: rows [rank * N_h/2, ...) of the tile
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              mbar_wait(&b_empty[bs], bp ^ 1);
              mbar_arrive_expect_tx(&b_full[bs], bytes);
              bulk_g2s(sB + (size_t)bs * SLOT, src + (size_t)rank * bytes, bytes, &b_full[bs]);
              src += tile_bytes;
              if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (PAIR: leader only; the peer's warp relays its weight barriers) =====
    uint32_t as = 0, ap = 0, bs = 0, bp = 0, acc = 0, accp = 0;
    if (PAIR && rank != 0) {
      for (int work = work0; work < num_work; work += wstep)
        for (int i = 0; i < nstages * nhalves * PARTS; ++i) {
          mbar_wait(&b_full[bs], bp);
          if (lane == 0) mbar_arrive_remote(&peer_b[bs], 0);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
    } else
    for (int work = work0; work < num_work; work += wstep) {
      if (PAIR) mbar_wait_cluster(&acc_empty[acc], accp ^ 1);
      else mbar_wait(&acc_empty[acc], accp ^ 1);
      tc_fence_after_sync();
      for (int st = 0; st < nstages; ++st) {
        if (PAIR) mbar_wait_cluster(&a_full[as], ap);
        else mbar_wait(&a_full[as], ap);
        const uint32_t a_hi = smem_base + as * A_BYTES, a_lo = a_hi + PART_BYTES;
        for (int h = 0; h < nhalves; ++h) {
          const uint32_t idesc = make_idesc_tf32(PAIR ? 2 * TILE_M : TILE_M, min(128, O - h * 128));
          const uint32_t s_hi = bs;
          mbar_wait(&b_full[bs], bp);
          if (PAIR) mbar_wait_cluster(&peer_b[bs], bp);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
          uint32_t s_lo = s_hi;
          if (PARTS == 2) {
            s_lo = bs;
            mbar_wait(&b_full[bs], bp);
            if (PAIR) mbar_wait_cluster(&peer_b[bs], bp);
            if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
          }
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t b_hi = sB_u32 + s_hi * SLOT, b_lo = sB_u32 + s_lo * SLOT;
            const uint32_t tmem_d = tmem_base + acc * acc_stride + (uint32_t)h * 128u;
            uint32_t accumulate = st > 0 ? 1u : 0u;
            auto mma = [&](uint32_t a, uint32_t b, uint32_t accu) {
              if (PAIR) umma_tf32_pair(tmem_d, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc, accu);
              else umma_tf32(tmem_d, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc, accu);
            };
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (PASSES == 3) {
                mma(a_lo + k4 * 32, b_hi + k4 * 32, accumulate);
                mma(a_hi + k4 * 32, b_lo + k4 * 32, 1u);
                accumulate = 1u;
              }
              mma(a_hi + k4 * 32, b_hi + k4 * 32, accumulate);
              accumulate = 1u;
            }
            if (PAIR) {
              umma_commit_pair(&b_empty[s_hi]);
              if (PARTS == 2) umma_commit_pair(&b_empty[s_lo]);
            } else {
              umma_commit(&b_empty[s_hi]);
              if (PARTS == 2) umma_commit(&b_empty[s_lo]);
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (PAIR) umma_commit_pair(&a_empty[as]);
          else umma_commit(&a_empty[as]);
        }
        __syncwarp();
        if (++as == (uint32_t)p.nsa) { as = 0; ap ^= 1; }
      }
      if (elect_one()) {
        if (PAIR) umma_commit_pair(&acc_full[acc]);
        else umma_commit(&acc_full[acc]);
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; accp ^= 1; }
    }
  } else if (warp < FIRST_PW) {
    // ===== epilogue: TMEM -> registers -> NCHW global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t acc = 0, accp = 0;
    for (int work = work0; work < num_work; work += wstep) {
      const int pi = find_range(p.map, work);
      const F32Prob& pr = p.pr[pi];
      const int tile = PAIR ? 2 * (work - p.map.start[pi]) + rank : work - p.map.start[pi];
      const int hw = pr.d.Ho * pr.d.Wo;
      mbar_wait(&acc_full[acc], accp);
      tc_fence_after_sync();
      const long long pix = (long long)tile * TILE_M + q * 32 + lane;
      const bool valid = pix < pr.mP;
      int n = 0, eho = 0, ewo = 0;
      if (valid) decode_pos(pr.d.Ho, pr.d.Wo, p.g.th, p.g.tw, pix, n, eho, ewo);
      const int rem = eho * pr.d.Wo + ewo;
      const float* bias = pr.bias;
      for (int c0 = 0; c0 < O; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16) + c0, r);
        tmem_ld_wait();
        if (valid) {
          const size_t d0 = ((size_t)n * O + c0) * hw + rem;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int o = c0 + j;
            if (o < O) {
              float v = __uint_as_float(r[j]);
              if (bias) v += __ldg(bias + o);
              pr.out[d0 + (size_t)j * hw] = v;
            }
          }
        }
      }
      tc_fence_before_sync();
      if (PAIR && rank != 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
      } else {
        mbar_arrive_warp(&acc_empty[acc]);
      }
      if (++acc == 2) { acc = 0; accp ^= 1; }
    }
  } else {
    // ===== gather producers: fp32 bilinear sampling into the swizzled A stage (hi [, lo] tiles) =====
    constexpr int RING = 4;
    static_assert(ITERS == RING, "one ring revolution per stage");
    const int pw = warp - FIRST_PW, r0 = pw * PIX_PER_WARP;
    const int grp = lane / LPP, lig = lane % LPP;
    FDesc* sD = reinterpret_cast<FDesc*>(sB + (size_t)p.nsb * SLOT);   // [taps][TILE_M]
    const uint32_t row_units = (uint32_t)(C / 4);
    uint32_t as = 0, ap = 0;
    for (int work = work0; work < num_work; work += wstep) {
      const int pi = find_range(p.map, work);
      const F32Prob& pr = p.pr[pi];
      const int tile = PAIR ? 2 * (work - p.map.start[pi]) + rank : work - p.map.start[pi];
      const uint4* xbase = reinterpret_cast<const uint4*>(pr.xp) + lig;
      {
        const Geo g = with_dims(p.g, pr.d);
        const int px = lane % PIX_PER_WARP;
        const long long pix = (long long)tile * TILE_M + r0 + px;
        const bool valid = pix < pr.mP;
        int n = 0, ho = 0, wo = 0;
        if (valid) decode_q(g, pix, n, ho, wo);
        __syncwarp();  // every lane is done reading the previous tile's descriptors
        constexpr int TPR = 32 / PIX_PER_WARP;            // taps per round
        constexpr int ROUNDS = (16 + TPR - 1) / TPR;      // taps <= 16
        const int hwo = g.Ho * g.Wo;
        float dy[ROUNDS], dx[ROUNDS], mk[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int tap = lane / PIX_PER_WARP + r * TPR;
          dy[r] = dx[r] = 0.f;
          mk[r] = 1.f;
          if (valid && tap < taps) {
            const float* o = pr.off + ((size_t)n * 2 * taps + 2 * tap) * hwo + ho * g.Wo + wo;
            dy[r] = __ldg(o);
            dx[r] = __ldg(o + hwo);
            if (pr.mask) mk[r] = __ldg(pr.mask + ((size_t)n * taps + tap) * hwo + ho * g.Wo + wo);
          }
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int tap = lane / PIX_PER_WARP + r * TPR;
          if (tap < taps) {
            FDesc d;
            make_fdesc(g, dy[r], dx[r], mk[r], valid, n, ho, wo, tap, row_units, d);
            FDesc* dst = sD + tap * TILE_M + r0 + px;
            *reinterpret_cast<uint4*>(dst->off) = *reinterpret_cast<const uint4*>(d.off);
            *reinterpret_cast<float4*>(dst->w) = *reinterpret_cast<const float4*>(d.w);
          }
        }
        __syncwarp();
      }
      uint4 v[RING][4];
      float4 wq[RING];
#define SDB_ISSUE(tap_, ch_, it_, slot_)                                                         \
      {                                                                                          \
        const FDesc* d_ = sD + (tap_) * TILE_M + r0 + (it_) * PPI + grp;                         \
        const uint4 o_ = *reinterpret_cast<const uint4*>(d_->off);                               \
        wq[slot_] = *reinterpret_cast<const float4*>(d_->w);                                     \
        const uint4* xb_ = xbase + (ch_) * (CPS / 4);                                            \
        v[slot_][0] = __ldg(xb_ + o_.x);                                                         \
        v[slot_][1] = __ldg(xb_ + o_.y);                                                         \
        v[slot_][2] = __ldg(xb_ + o_.z);                                                         \
        v[slot_][3] = __ldg(xb_ + o_.w);                                                         \
      }
#define SDB_LERP(f_, slot_)                                                                      \
        fmaf(wq[slot_].w, __uint_as_float(v[slot_][3].f_),                                       \
             fmaf(wq[slot_].z, __uint_as_float(v[slot_][2].f_),                                  \
                  fmaf(wq[slot_].y, __uint_as_float(v[slot_][1].f_), wq[slot_].x * __uint_as_float(v[slot_][0].f_))))
#pragma unroll
      for (int u = 0; u < RING; ++u) SDB_ISSUE(0, 0, u, u)
      int tap = 0, ch = 0;
      for (int st = 0; st < nstages; ++st) {
        int ntap = tap + 1, nch = ch;   // K order: 32-channel chunk outermost, taps inside
        if (ntap == taps) { ntap = 0; ++nch; }
        const bool has_next = st + 1 < nstages;
        mbar_wait(&a_empty[as], ap ^ 1);
        uint8_t* dst = sA + (size_t)as * A_BYTES;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const float a0 = SDB_LERP(x, it), a1 = SDB_LERP(y, it), a2 = SDB_LERP(z, it), a3 = SDB_LERP(w, it);
          uint4 hi;
          hi.x = rna_tf32(a0); hi.y = rna_tf32(a1); hi.z = rna_tf32(a2); hi.w = rna_tf32(a3);
          const uint32_t soff = sw128_offset(r0 + it * PPI + grp, lig);
          *reinterpret_cast<uint4*>(dst + soff) = hi;
          if (PASSES == 3) {
            uint4 lo;   // exact residual; the tensor core reads its top 19 bits
            lo.x = __float_as_uint(a0 - __uint_as_float(hi.x)); lo.y = __float_as_uint(a1 - __uint_as_float(hi.y));
            lo.z = __float_as_uint(a2 - __uint_as_float(hi.z)); lo.w = __float_as_uint(a3 - __uint_as_float(hi.w));
            *reinterpret_cast<uint4*>(dst + PART_BYTES + soff) = lo;
          }
          if (has_next) SDB_ISSUE(ntap, nch, it, it)
        }
        fence_proxy_async_smem();
        if (PAIR && rank != 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&a_full[as], 0);
        } else {
          mbar_arrive_warp(&a_full[as]);
        }
        if (++as == (uint32_t)p.nsa) { as = 0; ap ^= 1; }
        tap = ntap;
        ch = nch;
      }
#undef SDB_ISSUE
#undef SDB_LERP
    }
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc2(tmem_base, ncols);
    else tmem_dealloc(tmem_base, ncols);
  }
}

template <int PASSES, bool PAIR>
int launch_tf32(const F32Params& p, size_t smem, int grid, cudaStream_t st) {
  SDB_ENSURE_SMEM((dcn_fwd_tf32_kernel<PASSES, PAIR>), smem);
  ProfScope prof(SDB_OP_FORWARD, st);
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    SDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dcn_fwd_tf32_kernel<PASSES, PAIR>, p));
  } else {
    dcn_fwd_tf32_kernel<PASSES, PAIR><<<grid, NTHREADS, smem, st>>>(p);
  }
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

size_t xp_bytes(const Dims& d, int C) { return align_up((size_t)d.N * d.H * d.W * C * 4, 1024); }

}  // namespace

bool tf32_supported(const Geo& g, const char** why) {
  *why = "";
  if (g.groups != 1) { *why = "groups != 1"; return false; }
  if (g.dgroups != 1) { *why = "deformable_groups != 1"; return false; }
  if (g.C % 32 != 0) { *why = "C_in not a multiple of 32"; return false; }
  if (g.O % 16 != 0 || g.O < 16 || g.O > 256) { *why = "C_out must be a multiple of 16 in [16,256]"; return false; }
  if (g.taps() > 16) { *why = "more than 16 kernel taps (per-tile descriptors would not fit in shared memory)"; return false; }
  const long long pin = (long long)g.N * g.H * g.W;
  if (pin * (g.C / 4) >= (1LL << 32) || pin >= (1LL << 31) || g.P() + TILE_M >= (1LL << 31)) { *why = "tensor too large"; return false; }
  return true;
}

size_t tf32_prepared_weight_bytes(const Geo& g, int passes) { return prep_f32_layout(g.O, g.C, g.taps(), passes).total; }

int tf32_prepare_weights(const float* w, const float* bias, const Geo& g, int passes, void* prepared, cudaStream_t st) {
  const PrepF32 L = prep_f32_layout(g.O, g.C, g.taps(), passes);
  const long long total = (long long)g.O * g.taps() * (g.C / 4);
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  prep_weights_tf32_kernel<<<blocks, 256, 0, st>>>(w, bias, (uint8_t*)prepared, g.O, g.C, g.taps(), L.parts, L.bias_off);
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

// workspace of one forward call: the NHWC fp32 copy of every problem's input, then an image for every weight tensor that
// did not come prepared
size_t tf32_forward_workspace_bytes(const TcProblem* pb, int n, int nweights, const bool* have_prepared, const Geo& g,
                                    int passes) {
  size_t o = 0;
  for (int i = 0; i < n; ++i) o += xp_bytes(pb[i].d, g.C);
  for (int k = 0; k < nweights; ++k)
    if (!have_prepared[k]) o += tf32_prepared_weight_bytes(g, passes);
  return o;
}

// pb[i]: d, weight_id, x, off, mask, out.  weights / biases / prepared: per weight tensor (prepared[k] may be NULL).
int tf32_forward_all(const TcProblem* pb, int n, const void* const* weights, const void* const* biases,
                     const void* const* prepared, int nweights, const Geo& g, int passes, uint8_t* ws, cudaStream_t st) {
  F32Params p{};
  p.g = g;
  PackF32Table t{};
  t.C = g.C;
  size_t o = 0;
  const float* xp[MAX_PROBS];
  int m = 0, blocks = 0;
  for (int i = 0; i < n; ++i) {
    xp[i] = (const float*)(ws + o);
    o += xp_bytes(pb[i].d, g.C);
    const int HW = pb[i].d.H * pb[i].d.W;
    if (pb[i].d.N * HW == 0) continue;
    t.e[m].src = (const float*)pb[i].x; t.e[m].dst = (float*)xp[i]; t.e[m].HW = HW; t.e[m].nblk = (HW + 31) / 32;
    t.map.start[m] = blocks;
    blocks += pb[i].d.N * t.e[m].nblk;
    ++m;
  }
  t.map.n = m; t.map.start[m] = blocks;
  const uint8_t* img[MAX_WEIGHTS];
  const PrepF32 L = prep_f32_layout(g.O, g.C, g.taps(), passes);
  for (int k = 0; k < nweights; ++k) {
    if (prepared[k]) { img[k] = (const uint8_t*)prepared[k]; continue; }
    int rc = tf32_prepare_weights((const float*)weights[k], (const float*)biases[k], g, passes, ws + o, st);
    if (rc) return rc;
    img[k] = ws + o;
    o += L.total;
  }
  if (blocks) {
    pack_nhwc_f32_kernel<<<dim3(blocks, g.C / 32), 256, 0, st>>>(t); SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
  }
  p.map.n = n;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const Geo gi = with_dims(g, pb[i].d);
    F32Prob& q = p.pr[i];
    q.xp = xp[i]; q.off = pb[i].off; q.mask = pb[i].mask; q.wimg = img[pb[i].weight_id];
    q.bias = biases[pb[i].weight_id] ? (const float*)(img[pb[i].weight_id] + L.bias_off) : nullptr;
    q.out = (float*)pb[i].out; q.d = pb[i].d; q.mP = gi.P();
    p.map.start[i] = total;
    total += cdiv(gi.P(), TILE_M);
  }
  p.map.start[n] = total;
  if (total == 0) return SDB_OK;
  const bool pair = g_fwd_pair && g.O % 32 == 0;
  if (pair) {   // work items = pairs of tiles of one problem
    total = 0;
    for (int i = 0; i < n; ++i) {
      p.map.start[i] = total;
      total += cdiv(cdiv(with_dims(g, pb[i].d).P(), TILE_M), 2);
    }
    p.map.start[n] = total;
  }
  const size_t budget = 200 * 1024, d_bytes = (size_t)g.taps() * TILE_M * sizeof(FDesc);
  const size_t a_bytes = (size_t)(passes == 3 ? 2 : 1) * PART_BYTES, slot = pair ? B_SLOT / 2 : B_SLOT;
  p.nsa = passes == 3 ? 2 : 3;
  long long nsb = ((long long)budget - 1024 - (long long)d_bytes - (long long)(p.nsa * a_bytes)) / (long long)slot;
  if (nsb > MAX_B) nsb = MAX_B;
  SDB_REQUIRE(nsb >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for this geometry");
  p.nsb = (int)nsb;
  const size_t smem = p.nsa * a_bytes + p.nsb * slot + d_bytes + 1024;
  if (pair) {
    const int clusters = total < grid_sms() / 2 ? total : grid_sms() / 2;
    return passes == 3 ? launch_tf32<3, true>(p, smem, 2 * clusters, st) : launch_tf32<1, true>(p, smem, 2 * clusters, st);
  }
  const int grid = total < grid_sms() ? total : grid_sms();
  return passes == 3 ? launch_tf32<3, false>(p, smem, grid, st) : launch_tf32<1, false>(p, smem, grid, st);
}

}  // namespace sdb
