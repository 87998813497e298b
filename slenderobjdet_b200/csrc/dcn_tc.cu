// dcn_tc.cu -- deformable convolution on tcgen05 tensor cores (SDB_MATH_BF16), sm_100a only.
//
// Forward = implicit GEMM  out[p, o] = sum_{tap, c} col[p, (tap,c)] * W[o, (tap,c)]
//   M = 128 output pixels per CTA tile, N = C_out (<= 256), K = taps * C_in, tap-major K order.
//   * A operand (col) is never materialised in HBM: producer warps sample the input (NHWC bf16,
//     one 16-byte vector load per corner per lane = 8 channels) with fp32 bilinear arithmetic and
//     store bf16 rows straight into 128B-swizzled shared memory, the layout tcgen05.mma reads;
//   * B operand (weights) is pre-permuted once per call into that same swizzled tile format in
//     global memory, so one warp streams it with cp.async.bulk (TMA engine) + mbarrier tx counts;
//   * one thread issues tcgen05.mma (128 x C_out x 16, bf16 -> fp32) into a TMEM accumulator,
//     double-buffered so the epilogue of tile i overlaps the main loop of tile i+1;
//   * 4 epilogue warps drain TMEM with tcgen05.ld, add the bias and store NCHW (coalesced along
//     the pixel dimension, which is the TMEM lane dimension).
// The grid is persistent: one CTA per SM walking tiles round-robin, and ONE launch covers every problem
// of a call (FPN levels x convolutions; problem table = kernel parameter, dcn_tc_shared.cuh): the small
// pyramid levels are a handful of tiles each and would otherwise be separate launches of 2..66 CTAs.
//
// Reference semantics being reproduced: d2/layers/csrc/deformable/deform_conv_cuda_kernel.cu
// :96-130 (bilinear), :216-288 (im2col + validity), :785-868 (mask), deform_conv_cuda.cu:397-409 (GEMM).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "dcn_tc_shared.cuh"

namespace sdb {
namespace {
using namespace tc;
using namespace tcshared;

constexpr int NPW = 8;                       // gather producer warps
constexpr int FIRST_PW = 6;                  // warps: 0 weights, 1 mma, 2-5 epilogue, 6.. gather
constexpr int NTHREADS = (FIRST_PW + NPW) * 32;
constexpr int MAX_A_STAGES = 4, MAX_B_STAGES = 8;

// Weight images.  W [O][C][taps] is re-laid-out ONCE per weight version (sdb_dcn_prepare_weights) into the two
// operand images the kernels stream with cp.async.bulk, every tile already in the 128B-swizzled K-major layout
// tcgen05.mma reads:
//   image 0 (forward B operand)   per (channel chunk of `cps`, tap, 64-channel block): [O rows][64 c]
//   image 1 (dcol = dY W^T)       per (tap, channel chunk of `nch`, 64-o block):       [nch rows (c)][64 o], o >= O zero
//   image 2 (plain-convolution grad_input = conv(dY, W')): image 0 of W'[c][o][i][j] = W[o][c][KH-1-i][KW-1-j], i.e. the
//           forward B operand of the convolution with input / output channels swapped and the taps reversed
// blockIdx.y selects the image; bias -> fp32 behind them.
struct PrepLayout {
  size_t fwd_off, dgrad_off, bias_off, conv_off, total;
  int cps, nch, okb, cps_t;
  int which;   // images to write: 1 = forward, 2 = dcol (backward data), 4 = transposed convolution
};
template <typename T>
__global__ void __launch_bounds__(256) prep_weights_kernel(const T* __restrict__ w, const T* __restrict__ bias,
                                                           uint8_t* __restrict__ img, const PrepLayout L, int O, int C,
                                                           int taps) {
  // blockIdx.y walks the requested images: bit i of L.which = image i
  int which = 0;
  for (int k = blockIdx.y; ; ++which)
    if ((L.which >> which) & 1) { if (k == 0) break; --k; }
  if (which == 0) {
    const int kbps = L.cps / 64;
    const long long total = (long long)O * taps * (C / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c8 = (int)(i % (C / 8));
      const int tap = (int)((i / (C / 8)) % taps);
      const int o = (int)(i / ((long long)(C / 8) * taps));
      const int c = c8 * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = to_f32(w[((size_t)o * C + c + j) * taps + tap]);
      uint4 pk;
      pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
      pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
      const size_t tile = ((size_t)(c / L.cps) * taps + tap) * kbps + ((c % L.cps) >> 6);
      *reinterpret_cast<uint4*>(img + L.fwd_off + tile * ((size_t)O * 128) + sw128_offset(o, (c & 63) >> 3)) = pk;
    }
    if (blockIdx.x == 0)
      for (int o = threadIdx.x; o < O; o += blockDim.x)
        reinterpret_cast<float*>(img + L.bias_off)[o] = bias ? to_f32(bias[o]) : 0.f;
    return;
  }
  if (which == 2) {
    // W' : output channels C (rows), input channels O (K), taps reversed
    const int kbps = L.cps_t / 64;
    const long long total = (long long)C * taps * (O / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c8 = (int)(i % (O / 8));
      const int tap = (int)((i / (O / 8)) % taps);
      const int orow = (int)(i / ((long long)(O / 8) * taps));   // = input channel c of W
      const int ck = c8 * 8;                                      // = output channel o of W
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = to_f32(w[((size_t)(ck + j) * C + orow) * taps + (taps - 1 - tap)]);
      uint4 pk;
      pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
      pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
      const size_t tile = ((size_t)(ck / L.cps_t) * taps + tap) * kbps + ((ck % L.cps_t) >> 6);
      *reinterpret_cast<uint4*>(img + L.conv_off + tile * ((size_t)C * 128) + sw128_offset(orow, (ck & 63) >> 3)) = pk;
    }
    return;
  }
  const int o8n = L.okb * 8;   // 8-wide o chunks incl. zero padding
  const long long total = (long long)taps * C * o8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
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
    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
    pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
    const size_t tile = ((size_t)tap * (C / L.nch) + c / L.nch) * L.okb + (o8 >> 3);
    *reinterpret_cast<uint4*>(img + L.dgrad_off + tile * ((size_t)L.nch * 128) + sw128_offset(c % L.nch, o8 & 7)) = pk;
  }
}

// ------------------------------------------------------------------------------------------------
// sampling descriptor of one (output pixel, tap): 4 corner pixel indices + 4 weights (x mask)
// ------------------------------------------------------------------------------------------------
struct Sample {
  int idx[4];   // (n*H + y)*W + x of each corner (0 when the corner does not contribute)
  float w[4];   // bilinear weight * mask (0 when the corner does not contribute)
};

// raw (dy, dx, mask) of one (pixel, tap): fetched one tap ahead so the loads overlap the gather
struct RawOff {
  float dy, dx, m;
};
__device__ __forceinline__ RawOff fetch_raw(const Geo& g, const float* __restrict__ off,
                                            const float* __restrict__ mask, bool valid, int n, int ho, int wo,
                                            int tap) {
  RawOff r = {0.f, 0.f, 1.f};
  if (!valid || !off) return r;   // off == nullptr: plain convolution (zero offsets)
  const int hw = g.Ho * g.Wo, k2 = g.KH * g.KW;
  const float* o = off + ((size_t)n * 2 * k2 + 2 * tap) * hw + ho * g.Wo + wo;
  r.dy = __ldg(o);
  r.dx = __ldg(o + hw);
  if (mask) r.m = __ldg(mask + ((size_t)n * k2 + tap) * hw + ho * g.Wo + wo);
  return r;
}

__device__ __forceinline__ Sample make_sample(const Geo& g, const RawOff raw, bool valid, int n, int ho,
                                              int wo, int tap) {
  Sample s;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s.idx[k] = 0;
    s.w[k] = 0.f;
  }
  if (!valid) return s;
  const int i = tap / g.KW, j = tap - i * g.KW;
  const float h = (float)(ho * g.sh - g.ph + i * g.dh) + raw.dy;
  const float w = (float)(wo * g.sw - g.pw + j * g.dw) + raw.dx;
  if (!(h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W)) return s;
  const float m = raw.m;
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low, hh = 1.f - lh, hw_ = 1.f - lw;
  const bool t = h_low >= 0, b = h_high <= g.H - 1, l = w_low >= 0, r = w_high <= g.W - 1;
  const int base = n * g.H;
  if (t && l) { s.idx[0] = (base + h_low) * g.W + w_low;   s.w[0] = hh * hw_ * m; }
  if (t && r) { s.idx[1] = (base + h_low) * g.W + w_high;  s.w[1] = hh * lw * m; }
  if (b && l) { s.idx[2] = (base + h_high) * g.W + w_low;  s.w[2] = lh * hw_ * m; }
  if (b && r) { s.idx[3] = (base + h_high) * g.W + w_high; s.w[3] = lh * lw * m; }
  return s;
}

// one problem of a launch (a FPN level of one convolution)
struct FwdProb {
  const __nv_bfloat16* xp;  // NHWC bf16 input
  const float* off;
  const float* mask;
  const uint8_t* wimg;      // weight image of this problem's convolution
  const float* bias;        // fp32 [O] or nullptr
  void* out;                // NCHW, f32 or bf16: [N][O][Ho][Wo]
  uint8_t* col;             // export of the sampled columns: the A stages [tile][chunk][tap][128 x CPS bf16], or nullptr
  Dims d;                   // N, H, W, Ho, Wo of this problem
  long long mP;             // N * Ho * Wo
};
struct FwdParams {
  TileMap map;              // work item (128-pixel tile) -> problem
  FwdProb pr[MAX_PROBS];
  Geo g;                    // common geometry (N, H, W, Ho, Wo come from the problem)
  int nsa, nsb;
};


// LPP = lanes per pixel in the gather (8 channels per lane): channels per A stage CPS = 8*LPP.
// CONV: plain convolution (no offsets, no mask): one source row per (pixel, tap) -- one 16-byte load per lane instead of
// four, no interpolation arithmetic; the descriptor is (row offset, inside-the-image flag).
// PAIR: the kernel runs as clusters of two CTAs (cta_group::2).  A work item is a PAIR of 128-pixel tiles of one problem:
// each CTA gathers its own tile's A rows and streams HALF of every weight tile (C_out/2 rows), the leader's MMA warp
// issues M = 256 instructions for both, so every SM writes and fetches half the B bytes per FLOP.  The non-leader's MMA
// warp relays its CTA's a_full / b_full completions to the leader; commits are multicast to both CTAs.
template <int LPP, bool OUT_BF16, bool CONV, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1) dcn_fwd_tc_kernel(const __grid_constant__ FwdParams p) {
  constexpr int CPS = LPP * 8;           // channels per A stage
  constexpr int KBPS = CPS / 64;         // 64-channel k-blocks per A stage
  constexpr int A_BYTES = TILE_M * CPS * 2;
  constexpr int PPI = 32 / LPP;          // pixels per warp instruction
  constexpr int PIX_PER_WARP = TILE_M / NPW;
  static_assert(CPS % 64 == 0 && PIX_PER_WARP % PPI == 0 && PIX_PER_WARP <= 32, "bad gather split");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[MAX_A_STAGES], a_empty[MAX_A_STAGES];
  __shared__ __align__(8) uint64_t b_full[MAX_B_STAGES], b_empty[MAX_B_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t peer_b[MAX_B_STAGES];   // PAIR, leader: the peer's weight stage is full
  __shared__ uint32_t tmem_base_s;

  const int O = p.g.O, C = p.g.C, taps = p.g.KH * p.g.KW, nchunks = C / CPS;
  const int num_work = p.map.start[p.map.n];
  const int rank = PAIR ? (int)cluster_ctarank() : 0;           // 0 = leader
  const int work0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t B_TILE = (uint32_t)O * 128u;                   // one [O x 64] weight tile of the image
  const uint32_t B_BYTES = PAIR ? B_TILE / 2 : B_TILE;          // slot size = what this CTA loads of it (PAIR: C_out/2 rows)
  const uint32_t B_LOAD = B_BYTES;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sA = sm;
  uint8_t* sB = sm + (size_t)p.nsa * A_BYTES;
  // warp index via shfl = provably warp-uniform: role branches and loop counters stay in uniform registers,
  // so tcgen05.mma takes its descriptors from the uniform datapath without an ELECT/R2UR waterfall per instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint32_t acc_stride = 32;
  while ((int)acc_stride < O) acc_stride <<= 1;
  const uint32_t ncols = 2 * acc_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nsa; ++s) {
      mbar_init(&a_full[s], PAIR ? 2 * NPW : NPW);   // one arrival per gather warp (PAIR: of both CTAs, on the leader's barrier)
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.nsb; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], PAIR ? 8 : 4);   // one arrival per epilogue warp (of both CTAs)
    }
    if (PAIR)
      for (int s = 0; s < p.nsb; ++s) mbar_init(&peer_b[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc2(&tmem_base_s, ncols);
    else tmem_alloc(&tmem_base_s, ncols);
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // ===== weight producer: bulk async copies of pre-swizzled [O x 64] tiles =====
    if (lane == 0) {
      uint32_t bs = 0, bp = 0;
      const int nkb_total = taps * (C / 64);
      for (int work = work0; work < num_work; work += wstep) {
        const uint8_t* wsrc = p.pr[find_range(p.map, work)].wimg + (size_t)rank * B_LOAD;   // PAIR: rows [rank * O/2, ...)
        for (int kb = 0; kb < nkb_total; ++kb) {
          mbar_wait(&b_empty[bs], bp ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], B_LOAD);
          bulk_g2s(sB + (size_t)bs * B_BYTES, wsrc + (size_t)kb * B_TILE, B_LOAD, &b_full[bs]);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (PAIR: leader only; the peer's warp relays its CTA's full barriers) =====
    const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * TILE_M : TILE_M, O, 0, 0);
    uint32_t as = 0, ap = 0, bs = 0, bp = 0, acc = 0, accp = 0;
    if (PAIR && rank != 0) {
      for (int work = work0; work < num_work; work += wstep)
        for (int kb = 0; kb < taps * nchunks * KBPS; ++kb) {   // the weight tiles land on local barriers (complete_tx): relay them
          mbar_wait(&b_full[bs], bp);
          if (lane == 0) mbar_arrive_remote(&peer_b[bs], 0);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
    } else
    for (int work = work0; work < num_work; work += wstep) {
      if (PAIR) mbar_wait_cluster(&acc_empty[acc], accp ^ 1);
      else mbar_wait(&acc_empty[acc], accp ^ 1);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + acc * acc_stride;
      uint32_t accumulate = 0;
      for (int it = 0; it < taps * nchunks; ++it) {
        if (PAIR) mbar_wait_cluster(&a_full[as], ap);
        else mbar_wait(&a_full[as], ap);
        for (int kb = 0; kb < KBPS; ++kb) {
          mbar_wait(&b_full[bs], bp);
          if (PAIR) mbar_wait_cluster(&peer_b[bs], bp);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + as * A_BYTES + kb * (TILE_M * 128);
            const uint32_t b_addr = smem_base + p.nsa * A_BYTES + bs * B_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (PAIR)
                umma_bf16_pair(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32, 16, 1024),
                               make_smem_desc_sw128(b_addr + k4 * 32, 16, 1024), idesc, accumulate);
              else
                umma_bf16(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32, 16, 1024),
                          make_smem_desc_sw128(b_addr + k4 * 32, 16, 1024), idesc, accumulate);
              accumulate = 1;
            }
            if (PAIR) umma_commit_pair(&b_empty[bs]);
            else umma_commit(&b_empty[bs]);
          }
          __syncwarp();
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
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
      const FwdProb& pr = p.pr[pi];
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
              const size_t di = d0 + (size_t)j * hw;
              if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(pr.out)[di] = __float2bfloat16_rn(v);
              else          reinterpret_cast<float*>(pr.out)[di] = v;
            }
          }
        }
      }
      tc_fence_before_sync();
      if (PAIR && rank != 0) {      // the accumulator pair is released on the LEADER's barrier
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
      } else {
        mbar_arrive_warp(&acc_empty[acc]);
      }
      if (++acc == 2) { acc = 0; accp ^= 1; }
    }
  } else {
    // ===== gather producers: bilinear sampling straight into the swizzled A stage =====
    // (1) per tile, the descriptors (4 source rows + 4 weights) of this warp's 16 pixels for EVERY tap
    //     are computed from the offsets and go to shared memory once;
    // (2) the gather then runs as one continuous stream over (chunk, tap, pixel pair) with a 4-slot
    //     register ring: the four 16-byte loads of iteration i+4 are issued right after iteration i
    //     is consumed, across stage boundaries, so 16 loads per warp stay in flight instead of every
    //     warp paying the full memory latency once per stage in lock-step.
    constexpr int ITERS = PIX_PER_WARP / PPI;   // warp iterations per stage
    // plain convolution: one load per iteration instead of four, so the ring is twice as deep for the same bytes in flight
    constexpr int RING = (CONV && ITERS % 8 == 0) ? 8 : 4;
    static_assert(ITERS % RING == 0, "ring must divide the per-stage iteration count");
    const int pw = warp - FIRST_PW, r0 = pw * PIX_PER_WARP;
    const int grp = lane / LPP, lig = lane % LPP;
    GDesc* sD = reinterpret_cast<GDesc*>(sB + (size_t)p.nsb * B_BYTES);   // [taps][TILE_M]
    const int nstages = taps * nchunks;
    uint32_t as = 0, ap = 0;
    for (int work = work0; work < num_work; work += wstep) {
      const int pi = find_range(p.map, work);
      const FwdProb& pr = p.pr[pi];
      const int tile = PAIR ? 2 * (work - p.map.start[pi]) + rank : work - p.map.start[pi];
      const uint4* xbase = reinterpret_cast<const uint4*>(pr.xp) + lig;
      // saved columns (sdb_dcn_problem.columns): every sampled row also goes to HBM, in the operand layout of the stage
      // it is stored to, so the weight-gradient GEMM of the backward pass streams it back instead of sampling again
      uint8_t* colp = pr.col ? pr.col + (size_t)tile * nstages * A_BYTES + (lig >> 3) * (TILE_M * 128) : nullptr;
      if (PAIR && (long long)tile * TILE_M >= pr.mP) colp = nullptr;   // the all-invalid second tile of an odd problem
      {
        const Geo g = with_dims(p.g, pr.d);
        const int px = lane % PIX_PER_WARP;
        const long long pix = (long long)tile * TILE_M + r0 + px;
        const bool valid = pix < pr.mP;
        int n = 0, ho = 0, wo = 0;
        if (valid) decode_q(g, pix, n, ho, wo);
        __syncwarp();  // every lane is done reading the previous tile's descriptors
        // all offset loads of the tile first (one exposed latency instead of one per round); the next
        // tile's offsets are pulled into L2 now so that latency is an L2 hit
        constexpr int TPR = 32 / PIX_PER_WARP;            // taps per round
        constexpr int ROUNDS = (16 + TPR - 1) / TPR;      // taps <= 16
        RawOff raw[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int tap = lane / PIX_PER_WARP + r * TPR;
          if (!CONV) raw[r] = fetch_raw(g, pr.off, pr.mask, valid && tap < taps, n, ho, wo, tap < taps ? tap : 0);
        }
        if (!CONV) {
          const int nwork = work + wstep;
          if (nwork < num_work) {
            const int npi = find_range(p.map, nwork);
            const FwdProb& npr = p.pr[npi];
            const int ntile = PAIR ? 2 * (nwork - p.map.start[npi]) + rank : nwork - p.map.start[npi];
            const long long npix = (long long)ntile * TILE_M + r0 + px;
            if (npix < npr.mP) {
              const Geo ng = with_dims(p.g, npr.d);
              int nn, nho, nwo;
              decode_q(ng, npix, nn, nho, nwo);
              const int hwo = ng.Ho * ng.Wo;
              for (int tap = lane / PIX_PER_WARP; tap < taps; tap += TPR) {
                const float* o = npr.off + ((size_t)nn * 2 * taps + 2 * tap) * hwo + nho * ng.Wo + nwo;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(o + hwo));
                if (npr.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(npr.mask + ((size_t)nn * taps + tap) * hwo + nho * ng.Wo + nwo));
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int tap = lane / PIX_PER_WARP + r * TPR;
          if (CONV) {
            if (tap < taps) {
              const int i = tap / g.KW, j = tap - i * g.KW;
              const int h = ho * g.sh - g.ph + i * g.dh, w = wo * g.sw - g.pw + j * g.dw;
              const bool ok = valid && h >= 0 && h < g.H && w >= 0 && w < g.W;
              uint4 o;
              o.x = ok ? (uint32_t)((n * g.H + h) * g.W + w) * (uint32_t)(C / 8) : 0u;
              o.y = ok ? 1u : 0u;
              o.z = o.w = 0u;
              *reinterpret_cast<uint4*>((sD + tap * TILE_M + r0 + px)->off) = o;
            }
          } else if (tap < taps) {
            const Sample sm_ = make_sample(g, raw[r], valid, n, ho, wo, tap);
            uint4 o, w;
            o.x = (uint32_t)sm_.idx[0] * (uint32_t)(C / 8); o.y = (uint32_t)sm_.idx[1] * (uint32_t)(C / 8);
            o.z = (uint32_t)sm_.idx[2] * (uint32_t)(C / 8); o.w = (uint32_t)sm_.idx[3] * (uint32_t)(C / 8);
            w.x = pack_bf16x2(sm_.w[0], sm_.w[0]); w.y = pack_bf16x2(sm_.w[1], sm_.w[1]);
            w.z = pack_bf16x2(sm_.w[2], sm_.w[2]); w.w = pack_bf16x2(sm_.w[3], sm_.w[3]);
            GDesc* d = sD + tap * TILE_M + r0 + px;
            *reinterpret_cast<uint4*>(d->off) = o;
            *reinterpret_cast<uint4*>(d->w2) = w;
          }
        }
        __syncwarp();
      }
      uint4 v[RING][4], wq[RING];
      // issue the loads of iteration `it` of the stage (tap_, ch_) into ring slot `slot`
#define SDB_ISSUE(tap_, ch_, it_, slot_)                                                     \
      {                                                                                          \
        const GDesc* d_ = sD + (tap_) * TILE_M + r0 + (it_) * PPI + grp;                         \
        const uint4 o_ = *reinterpret_cast<const uint4*>(d_->off);                               \
        const uint4* xb_ = xbase + (ch_) * (CPS / 8);                                            \
        v[slot_][0] = __ldg(xb_ + o_.x);                                                         \
        if (CONV) {                                                                              \
          wq[slot_].x = o_.y;                                                                    \
        } else {                                                                                 \
          wq[slot_] = *reinterpret_cast<const uint4*>(d_->w2);                                   \
          v[slot_][1] = __ldg(xb_ + o_.y);                                                       \
          v[slot_][2] = __ldg(xb_ + o_.z);                                                       \
          v[slot_][3] = __ldg(xb_ + o_.w);                                                       \
        }                                                                                        \
      }
#define SDB_INTERP(a_, slot_)                                                                    \
        a_.x = bf2_fma(wq[slot_].w, v[slot_][3].x, bf2_fma(wq[slot_].z, v[slot_][2].x, bf2_fma(wq[slot_].y, v[slot_][1].x, bf2_mul(wq[slot_].x, v[slot_][0].x)))); \
        a_.y = bf2_fma(wq[slot_].w, v[slot_][3].y, bf2_fma(wq[slot_].z, v[slot_][2].y, bf2_fma(wq[slot_].y, v[slot_][1].y, bf2_mul(wq[slot_].x, v[slot_][0].y)))); \
        a_.z = bf2_fma(wq[slot_].w, v[slot_][3].z, bf2_fma(wq[slot_].z, v[slot_][2].z, bf2_fma(wq[slot_].y, v[slot_][1].z, bf2_mul(wq[slot_].x, v[slot_][0].z)))); \
        a_.w = bf2_fma(wq[slot_].w, v[slot_][3].w, bf2_fma(wq[slot_].z, v[slot_][2].w, bf2_fma(wq[slot_].y, v[slot_][1].w, bf2_mul(wq[slot_].x, v[slot_][0].w))));
#pragma unroll
      for (int u = 0; u < RING; ++u) SDB_ISSUE(0, 0, u, u)
      int tap = 0, ch = 0;
      for (int st = 0; st < nstages; ++st) {
        int ntap = tap + 1, nch = ch;   // K order: chunk outermost, taps inside (L1-friendly)
        if (ntap == taps) { ntap = 0; ++nch; }
        const bool has_next = st + 1 < nstages;
        mbar_wait(&a_empty[as], ap ^ 1);
        uint8_t* dst = sA + (size_t)as * A_BYTES + (lig >> 3) * (TILE_M * 128);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int slot = it % RING;
          uint4 a;
          if (CONV) {
            a = wq[slot].x ? v[slot][0] : make_uint4(0u, 0u, 0u, 0u);
          } else {
            SDB_INTERP(a, slot)
          }
          const uint32_t soff = sw128_offset(r0 + it * PPI + grp, lig & 7);
          *reinterpret_cast<uint4*>(dst + soff) = a;
          if (colp) __stcs(reinterpret_cast<uint4*>(colp + (size_t)st * A_BYTES + soff), a);   // streaming: keep x in L2
          if (it + RING < ITERS) {
            SDB_ISSUE(tap, ch, it + RING, slot)
          } else if (has_next) {
            SDB_ISSUE(ntap, nch, it + RING - ITERS, slot)
          }
        }
        // (fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC; moving it to the MMA warp, so that the gather
        // warps do not wait for the next stage's loads they have just issued, was measured: no difference)
        fence_proxy_async_smem();
        if (PAIR && rank != 0) {      // the pair's A stage is complete when both CTAs' warps have arrived on the LEADER's barrier
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
#undef SDB_INTERP
    }
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();   // both CTAs are done with the pair's tensor memory
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc2(tmem_base, ncols);
    else tmem_dealloc(tmem_base, ncols);
  }
}

template <int LPP, bool OUT_BF16, bool CONV, bool PAIR>
int launch_fwd(const FwdParams& p, size_t smem, int grid, cudaStream_t st) {
  SDB_ENSURE_SMEM((dcn_fwd_tc_kernel<LPP, OUT_BF16, CONV, PAIR>), smem);
  ProfScope prof(SDB_OP_FORWARD, st);
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    SDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dcn_fwd_tc_kernel<LPP, OUT_BF16, CONV, PAIR>, p));
    SDB_LAUNCHED(1);
  } else {
    dcn_fwd_tc_kernel<LPP, OUT_BF16, CONV, PAIR><<<grid, NTHREADS, smem, st>>>(p); SDB_LAUNCHED(1);
  }
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

// Stage counts from the shared-memory budget; returns the dynamic smem size (0 = does not fit).  200 KB of the 228 KB
// array go to shared memory: giving the L1 more (budgets of 140 / 170 KB, 64- instead of 128-channel stages) was
// measured to change the forward by -1 .. +7 % (profiles/r2_tuning.md) -- the gather is not bound by L1 capacity.
}  // namespace
// CTA-pair forward (sdb_set_forward_pair): on by default.  Measured on the benchmarked head: forward 215 -> 209 us, the
// plain-convolution forward 147 -> 137 us (shared-memory "bank conflict" arbitration -60 %, LSU wavefronts -11 %).
int g_fwd_pair = 1;
namespace {
size_t plan_smem(FwdParams& p, size_t a_bytes, size_t b_bytes, size_t d_bytes) {
  const size_t budget = 200 * 1024;
  p.nsa = 2;
  long long nsb = ((long long)budget - 1024 - (long long)d_bytes - (long long)(p.nsa * a_bytes)) / (long long)b_bytes;
  if (nsb > MAX_B_STAGES) nsb = MAX_B_STAGES;
  if (nsb < 2) return 0;
  p.nsb = (int)nsb;
  return p.nsa * a_bytes + p.nsb * b_bytes + d_bytes + 1024;
}

}  // namespace

int tc_lanes_per_pixel(const Geo& g) { return g.C % 128 == 0 ? 16 : 8; }

bool tc_supported(const Geo& g, const char** why) {
  *why = "";
  if (g.groups != 1) { *why = "groups != 1"; return false; }
  if (g.dgroups != 1) { *why = "deformable_groups != 1"; return false; }
  if (g.C % 64 != 0) { *why = "C_in not a multiple of 64"; return false; }
  if (g.O % 16 != 0 || g.O < 16 || g.O > 256) { *why = "C_out must be a multiple of 16 in [16,256]"; return false; }
  if (g.taps() > 16) { *why = "more than 16 kernel taps (per-tile descriptors would not fit in shared memory)"; return false; }
  const long long pin = (long long)g.N * g.H * g.W;
  // 32-bit row offsets (input rows, dcol rows in 16-byte units), int key / entry counts of the transposed index
  if (pin * (g.C / 8) >= (1LL << 32) || g.P() * g.O >= (1LL << 40) || (g.P() + TILE_M) * g.taps() * (g.C / 8) >= (1LL << 32) ||
      4 * g.P() * g.taps() + (pin + TILE_M) * (g.taps() + 9) >= (1LL << 31) || g.P() + TILE_M >= (1LL << 31)) { *why = "tensor too large"; return false; }
  return true;
}

size_t tc_packed_input_bytes(const Geo& g) { return align_up((size_t)g.N * g.H * g.W * g.C * 2, 1024); }
// sampled columns of one problem: one 128-pixel x C bf16 tile per (output tile, tap)
size_t tc_columns_bytes(const Geo& g) { return align_up((size_t)cdiv(g.P(), TILE_M) * g.taps() * TILE_M * g.C * 2, 1024); }

// ---- prepared weights -------------------------------------------------------------------------------------------
static PrepLayout prep_layout(const Geo& g) {
  PrepLayout L{};
  L.cps = tc_lanes_per_pixel(g) * 8;
  L.nch = g.C % 128 == 0 ? 128 : 64;
  L.okb = 2 * ((g.O + 127) / 128);
  size_t o = 0;
  L.fwd_off = o;   o = align_up(o + (size_t)g.taps() * g.C * g.O * 2, 1024);
  L.dgrad_off = o; o = align_up(o + (size_t)g.taps() * g.C * L.okb * 64 * 2, 1024);
  L.bias_off = o;  o = align_up(o + (size_t)g.O * 4, 1024);
  L.cps_t = g.O % 128 == 0 ? 128 : 64;
  L.conv_off = o;  o = align_up(o + (size_t)g.taps() * g.C * g.O * 2, 1024);
  L.total = o;
  return L;
}
size_t tc_prepared_weight_bytes(const Geo& g) { return prep_layout(g).total; }
TcWeightImages tc_weight_images(const Geo& g, const void* prepared, bool has_bias) {
  const PrepLayout L = prep_layout(g);
  const uint8_t* b = (const uint8_t*)prepared;
  TcWeightImages w;
  w.fwd = b + L.fwd_off; w.dgrad = b + L.dgrad_off; w.convt = b + L.conv_off;
  w.bias = has_bias ? (const float*)(b + L.bias_off) : nullptr;
  return w;
}
int tc_prepare_weights(const void* w, const void* bias, const Geo& g, int io_dtype, void* prepared, int which,
                       cudaStream_t st) {
  PrepLayout L = prep_layout(g);
  L.which = which & 7;
  const int nimg = (which & 1) + ((which >> 1) & 1) + ((which >> 2) & 1);
  if (nimg == 0) return SDB_OK;
  const long long total = (long long)g.taps() * g.C * L.okb * 8;
  const int blocks = (int)((total + 255) / 256 < 592 ? (total + 255) / 256 : 592);
  dim3 grid(blocks, nimg);
  if (io_dtype == SDB_F32)
    prep_weights_kernel<float><<<grid, 256, 0, st>>>((const float*)w, (const float*)bias, (uint8_t*)prepared, L, g.O, g.C, g.taps());
  else
    prep_weights_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, (uint8_t*)prepared, L, g.O, g.C, g.taps());
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

// ---- forward, all problems in one launch -----------------------------------------------------------------------
int tc_forward_multi(const TcProblem* pb, int n, const Geo& g, int io_dtype, cudaStream_t st) {
  FwdParams p{};
  p.g = g;
  p.map.n = n;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const Geo gi = with_dims(g, pb[i].d);
    FwdProb& q = p.pr[i];
    q.xp = (const __nv_bfloat16*)pb[i].xp; q.off = pb[i].off; q.mask = pb[i].mask; q.wimg = pb[i].w.fwd;
    q.bias = pb[i].w.bias; q.out = pb[i].out; q.col = pb[i].col; q.d = pb[i].d;
    q.mP = gi.P();
    p.map.start[i] = total;
    total += cdiv(gi.P(), TILE_M);
  }
  p.map.start[n] = total;
  if (total == 0) return SDB_OK;
  const int lpp = tc_lanes_per_pixel(g);
  const size_t a_bytes = (size_t)TILE_M * lpp * 8 * 2, b_bytes = (size_t)g.O * 128;
  const size_t d_bytes = (size_t)g.taps() * TILE_M * sizeof(GDesc);   // per-tile sampling descriptors
  const size_t smem = plan_smem(p, a_bytes, b_bytes, d_bytes);
  SDB_REQUIRE(smem > 0, SDB_ERR_UNSUPPORTED, "shared memory budget too small for this geometry");
  {
    const int rcw = tc_prep_wait(st);   // weight images prepared in this call, on the side stream
    if (rcw) return rcw;
  }
  const bool obf = io_dtype == SDB_BF16;
  const bool conv = pb[0].off == nullptr;   // plain convolution: every problem of the call (api.cu checks that they agree)
  // CTA pairs: work items are PAIRS of tiles of one problem (an odd tile count leaves the second CTA of the last pair an
  // all-invalid tile), C_out split in two halves of a multiple of 16 rows
  const bool pair = g_fwd_pair && g.O % 32 == 0 && lpp == 16;
  if (pair) {
    // half-size weight slots free shared memory for a third A stage (the two CTAs advance in lock-step: slack helps)
    p.nsa = 3;   // measured: 2 or 4 A stages (with 6 / 2 weight slots) are 1-2 % slower
    long long nsb = ((long long)(200 * 1024) - 1024 - (long long)d_bytes - (long long)(p.nsa * a_bytes)) / (long long)(b_bytes / 2);
    if (nsb > MAX_B_STAGES) nsb = MAX_B_STAGES;
    SDB_REQUIRE(nsb >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for this geometry");
    p.nsb = (int)nsb;
    const size_t smem2 = p.nsa * a_bytes + p.nsb * (b_bytes / 2) + d_bytes + 1024;
    total = 0;
    for (int i = 0; i < n; ++i) {
      p.map.start[i] = total;
      total += cdiv(cdiv(with_dims(g, pb[i].d).P(), TILE_M), 2);
    }
    p.map.start[n] = total;
    const int clusters = total < grid_sms() / 2 ? total : grid_sms() / 2;
    if (conv) return obf ? launch_fwd<16, true, true, true>(p, smem2, 2 * clusters, st) : launch_fwd<16, false, true, true>(p, smem2, 2 * clusters, st);
    return obf ? launch_fwd<16, true, false, true>(p, smem2, 2 * clusters, st) : launch_fwd<16, false, false, true>(p, smem2, 2 * clusters, st);
  }
  const int grid = total < grid_sms() ? total : grid_sms();
  if (conv) {
    if (lpp == 16) return obf ? launch_fwd<16, true, true, false>(p, smem, grid, st) : launch_fwd<16, false, true, false>(p, smem, grid, st);
    return obf ? launch_fwd<8, true, true, false>(p, smem, grid, st) : launch_fwd<8, false, true, false>(p, smem, grid, st);
  }
  if (lpp == 16) return obf ? launch_fwd<16, true, false, false>(p, smem, grid, st) : launch_fwd<16, false, false, false>(p, smem, grid, st);
  return obf ? launch_fwd<8, true, false, false>(p, smem, grid, st) : launch_fwd<8, false, false, false>(p, smem, grid, st);
}

}  // namespace sdb
