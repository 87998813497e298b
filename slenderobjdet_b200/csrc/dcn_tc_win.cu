// dcn_tc_win.cu -- deformable convolution forward on tcgen05 with the input halo window of every
// tile staged in shared memory by TMA (cp.async.bulk.tensor), sm_100a only.  Opt-in: SDB_TC_WIN=1.
//
// Why: in the L2-gather kernel (dcn_tc.cu) every (pixel, tap, corner) is a 16-byte load per lane that
// mostly misses L1: a 128-pixel tile pulls 2.4 MB of corner rows + 1.2 MB of weights through the L2->SM
// fabric and fills + reads every corner line in the L1 data array.  Here a tile is a compact th x tw patch
// of ONE image; for each 64-channel chunk the window producer issues one 4-D tensor-map copy of the
// patch's sampling window [BH][BW][64 ch] (out-of-image rows/columns zero-filled by TMA, which is exactly
// the reference's "corner outside the image reads 0" rule, deform_conv_cuda_kernel.cu:104-128) into one
// of two window buffers, and 16 gather warps read the four bilinear corners with LDS.128 (conflict-free:
// 8 lanes cover one pixel's 128 bytes).  Samples whose corners fall outside the window (|offset| > R,
// R = 3 px for C_out = 256) take a per-pixel path through global memory one stage ahead in the same
// register ring, so any offset is still exact.  Weight stages can be fetched half each by the two CTAs of
// a cluster and multicast.  K order: (chunk of 64 channels, tap); A stage = 128 px x 64 ch, 128B-swizzled;
// B stage = C_out x 64; two stages of each.
//
// Status (profiles/r1_win_trace.md): parity green; ~20 % slower than dcn_tc.cu on the RepPoints shapes
// because the gather's per-stage wait / fence / arrive is not overlapped with shared-memory reads yet.
//
// Reference semantics: d2/layers/csrc/deformable/deform_conv_cuda_kernel.cu:96-130 (bilinear),
// :216-288 (im2col + validity), :785-868 (mask), deform_conv_cuda.cu:397-409 (GEMM).
#include <cuda.h>   // CUtensorMap types; the encoder is fetched with cudaGetDriverEntryPoint (no -lcuda)
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "dcn_tc_shared.cuh"

namespace sdb {
namespace {
using namespace tc;
using namespace tcshared;

constexpr int W_NPW = 16;                      // gather warps (8 pixels x 64 channels per stage each)
constexpr int W_FIRST_PW = 7;                  // warps: 0 weights, 1 mma, 2-5 epilogue, 6 window TMA, 7.. gather
constexpr int W_NTHREADS = (W_FIRST_PW + W_NPW) * 32;
constexpr int W_NSA = 4;                       // A stages (16 KB each): the gather runs up to four stages ahead of the MMAs
constexpr int W_NSB = 2;                       // B (weight) stages, O x 128 B each
constexpr int W_A_BYTES = TILE_M * 128;        // one A stage: 128 pixels x 64 channels bf16
constexpr int W_PIXW = TILE_M / W_NPW;         // pixels per gather warp
constexpr int W_LPP = 4;                       // lanes per pixel: 2 x 8 channels each (the gather is issue-bound: fewer, fatter iterations)
constexpr int W_PPI = 32 / W_LPP;              // pixels per warp instruction
constexpr int W_ITERS = W_PIXW / W_PPI;
constexpr int W_RING = W_ITERS;                // register ring slots = one stage (one warp iteration = 8 pixels x 64 channels)
static_assert(W_PIXW % W_PPI == 0 && W_ITERS >= 1, "bad gather split");

struct WinParams {
  CUtensorMap tmap;          // NHWC bf16 input as (c, x, y, n), box (64, BW, BH, 1)
  const __nv_bfloat16* xp;   // same tensor, for out-of-window samples
  const float* off;
  const float* mask;
  const uint8_t* wimg;       // weight tiles ordered (chunk, tap), each [O][64] bf16 SW128
  const float* bias;
  void* out;
  Geo g;
  int tiles_x, tiles_y, num_tiles;
  int R, BH, BW;
  uint32_t win_bytes;
  int stagger_ns;            // per-warp start offset after each window wait (x (warp & 3)), ns
  int cl;                    // CTAs per cluster (1 or 2): weight stages are loaded half each and multicast
  int dbg;
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2,
                                            int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_dst),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// half of a weight stage -> the same offset in every CTA of the cluster, signalling each CTA's own barrier
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                            uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// arrive on the same barrier in every CTA of `mask` once all MMAs issued so far by this thread are done
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// W [O][C][taps] -> tiles ordered (64-channel chunk, tap, column half), each K-major 128B-swizzled
// [ncol_h rows (o)][64 c] bf16; also bias -> fp32.
template <typename T>
__global__ void __launch_bounds__(256) prep_weight_win_kernel(const T* __restrict__ w, const T* __restrict__ bias,
                                                              uint8_t* __restrict__ wimg, float* __restrict__ bias_f32,
                                                              int O, int C, int taps, int nh, int ncol_h) {
  const long long total = (long long)O * taps * (C / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % (C / 8));
    const int tap = (int)((i / (C / 8)) % taps);
    const int o = (int)(i / ((long long)(C / 8) * taps));
    const int c = c8 * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = to_f32(w[((size_t)o * C + c + j) * taps + tap]);
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]);
    pk.y = pack_bf16x2(v[2], v[3]);
    pk.z = pack_bf16x2(v[4], v[5]);
    pk.w = pack_bf16x2(v[6], v[7]);
    const size_t tile = ((size_t)(c >> 6) * taps + tap) * nh + o / ncol_h;
    *reinterpret_cast<uint4*>(wimg + tile * ((size_t)ncol_h * 128) + sw128_offset(o % ncol_h, (c & 63) >> 3)) = pk;
  }
  if (bias_f32 && blockIdx.x == 0)
    for (int o = threadIdx.x; o < O; o += blockDim.x) bias_f32[o] = bias ? to_f32(bias[o]) : 0.f;
}

// timing experiment (SDB_TC_DEBUG & 64): clock64 stamps of CTA 0's first two tiles, [role][tile][stage][slot]
} // namespace
__device__ unsigned long long g_trace[3][2][64][4];
namespace {
#define SDB_TRACE(role_, k_, st_, slot_)                                                        \
  if ((p.dbg & 64) && blockIdx.x == 0 && (k_) < 2 && (threadIdx.x & 31) == 0) g_trace[role_][k_][st_][slot_] = clock64();

// Pipeline: W_NSA = 4 A stages (128 px x 64 ch, written by the gather warps) and W_NSB = 2 B stages (O x 64
// weights, written by the TMA engine), each ring with its own full / empty barriers; ONE window buffer,
// reloaded per 64-channel chunk (a ~2 k-clock bubble per chunk, paid for a window reach of R = 5 px instead
// of 3 and for the two extra A stages that let the gather warps run decoupled).  Per stage the MMA thread does
// two waits, four 128 x O x 16 MMAs and two commits.  History of the alternatives: profiles/r1_win_trace.md.
template <bool OUT_BF16>
__global__ void __launch_bounds__(W_NTHREADS, 1) dcn_fwd_win_kernel(const __grid_constant__ WinParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[W_NSA], empty[W_NSA];       // A stages (gather warps <-> MMA)
  __shared__ __align__(8) uint64_t b_full[W_NSB], b_empty[W_NSB];   // B stages (weight producer <-> MMA)
  __shared__ __align__(8) uint64_t w_full[2], w_empty[2];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const Geo& g = p.g;
  const int O = g.O, taps = g.KH * g.KW, nchunks = g.C / 64;
  const uint32_t B_BYTES = (uint32_t)O * 128u;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sA = sm;
  uint8_t* sB = sA + W_NSA * W_A_BYTES;
  uint8_t* sWin = sB + (size_t)W_NSB * B_BYTES;
  uint4* sDesc = reinterpret_cast<uint4*>(sWin + (size_t)p.win_bytes);   // [taps][128]: off, w00|w01, w10|w11, -
  // warp index made provably warp-uniform (shfl): role branches and every loop counter below then live in
  // uniform registers, so tcgen05.mma gets its descriptors straight from the uniform datapath.  With
  // `threadIdx.x >> 5` the compiler wrapped each MMA in an ELECT / R2UR / BRA.U.ANY waterfall (~130 clk
  // per instruction, the tensor time of a whole 128x256x16 MMA).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint32_t acc_stride = 32;
  while ((int)acc_stride < O) acc_stride <<= 1;
  const uint32_t ncols = 2 * acc_stride;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int nstages = taps * nchunks;
  // Work split: every CTA of a cluster runs the SAME number of iterations, because each one loads its
  // share of every weight stage for all of them; iterations past the last tile only pass stages along.
  const int cl = p.cl;
  const uint32_t crank = cl > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << cl) - 1u);
  const int niter = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
#define SDB_TILE_OF(k_) (((k_) * ((int)gridDim.x / cl) + (int)blockIdx.x / cl) * cl + (int)crank)

  if (threadIdx.x == 0) {
    for (int s = 0; s < W_NSA; ++s) {
      mbar_init(&full[s], W_NPW);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < W_NSB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], cl);     // released by the MMA warp of every CTA in the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], W_NPW);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before_sync();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // peers' barriers are initialised before anything is multicast to them
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // ===== weight producer: bulk async copies of pre-swizzled [O x 64] tiles =====
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      const uint32_t part = B_BYTES / (uint32_t)cl;
      for (int k = 0; k < niter; ++k) {
        for (int kb = 0; kb < nstages; ++kb) {
          mbar_wait(&b_empty[s], ph ^ 1);
          if ((p.dbg & 64) && blockIdx.x == 0 && k < 2) g_trace[2][k][kb][0] = clock64();
          if (p.dbg & (1 | 128)) {
            mbar_arrive(&b_full[s]);
          } else {
            mbar_arrive_expect_tx(&b_full[s], B_BYTES);
            if (cl > 1)
              bulk_g2s_mc(sB + (size_t)s * B_BYTES + crank * part, p.wimg + (size_t)kb * B_BYTES + crank * part, part,
                          &b_full[s], cmask);
            else
              bulk_g2s(sB + (size_t)s * B_BYTES, p.wimg + (size_t)kb * B_BYTES, B_BYTES, &b_full[s]);
          }
          if (++s == W_NSB) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(TILE_M, O, 0, 0);
    const uint64_t adesc0 = make_smem_desc_sw128(smem_base, 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_base + W_NSA * W_A_BYTES, 16, 1024);
    uint32_t s = 0, ph = 0, sa = 0, pa = 0, acc = 0, accp = 0;   // s/ph: weight ring, sa/pa: A ring
    for (int k = 0; k < niter; ++k) {
      const bool real = SDB_TILE_OF(k) < p.num_tiles;
      if (real) {
        mbar_wait(&acc_empty[acc], accp ^ 1);
        tc_fence_after_sync();
      }
      const uint32_t tmem_d = tmem_base + acc * acc_stride;
      for (int it = 0; it < nstages; ++it) {
        SDB_TRACE(1, k, it, 0)
        mbar_wait(&b_full[s], ph);
        if (real) mbar_wait(&full[sa], pa);
        SDB_TRACE(1, k, it, 1)
        tc_fence_after_sync();
        if (elect_one()) {
          // descriptor start-address field counts 16-byte units: +2 per 16-element K step inside the swizzle atom
          const uint64_t ad = adesc0 + (uint64_t)(sa * (uint32_t)(W_A_BYTES >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(s * (B_BYTES >> 4));
          if (real && !(p.dbg & 1)) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem_d, ad + 2 * k4, bd + 2 * k4, idesc, (it | k4) ? 1u : 0u);
          }
          if (cl > 1) umma_commit_mc(&b_empty[s], cmask);
          else        umma_commit(&b_empty[s]);
          if (real) umma_commit(&empty[sa]);
        }
        __syncwarp();
        SDB_TRACE(1, k, it, 2)
        if (++s == W_NSB) { s = 0; ph ^= 1; }
        if (real && ++sa == W_NSA) { sa = 0; pa ^= 1; }
      }
      if (real) {
        if (elect_one()) umma_commit(&acc_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; accp ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue: TMEM -> registers -> NCHW global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int hw = g.Ho * g.Wo;
    uint32_t acc = 0, accp = 0;
    for (int k = 0; k < niter; ++k) {
      const int work = SDB_TILE_OF(k);
      if (work >= p.num_tiles) continue;
      const int n = work / tiles_per_img, trem = work - n * tiles_per_img;
      const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
      const int r = q * 32 + lane;
      const int ho = ty * g.th + r / g.tw, wo = tx * g.tw + r % g.tw;
      const bool valid = ho < g.Ho && wo < g.Wo;
      const int rem = ho * g.Wo + wo;
      mbar_wait(&acc_full[acc], accp);
      tc_fence_after_sync();
      for (int c0 = 0; c0 < O; c0 += 32) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16) + c0, rr);
        tmem_ld_wait();
        if (valid && !(p.dbg & 4)) {
          const size_t d0 = ((size_t)n * O + c0) * hw + rem;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int o = c0 + j;
            if (o < O) {
              float v = __uint_as_float(rr[j]);
              if (p.bias) v += __ldg(p.bias + o);
              const size_t di = d0 + (size_t)j * hw;
              if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(p.out)[di] = __float2bfloat16_rn(v);
              else          reinterpret_cast<float*>(p.out)[di] = v;
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive_warp(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; accp ^= 1; }
    }
  } else if (warp == 6) {
    // ===== window producer: one tensor-map copy per (tile, 64-channel chunk) into the single window buffer =====
    if (elect_one()) {
      uint32_t cnt = 0;
      for (int k = 0; k < niter; ++k) {
        const int work = SDB_TILE_OF(k);
        if (work >= p.num_tiles) continue;
        const int n = work / tiles_per_img, trem = work - n * tiles_per_img;
        const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
        const int wy0 = ty * g.th * g.sh - g.ph - p.R, wx0 = tx * g.tw * g.sw - g.pw - p.R;
        for (int ch = 0; ch < nchunks; ++ch, ++cnt) {
          mbar_wait(&w_empty[0], (cnt & 1) ^ 1);   // every gather warp is done with the previous chunk's window
          if (p.dbg & 8) {
            mbar_arrive(&w_full[0]);
          } else {
            mbar_arrive_expect_tx(&w_full[0], p.win_bytes);
            tma_load_4d(smem_u32(sWin), &p.tmap, ch * 64, wx0, wy0, n, &w_full[0]);
          }
        }
      }
    }
  } else {
    // ===== gather warps: bilinear sampling from the staged window into the swizzled A stage =====
    // One continuous stream of warp iterations (4 pixels x 64 channels each) over (chunk, tap, iteration)
    // through a 4-slot register ring: the corner loads of the NEXT stage are issued while this stage is
    // interpolated and stored, so out-of-window samples (global loads, L2 latency) are a full stage ahead
    // of their use instead of stalling all eight warps at every stage.
    // Each warp owns 16 pixels of every stage and runs at its own pace: with four A buffers the `empty` wait of
    // a stage is for MMAs issued four stages ago, so the warps drift apart and the shared-memory pipe stays
    // busy while individual warps wait / fence / arrive (in lock-step on two buffers those ~700 clk per stage
    // were dead time for everybody, profiles/r1_win_trace.md).
    const int pw = warp - W_FIRST_PW, r0 = pw * W_PIXW;
    const int rg = r0;
    const int grp = lane / W_LPP, lig = lane % W_LPP;
    // each lane owns two 16-byte chunks of its pixel's 128-byte row: chunks (lig, lig + 4), taken in opposite
    // order by odd pixels so the two pixels of a quarter-warp never hit the same banks in one LDS / STS
    const int ca = lig + 4 * (grp & 1), cb = lig + 4 * ((grp & 1) ^ 1);
    const uint32_t pitch = (uint32_t)p.BW * 128u;
    const uint4* xg = reinterpret_cast<const uint4*>(p.xp);
    const uint32_t c16 = (uint32_t)(g.C / 8);
    const uint32_t win0 = smem_u32(sWin);
    const uint32_t desc0 = smem_u32(sDesc) + (uint32_t)(rg + grp) * 16u;
    int tcount = 0;   // tiles done by this CTA: stage / chunk counters below continue across tiles
    for (int k = 0; k < niter; ++k) {
      const int work = SDB_TILE_OF(k);
      if (work >= p.num_tiles) continue;
      const int n = work / tiles_per_img, trem = work - n * tiles_per_img;
      const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
      const int wy0 = ty * g.th * g.sh - g.ph - p.R, wx0 = tx * g.tw * g.sw - g.pw - p.R;
      __syncwarp();   // every lane is done with the previous tile's descriptors (each warp builds and reads its own rows)
      // (1) sampling descriptors of this warp's pixels for every tap: window byte offset of corner
      //     (y0, x0) -- or, flagged, (y0, x0) itself when a corner is outside the window -- and the
      //     four bilinear weights (x mask) as bf16
      {
        constexpr int ROUNDS = (W_PIXW * 16 + 31) / 32;
        float rdy[ROUNDS], rdx[ROUNDS], rm[ROUNDS];
        const int hwo = g.Ho * g.Wo;
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
          const int i = lane + rd * 32;
          const int px = i % W_PIXW, tap = i / W_PIXW;
          const int r = r0 + px;
          const int ho = ty * g.th + r / g.tw, wo = tx * g.tw + r % g.tw;
          rdy[rd] = 0.f; rdx[rd] = 0.f; rm[rd] = 1.f;
          if (tap < taps && ho < g.Ho && wo < g.Wo) {
            const float* o = p.off + ((size_t)n * 2 * taps + 2 * tap) * hwo + ho * g.Wo + wo;
            rdy[rd] = __ldg(o);
            rdx[rd] = __ldg(o + hwo);
            if (p.mask) rm[rd] = __ldg(p.mask + ((size_t)n * taps + tap) * hwo + ho * g.Wo + wo);
          }
        }
        {   // pull the next tile's offsets into L2 while this tile is gathered
          const int nwork = SDB_TILE_OF(k + 1);
          if (nwork < p.num_tiles) {
            const int nn = nwork / tiles_per_img, ntrem = nwork - nn * tiles_per_img;
            const int nty = ntrem / p.tiles_x, ntx = ntrem - nty * p.tiles_x;
            for (int i = lane; i < W_PIXW * taps; i += 32) {
              const int px = i % W_PIXW, tap = i / W_PIXW;
              const int r = r0 + px;
              const int ho = nty * g.th + r / g.tw, wo = ntx * g.tw + r % g.tw;
              if (ho < g.Ho && wo < g.Wo && (px & 7) == 0) {
                const float* o = p.off + ((size_t)nn * 2 * taps + 2 * tap) * hwo + ho * g.Wo + wo;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(o + hwo));
                if (p.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + ((size_t)nn * taps + tap) * hwo + ho * g.Wo + wo));
              }
            }
          }
        }
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
          const int i = lane + rd * 32;
          const int px = i % W_PIXW, tap = i / W_PIXW;
          if (tap < taps) {
            const int r = r0 + px;
            const int ho = ty * g.th + r / g.tw, wo = tx * g.tw + r % g.tw;
            uint4 d = make_uint4(0u, 0u, 0u, 0u);
            if (ho < g.Ho && wo < g.Wo) {
              const int ki = tap / g.KW, kj = tap - ki * g.KW;
              const float h = (float)(ho * g.sh - g.ph + ki * g.dh) + rdy[rd];
              const float w = (float)(wo * g.sw - g.pw + kj * g.dw) + rdx[rd];
              if (h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W) {
                const float m = rm[rd];
                const int y0 = (int)floorf(h), x0 = (int)floorf(w);
                const float lh = h - y0, lw = w - x0, hh = 1.f - lh, hw_ = 1.f - lw;
                const bool t = y0 >= 0, b = y0 + 1 <= g.H - 1, l = x0 >= 0, rt = x0 + 1 <= g.W - 1;
                const float w00 = (t && l) ? hh * hw_ * m : 0.f, w01 = (t && rt) ? hh * lw * m : 0.f;
                const float w10 = (b && l) ? lh * hw_ * m : 0.f, w11 = (b && rt) ? lh * lw * m : 0.f;
                d.y = pack_bf16x2(w00, w01);
                d.z = pack_bf16x2(w10, w11);
                const int ry = y0 - wy0, rx = x0 - wx0;
                if (ry >= 0 && ry + 1 < p.BH && rx >= 0 && rx + 1 < p.BW) d.x = (uint32_t)(ry * p.BW + rx) * 128u;
                else d.x = 0x80000000u | ((uint32_t)(y0 + 1) << 16) | (uint32_t)(x0 + 1);
              }
            }
            sDesc[tap * TILE_M + r] = d;
          }
        }
        __syncwarp();
      }
      // (2) the gather stream
      uint4 v[W_RING][8];   // [slot][corner * 2 + chunk]
      uint32_t wa[W_RING], wbv[W_RING];
      // issue the four corner loads of iteration it_ of stage (tap_, ch_), whose window buffer starts at wbase_
#define SDB_WDESC(tap_, it_) lds128(desc0 + (uint32_t)((tap_) * TILE_M + (it_) * W_PPI) * 16u)
#define SDB_WISSUE(dd_, ch_, slot_, wbase_)                                                      \
      {                                                                                          \
        wa[slot_] = dd_.y;                                                                       \
        wbv[slot_] = dd_.z;                                                                      \
        if (!(dd_.x >> 31)) {                                                                    \
          const uint32_t a_ = (wbase_) + dd_.x + ca * 16, b_ = (wbase_) + dd_.x + cb * 16;       \
          v[slot_][0] = lds128(a_);                 v[slot_][1] = lds128(b_);                    \
          v[slot_][2] = lds128(a_ + 128);           v[slot_][3] = lds128(b_ + 128);              \
          v[slot_][4] = lds128(a_ + pitch);         v[slot_][5] = lds128(b_ + pitch);            \
          v[slot_][6] = lds128(a_ + pitch + 128);   v[slot_][7] = lds128(b_ + pitch + 128);      \
        } else {   /* a corner outside the staged window: the four corners come from global memory */ \
          const int y0_ = (int)((dd_.x >> 16) & 0x7fffu) - 1, x0_ = (int)(dd_.x & 0xffffu) - 1;  \
          const int ya_ = max(y0_, 0), yb_ = min(y0_ + 1, g.H - 1), xa_ = max(x0_, 0), xb_ = min(x0_ + 1, g.W - 1); \
          const uint32_t rowa_ = (uint32_t)(n * g.H + ya_) * (uint32_t)g.W, rowb_ = (uint32_t)(n * g.H + yb_) * (uint32_t)g.W; \
          const uint4* xb2_ = xg + (ch_) * 8;                                                    \
          const uint4* p0_ = xb2_ + (size_t)(rowa_ + xa_) * c16;                                 \
          const uint4* p1_ = xb2_ + (size_t)(rowa_ + xb_) * c16;                                 \
          const uint4* p2_ = xb2_ + (size_t)(rowb_ + xa_) * c16;                                 \
          const uint4* p3_ = xb2_ + (size_t)(rowb_ + xb_) * c16;                                 \
          v[slot_][0] = __ldg(p0_ + ca); v[slot_][1] = __ldg(p0_ + cb);                          \
          v[slot_][2] = __ldg(p1_ + ca); v[slot_][3] = __ldg(p1_ + cb);                          \
          v[slot_][4] = __ldg(p2_ + ca); v[slot_][5] = __ldg(p2_ + cb);                          \
          v[slot_][6] = __ldg(p3_ + ca); v[slot_][7] = __ldg(p3_ + cb);                          \
        }                                                                                        \
      }
      const int gs0 = tcount * nstages, gc0 = tcount * nchunks;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(&w_full[0], (uint32_t)(gc0 + ch) & 1u);   // this chunk's window has landed
        // all warps leave this wait together and have identical per-stage timelines, so without help they
        // hit the shared-memory pipe in the same phase and leave it idle in the same phase: stagger them
        if (p.stagger_ns > 0 && (pw & 3)) __nanosleep((unsigned)(p.stagger_ns * (pw & 3)));
        {
#pragma unroll
          for (int u = 0; u < W_RING; ++u) {
            const uint4 dd = SDB_WDESC(0, u);
            SDB_WISSUE(dd, ch, u, win0)
          }
        }
        bool rdy = false;   // readiness of this stage's A buffer, tested (non-blocking) one stage ahead
        for (int tap = 0; tap < taps; ++tap) {
          const int st = ch * taps + tap;
          const uint32_t gs = (uint32_t)(gs0 + st), sa = gs % W_NSA, use = gs / W_NSA;
          const bool has_next = tap + 1 < taps;   // the ring does not cross a chunk: the next window is not there yet
          if (pw == 0) SDB_TRACE(0, k, st, 0)
          if (!rdy) mbar_wait(&empty[sa], (use & 1) ^ 1);
          if (pw == 0) SDB_TRACE(0, k, st, 1)
          // the gather, not the MMAs, paces the kernel, so the next buffer is almost always free already: ask
          // now, look at the answer after this stage's work (a ready try_wait costs ~190 clk on the critical path)
          rdy = has_next && mbar_test(&empty[(gs + 1) % W_NSA], (((gs + 1) / W_NSA) & 1) ^ 1);
          uint8_t* dst = sA + (size_t)sa * W_A_BYTES;
          if (!(p.dbg & 16))
#pragma unroll
          for (int it = 0; it < W_ITERS; ++it) {
            const int sl = it % W_RING;
            const uint32_t w00 = __byte_perm(wa[sl], wa[sl], 0x1010), w01 = __byte_perm(wa[sl], wa[sl], 0x3232);
            const uint32_t w10 = __byte_perm(wbv[sl], wbv[sl], 0x1010), w11 = __byte_perm(wbv[sl], wbv[sl], 0x3232);
            uint4 a, b;
            a.x = bf2_fma(w11, v[sl][6].x, bf2_fma(w10, v[sl][4].x, bf2_fma(w01, v[sl][2].x, bf2_mul(w00, v[sl][0].x))));
            a.y = bf2_fma(w11, v[sl][6].y, bf2_fma(w10, v[sl][4].y, bf2_fma(w01, v[sl][2].y, bf2_mul(w00, v[sl][0].y))));
            a.z = bf2_fma(w11, v[sl][6].z, bf2_fma(w10, v[sl][4].z, bf2_fma(w01, v[sl][2].z, bf2_mul(w00, v[sl][0].z))));
            a.w = bf2_fma(w11, v[sl][6].w, bf2_fma(w10, v[sl][4].w, bf2_fma(w01, v[sl][2].w, bf2_mul(w00, v[sl][0].w))));
            b.x = bf2_fma(w11, v[sl][7].x, bf2_fma(w10, v[sl][5].x, bf2_fma(w01, v[sl][3].x, bf2_mul(w00, v[sl][1].x))));
            b.y = bf2_fma(w11, v[sl][7].y, bf2_fma(w10, v[sl][5].y, bf2_fma(w01, v[sl][3].y, bf2_mul(w00, v[sl][1].y))));
            b.z = bf2_fma(w11, v[sl][7].z, bf2_fma(w10, v[sl][5].z, bf2_fma(w01, v[sl][3].z, bf2_mul(w00, v[sl][1].z))));
            b.w = bf2_fma(w11, v[sl][7].w, bf2_fma(w10, v[sl][5].w, bf2_fma(w01, v[sl][3].w, bf2_mul(w00, v[sl][1].w))));
            if (!(p.dbg & 2)) {
              const uint32_t row = rg + it * W_PPI + grp;
              *reinterpret_cast<uint4*>(dst + sw128_offset(row, ca)) = a;
              *reinterpret_cast<uint4*>(dst + sw128_offset(row, cb)) = b;
            }
            if (has_next) {
              const uint4 dd = SDB_WDESC(tap + 1, it);
              SDB_WISSUE(dd, ch, sl, win0)
            }
          }
          if (pw == 0) SDB_TRACE(0, k, st, 2)
          fence_proxy_async_smem();
          mbar_arrive_warp(&full[sa]);
          if (pw == 0) SDB_TRACE(0, k, st, 3)
        }
        __syncwarp();   // every load from this chunk's window has been consumed
        if (lane == 0) mbar_arrive(&w_empty[0]);
      }
      ++tcount;
#undef SDB_WISSUE
#undef SDB_WDESC
    }
  }
#undef SDB_TILE_OF
  tc_fence_before_sync();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
    else
      cudaGetLastError();
  }
  return fn;
}

struct WinPlan {
  bool ok;
  int R, BH, BW;
  uint32_t win_bytes;
  size_t smem;
};
constexpr size_t W_SMEM_LIMIT = 232448 - 1024;   // 227 KB opt-in limit minus the static barriers

WinPlan plan_window(const Geo& g) {
  WinPlan w{};
  w.ok = false;
  // Opt-in (SDB_TC_WIN=1): on the RepPoints head shapes this kernel is still ~20 % slower than the L2-gather
  // kernel in dcn_tc.cu (profiles/r1_win_trace.md has the stage timelines and what bounds it).
  const char* en = getenv("SDB_TC_WIN");
  if (!en || atoi(en) == 0) return w;
  if (g.H >= 32767 || g.W >= 65535) return w;
  if (g.th * g.tw != TILE_M) return w;
  const size_t b_bytes = (size_t)g.O * 128;
  const size_t fixed = 1024 + (size_t)W_NSA * W_A_BYTES + (size_t)W_NSB * b_bytes + (size_t)g.taps() * TILE_M * 16;
  const int span_y = (g.th - 1) * g.sh + (g.KH - 1) * g.dh + 2, span_x = (g.tw - 1) * g.sw + (g.KW - 1) * g.dw + 2;
  int rmax = 8;
  if (const char* e = getenv("SDB_TC_WIN_R")) rmax = atoi(e);
  for (int R = rmax; R >= 1; --R) {
    const int BH = span_y + 2 * R, BW = span_x + 2 * R;
    if (BH > 256 || BW > 256) continue;
    const size_t wbytes = (size_t)BH * BW * 128;
    if (fixed + wbytes <= W_SMEM_LIMIT) {
      w.ok = true;
      w.R = R; w.BH = BH; w.BW = BW; w.win_bytes = (uint32_t)wbytes;
      w.smem = fixed + wbytes;
      return w;
    }
  }
  return w;
}

}  // namespace

bool tc_win_supported(const Geo& g) { return encode_tiled_fn() != nullptr && plan_window(g).ok; }

// forward through the window-staged kernel; xp = NHWC bf16 input (already packed), wimg/bias32 = workspace
int tc_forward_win(const __nv_bfloat16* xp, const float* off, const float* mask, const void* w, const void* bias,
                   uint8_t* wimg, float* bias32, void* out, const Geo& g, int io_dtype, cudaStream_t st) {
  const WinPlan pl = plan_window(g);
  SDB_REQUIRE(pl.ok, SDB_ERR_UNSUPPORTED, "window kernel does not fit this geometry");
  EncodeTiledFn enc = encode_tiled_fn();
  SDB_REQUIRE(enc != nullptr, SDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");

  const long long wtotal = (long long)g.O * g.taps() * (g.C / 8);
  const int wblocks = (int)((wtotal + 255) / 256 < 1184 ? (wtotal + 255) / 256 : 1184);
  if (io_dtype == SDB_F32)
    prep_weight_win_kernel<float><<<wblocks, 256, 0, st>>>((const float*)w, (const float*)bias, wimg, bias32, g.O, g.C, g.taps(), 1, g.O);
  else
    prep_weight_win_kernel<__nv_bfloat16><<<wblocks, 256, 0, st>>>((const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, wimg, bias32, g.O, g.C, g.taps(), 1, g.O);
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());

  WinParams p{};
  {
    const cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    const cuuint64_t strides[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
    const cuuint32_t box[4] = {64u, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xp, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDB_REQUIRE(r == CUDA_SUCCESS, SDB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  }
  p.xp = xp; p.off = off; p.mask = mask; p.wimg = wimg; p.bias = bias ? bias32 : nullptr; p.out = out; p.g = g;
  p.tiles_x = cdiv(g.Wo, g.tw); p.tiles_y = cdiv(g.Ho, g.th);
  p.num_tiles = g.N * p.tiles_x * p.tiles_y;
  p.R = pl.R; p.BH = pl.BH; p.BW = pl.BW; p.win_bytes = pl.win_bytes;
  if (const char* e = getenv("SDB_TC_DEBUG")) p.dbg = atoi(e);
  if (p.num_tiles == 0) return SDB_OK;
  p.stagger_ns = 0;   // measured: no effect (the shared-memory queue re-forms the convoy), kept as an experiment switch
  if (const char* e = getenv("SDB_TC_WIN_STAGGER")) p.stagger_ns = atoi(e);
  p.cl = 2;
  if (const char* e = getenv("SDB_TC_WIN_CL")) p.cl = atoi(e) == 1 ? 1 : 2;
  if (p.num_tiles < 2) p.cl = 1;
  int grid = (p.num_tiles + p.cl - 1) / p.cl * p.cl;
  const int max_grid = num_sms() / p.cl * p.cl;
  if (grid > max_grid) grid = max_grid;
  const bool obf = io_dtype == SDB_BF16;
  void (*kern)(const WinParams) = obf ? dcn_fwd_win_kernel<true> : dcn_fwd_win_kernel<false>;
  SDB_ENSURE_SMEM(kern, pl.smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(W_NTHREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(SDB_OP_FORWARD, st);
  SDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // namespace sdb

// scratch tooling (not part of the public ABI): copy the SDB_TC_DEBUG&64 clock trace to the host
extern "C" int sdb_debug_read_trace(unsigned long long* dst, int n) {
  const size_t bytes = sizeof(unsigned long long) * (size_t)(n < 3 * 2 * 64 * 4 ? n : 3 * 2 * 64 * 4);
  return (int)cudaMemcpyFromSymbol(dst, sdb::g_trace, bytes);
}
