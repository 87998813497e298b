// dcn_tc_bwd.cu -- backward of the deformable convolution on tcgen05 tensor cores (SDB_MATH_BF16).
//
// backward_data has two halves, neither of which uses atomics on tensors:
//  (1) grad_offset / grad_mask -- one fused persistent kernel:
//   dcol[p, (tap,c)] = sum_o dY[p,o] W[o,c,tap]        tcgen05 GEMM, M=128 pixels, N=128 channels,
//                                                      K = C_out, accumulator in TMEM (never in HBM)
//   4 drain warps move each 128x128 fp32 accumulator to a bf16 staging tile in smem (padded rows);
//   reduce warps (16 lanes x 8 channels per pixel) re-read the 4 input corners of (pixel, tap) and
//   reduce  sum_c dcol * d(bilinear)/d(y|x)  and  sum_c dcol * bilinear  over channels (packed bf16 dot
//   products, mixed-precision FMAs, a transposing shuffle reduction) -> grad_offset / grad_mask, plain
//   full-width stores (one owner per (pixel, tap)).
//  (2) grad_input -- the reference scatters w * dcol with fp32 atomics (K3/K6); here kernel (1) writes its
//   bf16 dcol staging tiles to HBM with one bulk copy each (the only pass over dY W^T: no second GEMM) and
//   grad_input is a gather of them over a transposed sampling index: dcn_tc_dx.cu.
//   Replaces G2 + K2/K5 + K3/K6 of the reference (deform_conv_cuda.cu:553-559,
//   deform_conv_cuda_kernel.cu:291-452, :870-1066).
//
// backward_weight: dW[o, c, tap] = sum_p dY[p,o] col[p,(tap,c)]  -- tcgen05 GEMM with BOTH operands
//   MN-major, K = pixels, accumulators stationary in TMEM for a CTA's whole pixel range, split partials
//   reduced (and permuted to [O][C][kH][kW]) by a second small kernel -- deterministic, no atomics.
//   Two variants: over the columns the forward pass saved (dcn_wgrad_col_tc_kernel: both operands stream in
//   by bulk copy, O x 256 accumulators) or, when they were not saved, re-sampling them with the forward's
//   gather producer (dcn_bwd_weight_tc_kernel).  Replaces K1' + G3 of the reference (deform_conv_cuda.cu:738-778).
//
// Order of a backward call (tc_backward_all): layout packs -> fork: transposed-index build on the side stream || weight-
// gradient GEMM -> its split reduce on the side stream || grad_offset kernel -> grad_input gather -> join.  The grad_offset
// kernel (and the forward) are persistent with a static tile schedule and run ALONE: any co-running kernel that holds SMs
// delays the whole launch.  grad_offset and the forward run as CTA pairs (cta_group::2, M = 256) by default.
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

// channels per CTA of the weight-gradient kernel over saved columns (UMMA N, <= 256; O x CG fp32 must fit in TMEM)
inline int wgrad_col_group(const Geo& g) { return g.C % 256 == 0 ? 256 : nch_of(g); }

// ------------------------------------------------------------------------------------------------
// packing kernels
// ------------------------------------------------------------------------------------------------
// dY [N][O][HWo] (T) -> image of 128-pixel tiles: tile t, o-block kb -> [128 rows][64 o] bf16,
// 128B-swizzled (K-major for dgrad's A operand, MN-major for wgrad's A operand).  Zero padded.
// one entry per problem: grid.x walks the tiles of all entries (TileMap), grid.y the 64-o blocks
struct PackGyTable {
  TileMap map;
  struct E { const void* gy; uint8_t* img; Dims d; int fast; } e[MAX_PROBS];
  Geo g;
};
template <typename T>
__device__ __forceinline__ void pack_gy_body(const PackGyTable& t, int ei, int okb, float (*s)[129]) {
  const Geo g = with_dims(t.g, t.e[ei].d);
  const T* __restrict__ gy = (const T*)t.e[ei].gy;
  uint8_t* __restrict__ img = t.e[ei].img;
  const int O = g.O, hw = g.Ho * g.Wo;
  const int tile = blockIdx.x - t.map.start[ei], kb = blockIdx.y;
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
template <typename T>
__global__ void __launch_bounds__(256) pack_gy_kernel(const __grid_constant__ PackGyTable t, int okb) {
  __shared__ float s[64][129];
  pack_gy_body<T>(t, find_range(t.map, blockIdx.x), okb, s);
}

// bf16 source whose row segments are 8-byte aligned (Wo, tw, Ho*Wo multiples of 4): a lane loads FOUR pixels of
// one channel with one 8-byte load -- a whole 128-pixel tile row per warp instruction, 8 per lane instead of 32
// two-byte loads -- and one pixel decode per thread.  Entries that do not qualify (`fast` == 0) take the generic body
// in the same launch.
__global__ void __launch_bounds__(256) pack_gy_bf16v_kernel(const __grid_constant__ PackGyTable t, int okb) {
  __shared__ __align__(16) float s_gen[64][129];
  const int ei = find_range(t.map, blockIdx.x);
  if (!t.e[ei].fast) {
    pack_gy_body<__nv_bfloat16>(t, ei, okb, s_gen);
    return;
  }
  const Geo g = with_dims(t.g, t.e[ei].d);
  const __nv_bfloat16* __restrict__ gy = (const __nv_bfloat16*)t.e[ei].gy;
  uint8_t* __restrict__ img = t.e[ei].img;
  const int O = g.O, hw = g.Ho * g.Wo;
  __nv_bfloat16 (*s)[136] = reinterpret_cast<__nv_bfloat16(*)[136]>(s_gen);   // [o][pixel (column ^ 32 for o >= 32)], 8 spare columns
  const int tile = blockIdx.x - t.map.start[ei], kb = blockIdx.y;
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

// grad_bias[o] += scale * sum over the problems of one weight tensor of sum_{n,h,w} dY; grid (O, weights)
struct BiasGradTable {
  int n;
  struct E { const void* gy; int N, hw, weight; } e[MAX_PROBS];
  float* gb[MAX_WEIGHTS];
};
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const __grid_constant__ BiasGradTable t, float scale, int O) {
  const int o = blockIdx.x, wid = blockIdx.y;
  if (!t.gb[wid]) return;
  float sum = 0.f;
  for (int e = 0; e < t.n; ++e) {
    if (t.e[e].weight != wid) continue;
    const T* gy = (const T*)t.e[e].gy;
    const int hw = t.e[e].hw;
    for (int n = 0; n < t.e[e].N; ++n)
      for (int i = threadIdx.x; i < hw; i += blockDim.x) sum += to_f32(gy[((size_t)n * O + o) * hw + i]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tt = 0.f;
    for (int i = 0; i < 8; ++i) tt += part[i];
    atomicAdd(t.gb[wid] + o, scale * tt);
  }
}

// ------------------------------------------------------------------------------------------------
// backward data kernel
// ------------------------------------------------------------------------------------------------
struct DgradProb {
  const __nv_bfloat16* xp;   // NHWC bf16 input
  const float* off;
  const float* mask;
  const uint8_t* gy_img;     // dY tiles
  const uint8_t* wt_img;     // W^T tiles of this problem's convolution
  float* goff;               // [N][2*taps][HWo] fp32, or nullptr
  float* gmask;              // [N][taps][HWo] fp32, or nullptr
  uint8_t* dcol;             // export of the bf16 dcol staging tiles [tile][tap][chunk][128][NCH] for grad_input, or nullptr
  Dims d;
};
struct DgradParams {
  TileMap map;               // tiles of all problems
  DgradProb pr[MAX_PROBS];
  Geo g;                     // common geometry
  int nsb, okb;
};

// Warp roles are aligned to warpgroups so that the register file can be re-split with setmaxnreg: warps 0-3 = bulk
// producer, MMA issuer, dcol exporter and one idle warp (40 registers each), warps 4-7 = TMEM drain (96), warps 8-15 = the reduce
// warps (184), which are the critical path of this kernel: at the 128 registers a 512-thread launch gives everybody
// ptxas could not overlap the dependent shuffle / FMA chains of two iterations and the drain warps waited for them
// 80 % of the time (profiles/r2_dcn_kernels_ncu_v1.txt).
constexpr int G_DRAIN0 = 4, G_SW0 = 8, G_THREADS = (G_SW0 + NSW) * 32;
template <int R>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

// acc += (s2.lo + s2.hi) * c, c = low / high bf16 half of c2: two mixed-precision FMAs (FHFMA.BF16), fp32 accumulation
__device__ __forceinline__ void fma_pair_lo(float& acc, uint32_t s2, uint32_t c2) {
  asm("{ .reg .b16 sl, sh, cl, ch;\n mov.b32 {sl, sh}, %1;\n mov.b32 {cl, ch}, %2;\n"
      "fma.rn.f32.bf16 %0, sl, cl, %0;\n fma.rn.f32.bf16 %0, sh, cl, %0;\n }" : "+f"(acc) : "r"(s2), "r"(c2));
}
__device__ __forceinline__ void fma_pair_hi(float& acc, uint32_t s2, uint32_t c2) {
  asm("{ .reg .b16 sl, sh, cl, ch;\n mov.b32 {sl, sh}, %1;\n mov.b32 {cl, ch}, %2;\n"
      "fma.rn.f32.bf16 %0, sl, ch, %0;\n fma.rn.f32.bf16 %0, sh, ch, %0;\n }" : "+f"(acc) : "r"(s2), "r"(c2));
}

// PAIR: clusters of two CTAs (cta_group::2, M = 256): a work item is a pair of tiles of one problem; each CTA loads its own
// dY tile and HALF of every W^T tile (NCH/2 rows), the leader's MMA warp issues for both, the peer's MMA warp relays its
// CTA's a_full / b_full completions; everything downstream of TMEM (drain, reduce, export) stays per CTA.
template <int NCH, bool PAIR>
__global__ void __launch_bounds__(G_THREADS, 1) dcn_bwd_data_tc_kernel(const __grid_constant__ DgradParams p) {
  constexpr int LPB = NCH / 8, PPI = 32 / LPB;
  constexpr uint32_t B_TILE = NCH * 128;             // one [NCH c][64 o] weight tile of the image
  constexpr uint32_t B_BYTES = PAIR ? B_TILE / 2 : B_TILE;   // what this CTA loads of it = slot size
  constexpr uint32_t STG_BYTES = stg_tile_bytes(NCH);   // bf16 staging tile (padded rows)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, a_empty;
  __shared__ __align__(8) uint64_t b_full[MAX_B_STAGES], b_empty[MAX_B_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], stg_full[2], stg_empty[2];
  __shared__ __align__(8) uint64_t peer_a, peer_b[MAX_B_STAGES];   // PAIR, leader: the peer's dY tile / weight slot is full
  __shared__ uint32_t tmem_base_s;

  const int C = p.g.C, taps = p.g.KH * p.g.KW, nch = C / NCH, units = taps * nch, okb = p.okb;
  const int num_tiles = p.map.start[p.map.n];   // PAIR: number of tile PAIRS
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int work0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
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
      mbar_init(&acc_empty[s], PAIR ? 8 : 4);   // one arrival per drain warp (of both CTAs)
      mbar_init(&stg_full[s], 4);
      mbar_init(&stg_empty[s], NSW + 1);  // one arrival per reduce warp + the exporter
    }
    if (PAIR) {
      mbar_init(&peer_a, 1);
      for (int s = 0; s < p.nsb; ++s) mbar_init(&peer_b[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc2(&tmem_base_s, NCOLS);
    else tmem_alloc(&tmem_base_s, NCOLS);
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp < G_DRAIN0) {
    setmaxnreg_dec<40>();
  }
  if (warp == 0) {
    // ===== bulk producer: dY tile once per tile, W^T tiles per (tap, chunk, o-block) =====
    if (lane == 0) {
      uint32_t bs = 0, bp = 0, ap = 0;
      for (int work = work0; work < num_tiles; work += wstep) {
        const int pi = find_range(p.map, work);
        const DgradProb& pr = p.pr[pi];
        int tile = PAIR ? 2 * (work - p.map.start[pi]) + rank : work - p.map.start[pi];
        // PAIR: the all-invalid second tile of an odd problem re-reads the first one (its results are never stored)
        if (PAIR && (long long)tile * TILE_M >= (long long)pr.d.N * pr.d.Ho * pr.d.Wo) --tile;
        mbar_wait(&a_empty, ap ^ 1);
        mbar_arrive_expect_tx(&a_full, A_BYTES);
        bulk_g2s(sA, pr.gy_img + (size_t)tile * A_BYTES, A_BYTES, &a_full);
        ap ^= 1;
        const uint8_t* wsrc = pr.wt_img + (size_t)rank * B_BYTES;   // PAIR: rows [rank * NCH/2, ...) of every tile
        for (int i = 0; i < units * okb; ++i) {
          mbar_wait(&b_empty[bs], bp ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
          bulk_g2s(sB + (size_t)bs * B_BYTES, wsrc + (size_t)i * B_TILE, B_BYTES, &b_full[bs]);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: dcol[128 x NCH] = dY_tile[128 x O] * W^T (PAIR: the leader, for both CTAs' tiles) =====
    const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * TILE_M : TILE_M, NCH, 0, 0);
    uint32_t bs = 0, bp = 0, acc = 0, accp = 0, ap = 0;
    if (PAIR && rank != 0) {
      // the peer's bulk copies complete on its own barriers: relay them to the leader
      for (int work = work0; work < num_tiles; work += wstep) {
        mbar_wait(&a_full, ap);
        ap ^= 1;
        if (lane == 0) mbar_arrive_remote(&peer_a, 0);
        for (int i = 0; i < units * okb; ++i) {
          mbar_wait(&b_full[bs], bp);
          if (lane == 0) mbar_arrive_remote(&peer_b[bs], 0);
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
      }
    } else
    for (int work = work0; work < num_tiles; work += wstep) {
      mbar_wait(&a_full, ap);
      if (PAIR) mbar_wait_cluster(&peer_a, ap);
      ap ^= 1;
      for (int u = 0; u < units; ++u) {
        if (PAIR) mbar_wait_cluster(&acc_empty[acc], accp ^ 1);
        else mbar_wait(&acc_empty[acc], accp ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * NCH;
        for (int kb = 0; kb < okb; ++kb) {
          mbar_wait(&b_full[bs], bp);
          if (PAIR) mbar_wait_cluster(&peer_b[bs], bp);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + kb * (TILE_M * 128);
            const uint32_t b_addr = smem_base + A_BYTES + bs * B_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (PAIR)
                umma_bf16_pair(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32, 16, 1024),
                               make_smem_desc_sw128(b_addr + k4 * 32, 16, 1024), idesc, (kb | k4) != 0);
              else
                umma_bf16(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32, 16, 1024),
                          make_smem_desc_sw128(b_addr + k4 * 32, 16, 1024), idesc, (kb | k4) != 0);
            }
            if (PAIR) umma_commit_pair(&b_empty[bs]);
            else umma_commit(&b_empty[bs]);
          }
          __syncwarp();
          if (++bs == (uint32_t)p.nsb) { bs = 0; bp ^= 1; }
        }
        if (elect_one()) {
          if (PAIR) umma_commit_pair(&acc_full[acc]);
          else umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; accp ^= 1; }
      }
      if (elect_one()) {
        if (PAIR) umma_commit_pair(&a_empty);
        else umma_commit(&a_empty);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===== dcol exporter: every finished staging tile goes to HBM as one bulk copy (the grad_input gather reads it) =====
    if (lane == 0) {
      uint32_t sb = 0, sp = 0;
      for (int work = work0; work < num_tiles; work += wstep) {
        const int pi = find_range(p.map, work);
        uint8_t* dst = p.pr[pi].dcol;
        const int tile = PAIR ? 2 * (work - p.map.start[pi]) + rank : work - p.map.start[pi];
        if (PAIR && (long long)tile * TILE_M >= (long long)p.pr[pi].d.N * p.pr[pi].d.Ho * p.pr[pi].d.Wo) dst = nullptr;
        if (dst) dst += (size_t)tile * units * STG_BYTES;
        for (int u = 0; u < units; ++u) {
          mbar_wait(&stg_full[sb], sp);
          if (dst) {
            bulk_s2g(dst + (size_t)u * STG_BYTES, sS + (size_t)sb * STG_BYTES, STG_BYTES);
            bulk_commit();
            bulk_wait_read_all();
          }
          mbar_arrive(&stg_empty[sb]);
          if (++sb == 2) { sb = 0; sp ^= 1; }
        }
      }
      bulk_wait_all();
    }
  } else if (warp < G_DRAIN0) {
    // idle warp of the first warpgroup (takes part in the register re-split and the final barrier only)
  } else if (warp < G_SW0) {
    // ===== drain: TMEM accumulator -> bf16 staging tile (row = pixel, swizzled 16-byte chunks) =====
    setmaxnreg_dec<96>();
    const int q = warp & 3;
    const uint32_t row = q * 32 + lane;
    uint32_t acc = 0, accp = 0;
    for (int work = work0; work < num_tiles; work += wstep) {
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
        fence_proxy_async_smem();            // the exporter's bulk copy reads the tile through the async proxy
        if (PAIR && rank != 0) {             // TMEM buffer may be overwritten: the pair's buffer is released on the leader
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
        } else {
          mbar_arrive_warp(&acc_empty[acc]);
        }
        mbar_arrive_warp(&stg_full[acc]);    // staging tile is ready (release)
        if (++acc == 2) { acc = 0; accp ^= 1; }
      }
    }
  } else {
    // ===== reduce warps: grad_offset / grad_mask = channel reductions of dcol against the corners =====
    // Same load pipeline as the forward gather: per (pixel, tap) a descriptor in shared memory (corner offsets +
    // the twelve bf16 coefficients with which the four corner dot products enter d/dy, d/dx and d/dmask, zero for a
    // corner outside the image; built one tap ahead by lanes 0..15), and a 4-slot register ring that keeps 16
    // sixteen-byte corner loads per lane in flight across unit boundaries.  Per (pixel, tap, 8 channels): four
    // packed-bf16 dot products S_k = <dcol, corner_k> (16 HFMA2, two bf16 partial sums each) and 24 mixed-precision
    // FMAs (bf16 partial sum x bf16 coefficient + fp32 accumulator) into this lane's share of the three quantities,
    // then a transposing shuffle reduction over the pixel's lane group.  The sums over channel chunks stay in
    // registers; one full-width store per quantity at the end of the tap.
    setmaxnreg_inc<184>();
    constexpr int ITERS = PIX_PER_WARP / PPI;
    constexpr int RING = 4;
    static_assert(ITERS % RING == 0, "ring must divide the per-unit iteration count");
    __shared__ uint4 s_od[NSW][2][PIX_PER_WARP][3];   // {off[4]}, {ay01, ay23, ax01, ax23}, {w01, w23, -, -}: bf16 pairs
    __shared__ int2 s_px[NSW][PIX_PER_WARP];          // (n, ho*Wo+wo) of the warp's pixels, n = -1 when padded
    const int sw = warp - G_SW0, r0 = sw * PIX_PER_WARP;
    const int grp = lane / LPB, lig = lane % LPB;
    uint32_t sb = 0, sp = 0;
    for (int work = work0; work < num_tiles; work += wstep) {
      const int pi = find_range(p.map, work);
      const DgradProb& pr = p.pr[pi];
      const int tile = PAIR ? 2 * (work - p.map.start[pi]) + rank : work - p.map.start[pi];
      if (!pr.goff && !pr.gmask) {   // problem that only wants dcol (grad_input without grad_offset): release the tiles
        for (int u = 0; u < units; ++u) {
          mbar_wait(&stg_full[sb], sp);
          mbar_arrive_warp(&stg_empty[sb]);
          if (++sb == 2) { sb = 0; sp ^= 1; }
        }
        continue;
      }
      const Geo g = with_dims(p.g, pr.d);
      const int hw = g.Ho * g.Wo;
      const uint4* xbase = reinterpret_cast<const uint4*>(pr.xp) + lig;
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
          // d(bilinear)/dh, /dw (get_coordinate_weight, deform_conv_cuda_kernel.cu:163-214) and the bilinear weights
          // themselves, as signed per-corner coefficients (x mask for the two coordinate gradients)
          const float lh = bs.lh, lw = bs.lw, m = bs.m;
          const float ay[4] = {-m * (1.f - lw), -m * lw, m * (1.f - lw), m * lw};
          const float ax[4] = {-m * (1.f - lh), m * (1.f - lh), -m * lh, m * lh};
          const float wk[4] = {(1.f - lh) * (1.f - lw), (1.f - lh) * lw, lh * (1.f - lw), lh * lw};
          uint32_t off4[4];
          float cy[4], cx[4], cw[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool on = bs.idx[k] >= 0;
            off4[k] = on ? (uint32_t)bs.idx[k] * (uint32_t)(C / 8) : 0u;
            cy[k] = on ? ay[k] : 0.f; cx[k] = on ? ax[k] : 0.f; cw[k] = on ? wk[k] : 0.f;
          }
          s_od[sw][tap_ & 1][lane][0] = make_uint4(off4[0], off4[1], off4[2], off4[3]);
          s_od[sw][tap_ & 1][lane][1] = make_uint4(pack_bf16x2(cy[0], cy[1]), pack_bf16x2(cy[2], cy[3]),
                                                   pack_bf16x2(cx[0], cx[1]), pack_bf16x2(cx[2], cx[3]));
          s_od[sw][tap_ & 1][lane][2] = make_uint4(pack_bf16x2(cw[0], cw[1]), pack_bf16x2(cw[2], cw[3]), 0u, 0u);
        }
      };
      build_desc(0, fetch_rawb(g, pr.off, pr.mask, valid, n, ho, wo, 0));
      RawB raw_next = fetch_rawb(g, pr.off, pr.mask, valid && taps > 1, n, ho, wo, 1);   // raw offsets run two taps ahead
      __syncwarp();
      // the pixel whose results this lane stores (see the transposing reduction below): its plane offsets
      const int2 my_px = s_px[sw][(lig >> 1) * PPI + grp];
      const int my_n = my_px.x;
      const size_t my_goff = (size_t)my_n * 2 * taps * hw + my_px.y, my_gmask = (size_t)my_n * taps * hw + my_px.y;
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
        raw_next = fetch_rawb(g, pr.off, pr.mask, valid && tap + 2 < taps, n, ho, wo, tap + 2);
        __syncwarp();
        // this lane's share of d/dy, d/dx and d/dmask of every pixel it serves, summed over the channel chunks in
        // registers: the cross-lane reduction runs once per (pixel, tap), not once per chunk
        float qA[ITERS], qB[ITERS], qC[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) qA[it] = qB[it] = qC[it] = 0.f;
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
            const uint4 cf = s_od[sw][tap & 1][px][1];
            const uint2 cw = *reinterpret_cast<const uint2*>(&s_od[sw][tap & 1][px][2]);
            const uint4 d = *reinterpret_cast<const uint4*>(stg + stg_offset<NCH>(r0 + px, lig));
            uint32_t S2[4];   // <dcol, corner_k> over this lane's 8 channels as two bf16 partial sums
#pragma unroll
            for (int k = 0; k < 4; ++k)
              S2[k] = bf2_fma(d.w, v[slot][k].w, bf2_fma(d.z, v[slot][k].z, bf2_fma(d.y, v[slot][k].y, bf2_mul(d.x, v[slot][k].x))));
            if (it + RING < ITERS) {
              SDB_ISSUE(tap, ch, it + RING, slot)
            } else if (has_next) {
              SDB_ISSUE(ntap, nchk, it + RING - ITERS, slot)
            }
            fma_pair_lo(qA[it], S2[0], cf.x); fma_pair_hi(qA[it], S2[1], cf.x);
            fma_pair_lo(qA[it], S2[2], cf.y); fma_pair_hi(qA[it], S2[3], cf.y);
            fma_pair_lo(qB[it], S2[0], cf.z); fma_pair_hi(qB[it], S2[1], cf.z);
            fma_pair_lo(qB[it], S2[2], cf.w); fma_pair_hi(qB[it], S2[3], cf.w);
            fma_pair_lo(qC[it], S2[0], cw.x); fma_pair_hi(qC[it], S2[1], cw.x);
            fma_pair_lo(qC[it], S2[2], cw.y); fma_pair_hi(qC[it], S2[3], cw.y);
          }
          mbar_arrive_warp(&stg_empty[sb]);
          if (++sb == 2) { sb = 0; sp ^= 1; }
        }
        // Transposing reduction over the LPB = 2 * ITERS lanes of a pixel group: every step sends half of the
        // remaining per-pixel values to the partner lane and keeps the other half, so after log2(ITERS) steps lane
        // `lig` holds ONE value per quantity -- that of pixel (lig >> 1) * PPI + grp -- summed over half the group, and
        // a last exchange with lane lig ^ 1 completes it: ITERS shuffles per quantity, results spread over all
        // lanes (one pixel per lane pair), so the stores below are one or two full-width instructions per tap.
        static_assert(LPB == 2 * ITERS, "transposing reduction needs two lanes per pixel of the group");
#pragma unroll
        for (int m = ITERS, nn = ITERS; m >= 2; m >>= 1, nn >>= 1) {
          const bool up = (lig & m) != 0;
#pragma unroll
          for (int j = 0; j < nn / 2; ++j) {
            const float sa = __shfl_xor_sync(0xffffffffu, up ? qA[j] : qA[j + nn / 2], m);
            const float sb_ = __shfl_xor_sync(0xffffffffu, up ? qB[j] : qB[j + nn / 2], m);
            const float sc = __shfl_xor_sync(0xffffffffu, up ? qC[j] : qC[j + nn / 2], m);
            qA[j] = (up ? qA[j + nn / 2] : qA[j]) + sa;
            qB[j] = (up ? qB[j + nn / 2] : qB[j]) + sb_;
            qC[j] = (up ? qC[j + nn / 2] : qC[j]) + sc;
          }
        }
        const float SA = qA[0] + __shfl_xor_sync(0xffffffffu, qA[0], 1);
        const float SB = qB[0] + __shfl_xor_sync(0xffffffffu, qB[0], 1);
        const float SC = qC[0] + __shfl_xor_sync(0xffffffffu, qC[0], 1);
        // even lane of a pair: d/dy and d/dmask, odd lane: d/dx of pixel (lig >> 1) * PPI + grp
        if (my_n >= 0) {
          if ((lig & 1) == 0) {
            if (pr.goff) pr.goff[my_goff + (size_t)(2 * tap) * hw] = SA;
            if (pr.gmask) pr.gmask[my_gmask + (size_t)tap * hw] = SC;
          } else if (pr.goff) {
            pr.goff[my_goff + (size_t)(2 * tap + 1) * hw] = SB;
          }
        }
      }
#undef SDB_ISSUE
    }
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc2(tmem_base, NCOLS);
    else tmem_dealloc(tmem_base, NCOLS);
  }
}

// ------------------------------------------------------------------------------------------------
// backward weight kernel
// ------------------------------------------------------------------------------------------------
struct WgradProb {
  const __nv_bfloat16* xp;
  const float* off;
  const float* mask;
  const uint8_t* gy_img;
  Dims d;
};
// The weight gradient of a convolution sums over ALL its problems (FPN levels): a CTA owns one (weight, tap,
// channel chunk, pixel split) and walks its share of the concatenated tile list of that weight's problems.
struct WgradParams {
  WgradProb pr[MAX_PROBS];
  TileMap wmap[MAX_WEIGHTS];         // per weight: tile ranges of its problems
  int pidx[MAX_WEIGHTS][MAX_PROBS];  // per weight: problem index of each range
  float* part[MAX_WEIGHTS];          // per weight: [splits][taps][O][C] fp32 partial sums
  int splits[MAX_WEIGHTS], tiles_per_split[MAX_WEIGHTS];
  int cta_start[MAX_WEIGHTS + 1];    // first CTA of each weight
  int nweights;
  Geo g;
  int okb, nsg, nsy;
};

template <int NCH>
__global__ void __launch_bounds__(BWD_THREADS, 1) dcn_bwd_weight_tc_kernel(const __grid_constant__ WgradParams p) {
  constexpr int LPB = NCH / 8;
  constexpr uint32_t G_BYTES = TILE_M * NCH * 2;   // gathered col tile [NCH/64 blocks][128 px][64 c]
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t g_full[4], g_empty[4], y_full[4], y_empty[4], acc_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ GDesc sdesc[NSW][2][PIX_PER_WARP];   // per gather warp, double buffered over tiles

  const int C = p.g.C, O = p.g.O, taps = p.g.KH * p.g.KW, nch = C / NCH, okb = p.okb, mh_n = okb / 2;
  const uint32_t Y_BYTES = (uint32_t)okb * (TILE_M * 128);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sG = sm;
  uint8_t* sY = sG + (size_t)p.nsg * G_BYTES;
  // warp index via shfl = provably warp-uniform: role branches and loop counters stay in uniform registers,
  // so tcgen05.mma takes its descriptors from the uniform datapath without an ELECT/R2UR waterfall per instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // work item of this CTA
  int wid = 0;
  while (wid + 1 < p.nweights && (int)blockIdx.x >= p.cta_start[wid + 1]) ++wid;
  const int lcta = blockIdx.x - p.cta_start[wid];
  const int nsplit = p.splits[wid];
  const int split = lcta % nsplit;
  const int ch = (lcta / nsplit) % nch;
  const int tap = lcta / (nsplit * nch);
  const TileMap& wm = p.wmap[wid];
  const int t0 = split * p.tiles_per_split[wid];
  const int t1 = min(wm.start[wm.n], t0 + p.tiles_per_split[wid]);
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
        const int r = find_range(wm, tile);
        const uint8_t* gy_img = p.pr[p.pidx[wid][r]].gy_img;
        mbar_wait(&y_empty[ys], yp ^ 1);
        mbar_arrive_expect_tx(&y_full[ys], Y_BYTES);
        bulk_g2s(sY + (size_t)ys * Y_BYTES, gy_img + (size_t)(tile - wm.start[r]) * Y_BYTES, Y_BYTES, &y_full[ys]);
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
          float4* dst = reinterpret_cast<float4*>(p.part[wid] + (((size_t)split * taps + tap) * O + o) * C + ch * NCH + c0);
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
    // problem of concatenated tile `tile_`, its geometry and this lane's pixel in it
    struct Loc { bool valid; int n, ho, wo, pi; };
    auto locate = [&](int tile_) {
      Loc L = {false, 0, 0, 0, 0};
      if (tile_ >= t1) return L;
      const int r = find_range(wm, tile_);
      L.pi = p.pidx[wid][r];
      const Geo g = with_dims(p.g, p.pr[L.pi].d);
      const long long pix = (long long)(tile_ - wm.start[r]) * TILE_M + r0 + lane;
      L.valid = lane < PIX_PER_WARP && pix < g.P();
      if (L.valid) decode_q(g, pix, L.n, L.ho, L.wo);
      return L;
    };
    auto fetch = [&](const Loc& L) {
      const Geo g = with_dims(p.g, p.pr[L.pi].d);
      return fetch_rawb(g, p.pr[L.pi].off, p.pr[L.pi].mask, L.valid, L.n, L.ho, L.wo, tap);
    };
    auto xptr = [&](int pi_) { return reinterpret_cast<const uint4*>(p.pr[pi_].xp + ch * NCH + lig * 8); };
    // forward-style descriptor (weights already x mask, zero outside) of this lane's pixel -> sdesc[buf_]
    auto build_desc = [&](const RawB raw_, const Loc& L, int buf_) {
      if (lane < PIX_PER_WARP) {
        const Geo g = with_dims(p.g, p.pr[L.pi].d);
        const BSample bs = make_bsample_raw(g, raw_, L.valid, L.n, L.ho, L.wo, tap);
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
    Loc loc = locate(t0);
    build_desc(fetch(loc), loc, t0 & 1);
    const uint4* x_cur = xptr(loc.pi);        // input of the tile being gathered
    loc = locate(t0 + 1);
    RawB raw = fetch(loc);                    // of tile t0 + 1
    const uint4* x_next = xptr(loc.pi);       // input of the tile after it (its loads are issued during this tile)
    __syncwarp();
    uint4 v[RING][4], wq[RING];
#define SDB_WISSUE(x16_, tile_, it_, slot_)                                                      \
    {                                                                                            \
      const GDesc* d_ = &sdesc[sw][(tile_) & 1][(it_) * PPI + grp];                              \
      const uint4 o_ = *reinterpret_cast<const uint4*>(d_->off);                                 \
      wq[slot_] = *reinterpret_cast<const uint4*>(d_->w2);                                       \
      v[slot_][0] = __ldg((x16_) + o_.x);                                                        \
      v[slot_][1] = __ldg((x16_) + o_.y);                                                        \
      v[slot_][2] = __ldg((x16_) + o_.z);                                                        \
      v[slot_][3] = __ldg((x16_) + o_.w);                                                        \
    }
    if (t0 < t1) {
#pragma unroll
      for (int u = 0; u < RING; ++u) SDB_WISSUE(x_cur, t0, u, u)
    }
    for (int tile = t0; tile < t1; ++tile) {
      const bool has_next = tile + 1 < t1;
      // descriptors of tile+1 (its raw offsets arrived during the previous tile), then raw offsets of tile+2
      __syncwarp();   // buffer (tile+1)&1 was read while tile-1 was gathered
      build_desc(raw, loc, (tile + 1) & 1);
      loc = locate(tile + 2);
      raw = fetch(loc);
      const uint4* x_after = xptr(loc.pi);
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
          SDB_WISSUE(x_cur, tile, it + RING, slot)
        } else if (has_next) {
          SDB_WISSUE(x_next, tile + 1, it + RING - ITERS, slot)
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(&g_full[gs]);
      if (++gs == (uint32_t)p.nsg) { gs = 0; gp ^= 1; }
      x_cur = x_next;
      x_next = x_after;
    }
#undef SDB_WISSUE
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// backward weight kernel over SAVED columns
// ------------------------------------------------------------------------------------------------
// When the forward pass saved its sampled columns (sdb_dcn_problem.columns: the A stages of dcn_fwd_tc_kernel, already
// in the 128B-swizzled operand layout), the weight gradient needs no sampling at all: dW[o, (tap, c)] = sum_p
// dY[p, o] col[p, (tap, c)] is a plain GEMM whose two operands stream in by bulk copy.  A CTA owns one (weight, tap,
// channel group of CG <= 256 channels, pixel split); the accumulators (O x CG fp32 = up to all 512 TMEM columns) stay
// in TMEM for the CTA's whole pixel range.  A stage holds HALF a tile (64 pixels = 4 K-steps): dY [okb][64 px][64 o] +
// col [CG/64][64 px][64 c], 8 KB per block, so three stages fit beside each other at O = C = 256.
struct WcolProb {
  const uint8_t* col;
  const uint8_t* gy_img;
};
struct WcolParams {
  WcolProb pr[MAX_PROBS];
  TileMap wmap[MAX_WEIGHTS];         // per weight: tile ranges of its problems
  int pidx[MAX_WEIGHTS][MAX_PROBS];  // per weight: problem index of each range
  float* part[MAX_WEIGHTS];          // per weight: [splits][taps][O][C] fp32 partial sums
  int splits[MAX_WEIGHTS], tiles_per_split[MAX_WEIGHTS];
  int cta_start[MAX_WEIGHTS + 1];
  int nweights;
  Geo g;
  int okb, cg, cps, ns;              // 64-o blocks of dY; channels per CTA; channels per saved A stage; stages
};
constexpr int WCOL_THREADS = 6 * 32;   // warps: 0 bulk producer, 1 mma, 2-5 epilogue
constexpr int WCOL_MAX_STAGES = 6;
constexpr uint32_t HALF_BLOCK = 64 * 128;   // bytes of one [64 px][64 ch] block

__global__ void __launch_bounds__(WCOL_THREADS, 1) dcn_wgrad_col_tc_kernel(const __grid_constant__ WcolParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[WCOL_MAX_STAGES], empty[WCOL_MAX_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;

  const int C = p.g.C, O = p.g.O, taps = p.g.KH * p.g.KW, okb = p.okb, mh_n = okb / 2, CG = p.cg;
  const int ncg = C / CG, cb_n = CG / 64;                      // channel groups; 64-channel blocks per group
  const int nchunks = C / p.cps, kbps = p.cps / 64;            // layout of the saved columns: [tile][chunk][tap][kbps][128][64]
  const uint32_t A_STAGE = (uint32_t)p.cps * TILE_M * 2;       // one saved A stage
  const uint32_t Y_BYTES = (uint32_t)okb * (TILE_M * 128);     // one dY tile
  const uint32_t S_Y = (uint32_t)okb * HALF_BLOCK, S_BYTES = S_Y + (uint32_t)cb_n * HALF_BLOCK;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  int wid = 0;
  while (wid + 1 < p.nweights && (int)blockIdx.x >= p.cta_start[wid + 1]) ++wid;
  const int lcta = blockIdx.x - p.cta_start[wid];
  const int nsplit = p.splits[wid];
  const int split = lcta % nsplit;
  const int cgi = (lcta / nsplit) % ncg;
  const int tap = lcta / (nsplit * ncg);
  const TileMap& wm = p.wmap[wid];
  const int t0 = split * p.tiles_per_split[wid];
  const int t1 = min(wm.start[wm.n], t0 + p.tiles_per_split[wid]);
  uint32_t ncols = 32;
  while ((int)ncols < mh_n * CG) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.ns; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
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
      uint32_t ss = 0, sp = 0;
      for (int tile = t0; tile < t1; ++tile) {
        const int r = find_range(wm, tile);
        const WcolProb& pr = p.pr[p.pidx[wid][r]];
        const int lt = tile - wm.start[r];
        const uint8_t* ysrc = pr.gy_img + (size_t)lt * Y_BYTES;
        const uint8_t* csrc = pr.col + ((size_t)lt * nchunks * taps + tap) * A_STAGE;
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&empty[ss], sp ^ 1);
          mbar_arrive_expect_tx(&full[ss], S_BYTES);
          uint8_t* dst = sm + (size_t)ss * S_BYTES;
          for (int kb = 0; kb < okb; ++kb)
            bulk_g2s(dst + kb * HALF_BLOCK, ysrc + (size_t)kb * (TILE_M * 128) + half * HALF_BLOCK, HALF_BLOCK, &full[ss]);
          for (int b = 0; b < cb_n; ++b) {
            const int cblk = cgi * cb_n + b;   // 64-channel block of C
            const uint8_t* src = csrc + (size_t)(cblk / kbps) * taps * A_STAGE + (size_t)(cblk % kbps) * (TILE_M * 128);
            bulk_g2s(dst + S_Y + b * HALF_BLOCK, src + half * HALF_BLOCK, HALF_BLOCK, &full[ss]);
          }
          if (++ss == (uint32_t)p.ns) { ss = 0; sp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // D[o, c] += sum_pixels dY[pix, o] * col[pix, c]; both operands MN-major, 64-element blocks HALF_BLOCK apart
    const uint32_t idesc = make_idesc_bf16(128, CG, 1, 1);
    uint32_t ss = 0, sp = 0, accumulate = 0;
    for (int i = 0; i < 2 * (t1 - t0); ++i) {
      mbar_wait(&full[ss], sp);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t y_addr = smem_base + ss * S_BYTES, c_addr = y_addr + S_Y;
        for (int s = 0; s < 4; ++s) {
          for (int mh = 0; mh < mh_n; ++mh)
            umma_bf16(tmem_base + mh * CG, make_smem_desc_sw128(y_addr + (2 * mh) * HALF_BLOCK + s * 2048, HALF_BLOCK, 1024),
                      make_smem_desc_sw128(c_addr + s * 2048, HALF_BLOCK, 1024), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(&empty[ss]);
      }
      __syncwarp();
      if (++ss == (uint32_t)p.ns) { ss = 0; sp ^= 1; }
    }
    if (elect_one()) umma_commit(&acc_full);
    __syncwarp();
  } else {
    // epilogue: partial dW tile -> workspace [split][tap][O][C]
    const int q = warp & 3;
    mbar_wait(&acc_full, 0);
    tc_fence_after_sync();
    for (int mh = 0; mh < mh_n; ++mh) {
      const int o = mh * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < CG; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + mh * CG + ((uint32_t)(q * 32) << 16) + c0, r);
        tmem_ld_wait();
        if (o < O && t1 > t0) {
          float4* dst = reinterpret_cast<float4*>(p.part[wid] + (((size_t)split * taps + tap) * O + o) * C + cgi * CG + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// grad_weight[o][c][tap] += scale * sum_split part[split][tap][o][c]; grid.y = weight
struct WgradReduceTable {
  const float* part[MAX_WEIGHTS];
  float* gw[MAX_WEIGHTS];
  int splits[MAX_WEIGHTS];
};
// One block per (o, 128-channel chunk): the split sums are formed with coalesced reads along c, transposed through
// shared memory and added to grad_weight [O][C][taps] as one contiguous run of 128 * taps floats (the direct write has
// a stride of `taps` floats between lanes: 9x the sectors).
__global__ void __launch_bounds__(128) wgrad_reduce_kernel(const __grid_constant__ WgradReduceTable t, float scale, int taps,
                                                           int O, int C) {
  __shared__ float s[128 * 16];   // taps <= 16
  const int wid = blockIdx.y;
  const float* __restrict__ part = t.part[wid];
  float* __restrict__ gw = t.gw[wid];
  const int splits = t.splits[wid];
  if (!gw || splits == 0) return;
  const int cchunks = (C + 127) / 128;
  const int o = blockIdx.x / cchunks, c0 = (blockIdx.x % cchunks) * 128;
  const int c = c0 + threadIdx.x;
  if (c < C)
    for (int tap = 0; tap < taps; ++tap) {
      float acc = 0.f;
      for (int sp = 0; sp < splits; ++sp) acc += part[(((size_t)sp * taps + tap) * O + o) * C + c];
      s[threadIdx.x * taps + tap] = acc;
    }
  __syncthreads();
  const int n = min(128, C - c0) * taps;
  float* dst = gw + ((size_t)o * C + c0) * taps;
  for (int i = threadIdx.x; i < n; i += 128) dst[i] += scale * s[i];
}

}  // namespace

// ---- fork / join onto a library-owned side stream -----------------------------------------------------------------
// The transposed-index build (a chain of seven small launches) depends only on the offsets and on grad_out, not on
// the grad_offset kernel, so it runs BESIDE that kernel on a side stream and joins before grad_input.  Event
// record / wait pairs are the fork-join pattern CUDA-graph capture understands, so the same code is captured into
// the caller's graph when the caller's stream is capturing.  One (stream, events) set per device, created lazily.
namespace {
struct Side {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;     // index build
  cudaEvent_t fork2 = nullptr, join2 = nullptr;   // weight-gradient split reduce
  cudaEvent_t fork3 = nullptr, join3 = nullptr;   // in-call weight-image preparation
};
// "this host thread has forked a weight preparation that no kernel has waited for yet": per thread, so that a call running on
// another thread (autograd's engine thread) cannot consume it; the events themselves may be shared -- they are all recorded
// on the one side stream, so waiting for a later record only waits for more
thread_local bool t_prep_pending = false;
Side* side_of_device() {
  static Side tab[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Side& d = tab[dev];
  if (!d.s) {
    if (cudaStreamCreateWithFlags(&d.s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.fork2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.join2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.fork3, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.join3, cudaEventDisableTiming) != cudaSuccess) {
      d.s = nullptr;
      cudaGetLastError();
      return nullptr;
    }
  }
  return &d;
}
}  // namespace

// In-call weight preparation (a call whose weights came without prepared images) runs on the side stream, beside the
// call's layout packs; the first kernel that reads the images waits for it (tc_prep_wait).
cudaStream_t tc_prep_begin(cudaStream_t st) {
  Side* side = side_of_device();
  if (!side) return st;
  if (cudaEventRecord(side->fork3, st) != cudaSuccess || cudaStreamWaitEvent(side->s, side->fork3, 0) != cudaSuccess) {
    cudaGetLastError();
    return st;
  }
  return side->s;
}
void tc_prep_end(cudaStream_t st, cudaStream_t used) {
  Side* side = side_of_device();
  if (!side || used == st) return;
  if (cudaEventRecord(side->join3, side->s) == cudaSuccess) t_prep_pending = true;
}
int tc_prep_wait(cudaStream_t st) {
  Side* side = side_of_device();
  if (!side || !t_prep_pending) return SDB_OK;
  t_prep_pending = false;
  SDB_CHECK_CUDA(cudaStreamWaitEvent(st, side->join3, 0));
  return SDB_OK;
}

// ---- workspace plan of a multi-problem call ------------------------------------------------------------------------
// Everything a call needs beyond its tensors lives in ONE caller-provided workspace, laid out as a pure function of
// the problem dimensions (so sdb_dcn_multi_workspace_bytes and the call agree):
//   per problem : xp (NHWC bf16 input, unless the caller passes x_packed), and for the backward gy_img (dY as
//                 swizzled tiles) + dcol (dY W^T as bf16 tiles, taps * C * 2 bytes per output pixel, written by the
//                 grad_offset kernel and read by the grad_input gather);
//   per group   : its slice of the transposed index (cnt / start share one key space, the entries one pool);
//   per weight  : the prepared operand images (unless the caller passes them) and the split-K partials of dW.
TcPlan tc_plan(const TcProblem* pb, int n, int nweights, const bool* have_prepared, const Geo& g, bool backward) {
  TcPlan P{};
  const int okb = okb_of(g);
  const bool conv = n > 0 && pb[0].off == nullptr;   // plain convolution: no dcol tiles, no transposed index
  P.conv = conv;
  size_t o = 0;
  for (int i = 0; i < n; ++i) {
    const Geo gi = with_dims(g, pb[i].d);
    P.xp_off[i] = o;
    if (!pb[i].xp) o = align_up(o + tc_packed_input_bytes(gi), 1024);
    if (backward) {
      const size_t tiles = (size_t)cdiv(gi.P(), TILE_M);
      P.gy_off[i] = o;  o = align_up(o + tiles * okb * (TILE_M * 128), 1024);
      P.dcol_off[i] = o;
      if (!conv) o = align_up(o + tiles * g.taps() * nch_chunks(g) * stg_tile_bytes(nch_of(g)), 1024);
      P.gyp_off[i] = o;
      if (conv) o = align_up(o + (size_t)gi.P() * g.O * 2, 1024);
    }
  }
  for (int w = 0; w < nweights; ++w) {
    P.prep_off[w] = o;
    if (!have_prepared[w] && !(conv && backward)) o = align_up(o + tc_prepared_weight_bytes(g), 1024);
    P.convw_off[w] = o;
    if (conv && backward) o = align_up(o + tc_prepared_weight_bytes(g), 1024);
  }
  if (backward) {
    // offset groups: canonical index = order of first appearance; problems with group < 0 are their own group
    int ng = 0;
    for (int i = 0; i < n; ++i) {
      int gi = -1;
      if (pb[i].group >= 0)
        for (int j = 0; j < i; ++j)
          if (pb[j].group == pb[i].group) { gi = P.group_of[j]; break; }
      if (gi < 0) { gi = ng; P.group_rep[ng++] = i; }
      P.group_of[i] = gi;
    }
    P.ngroups = ng;
    long long keys = 0, ents = 1;
    for (int k = 0; k < (conv ? 0 : ng); ++k) {
      const Geo gk = with_dims(g, pb[P.group_rep[k]].d);
      P.key_base[k] = keys;
      const long long pix_in = (long long)cdiv((long long)gk.N * gk.H * gk.W, TILE_M) * TILE_M;
      keys += pix_in * (g.taps() + 1);
      // every (output pixel, tap) reaches at most four input pixels; up to LIST_ALIGN - 1 pad entries per input pixel
      ents += 4 * gk.P() * g.taps() + (LIST_ALIGN - 1) * pix_in;
    }
    P.nkeys = keys;
    P.scan_blocks = cdiv(keys, SCAN_PER_BLOCK);
    if (!conv) {
      // cnt and the entry pool are cleared by ONE memset (pad entries must read as zero): keep them adjacent
      P.cnt_off = o;   o = align_up(o + (size_t)keys * 4, 1024);
      P.ent_cap = (long long)cdiv(ents, LIST_ALIGN) * LIST_ALIGN;
      P.ent_off = o;   o = align_up(o + (size_t)P.ent_cap * sizeof(CEntry), 1024);
      P.clear_bytes = o - P.cnt_off;
      P.start_off = o; o = align_up(o + (size_t)(keys + 1) * 4, 1024);
      P.bsum_off = o;  o = align_up(o + (size_t)P.scan_blocks * 4, 1024);
      P.blk_off = o;   o = align_up(o + (size_t)(P.ent_cap / LIST_ALIGN) * ENT_BLOCK_BYTES, 1024);
    }
    // weight-gradient split-K: all weights share the machine
    int tiles_w[MAX_WEIGHTS] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) tiles_w[pb[i].weight_id] += cdiv(with_dims(g, pb[i].d).P(), TILE_M);
    // two variants: CTAs per (tap, NCH-channel chunk) for the re-sampling kernel, per (tap, wgrad_col_group channels)
    // for the kernel over saved columns; the partials buffer is sized for the larger split count
    for (int v = 0; v < 2; ++v) {
      const int groups = v == 0 ? nch_chunks(g) : g.C / wgrad_col_group(g);
      int s = num_sms() / ((nweights > 0 ? nweights : 1) * g.taps() * groups);
      if (s < 1) s = 1;
      if (s > 64) s = 64;
      for (int w = 0; w < nweights; ++w) {
        int sw = s < tiles_w[w] ? s : tiles_w[w];
        if (sw < 1) sw = 1;
        const int tps = tiles_w[w] > 0 ? cdiv(tiles_w[w], sw) : 1;
        const int sp = tiles_w[w] > 0 ? cdiv(tiles_w[w], tps) : 0;   // no empty split
        if (v == 0) { P.tiles_per_split[w] = tps; P.splits[w] = sp; }
        else { P.tiles_per_split_col[w] = tps; P.splits_col[w] = sp; }
      }
    }
    for (int w = 0; w < nweights; ++w) {
      const int smax = P.splits[w] > P.splits_col[w] ? P.splits[w] : P.splits_col[w];
      P.part_off[w] = o;
      o = align_up(o + (size_t)smax * g.taps() * g.O * g.C * 4, 1024);
    }
  }
  P.total = o;
  return P;
}

// ---- forward: pack every input once, one kernel over all problems -----------------------------------------------
int tc_forward_all(TcProblem* pb, int n, const Geo& g, int io_dtype, cudaStream_t st) {
  PackJob jobs[MAX_PROBS];
  for (int i = 0; i < n; ++i) jobs[i] = PackJob{pb[i].x, pb[i].xp, pb[i].d.N, pb[i].d.H * pb[i].d.W};
  int rc = pack_nhwc_multi(jobs, n, g.C, g.C, io_dtype == SDB_BF16, st);
  if (rc) return rc;
  return tc_forward_multi(pb, n, g, io_dtype, st);
}

// CTA-pair grad_offset kernel (sdb_set_backward_pair)
int g_bwd_pair = 1;

// ---- backward: grad_offset / grad_mask, grad_input, grad_weight / grad_bias of all problems ------------------------
// dY is packed once per problem (tile image + NHWC rows) and serves the three kernels; the transposed index is built
// once per offset group; one launch per kernel over all problems.
int tc_backward_all(TcProblem* pb, int n, float* const* gw, float* const* gb, int nweights, const TcPlan& P, const Geo& g,
                    int io_dtype, float scale, bool pack_x, int accumulate_gx, bool grad_packed, uint8_t* base,
                    cudaStream_t st, int gather_phase) {
  const int NCH = nch_of(g), okb = okb_of(g);
  const bool bf = io_dtype == SDB_BF16;
  int rc;
  bool any_goff = false, any_gx = false, any_gw = false, any_gb = false;
  for (int i = 0; i < n; ++i) { any_goff |= pb[i].goff || pb[i].gmask; any_gx |= pb[i].gx != nullptr; }
  for (int w = 0; w < nweights; ++w) { any_gw |= gw[w] != nullptr; any_gb |= gb[w] != nullptr; }
  // gather_phase 3: weight gradients + the transposed index only (the index is built beside the weight-gradient GEMM; no
  // grad_offset kernel, no gather); 4: grad_offset + gather over the index a phase-3 call on the same table left behind
  const bool index_only = gather_phase == 3, index_ready = gather_phase == 4;
  if (index_only) any_goff = false;
  if (gather_phase == 2) {   // only the gather: dcol tiles and index were left in the workspace by a phase-1 call
    if (!any_gx) return SDB_OK;
    rc = tc_build_transposed_index(pb, n, P, g, base, st, true);
    if (rc) return rc;
    return tc_dx_multi(pb, n, g, io_dtype, accumulate_gx, st);
  }

  // (0) layouts: x -> NHWC (unless the forward exported it), dY -> tile image and NHWC rows
  if (pack_x && !grad_packed && (any_goff || any_gw)) {   // grad_input alone does not read x
    PackJob jobs[MAX_PROBS];
    for (int i = 0; i < n; ++i) jobs[i] = PackJob{pb[i].x, pb[i].xp, pb[i].d.N, pb[i].d.H * pb[i].d.W};
    rc = pack_nhwc_multi(jobs, n, g.C, g.C, bf, st);
    if (rc) return rc;
  }
  if (!grad_packed && (any_goff || any_gx || any_gw)) {
    PackGyTable t{};
    t.g = g;
    int m = 0, total = 0;
    for (int i = 0; i < n; ++i) {
      const Geo gi = with_dims(g, pb[i].d);
      const int hw = gi.Ho * gi.Wo;
      if (gi.P() == 0) continue;
      t.e[m].gy = pb[i].gy; t.e[m].img = pb[i].gy_img; t.e[m].d = pb[i].d;
      t.e[m].fast = bf && gi.Wo % 4 == 0 && g.tw % 4 == 0 && hw % 4 == 0 && ((size_t)pb[i].gy & 7) == 0;
      t.map.start[m] = total;
      total += cdiv(gi.P(), TILE_M);
      ++m;
    }
    t.map.n = m; t.map.start[m] = total;
    if (total > 0) {
      dim3 grid(total, okb);
      if (bf) pack_gy_bf16v_kernel<<<grid, 256, 0, st>>>(t, okb);
      else pack_gy_kernel<float><<<grid, 256, 0, st>>>(t, okb);
      SDB_LAUNCHED(1);
      SDB_CHECK_CUDA(cudaGetLastError());
    }
  }

  // (2a) transposed sampling index for grad_input, on the side stream.  Forked here -- after the layout packs, before the
  //      weight-gradient GEMM -- it runs beside that GEMM and not beside the statically scheduled grad_offset kernel, which
  //      every co-running kernel delays.  Measured (whole step): forked here 0.791 ms; forked before the packs 0.817 ms;
  //      forked after the GEMM (= beside grad_offset, the earlier arrangement) 0.828 ms.
  Side* side = (any_gx || any_gw) ? side_of_device() : nullptr;
  bool reduce_pending = false;
  cudaStream_t ist = st;   // stream of the index build
  if (side && any_gx) {
    SDB_CHECK_CUDA(cudaEventRecord(side->fork, st));
    SDB_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
    ist = side->s;
  }
  if (any_gx) {
    rc = tc_build_transposed_index(pb, n, P, g, base, ist, index_ready);
    if (rc) return rc;
    if (side) SDB_CHECK_CUDA(cudaEventRecord(side->join, side->s));
  }

  // (3) grad_weight (+ grad_bias) FIRST -- it needs only dY and the saved columns, its split reduce then runs beside the
  // data-gradient kernels, and the gather still follows grad_offset directly (it reads the L2-resident tail of dcol first) --: one CTA per (weight, tap, channel chunk, pixel split) over that weight's tiles
  bool all_col = any_gw;   // every problem that contributes to a wanted weight gradient brings its saved columns
  for (int i = 0; i < n; ++i)
    if (gw[pb[i].weight_id] && with_dims(g, pb[i].d).P() > 0 && !pb[i].col) all_col = false;
  if (any_gw && all_col) {
    WcolParams p{};
    p.g = g; p.okb = okb; p.nweights = nweights; p.cg = wgrad_col_group(g); p.cps = tc_lanes_per_pixel(g) * 8;
    int cta = 0, m = 0;
    int slot_of[MAX_PROBS];
    for (int i = 0; i < n; ++i) {
      slot_of[i] = -1;
      if (!gw[pb[i].weight_id] || with_dims(g, pb[i].d).P() == 0) continue;
      p.pr[m].col = pb[i].col; p.pr[m].gy_img = pb[i].gy_img;
      slot_of[i] = m++;
    }
    WgradReduceTable rt{};
    for (int w = 0; w < nweights; ++w) {
      TileMap& wm = p.wmap[w];
      int total = 0, r = 0;
      for (int i = 0; i < n; ++i) {
        if (pb[i].weight_id != w || slot_of[i] < 0) continue;
        p.pidx[w][r] = slot_of[i];
        wm.start[r++] = total;
        total += cdiv(with_dims(g, pb[i].d).P(), TILE_M);
      }
      wm.n = r; wm.start[r] = total;
      p.part[w] = (float*)(base + P.part_off[w]);
      p.tiles_per_split[w] = P.tiles_per_split_col[w];
      p.splits[w] = (gw[w] && total > 0) ? P.splits_col[w] : 0;
      p.cta_start[w] = cta;
      cta += g.taps() * (g.C / p.cg) * p.splits[w];
      rt.part[w] = p.part[w]; rt.gw[w] = gw[w]; rt.splits[w] = p.splits[w];
    }
    p.cta_start[nweights] = cta;
    if (cta > 0) {
      const size_t s_bytes = (size_t)(okb + p.cg / 64) * HALF_BLOCK;
      long long ns = ((long long)(208 * 1024) - 1024) / (long long)s_bytes;
      if (ns > WCOL_MAX_STAGES) ns = WCOL_MAX_STAGES;
      SDB_REQUIRE(ns >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for backward_weight");
      p.ns = (int)ns;
      const size_t smem = p.ns * s_bytes + 1024;
      SDB_ENSURE_SMEM(dcn_wgrad_col_tc_kernel, smem);
      {
        ProfScope prof(SDB_OP_BACKWARD_WEIGHT, st);
        dcn_wgrad_col_tc_kernel<<<cta, WCOL_THREADS, smem, st>>>(p);
        SDB_LAUNCHED(1);
      }
      SDB_CHECK_CUDA(cudaGetLastError());
      // the split reduce runs on the side stream (behind the index build), beside grad_offset / grad_input
      cudaStream_t rst = st;
      if (side) {
        SDB_CHECK_CUDA(cudaEventRecord(side->fork2, st));
        SDB_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork2, 0));
        rst = side->s;
      }
      wgrad_reduce_kernel<<<dim3(g.O * ((g.C + 127) / 128), nweights), 128, 0, rst>>>(rt, scale, g.taps(), g.O, g.C); SDB_LAUNCHED(1);
      if (side) { SDB_CHECK_CUDA(cudaEventRecord(side->join2, side->s)); reduce_pending = true; }
      SDB_CHECK_CUDA(cudaGetLastError());
    }
  } else if (any_gw) {
    WgradParams p{};
    p.g = g; p.okb = okb; p.nweights = nweights;
    int cta = 0, m = 0;
    int slot_of[MAX_PROBS];
    for (int i = 0; i < n; ++i) {
      slot_of[i] = -1;
      if (!gw[pb[i].weight_id] || with_dims(g, pb[i].d).P() == 0) continue;
      WgradProb& q = p.pr[m];
      q.xp = (const __nv_bfloat16*)pb[i].xp; q.off = pb[i].off; q.mask = pb[i].mask; q.gy_img = pb[i].gy_img; q.d = pb[i].d;
      slot_of[i] = m++;
    }
    WgradReduceTable rt{};
    for (int w = 0; w < nweights; ++w) {
      TileMap& wm = p.wmap[w];
      int total = 0, r = 0;
      for (int i = 0; i < n; ++i) {
        if (pb[i].weight_id != w || slot_of[i] < 0) continue;
        p.pidx[w][r] = slot_of[i];
        wm.start[r++] = total;
        total += cdiv(with_dims(g, pb[i].d).P(), TILE_M);
      }
      wm.n = r; wm.start[r] = total;
      p.part[w] = (float*)(base + P.part_off[w]);
      p.tiles_per_split[w] = P.tiles_per_split[w];
      p.splits[w] = (gw[w] && total > 0) ? P.splits[w] : 0;
      p.cta_start[w] = cta;
      cta += g.taps() * nch_chunks(g) * p.splits[w];
      rt.part[w] = p.part[w]; rt.gw[w] = gw[w]; rt.splits[w] = p.splits[w];
    }
    p.cta_start[nweights] = cta;
    if (cta > 0) {
      const size_t g_bytes = (size_t)TILE_M * NCH * 2, y_bytes = (size_t)okb * TILE_M * 128;
      p.nsy = 2;
      long long nsg = ((long long)(208 * 1024) - 1024 - (long long)(p.nsy * y_bytes)) / (long long)g_bytes;
      if (nsg > 4) nsg = 4;
      SDB_REQUIRE(nsg >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for backward_weight");
      p.nsg = (int)nsg;
      const size_t smem = p.nsg * g_bytes + p.nsy * y_bytes + 1024;
      if (NCH == 128) SDB_ENSURE_SMEM(dcn_bwd_weight_tc_kernel<128>, smem);
      else SDB_ENSURE_SMEM(dcn_bwd_weight_tc_kernel<64>, smem);
      {
        ProfScope prof(SDB_OP_BACKWARD_WEIGHT, st);
        if (NCH == 128) dcn_bwd_weight_tc_kernel<128><<<cta, BWD_THREADS, smem, st>>>(p);
        else dcn_bwd_weight_tc_kernel<64><<<cta, BWD_THREADS, smem, st>>>(p);
        SDB_LAUNCHED(1);
      }
      SDB_CHECK_CUDA(cudaGetLastError());
      // weights with no tile contribute nothing: the reduce of a weight with splits == 0 returns at once
      // the split reduce runs on the side stream (behind the index build), beside grad_offset / grad_input
      cudaStream_t rst = st;
      if (side) {
        SDB_CHECK_CUDA(cudaEventRecord(side->fork2, st));
        SDB_CHECK_CUDA(cudaStreamWaitEvent(side->s, side->fork2, 0));
        rst = side->s;
      }
      wgrad_reduce_kernel<<<dim3(g.O * ((g.C + 127) / 128), nweights), 128, 0, rst>>>(rt, scale, g.taps(), g.O, g.C); SDB_LAUNCHED(1);
      if (side) { SDB_CHECK_CUDA(cudaEventRecord(side->join2, side->s)); reduce_pending = true; }
      SDB_CHECK_CUDA(cudaGetLastError());
    }
  }
  // (1) grad_offset / grad_mask: dcol GEMM + channel reduction; the dcol tiles of the problems that want grad_input
  //     are exported for the gather (2b)
  if ((any_goff || any_gx) && !index_only) {
    DgradParams p{};
    p.g = g; p.okb = okb;
    int m = 0, total = 0;
    for (int i = 0; i < n; ++i) {
      if (!pb[i].goff && !pb[i].gmask && !pb[i].gx) continue;
      DgradProb& q = p.pr[m];
      q.xp = (const __nv_bfloat16*)pb[i].xp; q.off = pb[i].off; q.mask = pb[i].mask; q.gy_img = pb[i].gy_img;
      q.wt_img = pb[i].w.dgrad; q.goff = pb[i].goff; q.gmask = pb[i].mask ? pb[i].gmask : nullptr; q.d = pb[i].d;
      q.dcol = pb[i].gx ? pb[i].dcol : nullptr;
      p.map.start[m] = total;
      total += cdiv(with_dims(g, pb[i].d).P(), TILE_M);
      ++m;
    }
    p.map.n = m; p.map.start[m] = total;
    if (total > 0) {
      // CTA pairs (sdb_set_backward_pair): work items are PAIRS of tiles of one problem, half-size weight slots
      const bool pair = g_bwd_pair && NCH == 128;
      if (pair) {
        total = 0;
        for (int k = 0; k < m; ++k) {
          p.map.start[k] = total;
          total += cdiv(cdiv((long long)p.pr[k].d.N * p.pr[k].d.Ho * p.pr[k].d.Wo, TILE_M), 2);
        }
        p.map.start[m] = total;
      }
      const size_t a_bytes = (size_t)okb * TILE_M * 128, b_bytes = (size_t)NCH * 128 / (pair ? 2 : 1), stg = 2 * (size_t)stg_tile_bytes(NCH);
      long long nsb = ((long long)(208 * 1024) - 1024 - (long long)a_bytes - (long long)stg) / (long long)b_bytes;
      if (nsb > MAX_B_STAGES) nsb = MAX_B_STAGES;
      SDB_REQUIRE(nsb >= 2, SDB_ERR_UNSUPPORTED, "shared memory budget too small for backward_data");
      p.nsb = (int)nsb;
      const size_t smem = a_bytes + p.nsb * b_bytes + stg + 1024;
      rc = tc_prep_wait(st);   // weight images prepared in this call, on the side stream
      if (rc) return rc;
      {
        ProfScope prof(SDB_OP_BACKWARD_DATA, st);
        if (pair) {
          const int clusters = total < grid_sms() / 2 ? total : grid_sms() / 2;
          SDB_ENSURE_SMEM((dcn_bwd_data_tc_kernel<128, true>), smem);
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(G_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          SDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dcn_bwd_data_tc_kernel<128, true>, p));
        } else {
          const int grid = total < grid_sms() ? total : grid_sms();
          if (NCH == 128) {
            SDB_ENSURE_SMEM((dcn_bwd_data_tc_kernel<128, false>), smem);
            dcn_bwd_data_tc_kernel<128, false><<<grid, G_THREADS, smem, st>>>(p);
          } else {
            SDB_ENSURE_SMEM((dcn_bwd_data_tc_kernel<64, false>), smem);
            dcn_bwd_data_tc_kernel<64, false><<<grid, G_THREADS, smem, st>>>(p);
          }
        }
        SDB_LAUNCHED(1);
      }
      SDB_CHECK_CUDA(cudaGetLastError());
    }
  }

  // (2b) grad_input: gather of the dcol tiles over the transposed index
  if (any_gx) {
    if (side) SDB_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    if (gather_phase != 1 && !index_only) {
      rc = tc_dx_multi(pb, n, g, io_dtype, accumulate_gx, st);
      if (rc) return rc;
    }
  }

  if (any_gb) {
    BiasGradTable t{};
    int m = 0;
    for (int i = 0; i < n; ++i) {
      if (!gb[pb[i].weight_id]) continue;
      t.e[m].gy = pb[i].gy; t.e[m].N = pb[i].d.N; t.e[m].hw = pb[i].d.Ho * pb[i].d.Wo; t.e[m].weight = pb[i].weight_id;
      ++m;
    }
    t.n = m;
    for (int w = 0; w < nweights; ++w) t.gb[w] = gb[w];
    if (bf) bias_grad_kernel<__nv_bfloat16><<<dim3(g.O, nweights), 256, 0, st>>>(t, scale, g.O);
    else bias_grad_kernel<float><<<dim3(g.O, nweights), 256, 0, st>>>(t, scale, g.O);
    SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
  }
  if (reduce_pending) SDB_CHECK_CUDA(cudaStreamWaitEvent(st, side->join2, 0));   // the split reduce ran on the side stream
  return SDB_OK;
}

// ---- plain convolution (the towers' Conv2d 3x3, reppointsv2.py:644-675): offset == nullptr ------------------------------
// "same" convolutions only (stride 1, 2 * pad == dil * (k - 1)): grad_input is then the convolution of dY with the
// transposed, tap-reversed weights at the same geometry, which the forward kernel computes from weight image 2.
bool tc_conv_supported(const Geo& g, bool need_grad_input, const char** why) {
  if (!tc_supported(g, why)) return false;
  if (g.sh != 1 || g.sw != 1) { *why = "plain-convolution mode needs stride 1"; return false; }
  if (2 * g.ph != g.dh * (g.KH - 1) || 2 * g.pw != g.dw * (g.KW - 1)) { *why = "plain-convolution mode needs 'same' padding"; return false; }
  if (need_grad_input) {
    Geo t = g;
    t.C = g.O; t.O = g.C;
    if (!tc_supported(t, why)) return false;
  }
  return true;
}

int tc_conv_backward_all(TcProblem* pb, int n, const void* const* weights, float* const* gw, float* const* gb, int nweights,
                         const TcPlan& P, const Geo& g, int io_dtype, float scale, bool pack_x, uint8_t* base,
                         cudaStream_t st) {
  // (1) grad_weight / grad_bias: the generic path with the data gradients masked off (packs dY into tiles, then the GEMM
  //     over the saved columns -- or the re-sampling kernel, whose null offsets read as zero)
  void* gx[MAX_PROBS];
  bool any_gx = false;
  for (int i = 0; i < n; ++i) { gx[i] = pb[i].gx; any_gx |= gx[i] != nullptr; pb[i].gx = nullptr; pb[i].goff = nullptr; pb[i].gmask = nullptr; }
  int rc = tc_backward_all(pb, n, gw, gb, nweights, P, g, io_dtype, scale, pack_x, 0, false, base, st);
  for (int i = 0; i < n; ++i) pb[i].gx = gx[i];
  if (rc || !any_gx) return rc;
  // (2) grad_input = conv(dY, W'): dY -> NHWC bf16, weight image 2, forward kernel with C and O swapped
  Geo gt = g;
  gt.C = g.O; gt.O = g.C;
  PackJob jobs[MAX_PROBS];
  TcProblem q[MAX_PROBS];
  bool used[MAX_WEIGHTS] = {false, false, false, false};
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (!pb[i].gx || with_dims(g, pb[i].d).P() == 0) continue;
    used[pb[i].weight_id] = true;
    jobs[m] = PackJob{pb[i].gy, base + P.gyp_off[i], pb[i].d.N, pb[i].d.Ho * pb[i].d.Wo};
    q[m] = TcProblem{};
    q[m].d = Dims{pb[i].d.N, pb[i].d.Ho, pb[i].d.Wo, pb[i].d.H, pb[i].d.W};
    q[m].weight_id = pb[i].weight_id;
    q[m].xp = base + P.gyp_off[i];
    q[m].out = pb[i].gx;
    ++m;
  }
  if (m == 0) return SDB_OK;
  for (int w = 0; w < nweights; ++w) {
    if (!used[w]) continue;
    rc = tc_prepare_weights(weights[w], nullptr, g, io_dtype, base + P.convw_off[w], 4, st);
    if (rc) return rc;
  }
  for (int k = 0; k < m; ++k) {
    const TcWeightImages img = tc_weight_images(g, base + P.convw_off[q[k].weight_id], false);
    q[k].w.fwd = img.convt;
    q[k].w.bias = nullptr;
  }
  rc = pack_nhwc_multi(jobs, m, g.O, g.O, io_dtype == SDB_BF16, st);
  if (rc) return rc;
  return tc_forward_multi(q, m, gt, io_dtype, st);
}

}  // namespace sdb
