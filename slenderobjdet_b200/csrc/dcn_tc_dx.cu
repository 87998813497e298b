// dcn_tc_dx.cu -- grad_input of the tensor-core backward: the transposed sampling index and the gather over it.
//
// The reference scatters w * dcol into grad_input with fp32 atomics (deformable_col2im, deform_conv_cuda_kernel.cu:
// 291-376, :870-960).  Here the scatter is inverted once per offset group into a CSR index "which (output pixel, tap,
// weight) touch input pixel q" (count -> pad -> scan -> fill -> sort, ~36 entries per input pixel for 3x3), the
// grad_offset kernel (dcn_tc_bwd.cu) exports its bf16 dcol = dY W^T tiles to HBM, and grad_input becomes a pure gather
//   dX[q, c] = sum_{(p, tap, w) in list(q)} w * dcol[p, tap, c]
// with fp32 accumulation in a fixed (sorted) order: deterministic, no atomics, no fp32 NHWC accumulation buffer, no
// memset, no NHWC -> NCHW pass.  Own translation unit: the gather runs at the 64-register occupancy cliff and ptxas'
// allocation for it must not depend on unrelated kernels.
#include <type_traits>

#include "dcn_tc_shared.cuh"

namespace sdb {
namespace {
using namespace tc;
using namespace tcshared;

// ------------------------------------------------------------------------------------------------
// index of the transposed sampling pattern (grad_input as a gather, see file header)
// ------------------------------------------------------------------------------------------------
// key(q, tap) = q * (taps + 1) + tap, q = band-order position of the INPUT pixel (tile = q >> 7, row = q & 127).
// Plain CSR: the (output pixel p, weight w = bilinear x mask) pairs that reach q through `tap` are the entries
// start[key] .. start[key + 1]; because the keys of one input pixel are adjacent, the whole list of q -- every tap --
// is one contiguous run, which is what the gather walks.  The extra key q * (taps + 1) + taps holds 0..7 zero entries
// that pad the run to a multiple of eight entries (64 bytes), so the gather reads whole groups of eight with aligned
// 16-byte loads and no per-entry bounds test (a zero entry = row 0 with weight 0).  An entry is 8 bytes: the row of
// dcol it names, as an offset in 16-byte units into the problem's dcol tiles (channel chunk 0), and tap << 16 | bf16
// weight.  With stride 1 a key holds four entries on average (36 per input pixel for 3x3).
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
    f(q * (g.taps() + 1) + tap, wb);
  }
}

// One transposed index per OFFSET GROUP (problems that sample with the same offsets, e.g. the two DCNs of a
// RepPoints level, share it).  All groups live in one key space: group i owns keys [key_base, key_base + nkeys_i),
// so a single scan serves the whole call.  grid (blocks of 256 output pixels over all groups, taps).
struct CsrTable {
  TileMap map;   // blocks of 256 output pixels
  struct G { const float* off; const float* mask; Dims d; int key_base; } gr[MAX_PROBS];
  Geo g;
};
__global__ void __launch_bounds__(256) csr_count_kernel(const __grid_constant__ CsrTable t, int* __restrict__ cnt) {
  const int gi = find_range(t.map, blockIdx.x);
  const Geo g = with_dims(t.g, t.gr[gi].d);
  const int tap = blockIdx.y;
  const long long p = (long long)(blockIdx.x - t.map.start[gi]) * blockDim.x + threadIdx.x;
  if (p >= g.P()) return;
  const int hw = g.Ho * g.Wo;
  const int n = (int)(p / hw), r = (int)(p - (long long)n * hw);
  int* c = cnt + t.gr[gi].key_base;
  for_each_hit(g, t.gr[gi].off, t.gr[gi].mask, n, r / g.Wo, r % g.Wo, tap,
               [&](long long key, uint32_t) { atomicAdd(c + key, 1); });
}

// pad key of every input pixel: the number of zero entries that round its list up to LIST_ALIGN entries
__global__ void __launch_bounds__(256) csr_pad_kernel(int* __restrict__ cnt, int npix, int taps) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npix) return;
  int* c = cnt + (size_t)q * (taps + 1);
  int s = 0;
  for (int t = 0; t < taps; ++t) s += c[t];
  c[taps] = (-s) & (LIST_ALIGN - 1);
}

__global__ void __launch_bounds__(256) csr_block_sums_kernel(const int* __restrict__ cnt, int* __restrict__ bsum,
                                                             int nkeys) {
  const int base = blockIdx.x * SCAN_PER_BLOCK;
  int s = 0;
  for (int i = threadIdx.x; i < SCAN_PER_BLOCK; i += 256) {
    const int k = base + i;
    if (k < nkeys) s += cnt[k];
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
// start[k] = exclusive scan of the entry counts; start[nkeys] = total
__global__ void __launch_bounds__(256) csr_scan_final_kernel(const int* __restrict__ cnt, const int* __restrict__ bsum,
                                                             int* __restrict__ start, int nkeys) {
  __shared__ int wsum[8];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = bsum[blockIdx.x];
  __syncthreads();
  const int base = blockIdx.x * SCAN_PER_BLOCK;
  for (int i0 = 0; i0 < SCAN_PER_BLOCK; i0 += 256) {
    const int k = base + i0 + threadIdx.x;
    const int v = k < nkeys ? cnt[k] : 0;
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
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + wbase + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const __grid_constant__ CsrTable t, int* __restrict__ cnt_all,
                                                       const int* __restrict__ start_all, CEntry* __restrict__ ent) {
  const int gi = find_range(t.map, blockIdx.x);
  const Geo g = with_dims(t.g, t.gr[gi].d);
  const int tap = blockIdx.y;
  const long long p = (long long)(blockIdx.x - t.map.start[gi]) * blockDim.x + threadIdx.x;
  if (p >= g.P()) return;
  const int hw = g.Ho * g.Wo;
  const int n = (int)(p / hw), r = (int)(p - (long long)n * hw);
  const uint32_t pos = (uint32_t)encode_pos(g.Ho, g.Wo, g.th, g.tw, n, r / g.Wo, r % g.Wo);
  const uint32_t nchv = (uint32_t)nch_of(g), nchunks = (uint32_t)g.C / nchv;
  const uint32_t row16 = ((pos >> 7) * (uint32_t)g.taps() * nchunks + (uint32_t)tap * nchunks) * (stg_tile_bytes((int)nchv) / 16u) +
                         (pos & 127u) * stg_row_units((int)nchv);
  int* cnt = cnt_all + t.gr[gi].key_base;
  const int* start = start_all + t.gr[gi].key_base;
  for_each_hit(g, t.gr[gi].off, t.gr[gi].mask, n, r / g.Wo, r % g.Wo, tap, [&](long long key, uint32_t wb) {
    const int slot = atomicSub(cnt + key, 1) - 1;   // slots are handed out from the back
    CEntry e;
    e.row16 = row16;
    e.tw = ((uint32_t)tap << 16) | wb;
    ent[start[key] + slot] = e;
  });
}

// Canonical order.  csr_fill hands out list slots with atomics, so the order in which the fp32 sums of grad_input are
// formed would change from run to run.  This pass sorts the entries of every key by their output pixel (an output
// pixel reaches an (input pixel, tap) at most once, so the sort keys are unique): grad_input becomes
// bit-reproducible.  One thread per key; a key holds ~4 entries.
__global__ void __launch_bounds__(256) csr_sort_kernel(const int* __restrict__ start, CEntry* __restrict__ ent, int nkeys,
                                                       int taps) {
  const int key = blockIdx.x * blockDim.x + threadIdx.x;
  if (key >= nkeys || key % (taps + 1) == taps) return;   // pad keys hold zeros
  const int b = start[key], n = start[key + 1] - b;
  if (n < 2) return;
  unsigned long long* e = reinterpret_cast<unsigned long long*>(ent + b);   // row16 (monotonic in the output pixel) in the low word
  constexpr int CAP = 8;
  if (n <= CAP) {
    unsigned long long v[CAP];
#pragma unroll
    for (int i = 0; i < CAP; ++i) v[i] = i < n ? e[i] : ~0ull;
    // odd-even transposition sort on the low word (fully unrolled: registers, no local memory)
#pragma unroll
    for (int r = 0; r < CAP; ++r) {
#pragma unroll
      for (int i = r & 1; i + 1 < CAP; i += 2) {
        const bool sw = (uint32_t)v[i + 1] < (uint32_t)v[i] && v[i + 1] != ~0ull;
        const unsigned long long lo = sw ? v[i + 1] : v[i], hi = sw ? v[i] : v[i + 1];
        v[i] = lo; v[i + 1] = hi;
      }
    }
#pragma unroll
    for (int i = 0; i < CAP; ++i)
      if (i < n) e[i] = v[i];
    return;
  }
  for (int i = 1; i < n; ++i) {   // long list (many taps colliding on one input pixel): insertion sort in place
    const unsigned long long x = e[i];
    int j = i - 1;
    while (j >= 0 && (uint32_t)e[j] > (uint32_t)x) { e[j + 1] = e[j]; --j; }
    e[j + 1] = x;
  }
}

// Final form read by the gather: the sorted 8-byte entries regrouped into blocks of eight (8 x u32 dcol row, then 8 x
// bf16 weight: 48 bytes, dcn_tc_shared.cuh) -- three aligned 16-byte loads per batch in the gather and the weights
// already packed in pairs for its mixed-precision FMAs.  One thread per block; lists are padded to whole blocks with
// zero entries, so every block is written.
__global__ void __launch_bounds__(256) csr_pack_kernel(const CEntry* __restrict__ ent, uint4* __restrict__ blocks, int nblocks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint4* src = reinterpret_cast<const uint4*>(ent + (size_t)b * 8);   // two entries per uint4: (row16, tw, row16, tw)
  const uint4 e0 = src[0], e1 = src[1], e2 = src[2], e3 = src[3];
  uint4* dst = blocks + (size_t)b * 3;
  dst[0] = make_uint4(e0.x, e0.z, e1.x, e1.z);
  dst[1] = make_uint4(e2.x, e2.z, e3.x, e3.z);
  dst[2] = make_uint4((e0.y & 0xffffu) | (e0.w << 16), (e1.y & 0xffffu) | (e1.w << 16),
                      (e2.y & 0xffffu) | (e2.w << 16), (e3.y & 0xffffu) | (e3.w << 16));
}

// ------------------------------------------------------------------------------------------------
// grad_input: gather of the exported dcol tiles over the transposed index
// ------------------------------------------------------------------------------------------------
// dX[q, c] = sum over the list of q (all taps): w_e * dcol[p_e, tap_e, c], accumulated in fp32 in list order (sorted:
// bit-reproducible).  A CTA owns (128 input pixels in band order = a compact patch, one NCH-channel chunk); a group of
// LPB lanes serves one pixel with 8 channels per lane, so a list entry is one 16-byte load per lane and NCH*2
// contiguous bytes per group.  The group reads eight entries of its list with one coalesced load, broadcasts them with
// shuffles and issues the eight row loads together; latency is hidden by occupancy (one accumulator row of 8 floats
// per thread, <= 64 registers, 32 warps per SM), not by a software pipeline.  The four input pixels that share a dcol
// row (the four bilinear corners) sit in the same or a neighbouring warp and walk their lists in the same (tap,
// position) order, so the repeats are L1 / L2 hits.  HBM-bound on paper (dcol is read once: taps * C * 2 bytes per
// output pixel); the result leaves through a shared-memory transpose as NCHW rows.
struct DxProb {
  const uint8_t* dcol;
  const int* start;          // transposed index of the problem's offset group
  const uint32_t* ent;       // entry pool (blocks of eight, dcn_tc_shared.cuh)
  void* out;                 // NCHW grad_input (f32 or bf16)
  Dims d;
};
struct DxParams {
  TileMap map;               // work items: (input tile, channel chunk), chunk fastest
  DxProb pr[MAX_PROBS];
  Geo g;
  int accumulate;            // add to `out` (the single-call ABI accumulates into grad_x) instead of overwriting
};

// acc[0..7] += w * (8 bf16 of v), w = bf16 half HI (0 = low, 1 = high) of w2, fp32 accumulation: mixed-precision FMA
// (FHFMA.BF16) reads all bf16 halves in place
template <int HI>
__device__ __forceinline__ void fma8_bf16(float (&acc)[8], const uint4 v, uint32_t w2) {
#define SDB_FH(a0_, a1_, r_)                                                                   \
  if (HI)                                                                                      \
    asm("{ .reg .b16 lo, hi, wl, wh;\n mov.b32 {lo, hi}, %2;\n mov.b32 {wl, wh}, %3;\n"      \
        "fma.rn.f32.bf16 %0, lo, wh, %0;\n fma.rn.f32.bf16 %1, hi, wh, %1;\n }"               \
        : "+f"(a0_), "+f"(a1_) : "r"(r_), "r"(w2));                                            \
  else                                                                                         \
    asm("{ .reg .b16 lo, hi, wl, wh;\n mov.b32 {lo, hi}, %2;\n mov.b32 {wl, wh}, %3;\n"      \
        "fma.rn.f32.bf16 %0, lo, wl, %0;\n fma.rn.f32.bf16 %1, hi, wl, %1;\n }"               \
        : "+f"(a0_), "+f"(a1_) : "r"(r_), "r"(w2));
  SDB_FH(acc[0], acc[1], v.x) SDB_FH(acc[2], acc[3], v.y) SDB_FH(acc[4], acc[5], v.z) SDB_FH(acc[6], acc[7], v.w)
#undef SDB_FH
}

// 16-byte read-only load of row `row16` (16-byte units) behind this lane's base pointer, written so that the address
// is ONE wide multiply-add (IMAD.WIDE.U32) instead of a 64-bit add + shift chain
__device__ __forceinline__ uint4 ldg_row(const uint8_t* base, uint32_t row16) {
  return __ldg(reinterpret_cast<const uint4*>(base + (unsigned long long)row16 * 16ull));
}

template <int NCH, bool OUT_BF16, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 1024 / THREADS) dcn_dx_gather_kernel(const __grid_constant__ DxParams p) {
  constexpr int LPB = NCH / 8, PPI = 32 / LPB;
  constexpr int PIXW = TILE_M / (THREADS / 32), ROUNDS = PIXW / PPI;
  static_assert(ROUNDS >= 1 && LIST_ALIGN == 8, "list walk is written for groups of eight entries");
  constexpr size_t STG_BYTES = stg_tile_bytes(NCH);
  using ST = typename std::conditional<OUT_BF16, __nv_bfloat16, float>::type;
  extern __shared__ __align__(16) uint8_t s_raw[];
  ST* s_t = reinterpret_cast<ST*>(s_raw);  // [NCH][128] transpose buffer in the output type, column rotated by PPI * (c >> 3)
  __shared__ int2 s_px[TILE_M];            // (n, y*W + x) of the tile's pixels, n = -1 past the end

  // CTAs walk the work list BACKWARDS: the grad_offset kernel wrote dcol in forward order, so its last tiles are still
  // L2-resident when this kernel starts -- read that part first, before it is evicted (162 vs 165 us)
  const int work = p.map.start[p.map.n] - 1 - (int)blockIdx.x;
  const int pi = find_range(p.map, work);
  const DxProb& pr = p.pr[pi];
  const int C = p.g.C, taps = p.g.KH * p.g.KW, nch = C / NCH;
  const int local = work - p.map.start[pi];
  const int tile = local / nch, ch = local - tile * nch;
  const int H = pr.d.H, W = pr.d.W, hw = H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / LPB, lig = lane % LPB;
  const int r0 = warp * PIXW;

  if (threadIdx.x < TILE_M) {
    const long long q = (long long)tile * TILE_M + threadIdx.x;
    int n = -1, y = 0, x = 0;
    if (q < (long long)pr.d.N * hw) decode_pos(H, W, p.g.th, p.g.tw, q, n, y, x);
    s_px[threadIdx.x] = make_int2(n, y * W + x);
  }

  // list bounds of the tile's pixels (the first global-memory latency of every list, paid once per CTA)
  __shared__ int s_beg[TILE_M + 1];
  if (threadIdx.x <= TILE_M) s_beg[threadIdx.x] = __ldg(pr.start + ((size_t)tile * TILE_M + threadIdx.x) * (taps + 1));
  __syncthreads();

  // Pixel of (warp, round, lane group).  The pixels in flight at one time form a compact block of the 8 x 16 patch
  // (4 x 8 with 16 warps, 4 x 4 with 8): the four input pixels sharing a dcol row are 2 x 2 neighbours, so most of
  // them are in flight together and the repeats hit in L1.  Other shapes walk the tile linearly.
  auto pixel_of = [&](int r) {
    if (PPI == 2 && THREADS == 512) return ((r >> 1) * 4 + (warp >> 2)) * 16 + (r & 1) * 8 + (warp & 3) * 2 + grp;
    if (PPI == 2 && THREADS == 256) return ((r >> 2) * 4 + (warp >> 1)) * 16 + (r & 3) * 4 + (warp & 1) * 2 + grp;
    return r0 + r * PPI + grp;
  };
  // this lane's 16-byte column of the dcol rows of channel chunk `ch`
  const uint8_t* cb = pr.dcol + (size_t)ch * STG_BYTES + (size_t)lig * 16;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (int r = 0; r < ROUNDS; ++r) {
    const int px = pixel_of(r);
    const int beg = s_beg[px], nb = (s_beg[px + 1] - beg) >> 3;   // batches of eight entries
    const uint4* ep = reinterpret_cast<const uint4*>(pr.ent) + (size_t)(beg >> 3) * 3;  // blocks: rows 0-3, rows 4-7, 8 weights
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    uint4 ra = zero4, rb = zero4, wv = zero4;   // block of the batch being issued, fetched one batch ahead
    if (nb > 0) { ra = __ldg(ep); rb = __ldg(ep + 1); wv = __ldg(ep + 2); }
    for (int b = 0; b < nb; ++b) {
      uint4 v[8];
      // zero entries (padding) load row 0 and contribute nothing
      v[0] = ldg_row(cb, ra.x); v[1] = ldg_row(cb, ra.y); v[2] = ldg_row(cb, ra.z); v[3] = ldg_row(cb, ra.w);
      v[4] = ldg_row(cb, rb.x); v[5] = ldg_row(cb, rb.y); v[6] = ldg_row(cb, rb.z); v[7] = ldg_row(cb, rb.w);
      const uint4 w = wv;
      if (b + 1 < nb) { ra = __ldg(ep + 3 * (b + 1)); rb = __ldg(ep + 3 * (b + 1) + 1); wv = __ldg(ep + 3 * (b + 1) + 2); }
      fma8_bf16<0>(acc, v[0], w.x); fma8_bf16<1>(acc, v[1], w.x);
      fma8_bf16<0>(acc, v[2], w.y); fma8_bf16<1>(acc, v[3], w.y);
      fma8_bf16<0>(acc, v[4], w.z); fma8_bf16<1>(acc, v[5], w.z);
      fma8_bf16<0>(acc, v[6], w.w); fma8_bf16<1>(acc, v[7], w.w);
    }
    // [pixel][channel] registers -> transpose buffer
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int si = (lig * 8 + j) * TILE_M + ((px + PPI * lig) & (TILE_M - 1));
      if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(s_t)[si] = __float2bfloat16_rn(acc[j]);
      else reinterpret_cast<float*>(s_t)[si] = acc[j];
    }
  }
  __syncthreads();
  // NCHW rows: a warp store = 32 consecutive tile pixels of one channel
  {
    const int px = threadIdx.x & (TILE_M - 1);
    const int2 pxy = s_px[px];
    if (pxy.x >= 0) {
      const size_t o0 = ((size_t)pxy.x * C + (size_t)ch * NCH) * hw + pxy.y;
      for (int c = threadIdx.x >> 7; c < NCH; c += THREADS / TILE_M) {
        const int si = c * TILE_M + ((px + PPI * (c >> 3)) & (TILE_M - 1));
        const size_t di = o0 + (size_t)c * hw;
        if (OUT_BF16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(pr.out) + di;
          __nv_bfloat16 v = reinterpret_cast<const __nv_bfloat16*>(s_t)[si];
          if (p.accumulate) v = __float2bfloat16_rn(__bfloat162float(v) + __bfloat162float(*o));
          *o = v;
        } else {
          float* o = reinterpret_cast<float*>(pr.out) + di;
          float v = reinterpret_cast<const float*>(s_t)[si];
          if (p.accumulate) v += *o;
          *o = v;
        }
      }
    }
  }
}

}  // namespace

// the transposed sampling index (one per offset group) of a call, on stream `st`
int tc_build_transposed_index(TcProblem* pb, int n, const TcPlan& P, const Geo& g, uint8_t* base, cudaStream_t st,
                              bool pointers_only) {
  int* cnt = (int*)(base + P.cnt_off);
  int* start = (int*)(base + P.start_off);
  int* bsum = (int*)(base + P.bsum_off);
  CEntry* ent = (CEntry*)(base + P.ent_off);
  uint4* blocks = (uint4*)(base + P.blk_off);
  if (pointers_only) {   // the index of a previous call on the same table is still in the workspace (SDB_BWD_GATHER_ONLY)
    for (int i = 0; i < n; ++i) {
      pb[i].start = start + P.key_base[P.group_of[i]];
      pb[i].ent = blocks;
    }
    return SDB_OK;
  }
  CsrTable t{};
  t.g = g;
  int total = 0, m = 0;
  for (int k = 0; k < P.ngroups; ++k) {
    bool wanted = false;
    for (int i = 0; i < n; ++i) wanted |= P.group_of[i] == k && pb[i].gx;
    if (!wanted) continue;
    const TcProblem& r = pb[P.group_rep[k]];
    t.gr[m].off = r.off; t.gr[m].mask = r.mask; t.gr[m].d = r.d; t.gr[m].key_base = (int)P.key_base[k];
    t.map.start[m] = total;
    total += cdiv(with_dims(g, r.d).P(), 256);
    ++m;
  }
  t.map.n = m; t.map.start[m] = total;
  const int nkeys = (int)P.nkeys;
  SDB_CHECK_CUDA(cudaMemsetAsync(cnt, 0, P.clear_bytes, st));   // cnt and the entry pool in one fill
  dim3 hgrid(total, g.taps());
  csr_count_kernel<<<hgrid, 256, 0, st>>>(t, cnt);
  csr_pad_kernel<<<cdiv(nkeys / (g.taps() + 1), 256), 256, 0, st>>>(cnt, nkeys / (g.taps() + 1), g.taps());
  csr_block_sums_kernel<<<P.scan_blocks, 256, 0, st>>>(cnt, bsum, nkeys);
  csr_scan_top_kernel<<<1, 1024, 0, st>>>(bsum, P.scan_blocks);
  csr_scan_final_kernel<<<P.scan_blocks, 256, 0, st>>>(cnt, bsum, start, nkeys);
  csr_fill_kernel<<<hgrid, 256, 0, st>>>(t, cnt, start, ent);
  csr_sort_kernel<<<cdiv(nkeys, 256), 256, 0, st>>>(start, ent, nkeys, g.taps());
  // start[nkeys] = total entries, a multiple of LIST_ALIGN; the pack covers the worst case (unused blocks are zeros)
  const int nblocks = (int)(P.ent_cap / LIST_ALIGN);
  csr_pack_kernel<<<cdiv(nblocks, 256), 256, 0, st>>>(ent, blocks, nblocks);
  SDB_LAUNCHED(8);
  SDB_CHECK_CUDA(cudaGetLastError());
  for (int i = 0; i < n; ++i) {
    const long long kb = P.key_base[P.group_of[i]];
    pb[i].start = start + kb; pb[i].ent = blocks;
  }
  return SDB_OK;
}

// ---- grad_input of all problems that want it: one gather launch ---------------------------------------------------
int tc_dx_multi(const TcProblem* pb, int n, const Geo& g, int io_dtype, int accumulate, cudaStream_t st) {
  const int NCH = nch_of(g), nch = nch_chunks(g);
  DxParams p{};
  p.g = g; p.accumulate = accumulate;
  int m = 0, total = 0;
  for (int i = 0; i < n; ++i) {
    if (!pb[i].gx) continue;
    const long long pin = (long long)pb[i].d.N * pb[i].d.H * pb[i].d.W;
    if (pin == 0) continue;
    DxProb& q = p.pr[m];
    q.dcol = pb[i].dcol; q.start = pb[i].start; q.ent = (const uint32_t*)pb[i].ent;
    q.out = pb[i].gx; q.d = pb[i].d;
    p.map.start[m] = total;
    total += cdiv(pin, TILE_M) * nch;
    ++m;
  }
  p.map.n = m; p.map.start[m] = total;
  if (total == 0) return SDB_OK;
  const bool obf = io_dtype == SDB_BF16;
  const size_t smem = (size_t)NCH * TILE_M * (obf ? 2 : 4);
  ProfScope prof(3, st);   // slot 3 = grad_input (slender_b200.h)
  static const int dx_threads = getenv("SDB_DX_THREADS") ? atoi(getenv("SDB_DX_THREADS")) : 512;
#define SDB_DX_LAUNCH(NCH_, BF_)                                                                  \
  {                                                                                               \
    if (dx_threads == 256) {                                                                      \
      SDB_ENSURE_SMEM((dcn_dx_gather_kernel<NCH_, BF_, 256>), smem);                              \
      dcn_dx_gather_kernel<NCH_, BF_, 256><<<total, 256, smem, st>>>(p);                          \
    } else {                                                                                      \
      SDB_ENSURE_SMEM((dcn_dx_gather_kernel<NCH_, BF_, 512>), smem);                              \
      dcn_dx_gather_kernel<NCH_, BF_, 512><<<total, 512, smem, st>>>(p);                          \
    }                                                                                             \
  }
  if (NCH == 128) { if (obf) SDB_DX_LAUNCH(128, true) else SDB_DX_LAUNCH(128, false) }
  else            { if (obf) SDB_DX_LAUNCH(64, true) else SDB_DX_LAUNCH(64, false) }
#undef SDB_DX_LAUNCH
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // namespace sdb
