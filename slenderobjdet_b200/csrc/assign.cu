// assign.cu -- fused pairwise IoU + label assignment (Matcher / TopKMatcher), bit-exact ints.
//
// Replaces ~20 eager kernels and the [M,X,2] / [M,X] temporaries of the reference
// (d2/structures/boxes.py:316-348, d2/modeling/matcher.py:61-126,
// sd/modeling/matchers/topk_matcher.py:38-86) with two launches:
//   1. per-anchor kernel: IoU against all GT (GT boxes staged in shared memory), running
//      max / argmax (first max = lowest GT index, as torch.max on ties), threshold -> label;
//   2. per-GT kernel: k rounds of a block-wide arg-max over the anchors (value desc, index asc)
//      for TopKMatcher, or max + equality sweep for Matcher's low-quality matches.
// The IoU arithmetic uses the *_rn intrinsics in the reference's operation order so every IoU is
// bit-identical to the float32 CPU result; HBM-bound by design (no tensor cores).
#include <float.h>

#include "common.cuh"

namespace sdb {
namespace {

__device__ __forceinline__ float box_area(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// boxes.py:333-347, one rounding per operation, no fma contraction
__device__ __forceinline__ float iou_exact(const float4 a, float area_a, const float4 b, float area_b) {
  float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  w = fmaxf(w, 0.f);
  h = fmaxf(h, 0.f);
  const float inter = __fmul_rn(w, h);
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) : 0.f;
}

struct Thresh {
  float t[8];
  int8_t l[9];
  int n;
};

__device__ __forceinline__ int8_t label_of(const Thresh& th, float v) {
  // labels[i] for v in [t[i-1], t[i]) with t[-1] = -inf, t[n] = +inf (matcher.py:96-98)
  int8_t lab = 1;
  float lo = -INFINITY;
  for (int i = 0; i <= th.n; ++i) {
    const float hi = (i < th.n) ? th.t[i] : INFINITY;
    if (v >= lo && v < hi) lab = th.l[i];
    lo = hi;
  }
  return lab;
}

// ---- kernel 1: one thread per anchor ------------------------------------------------------------
template <bool FROM_BOXES>
__global__ void __launch_bounds__(256) per_anchor_kernel(const float4* __restrict__ gt,
                                                         const float4* __restrict__ anchors,
                                                         const float* __restrict__ q, int M, int X,
                                                         Thresh th, int64_t* __restrict__ matches,
                                                         int8_t* __restrict__ labels,
                                                         float* __restrict__ iou_out) {
  extern __shared__ float4 s_gt[];  // [M] boxes then [M] areas (as float)
  float* s_area = reinterpret_cast<float*>(s_gt + M);
  if (FROM_BOXES) {
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      const float4 b = gt[m];
      s_gt[m] = b;
      s_area[m] = box_area(b);
    }
    __syncthreads();
  }
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= X) return;
  float4 a = make_float4(0, 0, 0, 0);
  float area_a = 0.f;
  if (FROM_BOXES) {
    a = anchors[x];
    area_a = box_area(a);
  }
  float best = -1.f;
  int arg = 0;
  for (int m = 0; m < M; ++m) {
    float v;
    if (FROM_BOXES) {
      v = iou_exact(s_gt[m], s_area[m], a, area_a);
      if (iou_out) iou_out[(size_t)m * X + x] = v;
    } else {
      v = q[(size_t)m * X + x];
    }
    if (v > best) {
      best = v;
      arg = m;
    }
  }
  matches[x] = arg;
  labels[x] = label_of(th, best);
}

// ---- kernel 2: one CTA per GT -------------------------------------------------------------------
struct Key {
  float v;
  int i;
};
// "a ranks before b": larger value first, then smaller index
__device__ __forceinline__ bool before(const Key a, const Key b) {
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}

__device__ __forceinline__ Key block_best(Key k, Key* s_red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    Key o;
    o.v = __shfl_xor_sync(0xffffffffu, k.v, d);
    o.i = __shfl_xor_sync(0xffffffffu, k.i, d);
    if (before(o, k)) k = o;
  }
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // s_red reuse across rounds
  if ((threadIdx.x & 31) == 0) s_red[warp] = k;
  __syncthreads();
  Key r = s_red[0];
  for (int w = 1; w < nw; ++w)
    if (before(s_red[w], r)) r = s_red[w];
  return r;
}

// TOPK: the per-thread candidate lists need 64+ registers -> 512-thread blocks; the low-quality-match scan runs 1024
template <bool FROM_BOXES, bool TOPK>
__global__ void __launch_bounds__(TOPK ? 512 : 1024) per_gt_kernel(const float4* __restrict__ gt,
                                                      const float4* __restrict__ anchors,
                                                      const float* __restrict__ q, int M, int X,
                                                      int topk, int8_t* __restrict__ labels) {
  __shared__ Key s_red[32];
  const int m = blockIdx.x;
  float4 g = make_float4(0, 0, 0, 0);
  float area_g = 0.f;
  if (FROM_BOXES) {
    g = gt[m];
    area_g = box_area(g);
  }
  auto value = [&](int x) -> float {
    if (FROM_BOXES) {
      const float4 a = anchors[x];
      return iou_exact(g, area_g, a, box_area(a));
    }
    return q[(size_t)m * X + x];
  };
  const Key none = {-INFINITY, 0x7fffffff};
  constexpr int KMAX = 16;
  if (TOPK && topk > 0 && topk <= KMAX) {
    // one pass over the anchors: every thread keeps its own KMAX best keys sorted in registers (bubble
    // insertion, static indexing); the block then pops the best list head `topk` times.  Same total order
    // as q.topk(k, dim=1) with the lowest-index tie rule, without re-evaluating the IoU k times.
    Key loc[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) loc[j] = none;
    for (int x0 = threadIdx.x; x0 < X; x0 += 4 * blockDim.x) {
      float vals[4];   // four independent loads / IoUs in flight before the (serial) insertions
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x = x0 + u * blockDim.x;
        vals[u] = x < X ? value(x) : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x = x0 + u * blockDim.x;
        Key cur = {vals[u], x};
        if (x < X && before(cur, loc[KMAX - 1])) {
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            if (before(cur, loc[j])) {
              const Key t = loc[j];
              loc[j] = cur;
              cur = t;
            }
          }
        }
      }
    }
    for (int r = 0; r < topk; ++r) {
      const Key best = block_best(loc[0], s_red);
      if (best.i >= X) break;  // fewer than k anchors (host rejects this case)
      if (loc[0].i == best.i) {   // the owner records the label and pops its head
        labels[best.i] = 1;
#pragma unroll
        for (int j = 0; j + 1 < KMAX; ++j) loc[j] = loc[j + 1];
        loc[KMAX - 1] = none;
      }
    }
  } else if (topk > 0) {
    // k rounds: best key that ranks strictly after the previously selected one
    Key prev = {INFINITY, -1};
    for (int r = 0; r < topk; ++r) {
      Key best = none;
      for (int x = threadIdx.x; x < X; x += blockDim.x) {
        const Key k = {value(x), x};
        if (before(prev, k) && before(k, best)) best = k;
      }
      best = block_best(best, s_red);
      if (best.i >= X) break;  // fewer than k anchors (host rejects this case)
      if (threadIdx.x == 0) labels[best.i] = 1;
      prev = best;
    }
  } else {
    // low-quality matches (matcher.py:105-126): all anchors equal to this GT's best quality
    Key best = none;
    for (int x = threadIdx.x; x < X; x += blockDim.x) {
      const Key k = {value(x), x};
      if (before(k, best)) best = k;
    }
    best = block_best(best, s_red);
    for (int x = threadIdx.x; x < X; x += blockDim.x)
      if (value(x) == best.v) labels[x] = 1;
  }
}

// ---- per-GT top-k split over TOPK_SPLITS blocks (needs the workspace) ---------------------------------
// One block per GT leaves 48 SMs idle and is instruction-bound (profiles/r1_secondary_kernels.md).  Part 1:
// block (m, s) scans its 1/TOPK_SPLITS of the anchors with the same per-thread sorted lists and writes its
// k best keys; part 2: one warp-sized block per GT merges TOPK_SPLITS x k candidates.  Same total order
// (value desc, index asc) at both levels, so the result is identical to the single-block scan.
constexpr int TOPK_SPLITS = 8, TOPK_KMAX = 16;

template <bool FROM_BOXES>
__global__ void __launch_bounds__(256) per_gt_topk_part_kernel(const float4* __restrict__ gt,
                                                               const float4* __restrict__ anchors,
                                                               const float* __restrict__ q, int M, int X, int topk,
                                                               Key* __restrict__ cand) {
  __shared__ Key s_red[32];
  const int m = blockIdx.x, sp = blockIdx.y;
  const int chunk = (X + TOPK_SPLITS - 1) / TOPK_SPLITS, x_lo = sp * chunk, x_hi = min(X, x_lo + chunk);
  float4 g = make_float4(0, 0, 0, 0);
  float area_g = 0.f;
  if (FROM_BOXES) {
    g = gt[m];
    area_g = box_area(g);
  }
  const Key none = {-INFINITY, 0x7fffffff};
  Key loc[TOPK_KMAX];
#pragma unroll
  for (int j = 0; j < TOPK_KMAX; ++j) loc[j] = none;
  for (int x = x_lo + threadIdx.x; x < x_hi; x += blockDim.x) {
    float v;
    if (FROM_BOXES) {
      const float4 a = anchors[x];
      v = iou_exact(g, area_g, a, box_area(a));
    } else {
      v = q[(size_t)m * X + x];
    }
    Key cur = {v, x};
    if (before(cur, loc[TOPK_KMAX - 1])) {
#pragma unroll
      for (int j = 0; j < TOPK_KMAX; ++j) {
        if (before(cur, loc[j])) {
          const Key t = loc[j];
          loc[j] = cur;
          cur = t;
        }
      }
    }
  }
  Key* out = cand + ((size_t)m * TOPK_SPLITS + sp) * TOPK_KMAX;
  for (int r = 0; r < topk; ++r) {
    const Key best = block_best(loc[0], s_red);
    if (threadIdx.x == 0) out[r] = best;
    if (best.i < X && loc[0].i == best.i) {
#pragma unroll
      for (int j = 0; j + 1 < TOPK_KMAX; ++j) loc[j] = loc[j + 1];
      loc[TOPK_KMAX - 1] = none;
    }
  }
}

__global__ void __launch_bounds__(128) per_gt_topk_merge_kernel(const Key* __restrict__ cand, int X, int topk,
                                                                int8_t* __restrict__ labels) {
  __shared__ Key s_red[32];
  const int m = blockIdx.x, t = threadIdx.x;
  const Key none = {-INFINITY, 0x7fffffff};
  Key mine = none;   // thread t holds candidate (split t / KMAX, rank t % KMAX)
  if ((t % TOPK_KMAX) < topk) mine = cand[(size_t)m * TOPK_SPLITS * TOPK_KMAX + t];
  for (int r = 0; r < topk; ++r) {
    const Key best = block_best(mine, s_red);
    if (best.i >= X) break;
    if (mine.i == best.i) {
      labels[best.i] = 1;
      mine = none;
    }
  }
}

__global__ void fill_default_kernel(int X, int8_t lab0, int64_t* matches, int8_t* labels) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < X) {
    matches[x] = 0;
    labels[x] = lab0;
  }
}

__global__ void __launch_bounds__(256) pairwise_iou_kernel(const float4* __restrict__ b1,
                                                           const float4* __restrict__ b2, int N1,
                                                           int N2, float* __restrict__ iou) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N2) return;
  const float4 b = b2[j];
  const float ab = box_area(b);
  for (int i = blockIdx.y; i < N1; i += gridDim.y) {
    const float4 a = b1[i];
    iou[(size_t)i * N2 + j] = iou_exact(a, box_area(a), b, ab);
  }
}

int assign_impl(const float* gt, const float* anchors, const float* q, int M, int X,
                const float* thresholds, const int8_t* labels, int nth, int topk, int alq,
                int64_t* matches, int8_t* match_labels, float* iou_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  SDB_REQUIRE(M >= 0 && X >= 0, SDB_ERR_INVALID, "negative sizes M=%d X=%d", M, X);
  SDB_REQUIRE(nth >= 1 && nth <= 8, SDB_ERR_INVALID, "n_thresholds must be in [1,8], got %d", nth);
  SDB_REQUIRE(thresholds && labels && matches && match_labels, SDB_ERR_INVALID, "NULL argument");
  Thresh th;
  th.n = nth;
  for (int i = 0; i < nth; ++i) {
    th.t[i] = thresholds[i];
    SDB_REQUIRE(i == 0 ? thresholds[0] > 0 : thresholds[i] >= thresholds[i - 1], SDB_ERR_INVALID,
                "thresholds must be positive and ascending");
  }
  for (int i = 0; i <= nth; ++i) {
    SDB_REQUIRE(labels[i] >= -1 && labels[i] <= 1, SDB_ERR_INVALID, "labels must be in {-1,0,1}");
    th.l[i] = labels[i];
  }
  if (X == 0) return SDB_OK;
  if (M == 0) {  // topk_matcher.py:53-63
    fill_default_kernel<<<cdiv(X, 256), 256, 0, st>>>(X, labels[0], matches, match_labels); SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
    return SDB_OK;
  }
  SDB_REQUIRE(topk <= X, SDB_ERR_INVALID, "selected index k out of range (topk=%d > %d anchors)", topk, X);
  const bool from_boxes = q == nullptr;
  const size_t smem = from_boxes ? (size_t)M * (sizeof(float4) + sizeof(float)) : 0;
  SDB_REQUIRE(smem <= 200 * 1024, SDB_ERR_UNSUPPORTED, "too many GT boxes (%d) for one shared-memory stage", M);
  if (from_boxes) {
    if (smem > 48 * 1024)
      SDB_CHECK_CUDA(cudaFuncSetAttribute(per_anchor_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    per_anchor_kernel<true><<<cdiv(X, 256), 256, smem, st>>>(
        (const float4*)gt, (const float4*)anchors, nullptr, M, X, th, matches, match_labels, iou_out); SDB_LAUNCHED(1);
  } else {
    per_anchor_kernel<false><<<cdiv(X, 256), 256, 0, st>>>(nullptr, nullptr, q, M, X, th, matches,
                                                           match_labels, nullptr); SDB_LAUNCHED(1);
  }
  SDB_CHECK_CUDA(cudaGetLastError());
  if (topk > 0 && topk <= TOPK_KMAX && ws && ws_bytes >= sdb_assign_workspace_bytes(M, X, topk) &&
      sdb_assign_workspace_bytes(M, X, topk) > 0) {
    Key* cand = (Key*)ws;
    dim3 grid(M, TOPK_SPLITS);
    if (from_boxes) per_gt_topk_part_kernel<true><<<grid, 256, 0, st>>>((const float4*)gt, (const float4*)anchors, nullptr, M, X, topk, cand);
    else            per_gt_topk_part_kernel<false><<<grid, 256, 0, st>>>(nullptr, nullptr, q, M, X, topk, cand);
    per_gt_topk_merge_kernel<<<M, TOPK_SPLITS * TOPK_KMAX, 0, st>>>(cand, X, topk, match_labels);
    SDB_LAUNCHED(2);
    SDB_CHECK_CUDA(cudaGetLastError());
  } else if (topk > 0 || alq) {
    const int threads = X >= 8192 ? (topk > 0 ? 512 : 1024) : 256;
    if (from_boxes) {
      if (topk > 0) per_gt_kernel<true, true><<<M, threads, 0, st>>>((const float4*)gt, (const float4*)anchors, nullptr, M, X, topk, match_labels);
      else          per_gt_kernel<true, false><<<M, threads, 0, st>>>((const float4*)gt, (const float4*)anchors, nullptr, M, X, topk, match_labels);
    } else {
      if (topk > 0) per_gt_kernel<false, true><<<M, threads, 0, st>>>(nullptr, nullptr, q, M, X, topk, match_labels);
      else          per_gt_kernel<false, false><<<M, threads, 0, st>>>(nullptr, nullptr, q, M, X, topk, match_labels);
    }
    SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
  }
  return SDB_OK;
}

}  // namespace
}  // namespace sdb

using namespace sdb;

// ---- RepPoints point_targets (reppointsv2.py:370-428) -------------------------------------------------
namespace sdb {
namespace {
struct PtWs {
  int lvl_min, lvl_max;   // memset to 0x7f7f7f7f / 0x80808080 before the min/max pass
};
__device__ __forceinline__ int point_level(float stride) { return (int)log2f(stride); }   // torch.log2(s).int()

__global__ void __launch_bounds__(256) pt_level_range_kernel(const float* __restrict__ strides, int X, PtWs* ws) {
  int lo = 0x7fffffff, hi = (int)0x80000000;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < X; i += gridDim.x * blockDim.x) {
    const int l = point_level(strides[i]);
    lo = min(lo, l);
    hi = max(hi, l);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&ws->lvl_min, lo);
    atomicMax(&ws->lvl_max, hi);
  }
}

// one block per GT: nearest point of the GT's level, then claim it with atomicMin on (distance bits, GT index)
__global__ void __launch_bounds__(512) pt_claim_kernel(const float2* __restrict__ points, const float* __restrict__ strides,
                                                       const float4* __restrict__ gt, int X, float scale,
                                                       const PtWs* __restrict__ ws, unsigned long long* __restrict__ keys) {
  __shared__ Key s_red[32];
  const int m = blockIdx.x;
  const float4 b = gt[m];
  const float cx = __fdiv_rn(__fadd_rn(b.x, b.z), 2.f), cy = __fdiv_rn(__fadd_rn(b.y, b.w), 2.f);
  const float w = fmaxf(__fsub_rn(b.z, b.x), 1e-6f), h = fmaxf(__fsub_rn(b.w, b.y), 1e-6f);
  int lvl = (int)__fdiv_rn(__fadd_rn(log2f(__fdiv_rn(w, scale)), log2f(__fdiv_rn(h, scale))), 2.f);
  lvl = max(ws->lvl_min, min(ws->lvl_max, lvl));
  // "before" ranks larger values first: search the maximum of -distance, lowest point index on ties
  Key best = {-INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < X; i += blockDim.x) {
    if (point_level(strides[i]) != lvl) continue;
    const float2 p = points[i];
    const float dx = __fdiv_rn(__fsub_rn(p.x, cx), w), dy = __fdiv_rn(__fsub_rn(p.y, cy), h);
    const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const Key k = {-dist, i};
    if (before(k, best)) best = k;
  }
  best = block_best(best, s_red);
  if (threadIdx.x == 0 && best.i < X) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(-best.v) << 32) | (unsigned)m;
    atomicMin(keys + best.i, key);   // distances are >= 0: their bit patterns order like the values
  }
}

__global__ void __launch_bounds__(256) pt_write_kernel(const unsigned long long* __restrict__ keys, const float4* __restrict__ gt,
                                                       const int64_t* __restrict__ gt_labels, int X, int64_t num_classes,
                                                       float4* __restrict__ boxes, int64_t* __restrict__ labels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= X) return;
  const unsigned long long k = keys[i];
  if (k == ~0ull) {
    boxes[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    labels[i] = num_classes;
  } else {
    const int m = (int)(k & 0xffffffffu);
    boxes[i] = gt[m];
    labels[i] = gt_labels[m];
  }
}
}  // namespace
}  // namespace sdb

// ---- FCOS compute_targets_for_locations (fcos/utils.py:108-212) ------------------------------------------
namespace sdb {
namespace {
struct FcosLevels {
  int end[8];        // cumulative location count per level
  float radius[8];   // stride * center_sampling_radius
  int n;
};

// centerness of an ltrb target: kind 0 = sqrt(c) (fcos/utils.py:295-300), kind 1 = pow(c, min(w/h, h/w)), the
// FCOSRepPoints module's own definition (fcos_rpd_s1_topk.py:25-55); c = min(l,r)/max(l,r) * min(t,b)/max(t,b)
__device__ __forceinline__ float centerness_of(const float4 rg, int kind) {
  const float c = __fmul_rn(__fdiv_rn(fminf(rg.x, rg.z), fmaxf(rg.x, rg.z)), __fdiv_rn(fminf(rg.y, rg.w), fmaxf(rg.y, rg.w)));
  if (kind == 0) return __fsqrt_rn(c);
  const float r1 = __fdiv_rn(__fadd_rn(rg.x, rg.z), __fadd_rn(rg.y, rg.w));
  return powf(c, fminf(r1, __fdiv_rn(1.f, r1)));
}

// grid (ceil(X / 256), images): image n reads GT rows [n*Mpad, n*Mpad + count[n]) (count == nullptr: Mpad rows)
__global__ void __launch_bounds__(256) fcos_targets_kernel(const float2* __restrict__ loc, const float2* __restrict__ soi,
                                                           const float4* __restrict__ gt_all, const int64_t* __restrict__ gt_cls_all,
                                                           const int32_t* __restrict__ gt_count, int X, int Mpad,
                                                           const FcosLevels lv, int center_sampling, int ctr_kind,
                                                           int64_t num_classes, int64_t* __restrict__ out_cls_all,
                                                           float4* __restrict__ out_reg_all, int* __restrict__ out_gt_all,
                                                           float* __restrict__ out_ctr_all, uint8_t* __restrict__ out_topk_all) {
  extern __shared__ float4 s_gt[];   // M boxes, then M areas
  const int img = blockIdx.y;
  const int M = gt_count ? gt_count[img] : Mpad;
  const float4* gt = gt_all + (size_t)img * Mpad;
  const int64_t* gt_cls = gt_cls_all + (size_t)img * Mpad;
  int64_t* out_cls = out_cls_all + (size_t)img * X;
  float4* out_reg = out_reg_all + (size_t)img * X;
  int* out_gt = out_gt_all ? out_gt_all + (size_t)img * X : nullptr;
  float* out_ctr = out_ctr_all ? out_ctr_all + (size_t)img * X : nullptr;
  uint8_t* out_topk = out_topk_all ? out_topk_all + (size_t)img * X : nullptr;
  float* s_area = reinterpret_cast<float*>(s_gt + Mpad);
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const float4 b = gt[m];
    s_gt[m] = b;
    s_area[m] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));   // Boxes.area()
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= X) return;
  const float2 p = loc[i], so = soi[i];
  float rad = 0.f;
  if (center_sampling) {
    int l = 0;
    while (l + 1 < lv.n && i >= lv.end[l]) ++l;
    rad = lv.radius[l];
  }
  // get_sample_region's shortcut: sum over locations of GT 0's centre x == 0 -> "no gt", nothing is inside
  const bool no_gt = center_sampling && M > 0 && __fdiv_rn(__fadd_rn(s_gt[0].x, s_gt[0].z), 2.f) == 0.f;
  const float INF = 100000000.f;
  float best = __int_as_float(0x7f800000);
  int best_m = 0;
  for (int m = 0; m < M; ++m) {
    const float4 b = s_gt[m];
    const float l = __fsub_rn(p.x, b.x), t = __fsub_rn(p.y, b.y), r = __fsub_rn(b.z, p.x), bt = __fsub_rn(b.w, p.y);
    bool inside;
    if (center_sampling) {
      const float cx = __fdiv_rn(__fadd_rn(b.x, b.z), 2.f), cy = __fdiv_rn(__fadd_rn(b.y, b.w), 2.f);
      const float xmin = __fsub_rn(cx, rad), ymin = __fsub_rn(cy, rad), xmax = __fadd_rn(cx, rad), ymax = __fadd_rn(cy, rad);
      const float x0 = xmin > b.x ? xmin : b.x, y0 = ymin > b.y ? ymin : b.y;
      const float x1 = xmax > b.z ? b.z : xmax, y1 = ymax > b.w ? b.w : ymax;
      const float cl = __fsub_rn(p.x, x0), cr = __fsub_rn(x1, p.x), ct = __fsub_rn(p.y, y0), cb = __fsub_rn(y1, p.y);
      inside = !no_gt && fminf(fminf(cl, ct), fminf(cr, cb)) > 0.f;
    } else {
      inside = fminf(fminf(l, t), fminf(r, bt)) > 0.f;
    }
    const float mx = fmaxf(fmaxf(l, t), fmaxf(r, bt));
    const bool cared = mx >= so.x && mx <= so.y;
    const float v = (inside && cared) ? s_area[m] : INF;
    if (v < best) {
      best = v;
      best_m = m;
    }
  }
  const float4 b = M > 0 ? s_gt[best_m] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 rg = make_float4(__fsub_rn(p.x, b.x), __fsub_rn(p.y, b.y), __fsub_rn(b.z, p.x), __fsub_rn(b.w, p.y));
  out_reg[i] = rg;
  const int64_t cls = (best == INF || M == 0) ? num_classes : gt_cls[best_m];
  out_cls[i] = cls;
  if (out_gt) {   // top-k variant: the GT this location belongs to (foreground only) and its centerness (:270-273)
    const bool fg = cls >= 0 && cls != num_classes;
    out_gt[i] = fg ? best_m : -1;
    out_ctr[i] = centerness_of(rg, ctr_kind);
    out_topk[i] = 0;
  }
}

// one block per GT: the `topk` locations assigned to it with the highest centerness (all when it has fewer)
// grid (Mpad, images)
__global__ void __launch_bounds__(512) fcos_topk_kernel(const int* __restrict__ loc_gt_all, const float* __restrict__ ctr_all, int X,
                                                        int topk, uint8_t* __restrict__ out_topk_all) {
  __shared__ Key s_red[32];
  constexpr int KMAX = 16;
  const int m = blockIdx.x;
  const int* loc_gt = loc_gt_all + (size_t)blockIdx.y * X;
  const float* ctr = ctr_all + (size_t)blockIdx.y * X;
  uint8_t* out_topk = out_topk_all + (size_t)blockIdx.y * X;
  const Key none = {-INFINITY, 0x7fffffff};
  Key loc[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) loc[j] = none;
  for (int x = threadIdx.x; x < X; x += blockDim.x) {
    if (loc_gt[x] != m) continue;
    Key cur = {ctr[x], x};
    if (before(cur, loc[KMAX - 1])) {
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (before(cur, loc[j])) {
          const Key t = loc[j];
          loc[j] = cur;
          cur = t;
        }
      }
    }
  }
  for (int r = 0; r < topk; ++r) {
    const Key best = block_best(loc[0], s_red);
    if (best.i >= X) break;   // fewer than topk locations belong to this GT: all of them are marked already
    if (loc[0].i == best.i) {
      out_topk[best.i] = 1;
#pragma unroll
      for (int j = 0; j + 1 < KMAX; ++j) loc[j] = loc[j + 1];
      loc[KMAX - 1] = none;
    }
  }
}
}  // namespace
}  // namespace sdb

extern "C" {

static int fcos_targets_impl(const float* locations, const float* sizes_of_interest, const float* gt,
                             const int64_t* gt_classes, const int32_t* gt_counts, int32_t n_images, int32_t X, int32_t M,
                             const int32_t* num_points_per_level, const float* level_strides, int32_t n_levels,
                             float center_sampling_radius, int64_t num_classes, int32_t topk, int32_t centerness_kind,
                             int64_t* out_classes, float* out_reg, uint8_t* out_topk, void* workspace,
                             size_t workspace_bytes, void* stream) {
  using namespace sdb;
  SDB_REQUIRE(X >= 0 && n_images >= 0, SDB_ERR_INVALID, "negative sizes");
  SDB_REQUIRE(M > 0 || gt_counts != nullptr, SDB_ERR_INVALID, "fcos targets need at least one GT box (M=%d)", M);
  if (X == 0 || n_images == 0) return SDB_OK;
  SDB_REQUIRE(locations && sizes_of_interest && out_classes && out_reg && (M == 0 || (gt && gt_classes)), SDB_ERR_INVALID,
              "NULL argument");
  SDB_REQUIRE(centerness_kind == 0 || centerness_kind == 1, SDB_ERR_INVALID, "unknown centerness kind %d", centerness_kind);
  int* loc_gt = nullptr;
  float* ctr = nullptr;
  if (out_topk) {
    SDB_REQUIRE(topk > 0 && topk <= 16, SDB_ERR_UNSUPPORTED, "topk must be in [1,16], got %d", topk);
    SDB_REQUIRE(workspace && workspace_bytes >= (size_t)n_images * sdb_fcos_topk_workspace_bytes(X), SDB_ERR_WORKSPACE,
                "fcos top-k workspace too small");
    loc_gt = (int*)workspace;
    ctr = (float*)workspace + (size_t)n_images * X;
  }
  FcosLevels lv{};
  const bool cs = center_sampling_radius > 0.f;
  if (cs) {
    SDB_REQUIRE(num_points_per_level && level_strides && n_levels >= 1 && n_levels <= 8, SDB_ERR_INVALID,
                "center sampling needs 1..8 levels");
    int acc = 0;
    for (int l = 0; l < n_levels; ++l) {
      acc += num_points_per_level[l];
      lv.end[l] = acc;
      lv.radius[l] = (float)((double)level_strides[l] * (double)center_sampling_radius);
    }
    SDB_REQUIRE(acc == X, SDB_ERR_INVALID, "num_points_per_level sums to %d, expected %d", acc, X);
    lv.n = n_levels;
  }
  const size_t smem = (size_t)(M > 0 ? M : 1) * (sizeof(float4) + sizeof(float));
  SDB_REQUIRE(smem <= 48 * 1024, SDB_ERR_UNSUPPORTED, "too many GT boxes (%d) for one shared-memory stage", M);
  cudaStream_t st = (cudaStream_t)stream;
  fcos_targets_kernel<<<dim3(cdiv(X, 256), n_images), 256, smem, st>>>(
      (const float2*)locations, (const float2*)sizes_of_interest, (const float4*)gt, gt_classes, gt_counts, X, M, lv,
      cs ? 1 : 0, centerness_kind, num_classes, out_classes, (float4*)out_reg, loc_gt, ctr, out_topk);
  SDB_LAUNCHED(1);
  if (out_topk && M > 0) {
    fcos_topk_kernel<<<dim3(M, n_images), X >= 8192 ? 512 : 256, 0, st>>>(loc_gt, ctr, X, topk, out_topk);
    SDB_LAUNCHED(1);
  }
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

int sdb_fcos_location_targets(const float* locations, const float* sizes_of_interest, const float* gt,
                              const int64_t* gt_classes, int32_t X, int32_t M, const int32_t* num_points_per_level,
                              const float* level_strides, int32_t n_levels, float center_sampling_radius,
                              int64_t num_classes, int64_t* out_classes, float* out_reg, void* stream) {
  SDB_REQUIRE(M > 0, SDB_ERR_INVALID, "fcos targets need at least one GT box (M=%d)", M);
  return fcos_targets_impl(locations, sizes_of_interest, gt, gt_classes, nullptr, 1, X, M, num_points_per_level,
                           level_strides, n_levels, center_sampling_radius, num_classes, 0, 0, out_classes, out_reg, nullptr,
                           nullptr, 0, stream);
}

size_t sdb_fcos_topk_workspace_bytes(int32_t X) { return X > 0 ? (size_t)X * 8 : 0; }

int sdb_fcos_topk_location_targets(const float* locations, const float* sizes_of_interest, const float* gt,
                                   const int64_t* gt_classes, int32_t X, int32_t M,
                                   const int32_t* num_points_per_level, const float* level_strides, int32_t n_levels,
                                   float center_sampling_radius, int64_t num_classes, int32_t topk,
                                   int64_t* out_classes, float* out_reg, uint8_t* out_topk, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  SDB_REQUIRE(out_topk != nullptr, SDB_ERR_INVALID, "out_topk is NULL");
  SDB_REQUIRE(M > 0, SDB_ERR_INVALID, "fcos targets need at least one GT box (M=%d)", M);
  return fcos_targets_impl(locations, sizes_of_interest, gt, gt_classes, nullptr, 1, X, M, num_points_per_level,
                           level_strides, n_levels, center_sampling_radius, num_classes, topk, 0, out_classes, out_reg,
                           out_topk, workspace, workspace_bytes, stream);
}

int sdb_fcos_location_targets_batched(const float* locations, const float* sizes_of_interest, const float* gt,
                                      const int64_t* gt_classes, const int32_t* gt_counts, int32_t n_images, int32_t X,
                                      int32_t M_pad, const int32_t* num_points_per_level, const float* level_strides,
                                      int32_t n_levels, float center_sampling_radius, int64_t num_classes, int32_t topk,
                                      int32_t centerness_kind, int64_t* out_classes, float* out_reg, uint8_t* out_topk,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  SDB_REQUIRE(gt_counts != nullptr && M_pad >= 0, SDB_ERR_INVALID, "gt_counts is NULL or M_pad < 0");
  return fcos_targets_impl(locations, sizes_of_interest, gt, gt_classes, gt_counts, n_images, X, M_pad,
                           num_points_per_level, level_strides, n_levels, center_sampling_radius, num_classes, topk,
                           centerness_kind, out_classes, out_reg, topk > 0 ? out_topk : nullptr, workspace, workspace_bytes,
                           stream);
}

size_t sdb_point_targets_workspace_bytes(int32_t X) { return X > 0 ? 256 + (size_t)X * 8 : 0; }

int sdb_point_targets(const float* points, const float* strides, const float* gt, const int64_t* gt_labels,
                      int32_t X, int32_t M, float scale, int64_t num_classes, float* assigned_bboxes,
                      int64_t* assigned_labels, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace sdb;
  SDB_REQUIRE(X > 0 && M > 0, SDB_ERR_INVALID, "No gt or bboxes");   // reppointsv2.py:383-384
  SDB_REQUIRE(points && strides && gt && gt_labels && assigned_bboxes && assigned_labels, SDB_ERR_INVALID, "NULL argument");
  SDB_REQUIRE(workspace && workspace_bytes >= sdb_point_targets_workspace_bytes(X), SDB_ERR_WORKSPACE,
              "point_targets workspace too small");
  SDB_REQUIRE(scale > 0.f, SDB_ERR_INVALID, "point_base_scale must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  PtWs* ws = (PtWs*)workspace;
  unsigned long long* keys = (unsigned long long*)((uint8_t*)workspace + 256);
  SDB_CHECK_CUDA(cudaMemsetAsync(&ws->lvl_min, 0x7f, 4, st));
  SDB_CHECK_CUDA(cudaMemsetAsync(&ws->lvl_max, 0x80, 4, st));
  SDB_CHECK_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)X * 8, st));
  int blocks = cdiv(X, 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  pt_level_range_kernel<<<blocks, 256, 0, st>>>(strides, X, ws);
  pt_claim_kernel<<<M, 512, 0, st>>>((const float2*)points, strides, (const float4*)gt, X, scale, ws, keys);
  pt_write_kernel<<<cdiv(X, 256), 256, 0, st>>>(keys, (const float4*)gt, gt_labels, X, num_classes, (float4*)assigned_bboxes,
                                                assigned_labels);
  SDB_LAUNCHED(3);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

size_t sdb_assign_workspace_bytes(int32_t M, int32_t X, int32_t topk) {
  // candidate keys of the split per-GT top-k; small problems use the single-block scan and need none
  if (topk <= 0 || topk > 16 || M <= 0 || X < 8192) return 0;
  return (size_t)M * 8 * 16 * 8;   // M x TOPK_SPLITS x TOPK_KMAX x sizeof(Key)
}

int sdb_iou_assign(const float* gt, const float* anchors, int32_t M, int32_t X,
                   const float* thresholds, const int8_t* labels, int32_t n_thresholds, int32_t topk,
                   int32_t allow_low_quality, int64_t* matches, int8_t* match_labels, float* iou_out,
                   void* workspace, size_t workspace_bytes, void* stream) {
  SDB_REQUIRE((gt || M == 0) && (anchors || X == 0), SDB_ERR_INVALID, "NULL boxes");
  // thresholds / labels are HOST arrays (a handful of config scalars)
  return assign_impl(gt, anchors, nullptr, M, X, thresholds, labels, n_thresholds, topk,
                     allow_low_quality, matches, match_labels, iou_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int sdb_match_quality_assign(const float* q, int32_t M, int32_t X, const float* thresholds,
                             const int8_t* labels, int32_t n_thresholds, int32_t topk,
                             int32_t allow_low_quality, int64_t* matches, int8_t* match_labels,
                             void* workspace, size_t workspace_bytes, void* stream) {
  SDB_REQUIRE(q || M == 0 || X == 0, SDB_ERR_INVALID, "NULL quality matrix");
  static const float dummy = 0.f;
  return assign_impl(nullptr, nullptr, q ? q : &dummy, M, X, thresholds, labels, n_thresholds, topk,
                     allow_low_quality, matches, match_labels, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

int sdb_pairwise_iou(const float* boxes1, const float* boxes2, int32_t N1, int32_t N2, float* iou,
                     void* stream) {
  SDB_REQUIRE(N1 >= 0 && N2 >= 0, SDB_ERR_INVALID, "negative sizes");
  if (N1 == 0 || N2 == 0) return SDB_OK;
  SDB_REQUIRE(boxes1 && boxes2 && iou, SDB_ERR_INVALID, "NULL argument");
  dim3 grid(cdiv(N2, 256), N1 < 64 ? N1 : 64);
  pairwise_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes1,
                                                              (const float4*)boxes2, N1, N2, iou); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // extern "C"
