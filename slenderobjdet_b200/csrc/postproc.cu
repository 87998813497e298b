// postproc.cu -- inference post-processing of a dense point head, sm_100a.
//
// Replaces RepPointsV2.inference / inference_single_image
// (slender_det/modeling/meta_arch/reppoints/reppointsv2.py:486-603) for a whole batch without host round trips:
//   per (image, level): score = sigmoid(logit) over H*W*K cells; the `topk` best cells whose score exceeds the threshold
//                       (the reference sorts ALL cells, :572; here a histogram of the score bits finds the cut, only the
//                       cells at or above it are compacted and sorted); their boxes are decoded from the refined
//                       point sets (pts_to_bbox :328-366, x stride, + centre, clamped to the image :558-564);
//   per image        : candidates of all levels sorted by score, class-aware greedy NMS (IoU > threshold suppresses,
//                       detectron2/layers/nms.py:10-29 -> torchvision batched_nms) via a suppression bit matrix and
//                       a serial scan on the device, first `max_det` survivors.
// Bandwidth / latency-bound integer and float work: no tensor cores.  Reads the head's NCHW outputs in place (the
// reference permutes them to [HW, K] first); flat cell index = point * K + class as in the reference.
#include "common.cuh"

namespace sdb {
namespace {

constexpr int PP_MAX_LEVELS = 8;
constexpr int PP_MAX_IMAGES = 64;
constexpr int PP_CAP = 4096;          // candidates sorted per (image, level)
constexpr int PP_MAX_BINS = 16384;
constexpr int PP_ELEMS = 8;           // cells per thread in the scan kernels

struct PPLevel {
  const float* cls;    // [N, K, H, W]
  const float* pts;    // [N, 2*num_points, H, W]
  const float* ctr;    // [HW, 2] (x, y)
  int HW;
  float stride;
};
struct PPParams {
  PPLevel lv[PP_MAX_LEVELS];
  int blk_start[PP_MAX_LEVELS + 1];   // scan blocks per level
  int nlv, K, P2, transform;          // transform: 0 minmax, 1 partial_minmax, 2 moment
  float mt_w, mt_h;                   // exp(moment_transfer)
  float thr;
  uint32_t thr_bits;
  int shift, nbins, topk, max_det;
  float nms_thr;
  int img_h[PP_MAX_IMAGES], img_w[PP_MAX_IMAGES];
  // workspace
  int* hist;        // [N][nlv][nbins]
  int* cut;         // [N][nlv]
  int* cnt;         // [N][nlv]
  unsigned long long* keys;   // [N][nlv][PP_CAP]
  float4* lbox;     // [N][nlv][topk]
  float* lscore;    // [N][nlv][topk]
  int* lcls;        // [N][nlv][topk]
  int* lcount;      // [N][nlv]
  float4* sbox;     // [N][M]   M = nlv * topk, sorted by score
  float* sscore;    // [N][M]
  int* scls;        // [N][M]
  int* total;       // [N]
  unsigned long long* mask;   // [N][M][MW]
  int M, MW;
  int* overflow;    // [1]
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

__device__ __forceinline__ int find_level(const PPParams& p, int blk) {
  int l = 0;
  while (l + 1 < p.nlv && blk >= p.blk_start[l + 1]) ++l;
  return l;
}

// pass 1 (COMPACT = false): histogram of the score bits of the cells above the threshold;
// pass 2 (COMPACT = true) : cells at or above the cut bin -> (score bits, ~flat index) keys
template <bool COMPACT>
__global__ void __launch_bounds__(256) pp_scan_kernel(const __grid_constant__ PPParams p) {
  const int img = blockIdx.y, l = find_level(p, blockIdx.x);
  const PPLevel& L = p.lv[l];
  const long long cells = (long long)L.HW * p.K;
  const float* cls = L.cls + (size_t)img * cells;
  const int slot_il = img * p.nlv + l;
  const int cut = COMPACT ? p.cut[slot_il] : 0;
  const long long base = (long long)(blockIdx.x - p.blk_start[l]) * (256 * PP_ELEMS);
#pragma unroll
  for (int u = 0; u < PP_ELEMS; ++u) {
    const long long e = base + u * 256 + threadIdx.x;   // NCHW order: class * HW + point
    if (e >= cells) break;
    const float s = sigmoidf_ref(__ldg(cls + e));
    if (!(s > p.thr)) continue;
    int bin = (int)((__float_as_uint(s) - p.thr_bits) >> p.shift);
    bin = bin < p.nbins - 1 ? bin : p.nbins - 1;
    if (!COMPACT) {
      atomicAdd(p.hist + (size_t)slot_il * p.nbins + bin, 1);
    } else if (bin >= cut) {
      const int k = (int)(e / L.HW), pt = (int)(e - (long long)k * L.HW);
      const unsigned flat = (unsigned)pt * (unsigned)p.K + (unsigned)k;     // the reference's flattened [HW, K] index
      const int slot = atomicAdd(p.cnt + slot_il, 1);
      if (slot < PP_CAP) p.keys[(size_t)slot_il * PP_CAP + slot] = ((unsigned long long)__float_as_uint(s) << 32) | (0xffffffffu - flat);
      else *p.overflow = 1;
    }
  }
}

// one warp per (level, image): highest bin b with  sum_{bins >= b} >= topk  (0 when fewer than topk cells pass)
__global__ void __launch_bounds__(32) pp_cut_kernel(const __grid_constant__ PPParams p) {
  const int slot_il = blockIdx.y * p.nlv + blockIdx.x, lane = threadIdx.x;
  const int* h = p.hist + (size_t)slot_il * p.nbins;
  int running = 0, cut = 0;
  for (int top = p.nbins; top > 0; top -= 32) {
    const int b = top - 1 - lane;                       // lane 0 = highest bin of the chunk
    const int v = b >= 0 ? h[b] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, running + incl >= p.topk);
    if (hit) {
      cut = top - 1 - (__ffs(hit) - 1);
      break;
    }
    running += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    p.cut[slot_il] = cut > 0 ? cut : 0;
    p.cnt[slot_il] = 0;
  }
}

__device__ __forceinline__ void bitonic_desc(unsigned long long* k, int n, int tid, int nthreads) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < n / 2; t += nthreads) {
        const int i = 2 * t - (t & (stride - 1));       // lower index of the pair
        const int j = i + stride;
        const bool desc = ((i & size) == 0);
        const unsigned long long a = k[i], b = k[j];
        if ((a < b) == desc) { k[i] = b; k[j] = a; }
      }
    }
  }
  __syncthreads();
}

// one block per (level, image): sort the compacted keys, keep the first min(topk, n), decode their boxes
__global__ void __launch_bounds__(1024) pp_level_kernel(const __grid_constant__ PPParams p) {
  __shared__ unsigned long long keys[PP_CAP];
  const int l = blockIdx.x, img = blockIdx.y, slot_il = img * p.nlv + l;
  const PPLevel& L = p.lv[l];
  int n = p.cnt[slot_il];
  n = n < PP_CAP ? n : PP_CAP;
  for (int i = threadIdx.x; i < PP_CAP; i += blockDim.x) keys[i] = i < n ? p.keys[(size_t)slot_il * PP_CAP + i] : 0ull;
  bitonic_desc(keys, PP_CAP, threadIdx.x, blockDim.x);
  const int m = n < p.topk ? n : p.topk;
  if (threadIdx.x == 0) p.lcount[slot_il] = m;
  const float* pts = L.pts + (size_t)img * p.P2 * L.HW;
  const float iw = (float)p.img_w[img], ih = (float)p.img_h[img];
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    const unsigned long long key = keys[j];
    const unsigned flat = 0xffffffffu - (unsigned)(key & 0xffffffffu);
    const int pt = (int)(flat / (unsigned)p.K), k = (int)(flat % (unsigned)p.K);
    // pts_to_bbox (:328-366): channels 0, 2, 4, ... are x, 1, 3, 5, ... are y
    const int np = p.transform == 1 ? 4 : p.P2 / 2;
    float x0, y0, x1, y1;
    if (p.transform < 2) {
      x0 = y0 = INFINITY; x1 = y1 = -INFINITY;
      for (int q = 0; q < np; ++q) {
        const float x = __ldg(pts + (size_t)(2 * q) * L.HW + pt), y = __ldg(pts + (size_t)(2 * q + 1) * L.HW + pt);
        x0 = fminf(x0, x); x1 = fmaxf(x1, x); y0 = fminf(y0, y); y1 = fmaxf(y1, y);
      }
    } else {   // moment: mean +- std * exp(moment_transfer), unbiased std as torch.std
      float sx = 0.f, sy = 0.f;
      for (int q = 0; q < np; ++q) { sx += __ldg(pts + (size_t)(2 * q) * L.HW + pt); sy += __ldg(pts + (size_t)(2 * q + 1) * L.HW + pt); }
      const float mx = sx / (float)np, my = sy / (float)np;
      float vx = 0.f, vy = 0.f;
      for (int q = 0; q < np; ++q) {
        const float dx = __ldg(pts + (size_t)(2 * q) * L.HW + pt) - mx, dy = __ldg(pts + (size_t)(2 * q + 1) * L.HW + pt) - my;
        vx += dx * dx; vy += dy * dy;
      }
      const float hw = sqrtf(vx / (float)(np - 1)) * p.mt_w, hh = sqrtf(vy / (float)(np - 1)) * p.mt_h;
      x0 = mx - hw; y0 = my - hh; x1 = mx + hw; y1 = my + hh;
    }
    const float cx = __ldg(L.ctr + 2 * pt), cy = __ldg(L.ctr + 2 * pt + 1);
    float4 b;
    b.x = fminf(fmaxf(__fadd_rn(__fmul_rn(x0, L.stride), cx), 0.f), iw);
    b.y = fminf(fmaxf(__fadd_rn(__fmul_rn(y0, L.stride), cy), 0.f), ih);
    b.z = fminf(fmaxf(__fadd_rn(__fmul_rn(x1, L.stride), cx), 0.f), iw);
    b.w = fminf(fmaxf(__fadd_rn(__fmul_rn(y1, L.stride), cy), 0.f), ih);
    const size_t o = (size_t)slot_il * p.topk + j;
    p.lbox[o] = b;
    p.lscore[o] = __uint_as_float((unsigned)(key >> 32));
    p.lcls[o] = k;
  }
}

// one block per image: merge the levels' candidates into one list sorted by score (ties: level order, then rank)
__global__ void __launch_bounds__(1024) pp_merge_kernel(const __grid_constant__ PPParams p, int sortn) {
  extern __shared__ unsigned long long mkeys[];
  const int img = blockIdx.x;
  int start[PP_MAX_LEVELS + 1];
  start[0] = 0;
  for (int l = 0; l < p.nlv; ++l) start[l + 1] = start[l] + p.lcount[img * p.nlv + l];
  const int T = start[p.nlv];
  for (int i = threadIdx.x; i < sortn; i += blockDim.x) mkeys[i] = 0ull;
  __syncthreads();
  for (int l = 0; l < p.nlv; ++l) {
    const int m = start[l + 1] - start[l];
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const unsigned pos = (unsigned)(l * p.topk + j);
      const float s = p.lscore[(size_t)(img * p.nlv + l) * p.topk + j];
      mkeys[start[l] + j] = ((unsigned long long)__float_as_uint(s) << 32) | (0xffffffffu - pos);
    }
  }
  bitonic_desc(mkeys, sortn, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const unsigned pos = 0xffffffffu - (unsigned)(mkeys[i] & 0xffffffffu);
    const size_t src = (size_t)img * p.nlv * p.topk + pos;
    p.sbox[(size_t)img * p.M + i] = p.lbox[src];
    p.sscore[(size_t)img * p.M + i] = p.lscore[src];
    p.scls[(size_t)img * p.M + i] = p.lcls[src];
  }
  if (threadIdx.x == 0) p.total[img] = T;
}

// suppression bit matrix: bit j of mask[i][j / 64] = candidate j (j > i, same class) overlaps i with IoU > threshold
__global__ void __launch_bounds__(64) pp_nms_mask_kernel(const __grid_constant__ PPParams p) {
  const int img = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
  const int T = p.total[img];
  if (cb < rb || rb * 64 >= T || cb * 64 >= T) return;
  __shared__ float4 cbox[64];
  __shared__ int ccls[64];
  const float4* box = p.sbox + (size_t)img * p.M;
  const int* cls = p.scls + (size_t)img * p.M;
  const int j0 = cb * 64, nc = min(64, T - j0);
  if ((int)threadIdx.x < nc) { cbox[threadIdx.x] = box[j0 + threadIdx.x]; ccls[threadIdx.x] = cls[j0 + threadIdx.x]; }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= T) return;
  const float4 a = box[i];
  const int ca = cls[i];
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  unsigned long long bits = 0ull;
  for (int t = (rb == cb ? threadIdx.x + 1 : 0); t < nc; ++t) {
    if (ccls[t] != ca) continue;
    const float4 b = cbox[t];
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    if (__fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > p.nms_thr) bits |= 1ull << t;
  }
  p.mask[((size_t)img * p.M + i) * p.MW + cb] = bits;
}

// one block per image: greedy scan over the sorted candidates; the first max_det survivors are the detections
__global__ void __launch_bounds__(128) pp_nms_scan_kernel(const __grid_constant__ PPParams p, float* __restrict__ out_boxes,
                                                          float* __restrict__ out_scores, long long* __restrict__ out_classes,
                                                          int* __restrict__ out_count) {
  __shared__ unsigned long long remv[128];
  __shared__ int kept_s;
  const int img = blockIdx.x, T = p.total[img];
  const int words = (T + 63) / 64;
  for (int w = threadIdx.x; w < 128; w += blockDim.x) remv[w] = 0ull;
  if (threadIdx.x == 0) kept_s = 0;
  __syncthreads();
  for (int i = 0; i < T; ++i) {
    const bool alive = !((remv[i >> 6] >> (i & 63)) & 1ull);
    const int kept = kept_s;
    __syncthreads();
    if (alive) {
      if (threadIdx.x == 0) {
        const size_t o = (size_t)img * p.max_det + kept;
        const float4 b = p.sbox[(size_t)img * p.M + i];
        out_boxes[4 * o] = b.x; out_boxes[4 * o + 1] = b.y; out_boxes[4 * o + 2] = b.z; out_boxes[4 * o + 3] = b.w;
        out_scores[o] = p.sscore[(size_t)img * p.M + i];
        out_classes[o] = p.scls[(size_t)img * p.M + i];
        kept_s = kept + 1;
      }
      const unsigned long long* row = p.mask + ((size_t)img * p.M + i) * p.MW;
      for (int w = (i >> 6) + threadIdx.x; w < words; w += blockDim.x) remv[w] |= row[w];   // words below i/64 are never read again
    }
    __syncthreads();
    if (kept_s >= p.max_det) break;
  }
  if (threadIdx.x == 0) out_count[img] = kept_s;
}

struct PPWs {
  size_t hist, cut, cnt, keys, lbox, lscore, lcls, lcount, sbox, sscore, scls, total, mask, overflow, bytes;
};
inline size_t up(size_t v) { return (v + 255) / 256 * 256; }
PPWs pp_ws(int nlv, int n, int topk, int nbins) {
  PPWs w{};
  const size_t M = (size_t)nlv * topk, MW = (M + 63) / 64;
  size_t o = 0;
  w.hist = o;    o = up(o + (size_t)n * nlv * nbins * 4);
  w.cut = o;     o = up(o + (size_t)n * nlv * 4);
  w.cnt = o;     o = up(o + (size_t)n * nlv * 4);
  w.overflow = o; o = up(o + 4);
  w.keys = o;    o = up(o + (size_t)n * nlv * PP_CAP * 8);
  w.lbox = o;    o = up(o + (size_t)n * M * 16);
  w.lscore = o;  o = up(o + (size_t)n * M * 4);
  w.lcls = o;    o = up(o + (size_t)n * M * 4);
  w.lcount = o;  o = up(o + (size_t)n * nlv * 4);
  w.sbox = o;    o = up(o + (size_t)n * M * 16);
  w.sscore = o;  o = up(o + (size_t)n * M * 4);
  w.scls = o;    o = up(o + (size_t)n * M * 4);
  w.total = o;   o = up(o + (size_t)n * 4);
  w.mask = o;    o = up(o + (size_t)n * M * MW * 8);
  w.bytes = o;
  return w;
}
void pp_bins(float thr, uint32_t& thr_bits, int& shift, int& nbins) {
  const float t = thr > 0.f ? thr : 0.f;
  memcpy(&thr_bits, &t, 4);
  const uint32_t one = 0x3f800000u;
  shift = 13;
  while (((one - thr_bits) >> shift) + 2 > (uint32_t)PP_MAX_BINS) ++shift;
  nbins = (int)((one - thr_bits) >> shift) + 2;
}
}  // namespace
}  // namespace sdb

using namespace sdb;

extern "C" {

size_t sdb_points_postprocess_workspace_bytes(int32_t n_levels, int32_t n_images, int32_t topk, float score_thresh) {
  if (n_levels < 1 || n_levels > PP_MAX_LEVELS || n_images < 1 || n_images > PP_MAX_IMAGES || topk < 1) return 0;
  uint32_t tb; int sh, nb;
  pp_bins(score_thresh, tb, sh, nb);
  return pp_ws(n_levels, n_images, topk, nb).bytes;
}

int sdb_points_postprocess(const sdb_pp_level* levels, int32_t n_levels, int32_t n_images, int32_t num_classes,
                           int32_t num_points, int32_t transform, const float* moment_transfer, const int32_t* image_sizes,
                           float score_thresh, int32_t topk, float nms_thresh, int32_t max_det, float* out_boxes,
                           float* out_scores, int64_t* out_classes, int32_t* out_count, int32_t* out_overflow,
                           void* workspace, size_t workspace_bytes, void* stream) {
  SDB_REQUIRE(levels && n_levels >= 1 && n_levels <= PP_MAX_LEVELS, SDB_ERR_INVALID, "need 1..%d levels", PP_MAX_LEVELS);
  SDB_REQUIRE(n_images >= 1 && n_images <= PP_MAX_IMAGES, SDB_ERR_INVALID, "need 1..%d images per call", PP_MAX_IMAGES);
  SDB_REQUIRE(num_classes >= 1 && num_points >= 1 && transform >= 0 && transform <= 2, SDB_ERR_INVALID, "bad head description");
  SDB_REQUIRE(transform != 1 || num_points >= 4, SDB_ERR_INVALID, "partial_minmax needs at least 4 points");
  SDB_REQUIRE(transform != 2 || (moment_transfer && num_points >= 2), SDB_ERR_INVALID, "moment transform needs moment_transfer");
  SDB_REQUIRE(topk >= 1 && topk <= 2048 && (long long)n_levels * topk <= 8192, SDB_ERR_UNSUPPORTED,
              "topk must be <= 2048 and n_levels * topk <= 8192");
  SDB_REQUIRE(max_det >= 1 && image_sizes && out_boxes && out_scores && out_classes && out_count, SDB_ERR_INVALID, "NULL argument");
  PPParams p{};
  p.nlv = n_levels; p.K = num_classes; p.P2 = 2 * num_points; p.transform = transform;
  p.mt_w = transform == 2 ? expf(moment_transfer[0]) : 1.f;
  p.mt_h = transform == 2 ? expf(moment_transfer[1]) : 1.f;
  p.thr = score_thresh; p.topk = topk; p.max_det = max_det; p.nms_thr = nms_thresh;
  pp_bins(score_thresh, p.thr_bits, p.shift, p.nbins);
  int blocks = 0;
  for (int l = 0; l < n_levels; ++l) {
    const int HW = levels[l].H * levels[l].W;
    SDB_REQUIRE(levels[l].cls && levels[l].pts && levels[l].centers && HW > 0, SDB_ERR_INVALID, "level %d: NULL tensor or empty map", l);
    SDB_REQUIRE((long long)HW * num_classes < (1LL << 32), SDB_ERR_UNSUPPORTED, "level %d too large", l);
    p.lv[l] = PPLevel{levels[l].cls, levels[l].pts, levels[l].centers, HW, levels[l].stride};
    p.blk_start[l] = blocks;
    blocks += (int)(((long long)HW * num_classes + 256 * PP_ELEMS - 1) / (256 * PP_ELEMS));
  }
  p.blk_start[n_levels] = blocks;
  for (int i = 0; i < n_images; ++i) { p.img_h[i] = image_sizes[2 * i]; p.img_w[i] = image_sizes[2 * i + 1]; }
  const PPWs w = pp_ws(n_levels, n_images, topk, p.nbins);
  SDB_REQUIRE(workspace && workspace_bytes >= w.bytes, SDB_ERR_WORKSPACE, "postprocess workspace too small: %zu < %zu",
              workspace_bytes, w.bytes);
  uint8_t* b = (uint8_t*)workspace;
  p.hist = (int*)(b + w.hist); p.cut = (int*)(b + w.cut); p.cnt = (int*)(b + w.cnt); p.overflow = (int*)(b + w.overflow);
  p.keys = (unsigned long long*)(b + w.keys); p.lbox = (float4*)(b + w.lbox); p.lscore = (float*)(b + w.lscore);
  p.lcls = (int*)(b + w.lcls); p.lcount = (int*)(b + w.lcount); p.sbox = (float4*)(b + w.sbox);
  p.sscore = (float*)(b + w.sscore); p.scls = (int*)(b + w.scls); p.total = (int*)(b + w.total);
  p.mask = (unsigned long long*)(b + w.mask);
  p.M = n_levels * topk; p.MW = (p.M + 63) / 64;
  cudaStream_t st = (cudaStream_t)stream;
  // hist .. overflow are contiguous at the start of the workspace: one fill
  SDB_CHECK_CUDA(cudaMemsetAsync(b, 0, w.keys, st));
  SDB_CHECK_CUDA(cudaMemsetAsync(p.mask, 0, (size_t)n_images * p.M * p.MW * 8, st));
  dim3 sgrid(blocks, n_images);
  pp_scan_kernel<false><<<sgrid, 256, 0, st>>>(p);
  pp_cut_kernel<<<dim3(n_levels, n_images), 32, 0, st>>>(p);
  pp_scan_kernel<true><<<sgrid, 256, 0, st>>>(p);
  pp_level_kernel<<<dim3(n_levels, n_images), 1024, 0, st>>>(p);
  int sortn = 64;
  while (sortn < p.M) sortn <<= 1;
  SDB_ENSURE_SMEM(pp_merge_kernel, (size_t)sortn * 8);
  pp_merge_kernel<<<n_images, 1024, (size_t)sortn * 8, st>>>(p, sortn);
  pp_nms_mask_kernel<<<dim3(p.MW, p.MW, n_images), 64, 0, st>>>(p);
  pp_nms_scan_kernel<<<n_images, 128, 0, st>>>(p, out_boxes, out_scores, (long long*)out_classes, out_count);
  SDB_LAUNCHED(7);
  if (out_overflow) SDB_CHECK_CUDA(cudaMemcpyAsync(out_overflow, p.overflow, 4, cudaMemcpyDeviceToDevice, st));
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // extern "C"
