// gn.cu -- GroupNorm + ReLU of the head towers, forward and backward, fused and batched over FPN levels.
//
// The RepPoints / FCOS towers are stacks of  Conv2d(3x3, bias=False) -> GroupNorm(32, C) -> ReLU(inplace)
// (sd/modeling/meta_arch/reppoints/reppointsv2.py:644-675, applied per FPN level :733-736; fcos.py:494-538).  In eager
// PyTorch that is, per level and layer, a group_norm kernel pair plus a ReLU pass forward and three more passes
// backward.  Here ONE call covers every level (and both towers) of a layer:
//   forward : gn_stats_kernel   (per (image, group) slice: sum / sum of squares, split over CTAs, fp64 combine)
//             gn_apply_kernel   (y = relu((x - mean) * rstd * gamma + beta); also leaves mean / rstd for the backward)
//   backward: gn_bwd_sums_kernel  (per (image, channel) plane: a = sum dy', b = sum dy' * xhat, dy' = dy * [y > 0];
//                                  the ReLU mask is recomputed from x, so the activation need not be kept for it)
//             gn_bwd_apply_kernel (dx = rstd * (gamma * dy' - (xhat * s2 + s1) / m), s1 / s2 = group sums of gamma * a / b)
//             gn_bwd_params_kernel(dgamma / dbeta: fixed-order sums over images and levels -- bit-reproducible)
// All HBM-bound: forward reads x twice (the second pass hits L2 for all but the largest levels) and writes y once;
// backward reads (x, dy) twice and writes dx once.  Tensors are NCHW, float32 or bfloat16; statistics in fp32 / fp64.
// Semantics: torch.nn.functional.group_norm (biased variance, eps inside the square root) followed by relu.
#include "common.cuh"
#include "tc_common.cuh"
#include "dcn_tc_shared.cuh"

namespace sdb {
namespace {
using tcshared::TileMap;
using tcshared::find_range;
using tcshared::MAX_PROBS;

constexpr int MAX_PARAMS = 4;
constexpr int SLICE_ELEMS = 16384;   // elements of one (image, group) slice handled by one CTA of the two-pass kernels

struct GnEntry {
  const void* x;
  void* y;
  const void* gy;
  void* gx;
  float* stats;      // [N][G][2]: mean, rstd
  int N, HW, param, splits;
  int vec_slice;     // (image, group) slices start 16-byte aligned and hold whole vectors (all pointers aligned)
  int vec_plane;     // the same for (image, channel) planes
  long long part0;   // first partial of this tensor in the workspace
};
struct GnTable {
  TileMap map;       // CTA -> tensor
  GnEntry e[MAX_PROBS];
  const float* gamma[MAX_PARAMS];
  const float* beta[MAX_PARAMS];
  float* ggamma[MAX_PARAMS];
  float* gbeta[MAX_PARAMS];
  int np, C, G, relu;
  float eps;
};

template <typename T> __device__ __forceinline__ float ld(const T* p, long long i);
template <> __device__ __forceinline__ float ld<float>(const float* p, long long i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void st(T* p, long long i, float v);
template <> __device__ __forceinline__ void st<float>(float* p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

// 16-byte vectors: 8 bf16 / 4 fp32 per load (a warp instruction moves 512 contiguous bytes)
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
    v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    t.x = tc::pack_bf16x2(v[0], v[1]); t.y = tc::pack_bf16x2(v[2], v[3]);
    t.z = tc::pack_bf16x2(v[4], v[5]); t.w = tc::pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// y = fmaf(x, a, b) with a = rstd * gamma, b = beta - mean * a: the scale / shift form (one FMA per element).  The forward
// output, the ReLU mask recomputed by both backward kernels and xhat all come from these helpers, so the mask the backward
// sees is bit for bit the sign the forward saw.
struct Affine { float a, b; };
__device__ __forceinline__ Affine affine_of(float mean, float rstd, float gamma, float beta) {
  Affine f;
  f.a = rstd * gamma;
  f.b = fmaf(-mean, f.a, beta);
  return f;
}

// block-wide sums of two doubles (valid on every thread)
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[32], sb[32];
  __shared__ double ra, rb;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d);
    b += __shfl_xor_sync(0xffffffffu, b, d);
  }
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();   // previous use of the shared slots is over
  if ((threadIdx.x & 31) == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int i = 0; i < nw; ++i) { ta += sa[i]; tb += sb[i]; }
    ra = ta; rb = tb;
  }
  __syncthreads();
  a = ra; b = rb;
}

// CTA -> (tensor, image, group, split); the slice [lo, hi) of that (image, group)'s cpg * HW contiguous elements
struct Slice {
  int ei, n, g, split;
  long long base;   // element offset of the (image, group) block in the tensor
  int lo, hi, m;    // m = cpg * HW
};
__device__ __forceinline__ Slice slice_of(const GnTable& t) {
  Slice s;
  s.ei = find_range(t.map, blockIdx.x);
  const GnEntry& e = t.e[s.ei];
  const int local = blockIdx.x - t.map.start[s.ei];
  s.split = local % e.splits;
  const int ng = local / e.splits;
  s.g = ng % t.G;
  s.n = ng / t.G;
  const int cpg = t.C / t.G;
  s.m = cpg * e.HW;
  s.base = ((long long)s.n * t.C + (long long)s.g * cpg) * e.HW;
  const int per = ((s.m + e.splits - 1) / e.splits + 7) & ~7;   // whole 16-byte vectors
  s.lo = min(s.m, s.split * per);
  s.hi = min(s.m, s.lo + per);
  return s;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const __grid_constant__ GnTable t, double* __restrict__ part) {
  const Slice s = slice_of(t);
  const GnEntry& e = t.e[s.ei];
  const T* x = (const T*)e.x + s.base;
  float fs = 0.f, fq = 0.f;
  double ds = 0, dq = 0;
  int k = 0;
  if (e.vec_slice) {
    constexpr int VN = Vec<T>::N;
    for (int i = s.lo + threadIdx.x * VN; i < s.hi; i += 256 * VN) {
      float v[8];
      Vec<T>::load(x + i, v);
#pragma unroll
      for (int j = 0; j < VN; ++j) { fs += v[j]; fq = fmaf(v[j], v[j], fq); }
      if (++k == 4) { ds += fs; dq += fq; fs = fq = 0.f; k = 0; }   // short fp32 runs, fp64 across them
    }
  } else {
    for (int i = s.lo + threadIdx.x; i < s.hi; i += 256) {
      const float v = ld<T>(x, i);
      fs += v;
      fq = fmaf(v, v, fq);
      if (++k == 16) { ds += fs; dq += fq; fs = fq = 0.f; k = 0; }
    }
  }
  ds += fs; dq += fq;
  block_sum2(ds, dq);
  if (threadIdx.x == 0) {
    double* p = part + 2 * (e.part0 + ((long long)s.n * t.G + s.g) * e.splits + s.split);
    p[0] = ds; p[1] = dq;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const __grid_constant__ GnTable t, const double* __restrict__ part) {
  const Slice s = slice_of(t);
  const GnEntry& e = t.e[s.ei];
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    const double* p = part + 2 * (e.part0 + ((long long)s.n * t.G + s.g) * e.splits);
    double a = 0, q = 0;
    for (int k = 0; k < e.splits; ++k) { a += p[2 * k]; q += p[2 * k + 1]; }
    const double mean = a / s.m;
    double var = q / s.m - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)t.eps));
    sh[0] = (float)mean; sh[1] = rstd;
    if (s.split == 0 && e.stats) {
      e.stats[2 * (s.n * t.G + s.g)] = (float)mean;
      e.stats[2 * (s.n * t.G + s.g) + 1] = rstd;
    }
  }
  __syncthreads();
  const float mean = sh[0], rstd = sh[1];
  const int cpg = t.C / t.G, HW = e.HW;
  const float* gamma = t.gamma[e.param] + s.g * cpg;
  const float* beta = t.beta[e.param] + s.g * cpg;
  const T* x = (const T*)e.x + s.base;
  T* y = (T*)e.y + s.base;
  if (e.vec_slice) {
    constexpr int VN = Vec<T>::N;
    for (int i = s.lo + threadIdx.x * VN; i < s.hi; i += 256 * VN) {
      float v[8];
      Vec<T>::load(x + i, v);
      int c = i / HW, left = HW - (i - c * HW);   // elements of channel c from i on
      if (left >= VN) {                            // the whole vector lies in one channel (all but one vector per plane)
        const Affine f = affine_of(mean, rstd, __ldg(gamma + c), __ldg(beta + c));
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float r = fmaf(v[j], f.a, f.b);
          v[j] = t.relu ? fmaxf(r, 0.f) : r;
        }
      } else {
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          while (left == 0) { ++c; left = HW; }
          const Affine f = affine_of(mean, rstd, __ldg(gamma + c), __ldg(beta + c));
          const float r = fmaf(v[j], f.a, f.b);
          v[j] = t.relu ? fmaxf(r, 0.f) : r;
          --left;
        }
      }
      Vec<T>::store(y + i, v);
    }
  } else {
    for (int i = s.lo + threadIdx.x; i < s.hi; i += 256) {
      const int c = i / HW;
      const Affine f = affine_of(mean, rstd, __ldg(gamma + c), __ldg(beta + c));
      const float v = fmaf(ld<T>(x, i), f.a, f.b);
      st<T>(y, i, t.relu ? fmaxf(v, 0.f) : v);
    }
  }
}

// backward, pass 1: per (tensor, image, channel) plane: a = sum dy', b = sum dy' * xhat
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_sums_kernel(const __grid_constant__ GnTable t, float* __restrict__ ab) {
  const int ei = find_range(t.map, blockIdx.x);
  const GnEntry& e = t.e[ei];
  const int local = blockIdx.x - t.map.start[ei];   // n * C + c
  const int c = local % t.C, n = local / t.C;
  const int cpg = t.C / t.G, g = c / cpg;
  const float mean = e.stats[2 * (n * t.G + g)], rstd = e.stats[2 * (n * t.G + g) + 1];
  const Affine f = affine_of(mean, rstd, __ldg(t.gamma[e.param] + c), __ldg(t.beta[e.param] + c));
  const float nmr = -mean * rstd;   // xhat = fmaf(x, rstd, nmr)
  const T* x = (const T*)e.x + (long long)local * e.HW;
  const T* gy = (const T*)e.gy + (long long)local * e.HW;
  float fa = 0.f, fb = 0.f;
  double da = 0, db = 0;
  int k = 0;
  if (e.vec_plane) {
    constexpr int VN = Vec<T>::N;
    for (int i = threadIdx.x * VN; i < e.HW; i += 256 * VN) {
      float xv[8], dv[8];
      Vec<T>::load(x + i, xv);
      Vec<T>::load(gy + i, dv);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const float d = (t.relu && !(fmaf(xv[j], f.a, f.b) > 0.f)) ? 0.f : dv[j];
        fa += d;
        fb = fmaf(d, fmaf(xv[j], rstd, nmr), fb);
      }
      if (++k == 4) { da += fa; db += fb; fa = fb = 0.f; k = 0; }
    }
  } else {
    for (int i = threadIdx.x; i < e.HW; i += 256) {
      const float xs = ld<T>(x, i);
      const float d = (t.relu && !(fmaf(xs, f.a, f.b) > 0.f)) ? 0.f : ld<T>(gy, i);
      fa += d;
      fb = fmaf(d, fmaf(xs, rstd, nmr), fb);
      if (++k == 16) { da += fa; db += fb; fa = fb = 0.f; k = 0; }
    }
  }
  da += fa; db += fb;
  block_sum2(da, db);
  if (threadIdx.x == 0) {
    float* p = ab + 2 * (e.part0 + local);
    p[0] = (float)da; p[1] = (float)db;
  }
}

// backward, pass 2: dx over (image, group) slices
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const __grid_constant__ GnTable t, const float* __restrict__ ab) {
  const Slice s = slice_of(t);
  const GnEntry& e = t.e[s.ei];
  if (!e.gx) return;
  const int cpg = t.C / t.G, HW = e.HW;
  const float* gamma = t.gamma[e.param] + s.g * cpg;
  const float* beta = t.beta[e.param] + s.g * cpg;
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    const float* p = ab + 2 * (e.part0 + (long long)s.n * t.C + s.g * cpg);
    double s1 = 0, s2 = 0;
    for (int c = 0; c < cpg; ++c) { s1 += (double)gamma[c] * p[2 * c]; s2 += (double)gamma[c] * p[2 * c + 1]; }
    sh[0] = (float)(s1 / s.m); sh[1] = (float)(s2 / s.m);
  }
  __syncthreads();
  const float m1 = sh[0], m2 = sh[1];
  const float mean = e.stats[2 * (s.n * t.G + s.g)], rstd = e.stats[2 * (s.n * t.G + s.g) + 1];
  const float k1 = rstd * m1, k2 = rstd * m2, nmr = -mean * rstd;   // dx = a * dy' - k2 * xhat - k1
  const T* x = (const T*)e.x + s.base;
  const T* gy = (const T*)e.gy + s.base;
  T* gx = (T*)e.gx + s.base;
  if (e.vec_slice) {
    constexpr int VN = Vec<T>::N;
    for (int i = s.lo + threadIdx.x * VN; i < s.hi; i += 256 * VN) {
      float xv[8], dv[8];
      Vec<T>::load(x + i, xv);
      Vec<T>::load(gy + i, dv);
      int c = i / HW, left = HW - (i - c * HW);
      if (left >= VN) {
        const Affine f = affine_of(mean, rstd, __ldg(gamma + c), __ldg(beta + c));
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float d = (t.relu && !(fmaf(xv[j], f.a, f.b) > 0.f)) ? 0.f : dv[j];
          dv[j] = fmaf(f.a, d, fmaf(-k2, fmaf(xv[j], rstd, nmr), -k1));
        }
      } else {
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          while (left == 0) { ++c; left = HW; }
          const Affine f = affine_of(mean, rstd, __ldg(gamma + c), __ldg(beta + c));
          const float d = (t.relu && !(fmaf(xv[j], f.a, f.b) > 0.f)) ? 0.f : dv[j];
          dv[j] = fmaf(f.a, d, fmaf(-k2, fmaf(xv[j], rstd, nmr), -k1));
          --left;
        }
      }
      Vec<T>::store(gx + i, dv);
    }
  } else {
    for (int i = s.lo + threadIdx.x; i < s.hi; i += 256) {
      const int c = i / HW;
      const Affine f = affine_of(mean, rstd, __ldg(gamma + c), __ldg(beta + c));
      const float xs = ld<T>(x, i);
      const float d = (t.relu && !(fmaf(xs, f.a, f.b) > 0.f)) ? 0.f : ld<T>(gy, i);
      st<T>(gx, i, fmaf(f.a, d, fmaf(-k2, fmaf(xs, rstd, nmr), -k1)));
    }
  }
}

// backward, pass 3: dgamma[c] = sum over tensors of the parameter set, images: b;  dbeta[c] = ... a.  grid (np), fixed order.
__global__ void __launch_bounds__(256) gn_bwd_params_kernel(const __grid_constant__ GnTable t, const float* __restrict__ ab) {
  const int pid = blockIdx.x;
  if (!t.ggamma[pid] && !t.gbeta[pid]) return;
  for (int c = threadIdx.x; c < t.C; c += blockDim.x) {
    double a = 0, b = 0;
    for (int ei = 0; ei < t.map.n; ++ei) {
      const GnEntry& e = t.e[ei];
      if (e.param != pid) continue;
      for (int n = 0; n < e.N; ++n) {
        const float* p = ab + 2 * (e.part0 + (long long)n * t.C + c);
        a += p[0]; b += p[1];
      }
    }
    if (t.ggamma[pid]) t.ggamma[pid][c] += (float)b;
    if (t.gbeta[pid]) t.gbeta[pid][c] += (float)a;
  }
}

int splits_of(int m) {
  int s = (m + SLICE_ELEMS - 1) / SLICE_ELEMS;
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}

// builds the table; `planes`: CTA = (image, channel) plane (backward sums) instead of (image, group, split) slices
int build_table(const sdb_gn_tensor* ts, int n, const sdb_gn_params* ps, int np, int C, int G, float eps, int relu, bool planes,
                int elem_bytes, GnTable& t, long long* nparts) {
  SDB_REQUIRE(ts && n >= 1 && n <= MAX_PROBS, SDB_ERR_INVALID, "need 1..%d tensors, got %d", MAX_PROBS, n);
  SDB_REQUIRE(np >= 1 && np <= MAX_PARAMS, SDB_ERR_INVALID, "need 1..%d parameter sets, got %d", MAX_PARAMS, np);
  SDB_REQUIRE(C > 0 && G > 0 && C % G == 0, SDB_ERR_INVALID, "num_channels %d must be divisible by num_groups %d", C, G);
  t = GnTable{};
  t.np = np; t.C = C; t.G = G; t.relu = relu; t.eps = eps;
  for (int k = 0; k < np && ps; ++k) {
    t.gamma[k] = ps[k].gamma; t.beta[k] = ps[k].beta; t.ggamma[k] = ps[k].grad_gamma; t.gbeta[k] = ps[k].grad_beta;
  }
  int m = 0, total = 0;
  long long parts = 0;
  for (int i = 0; i < n; ++i) {
    SDB_REQUIRE(ts[i].N >= 0 && ts[i].HW > 0, SDB_ERR_INVALID, "tensor %d: bad size N=%d HW=%d", i, ts[i].N, ts[i].HW);
    SDB_REQUIRE(ts[i].param_id >= 0 && ts[i].param_id < np, SDB_ERR_INVALID, "tensor %d: param_id %d out of range", i, ts[i].param_id);
    SDB_REQUIRE((long long)(C / G) * ts[i].HW < (1LL << 31), SDB_ERR_UNSUPPORTED, "tensor %d: group slice too large", i);
    if (ts[i].N == 0) continue;
    GnEntry& e = t.e[m];
    e.x = ts[i].x; e.y = ts[i].y; e.gy = ts[i].grad_y; e.gx = ts[i].grad_x; e.stats = ts[i].stats;
    e.N = ts[i].N; e.HW = ts[i].HW; e.param = ts[i].param_id;
    e.splits = splits_of(C / G * ts[i].HW);
    {
      const size_t ptrs = (size_t)e.x | (size_t)e.y | (size_t)e.gy | (size_t)e.gx;
      const int vn = 16 / elem_bytes;
      e.vec_slice = (ptrs & 15) == 0 && ((long long)(C / G) * e.HW) % vn == 0;
      e.vec_plane = (ptrs & 15) == 0 && e.HW % vn == 0;
    }
    // one partial numbering serves both passes of a direction: forward (image, group, split), backward (image, channel)
    e.part0 = parts;
    const long long fwd = (long long)e.N * G * e.splits, bwd = (long long)e.N * C;
    parts += fwd > bwd ? fwd : bwd;
    t.map.start[m] = total;
    total += planes ? e.N * C : e.N * G * e.splits;
    ++m;
  }
  t.map.n = m; t.map.start[m] = total;
  *nparts = parts;
  return SDB_OK;
}

// the (image, group, split) CTA map of a table built with planes = true
void to_slices(GnTable& t) {
  int total = 0;
  for (int i = 0; i < t.map.n; ++i) {
    t.map.start[i] = total;
    total += t.e[i].N * t.G * t.e[i].splits;
  }
  t.map.start[t.map.n] = total;
}

}  // namespace
}  // namespace sdb

using namespace sdb;

extern "C" {

size_t sdb_gn_relu_workspace_bytes(const sdb_gn_tensor* tensors, int32_t n, int32_t C, int32_t G) {
  GnTable t;
  long long parts = 0;
  sdb_gn_params dummy[MAX_PARAMS] = {};
  int np = 1;
  for (int i = 0; tensors && i < n && i < MAX_PROBS; ++i)
    if (tensors[i].param_id >= np && tensors[i].param_id < MAX_PARAMS) np = tensors[i].param_id + 1;
  if (build_table(tensors, n, dummy, np, C, G, 1e-5f, 1, false, 2, t, &parts)) return 0;
  return (size_t)parts * 2 * sizeof(double) + 256;
}

int sdb_gn_relu_forward(const sdb_gn_tensor* tensors, int32_t n, const sdb_gn_params* params, int32_t np, int32_t C, int32_t G,
                        float eps, int32_t relu, int io_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  SDB_REQUIRE(io_dtype == SDB_F32 || io_dtype == SDB_BF16, SDB_ERR_INVALID, "unknown io_dtype %d", io_dtype);
  SDB_REQUIRE(params != nullptr, SDB_ERR_INVALID, "NULL parameter table");
  GnTable t;
  long long parts = 0;
  int rc = build_table(tensors, n, params, np, C, G, eps, relu, false, io_dtype == SDB_BF16 ? 2 : 4, t, &parts);
  if (rc) return rc;
  for (int k = 0; k < np; ++k) SDB_REQUIRE(params[k].gamma && params[k].beta, SDB_ERR_INVALID, "parameter set %d: gamma and beta must be non-NULL", k);
  for (int i = 0; i < t.map.n; ++i) SDB_REQUIRE(t.e[i].x && t.e[i].y, SDB_ERR_INVALID, "x and y must be non-NULL");
  const int total = t.map.start[t.map.n];
  if (total == 0) return SDB_OK;
  SDB_REQUIRE(workspace && workspace_bytes >= (size_t)parts * 2 * sizeof(double), SDB_ERR_WORKSPACE, "GroupNorm workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = (double*)workspace;
  if (io_dtype == SDB_BF16) {
    gn_stats_kernel<__nv_bfloat16><<<total, 256, 0, st>>>(t, part);
    gn_apply_kernel<__nv_bfloat16><<<total, 256, 0, st>>>(t, part);
  } else {
    gn_stats_kernel<float><<<total, 256, 0, st>>>(t, part);
    gn_apply_kernel<float><<<total, 256, 0, st>>>(t, part);
  }
  SDB_LAUNCHED(2);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

int sdb_gn_relu_backward(const sdb_gn_tensor* tensors, int32_t n, const sdb_gn_params* params, int32_t np, int32_t C, int32_t G,
                         float eps, int32_t relu, int io_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  SDB_REQUIRE(io_dtype == SDB_F32 || io_dtype == SDB_BF16, SDB_ERR_INVALID, "unknown io_dtype %d", io_dtype);
  SDB_REQUIRE(params != nullptr, SDB_ERR_INVALID, "NULL parameter table");
  GnTable t;
  long long parts = 0;
  int rc = build_table(tensors, n, params, np, C, G, eps, relu, true, io_dtype == SDB_BF16 ? 2 : 4, t, &parts);
  if (rc) return rc;
  for (int k = 0; k < np; ++k) SDB_REQUIRE(params[k].gamma && params[k].beta, SDB_ERR_INVALID, "parameter set %d: gamma and beta must be non-NULL", k);
  for (int i = 0; i < t.map.n; ++i)
    SDB_REQUIRE(t.e[i].x && t.e[i].gy && t.e[i].stats, SDB_ERR_INVALID, "x, grad_y and stats must be non-NULL");
  const int planes = t.map.start[t.map.n];
  if (planes == 0) return SDB_OK;
  SDB_REQUIRE(workspace && workspace_bytes >= (size_t)parts * 2 * sizeof(float), SDB_ERR_WORKSPACE, "GroupNorm workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* ab = (float*)workspace;
  if (io_dtype == SDB_BF16) gn_bwd_sums_kernel<__nv_bfloat16><<<planes, 256, 0, st>>>(t, ab);
  else gn_bwd_sums_kernel<float><<<planes, 256, 0, st>>>(t, ab);
  gn_bwd_params_kernel<<<np, 256, 0, st>>>(t, ab);
  to_slices(t);
  const int total = t.map.start[t.map.n];
  if (io_dtype == SDB_BF16) gn_bwd_apply_kernel<__nv_bfloat16><<<total, 256, 0, st>>>(t, ab);
  else gn_bwd_apply_kernel<float><<<total, 256, 0, st>>>(t, ab);
  SDB_LAUNCHED(3);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // extern "C"
