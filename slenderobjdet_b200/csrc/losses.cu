// losses.cu -- fused forward+gradient head losses (sum reduction), HBM-bound, no tensor cores.
//
//   sdb_sigmoid_focal_loss : fvcore sigmoid_focal_loss_jit(logits, one_hot, alpha, gamma, "sum")
//       without the dense one-hot target of sd/modeling/meta_arch/reppoints/reppointsv2.py:294-312
//       / fcos/fcos.py:289-297 -- one pass: read logits, write gradient, one atomic per CTA.
//   sdb_box_reg_loss       : sd/layers/iou_loss.py:4-77 (ltrb / xyxy; iou, linear_iou, giou),
//       sd/layers/smooth_l1_loss_with_weight.py:3-17, fvcore giou_loss
//       (meta/heads/anchor_head.py:369-376), with hand-derived gradients.
#include "common.cuh"

namespace sdb {
namespace {

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float part[32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) part[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < nw; ++i) t += part[i];
  return t;  // valid on thread 0
}

__device__ __forceinline__ void focal_elem(float x, bool pos, float alpha, float gamma, float& loss,
                                           float& grad) {
  // p = sigmoid(x); ce = BCE-with-logits(x, t); stable in fp32
  const float e = __expf(-fabsf(x));
  const float inv = 1.f / (1.f + e);
  const float p = x >= 0.f ? inv : e * inv;
  const float sp = log1pf(e);                     // softplus(-|x|)
  const float ce = fmaxf(x, 0.f) - (pos ? x : 0.f) + sp;
  const float q = pos ? 1.f - p : p;              // 1 - p_t
  const float mod = gamma == 2.f ? q * q : (q > 0.f ? powf(q, gamma) : 0.f);
  const float a = alpha >= 0.f ? (pos ? alpha : 1.f - alpha) : 1.f;
  loss = a * ce * mod;
  // d/dx: pos: a*mod*(-gamma*p*ce - (1-p));  neg: a*mod*(p + gamma*(1-p)*ce)
  grad = pos ? a * mod * (-gamma * p * ce - (1.f - p)) : a * mod * (p + gamma * (1.f - p) * ce);
}

template <int VEC>
__global__ void __launch_bounds__(256) focal_kernel(const float* __restrict__ logits,
                                                    const int64_t* __restrict__ cls, long long R,
                                                    int K, float alpha, float gamma, float gscale,
                                                    float* __restrict__ loss_sum,
                                                    float* __restrict__ grad) {
  const long long total = R * K / VEC;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e0 = i * VEC;
    const long long r = e0 / K;
    const int k0 = (int)(e0 - r * K);
    const int c = (int)__ldg(cls + r);  // values outside [0,K) = background
    float xs[VEC], gs[VEC];
    if (VEC == 4) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(logits) + i);
      xs[0] = v.x; xs[1] = v.y; xs[2] = v.z; xs[3] = v.w;
    } else {
      xs[0] = __ldcs(logits + i);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float l;
      focal_elem(xs[j], (k0 + j) == c, alpha, gamma, l, gs[j]);
      acc += l;
      gs[j] *= gscale;
    }
    if (grad) {
      if (VEC == 4)
        __stcs(reinterpret_cast<float4*>(grad) + i, make_float4(gs[0], gs[1], gs[2], gs[3]));
      else
        __stcs(grad + i, gs[0]);
    }
  }
  const float t = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss_sum, t);
}

// step(a<b) with autograd's 0.5 split at ties (torch.min / torch.max backward)
__device__ __forceinline__ float lt(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }

__global__ void __launch_bounds__(256) box_loss_kernel(const float4* __restrict__ pred,
                                                       const float4* __restrict__ target,
                                                       const float* __restrict__ weight,
                                                       long long R, int kind, int form, float beta,
                                                       float gscale, float* __restrict__ loss_sum,
                                                       float4* __restrict__ grad) {
  float acc = 0.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R;
       r += (long long)gridDim.x * blockDim.x) {
    const float4 P = pred[r], T = target[r];
    const float wgt = weight ? weight[r] : 1.f;
    const float p[4] = {P.x, P.y, P.z, P.w}, t[4] = {T.x, T.y, T.z, T.w};
    float loss = 0.f, g[4] = {0.f, 0.f, 0.f, 0.f};
    if (kind == SDB_LOSS_SMOOTH_L1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = p[j] - t[j], n = fabsf(d);
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        if (beta < 1e-5f) {
          loss += n;
          g[j] = sgn;
        } else if (n < beta) {
          loss += 0.5f * n * n / beta;
          g[j] = d / beta;
        } else {
          loss += n - 0.5f * beta;
          g[j] = sgn;
        }
      }
    } else if (kind == SDB_LOSS_GIOU_FVCORE) {
      const float eps = 1e-7f;
      const float xk1 = fmaxf(p[0], t[0]), yk1 = fmaxf(p[1], t[1]);
      const float xk2 = fminf(p[2], t[2]), yk2 = fminf(p[3], t[3]);
      const bool ov = (yk2 > yk1) && (xk2 > xk1);
      const float I = ov ? (xk2 - xk1) * (yk2 - yk1) : 0.f;
      const float pw = p[2] - p[0], ph = p[3] - p[1];
      const float U = pw * ph + (t[2] - t[0]) * (t[3] - t[1]) - I;
      const float xc1 = fminf(p[0], t[0]), yc1 = fminf(p[1], t[1]);
      const float xc2 = fmaxf(p[2], t[2]), yc2 = fmaxf(p[3], t[3]);
      const float Ac = (xc2 - xc1) * (yc2 - yc1);
      const float iou = I / (U + eps);
      loss = 1.f - (iou - (Ac - U) / (Ac + eps));
      const float dpa[4] = {-ph, -pw, ph, pw};
      float dI[4] = {0.f, 0.f, 0.f, 0.f};
      if (ov) {
        dI[0] = -lt(t[0], p[0]) * (yk2 - yk1);
        dI[1] = -lt(t[1], p[1]) * (xk2 - xk1);
        dI[2] = lt(p[2], t[2]) * (yk2 - yk1);
        dI[3] = lt(p[3], t[3]) * (xk2 - xk1);
      }
      const float dAc[4] = {-lt(p[0], t[0]) * (yc2 - yc1), -lt(p[1], t[1]) * (xc2 - xc1),
                            lt(t[2], p[2]) * (yc2 - yc1), lt(t[3], p[3]) * (xc2 - xc1)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dU = dpa[j] - dI[j];
        const float diou = (dI[j] * (U + eps) - I * dU) / ((U + eps) * (U + eps));
        const float dT = ((dAc[j] - dU) * (Ac + eps) - (Ac - U) * dAc[j]) / ((Ac + eps) * (Ac + eps));
        g[j] = -(diou - dT);
      }
    } else {
      float ta, pa, wi, hi, gw, gh, dpa[4], dwi[4], dhi[4], dgw[4], dgh[4];
      if (form == SDB_BOX_LTRB) {  // (l, t, r, b) distances, iou_loss.py:4-22
        ta = (t[0] + t[2]) * (t[1] + t[3]);
        pa = (p[0] + p[2]) * (p[1] + p[3]);
        wi = fminf(p[0], t[0]) + fminf(p[2], t[2]);
        gw = fmaxf(p[0], t[0]) + fmaxf(p[2], t[2]);
        hi = fminf(p[3], t[3]) + fminf(p[1], t[1]);
        gh = fmaxf(p[3], t[3]) + fmaxf(p[1], t[1]);
        const float sw = p[0] + p[2], sh = p[1] + p[3];
        dpa[0] = sh; dpa[1] = sw; dpa[2] = sh; dpa[3] = sw;
        dwi[0] = lt(p[0], t[0]); dwi[1] = 0.f; dwi[2] = lt(p[2], t[2]); dwi[3] = 0.f;
        dhi[0] = 0.f; dhi[1] = lt(p[1], t[1]); dhi[2] = 0.f; dhi[3] = lt(p[3], t[3]);
        dgw[0] = lt(t[0], p[0]); dgw[1] = 0.f; dgw[2] = lt(t[2], p[2]); dgw[3] = 0.f;
        dgh[0] = 0.f; dgh[1] = lt(t[1], p[1]); dgh[2] = 0.f; dgh[3] = lt(t[3], p[3]);
      } else {  // (x1, y1, x2, y2), iou_loss.py:40-62
        ta = (t[2] - t[0]) * (t[3] - t[1]);
        const float pw = p[2] - p[0], ph = p[3] - p[1];
        pa = pw * ph;
        wi = fminf(p[2], t[2]) - fmaxf(p[0], t[0]);
        gw = fmaxf(p[2], t[2]) - fminf(p[0], t[0]);
        hi = fminf(p[3], t[3]) - fmaxf(p[1], t[1]);
        gh = fmaxf(p[3], t[3]) - fminf(p[1], t[1]);
        dpa[0] = -ph; dpa[1] = -pw; dpa[2] = ph; dpa[3] = pw;
        dwi[0] = -lt(t[0], p[0]); dwi[1] = 0.f; dwi[2] = lt(p[2], t[2]); dwi[3] = 0.f;
        dhi[0] = 0.f; dhi[1] = -lt(t[1], p[1]); dhi[2] = 0.f; dhi[3] = lt(p[3], t[3]);
        dgw[0] = -lt(p[0], t[0]); dgw[1] = 0.f; dgw[2] = lt(t[2], p[2]); dgw[3] = 0.f;
        dgh[0] = 0.f; dgh[1] = -lt(p[1], t[1]); dgh[2] = 0.f; dgh[3] = lt(t[3], p[3]);
      }
      const float ac = gw * gh + 1e-7f;
      const float ai = wi * hi;
      const float au = ta + pa - ai;
      const float iou = (ai + 1.f) / (au + 1.f);
      const float giou = iou - (ac - au) / ac;
      loss = kind == SDB_LOSS_IOU ? -logf(iou) : (kind == SDB_LOSS_LINEAR_IOU ? 1.f - iou : 1.f - giou);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dai = dwi[j] * hi + wi * dhi[j];
        const float dau = dpa[j] - dai;
        const float diou = (dai * (au + 1.f) - (ai + 1.f) * dau) / ((au + 1.f) * (au + 1.f));
        if (kind == SDB_LOSS_IOU) {
          g[j] = -diou / iou;
        } else if (kind == SDB_LOSS_LINEAR_IOU) {
          g[j] = -diou;
        } else {
          const float dac = dgw[j] * gh + gw * dgh[j];
          g[j] = -(diou + (dau * ac - au * dac) / (ac * ac));
        }
      }
    }
    acc += loss * wgt;
    if (grad) {
      const float s = wgt * gscale;
      grad[r] = make_float4(g[0] * s, g[1] * s, g[2] * s, g[3] * s);
    }
  }
  const float tsum = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss_sum, tsum);
}

}  // namespace
}  // namespace sdb

using namespace sdb;

namespace sdb {
namespace {
// compute_centerness_targets (fcos/utils.py:295-300): one thread per row, IEEE division / sqrt and no fma
// contraction, in the reference's operation order -> bit-identical to torch's elementwise result.
// SLENDER: the FCOSRepPoints module's own definition (fcos_rpd_s1_topk.py:25-55): pow(c, min(w/h, h/w)) with
// w = l + r, h = t + b (powf is within a few ulp of torch's CPU pow, not bit-identical).
template <bool SLENDER>
__global__ void __launch_bounds__(256) centerness_kernel(const float4* __restrict__ ltrb, long long R,
                                                         float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ltrb[i];   // (l, t, r, b)
    const float lr = __fdiv_rn(fminf(v.x, v.z), fmaxf(v.x, v.z));
    const float tb = __fdiv_rn(fminf(v.y, v.w), fmaxf(v.y, v.w));
    const float c = __fmul_rn(lr, tb);
    if (SLENDER) {
      const float r1 = __fdiv_rn(__fadd_rn(v.x, v.z), __fadd_rn(v.y, v.w));
      out[i] = powf(c, fminf(r1, __fdiv_rn(1.f, r1)));
    } else {
      out[i] = __fsqrt_rn(c);
    }
  }
}

// dcn_offset = ((1 - gm) * p + gm * p) - base, one rounding per operation as in the reference's four torch ops;
// BWD: grad_pts = gm * grad_out.  Channel c of the output reads channel c ^ flip of the input.
template <bool BWD>
__global__ void __launch_bounds__(256) reppoints_offset_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                               long long total, int hw, int ks, float a, float b,
                                                               int flip) {
  const int K2 = 2 * ks * ks;
  const float pad = (float)((ks - 1) / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long plane = i / hw;
    const int c = (int)(plane % K2);
    const long long j = flip ? (plane - c + (c ^ 1)) * hw + (i - plane * hw) : i;   // element of the other tensor
    if (BWD) {
      dst[j] = __fmul_rn(b, src[i]);   // i indexes grad_out (channel c), j the point channel it came from
    } else {
      const float pv = src[j];
      const int k = c >> 1;
      const float base = (c & 1) ? (float)(k % ks) - pad : (float)(k / ks) - pad;
      dst[i] = __fsub_rn(__fadd_rn(__fmul_rn(a, pv), __fmul_rn(b, pv)), base);
    }
  }
}
}  // namespace
}  // namespace sdb

extern "C" {

int sdb_sigmoid_focal_loss(const float* logits, const int64_t* class_idx, int64_t R, int32_t K,
                           float alpha, float gamma, float grad_scale, float* loss_sum,
                           float* grad_logits, void* stream) {
  SDB_REQUIRE(R >= 0 && K > 0, SDB_ERR_INVALID, "bad focal-loss shape R=%lld K=%d", (long long)R, K);
  SDB_REQUIRE(loss_sum != nullptr, SDB_ERR_INVALID, "loss_sum is NULL");
  if (R == 0) return SDB_OK;
  SDB_REQUIRE(logits && class_idx, SDB_ERR_INVALID, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)R * K;
  const bool vec = (K % 4 == 0) && (((uintptr_t)logits | (uintptr_t)grad_logits) % 16 == 0);
  const long long work = vec ? total / 4 : total;
  long long blocks = (work + 255) / 256;
  const long long cap = 148LL * 8;  // persistent grid-stride: 8 CTAs per SM
  if (blocks > cap) blocks = cap;
  if (vec)
    focal_kernel<4><<<(int)blocks, 256, 0, st>>>(logits, class_idx, R, K, alpha, gamma, grad_scale, loss_sum, grad_logits);
  else
    focal_kernel<1><<<(int)blocks, 256, 0, st>>>(logits, class_idx, R, K, alpha, gamma, grad_scale, loss_sum, grad_logits); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

int sdb_box_reg_loss(const float* pred, const float* target, const float* weight, int64_t R,
                     int kind, int form, float beta, float grad_scale, float* loss_sum,
                     float* grad_pred, void* stream) {
  SDB_REQUIRE(R >= 0, SDB_ERR_INVALID, "negative row count");
  SDB_REQUIRE(kind >= SDB_LOSS_IOU && kind <= SDB_LOSS_GIOU_FVCORE, SDB_ERR_INVALID, "unknown loss kind %d", kind);
  SDB_REQUIRE(form == SDB_BOX_LTRB || form == SDB_BOX_XYXY, SDB_ERR_INVALID, "unknown box form %d", form);
  SDB_REQUIRE(loss_sum != nullptr, SDB_ERR_INVALID, "loss_sum is NULL");
  if (R == 0) return SDB_OK;
  SDB_REQUIRE(pred && target, SDB_ERR_INVALID, "NULL argument");
  long long blocks = (R + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  box_loss_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const float4*)pred, (const float4*)target, weight, R, kind, form, beta, grad_scale, loss_sum,
      (float4*)grad_pred); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

static int centerness_launch(bool slender, const float* reg_targets, int64_t R, float* out, void* stream) {
  SDB_REQUIRE(R >= 0, SDB_ERR_INVALID, "negative row count");
  if (R == 0) return SDB_OK;
  SDB_REQUIRE(reg_targets && out, SDB_ERR_INVALID, "NULL argument");
  long long blocks = (R + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (slender) sdb::centerness_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)reg_targets, R, out);
  else         sdb::centerness_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)reg_targets, R, out);
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}
int sdb_centerness_targets(const float* reg_targets, int64_t R, float* out, void* stream) {
  return centerness_launch(false, reg_targets, R, out, stream);
}
int sdb_slender_centerness_targets(const float* reg_targets, int64_t R, float* out, void* stream) {
  return centerness_launch(true, reg_targets, R, out, stream);
}

static int reppoints_offset_launch(bool bwd, const float* src, int32_t N, int32_t ks, int32_t H, int32_t W, float gm,
                                   int32_t flip, float* dst, void* stream) {
  SDB_REQUIRE(N >= 0 && ks > 0 && H > 0 && W > 0, SDB_ERR_INVALID, "bad sizes N=%d ks=%d H=%d W=%d", N, ks, H, W);
  const long long total = (long long)N * 2 * ks * ks * H * W;
  if (total == 0) return SDB_OK;
  SDB_REQUIRE(src && dst, SDB_ERR_INVALID, "NULL argument");
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const float a = (float)(1.0 - (double)gm), b = gm;   // python: (1 - gm) in double, each scalar cast to float32 by the op
  if (bwd) sdb::reppoints_offset_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, total, H * W, ks, a, b, flip != 0);
  else     sdb::reppoints_offset_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, total, H * W, ks, a, b, flip != 0);
  SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}
int sdb_reppoints_dcn_offset(const float* pts, int32_t N, int32_t ks, int32_t H, int32_t W, float gradient_mul,
                             int32_t flip_xy, float* out, void* stream) {
  return reppoints_offset_launch(false, pts, N, ks, H, W, gradient_mul, flip_xy, out, stream);
}
int sdb_reppoints_dcn_offset_backward(const float* grad_out, int32_t N, int32_t ks, int32_t H, int32_t W,
                                      float gradient_mul, int32_t flip_xy, float* grad_pts, void* stream) {
  return reppoints_offset_launch(true, grad_out, N, ks, H, W, gradient_mul, flip_xy, grad_pts, stream);
}

}  // extern "C"
