// dcn_simt.cu -- exact-fp32 SIMT deformable convolution (any shape / groups / deformable groups).
//
// This is the SDB_MATH_FP32 path: fused gather + register-tiled SGEMM, no column buffer in HBM.
// It exists for (a) fp32 parity at rel <= 1e-4 (BASELINE.json config 1) and (b) every geometry the
// tensor-core path does not cover.  The tcgen05 path lives in dcn_tc_*.cu.
//
// Semantics restated from the reference kernels (d2/layers/csrc/deformable/deform_conv_cuda_kernel.cu):
//   bilinear :96-130, validity test :273, gradient weight :132-161, coordinate weight :163-214,
//   grad_offset :366-452, grad_mask :1028-1064.  The reference materialises `columns` and calls
//   cuBLAS; here the sampled tile is produced straight into shared memory.
#include "common.cuh"

namespace sdb {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct Pix {
  int n, ho, wo, valid;
};

__device__ __forceinline__ Pix decode_pixel(const Geo& g, long long p) {
  Pix q;
  q.valid = p < g.P();
  const int hw = g.Ho * g.Wo;
  const long long pp = q.valid ? p : 0;
  q.n = (int)(pp / hw);
  const int r = (int)(pp - (long long)q.n * hw);
  q.ho = r / g.Wo;
  q.wo = r - q.ho * g.Wo;
  return q;
}

__device__ __forceinline__ void sample_pos(const Geo& g, const float* __restrict__ off, int n,
                                           int dg, int tap, int ho, int wo, float& h_im,
                                           float& w_im) {
  const int hw = g.Ho * g.Wo, k2 = g.KH * g.KW;
  const float* o = off + ((size_t)(n * g.dgroups + dg) * 2 * k2 + 2 * tap) * hw + ho * g.Wo + wo;
  const int i = tap / g.KW, j = tap - i * g.KW;
  h_im = (float)(ho * g.sh - g.ph + i * g.dh) + __ldg(o);
  w_im = (float)(wo * g.sw - g.pw + j * g.dw) + __ldg(o + hw);
}

__device__ __forceinline__ bool inside(const Geo& g, float h, float w) {
  return h > -1.f && w > -1.f && h < (float)g.H && w < (float)g.W;
}

// 4 corners + weights of a valid sampling position; invalid corners get index -1.
struct Corners {
  int idx[4];
  float wt[4];     // bilinear weights (hh*hw, hh*lw, lh*hw, lh*lw)
  float lh, lw;
};

__device__ __forceinline__ Corners corners_of(const Geo& g, float h, float w) {
  Corners c;
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  c.lh = h - h_low;
  c.lw = w - w_low;
  const float hh = 1.f - c.lh, hw = 1.f - c.lw;
  c.wt[0] = hh * hw;
  c.wt[1] = hh * c.lw;
  c.wt[2] = c.lh * hw;
  c.wt[3] = c.lh * c.lw;
  const bool t = h_low >= 0, b = h_high <= g.H - 1, l = w_low >= 0, r = w_high <= g.W - 1;
  c.idx[0] = (t && l) ? h_low * g.W + w_low : -1;
  c.idx[1] = (t && r) ? h_low * g.W + w_high : -1;
  c.idx[2] = (b && l) ? h_high * g.W + w_low : -1;
  c.idx[3] = (b && r) ? h_high * g.W + w_high : -1;
  return c;
}

__device__ __forceinline__ float bilinear(const float* __restrict__ im, const Corners& c) {
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = c.idx[q] >= 0 ? __ldg(im + c.idx[q]) : 0.f;
  return c.wt[0] * v[0] + c.wt[1] * v[1] + c.wt[2] * v[2] + c.wt[3] * v[3];
}

// one im2col element: x sampled at (pixel, channel c, tap), times mask
__device__ __forceinline__ float col_value(const Geo& g, const float* __restrict__ x,
                                           const float* __restrict__ off,
                                           const float* __restrict__ mask, const Pix& q, int c,
                                           int tap) {
  if (!q.valid) return 0.f;
  const int dg = c / (g.C / g.dgroups);
  float h, w;
  sample_pos(g, off, q.n, dg, tap, q.ho, q.wo, h, w);
  if (!inside(g, h, w)) return 0.f;
  const Corners cs = corners_of(g, h, w);
  float v = bilinear(x + ((size_t)q.n * g.C + c) * g.H * g.W, cs);
  if (mask)
    v *= __ldg(mask + ((size_t)(q.n * g.dgroups + dg) * g.KH * g.KW + tap) * g.Ho * g.Wo +
               q.ho * g.Wo + q.wo);
  return v;
}

// ------------------------------------------------------------------------------------------------
// forward: out[p, o] = sum_k col[p, k] * W[o, k]      tile 64 px x 64 o, 4x4 per thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) simt_fwd_kernel(const float* __restrict__ x,
                                                      const float* __restrict__ off,
                                                      const float* __restrict__ mask,
                                                      const float* __restrict__ w,
                                                      const float* __restrict__ bias,
                                                      float* __restrict__ out, Geo g) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k2 = g.KH * g.KW, Og = g.O / g.groups, Cg = g.C / g.groups, Kg = Cg * k2;
  const int otiles = (Og + BN - 1) / BN;
  const int grp = blockIdx.y / otiles, o0 = (blockIdx.y - grp * otiles) * BN;
  const long long p0 = (long long)blockIdx.x * BM;

  const Pix my = decode_pixel(g, p0 + (tid & 63));  // pixel this thread samples for
  float acc[4][4] = {};

  for (int k0 = 0; k0 < Kg; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = (tid >> 6) + 4 * r, k = k0 + kk;
      float v = 0.f;
      if (k < Kg) {
        const int cl = k / k2, tap = k - cl * k2;
        v = col_value(g, x, off, mask, my, grp * Cg + cl, tap);
      }
      As[kk][tid & 63] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int oo = tid >> 2, kk = (tid & 3) * 4 + r, k = k0 + kk;
      Bs[kk][oo] = (k < Kg && o0 + oo < Og) ? __ldg(w + (size_t)(grp * Og + o0 + oo) * Kg + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const Pix q = decode_pixel(g, p0 + tx * 4 + i);
    if (!q.valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ol = o0 + ty * 4 + j;
      if (ol >= Og) continue;
      const int o = grp * Og + ol;
      out[((size_t)q.n * g.O + o) * g.Ho * g.Wo + q.ho * g.Wo + q.wo] =
          acc[i][j] + (bias ? __ldg(bias + o) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward data: dcol[p, k] = sum_o dY[p, o] W[o, k]; then per element: scatter to grad_x,
// reduce into grad_offset / grad_mask.  Tile 64 px x 64 k.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) simt_bwd_data_kernel(
    const float* __restrict__ x, const float* __restrict__ off, const float* __restrict__ mask,
    const float* __restrict__ w, const float* __restrict__ gy, float* __restrict__ gx,
    float* __restrict__ goff, float* __restrict__ gmask, Geo g) {
  __shared__ __align__(16) float As[BK][BM];  // dY chunk [o][pixel]
  __shared__ __align__(16) float Bs[BK][BN];  // W chunk  [o][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k2 = g.KH * g.KW, Og = g.O / g.groups, Cg = g.C / g.groups, Kg = Cg * k2;
  const int ktiles = (Kg + BN - 1) / BN;
  const int grp = blockIdx.y / ktiles, k0 = (blockIdx.y - grp * ktiles) * BN;
  const long long p0 = (long long)blockIdx.x * BM;
  const int hw = g.Ho * g.Wo;

  const Pix my = decode_pixel(g, p0 + (tid & 63));
  float acc[4][4] = {};
  for (int oc = 0; oc < Og; oc += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = (tid >> 6) + 4 * r, ol = oc + kk;
      As[kk][tid & 63] =
          (my.valid && ol < Og)
              ? __ldg(gy + ((size_t)my.n * g.O + grp * Og + ol) * hw + my.ho * g.Wo + my.wo)
              : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = (tid >> 6) + 4 * r, ol = oc + kk, kc = tid & 63;
      Bs[kk][kc] = (ol < Og && k0 + kc < Kg) ? __ldg(w + (size_t)(grp * Og + ol) * Kg + k0 + kc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int cpd = g.C / g.dgroups;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const Pix q = decode_pixel(g, p0 + tx * 4 + i);
    if (!q.valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + ty * 4 + j;
      if (k >= Kg) continue;
      const int cl = k / k2, tap = k - cl * k2, c = grp * Cg + cl, dg = c / cpd;
      float h, wv;
      sample_pos(g, off, q.n, dg, tap, q.ho, q.wo, h, wv);
      if (!inside(g, h, wv)) continue;  // weights are 0 outside (:140-144, :172-176, :435-437)
      const Corners cs = corners_of(g, h, wv);
      const float* im = x + ((size_t)q.n * g.C + c) * g.H * g.W;
      float v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t] = cs.idx[t] >= 0 ? __ldg(im + cs.idx[t]) : 0.f;
      const size_t mpos = ((size_t)(q.n * g.dgroups + dg) * k2 + tap) * hw + q.ho * g.Wo + q.wo;
      const float m = mask ? __ldg(mask + mpos) : 1.f;
      const float dc = acc[i][j], dcm = dc * m;
      if (gx) {
        float* gim = gx + ((size_t)q.n * g.C + c) * g.H * g.W;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (cs.idx[t] >= 0) atomicAdd(gim + cs.idx[t], cs.wt[t] * dcm);
      }
      if (goff) {
        // d(bilinear)/dh and /dw (get_coordinate_weight :185-211)
        const float dh_ = -(1.f - cs.lw) * v[0] - cs.lw * v[1] + (1.f - cs.lw) * v[2] + cs.lw * v[3];
        const float dw_ = -(1.f - cs.lh) * v[0] + (1.f - cs.lh) * v[1] - cs.lh * v[2] + cs.lh * v[3];
        float* go = goff + ((size_t)(q.n * g.dgroups + dg) * 2 * k2 + 2 * tap) * hw + q.ho * g.Wo + q.wo;
        atomicAdd(go, dh_ * dcm);
        atomicAdd(go + hw, dw_ * dcm);
      }
      if (gmask)
        atomicAdd(gmask + mpos,
                  dc * (cs.wt[0] * v[0] + cs.wt[1] * v[1] + cs.wt[2] * v[2] + cs.wt[3] * v[3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward weight: dW[o, k] += scale * sum_p dY[p, o] col[p, k].  Tile 64 o x 64 k, pixels split
// over blockIdx.z, fp32 atomics to combine the splits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) simt_bwd_weight_kernel(
    const float* __restrict__ x, const float* __restrict__ off, const float* __restrict__ mask,
    const float* __restrict__ gy, float* __restrict__ gw, float scale, Geo g, int pix_per_split) {
  __shared__ __align__(16) float As[BK][BM];  // dY chunk  [pixel][o]
  __shared__ __align__(16) float Bs[BK][BN];  // col chunk [pixel][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k2 = g.KH * g.KW, Og = g.O / g.groups, Cg = g.C / g.groups, Kg = Cg * k2;
  const int otiles = (Og + BM - 1) / BM, ktiles = (Kg + BN - 1) / BN;
  const int grp = blockIdx.x / otiles, o0 = (blockIdx.x - grp * otiles) * BM;
  const int k0 = blockIdx.y * BN;
  (void)ktiles;
  const long long pbeg = (long long)blockIdx.z * pix_per_split;
  long long pend = pbeg + pix_per_split;
  if (pend > g.P()) pend = g.P();
  const int hw = g.Ho * g.Wo;

  float acc[4][4] = {};
  for (long long pc = pbeg; pc < pend; pc += BK) {
    const long long p = pc + (tid & 15);
    Pix q = decode_pixel(g, p);
    q.valid = q.valid && p < pend;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int oo = (tid >> 4) + 16 * r;
      As[tid & 15][oo] =
          (q.valid && o0 + oo < Og)
              ? __ldg(gy + ((size_t)q.n * g.O + grp * Og + o0 + oo) * hw + q.ho * g.Wo + q.wo)
              : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kc = (tid >> 4) + 16 * r, k = k0 + kc;
      float v = 0.f;
      if (k < Kg) {
        const int cl = k / k2, tap = k - cl * k2;
        v = col_value(g, x, off, mask, q, grp * Cg + cl, tap);
      }
      Bs[tid & 15][kc] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ol = o0 + tx * 4 + i;
    if (ol >= Og) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + ty * 4 + j;
      if (k < Kg) atomicAdd(gw + (size_t)(grp * Og + ol) * Kg + k, scale * acc[i][j]);
    }
  }
}

__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ gy,
                                                        float* __restrict__ gb, float scale,
                                                        int N, int O, int hw) {
  const int o = blockIdx.x;
  float s = 0.f;
  for (int n = 0; n < N; ++n)
    for (int i = threadIdx.x; i < hw; i += blockDim.x) s += gy[((size_t)n * O + o) * hw + i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(gb + o, scale * t);
  }
}

}  // namespace

int bias_grad_f32(const float* gy, float* gb, float scale, int N, int O, int hw, cudaStream_t st) {
  bias_grad_kernel<<<O, 256, 0, st>>>(gy, gb, scale, N, O, hw); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

int simt_forward(const float* x, const float* off, const float* mask, const float* w,
                 const float* bias, float* out, const Geo& g, cudaStream_t st) {
  const int Og = g.O / g.groups;
  dim3 grid(cdiv(g.P(), BM), g.groups * cdiv(Og, BN));
  ProfScope prof(SDB_OP_FORWARD, st);
  simt_fwd_kernel<<<grid, NT, 0, st>>>(x, off, mask, w, bias, out, g); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

int simt_backward_data(const float* x, const float* off, const float* mask, const float* w,
                       const float* gy, float* gx, float* goff, float* gmask, const Geo& g,
                       cudaStream_t st) {
  const int k2 = g.taps(), Kg = g.C / g.groups * k2;
  if (goff)
    SDB_CHECK_CUDA(cudaMemsetAsync(goff, 0, sizeof(float) * g.P() * g.dgroups * 2 * k2, st));
  if (gmask && mask)
    SDB_CHECK_CUDA(cudaMemsetAsync(gmask, 0, sizeof(float) * g.P() * g.dgroups * k2, st));
  dim3 grid(cdiv(g.P(), BM), g.groups * cdiv(Kg, BN));
  ProfScope prof(SDB_OP_BACKWARD_DATA, st);
  simt_bwd_data_kernel<<<grid, NT, 0, st>>>(x, off, mask, w, gy, gx, goff, mask ? gmask : nullptr, g); SDB_LAUNCHED(1);
  SDB_CHECK_CUDA(cudaGetLastError());
  return SDB_OK;
}

int bias_grad_f32(const float* gy, float* gb, float scale, int N, int O, int hw, cudaStream_t st);

int simt_backward_weight(const float* x, const float* off, const float* mask, const float* gy,
                         float* gw, float* gb, float scale, const Geo& g, cudaStream_t st) {
  const int k2 = g.taps(), Og = g.O / g.groups, Kg = g.C / g.groups * k2;
  if (gw) {
    const int tiles = g.groups * cdiv(Og, BM) * cdiv(Kg, BN);
    int splits = (int)((4 * 148 + tiles - 1) / tiles);  // ~4 waves of CTAs
    const long long P = g.P();
    if (splits > cdiv(P, 4 * BK)) splits = cdiv(P, 4 * BK);
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int pps = cdiv(P, splits);
    pps = (pps + BK - 1) / BK * BK;
    splits = cdiv(P, pps);
    dim3 grid(g.groups * cdiv(Og, BM), cdiv(Kg, BN), splits);
    ProfScope prof(SDB_OP_BACKWARD_WEIGHT, st);
    simt_bwd_weight_kernel<<<grid, NT, 0, st>>>(x, off, mask, gy, gw, scale, g, pps); SDB_LAUNCHED(1);
    SDB_CHECK_CUDA(cudaGetLastError());
  }
  if (gb) return bias_grad_f32(gy, gb, scale, g.N, g.O, g.HWo(), st);
  return SDB_OK;
}

}  // namespace sdb
