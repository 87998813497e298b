/*
 * oracle/dcn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's deformable-convolution algorithm
 * (DCN v1 and modulated v2, forward and backward).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.  The
 * product path (slenderobjdet_b200/) never links or calls it.
 *
 * Algorithm followed (paths relative to /root/reference/detectron2/detectron2/layers/csrc/deformable/):
 *   bilinear sample ............ deform_conv_cuda_kernel.cu:96-130
 *   input-gradient weight ...... deform_conv_cuda_kernel.cu:132-161
 *   coordinate-gradient weight . deform_conv_cuda_kernel.cu:163-214
 *   im2col indexing / validity . deform_conv_cuda_kernel.cu:216-288 (v1), :785-868 (v2, x mask)
 *   col2im (grad_input) ........ deform_conv_cuda_kernel.cu:291-363 (v1), :870-949 (v2)
 *   grad_offset / grad_mask .... deform_conv_cuda_kernel.cu:366-452 (v1), :951-1066 (v2)
 *   GEMMs (out, dcol, dW, db) .. deform_conv_cuda.cu:397-409, :553-559, :769-778, :1101-1114
 *
 * Parity pin: checked against the reference's only DCN known-answer test
 * (tests/test_deformable_conv.py:69-87) and against torchvision.ops.deform_conv2d
 * (the "CPU deform_conv2d path" BASELINE.json names) in tests/test_oracle_dcn.py.
 *
 * Numerics: sampling arithmetic is done in float exactly in the reference's
 * operation order; the GEMM reductions accumulate in double and round once, so
 * the oracle is independent of summation order.
 *
 * Layouts (all contiguous, float32):
 *   x       [N, C, H, W]
 *   offset  [N, dg*2*KH*KW, Ho, Wo]   channel 2*(i*KW+j) = dy, +1 = dx, per deformable group
 *   mask    [N, dg*KH*KW,   Ho, Wo]   or NULL (v1)
 *   weight  [O, C/groups, KH, KW]
 *   bias    [O] or NULL
 *   out     [N, O, Ho, Wo]
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int N, C, H, W, O, KH, KW, sh, sw, ph, pw, dh, dw, groups, dgroups, Ho, Wo;
} dcn_dims;

static dcn_dims mk(int N, int C, int H, int W, int O, int KH, int KW, int sh, int sw, int ph,
                   int pw, int dh, int dw, int groups, int dgroups) {
  dcn_dims d = {N, C, H, W, O, KH, KW, sh, sw, ph, pw, dh, dw, groups, dgroups, 0, 0};
  d.Ho = (H + 2 * ph - (dh * (KH - 1) + 1)) / sh + 1;
  d.Wo = (W + 2 * pw - (dw * (KW - 1) + 1)) / sw + 1;
  return d;
}

/* deform_conv_cuda_kernel.cu:96-130 */
static float bilinear(const float* im, int H, int W, float h, float w) {
  int h_low = (int)floorf(h), w_low = (int)floorf(w);
  int h_high = h_low + 1, w_high = w_low + 1;
  float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = im[h_low * W + w_low];
  if (h_low >= 0 && w_high <= W - 1) v2 = im[h_low * W + w_high];
  if (h_high <= H - 1 && w_low >= 0) v3 = im[h_high * W + w_low];
  if (h_high <= H - 1 && w_high <= W - 1) v4 = im[h_high * W + w_high];
  float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* deform_conv_cuda_kernel.cu:163-214; bp_dir 0 = d/dy, 1 = d/dx */
static float coord_weight(const float* im, int H, int W, float h, float w, int bp_dir) {
  if (h <= -1 || h >= H || w <= -1 || w >= W) return 0;
  int h_low = (int)floorf(h), w_low = (int)floorf(w);
  int h_high = h_low + 1, w_high = w_low + 1;
  float weight = 0;
  if (bp_dir == 0) {
    if (h_low >= 0 && w_low >= 0) weight += -1 * (w_low + 1 - w) * im[h_low * W + w_low];
    if (h_low >= 0 && w_high <= W - 1) weight += -1 * (w - w_low) * im[h_low * W + w_high];
    if (h_high <= H - 1 && w_low >= 0) weight += (w_low + 1 - w) * im[h_high * W + w_low];
    if (h_high <= H - 1 && w_high <= W - 1) weight += (w - w_low) * im[h_high * W + w_high];
  } else {
    if (h_low >= 0 && w_low >= 0) weight += -1 * (h_low + 1 - h) * im[h_low * W + w_low];
    if (h_low >= 0 && w_high <= W - 1) weight += (h_low + 1 - h) * im[h_low * W + w_high];
    if (h_high <= H - 1 && w_low >= 0) weight += -1 * (h - h_low) * im[h_high * W + w_low];
    if (h_high <= H - 1 && w_high <= W - 1) weight += (h - h_low) * im[h_high * W + w_high];
  }
  return weight;
}

/* sampling position of (output pixel, tap): deform_conv_cuda_kernel.cu:251-272 */
static void sample_pos(const dcn_dims* d, const float* offset, int n, int g, int i, int j, int ho,
                       int wo, float* h_im, float* w_im) {
  const int K2 = d->KH * d->KW, HW = d->Ho * d->Wo;
  const float* off = offset + ((size_t)(n * d->dgroups + g) * 2 * K2) * HW;
  float oh = off[(size_t)(2 * (i * d->KW + j)) * HW + ho * d->Wo + wo];
  float ow = off[(size_t)(2 * (i * d->KW + j) + 1) * HW + ho * d->Wo + wo];
  *h_im = (ho * d->sh - d->ph) + i * d->dh + oh;
  *w_im = (wo * d->sw - d->pw) + j * d->dw + ow;
}

/* columns[(c*KH*KW + tap), (ho,wo)] for one image: deform_conv_cuda_kernel.cu:216-288 / :785-868 */
static void im2col_one(const dcn_dims* d, const float* x, const float* offset, const float* mask,
                       int n, float* col) {
  const int K2 = d->KH * d->KW, HW = d->Ho * d->Wo, cpg = d->C / d->dgroups;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < d->C; ++c) {
    const int g = c / cpg;
    const float* im = x + ((size_t)n * d->C + c) * d->H * d->W;
    for (int i = 0; i < d->KH; ++i)
      for (int j = 0; j < d->KW; ++j) {
        float* dst = col + ((size_t)c * K2 + i * d->KW + j) * HW;
        for (int ho = 0; ho < d->Ho; ++ho)
          for (int wo = 0; wo < d->Wo; ++wo) {
            float h_im, w_im, val = 0;
            sample_pos(d, offset, n, g, i, j, ho, wo, &h_im, &w_im);
            if (h_im > -1 && w_im > -1 && h_im < d->H && w_im < d->W)
              val = bilinear(im, d->H, d->W, h_im, w_im);
            if (mask)
              val *= mask[((size_t)(n * d->dgroups + g) * K2 + i * d->KW + j) * HW + ho * d->Wo + wo];
            dst[ho * d->Wo + wo] = val;
          }
      }
  }
}

int dcn_oracle_out_hw(int H, int W, int KH, int KW, int sh, int sw, int ph, int pw, int dh, int dw,
                      int* Ho, int* Wo) {
  dcn_dims d = mk(1, 1, H, W, 1, KH, KW, sh, sw, ph, pw, dh, dw, 1, 1);
  *Ho = d.Ho;
  *Wo = d.Wo;
  return 0;
}

/* out = W . col (+ bias): deform_conv_cuda.cu:397-409, :904-926 */
int dcn_oracle_forward(const float* x, const float* offset, const float* mask, const float* weight,
                       const float* bias, float* out, int N, int C, int H, int W, int O, int KH,
                       int KW, int sh, int sw, int ph, int pw, int dh, int dw, int groups,
                       int dgroups) {
  dcn_dims d = mk(N, C, H, W, O, KH, KW, sh, sw, ph, pw, dh, dw, groups, dgroups);
  if (d.Ho <= 0 || d.Wo <= 0 || C % groups || O % groups || C % dgroups) return -1;
  const int K2 = KH * KW, HW = d.Ho * d.Wo, Cg = C / groups, Og = O / groups, Kg = Cg * K2;
  float* col = (float*)malloc((size_t)C * K2 * HW * sizeof(float));
  if (!col) return -2;
  for (int n = 0; n < N; ++n) {
    im2col_one(&d, x, offset, mask, n, col);
#pragma omp parallel for schedule(static)
    for (int o = 0; o < O; ++o) {
      const int g = o / Og;
      const float* wrow = weight + (size_t)o * Kg;
      const float* cg = col + (size_t)g * Kg * HW;
      double* acc = (double*)calloc(HW, sizeof(double));
      for (int k = 0; k < Kg; ++k) {
        const double wv = wrow[k];
        const float* cr = cg + (size_t)k * HW;
        for (int p = 0; p < HW; ++p) acc[p] += wv * cr[p];
      }
      float* dst = out + ((size_t)n * O + o) * HW;
      for (int p = 0; p < HW; ++p) dst[p] = (float)(acc[p] + (bias ? (double)bias[o] : 0.0));
      free(acc);
    }
  }
  free(col);
  return 0;
}

/*
 * All five gradients.  Any output pointer may be NULL (skipped).  Outputs are
 * OVERWRITTEN (the oracle does not accumulate into caller buffers).
 */
int dcn_oracle_backward(const float* x, const float* offset, const float* mask,
                        const float* weight, const float* grad_out, float* grad_x,
                        float* grad_offset, float* grad_mask, float* grad_weight,
                        float* grad_bias, int N, int C, int H, int W, int O, int KH, int KW,
                        int sh, int sw, int ph, int pw, int dh, int dw, int groups, int dgroups) {
  dcn_dims d = mk(N, C, H, W, O, KH, KW, sh, sw, ph, pw, dh, dw, groups, dgroups);
  if (d.Ho <= 0 || d.Wo <= 0 || C % groups || O % groups || C % dgroups) return -1;
  const int K2 = KH * KW, HW = d.Ho * d.Wo, Cg = C / groups, Og = O / groups, Kg = Cg * K2;
  const int cpg = C / dgroups;
  const size_t ncol = (size_t)C * K2 * HW;
  float* col = (float*)malloc(ncol * sizeof(float));
  float* dcol = (float*)malloc(ncol * sizeof(float));
  double* gx = grad_x ? (double*)calloc((size_t)N * C * H * W, sizeof(double)) : NULL;
  double* gw = grad_weight ? (double*)calloc((size_t)O * Kg, sizeof(double)) : NULL;
  double* gb = grad_bias ? (double*)calloc(O, sizeof(double)) : NULL;
  if (!col || !dcol) return -2;

  for (int n = 0; n < N; ++n) {
    const float* dy = grad_out + (size_t)n * O * HW;
    /* dcol = W^T . dY  (deform_conv_cuda.cu:553-559) */
#pragma omp parallel for schedule(static)
    for (int kk = 0; kk < C * K2; ++kk) {
      const int c = kk / K2, g = c / Cg, k = (c - g * Cg) * K2 + kk % K2;
      double* acc = (double*)calloc(HW, sizeof(double));
      for (int o = g * Og; o < (g + 1) * Og; ++o) {
        const double wv = weight[(size_t)o * Kg + k];
        const float* dr = dy + (size_t)o * HW;
        for (int p = 0; p < HW; ++p) acc[p] += wv * dr[p];
      }
      float* dst = dcol + (size_t)kk * HW;
      for (int p = 0; p < HW; ++p) dst[p] = (float)acc[p];
      free(acc);
    }

    /* grad_offset / grad_mask: one value per (dg, tap, dir, pixel), summed over the
       group's channels (deform_conv_cuda_kernel.cu:366-452, :951-1066) */
    if (grad_offset || grad_mask) {
#pragma omp parallel for schedule(static)
      for (int gt = 0; gt < dgroups * K2; ++gt) {
        const int g = gt / K2, tap = gt % K2, i = tap / KW, j = tap % KW;
        for (int ho = 0; ho < d.Ho; ++ho)
          for (int wo = 0; wo < d.Wo; ++wo) {
            float h_im, w_im;
            sample_pos(&d, offset, n, g, i, j, ho, wo, &h_im, &w_im);
            const int inside = !(h_im <= -1 || w_im <= -1 || h_im >= H || w_im >= W);
            const float m =
                mask ? mask[((size_t)(n * dgroups + g) * K2 + tap) * HW + ho * d.Wo + wo] : 1.f;
            double gy = 0, gxx = 0, gm = 0;
            if (inside)
              for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float* im = x + ((size_t)n * C + c) * H * W;
                const float dc = dcol[((size_t)c * K2 + tap) * HW + ho * d.Wo + wo];
                gy += (double)(coord_weight(im, H, W, h_im, w_im, 0) * dc * m);
                gxx += (double)(coord_weight(im, H, W, h_im, w_im, 1) * dc * m);
                gm += (double)(dc * bilinear(im, H, W, h_im, w_im));
              }
            const size_t pix = (size_t)ho * d.Wo + wo;
            if (grad_offset) {
              float* go = grad_offset + ((size_t)(n * dgroups + g) * 2 * K2) * HW;
              go[(size_t)(2 * tap) * HW + pix] = (float)gy;
              go[(size_t)(2 * tap + 1) * HW + pix] = (float)gxx;
            }
            if (grad_mask)
              grad_mask[((size_t)(n * dgroups + g) * K2 + tap) * HW + pix] = (float)gm;
          }
      }
    }

    /* grad_input: bilinear scatter of dcol (x mask) (deform_conv_cuda_kernel.cu:291-363, :132-161) */
    if (gx) {
#pragma omp parallel for schedule(static)
      for (int c = 0; c < C; ++c) {
        const int g = c / cpg;
        double* gim = gx + ((size_t)n * C + c) * H * W;
        for (int tap = 0; tap < K2; ++tap) {
          const int i = tap / KW, j = tap % KW;
          for (int ho = 0; ho < d.Ho; ++ho)
            for (int wo = 0; wo < d.Wo; ++wo) {
              float h_im, w_im;
              sample_pos(&d, offset, n, g, i, j, ho, wo, &h_im, &w_im);
              if (h_im <= -1 || w_im <= -1 || h_im >= H || w_im >= W) continue;
              float top = dcol[((size_t)c * K2 + tap) * HW + ho * d.Wo + wo];
              if (mask)
                top *= mask[((size_t)(n * dgroups + g) * K2 + tap) * HW + ho * d.Wo + wo];
              const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
              for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                  const int hh = h_low + a, ww = w_low + b;
                  if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                  /* get_gradient_weight, the four cases of :151-159 */
                  const float wh = a ? (h_im + 1 - hh) : (hh + 1 - h_im);
                  const float wwt = b ? (w_im + 1 - ww) : (ww + 1 - w_im);
                  gim[hh * W + ww] += (double)(wh * wwt * top);
                }
            }
        }
      }
    }

    /* grad_weight += dY . col^T ; grad_bias += sum dY  (deform_conv_cuda.cu:769-778, :1101-1114) */
    if (gw) {
      im2col_one(&d, x, offset, mask, n, col);
#pragma omp parallel for schedule(static)
      for (int o = 0; o < O; ++o) {
        const int g = o / Og;
        const float* dr = dy + (size_t)o * HW;
        for (int k = 0; k < Kg; ++k) {
          const float* cr = col + ((size_t)g * Kg + k) * HW;
          double acc = 0;
          for (int p = 0; p < HW; ++p) acc += (double)dr[p] * cr[p];
          gw[(size_t)o * Kg + k] += acc;
        }
      }
    }
    if (gb)
      for (int o = 0; o < O; ++o) {
        double acc = 0;
        for (int p = 0; p < HW; ++p) acc += dy[(size_t)o * HW + p];
        gb[o] += acc;
      }
  }

  if (gx) {
    for (size_t t = 0; t < (size_t)N * C * H * W; ++t) grad_x[t] = (float)gx[t];
    free(gx);
  }
  if (gw) {
    for (size_t t = 0; t < (size_t)O * Kg; ++t) grad_weight[t] = (float)gw[t];
    free(gw);
  }
  if (gb) {
    for (int o = 0; o < O; ++o) grad_bias[o] = (float)gb[o];
    free(gb);
  }
  free(col);
  free(dcol);
  return 0;
}
