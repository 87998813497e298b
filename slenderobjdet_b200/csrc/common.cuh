// common.cuh -- shared host/device helpers for libslender_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/slender_b200.h"

namespace sdb {

// thread-local error text behind sdb_last_error()
void set_error(const char* fmt, ...);

// every kernel launch goes through this counter (sdb_launch_count)
extern long long g_launches;
#define SDB_LAUNCHED(n) (sdb::g_launches += (n))

#define SDB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      sdb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                     __LINE__);                                                           \
      return SDB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define SDB_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      sdb::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// Geometry with derived output size, passed by value to kernels.
struct Geo {
  int N, C, H, W, O, KH, KW, sh, sw, ph, pw, dh, dw, groups, dgroups, Ho, Wo;
  int th, tw;  // tensor-core path: output pixels are walked in th x tw blocks (see decode_q)
  __host__ __device__ int taps() const { return KH * KW; }
  __host__ __device__ int HWo() const { return Ho * Wo; }
  __host__ __device__ long long P() const { return (long long)N * Ho * Wo; }
};

inline Geo make_geo(const sdb_dcn_geom& g) {
  Geo d{g.N, g.C_in, g.H, g.W, g.C_out, g.kH, g.kW, g.sH, g.sW, g.pH, g.pW, g.dH, g.dW,
        g.groups, g.deformable_groups, 0, 0, 8, 16};
  d.Ho = (g.H + 2 * g.pH - (g.dH * (g.kH - 1) + 1)) / (g.sH > 0 ? g.sH : 1) + 1;
  d.Wo = (g.W + 2 * g.pW - (g.dW * (g.kW - 1) + 1)) / (g.sW > 0 ? g.sW : 1) + 1;
  return d;
}

int check_geom(const sdb_dcn_geom* g);  // shape_check restated; sets error text

// RAII event pair around a launcher's dominant kernel (no-op unless sdb_profile_enable(1))
struct ProfScope {
  int slot;
  cudaStream_t st;
  int id;
  ProfScope(int slot, cudaStream_t st);
  ~ProfScope();
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (kernel, device) and raised only when a launch
// needs more: the attribute call is a driver round trip on every launch otherwise (the public-API step is
// host-bound).  Implemented in api.cu.
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);
#define SDB_ENSURE_SMEM(kernel, bytes) SDB_CHECK_CUDA(sdb::ensure_dynamic_smem((const void*)(kernel), (bytes)))

// ---- launchers implemented per translation unit ------------------------------------------------
// SIMT fp32 path (dcn_simt.cu)
int simt_forward(const float* x, const float* off, const float* mask, const float* w,
                 const float* bias, float* out, const Geo& g, cudaStream_t st);
int simt_backward_data(const float* x, const float* off, const float* mask, const float* w,
                       const float* gy, float* gx, float* goff, float* gmask, const Geo& g,
                       cudaStream_t st);
int simt_backward_weight(const float* x, const float* off, const float* mask, const float* gy,
                         float* gw, float* gb, float scale, const Geo& g, cudaStream_t st);

// tcgen05 bf16 path (dcn_tc.cu, dcn_tc_bwd.cu; multi-problem launchers declared in dcn_tc_shared.cuh)
bool tc_supported(const Geo& g, const char** why);
size_t tc_packed_input_bytes(const Geo& g);
size_t tc_columns_bytes(const Geo& g);

}  // namespace sdb
