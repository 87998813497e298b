// api.cu -- extern "C" entry points of libslender_b200.so (see include/slender_b200.h).
#include <stdarg.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sdb {

static thread_local char g_err[512] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
  struct Entry { const void* k; int dev; size_t bytes; };
  static Entry tab[256];
  static int n = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n; ++i)
    if (tab[i].k == kernel && tab[i].dev == dev) {
      if (bytes <= tab[i].bytes) return cudaSuccess;
      const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e == cudaSuccess) tab[i].bytes = bytes;
      return e;
    }
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && n < 256) tab[n++] = Entry{kernel, dev, bytes};
  return e;
}

// shape_check restated (d2/layers/csrc/deformable/deform_conv_cuda.cu:140-270)
int check_geom(const sdb_dcn_geom* g) {
  SDB_REQUIRE(g != nullptr, SDB_ERR_INVALID, "geometry pointer is NULL");
  SDB_REQUIRE(g->N >= 0 && g->C_in > 0 && g->H > 0 && g->W > 0 && g->C_out > 0, SDB_ERR_INVALID,
              "non-positive tensor size: N=%d C_in=%d H=%d W=%d C_out=%d", g->N, g->C_in, g->H,
              g->W, g->C_out);
  SDB_REQUIRE(g->kH > 0 && g->kW > 0, SDB_ERR_INVALID,
              "kernel size should be greater than zero, but got kH: %d kW: %d", g->kH, g->kW);
  SDB_REQUIRE(g->sH > 0 && g->sW > 0, SDB_ERR_INVALID,
              "stride should be greater than zero, but got dH: %d dW: %d", g->sH, g->sW);
  SDB_REQUIRE(g->dH > 0 && g->dW > 0, SDB_ERR_INVALID,
              "dilation should be greater than 0, but got dilationH: %d dilationW: %d", g->dH, g->dW);
  SDB_REQUIRE(g->pH >= 0 && g->pW >= 0, SDB_ERR_INVALID, "negative padding");
  SDB_REQUIRE(g->groups > 0 && g->deformable_groups > 0, SDB_ERR_INVALID, "groups must be positive");
  SDB_REQUIRE(g->C_in % g->groups == 0 && g->C_out % g->groups == 0, SDB_ERR_INVALID,
              "channels (%d -> %d) not divisible by groups %d", g->C_in, g->C_out, g->groups);
  SDB_REQUIRE(g->C_in % g->deformable_groups == 0, SDB_ERR_INVALID,
              "input channels must divide deformable group size");
  const Geo d = make_geo(*g);
  SDB_REQUIRE(d.Ho > 0 && d.Wo > 0, SDB_ERR_INVALID,
              "Given input size: (%d x %d x %d). Calculated output size: (%d x %d x %d). Output size is too small",
              g->C_in, g->H, g->W, g->C_out, d.Ho, d.Wo);
  return SDB_OK;
}

static int check_io(int io_dtype, int math) {
  SDB_REQUIRE(io_dtype == SDB_F32 || io_dtype == SDB_BF16, SDB_ERR_INVALID, "unknown io_dtype %d", io_dtype);
  SDB_REQUIRE(math == SDB_MATH_FP32 || math == SDB_MATH_BF16, SDB_ERR_INVALID, "unknown math mode %d", math);
  SDB_REQUIRE(!(math == SDB_MATH_FP32 && io_dtype != SDB_F32), SDB_ERR_UNSUPPORTED,
              "SDB_MATH_FP32 needs float32 tensors (bf16 tensors use SDB_MATH_BF16)");
  return SDB_OK;
}

static int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return SDB_ERR_CUDA;
  }
  static thread_local int checked_dev = -1;
  if (checked_dev != dev) {
    cudaDeviceProp p;
    SDB_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
    SDB_REQUIRE(p.major == 10, SDB_ERR_UNSUPPORTED,
                "libslender_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
    checked_dev = dev;
  }
  return SDB_OK;
}

// ---- profiling --------------------------------------------------------------------------------
namespace {
constexpr int kProfSlots = 4, kProfMax = 8192;   // forward, grad_offset, grad_weight, grad_input
struct ProfState {
  bool on = false;
  int n[kProfSlots] = {0, 0, 0, 0};
  cudaEvent_t* ev[kProfSlots] = {nullptr, nullptr, nullptr, nullptr};  // 2 events per launch
} g_prof;
}  // namespace

ProfScope::ProfScope(int slot_, cudaStream_t st_) : slot(slot_), st(st_), id(-1) {
  if (!g_prof.on || slot < 0 || slot >= kProfSlots || g_prof.n[slot] >= kProfMax) return;
  if (!g_prof.ev[slot]) g_prof.ev[slot] = (cudaEvent_t*)calloc(2 * kProfMax, sizeof(cudaEvent_t));
  id = g_prof.n[slot]++;
  cudaEvent_t* e = g_prof.ev[slot] + 2 * id;
  if (!e[0]) {
    cudaEventCreate(&e[0]);
    cudaEventCreate(&e[1]);
  }
  cudaEventRecord(e[0], st);
}
ProfScope::~ProfScope() {
  if (id >= 0) cudaEventRecord(g_prof.ev[slot][2 * id + 1], st);
}

}  // namespace sdb

using namespace sdb;

extern "C" {

const char* sdb_last_error(void) { return g_err; }
int sdb_abi_version(void) { return SDB_ABI_VERSION; }

long long sdb_launch_count(void) { return g_launches; }
int sdb_profile_enable(int on) {
  g_prof.on = on != 0;
  return SDB_OK;
}
int sdb_profile_reset(void) {
  for (int s = 0; s < kProfSlots; ++s) g_prof.n[s] = 0;
  return SDB_OK;
}
int sdb_profile_read(int slot, float* total_ms, int* launches) {
  SDB_REQUIRE(slot >= 0 && slot < kProfSlots && total_ms && launches, SDB_ERR_INVALID, "bad profile slot");
  float sum = 0.f;
  for (int i = 0; i < g_prof.n[slot]; ++i) {
    cudaEvent_t* e = g_prof.ev[slot] + 2 * i;
    SDB_CHECK_CUDA(cudaEventSynchronize(e[1]));
    float ms = 0.f;
    SDB_CHECK_CUDA(cudaEventElapsedTime(&ms, e[0], e[1]));
    sum += ms;
  }
  *total_ms = sum;
  *launches = g_prof.n[slot];
  return SDB_OK;
}

int sdb_dcn_output_size(const sdb_dcn_geom* g, int32_t* Ho, int32_t* Wo) {
  int rc = check_geom(g);
  if (rc) return rc;
  const Geo d = make_geo(*g);
  if (Ho) *Ho = d.Ho;
  if (Wo) *Wo = d.Wo;
  return SDB_OK;
}

int sdb_dcn_supported(const sdb_dcn_geom* g, int io_dtype, int math) {
  if (check_geom(g) || check_io(io_dtype, math)) return 0;
  if (math == SDB_MATH_FP32) return 1;
  const char* why = "";
  if (!tc_supported(make_geo(*g), &why)) {
    set_error("SDB_MATH_BF16 unsupported for this geometry: %s", why);
    return 0;
  }
  return 1;
}

size_t sdb_dcn_workspace_bytes(int op, const sdb_dcn_geom* g, int io_dtype, int math) {
  if (check_geom(g) || check_io(io_dtype, math) || math == SDB_MATH_FP32) return 0;
  return tc_workspace_bytes(op, make_geo(*g), io_dtype);
}

size_t sdb_dcn_packed_input_bytes(const sdb_dcn_geom* g, int math) {
  if (check_geom(g) || math != SDB_MATH_BF16) return 0;
  return tc_packed_input_bytes(make_geo(*g));
}

#define SDB_PROLOGUE()                         \
  int rc = check_geom(g);                      \
  if (rc) return rc;                           \
  rc = check_io(io_dtype, math);               \
  if (rc) return rc;                           \
  rc = require_device();                       \
  if (rc) return rc;                           \
  const Geo d = make_geo(*g);                  \
  cudaStream_t st = (cudaStream_t)stream;      \
  if (d.N == 0) return SDB_OK;

int sdb_dcn_forward(const void* x, const float* offset, const float* mask, const void* weight,
                    const void* bias, void* out, const sdb_dcn_geom* g, int io_dtype, int math,
                    void* workspace, size_t workspace_bytes, void* x_packed_out, void* stream) {
  SDB_PROLOGUE();
  SDB_REQUIRE(x && offset && weight && out, SDB_ERR_INVALID, "x, offset, weight and out must be non-NULL");
  if (math == SDB_MATH_FP32)
    return simt_forward((const float*)x, offset, mask, (const float*)weight, (const float*)bias,
                        (float*)out, d, st);
  const char* why = "";
  SDB_REQUIRE(tc_supported(d, &why), SDB_ERR_UNSUPPORTED, "SDB_MATH_BF16 unsupported: %s", why);
  return tc_forward(x, offset, mask, weight, bias, out, d, io_dtype, workspace, workspace_bytes,
                    x_packed_out, st);
}

int sdb_dcn_backward_data(const void* x, const float* offset, const float* mask, const void* weight,
                          const void* grad_out, void* grad_x, float* grad_offset, float* grad_mask,
                          const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                          size_t workspace_bytes, const void* x_packed, void* stream) {
  SDB_PROLOGUE();
  SDB_REQUIRE(x && offset && weight && grad_out, SDB_ERR_INVALID,
              "x, offset, weight and grad_out must be non-NULL");
  if (math == SDB_MATH_FP32)
    return simt_backward_data((const float*)x, offset, mask, (const float*)weight,
                              (const float*)grad_out, (float*)grad_x, grad_offset, grad_mask, d, st);
  const char* why = "";
  SDB_REQUIRE(tc_supported(d, &why), SDB_ERR_UNSUPPORTED, "SDB_MATH_BF16 unsupported: %s", why);
  return tc_backward_data(x, offset, mask, weight, grad_out, grad_x, grad_offset, grad_mask, d,
                          io_dtype, workspace, workspace_bytes, x_packed, st);
}

int sdb_dcn_backward_weight(const void* x, const float* offset, const float* mask,
                            const void* grad_out, float* grad_weight, float* grad_bias, float scale,
                            const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                            size_t workspace_bytes, const void* x_packed, void* stream) {
  SDB_PROLOGUE();
  SDB_REQUIRE(x && offset && grad_out, SDB_ERR_INVALID, "x, offset and grad_out must be non-NULL");
  if (math == SDB_MATH_FP32)
    return simt_backward_weight((const float*)x, offset, mask, (const float*)grad_out, grad_weight,
                                grad_bias, scale, d, st);
  const char* why = "";
  SDB_REQUIRE(tc_supported(d, &why), SDB_ERR_UNSUPPORTED, "SDB_MATH_BF16 unsupported: %s", why);
  return tc_backward_weight(x, offset, mask, grad_out, grad_weight, grad_bias, scale, d, io_dtype,
                            workspace, workspace_bytes, x_packed, st);
}

}  // extern "C"
