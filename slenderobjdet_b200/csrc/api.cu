// api.cu -- extern "C" entry points of libslender_b200.so (see include/slender_b200.h).
#include <stdarg.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "dcn_tc_shared.cuh"

namespace sdb {

static thread_local char g_err[512] = "";
long long g_launches = 0;
namespace tcshared { int g_sm_reserve = 0; }
extern int g_fwd_pair, g_bwd_pair;

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
  SDB_REQUIRE(math == SDB_MATH_FP32 || math == SDB_MATH_BF16 || math == SDB_MATH_TF32 || math == SDB_MATH_TF32X3,
              SDB_ERR_INVALID, "unknown math mode %d", math);
  SDB_REQUIRE(!(math != SDB_MATH_BF16 && io_dtype != SDB_F32), SDB_ERR_UNSUPPORTED,
              "SDB_MATH_FP32 / TF32 / TF32X3 need float32 tensors (bf16 tensors use SDB_MATH_BF16)");
  return SDB_OK;
}
static inline bool is_tf32(int math) { return math == SDB_MATH_TF32 || math == SDB_MATH_TF32X3; }
static inline int tf32_passes(int math) { return math == SDB_MATH_TF32X3 ? 3 : 1; }

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
int sdb_set_forward_pair(int on) {
  g_fwd_pair = on != 0;
  return SDB_OK;
}
int sdb_set_backward_pair(int on) {
  g_bwd_pair = on != 0;
  return SDB_OK;
}
int sdb_set_sm_reserve(int n) {
  SDB_REQUIRE(n >= 0 && n < 128, SDB_ERR_INVALID, "SM reserve must be in [0, 128), got %d", n);
  tcshared::g_sm_reserve = n;
  return SDB_OK;
}
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
  if (is_tf32(math) ? !tf32_supported(make_geo(*g), &why) : !tc_supported(make_geo(*g), &why)) {
    set_error("%s unsupported for this geometry: %s", is_tf32(math) ? "SDB_MATH_TF32" : "SDB_MATH_BF16", why);
    return 0;
  }
  return 1;
}

}  // extern "C"

// ---- multi-problem plumbing ---------------------------------------------------------------------------------------
namespace {
struct MultiCall {
  TcProblem pb[tcshared::MAX_PROBS];
  bool have_prep[tcshared::MAX_WEIGHTS];
  float* gw[tcshared::MAX_WEIGHTS];
  float* gb[tcshared::MAX_WEIGHTS];
  TcPlan plan;
};

// validate the tables and translate them (workspace pointers are resolved later, by resolve())
int build_call(const sdb_dcn_problem* probs, int n, const sdb_dcn_weights* w, int nw, const Geo& g, bool backward,
               MultiCall& mc) {
  SDB_REQUIRE(probs && n >= 1 && n <= tcshared::MAX_PROBS, SDB_ERR_INVALID, "need 1..%d problems, got %d", tcshared::MAX_PROBS, n);
  SDB_REQUIRE(w && nw >= 1 && nw <= tcshared::MAX_WEIGHTS, SDB_ERR_INVALID, "need 1..%d weight tensors, got %d", tcshared::MAX_WEIGHTS, nw);
  for (int k = 0; k < nw; ++k) {
    mc.have_prep[k] = w[k].prepared != nullptr;
    mc.gw[k] = backward ? w[k].grad_weight : nullptr;
    mc.gb[k] = backward ? w[k].grad_bias : nullptr;
  }
  for (int i = 0; i < n; ++i) {
    const sdb_dcn_problem& q = probs[i];
    SDB_REQUIRE(q.N >= 0 && q.H > 0 && q.W > 0, SDB_ERR_INVALID, "problem %d: bad size N=%d H=%d W=%d", i, q.N, q.H, q.W);
    SDB_REQUIRE(q.weight_id >= 0 && q.weight_id < nw, SDB_ERR_INVALID, "problem %d: weight_id %d out of range", i, q.weight_id);
    Geo gi = g;
    gi.N = q.N; gi.H = q.H; gi.W = q.W;
    gi.Ho = (q.H + 2 * g.ph - (g.dh * (g.KH - 1) + 1)) / g.sh + 1;
    gi.Wo = (q.W + 2 * g.pw - (g.dw * (g.KW - 1) + 1)) / g.sw + 1;
    SDB_REQUIRE(gi.Ho > 0 && gi.Wo > 0, SDB_ERR_INVALID, "problem %d: output size (%d x %d) is too small", i, gi.Ho, gi.Wo);
    const char* why = "";
    SDB_REQUIRE(tc_supported(gi, &why), SDB_ERR_UNSUPPORTED, "problem %d: tensor-core path unsupported: %s", i, why);
    // plain-convolution mode: offset == NULL for EVERY problem of the call (the towers' Conv2d, reppointsv2.py:644-675)
    SDB_REQUIRE((q.offset == nullptr) == (probs[0].offset == nullptr), SDB_ERR_INVALID,
                "problem %d: either every problem has an offset tensor or none has (plain convolution)", i);
    if (q.offset == nullptr) {
      SDB_REQUIRE(q.mask == nullptr && q.grad_offset == nullptr && q.grad_mask == nullptr, SDB_ERR_INVALID,
                  "problem %d: plain convolution takes no mask and yields no offset / mask gradient", i);
      SDB_REQUIRE(tc_conv_supported(gi, backward && q.grad_x != nullptr, &why), SDB_ERR_UNSUPPORTED,
                  "problem %d: plain convolution unsupported: %s", i, why);
    }
    TcProblem& t = mc.pb[i];
    t = TcProblem{};
    t.d = Dims{q.N, q.H, q.W, gi.Ho, gi.Wo};
    t.weight_id = q.weight_id; t.group = q.offset_group;
    t.x = q.x; t.off = q.offset; t.mask = q.mask; t.out = q.out; t.xp = q.x_packed; t.col = (uint8_t*)q.columns;
    t.gy = q.grad_out; t.gx = backward ? q.grad_x : nullptr; t.goff = backward ? q.grad_offset : nullptr;
    t.gmask = backward ? q.grad_mask : nullptr;
    if (q.offset_group >= 0)
      for (int j = 0; j < i; ++j)
        if (probs[j].offset_group == q.offset_group)
          SDB_REQUIRE(probs[j].N == q.N && probs[j].H == q.H && probs[j].W == q.W && probs[j].offset == q.offset &&
                          probs[j].mask == q.mask,
                      SDB_ERR_INVALID, "problems %d and %d share offset_group %d but not their offset / mask / size", j, i,
                      q.offset_group);
  }
  mc.plan = tc_plan(mc.pb, n, nw, mc.have_prep, g, backward);
  return SDB_OK;
}

// point every problem at its workspace slices and weight images; prepare the images the call still needs
int resolve(MultiCall& mc, int n, const sdb_dcn_weights* w, int nw, const Geo& g, int io_dtype, bool backward, uint8_t* base,
            bool* packed_from_caller, cudaStream_t st) {
  TcWeightImages img[tcshared::MAX_WEIGHTS];
  cudaStream_t pst = st;   // images the call has to build itself: on the side stream, beside the layout packs
  bool forked = false;
  for (int k = 0; k < nw; ++k) {
    const void* prep = w[k].prepared;
    if (!prep) {
      int which = 0;   // only the images this call reads
      if (!backward) which = 1;
      else if (!mc.plan.conv) {
        for (int i = 0; i < n; ++i)
          if (mc.pb[i].weight_id == k && (mc.pb[i].goff || mc.pb[i].gmask || mc.pb[i].gx)) which |= 2;
      }
      if (which && !forked) { pst = tc_prep_begin(st); forked = true; }
      int rc = tc_prepare_weights(w[k].weight, w[k].bias, g, io_dtype, base + mc.plan.prep_off[k], which, pst);
      if (rc) return rc;
      prep = base + mc.plan.prep_off[k];
    }
    img[k] = tc_weight_images(g, prep, w[k].bias != nullptr);
  }
  if (forked) tc_prep_end(st, pst);
  *packed_from_caller = true;
  for (int i = 0; i < n; ++i) {
    TcProblem& t = mc.pb[i];
    t.w = img[t.weight_id];
    if (!t.xp) { t.xp = base + mc.plan.xp_off[i]; *packed_from_caller = false; }
    if (backward) {
      t.gy_img = base + mc.plan.gy_off[i];
      t.dcol = base + mc.plan.dcol_off[i];
    }
  }
  return SDB_OK;
}

int multi_fp32(const sdb_dcn_problem* probs, int n, const sdb_dcn_weights* w, const Geo& g, bool backward, float scale,
               int flags, cudaStream_t st) {
  for (int i = 0; i < n; ++i) {
    const sdb_dcn_problem& q = probs[i];
    const sdb_dcn_weights& ww = w[q.weight_id];
    Geo gi = g;
    gi.N = q.N; gi.H = q.H; gi.W = q.W;
    gi.Ho = (q.H + 2 * g.ph - (g.dh * (g.KH - 1) + 1)) / g.sh + 1;
    gi.Wo = (q.W + 2 * g.pw - (g.dw * (g.KW - 1) + 1)) / g.sw + 1;
    SDB_REQUIRE(gi.Ho > 0 && gi.Wo > 0, SDB_ERR_INVALID, "problem %d: output size is too small", i);
    if (gi.N == 0) continue;
    int rc = SDB_OK;
    if (!backward) {
      rc = simt_forward((const float*)q.x, q.offset, q.mask, (const float*)ww.weight, (const float*)ww.bias, (float*)q.out, gi, st);
    } else {
      const bool do_data = !(flags & SDB_BWD_WEIGHT_ONLY), do_w = !(flags & SDB_BWD_DATA_ONLY);
      if (do_data && q.grad_x) SDB_CHECK_CUDA(cudaMemsetAsync(q.grad_x, 0, (size_t)gi.N * gi.C * gi.H * gi.W * 4, st));   // overwritten
      if (do_data && (q.grad_x || q.grad_offset || q.grad_mask))
        rc = simt_backward_data((const float*)q.x, q.offset, q.mask, (const float*)ww.weight, (const float*)q.grad_out,
                                (float*)q.grad_x, q.grad_offset, q.mask ? q.grad_mask : nullptr, gi, st);
      if (!rc && do_w && (ww.grad_weight || ww.grad_bias))
        rc = simt_backward_weight((const float*)q.x, q.offset, q.mask, (const float*)q.grad_out, ww.grad_weight,
                                  ww.grad_bias, scale, gi, st);
    }
    if (rc) return rc;
  }
  return SDB_OK;
}

// kind::tf32 forward (dcn_tf32.cu): validate the table, translate it, run.  `ws` == nullptr: only the size is wanted.
int tf32_call(const sdb_dcn_problem* probs, int n, const sdb_dcn_weights* w, int nw, const Geo& g, int passes, uint8_t* ws,
              size_t ws_bytes, size_t* need, bool size_only, cudaStream_t st) {
  SDB_REQUIRE(probs && n >= 1 && n <= tcshared::MAX_PROBS, SDB_ERR_INVALID, "need 1..%d problems, got %d", tcshared::MAX_PROBS, n);
  SDB_REQUIRE(w && nw >= 1 && nw <= tcshared::MAX_WEIGHTS, SDB_ERR_INVALID, "need 1..%d weight tensors, got %d", tcshared::MAX_WEIGHTS, nw);
  TcProblem pb[tcshared::MAX_PROBS];
  bool have_prep[tcshared::MAX_WEIGHTS];
  const void *wt[tcshared::MAX_WEIGHTS], *bs[tcshared::MAX_WEIGHTS], *prep[tcshared::MAX_WEIGHTS];
  for (int k = 0; k < nw; ++k) {
    have_prep[k] = w[k].prepared != nullptr;
    wt[k] = w[k].weight; bs[k] = w[k].bias; prep[k] = w[k].prepared;
    SDB_REQUIRE(size_only || wt[k] || prep[k], SDB_ERR_INVALID, "weight %d: neither weight nor prepared image given", k);
  }
  for (int i = 0; i < n; ++i) {
    const sdb_dcn_problem& q = probs[i];
    SDB_REQUIRE(q.N >= 0 && q.H > 0 && q.W > 0, SDB_ERR_INVALID, "problem %d: bad size N=%d H=%d W=%d", i, q.N, q.H, q.W);
    SDB_REQUIRE(q.weight_id >= 0 && q.weight_id < nw, SDB_ERR_INVALID, "problem %d: weight_id %d out of range", i, q.weight_id);
    Geo gi = g;
    gi.N = q.N; gi.H = q.H; gi.W = q.W;
    gi.Ho = (q.H + 2 * g.ph - (g.dh * (g.KH - 1) + 1)) / g.sh + 1;
    gi.Wo = (q.W + 2 * g.pw - (g.dw * (g.KW - 1) + 1)) / g.sw + 1;
    SDB_REQUIRE(gi.Ho > 0 && gi.Wo > 0, SDB_ERR_INVALID, "problem %d: output size (%d x %d) is too small", i, gi.Ho, gi.Wo);
    const char* why = "";
    SDB_REQUIRE(tf32_supported(gi, &why), SDB_ERR_UNSUPPORTED, "problem %d: tf32 tensor-core path unsupported: %s", i, why);
    TcProblem& t = pb[i];
    t = TcProblem{};
    t.d = Dims{q.N, q.H, q.W, gi.Ho, gi.Wo};
    t.weight_id = q.weight_id;
    t.x = q.x; t.off = q.offset; t.mask = q.mask; t.out = q.out;
  }
  *need = tf32_forward_workspace_bytes(pb, n, nw, have_prep, g, passes);
  if (size_only) return SDB_OK;
  SDB_REQUIRE(*need == 0 || (ws && ws_bytes >= *need), SDB_ERR_WORKSPACE, "forward workspace too small: %zu < %zu", ws_bytes, *need);
  return tf32_forward_all(pb, n, wt, bs, prep, nw, g, passes, ws, st);
}

// the single-problem entry points are one-row tables
struct Single {
  sdb_dcn_problem p;
  sdb_dcn_weights w;
};
Single single_of(const sdb_dcn_geom* g) {
  Single s{};
  s.p.N = g->N; s.p.H = g->H; s.p.W = g->W; s.p.weight_id = 0; s.p.offset_group = -1;
  return s;
}
// the common geometry of a table call: N / H / W of `g` are ignored (every problem brings its own)
sdb_dcn_geom common_geom(const sdb_dcn_geom* g) {
  sdb_dcn_geom c = *g;
  c.N = 1; c.H = 1024; c.W = 1024;   // placeholder extent that passes every size limit
  return c;
}
}  // namespace

extern "C" {

#define SDB_MULTI_PROLOGUE()                                                                       \
  SDB_REQUIRE(g != nullptr, SDB_ERR_INVALID, "geometry pointer is NULL");                          \
  const sdb_dcn_geom cg = common_geom(g);                                                          \
  int rc = check_geom(&cg);                                                                        \
  if (rc) return rc;                                                                               \
  rc = check_io(io_dtype, math);                                                                   \
  if (rc) return rc;                                                                               \
  rc = require_device();                                                                           \
  if (rc) return rc;                                                                               \
  const Geo d = make_geo(cg);                                                                      \
  cudaStream_t st = (cudaStream_t)stream;

size_t sdb_dcn_prepared_weight_bytes(const sdb_dcn_geom* g, int io_dtype, int math) {
  if (!g) return 0;
  const sdb_dcn_geom cg = common_geom(g);
  if (check_geom(&cg) || check_io(io_dtype, math) || math == SDB_MATH_FP32) return 0;
  if (is_tf32(math)) return tf32_prepared_weight_bytes(make_geo(cg), tf32_passes(math));
  return tc_prepared_weight_bytes(make_geo(cg));
}

int sdb_dcn_prepare_weights(const void* weight, const void* bias, const sdb_dcn_geom* g, int io_dtype, int math,
                            void* prepared, void* stream) {
  SDB_MULTI_PROLOGUE();
  if (math == SDB_MATH_FP32) return SDB_OK;
  SDB_REQUIRE(weight && prepared, SDB_ERR_INVALID, "weight and prepared must be non-NULL");
  const char* why = "";
  if (is_tf32(math)) {
    SDB_REQUIRE(tf32_supported(d, &why), SDB_ERR_UNSUPPORTED, "SDB_MATH_TF32 unsupported: %s", why);
    return tf32_prepare_weights((const float*)weight, (const float*)bias, d, tf32_passes(math), prepared, st);
  }
  SDB_REQUIRE(tc_supported(d, &why), SDB_ERR_UNSUPPORTED, "SDB_MATH_BF16 unsupported: %s", why);
  return tc_prepare_weights(weight, bias, d, io_dtype, prepared, 3, st);
}

size_t sdb_dcn_multi_workspace_bytes(const sdb_dcn_problem* problems, int32_t n, const sdb_dcn_weights* weights,
                                     int32_t nw, const sdb_dcn_geom* g, int io_dtype, int math, int backward) {
  if (!g) return 0;
  const sdb_dcn_geom cg = common_geom(g);
  if (check_geom(&cg) || check_io(io_dtype, math) || math == SDB_MATH_FP32) return 0;
  if (is_tf32(math)) {   // the backward of the tf32 modes is the exact fp32 path: no workspace
    size_t need = 0;
    if (backward || tf32_call(problems, n, weights, nw, make_geo(cg), tf32_passes(math), nullptr, 0, &need, true, nullptr)) return 0;
    return need;
  }
  MultiCall mc;
  if (build_call(problems, n, weights, nw, make_geo(cg), backward != 0, mc)) return 0;
  return mc.plan.total;
}

int sdb_dcn_forward_multi(const sdb_dcn_problem* problems, int32_t n, const sdb_dcn_weights* weights, int32_t nw,
                          const sdb_dcn_geom* g, int io_dtype, int math, void* workspace, size_t workspace_bytes,
                          void* stream) {
  SDB_MULTI_PROLOGUE();
  SDB_REQUIRE(problems && weights, SDB_ERR_INVALID, "NULL table");
  for (int i = 0; i < n; ++i)
    SDB_REQUIRE(problems[i].N == 0 || (problems[i].x && problems[i].out && (problems[i].offset || math == SDB_MATH_BF16)),
                SDB_ERR_INVALID, "problem %d: x, offset and out must be non-NULL (offset == NULL, plain convolution, needs SDB_MATH_BF16)", i);
  if (math == SDB_MATH_FP32) return multi_fp32(problems, n, weights, d, false, 1.f, 0, st);
  if (is_tf32(math)) {
    size_t need = 0;
    return tf32_call(problems, n, weights, nw, d, tf32_passes(math), (uint8_t*)workspace, workspace_bytes, &need, false, st);
  }
  MultiCall mc;
  rc = build_call(problems, n, weights, nw, d, false, mc);
  if (rc) return rc;
  SDB_REQUIRE(mc.plan.total == 0 || (workspace && workspace_bytes >= mc.plan.total), SDB_ERR_WORKSPACE,
              "forward workspace too small: %zu < %zu", workspace_bytes, mc.plan.total);
  bool from_caller;
  rc = resolve(mc, n, weights, nw, d, io_dtype, false, (uint8_t*)workspace, &from_caller, st);
  if (rc) return rc;
  rc = tc_forward_all(mc.pb, n, d, io_dtype, st);
  const int rcw = tc_prep_wait(st);   // no-op unless no kernel consumed the in-call weight preparation
  return rc ? rc : rcw;
}

static int backward_multi_impl(const sdb_dcn_problem* problems, int32_t n, const sdb_dcn_weights* weights, int32_t nw,
                               const Geo& d, int io_dtype, float scale, int accumulate_gx, int flags, void* workspace,
                               size_t workspace_bytes, cudaStream_t st) {
  MultiCall mc;
  int rc = build_call(problems, n, weights, nw, d, true, mc);
  if (rc) return rc;
  SDB_REQUIRE(workspace && workspace_bytes >= mc.plan.total, SDB_ERR_WORKSPACE, "backward workspace too small: %zu < %zu",
              workspace_bytes, mc.plan.total);
  if (flags & SDB_BWD_DATA_ONLY)
    for (int k = 0; k < nw; ++k) mc.gw[k] = mc.gb[k] = nullptr;
  if (flags & SDB_BWD_WEIGHT_ONLY)   // BUILD_INDEX keeps grad_x: it tells which offset groups need a transposed index
    for (int i = 0; i < n; ++i) {
      if (!(flags & SDB_BWD_BUILD_INDEX)) mc.pb[i].gx = nullptr;
      mc.pb[i].goff = nullptr, mc.pb[i].gmask = nullptr;
    }
  bool from_caller;
  if (!(flags & SDB_BWD_GATHER_ONLY)) {   // the gather reads no weight image
    rc = resolve(mc, n, weights, nw, d, io_dtype, true, (uint8_t*)workspace, &from_caller, st);
    if (rc) return rc;
  } else {
    for (int i = 0; i < n; ++i) mc.pb[i].dcol = (uint8_t*)workspace + mc.plan.dcol_off[i];
  }
  // x_packed given for SOME problems only: the pack launch covers the ones that live in the workspace
  bool pack_any = false;
  for (int i = 0; i < n; ++i) {
    const bool own = problems[i].x_packed == nullptr;
    pack_any |= own;
    if (!own) mc.pb[i].x = nullptr;   // pack_nhwc_multi skips NULL sources
  }
  if (mc.plan.conv) {
    SDB_REQUIRE(!accumulate_gx && !(flags & (SDB_BWD_GRAD_PACKED | SDB_BWD_NO_GATHER | SDB_BWD_GATHER_ONLY | SDB_BWD_BUILD_INDEX | SDB_BWD_INDEX_READY)), SDB_ERR_UNSUPPORTED,
                "plain convolution: grad_x is overwritten, and the phased-backward flags are not supported");
    const void* wt[tcshared::MAX_WEIGHTS];
    for (int k = 0; k < nw; ++k) {
      wt[k] = weights[k].weight;
      bool needs = false;
      for (int i = 0; i < n; ++i) needs |= mc.pb[i].weight_id == k && mc.pb[i].gx != nullptr;
      SDB_REQUIRE(!needs || wt[k], SDB_ERR_INVALID, "weight %d: the weight tensor is needed for grad_x", k);
    }
    rc = tc_conv_backward_all(mc.pb, n, wt, mc.gw, mc.gb, nw, mc.plan, d, io_dtype, scale, pack_any, (uint8_t*)workspace, st);
    const int rcw = tc_prep_wait(st);
    return rc ? rc : rcw;
  }
  const int phase = (flags & SDB_BWD_NO_GATHER) ? 1 : (flags & SDB_BWD_GATHER_ONLY) ? 2 : (flags & SDB_BWD_BUILD_INDEX) ? 3
                    : (flags & SDB_BWD_INDEX_READY) ? 4 : 0;
  rc = tc_backward_all(mc.pb, n, mc.gw, mc.gb, nw, mc.plan, d, io_dtype, scale, pack_any, accumulate_gx,
                       (flags & SDB_BWD_GRAD_PACKED) != 0, (uint8_t*)workspace, st, phase);
  const int rcw = tc_prep_wait(st);   // no-op unless no kernel consumed the in-call weight preparation
  return rc ? rc : rcw;
}

int sdb_dcn_backward_multi(const sdb_dcn_problem* problems, int32_t n, const sdb_dcn_weights* weights, int32_t nw,
                           const sdb_dcn_geom* g, int io_dtype, int math, float scale, int flags, void* workspace,
                           size_t workspace_bytes, void* stream) {
  // argument checks that need no device come first
  SDB_REQUIRE((flags & ~127) == 0 && (flags & 3) != 3 && (flags & 24) != 24 && !((flags & 24) && (flags & SDB_BWD_WEIGHT_ONLY)) &&
                  !((flags & SDB_BWD_BUILD_INDEX) && (flags & ~(SDB_BWD_BUILD_INDEX | SDB_BWD_WEIGHT_ONLY))) &&
                  !((flags & SDB_BWD_BUILD_INDEX) && !(flags & SDB_BWD_WEIGHT_ONLY)) &&
                  !((flags & SDB_BWD_INDEX_READY) && (flags & (24 | SDB_BWD_WEIGHT_ONLY))),
              SDB_ERR_INVALID, "bad backward flags %d", flags);
  SDB_MULTI_PROLOGUE();
  SDB_REQUIRE(!(flags & 120) || math == SDB_MATH_BF16, SDB_ERR_UNSUPPORTED, "the phased-backward flags need SDB_MATH_BF16");
  SDB_REQUIRE(problems && weights, SDB_ERR_INVALID, "NULL table");
  for (int i = 0; i < n; ++i)
    SDB_REQUIRE(problems[i].N == 0 || (problems[i].x && problems[i].grad_out && (problems[i].offset || math == SDB_MATH_BF16)),
                SDB_ERR_INVALID, "problem %d: x, offset and grad_out must be non-NULL (offset == NULL, plain convolution, needs SDB_MATH_BF16)", i);
  if (math == SDB_MATH_FP32 || is_tf32(math)) return multi_fp32(problems, n, weights, d, true, scale, flags, st);
  return backward_multi_impl(problems, n, weights, nw, d, io_dtype, scale, 0, flags, workspace, workspace_bytes, st);
}

// ---- single-problem entry points (the reference's call granularity) --------------------------------------------
size_t sdb_dcn_workspace_bytes(int op, const sdb_dcn_geom* g, int io_dtype, int math) {
  if (check_geom(g) || check_io(io_dtype, math) || math == SDB_MATH_FP32) return 0;
  Single s = single_of(g);
  s.p.offset = reinterpret_cast<const float*>(sizeof(float));   // size query of a DEFORMABLE convolution (NULL = plain convolution); never read
  return sdb_dcn_multi_workspace_bytes(&s.p, 1, &s.w, 1, g, io_dtype, math, op != SDB_OP_FORWARD);
}

size_t sdb_dcn_packed_input_bytes(const sdb_dcn_geom* g, int math) {
  if (check_geom(g) || math != SDB_MATH_BF16) return 0;
  return tc_packed_input_bytes(make_geo(*g));
}

size_t sdb_dcn_columns_bytes(const sdb_dcn_geom* g, int math) {
  if (check_geom(g) || math != SDB_MATH_BF16) return 0;
  return tc_columns_bytes(make_geo(*g));
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
  Single s = single_of(g);
  s.p.x = x; s.p.offset = offset; s.p.mask = mask; s.p.out = out; s.p.x_packed = x_packed_out;
  s.w.weight = weight; s.w.bias = bias;
  return sdb_dcn_forward_multi(&s.p, 1, &s.w, 1, g, io_dtype, math, workspace, workspace_bytes, stream);
}

int sdb_dcn_backward_data(const void* x, const float* offset, const float* mask, const void* weight,
                          const void* grad_out, void* grad_x, float* grad_offset, float* grad_mask,
                          const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                          size_t workspace_bytes, const void* x_packed, void* stream) {
  SDB_PROLOGUE();
  SDB_REQUIRE(x && offset && weight && grad_out, SDB_ERR_INVALID,
              "x, offset, weight and grad_out must be non-NULL");
  if (math == SDB_MATH_FP32 || is_tf32(math))
    return simt_backward_data((const float*)x, offset, mask, (const float*)weight,
                              (const float*)grad_out, (float*)grad_x, grad_offset, grad_mask, d, st);
  Single s = single_of(g);
  s.p.x = x; s.p.offset = offset; s.p.mask = mask; s.p.x_packed = (void*)x_packed; s.p.grad_out = grad_out;
  s.p.grad_x = grad_x; s.p.grad_offset = grad_offset; s.p.grad_mask = grad_mask;
  s.w.weight = weight;
  // this entry point ACCUMULATES into grad_x (the reference's contract, deform_conv.py:89)
  return backward_multi_impl(&s.p, 1, &s.w, 1, d, io_dtype, 1.f, 1, 0, workspace, workspace_bytes, st);
}

int sdb_dcn_backward_weight(const void* x, const float* offset, const float* mask,
                            const void* grad_out, float* grad_weight, float* grad_bias, float scale,
                            const sdb_dcn_geom* g, int io_dtype, int math, void* workspace,
                            size_t workspace_bytes, const void* x_packed, void* stream) {
  SDB_PROLOGUE();
  SDB_REQUIRE(x && offset && grad_out, SDB_ERR_INVALID, "x, offset and grad_out must be non-NULL");
  if (math == SDB_MATH_FP32 || is_tf32(math))
    return simt_backward_weight((const float*)x, offset, mask, (const float*)grad_out, grad_weight,
                                grad_bias, scale, d, st);
  Single s = single_of(g);
  s.p.x = x; s.p.offset = offset; s.p.mask = mask; s.p.x_packed = (void*)x_packed; s.p.grad_out = grad_out;
  s.w.grad_weight = grad_weight; s.w.grad_bias = grad_bias;   // no operand image of the weights is read here
  return backward_multi_impl(&s.p, 1, &s.w, 1, d, io_dtype, scale, 0, 0, workspace, workspace_bytes, st);
}

}  // extern "C"
