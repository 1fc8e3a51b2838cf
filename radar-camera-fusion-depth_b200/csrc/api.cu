// C-ABI entry points that dispatch between the convolution engines, plus version / error plumbing.
#include <stdarg.h>
#include <string.h>
#include "conv_common.cuh"

namespace rcfd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local char g_kernel[96] = "";

void note_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_kernel, sizeof(g_kernel), fmt, ap);
  va_end(ap);
}

int conv_simt_launch(const ConvKP& p, int dtype, cudaStream_t st);
int conv_wgrad_simt_launch(const ConvKP& p, float* dw, int dtype, cudaStream_t st);
bool conv_tc_supported(const ConvKP& p, int dtype);
int conv_tc_launch(const ConvKP& p, cudaStream_t st);
bool conv_strip_up_supported(const ConvKP& p, int dtype);
bool conv_strip_up_preferred(const ConvKP& p, int dtype);
int conv_strip_up_launch(const ConvKP& p, cudaStream_t st);
bool conv_strip_supported(const ConvKP& p, int dtype);
bool conv_strip_preferred(const ConvKP& p, int dtype);
int conv_strip_launch(const ConvKP& p, cudaStream_t st);
extern int g_strip_desc_mode;
extern int g_tma_bn_cap;
extern int g_tma_pair;
extern int g_tma_split_k;
extern int g_wgrad_tma_v2;
extern int g_wgrad_tma_cs_max;
extern int g_wgrad_tma_groups;
extern int g_bn_vectors_per_thread;
extern int g_bn_reduce_rows_per_thread;
extern int g_bn_fwd_vectors_per_thread;
extern int g_ew_vectors_per_thread;
extern int g_strip_max_waste;
extern int g_strip_input_stationary;
extern int g_strip_up_max_waste;
bool conv_tma_supported(const ConvKP& p, int dtype);
int conv_tma_launch(const ConvKP& p, cudaStream_t st);
bool wgrad_tc_supported(const ConvKP& p, int dtype);
bool wgrad_tma_supported(const ConvKP& p, int dtype);
int wgrad_tma_launch(const ConvKP& p, float* dw, cudaStream_t st);
int wgrad_tc_launch(const ConvKP& p, float* dw, cudaStream_t st);
bool wgrad_strip_supported(const ConvKP& p, int dtype);
bool wgrad_strip_preferred(const ConvKP& p, int dtype);
int64_t wgrad_strip_workspace(const ConvKP& p, int dtype);
int wgrad_strip_launch(const ConvKP& p, float* dw, void* workspace, int64_t workspace_bytes, cudaStream_t st);

}  // namespace rcfd

using namespace rcfd;

extern "C" {

const char* rcfd_version(void) { return "rcfd-b200 0.1.0"; }
const char* rcfd_arch(void) { return "sm_100a"; }
const char* rcfd_last_error(void) { return g_err; }
const char* rcfd_last_kernel(void) { return g_kernel; }

int rcfd_conv2d_fwd(const rcfd_conv_desc* d, void* stream) {
  ConvKP p;
  int rc = make_conv_kp(d, &p);
  if (rc != RCFD_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int engine = d->engine;
  if (engine == RCFD_ENGINE_AUTO && (conv_strip_preferred(p, d->dtype) || conv_strip_up_preferred(p, d->dtype)))
    engine = RCFD_ENGINE_STRIP;
  if (engine == RCFD_ENGINE_STRIP) {
    if (conv_strip_up_supported(p, d->dtype)) return conv_strip_up_launch(p, st);
    if (!conv_strip_supported(p, d->dtype)) {
      set_error("conv: not a case of the row-streaming engine (bf16, 3x3 / stride 1 / pad 1, one source with 16, 32 or 64 channels)");
      return RCFD_EUNSUPPORTED;
    }
    return conv_strip_launch(p, st);
  }
  if (engine == RCFD_ENGINE_AUTO)
    engine = conv_tma_supported(p, d->dtype) ? RCFD_ENGINE_TMA
                                             : (conv_tc_supported(p, d->dtype) ? RCFD_ENGINE_TCGEN05 : RCFD_ENGINE_SIMT);
  if (engine == RCFD_ENGINE_TMA) {
    if (!conv_tma_supported(p, d->dtype)) {
      set_error("conv: shape/dtype not supported by the TMA engine (bf16, stride 1/2, no up-sampling / zero insertion, channels %% 16 == 0)");
      return RCFD_EUNSUPPORTED;
    }
    return conv_tma_launch(p, st);
  }
  if (engine == RCFD_ENGINE_TCGEN05) {
    if (!conv_tc_supported(p, d->dtype)) {
      set_error("conv: shape/dtype not supported by the tcgen05 engine (bf16, channels %% 8 == 0, cout %% 16 == 0, cout <= 256)");
      return RCFD_EUNSUPPORTED;
    }
    return conv_tc_launch(p, st);
  }
  if (engine != RCFD_ENGINE_SIMT) { set_error("conv: bad engine %d", engine); return RCFD_EINVAL; }
  note_kernel("conv_simt_kernel");
  return conv_simt_launch(p, d->dtype, st);
}

int rcfd_set_option(const char* key, int32_t value) {
  RCFD_CHECK_ARG(key != nullptr, "set_option: null key");
  if (strcmp(key, "strip_desc_mode") == 0) { g_strip_desc_mode = value; return RCFD_OK; }
  if (strcmp(key, "tma_bn_cap") == 0) { g_tma_bn_cap = value; return RCFD_OK; }
  if (strcmp(key, "tma_pair") == 0) { g_tma_pair = value; return RCFD_OK; }
  if (strcmp(key, "tma_split_k") == 0) { g_tma_split_k = value; return RCFD_OK; }
  if (strcmp(key, "wgrad_tma_v2") == 0) { g_wgrad_tma_v2 = value; return RCFD_OK; }
  if (strcmp(key, "wgrad_tma_cs_max") == 0) { g_wgrad_tma_cs_max = value; return RCFD_OK; }
  if (strcmp(key, "wgrad_tma_groups") == 0) { g_wgrad_tma_groups = value; return RCFD_OK; }
  if (strcmp(key, "bn_reduce_rows_per_thread") == 0) { g_bn_reduce_rows_per_thread = value < 1 ? 1 : value; return RCFD_OK; }
  if (strcmp(key, "bn_vectors_per_thread") == 0) { g_bn_vectors_per_thread = value < 1 ? 1 : value; return RCFD_OK; }
  if (strcmp(key, "bn_fwd_vectors_per_thread") == 0) { g_bn_fwd_vectors_per_thread = value < 1 ? 1 : value; return RCFD_OK; }
  if (strcmp(key, "ew_vectors_per_thread") == 0) { g_ew_vectors_per_thread = value < 1 ? 1 : value; return RCFD_OK; }
  if (strcmp(key, "strip_input_stationary") == 0) { g_strip_input_stationary = value; return RCFD_OK; }
  if (strcmp(key, "strip_max_waste") == 0) { g_strip_max_waste = value; return RCFD_OK; }
  if (strcmp(key, "strip_up_max_waste") == 0) { g_strip_up_max_waste = value; return RCFD_OK; }
  set_error("set_option: unknown key %s", key);
  return RCFD_EINVAL;
}

// bytes of device scratch rcfd_conv2d_wgrad needs for this descriptor (0 for most shapes; the row-streaming
// engine behind an up-sampling accumulates 16 sub-pixel matrices before folding them into the 9 taps)
int64_t rcfd_conv2d_wgrad_workspace(const rcfd_conv_desc* d) {
  ConvKP p;
  if (make_conv_kp(d, &p) != RCFD_OK) return 0;
  if (d->engine == RCFD_ENGINE_STRIP || (d->engine == RCFD_ENGINE_AUTO && wgrad_strip_preferred(p, d->dtype)))
    return wgrad_strip_workspace(p, d->dtype);
  return 0;
}

int rcfd_conv2d_wgrad(const rcfd_conv_desc* d, float* dw, void* workspace, int64_t workspace_bytes, void* stream) {
  ConvKP p;
  int rc = make_conv_kp(d, &p);
  if (rc != RCFD_OK) return rc;
  RCFD_CHECK_ARG(dw != nullptr, "wgrad: null dw");
  int engine = d->engine;
  if (engine == RCFD_ENGINE_AUTO && wgrad_strip_preferred(p, d->dtype)) engine = RCFD_ENGINE_STRIP;
  if (engine == RCFD_ENGINE_STRIP) {
    if (!wgrad_strip_supported(p, d->dtype)) {
      set_error("wgrad: not a case of the row-streaming engine (bf16, 3x3 / stride 1 / pad 1, sources with 32 or 64 channels, cout 16 / 32 / 64)");
      return RCFD_EUNSUPPORTED;
    }
    return wgrad_strip_launch(p, dw, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  if (engine == RCFD_ENGINE_AUTO)
    // TMA feeds one box ROW (<= 128 B) per ~5 cycles: with < 64 channels per box the rows are
    // 32-64 B and the cp.async gather engine is faster (measured: profiles/r1_progress_log.md)
    engine = (wgrad_tma_supported(p, d->dtype) && p.c0 % 64 == 0 && p.c1 % 64 == 0)
                 ? RCFD_ENGINE_TMA
                 : (wgrad_tc_supported(p, d->dtype) ? RCFD_ENGINE_TCGEN05 : RCFD_ENGINE_SIMT);
  if (engine == RCFD_ENGINE_TMA) {
    if (!wgrad_tma_supported(p, d->dtype)) {
      set_error("wgrad: shape/dtype not supported by the TMA engine (bf16, stride 1/2, no up-sampling, channels %% 16 == 0)");
      return RCFD_EUNSUPPORTED;
    }
    return wgrad_tma_launch(p, dw, (cudaStream_t)stream);
  }
  if (engine == RCFD_ENGINE_TCGEN05) {
    if (!wgrad_tc_supported(p, d->dtype)) {
      set_error("wgrad: shape/dtype not supported by the tcgen05 engine (bf16, channels %% 8 == 0)");
      return RCFD_EUNSUPPORTED;
    }
    return wgrad_tc_launch(p, dw, (cudaStream_t)stream);
  }
  note_kernel("conv_wgrad_simt_kernel");
  return conv_wgrad_simt_launch(p, dw, d->dtype, (cudaStream_t)stream);
}

}  // extern "C"
