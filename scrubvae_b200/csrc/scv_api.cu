// C-ABI glue: error text, device info, precision dispatch of scv_gemm / scv_wgrad.
#include <stdarg.h>
#include "scv_common.cuh"

namespace scv {

static thread_local char g_err[512] = "";
int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

int gemm_ffma(const scv_gemm_t* p, cudaStream_t st);
int wgrad_ffma(const scv_wgrad_t* p, cudaStream_t st);
int gemm_ffma_group(const scv_gemm_t* p, int n, cudaStream_t st);
int wgrad_ffma_group(const scv_wgrad_t* p, int n, cudaStream_t st);
int gemm_tc(const scv_gemm_t* p, cudaStream_t st);    // returns 1 if the shape is not taken
int wgrad_tc(const scv_wgrad_t* p, cudaStream_t st);  // returns 1 if the shape is not taken

}  // namespace scv

extern "C" {

int scv_version(void) { return 100; }
const char* scv_last_error(void) { return scv::g_err; }
int64_t scv_launch_count(void) { return scv::g_launches; }

int scv_gemm(const scv_gemm_t* p, void* stream) {
  SCV_REQUIRE(p && p->A && p->W && p->Y, "scv_gemm: null pointer");
  SCV_REQUIRE(p->B > 0 && p->Lo > 0 && p->K > 0 && p->N > 0, "scv_gemm: empty problem");
  SCV_REQUIRE(p->n_last >= 0 && p->n_last <= p->N, "scv_gemm: n_last out of range");
  SCV_REQUIRE(!p->bias || (p->bias_mod > 0 && p->bias_n <= p->N), "scv_gemm: bad bias_mod/bias_n");
  if (p->precision != SCV_PREC_FP32) {
    int r = scv::gemm_tc(p, (cudaStream_t)stream);
    if (r != 1) return r;
    SCV_REQUIRE(p->precision != SCV_PREC_BF16, "scv_gemm: bf16 operands need the tensor-core path, which declined this shape "
                "(N >= 16, K %% 8 == 0, strides %% 8 == 0, 16-byte aligned pointers, Y / R / bias float4-aligned)");
  }
  int rc = scv::gemm_ffma(p, (cudaStream_t)stream);
  if (rc == 0 && p->bnr_sums) {
    // the FFMA path has no fused reduction: the stand-alone pass over X and the Y just written, same sums
    const int64_t C = p->bnr_c;
    SCV_REQUIRE(C > 0 && p->N % C == 0 && (p->Lo == 1 || (p->y_ls == p->N && p->bnr_ls == p->N)),
                "scv_gemm: fused BatchNorm reduction needs rows that are contiguous groups of bnr_c channels");
    scv_bnact_bwd_t q;
    memset(&q, 0, sizeof(q));
    q.X = p->bnr_x; q.x_bs = p->bnr_bs; q.x_ls = C;
    q.B = p->B; q.L = p->Lo * (p->N / C) - (p->N - p->n_last) / C; q.C = C;
    q.fold = 1; q.count = 1.0; q.eps = 0.0;
    q.slope = p->bnr_slope;
    q.dO = p->Y; q.o_bs = p->y_bs; q.o_ls = C;
    q.sums = p->bnr_sums;
    q.mode = (p->bnr_chan ? 1 : 0) | (p->bnr_slope ? 2 : 0) | 4;
    q.chan = p->bnr_chan;
    rc = scv_bnact_bwd_reduce(&q, stream);
  }
  return rc;
}

int scv_gemm_group(const scv_gemm_t* p, int64_t n, void* stream) {
  SCV_REQUIRE(p && n >= 1, "scv_gemm_group: no problems");
  for (int64_t i = 0; i < n; ++i) {
    SCV_REQUIRE(p[i].A && p[i].W && p[i].Y, "scv_gemm_group: null pointer (problem %lld)", (long long)i);
    SCV_REQUIRE(p[i].B > 0 && p[i].Lo > 0 && p[i].K > 0 && p[i].N > 0, "scv_gemm_group: empty problem");
    SCV_REQUIRE(p[i].n_last >= 0 && p[i].n_last <= p[i].N, "scv_gemm_group: n_last out of range");
    SCV_REQUIRE(!p[i].bias || (p[i].bias_mod > 0 && p[i].bias_n <= p[i].N), "scv_gemm_group: bad bias_mod/bias_n");
  }
  return scv::gemm_ffma_group(p, (int)n, (cudaStream_t)stream);
}

int scv_wgrad_group(const scv_wgrad_t* p, int64_t n, void* stream) {
  SCV_REQUIRE(p && n >= 1, "scv_wgrad_group: no problems");
  for (int64_t i = 0; i < n; ++i) {
    SCV_REQUIRE(p[i].A && p[i].dY && p[i].dW, "scv_wgrad_group: null pointer (problem %lld)", (long long)i);
    SCV_REQUIRE(p[i].B > 0 && p[i].Lo > 0 && p[i].K > 0 && p[i].N > 0, "scv_wgrad_group: empty problem");
    SCV_REQUIRE(!p[i].dbias || (p[i].bias_mod > 0 && p[i].bias_n <= p[i].N), "scv_wgrad_group: bad bias_mod/bias_n");
  }
  return scv::wgrad_ffma_group(p, (int)n, (cudaStream_t)stream);
}

int scv_wgrad(const scv_wgrad_t* p, void* stream) {
  SCV_REQUIRE(p && p->A && p->dY && p->dW, "scv_wgrad: null pointer");
  SCV_REQUIRE(p->B > 0 && p->Lo > 0 && p->K > 0 && p->N > 0, "scv_wgrad: empty problem");
  SCV_REQUIRE(!p->dbias || (p->bias_mod > 0 && p->bias_n <= p->N), "scv_wgrad: bad bias_mod/bias_n");
  if (p->precision != SCV_PREC_FP32) {
    int r = scv::wgrad_tc(p, (cudaStream_t)stream);
    if (r != 1) return r;
    SCV_REQUIRE(p->precision != SCV_PREC_BF16, "scv_wgrad: bf16 operands need the tensor-core path, which declined this shape");
  }
  return scv::wgrad_ffma(p, (cudaStream_t)stream);
}

}  // extern "C"
