// C-ABI entry points that expose the raw tcgen05 GEMM (used by the unit tests and the micro-benchmarks).
#include "gemm_launch.cuh"
#include "../../include/srggnn.h"

using namespace srg;

namespace srg {
int query_device(DeviceInfo* d) {
  int dev = 0;
  SRG_CUDA(cudaGetDevice(&dev));
  if (d->device == dev && d->num_sms > 0) return SRG_OK;
  cudaDeviceProp prop;
  SRG_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return set_error(SRG_ERR_UNSUPPORTED, "srggnn needs an sm_100a GPU (found sm_%d%d)", prop.major, prop.minor);
  d->device = dev;
  d->num_sms = prop.multiProcessorCount;
  return SRG_OK;
}
}  // namespace srg

extern "C" {

const char* srg_last_error(void) { return g_last_error; }

long long srg_launch_count(void) { return g_launches.load(); }

int srg_profile_begin(void) {
  g_prof.n = 0;
  g_prof.enabled = true;
  return SRG_OK;
}

// Stops recording, waits for the device, and accumulates per kernel kind: milliseconds, algorithmic FLOPs, launches.
int srg_profile_end(int max_kinds, double* ms, double* flops, long long* launches) {
  g_prof.enabled = false;
  SRG_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < max_kinds; ++k) {
    ms[k] = 0.0;
    flops[k] = 0.0;
    launches[k] = 0;
  }
  for (int i = 0; i < g_prof.n; ++i) {
    float t = 0.f;
    SRG_CUDA(cudaEventElapsedTime(&t, g_prof.ev0[i], g_prof.ev1[i]));
    const int k = g_prof.kind[i];
    if (k >= 0 && k < max_kinds) {
      ms[k] += t;
      // executed FLOPs: with device-resident sizes, the rows / K the kernel actually processed
      double f = g_prof.flops[i];
      const int m_rt = g_prof.rt ? g_prof.rt[2 * i] : -1, k_rt = g_prof.rt ? g_prof.rt[2 * i + 1] : -1;
      if (m_rt >= 0) f = g_prof.mn[i] * m_rt * g_prof.kb_host[i] * kBlockK;
      if (k_rt >= 0) f = g_prof.mn[i] * g_prof.m_host[i] * static_cast<double>(g_prof.nseg[i]) * k_rt;
      flops[k] += f;
      launches[k] += 1;
    }
  }
  g_prof.n = 0;
  return SRG_OK;
}

const char* srg_profile_kind_name(int kind) {
  static const char* epi[] = {"store_bf16", "store_f32", "gru_zr", "gru_h", "logits", "bwd_drh", "bwd_dh", "?"};
  static thread_local char buf[64];
  const int e = kind / 4;
  snprintf(buf, sizeof(buf), "gemm_%s_%s%s", epi[e & 7], (kind & 2) ? "aT" : "a", (kind & 1) ? "bT" : "b");
  return buf;
}

int srg_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* C, int64_t ldc,
                  int c_dtype, int M, int N, int K, const float* bias, float alpha, int cg, int k_splits, int reduce,
                  void* stream) {
  static thread_local DeviceInfo dev;
  SRG_TRY(query_device(&dev));
  GemmProblem p;
  p.cg = cg;
  p.a_mn = a_mn != 0;
  p.b_mn = b_mn != 0;
  p.M = M;
  p.N = N;
  p.nseg = 1;
  p.seg[0].a = p.a_mn ? mat(A, K, M, lda, DT_BF16) : mat(A, M, K, lda, DT_BF16);
  p.seg[0].k_off = 0;
  p.seg[0].k_len = K;
  p.b = p.b_mn ? mat(B, K, N, ldb, DT_BF16) : mat(B, N, K, ldb, DT_BF16);
  p.alpha = alpha;
  p.bias = bias;
  p.bias_scale = 1.f;
  p.k_splits = k_splits;
  if (c_dtype == SRG_DT_F32) {
    p.epi = EPI_STORE_F32;
    p.flags = reduce ? FLAG_REDUCE : 0;
    p.io[0] = mat(C, M, N, ldc, DT_F32);
  } else if (c_dtype == SRG_DT_BF16) {
    p.epi = EPI_STORE_BF16;
    p.io[0] = mat(C, M, N, ldc, DT_BF16);
  } else {
    return set_error(SRG_ERR_ARG, "srg_gemm_bf16: bad c_dtype %d", c_dtype);
  }
  return run_gemm(p, dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
