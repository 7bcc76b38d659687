// Host-side plumbing shared by the C-ABI entry points: error reporting, TMA descriptor encoding,
// and the GEMM launcher that maps a problem description onto one gemm_kernel instantiation.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <string>
#include "gemm.cuh"

namespace srg {

// ---------------------------------------------------------------- errors (no exceptions cross the ABI)
inline thread_local char g_last_error[512] = "";

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

enum : int {
  SRG_OK = 0,
  SRG_ERR_ARG = 1,
  SRG_ERR_CUDA = 2,
  SRG_ERR_UNSUPPORTED = 3,
  SRG_ERR_WORKSPACE = 4,
};

#define SRG_CUDA(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return ::srg::set_error(::srg::SRG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                     \
  } while (0)

#define SRG_CHECK(cond, ...)                                              \
  do {                                                                    \
    if (!(cond)) return ::srg::set_error(::srg::SRG_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define SRG_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// Entry points run on the device their handle was created on, whatever the caller's current device is.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev && dev >= 0) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
constexpr int kMaxDevices = 64;

// ---------------------------------------------------------------- launch accounting / live kernel timing
// g_launches counts every kernel this library launches (bench.py reports it as "gpu_launches").
inline std::atomic<long long> g_launches{0};

// Optional CUDA-event bracket around every GEMM launch, grouped by kernel kind, so that bench.py can report the
// duration of the dominant kernel measured inside the timed region, on the launching stream.
constexpr int kProfMaxRecords = 16384;
constexpr int kProfKinds = 32;
struct ProfState {
  bool enabled = false;
  int n = 0;
  cudaEvent_t ev0[kProfMaxRecords];
  cudaEvent_t ev1[kProfMaxRecords];
  int created = 0;
  int kind[kProfMaxRecords];
  double flops[kProfMaxRecords];       // 2*M*N*K with the host-side (upper bound) sizes
  // device-resident sizes (GemmArgs::m_dev / k_dev) of the launch, copied to pinned host memory on the launching stream
  // right after the kernel, so that the executed FLOPs are counted with the row counts the kernel really saw
  int* rt = nullptr;                   // pinned [kProfMaxRecords][2] = {M, K per segment}, -1 = static
  double mn[kProfMaxRecords];          // 2 * N * (K / M for the two cases below)
  int m_host[kProfMaxRecords], nseg[kProfMaxRecords], kb_host[kProfMaxRecords];
};
inline ProfState g_prof;
#ifdef SRG_EPI_TIMING
inline unsigned long long* g_epi_t_dev = nullptr;   // one instance per process (inline variable), allocated on first use
#endif
inline int prof_kind(int epi, bool a_mn, bool b_mn) { return epi * 4 + (a_mn ? 2 : 0) + (b_mn ? 1 : 0); }

// ---------------------------------------------------------------- TMA descriptors
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

enum DType : int { DT_NONE = 0, DT_F32 = 1, DT_BF16 = 2 };

// Row-major 2-D tensor [rows, cols] with leading dimension `ld` (elements); box = [box_rows, box_cols];
// SWIZZLE_128B when a box row spans 128 bytes, SWIZZLE_64B when it spans 64 bytes.
inline int make_tmap(CUtensorMap* out, const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows, int box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(SRG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = (dtype == DT_F32) ? 4 : 2;
  SRG_CHECK(box_cols * es == 128 || box_cols * es == 64, "TMA box must span 128 or 64 bytes (got %d)", box_cols * es);
  const CUtensorMapSwizzle swz = (box_cols * es == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  SRG_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  SRG_CHECK(((ld * es) & 15) == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%lld)", (long long)ld);
  SRG_CHECK(rows > 0 && cols > 0, "empty tensor for TMA (%lld x %lld)", (long long)rows, (long long)cols);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SRG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
                     (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
  return SRG_OK;
}

// ---------------------------------------------------------------- GEMM problem description
struct Mat {
  const void* ptr = nullptr;
  int64_t rows = 0, cols = 0, ld = 0;  // row-major storage
  int dtype = DT_NONE;
};
inline Mat mat(const void* p, int64_t rows, int64_t cols, int64_t ld, int dtype) {
  Mat m;
  m.ptr = p; m.rows = rows; m.cols = cols; m.ld = ld; m.dtype = dtype;
  return m;
}

struct GemmSeg {
  Mat a;        // K-major: [M, >=k_off+K_s] ; MN-major: [>=k_off+K_s, M]
  int k_off;    // start along K inside `a`
  int k_len;    // multiple of 64
};

struct GemmProblem {
  int cg = 2;           // CTAs per UMMA (1 or 2)
  bool a_mn = false, b_mn = false;
  int epi = EPI_STORE_F32;
  bool f32 = false;     // fp32-parity arithmetic in the epilogue
  int M = 0, N = 0;     // N multiple of the column tile
  // device-resident sizes (see GemmArgs): with m_dev, M is the upper bound the operands are allocated for; with k_dev
  // every segment's k_len is the upper bound of its device-resident length
  const int* m_dev = nullptr;
  const int* k_dev = nullptr;
  int nseg = 0;
  GemmSeg seg[kMaxSeg];
  Mat b;                // K-major: [N, Ktot]; MN-major: [Ktot, N]
  Mat io[kMaxIoMaps];
  float alpha = 1.f;
  const float* bias = nullptr;
  float bias_scale = 1.f;
  int n_split = 0, n_valid = 0;
  float* stats = nullptr;
  int flags = 0;
  int k_splits = 1;
  int corr_seg_begin = -1;  // first A segment of the split-correction terms (fp32-parity problems), -1 = none
  int max_clusters = 0;  // 0 = all SMs
};

struct DeviceInfo {
  int device = -1;
  int num_sms = 0;
};

// Encode the TMA descriptors of a problem and launch the matching gemm_kernel instantiation.  Defined in
// gemm_launch.cuh, which only capi_gemm.cu includes, so that every kernel is instantiated (and compiled) once.
int run_gemm(const GemmProblem& p, const DeviceInfo& dev, cudaStream_t stream);

}  // namespace srg
