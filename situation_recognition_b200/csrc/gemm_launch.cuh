// GEMM launcher: maps a problem description onto one gemm_kernel instantiation (descriptor encoding + dispatch).
// Included by exactly one translation unit (capi_gemm.cu); everybody else calls run_gemm() declared in host.cuh.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "host.cuh"

namespace srg {

template <int CG, int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool F32>
inline int launch_gemm_inst(const GemmProblem& p, const GemmMaps& maps, const GemmArgs& args, const DeviceInfo& dev,
                            cudaStream_t stream) {
  using Cfg = GemmCfg<CG, BLOCK_N, EPI, F32>;
  auto kern = gemm_kernel<CG, BLOCK_N, A_MN, B_MN, EPI, F32>;
  // the attribute is per device (and per function): one flag per device this process has launched on
  static bool attr_done[kMaxDevices] = {};
  SRG_CHECK(dev.device >= 0 && dev.device < kMaxDevices, "device index %d out of range", dev.device);
  if (!attr_done[dev.device]) {
    SRG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done[dev.device] = true;
  }
  const int num_m_tiles = (p.M + kTileM * CG - 1) / (kTileM * CG);
  const int total_work = num_m_tiles * (p.N / BLOCK_N) * args.k_splits;
  int clusters = dev.num_sms / CG;
  if (p.max_clusters > 0 && p.max_clusters < clusters) clusters = p.max_clusters;
  if (total_work < clusters) clusters = total_work;
  if (clusters <= 0) return SRG_OK;

  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * CG, 1, 1);
  cfg.blockDim = dim3(kNumThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CG > 1) ? 1 : 0;
  int rec = -1;
  if (g_prof.enabled && g_prof.n < kProfMaxRecords) {
    rec = g_prof.n++;
    if (rec >= g_prof.created) {
      SRG_CUDA(cudaEventCreate(&g_prof.ev0[rec]));
      SRG_CUDA(cudaEventCreate(&g_prof.ev1[rec]));
      g_prof.created = rec + 1;
    }
    if (g_prof.rt == nullptr)
      SRG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_prof.rt), sizeof(int) * 2 * kProfMaxRecords, cudaHostAllocDefault));
    g_prof.kind[rec] = prof_kind(EPI, A_MN, B_MN);
    g_prof.flops[rec] = 2.0 * p.M * static_cast<double>(p.N) * args.total_kb * kBlockK;
    g_prof.m_host[rec] = p.M;
    g_prof.nseg[rec] = p.nseg;
    g_prof.kb_host[rec] = args.total_kb;
    g_prof.mn[rec] = 2.0 * static_cast<double>(p.N);
    g_prof.rt[2 * rec] = g_prof.rt[2 * rec + 1] = -1;
    SRG_CUDA(cudaEventRecord(g_prof.ev0[rec], stream));
  }
  SRG_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, args));
  if (rec >= 0) {
    SRG_CUDA(cudaEventRecord(g_prof.ev1[rec], stream));
    if (p.m_dev != nullptr)
      SRG_CUDA(cudaMemcpyAsync(&g_prof.rt[2 * rec], p.m_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (p.k_dev != nullptr)
      SRG_CUDA(cudaMemcpyAsync(&g_prof.rt[2 * rec + 1], p.k_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return SRG_OK;
}

template <int CG>
inline int tile_n() { return CG == 2 ? 256 : 128; }

template <int CG, int BLOCK_N>
inline int dispatch_gemm(const GemmProblem& p, const GemmMaps& maps, const GemmArgs& args, const DeviceInfo& dev,
                         cudaStream_t stream) {
#define SRG_CASE(AMN, BMN, E, F)                                                                     \
  if (p.a_mn == AMN && p.b_mn == BMN && p.epi == E && p.f32 == F)                                    \
    return launch_gemm_inst<CG, BLOCK_N, AMN, BMN, E, F>(p, maps, args, dev, stream);
  SRG_CASE(false, false, EPI_STORE_BF16, false)
  SRG_CASE(false, false, EPI_STORE_F32, false)
  SRG_CASE(false, false, EPI_ZR, false)
  SRG_CASE(false, false, EPI_ZR, true)
  SRG_CASE(false, false, EPI_H, false)
  SRG_CASE(false, false, EPI_H, true)
  SRG_CASE(false, false, EPI_LOGITS, false)
  SRG_CASE(false, false, EPI_LOGITS, true)
  SRG_CASE(false, true, EPI_STORE_BF16, false)
  SRG_CASE(false, true, EPI_STORE_BF16, true)
  SRG_CASE(false, true, EPI_STORE_F32, false)
  SRG_CASE(false, true, EPI_DRH, false)
  SRG_CASE(false, true, EPI_DH, false)
  SRG_CASE(true, true, EPI_STORE_F32, false)
#undef SRG_CASE
  return set_error(SRG_ERR_UNSUPPORTED, "no gemm kernel for a_mn=%d b_mn=%d epi=%d f32=%d", (int)p.a_mn, (int)p.b_mn,
                   p.epi, (int)p.f32);
}

inline int io_box_cols(int dtype) { return dtype == DT_F32 ? 32 : 64; }

// Encode descriptors and launch.
int run_gemm(const GemmProblem& p, const DeviceInfo& dev, cudaStream_t stream) {
  SRG_CHECK(p.cg == 1 || p.cg == 2, "cg must be 1 or 2");
  const int block_n = (p.cg == 2) ? 256 : 128;
  SRG_CHECK(p.M > 0, "gemm: M must be positive");
  SRG_CHECK(p.N > 0 && p.N % block_n == 0, "gemm: N=%d must be a positive multiple of %d", p.N, block_n);
  SRG_CHECK(p.nseg >= 1 && p.nseg <= kMaxSeg, "gemm: bad segment count %d", p.nseg);
  SRG_CHECK(p.k_splits >= 1, "gemm: k_splits must be >= 1");
  SRG_CHECK(p.k_splits == 1 || (p.epi == EPI_STORE_F32 && (p.flags & FLAG_REDUCE)),
            "gemm: split-K needs the fp32 reduce epilogue");

  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  GemmArgs args;
  memset(&args, 0, sizeof(args));
  args.M = p.M;
  args.N = p.N;
  args.m_dev = p.m_dev;
  args.k_dev = p.k_dev;
  args.nseg = p.nseg;
  if (p.k_dev != nullptr) {
    SRG_CHECK(p.flags & FLAG_BK_A, "gemm: device-resident K needs B to follow A's K coordinate");
    for (int s = 1; s < p.nseg; ++s)
      SRG_CHECK(p.seg[s].k_len == p.seg[0].k_len, "gemm: device-resident K needs equally long segments");
  }
  if (p.flags & FLAG_BK_A) {
    SRG_CHECK(p.a_mn && p.b_mn, "gemm: FLAG_BK_A is for MN-major operands (weight gradients)");
    for (int s = 0; s < p.nseg; ++s)
      SRG_CHECK(p.seg[s].a.ptr == p.seg[0].a.ptr, "gemm: FLAG_BK_A needs every segment in the same A matrix");
  }

  // A segments: distinct matrices share a descriptor slot
  const void* aptr[kMaxAMaps];
  int n_amaps = 0;
  int ktot = 0;
  args.corr_kb_begin = 1 << 30;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& sg = p.seg[s];
    // a K tail (< 64) is legal only in the last segment: TMA zero-fills beyond the tensor extent of A and B
    SRG_CHECK(sg.k_len > 0 && (sg.k_len % kBlockK == 0 || s == p.nseg - 1), "gemm: segment %d K=%d not a multiple of 64",
              s, sg.k_len);
    SRG_CHECK(sg.a.dtype == DT_BF16, "gemm: A must be bf16");
    int mi = -1;
    for (int j = 0; j < n_amaps; ++j)
      if (aptr[j] == sg.a.ptr) mi = j;
    if (mi < 0) {
      SRG_CHECK(n_amaps < kMaxAMaps, "gemm: more than %d distinct A matrices", kMaxAMaps);
      mi = n_amaps++;
      aptr[mi] = sg.a.ptr;
      if (!p.a_mn) {
        SRG_CHECK(sg.a.rows >= p.M, "gemm: A segment %d has %lld rows < M=%d", s, (long long)sg.a.rows, p.M);
        SRG_TRY(make_tmap(&maps.a[mi], sg.a.ptr, DT_BF16, p.M, sg.a.cols, sg.a.ld, kTileM, kBlockK));
      } else {
        SRG_CHECK(sg.a.cols >= p.M, "gemm: MN-major A segment %d has %lld cols < M=%d", s, (long long)sg.a.cols, p.M);
        SRG_TRY(make_tmap(&maps.a[mi], sg.a.ptr, DT_BF16, sg.a.rows, p.M, sg.a.ld, kBlockK, 64));
      }
    }
    const int64_t kext = p.a_mn ? sg.a.rows : sg.a.cols;
    SRG_CHECK(sg.k_off >= 0 && sg.k_off + sg.k_len <= kext, "gemm: segment %d K range [%d,%d) outside A (%lld)", s,
              sg.k_off, sg.k_off + sg.k_len, (long long)kext);
    if (s == p.corr_seg_begin) args.corr_kb_begin = ktot / kBlockK;
    args.seg_map[s] = mi;
    args.seg_acol[s] = sg.k_off;
    args.seg_kb[s] = (sg.k_len + kBlockK - 1) / kBlockK;
    ktot += sg.k_len;
    if (sg.k_len % kBlockK != 0)
      SRG_CHECK(sg.k_off + sg.k_len == kext, "gemm: a K tail must end at the tensor extent (zero fill)");
  }
  for (int j = n_amaps; j < kMaxAMaps; ++j) maps.a[j] = maps.a[0];
  args.total_kb = (ktot + kBlockK - 1) / kBlockK;
  // every split must own at least one k-block (an empty split would never signal its accumulator)
  int k_splits = p.k_splits < args.total_kb ? p.k_splits : args.total_kb;
  {
    const int per = (args.total_kb + k_splits - 1) / k_splits;
    k_splits = (args.total_kb + per - 1) / per;
  }

  SRG_CHECK(p.b.dtype == DT_BF16, "gemm: B must be bf16");
  if (!p.b_mn) {
    SRG_CHECK(p.b.rows >= p.N && p.b.cols >= ktot, "gemm: B [%lld,%lld] smaller than [N=%d,K=%d]",
              (long long)p.b.rows, (long long)p.b.cols, p.N, ktot);
    SRG_CHECK(ktot % kBlockK == 0 || p.b.cols == ktot, "gemm: K tail needs B cols == K");
    SRG_TRY(make_tmap(&maps.b, p.b.ptr, DT_BF16, p.N, ktot, p.b.ld, block_n / p.cg, kBlockK));
  } else if (p.flags & FLAG_BK_A) {
    int64_t kmax = 0;
    for (int s = 0; s < p.nseg; ++s) kmax = std::max<int64_t>(kmax, p.seg[s].k_off + p.seg[s].k_len);
    SRG_CHECK(p.b.rows >= kmax && p.b.cols >= p.N, "gemm: MN-major B [%lld,%lld] smaller than [K=%lld,N=%d]",
              (long long)p.b.rows, (long long)p.b.cols, (long long)kmax, p.N);
    SRG_TRY(make_tmap(&maps.b, p.b.ptr, DT_BF16, p.b.rows, p.N, p.b.ld, kBlockK, 64));
  } else {
    SRG_CHECK(p.b.rows >= ktot && p.b.cols >= p.N, "gemm: MN-major B [%lld,%lld] smaller than [K=%d,N=%d]",
              (long long)p.b.rows, (long long)p.b.cols, ktot, p.N);
    SRG_CHECK(ktot % kBlockK == 0 || p.b.rows == ktot, "gemm: K tail needs B rows == K");
    SRG_TRY(make_tmap(&maps.b, p.b.ptr, DT_BF16, ktot, p.N, p.b.ld, kBlockK, 64));
  }
  bool have_io0 = false;
  for (int i = 0; i < kMaxIoMaps; ++i) {
    if (p.io[i].dtype == DT_NONE || p.io[i].ptr == nullptr) continue;
    // the software-pipelined epilogues move [128 rows x 32 columns] boxes per CTA (bf16: 64-byte rows), the others
    // [32 rows x 128 B] boxes per warp
    const bool pipe = !p.f32 && (p.epi == EPI_ZR || p.epi == EPI_H || p.epi == EPI_DH || p.epi == EPI_DRH);
    SRG_TRY(make_tmap(&maps.io[i], p.io[i].ptr, p.io[i].dtype, p.io[i].rows, p.io[i].cols, p.io[i].ld,
                      pipe ? kTileM : 32, pipe ? 32 : io_box_cols(p.io[i].dtype)));
    if (i == 0) have_io0 = true;
  }
  (void)have_io0;

  args.k_splits = k_splits;
  args.alpha = p.alpha;
  args.bias = p.bias;
  args.bias_scale = p.bias_scale;
  args.n_split = p.n_split;
  args.n_valid = p.n_valid;
  args.stats = p.stats;
  args.flags = p.flags;
  // weights stay in L2 (evict_last) unless SRG_B_EVICT_LAST=0; the weight-gradient GEMMs have activations on both sides
  static const bool b_evict_last = [] { const char* v = getenv("SRG_B_EVICT_LAST"); return !(v && v[0] == '0'); }();
  args.b_hint = (b_evict_last && !(p.a_mn && p.b_mn)) ? 0x14F0000000000000ull : 0ull;
#ifdef SRG_EPI_TIMING
  if (g_epi_t_dev == nullptr) {
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_epi_t_dev), 64 * sizeof(unsigned long long)));
    SRG_CUDA(cudaMemset(g_epi_t_dev, 0, 64 * sizeof(unsigned long long)));
  }
  args.epi_t = g_epi_t_dev;
#endif

  if (p.cg == 2) return dispatch_gemm<2, 256>(p, maps, args, dev, stream);
  return dispatch_gemm<1, 128>(p, maps, args, dev, stream);
}

}  // namespace srg
