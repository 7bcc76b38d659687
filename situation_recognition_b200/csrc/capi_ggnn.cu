// C-ABI of the GGNN role-graph stage (include/srggnn.h): handle, weight packing, forward and backward
// orchestration.  Every contraction is one launch of the tcgen05 kernel in gemm.cuh; everything HBM-bound
// is in kernels_elementwise.cu.
#include <new>
#include "elementwise.cuh"
#include "host.cuh"
#include "../../include/srggnn.h"

using namespace srg;

namespace srg {
int query_device(DeviceInfo* d);
}

namespace {
constexpr int kMaxT = 8;
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
}  // namespace

struct srg_handle {
  int device = 0, D = 0, R = 0, T = 0, V = 0, n_roles = 0, L = 0;
  int Vpad = 0, Lpad = 0;
  int cg = 2;
  DeviceInfo dev;
  int32_t* d_verb2roles = nullptr;
  int32_t* d_role_count = nullptr;
  bool tables_set = false;
  int* d_bad = nullptr;
  // packed tensor-core operands
  int packed_prec = -1;
  int alloc_split = 0;  // K replication factor the buffers were allocated for (1 or 3)
  bf16 *Wp = nullptr, *Wzr = nullptr, *Wh = nullptr, *Wcn = nullptr, *Wcv = nullptr;
  bf16 *Wm_stack = nullptr, *UW_stack = nullptr, *Uh = nullptr;
  float *bp = nullptr, *bzr = nullptr, *bh = nullptr, *bcn = nullptr, *bcv = nullptr;
};

namespace {

// ------------------------------------------------------------------------------------------------ workspace
struct StepBufs {
  bf16 *a_hi = nullptr, *a_lo = nullptr, *m_hi = nullptr, *m_lo = nullptr, *rh_hi = nullptr, *rh_lo = nullptr;
  bf16 *r = nullptr, *hc = nullptr;
  void* z = nullptr;
};

struct PathBufs {
  int M = 0;
  float* h32 = nullptr;
  bf16* hb_hi[kMaxT + 1] = {};
  bf16* hb_lo[kMaxT + 1] = {};
  StepBufs st[kMaxT];
  bf16 *xd_hi = nullptr, *xd_lo = nullptr;
  float* stats = nullptr;
  float* mask = nullptr;
  // backward
  float *dh = nullptr, *dh_acc = nullptr, *da = nullptr;
  bf16 *dpre_z[2] = {}, *dpre_h[2] = {}, *dpre_r = nullptr, *dm = nullptr, *adm = nullptr, *dlb = nullptr;
  size_t bytes = 0;
};

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {
    // absolute 1024-byte alignment (the size query reserves the worst-case slack for it)
    if (base) off = (1024 - (reinterpret_cast<uintptr_t>(base) & 1023)) & 1023;
  }
  template <typename T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += (n * sizeof(T) + 1023) & ~static_cast<size_t>(1023);
    return p;
  }
};

// Lay the buffers of one path out in `ws` (ws == nullptr: only compute the size).
PathBufs carve(const srg_handle* h, int mode, int B, int prec, int save, void* ws) {
  PathBufs pb;
  const int D = h->D, T = h->T;
  const int M = (mode == SRG_MODE_NOUN) ? B * h->R : B;
  const size_t MD = static_cast<size_t>(M) * D;
  const bool f32 = (prec == SRG_PREC_FP32);
  const int npad = (mode == SRG_MODE_NOUN) ? h->Lpad : h->Vpad;
  Bump bump(ws);
  pb.M = M;
  pb.h32 = bump.take<float>(MD);
  const int n_hb = save ? T + 1 : 2;
  bf16* hbh[kMaxT + 1];
  bf16* hbl[kMaxT + 1];
  for (int i = 0; i < n_hb; ++i) {
    hbh[i] = bump.take<bf16>(MD);
    hbl[i] = f32 ? bump.take<bf16>(MD) : nullptr;
  }
  for (int t = 0; t <= T; ++t) {
    pb.hb_hi[t] = hbh[save ? t : (t & 1)];
    pb.hb_lo[t] = hbl[save ? t : (t & 1)];
  }
  const int n_st = save ? T : 1;
  StepBufs sts[kMaxT];
  for (int i = 0; i < n_st; ++i) {
    StepBufs& s = sts[i];
    if (mode == SRG_MODE_NOUN) {
      s.a_hi = bump.take<bf16>(MD);
      s.a_lo = f32 ? bump.take<bf16>(MD) : nullptr;
    }
    s.m_hi = bump.take<bf16>(MD);
    s.m_lo = f32 ? bump.take<bf16>(MD) : nullptr;
    s.rh_hi = bump.take<bf16>(MD);
    s.rh_lo = f32 ? bump.take<bf16>(MD) : nullptr;
    s.z = f32 ? static_cast<void*>(bump.take<float>(MD)) : static_cast<void*>(bump.take<bf16>(MD));
    if (save) {
      s.r = bump.take<bf16>(MD);
      s.hc = bump.take<bf16>(MD);
    }
  }
  for (int t = 0; t < T; ++t) pb.st[t] = sts[save ? t : 0];
  pb.xd_hi = bump.take<bf16>(MD);
  pb.xd_lo = f32 ? bump.take<bf16>(MD) : nullptr;
  pb.stats = bump.take<float>(static_cast<size_t>(M) * (npad / 128) * 2);
  if (mode == SRG_MODE_NOUN) pb.mask = bump.take<float>(static_cast<size_t>(B) * h->R * h->R);
  if (save) {
    pb.dh = bump.take<float>(MD);
    pb.dh_acc = bump.take<float>(MD);
    pb.da = bump.take<float>(MD);
    for (int i = 0; i < 2; ++i) {
      pb.dpre_z[i] = bump.take<bf16>(MD);
      pb.dpre_h[i] = bump.take<bf16>(MD);
    }
    pb.dpre_r = bump.take<bf16>(MD);
    pb.dm = bump.take<bf16>(MD);
    pb.adm = (mode == SRG_MODE_NOUN) ? bump.take<bf16>(MD) : nullptr;
    pb.dlb = bump.take<bf16>(static_cast<size_t>(M) * npad);
  }
  pb.bytes = bump.off + 1024;
  return pb;
}

int check_ws(const srg_handle* h, int mode, int B, int prec, int save, void* ws, size_t ws_bytes, PathBufs* out) {
  SRG_CHECK(h != nullptr, "null handle");
  SRG_CHECK(B > 0, "batch must be positive (got %d)", B);
  SRG_CHECK(prec == SRG_PREC_BF16 || prec == SRG_PREC_FP32, "bad precision %d", prec);
  SRG_CHECK(!(save && prec != SRG_PREC_BF16), "save_for_backward requires SRG_PREC_BF16 (the fp32 mode is forward-only)");
  SRG_CHECK(ws != nullptr, "null workspace");
  SRG_CHECK((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
  *out = carve(h, mode, B, prec, save, ws);
  const size_t need = carve(h, mode, B, prec, save, nullptr).bytes;
  if (need > ws_bytes)
    return set_error(SRG_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  return SRG_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM helpers
GemmProblem base_problem(const srg_handle* h, int M, int N) {
  GemmProblem p;
  p.cg = h->cg;
  p.M = M;
  p.N = N;
  return p;
}

void add_seg(GemmProblem& p, const bf16* a, int64_t rows, int64_t cols, int k_off, int k_len) {
  GemmSeg& s = p.seg[p.nseg++];
  s.a = mat(a, rows, cols, cols, DT_BF16);
  s.k_off = k_off;
  s.k_len = k_len;
}

// A = [x0 | x1 | ...] (each [M, D]) -- in fp32 mode replicated as (hi.., hi.., lo..) to meet B = [hi | lo | hi]
void add_split_segs(GemmProblem& p, int M, int D, bool f32, const bf16* const* hi, const bf16* const* lo, int n) {
  for (int i = 0; i < n; ++i) add_seg(p, hi[i], M, D, 0, D);
  if (f32) {
    for (int i = 0; i < n; ++i) add_seg(p, hi[i], M, D, 0, D);
    for (int i = 0; i < n; ++i) add_seg(p, lo[i], M, D, 0, D);
  }
}

int wgrad_splits(const srg_handle* h, int out_rows, int out_cols, int K) {
  const int bn = (h->cg == 2) ? 256 : 128;
  const int tiles = ((out_rows + 128 * h->cg - 1) / (128 * h->cg)) * (out_cols / bn);
  const int clusters = h->dev.num_sms / h->cg;
  int s = (clusters * 7 + tiles - 1) / tiles;  // ~7 waves
  const int kb = (K + kBlockK - 1) / kBlockK;
  if (s > kb) s = kb;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

// dW[out, in] += dY^T[out, M] * X[M, in]   (both operands MN-major, split-K, TMA reduce-add)
int wgrad(const srg_handle* h, const bf16* dY, int64_t dy_cols, int out_rows, const bf16* X, int in_cols, int M,
          float* dW, cudaStream_t s) {
  if (dW == nullptr) return SRG_OK;
  GemmProblem p = base_problem(h, out_rows, in_cols);
  p.a_mn = true;
  p.b_mn = true;
  p.nseg = 1;
  p.seg[0].a = mat(dY, M, dy_cols, dy_cols, DT_BF16);
  p.seg[0].k_off = 0;
  p.seg[0].k_len = M;
  p.b = mat(X, M, in_cols, in_cols, DT_BF16);
  p.epi = EPI_STORE_F32;
  p.flags = FLAG_REDUCE;
  p.io[0] = mat(dW, out_rows, in_cols, in_cols, DT_F32);
  p.k_splits = wgrad_splits(h, out_rows, in_cols, M);
  return run_gemm(p, h->dev, s);
}

// ------------------------------------------------------------------------------------------------ forward
int ggnn_steps(srg_handle* h, int mode, PathBufs& pb, float* h32, const float* mask, int B, int prec, int save,
               cudaStream_t s) {
  const int D = h->D, T = h->T, M = pb.M;
  const bool f32 = (prec == SRG_PREC_FP32);
  const int kmul = f32 ? 3 : 1;
  for (int t = 0; t < T; ++t) {
    StepBufs& st = pb.st[t];
    const bf16* ain_hi = pb.hb_hi[t];
    const bf16* ain_lo = pb.hb_lo[t];
    if (mode == SRG_MODE_NOUN) {
      SRG_TRY(launch_aggregate(h32, mask, B, h->R, D, st.a_hi, st.a_lo, s));
      ain_hi = st.a_hi;
      ain_lo = st.a_lo;
    }
    {  // message: m = a W_p^T + (R | 1) b_p      (model.py:62-64 | 67-77)
      GemmProblem p = base_problem(h, M, D);
      add_split_segs(p, M, D, f32, &ain_hi, &ain_lo, 1);
      p.b = mat(h->Wp, D, static_cast<int64_t>(D) * kmul, static_cast<int64_t>(D) * kmul, DT_BF16);
      p.epi = EPI_STORE_BF16;
      p.bias = h->bp;
      p.bias_scale = (mode == SRG_MODE_NOUN) ? static_cast<float>(h->R) : 1.0f;
      p.flags = f32 ? FLAG_LO : 0;
      p.io[0] = mat(st.m_hi, M, D, D, DT_BF16);
      if (f32) p.io[1] = mat(st.m_lo, M, D, D, DT_BF16);
      SRG_TRY(run_gemm(p, h->dev, s));
    }
    {  // gates: [z | r] = sigmoid([m | h] [W_z U_z ; W_r U_r]^T + b), rh = r * h      (model.py:80-81)
      GemmProblem p = base_problem(h, M, 2 * D);
      const bf16* hi[2] = {st.m_hi, pb.hb_hi[t]};
      const bf16* lo[2] = {st.m_lo, pb.hb_lo[t]};
      add_split_segs(p, M, D, f32, hi, lo, 2);
      p.b = mat(h->Wzr, 2 * D, static_cast<int64_t>(2 * D) * kmul, static_cast<int64_t>(2 * D) * kmul, DT_BF16);
      p.epi = EPI_ZR;
      p.f32 = f32;
      p.bias = h->bzr;
      p.n_split = D;
      p.flags = (f32 ? FLAG_LO : 0) | (save ? FLAG_STASH : 0);
      p.io[0] = mat(st.z, M, D, D, f32 ? DT_F32 : DT_BF16);
      p.io[1] = mat(h32, M, D, D, DT_F32);
      p.io[2] = mat(st.rh_hi, M, D, D, DT_BF16);
      if (f32) p.io[3] = mat(st.rh_lo, M, D, D, DT_BF16);
      if (save) p.io[4] = mat(st.r, M, D, D, DT_BF16);
      SRG_TRY(run_gemm(p, h->dev, s));
    }
    {  // candidate + update: h' = h + z * (tanh([m | rh] [W_h U_h]^T + b) - h)      (model.py:82-84)
      GemmProblem p = base_problem(h, M, D);
      const bf16* hi[2] = {st.m_hi, st.rh_hi};
      const bf16* lo[2] = {st.m_lo, st.rh_lo};
      add_split_segs(p, M, D, f32, hi, lo, 2);
      p.b = mat(h->Wh, D, static_cast<int64_t>(2 * D) * kmul, static_cast<int64_t>(2 * D) * kmul, DT_BF16);
      p.epi = EPI_H;
      p.f32 = f32;
      p.bias = h->bh;
      p.flags = save ? FLAG_STASH : 0;
      p.io[0] = mat(h32, M, D, D, DT_F32);
      p.io[1] = mat(st.z, M, D, D, f32 ? DT_F32 : DT_BF16);
      p.io[2] = mat(pb.hb_hi[t + 1], M, D, D, DT_BF16);
      if (f32) p.io[3] = mat(pb.hb_lo[t + 1], M, D, D, DT_BF16);
      if (save) p.io[4] = mat(st.hc, M, D, D, DT_BF16);
      SRG_TRY(run_gemm(p, h->dev, s));
    }
  }
  return SRG_OK;
}

int classifier_forward(srg_handle* h, int mode, PathBufs& pb, const float* h32, const uint8_t* keep, float drop_p,
                       float* logits, int64_t ldl, int prec, cudaStream_t s) {
  const int D = h->D, M = pb.M;
  const bool f32 = (prec == SRG_PREC_FP32);
  const int kmul = f32 ? 3 : 1;
  const int ncls = (mode == SRG_MODE_NOUN) ? h->L : h->V;
  const int npad = (mode == SRG_MODE_NOUN) ? h->Lpad : h->Vpad;
  SRG_CHECK(ldl >= ncls && (ldl % 4) == 0, "logits leading dimension %lld must be >= %d and a multiple of 4",
            (long long)ldl, ncls);
  SRG_CHECK(drop_p >= 0.f && drop_p < 1.f, "dropout probability %f out of range", drop_p);
  const bf16* x_hi = pb.hb_hi[h->T];
  const bf16* x_lo = pb.hb_lo[h->T];
  if (keep != nullptr && drop_p > 0.f) {
    SRG_TRY(launch_dropout_cast(h32, keep, 1.0f / (1.0f - drop_p), static_cast<int64_t>(M) * D, pb.xd_hi, pb.xd_lo, s));
    x_hi = pb.xd_hi;
    x_lo = pb.xd_lo;
  }
  GemmProblem p = base_problem(h, M, npad);
  add_split_segs(p, M, D, f32, &x_hi, &x_lo, 1);
  const bf16* W = (mode == SRG_MODE_NOUN) ? h->Wcn : h->Wcv;
  p.b = mat(W, npad, static_cast<int64_t>(D) * kmul, static_cast<int64_t>(D) * kmul, DT_BF16);
  p.epi = EPI_LOGITS;
  p.bias = (mode == SRG_MODE_NOUN) ? h->bcn : h->bcv;
  p.n_valid = ncls;
  p.stats = pb.stats;
  p.io[0] = mat(logits, M, ldl < npad ? ldl : npad, ldl, DT_F32);
  return run_gemm(p, h->dev, s);
}

// ------------------------------------------------------------------------------------------------ backward
struct PathGrads {
  float *Wc, *bc;
};

int path_backward(srg_handle* h, int mode, PathBufs& pb, const float* dlogits, int64_t ldl, int B,
                  const uint8_t* keep, float drop_p, const srg_grads* g, cudaStream_t s) {
  const int D = h->D, T = h->T, M = pb.M, R = h->R;
  const int ncls = (mode == SRG_MODE_NOUN) ? h->L : h->V;
  const int npad = (mode == SRG_MODE_NOUN) ? h->Lpad : h->Vpad;
  const bf16* Wc = (mode == SRG_MODE_NOUN) ? h->Wcn : h->Wcv;
  float* gWc = (mode == SRG_MODE_NOUN) ? g->Wc_noun : g->Wc_verb;
  float* gbc = (mode == SRG_MODE_NOUN) ? g->bc_noun : g->bc_verb;
  const int64_t MD = static_cast<int64_t>(M) * D;
  const bool use_drop = (keep != nullptr && drop_p > 0.f);
  const bf16* x = use_drop ? pb.xd_hi : pb.hb_hi[T];

  // ---- classifier (model.py:105-111,152,168)
  SRG_TRY(launch_cast_pad(dlogits, ldl, M, ncls, npad, pb.dlb, s));
  SRG_TRY(launch_colsum(pb.dlb, npad, M, ncls, gbc, 1.f, nullptr, 0.f, s));
  SRG_TRY(wgrad(h, pb.dlb, npad, ncls, x, D, M, gWc, s));
  {
    GemmProblem p = base_problem(h, M, D);
    add_seg(p, pb.dlb, M, npad, 0, npad);
    p.b_mn = true;
    p.b = mat(Wc, npad, D, D, DT_BF16);
    p.epi = EPI_STORE_F32;
    p.io[0] = mat(use_drop ? pb.da : pb.dh, M, D, D, DT_F32);
    SRG_TRY(run_gemm(p, h->dev, s));
    if (use_drop) SRG_TRY(launch_dropout_bwd(pb.da, keep, 1.0f / (1.0f - drop_p), MD, pb.dh, s));
  }

  // dL/dh' of the last step comes from the classifier; from there on every step's "pre" work (GRU-gate derivatives of
  // step t-1) is fused into the epilogue of the GEMM that completes dL/dh of step t (EPI_DH).
  float* dh_acc = pb.dh_acc;
  int cur = 0;
  SRG_TRY(launch_gru_bwd_pre(pb.dh, static_cast<const bf16*>(pb.st[T - 1].z), pb.st[T - 1].hc, pb.hb_hi[T - 1], MD,
                             pb.dpre_z[cur], pb.dpre_h[cur], dh_acc, s));
  for (int t = T - 1; t >= 0; --t) {
    StepBufs& st = pb.st[t];
    bf16* dpre_z = pb.dpre_z[cur];
    bf16* dpre_h = pb.dpre_h[cur];
    {  // d(r*h) = dpre_h U_h ; fused: dpre_r = drh*h*r*(1-r), dh_acc += drh*r
      GemmProblem p = base_problem(h, M, D);
      add_seg(p, dpre_h, M, D, 0, D);
      p.b_mn = true;
      p.b = mat(h->Uh, D, D, D, DT_BF16);
      p.epi = EPI_DRH;
      p.io[0] = mat(pb.hb_hi[t], M, D, D, DT_BF16);
      p.io[1] = mat(st.r, M, D, D, DT_BF16);
      p.io[2] = mat(pb.dpre_r, M, D, D, DT_BF16);
      p.io[3] = mat(dh_acc, M, D, D, DT_F32);
      SRG_TRY(run_gemm(p, h->dev, s));
    }
    {  // dm = [dpre_h | dpre_z | dpre_r] [W_h ; W_z ; W_r]
      GemmProblem p = base_problem(h, M, D);
      add_seg(p, dpre_h, M, D, 0, D);
      add_seg(p, dpre_z, M, D, 0, D);
      add_seg(p, pb.dpre_r, M, D, 0, D);
      p.b_mn = true;
      p.b = mat(h->Wm_stack, 3 * D, D, D, DT_BF16);
      p.epi = EPI_STORE_BF16;
      p.io[0] = mat(pb.dm, M, D, D, DT_BF16);
      SRG_TRY(run_gemm(p, h->dev, s));
    }
    // the aggregation is linear and commutes with W_p: push dm through its transpose first (noun mode)
    const bf16* adm = pb.dm;
    if (mode == SRG_MODE_NOUN) {
      SRG_TRY(launch_aggregate_t_bf16(pb.dm, pb.mask, B, R, D, pb.adm, s));
      adm = pb.adm;
    }
    {  // dL/dh_t = dh_acc + [dpre_z | dpre_r | adm] [U_z ; U_r ; W_p], then the gate derivatives of step t-1
      GemmProblem p = base_problem(h, M, D);
      add_seg(p, dpre_z, M, D, 0, D);
      add_seg(p, pb.dpre_r, M, D, 0, D);
      add_seg(p, adm, M, D, 0, D);
      p.b_mn = true;
      p.b = mat(h->UW_stack, 3 * D, D, D, DT_BF16);
      p.epi = EPI_DH;
      p.io[0] = mat(dh_acc, M, D, D, DT_F32);
      if (t > 0) {
        p.flags = FLAG_NEXT;
        p.io[1] = mat(pb.st[t - 1].z, M, D, D, DT_BF16);
        p.io[2] = mat(pb.st[t - 1].hc, M, D, D, DT_BF16);
        p.io[3] = mat(pb.hb_hi[t - 1], M, D, D, DT_BF16);
        p.io[4] = mat(pb.dpre_z[cur ^ 1], M, D, D, DT_BF16);
        p.io[5] = mat(pb.dpre_h[cur ^ 1], M, D, D, DT_BF16);
      }
      SRG_TRY(run_gemm(p, h->dev, s));
    }
    // ---- weight gradients of this step (the 7 linears are shared by all steps and both paths: accumulate)
    const bf16* msg_in = (mode == SRG_MODE_NOUN) ? st.a_hi : pb.hb_hi[t];
    SRG_TRY(wgrad(h, pb.dm, D, D, msg_in, D, M, g->W_p, s));
    SRG_TRY(wgrad(h, dpre_z, D, D, st.m_hi, D, M, g->W_z, s));
    SRG_TRY(wgrad(h, dpre_z, D, D, pb.hb_hi[t], D, M, g->U_z, s));
    SRG_TRY(wgrad(h, pb.dpre_r, D, D, st.m_hi, D, M, g->W_r, s));
    SRG_TRY(wgrad(h, pb.dpre_r, D, D, pb.hb_hi[t], D, M, g->U_r, s));
    SRG_TRY(wgrad(h, dpre_h, D, D, st.m_hi, D, M, g->W_h, s));
    SRG_TRY(wgrad(h, dpre_h, D, D, st.rh_hi, D, M, g->U_h, s));
    ColsumJob jobs[4] = {{dpre_z, g->b_Wz, g->b_Uz, 1.f},
                         {pb.dpre_r, g->b_Wr, g->b_Ur, 1.f},
                         {dpre_h, g->b_Wh, g->b_Uh, 1.f},
                         {pb.dm, g->b_p, nullptr, (mode == SRG_MODE_NOUN) ? static_cast<float>(R) : 1.f}};
    SRG_TRY(launch_colsum_multi(jobs, 4, D, M, D, s));
    cur ^= 1;
  }
  float* dh = dh_acc;
  pb.dh = dh;  // gradient w.r.t. the initial node states
  return SRG_OK;
}

int ensure_pack_buffers(srg_handle* h, int split) {
  if (h->alloc_split >= split) return SRG_OK;
  const size_t D = h->D;
  auto re = [&](bf16** p, size_t n) -> int {
    if (*p) cudaFree(*p);
    *p = nullptr;
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(bf16)));
    return SRG_OK;
  };
  SRG_TRY(re(&h->Wp, D * D * split));
  SRG_TRY(re(&h->Wzr, 2 * D * 2 * D * split));
  SRG_TRY(re(&h->Wh, D * 2 * D * split));
  SRG_TRY(re(&h->Wcn, static_cast<size_t>(h->Lpad) * D * split));
  SRG_TRY(re(&h->Wcv, static_cast<size_t>(h->Vpad) * D * split));
  if (!h->Wm_stack) {
    SRG_TRY(re(&h->Wm_stack, 3 * D * D));
    SRG_TRY(re(&h->UW_stack, 3 * D * D));
    SRG_TRY(re(&h->Uh, D * D));
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->bp), D * sizeof(float)));
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->bzr), 2 * D * sizeof(float)));
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->bh), D * sizeof(float)));
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->bcn), h->Lpad * sizeof(float)));
    SRG_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->bcv), h->Vpad * sizeof(float)));
  }
  h->alloc_split = split;
  return SRG_OK;
}

}  // namespace

// ================================================================================================== C ABI
extern "C" {

int srg_version(void) { return 100; }

int srg_create(srg_handle** out, int device, int D, int R, int T, int n_verbs, int n_roles, int n_labels) {
  SRG_CHECK(out != nullptr, "srg_create: null output pointer");
  SRG_CHECK(D > 0 && D % 256 == 0, "hidden size D=%d must be a positive multiple of 256", D);
  SRG_CHECK(R >= 1 && R <= 8, "max role count R=%d must be in [1, 8]", R);
  SRG_CHECK(T >= 1 && T <= kMaxT, "propagation steps T=%d must be in [1, %d]", T, kMaxT);
  SRG_CHECK(n_verbs > 0 && n_roles > 0 && n_labels > 0, "vocabulary sizes must be positive");
  SRG_CUDA(cudaSetDevice(device));
  srg_handle* h = new (std::nothrow) srg_handle();
  SRG_CHECK(h != nullptr, "out of host memory");
  h->device = device;
  h->D = D; h->R = R; h->T = T; h->V = n_verbs; h->n_roles = n_roles; h->L = n_labels;
  h->Vpad = round_up(n_verbs, 256);
  h->Lpad = round_up(n_labels, 256);
  int rc = query_device(&h->dev);
  if (rc != SRG_OK) { delete h; return rc; }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->d_verb2roles), sizeof(int32_t) * n_verbs * R);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->d_role_count), sizeof(int32_t) * n_verbs);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->d_bad), sizeof(int));
  if (e != cudaSuccess) {
    delete h;
    return set_error(SRG_ERR_CUDA, "srg_create: cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return SRG_OK;
}

int srg_destroy(srg_handle* h) {
  if (!h) return SRG_OK;
  void* ptrs[] = {h->d_verb2roles, h->d_role_count, h->d_bad, h->Wp, h->Wzr, h->Wh, h->Wcn, h->Wcv, h->Wm_stack,
                  h->UW_stack, h->Uh, h->bp, h->bzr, h->bh, h->bcn, h->bcv};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete h;
  return SRG_OK;
}

int srg_set_cta_group(srg_handle* h, int cta_group) {
  SRG_CHECK(h != nullptr, "null handle");
  SRG_CHECK(cta_group == 1 || cta_group == 2, "cta_group must be 1 or 2");
  h->cg = cta_group;
  return SRG_OK;
}

int srg_set_tables(srg_handle* h, const int32_t* verb2roles, const int32_t* role_count) {
  SRG_CHECK(h && verb2roles && role_count, "srg_set_tables: null argument");
  for (int v = 0; v < h->V; ++v) {
    SRG_CHECK(role_count[v] >= 0 && role_count[v] <= h->R, "role_count[%d]=%d outside [0,%d]", v, role_count[v], h->R);
    for (int r = 0; r < h->R; ++r) {
      const int id = verb2roles[v * h->R + r];
      SRG_CHECK(id >= 0 && id <= h->n_roles, "verb2roles[%d,%d]=%d outside [0,%d]", v, r, id, h->n_roles);
    }
  }
  SRG_CUDA(cudaMemcpy(h->d_verb2roles, verb2roles, sizeof(int32_t) * h->V * h->R, cudaMemcpyHostToDevice));
  SRG_CUDA(cudaMemcpy(h->d_role_count, role_count, sizeof(int32_t) * h->V, cudaMemcpyHostToDevice));
  h->tables_set = true;
  return SRG_OK;
}

int srg_gather_mask(srg_handle* h, const int64_t* verb, int B, int64_t* role_idx, float* mask, int* bad_verb,
                    void* stream) {
  SRG_CHECK(h != nullptr && h->tables_set, "srg_gather_mask: call srg_set_tables first");
  SRG_CHECK(B >= 0, "negative batch");
  if (B == 0) return SRG_OK;
  SRG_CHECK(verb != nullptr, "null verb pointer");
  return launch_gather_mask(h->d_verb2roles, h->d_role_count, h->V, h->R, verb, B, role_idx, mask, bad_verb,
                            static_cast<cudaStream_t>(stream));
}

int srg_pack_weights(srg_handle* h, const srg_params* p, int precision, void* stream) {
  SRG_CHECK(h && p, "srg_pack_weights: null argument");
  SRG_CHECK(precision == SRG_PREC_BF16 || precision == SRG_PREC_FP32, "bad precision %d", precision);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int D = h->D;
  const int split = (precision == SRG_PREC_FP32) ? 3 : 1;
  SRG_TRY(ensure_pack_buffers(h, split));
  // K layout per weight group: [hi-block | lo-block | hi-block] (fp32 mode) or [hi-block] (bf16 mode);
  // a block holds the concatenated inputs of that GEMM, e.g. [W_z | U_z].
  auto pack = [&](const float* src, int rows, int rows_pad, bf16* dst, int64_t ld, int64_t row_off, int block_cols,
                  int col_in_block) -> int {
    bf16* d = dst + row_off * ld;
    SRG_TRY(launch_pack_weight(src, rows, D, rows_pad, d, ld, col_in_block, 0, s));
    if (split == 3) {
      SRG_TRY(launch_pack_weight(src, rows, D, rows_pad, d, ld, block_cols + col_in_block, 1, s));
      SRG_TRY(launch_pack_weight(src, rows, D, rows_pad, d, ld, 2 * block_cols + col_in_block, 0, s));
    }
    return SRG_OK;
  };
  const int64_t ld1 = static_cast<int64_t>(D) * split, ld2 = static_cast<int64_t>(2 * D) * split;
  SRG_TRY(pack(p->W_p, D, D, h->Wp, ld1, 0, D, 0));
  SRG_TRY(pack(p->W_z, D, D, h->Wzr, ld2, 0, 2 * D, 0));
  SRG_TRY(pack(p->U_z, D, D, h->Wzr, ld2, 0, 2 * D, D));
  SRG_TRY(pack(p->W_r, D, D, h->Wzr, ld2, D, 2 * D, 0));
  SRG_TRY(pack(p->U_r, D, D, h->Wzr, ld2, D, 2 * D, D));
  SRG_TRY(pack(p->W_h, D, D, h->Wh, ld2, 0, 2 * D, 0));
  SRG_TRY(pack(p->U_h, D, D, h->Wh, ld2, 0, 2 * D, D));
  SRG_TRY(pack(p->Wc_noun, h->L, h->Lpad, h->Wcn, ld1, 0, D, 0));
  SRG_TRY(pack(p->Wc_verb, h->V, h->Vpad, h->Wcv, ld1, 0, D, 0));
  if (precision == SRG_PREC_BF16) {
    // backward operands (MN-major B = stacks along K)
    SRG_TRY(launch_pack_weight(p->W_h, D, D, D, h->Wm_stack, D, 0, 0, s));
    SRG_TRY(launch_pack_weight(p->W_z, D, D, D, h->Wm_stack + static_cast<size_t>(D) * D, D, 0, 0, s));
    SRG_TRY(launch_pack_weight(p->W_r, D, D, D, h->Wm_stack + 2 * static_cast<size_t>(D) * D, D, 0, 0, s));
    SRG_TRY(launch_pack_weight(p->U_z, D, D, D, h->UW_stack, D, 0, 0, s));
    SRG_TRY(launch_pack_weight(p->U_r, D, D, D, h->UW_stack + static_cast<size_t>(D) * D, D, 0, 0, s));
    SRG_TRY(launch_pack_weight(p->W_p, D, D, D, h->UW_stack + 2 * static_cast<size_t>(D) * D, D, 0, 0, s));
    SRG_TRY(launch_pack_weight(p->U_h, D, D, D, h->Uh, D, 0, 0, s));
  }
  SRG_TRY(launch_pack_bias(p->b_p, nullptr, D, D, h->bp, s));
  SRG_TRY(launch_pack_bias(p->b_Wz, p->b_Uz, D, D, h->bzr, s));
  SRG_TRY(launch_pack_bias(p->b_Wr, p->b_Ur, D, D, h->bzr + D, s));
  SRG_TRY(launch_pack_bias(p->b_Wh, p->b_Uh, D, D, h->bh, s));
  SRG_TRY(launch_pack_bias(p->bc_noun, nullptr, h->L, h->Lpad, h->bcn, s));
  SRG_TRY(launch_pack_bias(p->bc_verb, nullptr, h->V, h->Vpad, h->bcv, s));
  h->packed_prec = precision;
  return SRG_OK;
}

size_t srg_workspace_bytes(srg_handle* h, int mode, int B, int precision, int save_for_backward) {
  if (!h || B <= 0) return 0;
  return carve(h, mode, B, precision, save_for_backward, nullptr).bytes;
}

int srg_nouns_forward(srg_handle* h, const float* feat, const int64_t* verb, int B, const float* role_emb,
                      const float* verb_emb, const uint8_t* keep, float drop_p, float* logits, int64_t ldl,
                      int precision, int save_for_backward, void* workspace, size_t workspace_bytes, void* stream) {
  PathBufs pb;
  SRG_TRY(check_ws(h, SRG_MODE_NOUN, B, precision, save_for_backward, workspace, workspace_bytes, &pb));
  SRG_CHECK(h->tables_set, "srg_nouns_forward: call srg_set_tables first");
  SRG_CHECK(h->packed_prec == precision, "srg_nouns_forward: weights are not packed for precision %d", precision);
  SRG_CHECK(feat && verb && role_emb && verb_emb && logits, "srg_nouns_forward: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SRG_TRY(launch_gather_mask(h->d_verb2roles, h->d_role_count, h->V, h->R, verb, B, nullptr, pb.mask, nullptr, s));
  SRG_TRY(launch_node_init_noun(feat, role_emb, verb_emb, verb, h->d_verb2roles, h->V, B, h->R, h->D, pb.h32,
                                pb.hb_hi[0], pb.hb_lo[0], s));
  SRG_TRY(ggnn_steps(h, SRG_MODE_NOUN, pb, pb.h32, pb.mask, B, precision, save_for_backward, s));
  return classifier_forward(h, SRG_MODE_NOUN, pb, pb.h32, keep, drop_p, logits, ldl, precision, s);
}

int srg_verb_forward(srg_handle* h, const float* feat, int B, const uint8_t* keep, float drop_p, float* logits,
                     int64_t ldl, int precision, int save_for_backward, void* workspace, size_t workspace_bytes,
                     void* stream) {
  PathBufs pb;
  SRG_TRY(check_ws(h, SRG_MODE_VERB, B, precision, save_for_backward, workspace, workspace_bytes, &pb));
  SRG_CHECK(h->packed_prec == precision, "srg_verb_forward: weights are not packed for precision %d", precision);
  SRG_CHECK(feat && logits, "srg_verb_forward: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SRG_TRY(launch_node_init_verb(feat, B, h->D, pb.h32, pb.hb_hi[0], pb.hb_lo[0], s));
  SRG_TRY(ggnn_steps(h, SRG_MODE_VERB, pb, pb.h32, nullptr, B, precision, save_for_backward, s));
  return classifier_forward(h, SRG_MODE_VERB, pb, pb.h32, keep, drop_p, logits, ldl, precision, s);
}

int srg_ggnn_forward(srg_handle* h, int mode, float* hidden, const float* mask, int B, int precision,
                     int save_for_backward, void* workspace, size_t workspace_bytes, void* stream) {
  PathBufs pb;
  SRG_CHECK(mode == SRG_MODE_NOUN || mode == SRG_MODE_VERB, "bad mode %d", mode);
  SRG_TRY(check_ws(h, mode, B, precision, save_for_backward, workspace, workspace_bytes, &pb));
  SRG_CHECK(h->packed_prec == precision, "srg_ggnn_forward: weights are not packed for precision %d", precision);
  SRG_CHECK(hidden != nullptr, "null hidden state");
  SRG_CHECK(mode == SRG_MODE_VERB || mask != nullptr, "noun mode needs a mask");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SRG_TRY(launch_split_cast(hidden, static_cast<int64_t>(pb.M) * h->D, pb.hb_hi[0], pb.hb_lo[0], s));
  return ggnn_steps(h, mode, pb, hidden, mask, B, precision, save_for_backward, s);
}

int srg_count_targets(srg_handle* h, const int64_t* gt_nouns, int B, float* counts, void* stream) {
  SRG_CHECK(h && gt_nouns && counts, "srg_count_targets: null argument");
  return launch_count_targets(gt_nouns, B, h->R, h->L, counts, static_cast<cudaStream_t>(stream));
}

int srg_nouns_loss(srg_handle* h, const float* logits, int64_t ldl, const int64_t* gt_nouns, int B,
                   const float* counts, float* loss, float* dlogits, float grad_scale, void* stream) {
  SRG_CHECK(h && logits && gt_nouns && counts && loss, "srg_nouns_loss: null argument");
  SRG_CHECK(ldl >= h->L, "srg_nouns_loss: ldl %lld < n_labels %d", (long long)ldl, h->L);
  return launch_nouns_ce(logits, ldl, h->L, gt_nouns, B, h->R, counts, loss, dlogits, grad_scale,
                         static_cast<cudaStream_t>(stream));
}

int srg_verb_loss(srg_handle* h, const float* logits, int64_t ldl, const int64_t* gt_verb, int B, float inv_batch,
                  float* loss, float* dlogits, float grad_scale, void* stream) {
  SRG_CHECK(h && logits && gt_verb && loss, "srg_verb_loss: null argument");
  SRG_CHECK(ldl >= h->V, "srg_verb_loss: ldl %lld < n_verbs %d", (long long)ldl, h->V);
  return launch_verb_ce(logits, ldl, h->V, gt_verb, B, inv_batch, loss, dlogits, grad_scale,
                        static_cast<cudaStream_t>(stream));
}

int srg_nouns_backward(srg_handle* h, const float* dlogits, int64_t ldl, const float* feat, const int64_t* verb, int B,
                       const float* role_emb, const float* verb_emb, const uint8_t* keep, float drop_p,
                       const srg_grads* g, void* workspace, size_t workspace_bytes, void* stream) {
  PathBufs pb;
  SRG_TRY(check_ws(h, SRG_MODE_NOUN, B, SRG_PREC_BF16, 1, workspace, workspace_bytes, &pb));
  SRG_CHECK(h->packed_prec == SRG_PREC_BF16, "backward needs bf16-packed weights");
  SRG_CHECK(dlogits && feat && verb && role_emb && verb_emb && g, "srg_nouns_backward: null argument");
  SRG_CHECK(ldl >= h->L, "srg_nouns_backward: ldl %lld < n_labels %d", (long long)ldl, h->L);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SRG_TRY(path_backward(h, SRG_MODE_NOUN, pb, dlogits, ldl, B, keep, drop_p, g, s));
  if (g->role_emb != nullptr && g->verb_emb != nullptr)
    SRG_TRY(launch_node_init_bwd(pb.dh, pb.hb_hi[0], feat, role_emb, verb_emb, verb, h->d_verb2roles, h->n_roles, B,
                                 h->R, h->D, g->role_emb, g->verb_emb, s));
  return SRG_OK;
}

int srg_verb_backward(srg_handle* h, const float* dlogits, int64_t ldl, int B, const uint8_t* keep, float drop_p,
                      const srg_grads* g, void* workspace, size_t workspace_bytes, void* stream) {
  PathBufs pb;
  SRG_TRY(check_ws(h, SRG_MODE_VERB, B, SRG_PREC_BF16, 1, workspace, workspace_bytes, &pb));
  SRG_CHECK(h->packed_prec == SRG_PREC_BF16, "backward needs bf16-packed weights");
  SRG_CHECK(dlogits && g, "srg_verb_backward: null argument");
  SRG_CHECK(ldl >= h->V, "srg_verb_backward: ldl %lld < n_verbs %d", (long long)ldl, h->V);
  return path_backward(h, SRG_MODE_VERB, pb, dlogits, ldl, B, keep, drop_p, g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
