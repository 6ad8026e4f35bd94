// Krylov layer: what PETSc KSP/PC does under solver.solve() (solvers/solver.py:71) for the DPP
// system -- KSPCG (preconditioned norm), KSPGMRES (left PC, classical Gram-Schmidt, restart),
// PCJACOBI / point-block Jacobi / PCFIELDSPLIT (additive, multiplicative), and the
// scale-splitting block Picard iteration on the dpp_delayed_form split (forms/dpp.py:135-205).
// Semantics follow SURVEY Appendix A.3-A.6; tests/ compare against the CPU restatement.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "vector_ops.cuh"

namespace dpp {

// A captured CUDA graph of one launch-bound inner loop (a batch of CG iterations, a GMRES restart cycle):
// every kernel argument of such a loop is the same from one batch / cycle to the next (scalars live in
// device memory), so it is captured once and replayed; `key` ties it to the solver configuration.
struct GraphSlot {
  cudaGraphExec_t exec = nullptr;
  unsigned long long key = 0;
  long long nodes = 0;
  bool broken = false;  // capture failed once on this handle: keep launching directly
};

struct Krylov {
  int64_t nvec = 0;  // 2 * n_nodes
  GraphSlot cg_graph[2], gmres_graph;
  double *b = nullptr, *x = nullptr, *r = nullptr, *p = nullptr, *w = nullptr, *z = nullptr, *t = nullptr;
  double *u0 = nullptr;
  // single-field work vectors for block solves
  double *bi = nullptr, *xi = nullptr, *ri = nullptr, *pi = nullptr, *wi = nullptr, *zi = nullptr;
  double *dinv = nullptr;        // 1/diag(A_bc)  [2n]
  double *pb = nullptr;          // point-block inverse: i00,i01,i11  [3n]
  bool pb_valid = false;
  std::vector<double*> V;        // GMRES basis
  double* d_gm = nullptr;        // device-resident GMRES state (gmres_run_device)
  double* h_gm = nullptr;        // pinned mirror of its scalar head
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double* h_poll = nullptr;      // pinned [2][S_SLOT_SIZE]: scalar snapshots behind each CG batch (cg_run)
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};
  int64_t inner_its = 0;
  int64_t apply_count = 0;
};

namespace {

struct OpSpec {
  int nf;      // 2: monolithic A_bc ; 1: block (row,col)
  int row, col;
  int mode;    // dpp_operator_mode
  int kind = 0;        // 0: (block of) the DPP operator with Dirichlet elimination; 1: unconstrained nodal mass matrix (nf = 1)
  int premasked = 1;   // inputs are exactly zero on eliminated rows (all Krylov vectors); 0: the apply masks them itself
};

struct Tol {
  double rtol, atol, dtol;
  int max_it;
};

struct KspOut {
  int its = 0;
  int reason = 0;
  double rnorm = 0.0;
  std::vector<double> hist;
};

__global__ void k_reciprocal(long long n, const double* __restrict__ d, double* __restrict__ o) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = 1.0 / d[i];
}

__global__ void k_pb_blocks(long long n, const double* __restrict__ diag, const double* __restrict__ mdiag,
                            const uint8_t* __restrict__ mask, double boff, double* __restrict__ pb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = diag[i], c = diag[n + i];
    const double b = (mask[i] || mask[n + i]) ? 0.0 : boff * mdiag[i];
    const double det = a * c - b * b;
    pb[i] = c / det;
    pb[n + i] = -b / det;
    pb[2 * n + i] = a / det;
  }
}

__global__ void k_sub(long long n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ o) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] - b[i];
}

int ew_blocks(const dpp_context* ctx, long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16));
}

unsigned long long mix(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h;
}

// run `body` (which only enqueues work on ctx->stream) through a cached CUDA graph
template <class F>
int run_graphed(dpp_context* ctx, GraphSlot& gs, unsigned long long key, F&& body) {
  if (gs.broken || ctx->world > 1 || getenv("DPP_NO_GRAPH") != nullptr) return body();
  if (gs.exec != nullptr && gs.key == key) {
    DPP_CUDA(cudaGraphLaunch(gs.exec, ctx->stream));
    ctx->launches += gs.nodes;
    return DPP_OK;
  }
  if (gs.exec != nullptr) {
    cudaGraphExecDestroy(gs.exec);
    gs.exec = nullptr;
  }
  const long long l0 = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    gs.broken = true;
    return body();
  }
  const int rc = body();
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
  if (rc != DPP_OK || e != cudaSuccess || g == nullptr) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    gs.broken = true;
    ctx->launches = l0;
    return body();  // nothing was executed during the failed capture
  }
  gs.nodes = ctx->launches - l0;
  const cudaError_t ei = cudaGraphInstantiate(&gs.exec, g, 0);
  cudaGraphDestroy(g);
  if (ei != cudaSuccess) {
    cudaGetLastError();
    gs.exec = nullptr;
    gs.broken = true;
    ctx->launches = l0;
    return body();
  }
  gs.key = key;
  DPP_CUDA(cudaGraphLaunch(gs.exec, ctx->stream));
  return DPP_OK;
}

VecLayout layout(const dpp_context* ctx, int nf) {
  return VecLayout{nf, ctx->n_nodes, ctx->owned_begin, ctx->owned_end};
}

// y = Op x ; optional fused <x,y> partials ; x/y are base pointers (field stride n_nodes)
int apply_spec(dpp_context* ctx, const OpSpec& op, const double* x, double* y, bool want_dot, const double* skip,
               int* nblocks) {
  const int64_t n = ctx->n_nodes;
  if (op.mode == DPP_OP_ASSEMBLED) {
    if (op.nf != 2) {
      ctx->set_error("assembled operator mode supports the monolithic operator only");
      return DPP_ERR_INVALID;
    }
    ctx->krylov->apply_count++;
    return csr_spmv(ctx, x, y, want_dot ? ctx->d_partials : nullptr, nblocks, skip);
  }
  OpArgs a{};
  a.nf = op.nf;
  a.identity_on_masked = 1;
  a.owned_begin = ctx->owned_begin;
  a.owned_end = ctx->owned_end;
  a.dot_partials = want_dot ? ctx->d_partials : nullptr;
  a.skip_flag = skip;
  a.input_premasked = op.premasked;  // Krylov vectors are exactly zero on eliminated rows/columns
  if (op.kind == 1) {  // M x, no boundary conditions (L2 projections)
    a.c = Coef{};
    a.c.cM[0][0] = 1.0;
    a.x[0] = x;
    a.y[0] = y;
  } else if (op.nf == 2) {
    a.c = dpp_coef(ctx);
    for (int f = 0; f < 2; ++f) {
      a.x[f] = x + f * n;
      a.y[f] = y + f * n;
      a.in_mask[f] = a.out_mask[f] = ctx->d_mask + f * n;
    }
  } else {
    a.c = block_coef(ctx, op.row, op.col);
    a.x[0] = x;
    a.y[0] = y;
    a.in_mask[0] = ctx->d_mask + op.col * n;
    a.out_mask[0] = ctx->d_mask + op.row * n;
    a.identity_on_masked = (op.row == op.col) ? 1 : 0;
  }
  ctx->krylov->apply_count++;
  return op_apply(ctx, a, nblocks);
}

int halo(dpp_context* ctx, double* x, int nf) {
  if (ctx->world <= 1) return DPP_OK;
  double* f[2] = {x, x + ctx->n_nodes};
  return comm_halo_exchange(ctx, f, nf);
}

// ------------------------------------------------------------------------------------------------
// KSPCG with fused Jacobi / no preconditioner: the whole iteration runs from device scalars, the
// host polls the converged flag every `check_every` iterations (kernels launched past convergence
// are no-ops, so the reported iteration count is exact).
// ------------------------------------------------------------------------------------------------
struct CgWork {
  double *r, *p, *w, *z;
};

struct Pc;  // fwd
int pc_apply(dpp_context* ctx, Pc& pc, const double* r, double* z);

struct Pc {
  int type = DPP_PC_NONE;
  const double* dinv = nullptr;  // Jacobi (fusable)
  const dpp_options* opt = nullptr;
  bool fusable() const { return type == DPP_PC_NONE || type == DPP_PC_JACOBI; }
};

int cg_run(dpp_context* ctx, const OpSpec& op, Pc& pc, const double* b, double* x, CgWork wk, const Tol& tol, int slot,
           int check_every, int hist_cap, KspOut* out) {
  const VecLayout L = layout(ctx, op.nf);
  const int64_t len = (int64_t)op.nf * ctx->n_nodes;
  const bool fused = pc.fusable();
  const double* dinv = (pc.type == DPP_PC_JACOBI) ? pc.dinv : nullptr;
  double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  if (hist_cap > ctx->hist_cap[slot]) {
    if (ctx->d_hist[slot]) cudaFree(ctx->d_hist[slot]);
    ctx->d_hist[slot] = nullptr;
    DPP_CHECK(dev_alloc(ctx, &ctx->d_hist[slot], hist_cap));
    ctx->hist_cap[slot] = hist_cap;
  }
  DPP_CHECK(scalars_init(ctx, slot, tol.rtol, tol.atol, tol.dtol, tol.max_it, std::min(hist_cap, ctx->hist_cap[slot])));
  DPP_CHECK(vec_zero(ctx, x, len));
  DPP_CHECK(vec_copy(ctx, wk.r, b, len));
  const double* h = ctx->h_scalars + (size_t)slot * S_SLOT_SIZE;
  if (fused && op.kind == 0 && cg_fused_available(ctx, op.nf, op.mode, pc.type)) {
    // two kernels per iteration (cg_fused_uniform.cu): p, x updates live inside the apply kernel
    const Coef coef = op.nf == 2 ? dpp_coef(ctx) : block_coef(ctx, op.row, op.col);
    double* dtab = ctx->d_dtab + (size_t)slot * 128;
    const int fld[2] = {op.nf == 2 ? 0 : op.row, 1};
    DPP_CHECK(cg_fused_table(ctx, coef, op.nf, pc.type, fld, dtab));
    DPP_CHECK(cg_fused_begin(ctx, op.nf, b));
    DPP_CHECK(cg_fused_halo_r(ctx, op.nf, false, slot));
    DPP_CHECK(cg_fused_rz_init(ctx, op.nf, fld, slot, dtab));
    DPP_CHECK(scalars_fetch(ctx, slot));
    const int every = std::max(1, check_every);
    long long kk = 0;
    auto batch = [&](long long kk0) -> int {
      for (int k = 0; k < every; ++k) {
        DPP_CHECK(cg_fused_apply(ctx, op.nf, coef, kk0 + k, fld, slot, dtab));
        DPP_CHECK(cg_fused_r_update(ctx, op.nf, coef, kk0 + k, fld, slot, dtab));
        DPP_CHECK(cg_fused_halo_r(ctx, op.nf, true, slot));
      }
      return DPP_OK;
    };
    // the p ping-pong repeats with period 2: batches of an even number of iterations are identical launch
    // sequences -> replayed as one CUDA graph after the first (directly launched) batch
    unsigned long long key = mix(mix(mix(ctx->state_gen, (unsigned long long)op.nf * 16 + op.row * 4 + pc.type),
                                     (unsigned long long)every * 4 + cg_fused_variant(ctx)),
                                 (unsigned long long)(uintptr_t)hist_device(ctx, slot));
    // Polling without a pipeline bubble: batch k+1 is enqueued BEFORE the host waits for the scalars of
    // batch k (copied into a pinned double buffer behind each batch).  Kernels launched past convergence
    // are no-ops on every rank alike (all ranks hold bit-identical scalars), so the speculative batch
    // changes nothing; it costs a few microseconds once per solve instead of a host round trip per batch.
    Krylov* K = ctx->krylov;
    if (!K->h_poll) {
      DPP_CUDA(cudaMallocHost((void**)&K->h_poll, sizeof(double) * 2 * S_SLOT_SIZE));
      for (auto& e : K->ev_poll) DPP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    auto enqueue = [&](int b) -> int {
      // the launch sequence of a batch repeats when `every` is a multiple of the direction-ring length (16; the
      // classic two-buffer ping-pong needs an even count)
      if (kk == 0 || (every % 16) != 0) DPP_CHECK(batch(kk));
      else DPP_CHECK(run_graphed(ctx, K->cg_graph[slot], key, [&]() { return batch(kk); }));
      kk += every;
      DPP_CUDA(cudaMemcpyAsync(K->h_poll + (size_t)b * S_SLOT_SIZE, S, sizeof(double) * S_SLOT_SIZE, cudaMemcpyDeviceToHost,
                               ctx->stream));
      DPP_CUDA(cudaEventRecord(K->ev_poll[b], ctx->stream));
      return DPP_OK;
    };
    if (h[S_REASON] == 0.0) {
      int cur = 0;
      DPP_CHECK(enqueue(cur));
      K->apply_count += every;
      while (true) {
        DPP_CHECK(enqueue(cur ^ 1));   // speculative
        DPP_CUDA(cudaEventSynchronize(K->ev_poll[cur]));
        if (K->h_poll[(size_t)cur * S_SLOT_SIZE + S_REASON] != 0.0) break;
        K->apply_count += every;       // the batch just enqueued was needed
        cur ^= 1;
      }
      DPP_CHECK(scalars_fetch(ctx, slot));   // final state (also drains the speculative no-op batch)
    }
    DPP_CHECK(cg_fused_x_finalize(ctx, op.nf, (long long)h[S_ITS], slot, x));
    out->its = (int)h[S_ITS];
    out->reason = (int)h[S_REASON];
    out->rnorm = h[S_RNORM];
    const int nh = std::min(out->its + 1, std::min(hist_cap, ctx->hist_cap[slot]));
    out->hist.resize(std::max(nh, 0));
    if (nh > 0) {
      DPP_CUDA(cudaMemcpyAsync(out->hist.data(), ctx->d_hist[slot], sizeof(double) * nh, cudaMemcpyDeviceToHost, ctx->stream));
      DPP_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return DPP_OK;
  }
  if (fused) {
    DPP_CHECK(vec_pointwise_mult(ctx, L, dinv, wk.r, wk.z));
  } else {
    DPP_CHECK(pc_apply(ctx, pc, wk.r, wk.z));
  }
  DPP_CHECK(vec_dot2(ctx, L, wk.r, wk.z, wk.z, wk.z, slot, POST_CG_INIT));
  DPP_CHECK(scalars_fetch(ctx, slot));
  const int every = fused ? std::max(1, check_every) : 1;
  while (h[S_REASON] == 0.0) {
    for (int k = 0; k < every; ++k) {
      DPP_CHECK(cg_p_update(ctx, L, wk.p, wk.r, dinv, fused ? nullptr : wk.z, slot));
      DPP_CHECK(halo(ctx, wk.p, op.nf));
      int nb = 0;
      DPP_CHECK(apply_spec(ctx, op, wk.p, wk.w, true, S + S_REASON, &nb));
      DPP_CHECK(reduce_partials(ctx, nb, 1, slot, POST_CG_PAP));
      DPP_CHECK(cg_xr_update(ctx, L, x, wk.r, wk.p, wk.w, dinv, fused, slot, POST_CG_RZ));
      if (!fused) {
        DPP_CHECK(pc_apply(ctx, pc, wk.r, wk.z));
        DPP_CHECK(vec_dot2(ctx, L, wk.r, wk.z, wk.z, wk.z, slot, POST_CG_RZ));
      }
    }
    DPP_CHECK(scalars_fetch(ctx, slot));
  }
  out->its = (int)h[S_ITS];
  out->reason = (int)h[S_REASON];
  out->rnorm = h[S_RNORM];
  const int nh = std::min(out->its + 1, std::min(hist_cap, ctx->hist_cap[slot]));
  out->hist.resize(std::max(nh, 0));
  if (nh > 0) {
    DPP_CUDA(cudaMemcpyAsync(out->hist.data(), ctx->d_hist[slot], sizeof(double) * nh, cudaMemcpyDeviceToHost, ctx->stream));
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return DPP_OK;
}

// ------------------------------------------------------------------------------------------------
// Preconditioners
// ------------------------------------------------------------------------------------------------
int solve_block(dpp_context* ctx, const dpp_options* opt, int f, const double* rhs, double* out) {
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  const VecLayout L1 = layout(ctx, 1);
  const double* dinv_f = (opt->inner_pc_type == DPP_PC_JACOBI) ? K->dinv + f * n : nullptr;
  if (opt->inner_ksp_type == DPP_INNER_PREONLY) return vec_pointwise_mult(ctx, L1, dinv_f, rhs, out);
  OpSpec op{1, f, f, DPP_OP_MATRIX_FREE};
  Pc pc;
  pc.type = dinv_f ? DPP_PC_JACOBI : DPP_PC_NONE;
  pc.dinv = dinv_f;
  Tol tol{opt->inner_rtol, opt->inner_atol, opt->dtol, opt->inner_max_it};
  KspOut o;
  CgWork wk{K->ri, K->pi, K->wi, K->zi};
  DPP_CHECK(cg_run(ctx, op, pc, rhs, out, wk, tol, 1, opt->check_every, 0, &o));
  K->inner_its += o.its;
  return DPP_OK;
}

// out = rhs_row - A[row][col] xcol     (off-diagonal block: P_row (-beta/mu M) P_col)
int coupled_rhs(dpp_context* ctx, int row, int col, const double* rhs_row, double* xcol, double* out, double* tmp) {
  OpSpec op{1, row, col, DPP_OP_MATRIX_FREE};
  DPP_CHECK(halo(ctx, xcol, 1));
  int nb = 0;
  DPP_CHECK(apply_spec(ctx, op, xcol, tmp, false, nullptr, &nb));
  k_sub<<<ew_blocks(ctx, ctx->n_nodes), 256, 0, ctx->stream>>>(ctx->n_nodes, rhs_row, tmp, out);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

int pc_apply(dpp_context* ctx, Pc& pc, const double* r, double* z) {
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  const VecLayout L2 = layout(ctx, 2);
  switch (pc.type) {
    case DPP_PC_NONE:
      return vec_copy(ctx, z, r, 2 * n);
    case DPP_PC_JACOBI:
      return vec_pointwise_mult(ctx, L2, pc.dinv, r, z);
    case DPP_PC_PBJACOBI:
      return vec_pbjacobi(ctx, L2, K->pb, K->pb + n, K->pb + 2 * n, r, z);
    case DPP_PC_FIELDSPLIT: {
      // y0 = A00^-1 r0 ; y1 = A11^-1 (r1 - A10 y0)  [multiplicative]  (solvers/parameters.py:30-57)
      DPP_CHECK(solve_block(ctx, pc.opt, 0, r, z));
      const double* rhs1 = r + n;
      if (pc.opt->fieldsplit_type == DPP_FS_MULTIPLICATIVE) {
        DPP_CHECK(coupled_rhs(ctx, 1, 0, r + n, z, K->bi, K->xi));
        rhs1 = K->bi;
      }
      return solve_block(ctx, pc.opt, 1, rhs1, z + n);
    }
  }
  ctx->set_error("unknown pc_type");
  return DPP_ERR_INVALID;
}

int pc_setup(dpp_context* ctx, const dpp_options* opt, Pc* pc) {
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  pc->type = opt->pc_type;
  pc->opt = opt;
  const bool need_diag = opt->pc_type == DPP_PC_JACOBI || opt->pc_type == DPP_PC_PBJACOBI ||
                         ((opt->pc_type == DPP_PC_FIELDSPLIT || opt->ksp_type == DPP_KSP_PICARD) &&
                          opt->inner_pc_type == DPP_PC_JACOBI);
  if (need_diag) {
    if (!ctx->diag_valid) {
      DPP_CHECK(op_diagonal(ctx));
      k_reciprocal<<<ew_blocks(ctx, 2 * n), 256, 0, ctx->stream>>>(2 * n, ctx->d_diag, K->dinv);
      ctx->launches++;
      DPP_CUDA(cudaGetLastError());
      ctx->diag_valid = true;
      K->pb_valid = false;
    }
    pc->dinv = K->dinv;
  }
  if (opt->pc_type == DPP_PC_PBJACOBI && !K->pb_valid) {
    if (!K->pb) DPP_CHECK(dev_alloc(ctx, &K->pb, 3 * n));
    // nodal mass diagonal (no masking) into K->t
    Coef cm{};
    cm.cM[0][0] = 1.0;
    cm.cM[1][1] = 1.0;
    const uint8_t* save = ctx->d_mask;
    ctx->d_mask = nullptr;
    int rc = (ctx->family == DPP_KERNEL_STRUCTURED) ? structured_diagonal(ctx, cm, K->t) : general_diagonal(ctx, cm, K->t);
    ctx->d_mask = const_cast<uint8_t*>(save);
    DPP_CHECK(rc);
    k_pb_blocks<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(n, ctx->d_diag, K->t, ctx->d_mask, -ctx->beta / ctx->mu, K->pb);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
    K->pb_valid = true;
  }
  return DPP_OK;
}

// ------------------------------------------------------------------------------------------------
// KSPGMRES(m): left preconditioning, classical Gram-Schmidt, Givens residual recurrence
// ------------------------------------------------------------------------------------------------
struct DefaultTest {
  double rtol, atol, dtol, rnorm0 = 0, ttol = 0;
  int operator()(int its, double rnorm) {
    if (its == 0) {
      rnorm0 = rnorm;
      ttol = std::max(rtol * rnorm, atol);
    }
    if (!std::isfinite(rnorm)) return DPP_DIVERGED_NANORINF;
    if (rnorm <= ttol) return rnorm < atol ? DPP_CONVERGED_ATOL : DPP_CONVERGED_RTOL;
    if (rnorm >= dtol * rnorm0) return DPP_DIVERGED_DTOL;
    return 0;
  }
};

int norm2_host(dpp_context* ctx, const VecLayout& L, const double* v, int slot, double* out) {
  DPP_CHECK(vec_dot2(ctx, L, v, v, nullptr, nullptr, slot, POST_NONE));
  DPP_CHECK(scalars_fetch(ctx, slot));
  *out = std::sqrt(ctx->h_scalars[(size_t)slot * S_SLOT_SIZE + S_TMP]);
  return DPP_OK;
}

int gmres_run(dpp_context* ctx, const OpSpec& op, Pc& pc, const double* b, double* x, const Tol& tol, int restart,
              KspOut* out) {
  Krylov* K = ctx->krylov;
  const VecLayout L = layout(ctx, 2);
  const int64_t len = 2 * ctx->n_nodes;
  const int slot = 0;
  restart = std::max(1, std::min(restart, kMaxGmresRestart));
  while ((int)K->V.size() < restart + 1) {
    double* v = nullptr;
    DPP_CHECK(dev_alloc(ctx, &v, len));
    DPP_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * len, ctx->stream));
    K->V.push_back(v);
  }
  DefaultTest test{tol.rtol, tol.atol, tol.dtol};
  const double* h = ctx->h_scalars + (size_t)slot * S_SLOT_SIZE;
  DPP_CHECK(vec_zero(ctx, x, len));
  int its = 0, reason = 0;
  double res = 0.0;
  bool first = true;
  std::vector<double> H((size_t)(restart + 2) * (restart + 1)), cc(restart + 1), ss(restart + 1), grs(restart + 2);
  auto Hm = [&](int r, int c) -> double& { return H[(size_t)r * (restart + 1) + c]; };
  const bool pc_none = pc.type == DPP_PC_NONE;
  while (true) {
    // V0 = M^-1 (b - A x)
    if (first) {
      if (pc_none) DPP_CHECK(vec_copy(ctx, K->V[0], b, len));
      else DPP_CHECK(pc_apply(ctx, pc, b, K->V[0]));
    } else {
      DPP_CHECK(halo(ctx, x, 2));
      int nb = 0;
      DPP_CHECK(apply_spec(ctx, op, x, K->w, false, nullptr, &nb));
      DPP_CHECK(vec_axpby(ctx, L, 1.0, b, -1.0, K->w));  // w = b - A x
      if (pc_none) DPP_CHECK(vec_copy(ctx, K->V[0], K->w, len));
      else DPP_CHECK(pc_apply(ctx, pc, K->w, K->V[0]));
    }
    first = false;
    DPP_CHECK(norm2_host(ctx, L, K->V[0], slot, &res));
    out->hist.push_back(res);
    if (res == 0.0) { reason = DPP_CONVERGED_ATOL; break; }
    reason = test(its, res);
    if (reason) break;
    if (its >= tol.max_it) { reason = DPP_DIVERGED_ITS; break; }
    DPP_CHECK(vec_scale_into(ctx, L, 1.0 / res, K->V[0], K->V[0]));
    std::fill(H.begin(), H.end(), 0.0);
    grs[0] = res;
    int it = 0;
    while (!reason && it < restart && its < tol.max_it) {
      if (it) out->hist.push_back(res);
      double* vnew = K->V[it + 1];
      DPP_CHECK(halo(ctx, K->V[it], 2));
      int nb = 0;
      if (pc_none) {
        DPP_CHECK(apply_spec(ctx, op, K->V[it], vnew, false, nullptr, &nb));
      } else {
        DPP_CHECK(apply_spec(ctx, op, K->V[it], K->w, false, nullptr, &nb));
        DPP_CHECK(pc_apply(ctx, pc, K->w, vnew));
      }
      DPP_CHECK(gmres_mdot(ctx, L, K->V.data(), it + 1, vnew, slot));
      DPP_CHECK(gmres_maxpy_norm(ctx, L, K->V.data(), it + 1, vnew, slot));
      DPP_CHECK(scalars_fetch(ctx, slot));
      const double tt = std::sqrt(h[S_TMP + kGmresNormOffset]);
      bool hapend = false;
      const double hapbnd = std::min(std::fabs(tt / grs[it]), 1e-30);
      if (tt < hapbnd) hapend = true;
      else DPP_CHECK(vec_scale_into(ctx, L, 1.0 / tt, vnew, vnew));
      for (int j = 0; j <= it; ++j) Hm(j, it) = h[S_TMP + j];
      Hm(it + 1, it) = tt;
      for (int j = 0; j < it; ++j) {
        const double t = Hm(j, it);
        Hm(j, it) = cc[j] * t + ss[j] * Hm(j + 1, it);
        Hm(j + 1, it) = cc[j] * Hm(j + 1, it) - ss[j] * t;
      }
      if (!hapend) {
        const double t = std::sqrt(Hm(it, it) * Hm(it, it) + Hm(it + 1, it) * Hm(it + 1, it));
        if (t == 0.0) { reason = DPP_DIVERGED_BREAKDOWN; break; }
        cc[it] = Hm(it, it) / t;
        ss[it] = Hm(it + 1, it) / t;
        grs[it + 1] = -ss[it] * grs[it];
        grs[it] = cc[it] * grs[it];
        Hm(it, it) = cc[it] * Hm(it, it) + ss[it] * Hm(it + 1, it);
        res = std::fabs(grs[it + 1]);
      } else {
        res = 0.0;
      }
      ++it;
      ++its;
      reason = test(its, res);
      if (hapend && !reason) reason = DPP_DIVERGED_BREAKDOWN;
    }
    if (it && (reason || its >= tol.max_it)) out->hist.push_back(res);
    if (it) {
      std::vector<double> y(it);
      for (int k = it - 1; k >= 0; --k) {
        double s = grs[k];
        for (int j = k + 1; j < it; ++j) s -= Hm(k, j) * y[j];
        y[k] = s / Hm(k, k);
      }
      DPP_CHECK(vec_maxpy_host(ctx, L, K->V.data(), it, y.data(), x));
      DPP_CUDA(cudaStreamSynchronize(ctx->stream));  // y is a stack-lifetime host buffer
    }
    if (reason) break;
    if (its >= tol.max_it) { reason = DPP_DIVERGED_ITS; break; }
  }
  out->its = its;
  out->reason = reason;
  out->rnorm = res;
  return DPP_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-resident KSPGMRES(m): the Hessenberg / Givens recurrence, PETSc's convergence test and the
// history live in a device state block updated by single-thread kernels after each reduction; the host
// launches a whole restart cycle (kernels past convergence are no-ops) and synchronises once per cycle
// instead of once per iteration.  Same arithmetic, in the same order, as gmres_run below (the host-driven
// version, kept for preconditioners that synchronise themselves: fieldsplit with inner Krylov solves).
// ------------------------------------------------------------------------------------------------
enum { G_IT = 0, G_ITS, G_RES, G_REASON, G_RNORM0, G_TTOL, G_RTOL, G_ATOL, G_DTOL, G_MAXIT, G_INV, G_HAPEND, G_HISTCAP,
       G_M, G_NV, G_CC = 16, G_SS = 48, G_GRS = 80, G_Y = 114, G_H = 160, G_SIZE = 160 + 32 * 31 };

__device__ __forceinline__ int gmres_test(double* G, int its, double rnorm) {  // KSPConvergedDefault
  if (its == 0) {
    G[G_RNORM0] = rnorm;
    const double t = G[G_RTOL] * rnorm;
    G[G_TTOL] = t > G[G_ATOL] ? t : G[G_ATOL];
  }
  if (!(rnorm == rnorm) || isinf(rnorm)) return DPP_DIVERGED_NANORINF;
  if (rnorm <= G[G_TTOL]) return rnorm < G[G_ATOL] ? DPP_CONVERGED_ATOL : DPP_CONVERGED_RTOL;
  if (rnorm >= G[G_DTOL] * G[G_RNORM0]) return DPP_DIVERGED_DTOL;
  return 0;
}

// start of a restart cycle: S[S_TMP] = ||V0||^2 of the (preconditioned) residual
__global__ void k_gmres_cycle_start(double* G, const double* S, double* hist) {
  if (G[G_REASON] != 0.0) return;
  const int its = (int)G[G_ITS], m = (int)G[G_M];
  const double res = sqrt(S[S_TMP]);
  if (hist != nullptr && its < (int)G[G_HISTCAP]) hist[its] = res;
  int reason = 0;
  if (res == 0.0) reason = DPP_CONVERGED_ATOL;
  else {
    reason = gmres_test(G, its, res);
    if (!reason && its >= (int)G[G_MAXIT]) reason = DPP_DIVERGED_ITS;
  }
  G[G_RES] = res;
  G[G_REASON] = (double)reason;
  G[G_INV] = res != 0.0 ? 1.0 / res : 0.0;
  G[G_HAPEND] = 0.0;
  G[G_IT] = 0.0;
  G[G_GRS] = res;
  for (int q = 0; q < (m + 2) * (m + 1); ++q) G[G_H + q] = 0.0;
}

// after the Gram-Schmidt step of cycle iteration `it`: h = S[S_TMP .. S_TMP+it], ||w||^2 = S[S_TMP+36]
__global__ void k_gmres_post(double* G, const double* S, double* hist) {
  if (G[G_REASON] != 0.0) return;
  const int it = (int)G[G_IT], m = (int)G[G_M];
  const int ld = m + 1;
  double* H = G + G_H;
  double* cc = G + G_CC;
  double* ss = G + G_SS;
  double* grs = G + G_GRS;
  const double tt = sqrt(S[S_TMP + kGmresNormOffset]);
  bool hapend = false;
  const double hapbnd = fmin(fabs(tt / grs[it]), 1e-30);
  if (tt < hapbnd) hapend = true;
  else G[G_INV] = 1.0 / tt;
  G[G_HAPEND] = hapend ? 1.0 : 0.0;
  for (int j = 0; j <= it; ++j) H[j * ld + it] = S[S_TMP + j];
  H[(it + 1) * ld + it] = tt;
  for (int j = 0; j < it; ++j) {
    const double t = H[j * ld + it];
    H[j * ld + it] = cc[j] * t + ss[j] * H[(j + 1) * ld + it];
    H[(j + 1) * ld + it] = cc[j] * H[(j + 1) * ld + it] - ss[j] * t;
  }
  double res;
  if (!hapend) {
    const double a = H[it * ld + it], b = H[(it + 1) * ld + it];
    const double t = sqrt(a * a + b * b);
    if (t == 0.0) {
      G[G_REASON] = (double)DPP_DIVERGED_BREAKDOWN;
      return;
    }
    cc[it] = a / t;
    ss[it] = b / t;
    grs[it + 1] = -ss[it] * grs[it];
    grs[it] = cc[it] * grs[it];
    H[it * ld + it] = cc[it] * a + ss[it] * b;
    res = fabs(grs[it + 1]);
  } else {
    res = 0.0;
  }
  const int its = (int)G[G_ITS] + 1;
  int reason = gmres_test(G, its, res);
  if (hapend && !reason) reason = DPP_DIVERGED_BREAKDOWN;
  if (!reason && its >= (int)G[G_MAXIT]) reason = DPP_DIVERGED_ITS;  // the cycle stops here (ksp_max_it)
  G[G_IT] = (double)(it + 1);
  G[G_ITS] = (double)its;
  G[G_RES] = res;
  G[G_REASON] = (double)reason;
  if (hist != nullptr && its < (int)G[G_HISTCAP]) hist[its] = res;
}

// end of a cycle: y = H^-1 g for the `it` completed iterations (runs whatever the reason)
__global__ void k_gmres_solve_y(double* G) {
  const int it = (int)G[G_IT], m = (int)G[G_M];
  const int ld = m + 1;
  const double* H = G + G_H;
  const double* grs = G + G_GRS;
  double* y = G + G_Y;
  for (int k = it - 1; k >= 0; --k) {
    double s = grs[k];
    for (int j = k + 1; j < it; ++j) s -= H[k * ld + j] * y[j];
    y[k] = s / H[k * ld + k];
  }
  G[G_NV] = (double)it;
}

int gmres_run_device(dpp_context* ctx, const OpSpec& op, Pc& pc, const double* b, double* x, const Tol& tol, int restart,
                     int hist_cap, KspOut* out) {
  Krylov* K = ctx->krylov;
  const VecLayout L = layout(ctx, 2);
  const int64_t len = 2 * ctx->n_nodes;
  const int slot = 0;
  restart = std::max(1, std::min(restart, kMaxGmresRestart));
  while ((int)K->V.size() < restart + 1) {
    double* v = nullptr;
    DPP_CHECK(dev_alloc(ctx, &v, len));
    DPP_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * len, ctx->stream));
    K->V.push_back(v);
  }
  if (!K->d_gm) {
    DPP_CHECK(dev_alloc(ctx, &K->d_gm, G_SIZE));
    DPP_CUDA(cudaMallocHost((void**)&K->h_gm, sizeof(double) * 16));
  }
  if (hist_cap > ctx->hist_cap[slot]) {
    if (ctx->d_hist[slot]) cudaFree(ctx->d_hist[slot]);
    ctx->d_hist[slot] = nullptr;
    DPP_CHECK(dev_alloc(ctx, &ctx->d_hist[slot], hist_cap));
    ctx->hist_cap[slot] = hist_cap;
  }
  double* hist = hist_cap > 0 ? ctx->d_hist[slot] : nullptr;
  double* G = K->d_gm;
  const double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  std::vector<double> g0(G_SIZE, 0.0);
  g0[G_RTOL] = tol.rtol; g0[G_ATOL] = tol.atol; g0[G_DTOL] = tol.dtol; g0[G_MAXIT] = (double)tol.max_it;
  g0[G_HISTCAP] = (double)std::min(hist_cap, ctx->hist_cap[slot]); g0[G_M] = (double)restart;
  DPP_CUDA(cudaMemcpyAsync(G, g0.data(), sizeof(double) * G_SIZE, cudaMemcpyHostToDevice, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));  // g0 is a stack-lifetime host buffer
  DPP_CHECK(vec_zero(ctx, x, len));
  const bool pc_none = pc.type == DPP_PC_NONE;
  const double* skip = G + G_REASON;
  bool first = true;
  int its = 0, reason = 0, n_cycles = 0;
  double res = 0.0;
  while (true) {
    // V0 = M^-1 (b - A x)
    if (first) {
      if (pc_none) DPP_CHECK(vec_copy(ctx, K->V[0], b, len));
      else DPP_CHECK(pc_apply(ctx, pc, b, K->V[0]));
    } else {
      DPP_CHECK(halo(ctx, x, 2));
      int nb = 0;
      DPP_CHECK(apply_spec(ctx, op, x, K->w, false, nullptr, &nb));
      DPP_CHECK(vec_axpby(ctx, L, 1.0, b, -1.0, K->w));  // w = b - A x
      if (pc_none) DPP_CHECK(vec_copy(ctx, K->V[0], K->w, len));
      else DPP_CHECK(pc_apply(ctx, pc, K->w, K->V[0]));
    }
    first = false;
    DPP_CHECK(vec_dot2(ctx, L, K->V[0], K->V[0], nullptr, nullptr, slot, POST_NONE));
    k_gmres_cycle_start<<<1, 1, 0, ctx->stream>>>(G, S, hist);
    ctx->launches++;
    auto cycle = [&]() -> int {
      DPP_CHECK(vec_scale_dev(ctx, L, K->V[0], G + G_INV, skip, nullptr));
      for (int it = 0; it < restart; ++it) {
        double* vnew = K->V[it + 1];
        DPP_CHECK(halo(ctx, K->V[it], 2));
        int nb = 0;
        if (pc_none) {
          DPP_CHECK(apply_spec(ctx, op, K->V[it], vnew, false, skip, &nb));
        } else {
          DPP_CHECK(apply_spec(ctx, op, K->V[it], K->w, false, skip, &nb));
          DPP_CHECK(pc_apply(ctx, pc, K->w, vnew));  // pointwise preconditioners only: harmless past convergence
        }
        DPP_CHECK(gmres_mdot(ctx, L, K->V.data(), it + 1, vnew, slot, skip));
        DPP_CHECK(gmres_maxpy_norm(ctx, L, K->V.data(), it + 1, vnew, slot, skip));
        k_gmres_post<<<1, 1, 0, ctx->stream>>>(G, S, hist);
        ctx->launches++;
        DPP_CHECK(vec_scale_dev(ctx, L, vnew, G + G_INV, skip, G + G_HAPEND));
      }
      k_gmres_solve_y<<<1, 1, 0, ctx->stream>>>(G);
      ctx->launches++;
      DPP_CHECK(vec_maxpy_dev(ctx, L, K->V.data(), restart, G + G_Y, G + G_NV, x));
      return DPP_OK;
    };
    // every restart cycle is the same launch sequence (state and scalars live on the device): one CUDA graph
    const unsigned long long key = mix(mix(mix(ctx->state_gen, (unsigned long long)restart * 64 + pc.type * 4 + op.mode),
                                           (unsigned long long)(uintptr_t)hist), (unsigned long long)(uintptr_t)x);
    if (n_cycles == 0) DPP_CHECK(cycle());   // first cycle directly: lazy allocations / attributes happen here
    else DPP_CHECK(run_graphed(ctx, K->gmres_graph, key, cycle));
    ++n_cycles;
    DPP_CUDA(cudaMemcpyAsync(K->h_gm, G, sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
    its = (int)K->h_gm[G_ITS];
    reason = (int)K->h_gm[G_REASON];
    res = K->h_gm[G_RES];
    if (reason) break;
    if (its >= tol.max_it) { reason = DPP_DIVERGED_ITS; break; }
  }
  out->its = its;
  out->reason = reason;
  out->rnorm = res;
  const int nh = std::min(its + 1, std::min(hist_cap, ctx->hist_cap[slot]));
  out->hist.resize(std::max(nh, 0));
  if (nh > 0) {
    DPP_CUDA(cudaMemcpyAsync(out->hist.data(), ctx->d_hist[slot], sizeof(double) * nh, cudaMemcpyDeviceToHost, ctx->stream));
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return DPP_OK;
}

// ------------------------------------------------------------------------------------------------
// Block Picard (scale splitting), SURVEY A.6
// ------------------------------------------------------------------------------------------------
int picard_run(dpp_context* ctx, const dpp_options* opt, const double* b, double* d, const Tol& tol, double bnorm,
               KspOut* out) {
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  const VecLayout L2 = layout(ctx, 2);
  DPP_CHECK(vec_zero(ctx, d, 2 * n));
  const double ttol = std::max(tol.rtol * bnorm, tol.atol);
  double fnorm = bnorm;
  int its = 0;
  out->hist.push_back(fnorm);
  OpSpec mono{2, 0, 0, DPP_OP_MATRIX_FREE};
  while (fnorm > ttol && its < tol.max_it && std::isfinite(fnorm)) {
    // (k1K+bM) d0 = b0 - A01 d1
    DPP_CHECK(coupled_rhs(ctx, 0, 1, b, d + n, K->bi, K->xi));
    DPP_CHECK(solve_block(ctx, opt, 0, K->bi, K->t));
    DPP_CHECK(vec_copy(ctx, d, K->t, n));
    // (k2K+bM) d1 = b1 - A10 d0
    DPP_CHECK(coupled_rhs(ctx, 1, 0, b + n, d, K->bi, K->xi));
    DPP_CHECK(solve_block(ctx, opt, 1, K->bi, K->t));
    DPP_CHECK(vec_copy(ctx, d + n, K->t, n));
    // monolithic residual
    DPP_CHECK(halo(ctx, d, 2));
    int nb = 0;
    DPP_CHECK(apply_spec(ctx, mono, d, K->w, false, nullptr, &nb));
    DPP_CHECK(vec_axpby(ctx, L2, 1.0, b, -1.0, K->w));
    DPP_CHECK(norm2_host(ctx, L2, K->w, 0, &fnorm));
    out->hist.push_back(fnorm);
    ++its;
  }
  out->its = its;
  out->rnorm = fnorm;
  out->reason = !std::isfinite(fnorm) ? DPP_DIVERGED_NANORINF
                                      : (fnorm <= ttol ? (fnorm < tol.atol ? DPP_CONVERGED_ATOL : DPP_CONVERGED_RTOL)
                                                       : DPP_DIVERGED_ITS);
  return DPP_OK;
}

int ensure_work(dpp_context* ctx) {
  if (ctx->krylov && ctx->krylov->nvec == 2 * ctx->n_nodes) return DPP_OK;
  if (!ctx->krylov) ctx->krylov = new Krylov();
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  K->nvec = 2 * n;
  double** two[] = {&K->b, &K->x, &K->r, &K->p, &K->w, &K->z, &K->t, &K->u0, &K->dinv};
  for (double** v : two) {
    DPP_CHECK(dev_alloc(ctx, v, 2 * n));
    DPP_CUDA(cudaMemsetAsync(*v, 0, sizeof(double) * 2 * n, ctx->stream));
  }
  double** one[] = {&K->bi, &K->xi, &K->ri, &K->pi, &K->wi, &K->zi};
  for (double** v : one) {
    DPP_CHECK(dev_alloc(ctx, v, n));
    DPP_CUDA(cudaMemsetAsync(*v, 0, sizeof(double) * n, ctx->stream));
  }
  for (auto& e : K->ev) DPP_CUDA(cudaEventCreate(&e));
  return DPP_OK;
}

}  // namespace

double* krylov_scratch_vector(dpp_context* ctx) { return ctx->krylov ? ctx->krylov->t : nullptr; }

int krylov_work_vectors(dpp_context* ctx, double** a, double** b) {
  DPP_CHECK(ensure_work(ctx));
  *a = ctx->krylov->p;
  *b = ctx->krylov->w;
  return DPP_OK;
}

__global__ void k_fill_work(long long n, double* __restrict__ x, const uint8_t* __restrict__ mask, unsigned long long salt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = ((unsigned long long)i + salt) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    x[i] = (mask != nullptr && mask[i]) ? 0.0 : (double)(z & 0xFFFFF) / 1048576.0 - 0.5;
  }
}

// ------------------------------------------------------------------------------------------------
// L2 projection solve M u = rhs on the scalar space (no boundary conditions): Jacobi-CG on the nodal mass
// matrix, the linear solve inside fd.project (utils/postprocessing.py:63).  rhs / out: device, [n_nodes].
// ------------------------------------------------------------------------------------------------
int krylov_mass_solve(dpp_context* ctx, const double* d_rhs, double* d_out, double rtol, int max_it, int* its,
                      double* rnorm, int* reason) {
  DPP_CHECK(ensure_work(ctx));
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  // nodal mass diagonal (no masking) -> K->t[0..n), reciprocal -> K->t[n..2n)
  Coef cm{};
  cm.cM[0][0] = 1.0;
  cm.cM[1][1] = 1.0;
  uint8_t* save = ctx->d_mask;
  ctx->d_mask = nullptr;
  const int rc = (ctx->family == DPP_KERNEL_STRUCTURED) ? structured_diagonal(ctx, cm, K->t) : general_diagonal(ctx, cm, K->t);
  ctx->d_mask = save;
  DPP_CHECK(rc);
  k_reciprocal<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(n, K->t, K->t + n);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  OpSpec op{1, 0, 0, DPP_OP_MATRIX_FREE, 1, 1};
  Pc pc;
  pc.type = DPP_PC_JACOBI;
  pc.dinv = K->t + n;
  Tol tol{rtol, 1e-300, 1e10, max_it};
  KspOut o;
  CgWork wk{K->ri, K->pi, K->wi, K->zi};
  DPP_CHECK(cg_run(ctx, op, pc, d_rhs, d_out, wk, tol, 1, 4, 0, &o));
  if (its) *its = o.its;
  if (rnorm) *rnorm = o.rnorm;
  if (reason) *reason = o.reason;
  return DPP_OK;
}

// ------------------------------------------------------------------------------------------------
// Lanczos tridiagonalisation of the symmetric operator A_bc (which = 0), or of its diagonal blocks
// A00 (1) / A11 (2), from a pseudo-random start vector that also has components on the eliminated rows
// (their eigenvalue 1 belongs to the spectrum the reference's dense SVD sees, solvers/conditioning.py:
// 105-218).  alpha[j] = <v_j, A v_j>, beta[j] = ||A v_j - alpha_j v_j - beta_{j-1} v_{j-1}||.  No
// re-orthogonalisation: the extreme Ritz values, which is all a condition number needs, converge regardless.
// ------------------------------------------------------------------------------------------------
int krylov_lanczos(dpp_context* ctx, int which, int steps, unsigned long long seed, double* alpha, double* beta, int* done) {
  if (!ctx->have_params) {
    ctx->set_error("dpp_lanczos: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  DPP_CHECK(ensure_work(ctx));
  Krylov* K = ctx->krylov;
  const int nf = which == 0 ? 2 : 1;
  const int fld = which == 2 ? 1 : 0;
  const int64_t n = ctx->n_nodes, len = nf * n;
  const VecLayout L = layout(ctx, nf);
  OpSpec op{nf, fld, fld, DPP_OP_MATRIX_FREE, 0, 0};
  double *v = K->r, *vp = K->p, *w = K->w;
  const double* h = ctx->h_scalars;   // slot 0
  k_fill_work<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(len, v, nullptr, 7919ull * (seed + 1));
  ctx->launches++;
  DPP_CHECK(vec_zero(ctx, vp, len));
  DPP_CHECK(vec_dot2(ctx, L, v, v, v, v, 0, POST_NONE));
  DPP_CHECK(scalars_fetch(ctx, 0));
  DPP_CHECK(vec_axpby(ctx, L, 0.0, v, 1.0 / std::sqrt(h[S_TMP]), v));
  double b_prev = 0.0;
  int j = 0;
  for (; j < steps; ++j) {
    int nb = 0;
    DPP_CHECK(apply_spec(ctx, op, v, w, false, nullptr, &nb));
    // (the kernels' fused <x, Ax> skips the identity rows, which are fixed up afterwards: separate dot)
    DPP_CHECK(vec_dot2(ctx, L, v, w, v, w, 0, POST_NONE));
    DPP_CHECK(scalars_fetch(ctx, 0));
    const double a = h[S_TMP];
    alpha[j] = a;
    DPP_CHECK(vec_axpby(ctx, L, -a, v, 1.0, w));
    if (j > 0) DPP_CHECK(vec_axpby(ctx, L, -b_prev, vp, 1.0, w));
    DPP_CHECK(vec_dot2(ctx, L, w, w, w, w, 0, POST_NONE));
    DPP_CHECK(scalars_fetch(ctx, 0));
    const double b = std::sqrt(h[S_TMP]);
    beta[j] = b;
    if (!(b > 1e-14 * std::fabs(a))) { ++j; break; }   // invariant subspace found
    DPP_CHECK(vec_axpby(ctx, L, 1.0 / b, w, 0.0, vp));   // vp <- next v
    std::swap(v, vp);
    b_prev = b;
  }
  *done = j;
  return DPP_OK;
}

int krylov_time_cg_kernels(dpp_context* ctx, int warmup, int reps, double* apply_ms, double* update_ms,
                           double* matvec_ms, int nf, int field) {
  if (!ctx->have_params) {
    ctx->set_error("dpp_time_cg_kernels: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  if (!cg_fused_available(ctx, nf, DPP_OP_MATRIX_FREE, DPP_PC_JACOBI)) {
    ctx->set_error("dpp_time_cg_kernels: the handle does not run the fused uniform-grid CG path");
    return DPP_ERR_INVALID;
  }
  DPP_CHECK(ensure_work(ctx));
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  const int slot = 0;
  double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  // pseudo-random work vectors (zero on constrained rows, like every Krylov vector): r, p0, p1, x
  const int which[4] = {0, 1, 2, 4};
  for (int v = 0; v < 4; ++v) {
    k_fill_work<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(2 * n, K->t, ctx->d_mask, 7919ull * (v + 1));
    ctx->launches++;
    DPP_CHECK(cg_fused_pad_from(ctx, which[v], K->t));
  }
  DPP_CHECK(scalars_init(ctx, slot, 1e-8, 1e-12, 1e4, 1 << 30, 0));
  double* hs = ctx->h_scalars + (size_t)slot * S_SLOT_SIZE;
  hs[S_ITS] = 1.0; hs[S_RZ] = 1.0; hs[S_RZ_OLD] = 2.0; hs[S_PAP] = 1.0; hs[S_ALPHA] = 1e-3; hs[S_XPEND] = 1.0;
  hs[S_RTOL] = 0.0; hs[S_ATOL] = 0.0; hs[S_DTOL] = 1e300; hs[S_MAXIT] = 1e18;
  hs[S_RNORM0] = 1e300; hs[S_TTOL] = 0.0;  // the folded convergence test never fires while timing
  DPP_CUDA(cudaMemcpyAsync(S, hs, sizeof(double) * S_SLOT_SIZE, cudaMemcpyHostToDevice, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  // nf = 2: the monolithic operator; nf = 1: the diagonal block of `field` (Picard / fieldsplit block solves)
  const Coef coef = nf == 2 ? dpp_coef(ctx) : block_coef(ctx, field, field);
  double* dtab = ctx->d_dtab;
  const int fld[2] = {nf == 2 ? 0 : field, 1};
  DPP_CHECK(cg_fused_table(ctx, coef, nf, DPP_PC_JACOBI, fld, dtab));
  cudaEvent_t e0, e1;
  DPP_CUDA(cudaEventCreate(&e0));
  DPP_CUDA(cudaEventCreate(&e1));
  float ms = 0;
  int nb = 0;
  double* outs[3] = {apply_ms, update_ms, matvec_ms};
  for (int pass = 0; pass < (nf == 2 ? 3 : 2); ++pass) {
    for (int i = -warmup; i < reps; ++i) {
      if (i == 0) DPP_CUDA(cudaEventRecord(e0, ctx->stream));
      if (pass == 0) DPP_CHECK(cg_fused_apply(ctx, nf, coef, i + warmup, fld, slot, dtab));
      else if (pass == 1) DPP_CHECK(cg_fused_r_update(ctx, nf, coef, i + warmup, fld, slot, dtab));
      else if (ctx->grid.band == 1) DPP_CHECK(cg_fused_plain_apply(ctx, 2, coef, true, &nb));
      else {   // degree 2: the stand-alone stencil kernel on the C-ABI layout (apply_structured_q2.cu)
        OpSpec op{};
        op.nf = 2; op.mode = DPP_OP_MATRIX_FREE; op.premasked = 1;
        DPP_CHECK(apply_spec(ctx, op, K->t, K->r, true, nullptr, &nb));
      }
    }
    DPP_CUDA(cudaEventRecord(e1, ctx->stream));
    DPP_CUDA(cudaEventSynchronize(e1));
    DPP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *outs[pass] = (double)ms / reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return DPP_OK;
}

int krylov_solve(dpp_context* ctx, const dpp_options* opt, double* u_host, dpp_result* res, double* hist_host,
                 int32_t hist_cap) {
  if (!ctx->have_params) {
    ctx->set_error("dpp_solve: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  if (opt->ksp_type == DPP_KSP_GMRES && (opt->gmres_restart < 1 || opt->gmres_restart > kMaxGmresRestart)) {
    // the device-resident Hessenberg / Givens state is sized for restart <= 30 (PETSc's default): refuse
    // rather than silently run GMRES(30) with different iteration counts than PETSc would report
    ctx->set_error("ksp_gmres_restart must be in [1, " + std::to_string(kMaxGmresRestart) + "]");
    return DPP_ERR_INVALID;
  }
  DPP_CHECK(ensure_work(ctx));
  Krylov* K = ctx->krylov;
  const int64_t n = ctx->n_nodes;
  const VecLayout L2 = layout(ctx, 2);
  K->inner_its = 0;
  K->apply_count = 0;
  if (opt->operator_mode == DPP_OP_ASSEMBLED && !csr_valid(ctx)) {
    int64_t nnz = 0;
    DPP_CHECK(csr_assemble(ctx, &nnz));
  }
  DPP_CUDA(cudaEventRecord(K->ev[0], ctx->stream));
  // lifting: u0 = g on Gamma; b = -(A u0) on interior rows, 0 on Gamma
  {
    OpArgs a{};
    a.nf = 2;
    a.c = dpp_coef(ctx);
    for (int f = 0; f < 2; ++f)
      for (int g = 0; g < 2; ++g) {
        a.c.cK[f][g] = -a.c.cK[f][g];
        a.c.cM[f][g] = -a.c.cM[f][g];
      }
    for (int f = 0; f < 2; ++f) {
      a.x[f] = ctx->d_g + f * n;
      a.y[f] = K->b + f * n;
      a.in_mask[f] = nullptr;
      a.out_mask[f] = ctx->d_mask + f * n;
    }
    a.identity_on_masked = 0;
    a.owned_begin = ctx->owned_begin;
    a.owned_end = ctx->owned_end;
    int nb = 0;
    DPP_CHECK(op_apply(ctx, a, &nb));
  }
  double bnorm = 0.0;
  DPP_CHECK(norm2_host(ctx, L2, K->b, 0, &bnorm));
  Pc pc;
  DPP_CHECK(pc_setup(ctx, opt, &pc));
  DPP_CUDA(cudaEventRecord(K->ev[1], ctx->stream));

  Tol tol{opt->rtol, opt->atol, opt->dtol, opt->max_it};
  OpSpec mono{2, 0, 0, opt->operator_mode};
  KspOut out;
  int rc = DPP_OK;
  switch (opt->ksp_type) {
    case DPP_KSP_CG: {
      CgWork wk{K->r, K->p, K->w, K->z};
      rc = cg_run(ctx, mono, pc, K->b, K->x, wk, tol, 0, opt->check_every, std::max(hist_cap, 0), &out);
      break;
    }
    case DPP_KSP_GMRES:
      // pointwise preconditioners: whole restart cycles run from device state; fieldsplit (inner Krylov
      // solves that synchronise with the host anyway) keeps the host-driven recurrence
      if (opt->pc_type != DPP_PC_FIELDSPLIT && getenv("DPP_GMRES_HOST") == nullptr)
        rc = gmres_run_device(ctx, mono, pc, K->b, K->x, tol, opt->gmres_restart, std::max(hist_cap, 0), &out);
      else
        rc = gmres_run(ctx, mono, pc, K->b, K->x, tol, opt->gmres_restart, &out);
      break;
    case DPP_KSP_PICARD:
      rc = picard_run(ctx, opt, K->b, K->x, tol, bnorm, &out);
      break;
    default:
      ctx->set_error("unknown ksp_type");
      rc = DPP_ERR_INVALID;
  }
  DPP_CHECK(rc);
  if (ctx->world > 1 && comm_halo_failed(ctx)) {
    ctx->set_error("a neighbour rank never delivered its halo planes (peer-memory halo inbox timed out)");
    return DPP_ERR_NCCL;
  }
  if (out.reason == DPP_DIVERGED_COMM_TIMEOUT) {
    ctx->set_error("a peer rank never arrived at a Krylov reduction (peer-memory mailbox timed out)");
    return DPP_ERR_NCCL;
  }
  DPP_CUDA(cudaEventRecord(K->ev[2], ctx->stream));
  // u = u0 + d
  if (!ctx->d_solution) DPP_CHECK(dev_alloc(ctx, &ctx->d_solution, 2 * n));
  DPP_CHECK(vec_copy(ctx, ctx->d_solution, ctx->d_g, 2 * n));
  DPP_CHECK(vec_axpby(ctx, L2, 1.0, K->x, 1.0, ctx->d_solution));
  DPP_CHECK(halo(ctx, ctx->d_solution, 2));  // ghost planes of the returned vector are consistent
  if (u_host) {
    const double* src = ctx->d_solution;
    if (ctx->d_perm) {  // hand the solution back in the caller's numbering
      DPP_CHECK(perm_to_user(ctx, ctx->d_solution, K->t, 2));
      src = K->t;
    }
    DPP_CUDA(cudaMemcpyAsync(u_host, src, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms_setup = 0, ms_solve = 0;
  cudaEventElapsedTime(&ms_setup, K->ev[0], K->ev[1]);
  cudaEventElapsedTime(&ms_solve, K->ev[1], K->ev[2]);
  if (res) {
    res->iterations = out.its;
    res->converged_reason = out.reason;
    res->inner_iterations = (int32_t)K->inner_its;
    res->residual_norm = out.rnorm;
    res->rhs_norm = bnorm;
    res->solve_ms = ms_solve;
    res->setup_ms = ms_setup;
    res->apply_ms = 0.0;
    res->apply_count = K->apply_count;
    res->history_len = 0;
    if (hist_host && hist_cap > 0) {
      const int nh = (int)std::min<size_t>(out.hist.size(), (size_t)hist_cap);
      std::memcpy(hist_host, out.hist.data(), sizeof(double) * nh);
      res->history_len = nh;
    }
  }
  return DPP_OK;
}

void krylov_destroy(dpp_context* ctx) {
  Krylov* K = ctx->krylov;
  if (!K) return;
  double* vs[] = {K->b, K->x, K->r, K->p, K->w, K->z, K->t, K->u0, K->dinv, K->bi, K->xi, K->ri, K->pi, K->wi, K->zi, K->pb};
  for (double* v : vs)
    if (v) cudaFree(v);
  for (double* v : K->V) cudaFree(v);
  for (GraphSlot* gsl : {&K->cg_graph[0], &K->cg_graph[1], &K->gmres_graph})
    if (gsl->exec) cudaGraphExecDestroy(gsl->exec);
  if (K->d_gm) cudaFree(K->d_gm);
  if (K->h_gm) cudaFreeHost(K->h_gm);
  if (K->h_poll) cudaFreeHost(K->h_poll);
  for (auto& e : K->ev_poll)
    if (e) cudaEventDestroy(e);
  for (auto& e : K->ev)
    if (e) cudaEventDestroy(e);
  delete K;
  ctx->krylov = nullptr;
}

}  // namespace dpp
