// C ABI entry points (include/dpp_b200.h).  Thin: argument checks, uploads, dispatch.
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "dpp_internal.cuh"
#include "vector_ops.cuh"

namespace dpp {
int krylov_work_vectors(dpp_context* ctx, double** a, double** b);
double* krylov_scratch_vector(dpp_context* ctx);  // [2*n_nodes], valid after krylov_work_vectors / a solve
}

static std::string g_create_error;

namespace {

__global__ void k_fill_pseudo(long long n, double* __restrict__ x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    x[i] = (double)(z & 0xFFFFF) / 1048576.0 - 0.5;
  }
}

__global__ void k_zero_masked(long long n, const uint8_t* __restrict__ m, double* __restrict__ x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (m[i]) x[i] = 0.0;
}

__global__ void k_scatter_bc(long long n, const int32_t* __restrict__ nodes, const double* __restrict__ vals,
                             uint8_t* __restrict__ mask, double* __restrict__ g) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    mask[nodes[i]] = 1;
    g[nodes[i]] = vals[i];
  }
}

__global__ void k_map_ids(long long n, int32_t* __restrict__ ids, const int32_t* __restrict__ perm) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ids[i] = perm[ids[i]];
}

// TO_INTERNAL: internal[perm[u]] = user[u] ; else user[u] = internal[perm[u]]   (per field)
template <bool TO_INTERNAL>
__global__ void k_permute(long long n, int nf, const int32_t* __restrict__ perm, const double* __restrict__ src,
                          double* __restrict__ dst) {
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n; u += (long long)gridDim.x * blockDim.x) {
    const long long q = perm[u];
    for (int f = 0; f < nf; ++f) {
      if (TO_INTERNAL) dst[f * n + q] = src[f * n + u];
      else dst[f * n + u] = src[f * n + q];
    }
  }
}

int fail_create(dpp_context* ctx, int rc, const std::string& msg) {
  g_create_error = msg.empty() && ctx ? ctx->err : msg;
  if (ctx) dpp_destroy(ctx);
  return rc;
}

}  // namespace

namespace dpp {
int perm_to_internal(dpp_context* ctx, const double* user, double* internal, int nf) {
  const int blocks = (int)std::min<long long>((ctx->n_nodes + 255) / 256, (long long)ctx->sm_count * 16);
  k_permute<true><<<blocks, 256, 0, ctx->stream>>>(ctx->n_nodes, nf, ctx->d_perm, user, internal);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}
int perm_to_user(dpp_context* ctx, const double* internal, double* user, int nf) {
  const int blocks = (int)std::min<long long>((ctx->n_nodes + 255) / 256, (long long)ctx->sm_count * 16);
  k_permute<false><<<blocks, 256, 0, ctx->stream>>>(ctx->n_nodes, nf, ctx->d_perm, internal, user);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}
}  // namespace dpp

extern "C" {

int dpp_set_numbering(dpp_handle ctx, const int32_t* map) {
  if (!ctx || !map) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->world > 1) {
    ctx->set_error("dpp_set_numbering: not available on slab-partitioned handles");
    return DPP_ERR_INVALID;
  }
  std::vector<uint8_t> seen((size_t)ctx->n_nodes, 0);
  for (int64_t i = 0; i < ctx->n_nodes; ++i) {
    if (map[i] < 0 || map[i] >= ctx->n_nodes || seen[map[i]]) {
      ctx->set_error("dpp_set_numbering: the map is not a permutation of the nodes");
      return DPP_ERR_INVALID;
    }
    seen[map[i]] = 1;
  }
  if (!ctx->d_perm) DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_perm, ctx->n_nodes));
  DPP_CUDA(cudaMemcpy(ctx->d_perm, map, sizeof(int32_t) * ctx->n_nodes, cudaMemcpyHostToDevice));
  // Dirichlet data uploaded before the map would be in the wrong numbering: start clean
  for (int f = 0; f < 2; ++f) {
    DPP_CUDA(cudaMemsetAsync(ctx->d_mask + f * ctx->n_nodes, 0, ctx->n_nodes, ctx->stream));
    DPP_CUDA(cudaMemsetAsync(ctx->d_g + f * ctx->n_nodes, 0, sizeof(double) * ctx->n_nodes, ctx->stream));
    ctx->n_bc[f] = 0;
    ctx->bc_gen[f]++;
    ctx->have_bc[f] = false;
  }
  ctx->invalidate();
  dpp::csr_invalidate(ctx);
  return DPP_OK;
}

int dpp_create(dpp_handle* h, int device, int dim, int degree, int64_t n_nodes, int64_t n_cells, int nodes_per_cell,
               const int32_t* cnm, int64_t n_coord_nodes, const double* coords, const int32_t* ccnm) {
  if (!h) return DPP_ERR_INVALID;
  *h = nullptr;
  if ((dim != 2 && dim != 3) || (degree != 1 && degree != 2) || n_nodes <= 0 || n_cells <= 0 || !cnm || !coords || !ccnm)
    return fail_create(nullptr, DPP_ERR_INVALID, "dpp_create: invalid argument");
  int npc = 1;
  for (int d = 0; d < dim; ++d) npc *= degree + 1;
  if (nodes_per_cell != npc) return fail_create(nullptr, DPP_ERR_INVALID, "dpp_create: nodes_per_cell != (degree+1)^dim");
  if (n_nodes >= (1LL << 31) || n_cells >= (1LL << 31))
    return fail_create(nullptr, DPP_ERR_INVALID, "dpp_create: int32 node/cell ids");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
    return fail_create(nullptr, DPP_ERR_NO_DEVICE,
                       "dpp_create: no CUDA device available (libdppb200 has no CPU fallback)");
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    return fail_create(nullptr, DPP_ERR_NO_DEVICE, "dpp_create: device is not sm_100 or newer");
  dpp_context* ctx = new (std::nothrow) dpp_context();
  if (!ctx) return fail_create(nullptr, DPP_ERR_INVALID, "dpp_create: out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->dim = dim;
  ctx->degree = degree;
  ctx->npc = npc;
  ctx->nvc = 1 << dim;
  ctx->n_nodes = n_nodes;
  ctx->n_cells = n_cells;
  ctx->n_coord_nodes = n_coord_nodes;
  ctx->owned_begin = 0;
  ctx->owned_end = n_nodes;
  auto body = [&]() -> int {
    DPP_CUDA(cudaSetDevice(device));
    DPP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_cnm, n_cells * npc));
    DPP_CUDA(cudaMemcpyAsync(ctx->d_cnm, cnm, sizeof(int32_t) * n_cells * npc, cudaMemcpyHostToDevice, ctx->stream));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_coords, n_coord_nodes * dim));
    DPP_CUDA(cudaMemcpyAsync(ctx->d_coords, coords, sizeof(double) * n_coord_nodes * dim, cudaMemcpyHostToDevice, ctx->stream));
    if (degree == 1 && ccnm == cnm) {
      ctx->d_ccnm = ctx->d_cnm;
      ctx->ccnm_alias = true;
    } else {
      DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_ccnm, n_cells * ctx->nvc));
      DPP_CUDA(cudaMemcpyAsync(ctx->d_ccnm, ccnm, sizeof(int32_t) * n_cells * ctx->nvc, cudaMemcpyHostToDevice, ctx->stream));
    }
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_mask, 2 * n_nodes));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_g, 2 * n_nodes));
    DPP_CUDA(cudaMemsetAsync(ctx->d_mask, 0, 2 * n_nodes, ctx->stream));
    DPP_CUDA(cudaMemsetAsync(ctx->d_g, 0, sizeof(double) * 2 * n_nodes, ctx->stream));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_partials, (int64_t)dpp::kMaxPartialBlocks * dpp::kMaxDotWidth));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_scalars, dpp::kNumScalars));
    DPP_CUDA(cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * dpp::kNumScalars, ctx->stream));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_dtab, 2 * 128));
    DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_counters, 4));
    DPP_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned) * 4, ctx->stream));
    DPP_CUDA(cudaMallocHost((void**)&ctx->h_scalars, sizeof(double) * dpp::kNumScalars));
    std::memset(ctx->h_scalars, 0, sizeof(double) * dpp::kNumScalars);
    DPP_CHECK(dpp::structured_detect_and_setup(ctx, cnm, coords, ccnm));
    ctx->family = ctx->structured_ok ? DPP_KERNEL_STRUCTURED : DPP_KERNEL_GENERAL;
    if (ctx->family == DPP_KERNEL_GENERAL) DPP_CHECK(dpp::general_setup(ctx, cnm));
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
    return DPP_OK;
  };
  const int rc = body();
  if (rc != DPP_OK) return fail_create(ctx, rc, "");
  *h = ctx;
  return DPP_OK;
}

void dpp_destroy(dpp_handle ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  dpp::krylov_destroy(ctx);
  dpp::cg_fused_destroy(ctx);
  dpp::csr_destroy(ctx);
  dpp::cells_destroy(ctx);
  dpp::comm_destroy(ctx);
  void* ptrs[] = {ctx->d_cnm, ctx->d_coords, ctx->ccnm_alias ? nullptr : ctx->d_ccnm, ctx->d_tables, ctx->d_adj_ptr,
                  ctx->d_adj_cell, ctx->d_adj_loc, ctx->d_cell_geom, ctx->d_mask, ctx->d_g, ctx->d_solution, ctx->d_diag,
                  ctx->d_partials, ctx->d_scalars, ctx->d_dtab, ctx->d_counters, ctx->d_hist[0], ctx->d_hist[1], ctx->d_premask, ctx->d_bc_nodes[0], ctx->d_bc_nodes[1], ctx->d_bc_vals, ctx->d_perm};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
}

const char* dpp_last_error(dpp_handle ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dpp_get_info(dpp_handle ctx, dpp_info* info) {
  if (!ctx || !info) return DPP_ERR_INVALID;
  std::memset(info, 0, sizeof(*info));
  info->kernel_family = ctx->family;
  info->dim = ctx->dim;
  info->degree = ctx->degree;
  if (ctx->structured_ok)
    for (int a = 0; a < 3; ++a) info->grid_nodes[a] = ctx->grid.n[a];
  info->n_nodes = ctx->n_nodes;
  info->n_cells = ctx->n_cells;
  info->n_owned_nodes = ctx->owned_end - ctx->owned_begin;
  info->rank = ctx->rank;
  info->world = ctx->world;
  info->sm_count = ctx->sm_count;
  info->peer_memory = (dpp::comm_ipc_ready(ctx) ? 1 : 0) | (dpp::comm_ipc_halo_ready(ctx) ? 2 : 0) |
                      (dpp::comm_ipc_box_ready(ctx) ? 4 : 0);
  info->device_bytes = ctx->device_bytes;
  return DPP_OK;
}

int dpp_fused_cg_supported(dpp_handle ctx) {
  if (!ctx) return DPP_ERR_INVALID;
  return dpp::cg_fused_available(ctx, 2, DPP_OP_MATRIX_FREE, DPP_PC_JACOBI) ? 1 : 0;
}

int dpp_set_fused_cg(dpp_handle ctx, int enable) {
  if (!ctx) return DPP_ERR_INVALID;
  ctx->fused_cg_disabled = enable == 0;
  ctx->invalidate();
  return DPP_OK;
}

int dpp_force_kernel_family(dpp_handle ctx, int family) {
  if (!ctx) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (family == DPP_KERNEL_STRUCTURED) {
    if (!ctx->structured_ok) {
      ctx->set_error("structured kernel family unavailable for this mesh");
      return DPP_ERR_INVALID;
    }
    ctx->family = family;
  } else if (family == DPP_KERNEL_GENERAL) {
    if (!ctx->general_ready) {
      std::vector<int32_t> cnm((size_t)ctx->n_cells * ctx->npc);
      DPP_CUDA(cudaMemcpy(cnm.data(), ctx->d_cnm, sizeof(int32_t) * cnm.size(), cudaMemcpyDeviceToHost));
      DPP_CHECK(dpp::general_setup(ctx, cnm.data()));
    }
    ctx->family = family;
  } else {
    ctx->set_error("unknown kernel family");
    return DPP_ERR_INVALID;
  }
  ctx->invalidate();
  return DPP_OK;
}

int dpp_set_params(dpp_handle ctx, double k1, double k2, double beta, double mu) {
  if (!ctx) return DPP_ERR_INVALID;
  if (!(mu != 0.0)) {
    ctx->set_error("dpp_set_params: mu must be non-zero");
    return DPP_ERR_INVALID;
  }
  ctx->k1 = k1; ctx->k2 = k2; ctx->beta = beta; ctx->mu = mu;
  ctx->have_params = true;
  ctx->invalidate();
  dpp::csr_invalidate(ctx);
  return DPP_OK;
}

int dpp_set_dirichlet(dpp_handle ctx, int field, int64_t n, const int32_t* nodes, const double* values) {
  if (!ctx || (field != 0 && field != 1) || n < 0 || (n > 0 && (!nodes || !values))) {
    if (ctx) ctx->set_error("dpp_set_dirichlet: invalid argument");
    return DPP_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  for (int64_t i = 0; i < n; ++i)
    if (nodes[i] < 0 || nodes[i] >= ctx->n_nodes) {
      ctx->set_error("dpp_set_dirichlet: node id out of range");
      return DPP_ERR_INVALID;
    }
  const int64_t nn = ctx->n_nodes;
  DPP_CUDA(cudaMemsetAsync(ctx->d_mask + field * nn, 0, nn, ctx->stream));
  DPP_CUDA(cudaMemsetAsync(ctx->d_g + field * nn, 0, sizeof(double) * nn, ctx->stream));
  if (n > 0) {
    // device buffers are kept and only grow: no cudaMalloc/cudaFree on the per-solve path
    if (n > ctx->bc_cap[field]) {
      if (ctx->d_bc_nodes[field]) cudaFree(ctx->d_bc_nodes[field]);
      ctx->d_bc_nodes[field] = nullptr;
      ctx->bc_cap[field] = 0;
      DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_bc_nodes[field], n));
      ctx->bc_cap[field] = n;
    }
    if (n > ctx->bc_vals_cap) {
      if (ctx->d_bc_vals) cudaFree(ctx->d_bc_vals);
      ctx->d_bc_vals = nullptr;
      ctx->bc_vals_cap = 0;
      DPP_CHECK(dpp::dev_alloc(ctx, &ctx->d_bc_vals, n));
      ctx->bc_vals_cap = n;
    }
    int32_t* d_nodes = ctx->d_bc_nodes[field];  // kept: row-elimination fix-up of the uniform-grid apply
    DPP_CUDA(cudaMemcpyAsync(d_nodes, nodes, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    DPP_CUDA(cudaMemcpyAsync(ctx->d_bc_vals, values, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4096);
    if (ctx->d_perm) {  // caller numbering -> internal numbering
      k_map_ids<<<blocks, 256, 0, ctx->stream>>>(n, d_nodes, ctx->d_perm);
      ctx->launches++;
    }
    k_scatter_bc<<<blocks, 256, 0, ctx->stream>>>(n, d_nodes, ctx->d_bc_vals, ctx->d_mask + field * nn, ctx->d_g + field * nn);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller may reuse its host buffers
  }
  ctx->n_bc[field] = n;
  ctx->bc_gen[field]++;
  ctx->have_bc[field] = n > 0;
  ctx->invalidate();
  dpp::csr_invalidate(ctx);
  return DPP_OK;
}

static int apply_common(dpp_context* ctx, const double* x, double* y, int mode, bool want_dot, int* nb,
                        bool premasked = false) {
  if (!ctx->have_params) {
    ctx->set_error("dpp_apply: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  const int64_t n = ctx->n_nodes;
  if (mode == DPP_OP_ASSEMBLED) {
    if (!dpp::csr_valid(ctx)) {
      int64_t nnz = 0;
      DPP_CHECK(dpp::csr_assemble(ctx, &nnz));
    }
    return dpp::csr_spmv(ctx, x, y, want_dot ? ctx->d_partials : nullptr, nb, nullptr);
  }
  dpp::OpArgs a{};
  a.nf = 2;
  a.c = dpp::dpp_coef(ctx);
  for (int f = 0; f < 2; ++f) {
    a.x[f] = x + f * n;
    a.y[f] = y + f * n;
    a.in_mask[f] = a.out_mask[f] = ctx->d_mask + f * n;
  }
  a.identity_on_masked = 1;
  a.owned_begin = ctx->owned_begin;
  a.owned_end = ctx->owned_end;
  a.dot_partials = want_dot ? ctx->d_partials : nullptr;
  a.input_premasked = premasked ? 1 : 0;
  return dpp::op_apply(ctx, a, nb);
}

int dpp_apply_dev(dpp_handle ctx, const double* x, double* y, int mode) {
  if (!ctx || !x || !y) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int nb = 0;
  return apply_common(ctx, x, y, mode, false, &nb);
}

int dpp_apply_host(dpp_handle ctx, const double* x, double* y, int mode) {
  if (!ctx || !x || !y) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  double *dx = nullptr, *dy = nullptr;
  DPP_CHECK(dpp::krylov_work_vectors(ctx, &dx, &dy));
  const size_t bytes = sizeof(double) * 2 * ctx->n_nodes;
  double* dt = ctx->d_perm ? dpp::krylov_scratch_vector(ctx) : nullptr;
  if (ctx->d_perm) {
    DPP_CUDA(cudaMemcpyAsync(dt, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DPP_CHECK(dpp::perm_to_internal(ctx, dt, dx, 2));
  } else {
    DPP_CUDA(cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  DPP_CUDA(cudaMemsetAsync(dy, 0, bytes, ctx->stream));
  int nb = 0;
  DPP_CHECK(apply_common(ctx, dx, dy, mode, false, &nb));
  if (ctx->d_perm) {
    DPP_CHECK(dpp::perm_to_user(ctx, dy, dt, 2));
    dy = dt;
  }
  DPP_CUDA(cudaMemcpyAsync(y, dy, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DPP_OK;
}

int dpp_get_diagonal_host(dpp_handle ctx, double* diag) {
  if (!ctx || !diag) return DPP_ERR_INVALID;
  if (!ctx->have_params) {
    ctx->set_error("dpp_get_diagonal: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  cudaSetDevice(ctx->device);
  DPP_CHECK(dpp::op_diagonal(ctx));
  const double* src = ctx->d_diag;
  if (ctx->d_perm) {
    double *dx = nullptr, *dy = nullptr;
    DPP_CHECK(dpp::krylov_work_vectors(ctx, &dx, &dy));
    DPP_CHECK(dpp::perm_to_user(ctx, ctx->d_diag, dpp::krylov_scratch_vector(ctx), 2));
    src = dpp::krylov_scratch_vector(ctx);
  }
  DPP_CUDA(cudaMemcpyAsync(diag, src, sizeof(double) * 2 * ctx->n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->diag_valid = false;  // Krylov keeps its own reciprocal; recompute lazily
  return DPP_OK;
}

int dpp_assemble_csr(dpp_handle ctx, int64_t* nnz) {
  if (!ctx) return DPP_ERR_INVALID;
  if (!ctx->have_params) {
    ctx->set_error("dpp_assemble_csr: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  cudaSetDevice(ctx->device);
  return dpp::csr_assemble(ctx, nnz);
}

int dpp_get_csr_host(dpp_handle ctx, int64_t* indptr, int32_t* indices, double* data) {
  if (!ctx) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return dpp::csr_export(ctx, indptr, indices, data);
}

int dpp_get_csr_block_host(dpp_handle ctx, int row_field, int col_field, int64_t* indptr, int32_t* indices, double* data) {
  if (!ctx || row_field < 0 || row_field > 1 || col_field < 0 || col_field > 1) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return dpp::csr_export_block(ctx, row_field, col_field, indptr, indices, data);
}

int dpp_time_assembly(dpp_handle ctx, int reps, double* symbolic_ms, double* numeric_ms, int64_t* nnz) {
  if (!ctx) return DPP_ERR_INVALID;
  if (!ctx->have_params) {
    ctx->set_error("dpp_time_assembly: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  cudaSetDevice(ctx->device);
  return dpp::csr_time_phases(ctx, reps, symbolic_ms, numeric_ms, nnz);
}

void dpp_default_options(dpp_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->ksp_type = DPP_KSP_CG;
  o->pc_type = DPP_PC_JACOBI;
  o->fieldsplit_type = DPP_FS_MULTIPLICATIVE;
  o->inner_ksp_type = DPP_INNER_CG;
  o->inner_pc_type = DPP_PC_JACOBI;
  o->operator_mode = DPP_OP_MATRIX_FREE;
  o->max_it = 50000;       /* solvers/parameters.py:1 */
  o->gmres_restart = 30;   /* PETSc default */
  o->inner_max_it = 10000;
  o->check_every = 16;   /* = the direction-ring length of the fused CG: batches replay as one CUDA graph */
  o->rtol = 1e-8;          /* solvers/parameters.py:14 */
  o->atol = 1e-12;         /* solvers/parameters.py:15 */
  o->dtol = 1e4;           /* PETSc default */
  o->inner_rtol = 1e-10;
  o->inner_atol = 1e-50;
}

int dpp_solve(dpp_handle ctx, const dpp_options* opt, double* u_host, dpp_result* result, double* hist, int32_t hist_cap) {
  if (!ctx) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  dpp_options def;
  if (!opt) {
    dpp_default_options(&def);
    opt = &def;
  }
  if (opt->max_it < 0 || opt->gmres_restart < 1 || opt->check_every < 1) {
    ctx->set_error("dpp_solve: invalid options");
    return DPP_ERR_INVALID;
  }
  return dpp::krylov_solve(ctx, opt, u_host, result, hist, hist_cap);
}

const double* dpp_solution_dev(dpp_handle ctx) { return ctx ? ctx->d_solution : nullptr; }

int dpp_time_apply(dpp_handle ctx, int mode, int warmup, int reps, int with_dot, double* mean_ms) {
  if (!ctx || reps <= 0 || !mean_ms) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  double *dx = nullptr, *dy = nullptr;
  DPP_CHECK(dpp::krylov_work_vectors(ctx, &dx, &dy));
  const long long len = 2 * ctx->n_nodes;
  k_fill_pseudo<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(len, dx);
  k_zero_masked<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(len, ctx->d_mask, dx);  // as every Krylov vector
  ctx->launches += 2;
  DPP_CUDA(cudaGetLastError());
  cudaEvent_t e0, e1;
  DPP_CUDA(cudaEventCreate(&e0));
  DPP_CUDA(cudaEventCreate(&e1));
  int nb = 0;
  for (int i = 0; i < warmup; ++i) DPP_CHECK(apply_common(ctx, dx, dy, mode, with_dot != 0, &nb, true));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  DPP_CUDA(cudaEventRecord(e0, ctx->stream));
  for (int i = 0; i < reps; ++i) DPP_CHECK(apply_common(ctx, dx, dy, mode, with_dot != 0, &nb, true));
  DPP_CUDA(cudaEventRecord(e1, ctx->stream));
  DPP_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  DPP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *mean_ms = (double)ms / reps;
  return DPP_OK;
}

int dpp_time_cg_kernels(dpp_handle ctx, int warmup, int reps, double* apply_ms, double* update_ms, double* matvec_ms) {
  if (!ctx || reps <= 0 || warmup < 0 || !apply_ms || !update_ms || !matvec_ms) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return dpp::krylov_time_cg_kernels(ctx, warmup, reps, apply_ms, update_ms, matvec_ms);
}

int dpp_time_cg_block_kernels(dpp_handle ctx, int field, int warmup, int reps, double* apply_ms, double* update_ms) {
  if (!ctx || reps <= 0 || warmup < 0 || !apply_ms || !update_ms || field < 0 || field > 1) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  double unused = 0.0;
  return dpp::krylov_time_cg_kernels(ctx, warmup, reps, apply_ms, update_ms, &unused, 1, field);
}

int dpp_error_norms(dpp_handle ctx, const double* u_host, const double* exact_host, int nq, double* out4) {
  if (!ctx || !out4) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!ctx->have_params) {
    ctx->set_error("dpp_error_norms: call dpp_set_params first");
    return DPP_ERR_STATE;
  }
  if (ctx->world > 1) {
    ctx->set_error("dpp_error_norms: single-GPU handles only");
    return DPP_ERR_INVALID;
  }
  double *dx = nullptr, *dy = nullptr;
  DPP_CHECK(dpp::krylov_work_vectors(ctx, &dx, &dy));   // dx: u, dy: exact, scratch: staging for the numbering map
  double* dt = dpp::krylov_scratch_vector(ctx);
  const size_t bytes = sizeof(double) * 2 * ctx->n_nodes;
  const double* du = ctx->d_solution;
  if (u_host) {
    if (ctx->d_perm) {
      DPP_CUDA(cudaMemcpyAsync(dt, u_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
      DPP_CHECK(dpp::perm_to_internal(ctx, dt, dx, 2));
    } else {
      DPP_CUDA(cudaMemcpyAsync(dx, u_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    du = dx;
  } else if (!du) {
    ctx->set_error("dpp_error_norms: no solution on the device yet");
    return DPP_ERR_STATE;
  }
  const double* de = nullptr;
  if (exact_host) {
    if (ctx->d_perm) {
      DPP_CUDA(cudaMemcpyAsync(dt, exact_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
      DPP_CHECK(dpp::perm_to_internal(ctx, dt, dy, 2));
    } else {
      DPP_CUDA(cudaMemcpyAsync(dy, exact_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    de = dy;
  }
  return dpp::error_norms(ctx, du, de, nq, out4);
}

int dpp_darcy_velocity(dpp_handle ctx, const double* p_host, int field, double conductivity, double rtol, int32_t max_it,
                       double* velocity_host, int32_t* iterations) {
  if (!ctx || !velocity_host || field < 0 || field > 1 || !(rtol > 0.0) || max_it < 1) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->world > 1) {
    ctx->set_error("dpp_darcy_velocity: single-GPU handles only");
    return DPP_ERR_INVALID;
  }
  if (!p_host && !ctx->d_solution) {
    ctx->set_error("dpp_darcy_velocity: no solution on the device yet");
    return DPP_ERR_STATE;
  }
  const int64_t n = ctx->n_nodes;
  const int dim = ctx->dim;
  double* buf = nullptr;   // [p | rhs (dim) | velocity (dim) | staging (dim)]
  DPP_CUDA(cudaMalloc((void**)&buf, sizeof(double) * (size_t)(1 + 3 * dim) * n));
  double *dp = buf, *drhs = buf + n, *dvel = buf + (1 + dim) * n, *dstage = buf + (1 + 2 * dim) * n;
  int rc = DPP_OK;
  auto run = [&]() -> int {
    if (p_host) {
      if (ctx->d_perm) {
        DPP_CUDA(cudaMemcpyAsync(dstage, p_host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        DPP_CHECK(dpp::perm_to_internal(ctx, dstage, dp, 1));
      } else {
        DPP_CUDA(cudaMemcpyAsync(dp, p_host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
      }
    } else {
      DPP_CUDA(cudaMemcpyAsync(dp, ctx->d_solution + (size_t)field * n, sizeof(double) * n, cudaMemcpyDeviceToDevice,
                               ctx->stream));
    }
    DPP_CHECK(dpp::darcy_velocity(ctx, dp, conductivity, rtol, max_it, drhs, dvel, iterations, nullptr));
    const double* src = dvel;
    if (ctx->d_perm) {
      for (int c = 0; c < dim; ++c) DPP_CHECK(dpp::perm_to_user(ctx, dvel + c * n, dstage + c * n, 1));
      src = dstage;
    }
    DPP_CUDA(cudaMemcpyAsync(velocity_host, src, sizeof(double) * dim * n, cudaMemcpyDeviceToHost, ctx->stream));
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
    return DPP_OK;
  };
  rc = run();
  cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  return rc;
}

int dpp_lanczos(dpp_handle ctx, int which, int32_t steps, uint64_t seed, double* alpha_host, double* beta_host,
                int32_t* steps_done) {
  if (!ctx || !alpha_host || !beta_host || !steps_done || which < 0 || which > 2 || steps < 1) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->world > 1) {
    ctx->set_error("dpp_lanczos: single-GPU handles only");
    return DPP_ERR_INVALID;
  }
  int done = 0;
  DPP_CHECK(dpp::krylov_lanczos(ctx, which, steps, seed, alpha_host, beta_host, &done));
  *steps_done = done;
  return DPP_OK;
}

int dpp_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return DPP_ERR_INVALID;
  return cudaMallocHost(ptr, (size_t)bytes) == cudaSuccess ? DPP_OK : DPP_ERR_CUDA;
}

int dpp_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? DPP_OK : DPP_ERR_CUDA; }

int dpp_plan_x_segments(int tiles, int planes, int resident_ctas, int max_ctas) {
  if (tiles < 1 || planes < 1 || resident_ctas < 1 || max_ctas < tiles) return DPP_ERR_INVALID;
  return dpp::choose_x_segments(tiles, planes, resident_ctas, max_ctas);
}

int dpp_kernel_launch_count(dpp_handle ctx, int64_t* launches) {
  if (!ctx || !launches) return DPP_ERR_INVALID;
  *launches = ctx->launches;
  return DPP_OK;
}

}  // extern "C"
