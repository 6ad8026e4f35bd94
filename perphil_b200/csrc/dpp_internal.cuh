// Internal declarations shared by the translation units of libdppb200.so.
// Nothing here crosses the C ABI (include/dpp_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dpp_b200.h"

#define DPP_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ctx->set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                     std::to_string(__LINE__) + ")");                                           \
      return DPP_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define DPP_CHECK(expr)                  \
  do {                                   \
    int rc__ = (expr);                   \
    if (rc__ != DPP_OK) return rc__;     \
  } while (0)

namespace dpp {

constexpr int kMaxBand = 5;  // 2*P+1 for P<=2

// y_f = sum_g (cK[f][g] K + cM[f][g] M) x_g  : the 2x2 block structure of dpp_form
struct Coef {
  double cK[2][2];
  double cM[2][2];
};

// One application of a (block of the) DPP operator with Firedrake DirichletBC semantics.
struct OpArgs {
  int nf;                      // 1 or 2 fields
  const double* x[2];
  double* y[2];
  const uint8_t* in_mask[2];   // !=0 -> input treated as 0 (column elimination); may be null
  const uint8_t* out_mask[2];  // !=0 -> row replaced (row elimination); may be null
  int identity_on_masked;      // 1: y = x on eliminated rows (diagonal 1); 0: y = 0
  Coef c;
  double* dot_partials;        // optional: per-block partial sums of sum_f <x_f, y_f> over owned rows
  int64_t owned_begin, owned_end;  // rows (local node ids) to compute
  const double* skip_flag;     // optional device scalar: != 0 -> the launch is a no-op
  int input_premasked;         // caller guarantees x == 0 wherever in_mask != 0 (all Krylov vectors)
};

// Tensor-grid description for the structured family. Axis 0 = x (slowest), 2 = z (contiguous).
// 2-D meshes are stored with a 1-node dummy axis 0 (m = 1, k = 0).
struct GridDesc {
  int n[3];                 // nodes per axis
  int band;                 // P: half bandwidth of the 1-D matrices (degree)
  const double* m1d[3];     // device: [n[a]][2P+1] assembled 1-D mass rows (zero outside domain)
  const double* k1d[3];     // device: [n[a]][2P+1] assembled 1-D stiffness rows
};

struct Krylov;  // krylov.cu
struct FusedState;  // cg_fused_uniform.cu
struct CsrMatrix;
struct Comm;
struct CellBlocks;  // apply_cells.cu

}  // namespace dpp

struct dpp_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int64_t device_bytes = 0;

  // mesh
  int dim = 0, degree = 0, npc = 0, nvc = 0;
  int64_t n_nodes = 0, n_cells = 0, n_coord_nodes = 0;
  int32_t* d_cnm = nullptr;       // [n_cells*npc]
  double* d_coords = nullptr;     // [n_coord_nodes*dim]
  int32_t* d_ccnm = nullptr;      // [n_cells*nvc] (aliases d_cnm when degree==1 and same array)
  bool ccnm_alias = false;

  int32_t* d_perm = nullptr;      // optional user -> internal node map (dpp_set_numbering)

  // kernel family
  int family = DPP_KERNEL_GENERAL;
  bool structured_ok = false;
  dpp::GridDesc grid{};
  double* d_tables = nullptr;     // backing store of grid.m1d/k1d
  std::vector<double> h_axis[3];  // 1-D vertex coordinates per axis (structured)
  bool grid_uniform = false;      // equal spacing on every axis -> apply_structured_uniform.cu
  bool q2_uniform = false;        // degree 2 lattice with equal cell size on every axis (uni_h)
  double uni_h[3] = {1.0, 1.0, 1.0};
  bool force_table_kernel = false;
  bool fused_cg_disabled = false;   // dpp_set_fused_cg(h, 0): ranks of a slab run agree on ONE protocol
  double uni_m_off[3] = {0, 0, 0}, uni_k_off[3] = {0, 0, 0};
  double uni_mxc[2] = {0, 0}, uni_kxc[2] = {0, 0};  // axis-0 centre entries [interior, boundary]

  // general family: node -> (cell, local index) adjacency, per-cell geometry
  int64_t* d_adj_ptr = nullptr;   // [n_nodes+1]
  int32_t* d_adj = nullptr;       // packed cell*32 + local   (cell < 2^26) or two arrays; see apply_general.cu
  int32_t* d_adj_cell = nullptr;
  uint8_t* d_adj_loc = nullptr;
  double* d_cell_geom = nullptr;  // [n_cells*8]: affine metric (6) + detJ + flag
  bool general_ready = false;
  dpp::CellBlocks* cells = nullptr;   // cell-block decomposition of the element-based Q1 hex kernel (apply_cells.cu)

  // parameters
  bool have_params = false;
  double k1 = 0, k2 = 0, beta = 0, mu = 1;

  // Dirichlet data
  uint8_t* d_mask = nullptr;      // [2*n_nodes]
  double* d_g = nullptr;          // [2*n_nodes]  (0 where unconstrained)
  bool have_bc[2] = {false, false};
  int32_t* d_bc_nodes[2] = {nullptr, nullptr};  // constrained node ids per field
  int64_t n_bc[2] = {0, 0};
  int64_t bc_cap[2] = {0, 0};     // capacity of d_bc_nodes
  double* d_bc_vals = nullptr;    // upload staging for Dirichlet values
  int64_t bc_vals_cap = 0;
  int64_t bc_gen[2] = {0, 0};     // bumped by every dpp_set_dirichlet of that field

  // partition
  int rank = 0, world = 1;
  int64_t owned_begin = 0, owned_end = 0;
  int dom_lo = 1, dom_hi = 1;     // first / last stored x-plane lies on the domain boundary (else: ghost plane)
  dpp::Comm* comm = nullptr;

  // work vectors / solver state (krylov.cu)
  dpp::Krylov* krylov = nullptr;
  dpp::FusedState* fused = nullptr;  // padded vectors + TMA descriptors of the fused CG path
  dpp::CsrMatrix* csr = nullptr;
  double* d_solution = nullptr;   // [2*n_nodes]
  double* d_diag = nullptr;       // [2*n_nodes] diag(A_bc), valid when diag_valid
  double* d_premask = nullptr;    // [2*n_nodes] scratch: input with eliminated columns zeroed
  bool diag_valid = false;

  // reduction scratch
  double* d_partials = nullptr;   // [kMaxPartialBlocks * kMaxDotWidth]
  double* d_scalars = nullptr;    // device scalar block
  double* h_scalars = nullptr;    // pinned mirror
  unsigned* d_counters = nullptr; // arrival counters of the folded reductions [4]
  double* d_dtab = nullptr;       // [2 slots][2 fields][64] reciprocal-diagonal class tables of the fused CG
  double* d_hist[2] = {nullptr, nullptr};  // residual history per solver slot
  int hist_cap[2] = {0, 0};

  void set_error(const std::string& s) { err = s; }
  unsigned long long state_gen = 0;  // bumped whenever parameters / BCs / numbering / partition change
  void invalidate() { diag_valid = false; ++state_gen; }
};

namespace dpp {

constexpr int kMaxPartialBlocks = 4096;
constexpr int kMaxDotWidth = 40;   // >= gmres restart + 2
constexpr int kMaxGmresRestart = 30;   // Givens / Hessenberg state of the device-resident GMRES (krylov.cu)
constexpr int kNumScalars = 256;

// Number of x-segments for the plane-streaming kernels: CTAs are (tile, segment) items dispatched by the
// hardware in linear order (tile fastest), so all tiles of a segment run concurrently (halo rows/columns
// of neighbouring tiles then hit in L2 -- a fully persistent partition loses that, measured: 2x DRAM
// reads).  Cost model: rounds of `capacity` resident CTAs x (planes per segment + 2 redundant planes +
// prologue + drift penalty); returns the segment count with the smallest cost.
inline int choose_x_segments(int tiles, int nown, int capacity, int max_ctas, int redundant = 2) {
  int best = 1;
  long long best_cost = -1;
  for (int nseg = 1; nseg <= nown; ++nseg) {
    if ((long long)tiles * nseg > max_ctas && nseg > 1) break;
    const int len = (nown + nseg - 1) / nseg;
    if (len < 4 && nseg > 1) break;
    const long long rounds = ((long long)tiles * nseg + capacity - 1) / capacity;
    // long runs let neighbouring tiles drift apart (no synchronisation between CTAs): beyond ~32 planes their
    // halo rows/columns start to miss in L2 (measured at 256^3: 2 segments of 128 planes 327 us, 9 of 29 303 us)
    const long long drift = len > 32 ? (len - 32) / 4 : 0;
    const long long cost = rounds * (len + redundant + 2 + drift);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nseg; }
  }
  return best;
}

// ---- apply_structured.cu
int structured_detect_and_setup(dpp_context* ctx, const int32_t* cnm_host, const double* coords_host,
                                const int32_t* ccnm_host);
int structured_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks);
int structured_diagonal(dpp_context* ctx, const Coef& c, double* d_diag /*[2*n_nodes]*/);
int structured_apply_uniform(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks);
// row elimination after a uniform-grid apply: y_f[node] = identity ? xid_f[node] : 0 on the constrained
// nodes of mask field fld[f] (owned rows only)
int structured_fix_rows(dpp_context* ctx, int nf, const int* fld, double* const* y, const double* const* xid,
                        int identity, const double* skip_flag);

// ---- apply_general.cu
int general_setup(dpp_context* ctx, const int32_t* cnm_host);
int general_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks);
int general_diagonal(dpp_context* ctx, const Coef& c, double* d_diag);

// ---- apply_cells.cu: element-based kernel for general Q1 hexahedral meshes
bool cells_supported(const dpp_context* ctx);
bool cells_ready(const dpp_context* ctx);
int cells_setup(dpp_context* ctx);
int cells_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks);
int cells_stats(const dpp_context* ctx, int64_t* n_blocks, int64_t* total_slots, int64_t* affine_cells);
void cells_destroy(dpp_context* ctx);

// ---- operator.cu : dispatch + DPP coefficient blocks
Coef dpp_coef(const dpp_context* ctx);                    // monolithic 2x2
Coef block_coef(const dpp_context* ctx, int row, int col); // single block as nf=1 coefficient
int op_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks);
int op_diagonal(dpp_context* ctx);                        // fills ctx->d_diag with diag(A_bc)

// ---- vector_ops.cu
int vec_zero(dpp_context* ctx, double* x, int64_t n);
int vec_copy(dpp_context* ctx, double* dst, const double* src, int64_t n);

// ---- comm.cu
int comm_halo_exchange(dpp_context* ctx, double* const* fields, int nf);
int comm_allreduce_sum(dpp_context* ctx, double* d_vals, int n);
// ghost x-planes of a plane-contiguous (padded) vector: owned boundary planes -> the slab neighbours
int comm_halo_planes(dpp_context* ctx, double* base, int nf, long long field_stride, long long plane_elems, int i_begin,
                     int i_end);
void cg_fused_destroy(dpp_context* ctx);

// peer-memory (CUDA IPC) fast path, comm.cu
constexpr int kMaxIpcRanks = 16;
constexpr int kMboxEntry = 8;      // doubles per mailbox entry: 7 values + sequence flag
constexpr int kMboxWords = 16;     // 8-byte words per entry: tagged-word protocol = 2 words per value (cg_device.cuh)
// wide entries behind the narrow ones and the sequence counter (same allocation, same IPC handle): the classical
// Gram-Schmidt dot products of GMRES(30), up to 32 values per reduction (vector_ops.cu: k_reduce_partials_wide)
constexpr int kMboxWideVals = 32;
constexpr int kMboxWideWords = 2 * kMboxWideVals;
constexpr size_t kMboxWideOffset = 2 * (size_t)kMaxIpcRanks * kMboxWords + 2;
constexpr size_t kMboxDoubles = kMboxWideOffset + 2 * (size_t)kMaxIpcRanks * kMboxWideWords;
struct IpcReduce {                 // kernel argument of the mailbox allreduce
  double* local;                   // [2 slots][world][kMboxWords]
  double* peer[kMaxIpcRanks];      // the same array of every rank (peer[rank] == local)
  int rank, world;                 // world == 1: no exchange
  int ll;                          // 1: tagged words (value = flag), 0: values + fence + flag word;
                                   // measurement: 2 = no system fences at all, 3 = fence in every reduction
  unsigned long long* seq_dev;     // own device counter: number of exchanges executed so far
};
struct FoldArgs {                  // reduction epilogue folded into the producing kernel ("last block done")
  int enabled;
  unsigned* counter;               // self-resetting arrival counter
  double* S;                       // scalar slot (read-write)
  double* hist;
  int post;
  IpcReduce ipc;
  double* xring;                   // deferred x update: step lengths + tags of the direction ring (else null)
};
struct IpcHalo {                   // kernel argument of the halo push (padded layout)
  double* peer_r[2];               // lower / upper neighbour's residual vector (null: none)
  long long peer_field[2];         // their padded field stride
  long long peer_ghost_off[2];     // offset of the ghost plane that mirrors my boundary plane
  int debug_fence_all;             // DPP_DEBUG_FENCE_ALL: system fence in every block (measurement)
};
bool comm_ipc_ready(const dpp_context* ctx);        // mailbox all-reduce
bool comm_ipc_halo_ready(const dpp_context* ctx);   // + halo push into the neighbours' residual vectors
bool comm_ipc_box_ready(const dpp_context* ctx);    // + halo inboxes for generic vectors (comm_halo_exchange)
int comm_halo_failed(dpp_context* ctx);             // a k_halo_get timed out
IpcReduce comm_ipc_reduce_args(dpp_context* ctx);   // world == 1 when the mailbox path is not active
IpcHalo comm_ipc_halo(const dpp_context* ctx);
// residual buffer registration (cg_fused_uniform.cu owns the memory)
double* cg_fused_r_buffer(dpp_context* ctx, long long* field, long long* plane);
void comm_destroy(dpp_context* ctx);

// ---- assemble_csr.cu
int csr_assemble(dpp_context* ctx, int64_t* nnz);
int csr_export(dpp_context* ctx, int64_t* indptr, int32_t* indices, double* data);
int csr_export_block(dpp_context* ctx, int fr, int fc, int64_t* indptr, int32_t* indices, double* data);
int csr_time_phases(dpp_context* ctx, int reps, double* symbolic_ms, double* numeric_ms, int64_t* nnz);
int csr_spmv(dpp_context* ctx, const double* x, double* y, double* dot_partials, int* n_partial_blocks,
             const double* skip_flag);
void csr_invalidate(dpp_context* ctx);
bool csr_valid(const dpp_context* ctx);
void csr_destroy(dpp_context* ctx);

// ---- numbering map helpers (dpp_api.cu): internal <-> caller numbering of field-blocked vectors
// ---- darcy.cu / krylov.cu: post-processing and spectrum estimates
int darcy_velocity(dpp_context* ctx, const double* d_p, double conductivity, double rtol, int max_it, double* d_rhs,
                   double* d_vel, int32_t* iterations, double* residuals);
int krylov_lanczos(dpp_context* ctx, int which, int steps, unsigned long long seed, double* alpha, double* beta, int* done);
int perm_to_internal(dpp_context* ctx, const double* user, double* internal, int nf);
int perm_to_user(dpp_context* ctx, const double* internal, double* user, int nf);

// ---- error_norms.cu
int error_norms(dpp_context* ctx, const double* d_u, const double* d_exact, int nq, double out[4]);

// ---- krylov.cu
int krylov_solve(dpp_context* ctx, const dpp_options* opt, double* u_host, dpp_result* res,
                 double* hist_host, int32_t hist_cap);
void krylov_destroy(dpp_context* ctx);
int krylov_time_cg_kernels(dpp_context* ctx, int warmup, int reps, double* apply_ms, double* update_ms,
                           double* matvec_ms, int nf = 2, int field = 0);

template <typename T>
inline int dev_alloc(dpp_context* ctx, T** p, int64_t count) {
  size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void**)p, bytes);
  if (e != cudaSuccess) {
    ctx->set_error(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    return DPP_ERR_CUDA;
  }
  ctx->device_bytes += (int64_t)bytes;
  return DPP_OK;
}

}  // namespace dpp
