/*
 * dpp_b200.h -- C ABI of libdppb200.so: the B200 (sm_100a) implementation of perphil's hot path,
 * "assemble + solve the linear double-porosity/permeability (DPP) pressure system".
 *
 * perphil (reference, /root/reference) has no FFI of its own: its hot path is
 *     perphil.solvers.solver.solve_dpp            (src/perphil/solvers/solver.py:30-76)
 *     perphil.forms.dpp.dpp_form                  (src/perphil/forms/dpp.py:95-132)
 *     perphil.solvers.conditioning.get_matrix_data_from_form   (src/perphil/solvers/conditioning.py:66-102)
 * and everything below those calls runs inside Firedrake/PETSc.  The entry points declared here
 * are what a ctypes binding for that path binds instead; each one names the
 * reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative dpp_status; dpp_last_error() gives text;
 *   - plain pointers and sizes only; the caller owns every HOST buffer it passes, the library
 *     copies what it needs to the device and owns all device memory until dpp_destroy();
 *   - pointers named *_host are host memory; pointers named *_dev are device memory of the
 *     handle's GPU; calls are blocking unless stated; one handle = one GPU = one CUDA stream;
 *   - DOF layout everywhere: field-blocked [p1(0..n_nodes-1) ; p2(0..n_nodes-1)], node numbering
 *     exactly as supplied in cell_node_map (iterative_bench.py:323-324 relies on this layout);
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef DPP_B200_H
#define DPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpp_context* dpp_handle;

typedef enum {
  DPP_OK = 0,
  DPP_ERR_INVALID = -1,     /* bad argument / unsupported combination           */
  DPP_ERR_CUDA = -2,        /* CUDA runtime error                               */
  DPP_ERR_STATE = -3,       /* call order (e.g. solve before set_params)        */
  DPP_ERR_NCCL = -4,        /* NCCL error                                       */
  DPP_ERR_NO_DEVICE = -5    /* no usable sm_100 GPU: the library never falls back */
} dpp_status;

/* KSP / PC vocabulary mirrors the PETSc option values used in solvers/parameters.py:1-57 */
typedef enum { DPP_KSP_CG = 0, DPP_KSP_GMRES = 1, DPP_KSP_PICARD = 2 } dpp_ksp_type;
typedef enum { DPP_PC_NONE = 0, DPP_PC_JACOBI = 1, DPP_PC_PBJACOBI = 2, DPP_PC_FIELDSPLIT = 3 } dpp_pc_type;
typedef enum { DPP_FS_ADDITIVE = 0, DPP_FS_MULTIPLICATIVE = 1 } dpp_fieldsplit_type;
typedef enum { DPP_INNER_PREONLY = 0, DPP_INNER_CG = 1 } dpp_inner_ksp_type;
typedef enum { DPP_OP_MATRIX_FREE = 0, DPP_OP_ASSEMBLED = 1 } dpp_operator_mode;
/* which matrix-free kernel family serves the handle (reported by dpp_get_info) */
typedef enum { DPP_KERNEL_GENERAL = 0, DPP_KERNEL_STRUCTURED = 1 } dpp_kernel_family;

/* PETSc KSPConvergedReason values (positive = converged) */
enum {
  DPP_CONVERGED_RTOL = 2, DPP_CONVERGED_ATOL = 3, DPP_CONVERGED_ITS = 4,
  DPP_DIVERGED_ITS = -3, DPP_DIVERGED_DTOL = -4, DPP_DIVERGED_BREAKDOWN = -5,
  DPP_DIVERGED_INDEFINITE_MAT = -10, DPP_DIVERGED_NANORINF = -9,
  DPP_DIVERGED_COMM_TIMEOUT = -100  /* a peer never arrived at a reduction (not a PETSc reason) */
};

typedef struct {
  int32_t ksp_type;          /* dpp_ksp_type         <- "ksp_type" / "snes_type" (Picard)         */
  int32_t pc_type;           /* dpp_pc_type          <- "pc_type"                                 */
  int32_t fieldsplit_type;   /* dpp_fieldsplit_type  <- "pc_fieldsplit_type"                      */
  int32_t inner_ksp_type;    /* dpp_inner_ksp_type   <- "fieldsplit_{0,1}" -> "ksp_type"          */
  int32_t inner_pc_type;     /* NONE | JACOBI        <- "fieldsplit_{0,1}" -> "pc_type"           */
  int32_t operator_mode;     /* dpp_operator_mode    <- "mat_type": "matfree" | "aij"             */
  int32_t max_it;            /* "ksp_max_it" (50000 in solvers/parameters.py:1)                   */
  int32_t gmres_restart;     /* "ksp_gmres_restart" (PETSc default 30)                            */
  int32_t inner_max_it;
  int32_t check_every;       /* host polls the device convergence flag every N iterations (>=1)   */
  double rtol;               /* "ksp_rtol" 1e-8 (parameters.py:14)                                */
  double atol;               /* "ksp_atol" 1e-12                                                  */
  double dtol;               /* PETSc default 1e4                                                 */
  double inner_rtol;
  double inner_atol;
} dpp_options;

typedef struct {
  int32_t iterations;        /* KSP its (outer)        -> Solution.iteration_number (solver.py:73) */
  int32_t converged_reason;  /* KSPConvergedReason                                                 */
  int32_t inner_iterations;  /* total inner (fieldsplit / Picard block) CG iterations              */
  int32_t history_len;       /* entries written to residual_history                                */
  double residual_norm;      /* KSP residual norm      -> Solution.residual_error (solver.py:74)   */
  double rhs_norm;           /* ||b||_2 of the lifted system = "0 SNES Function norm"              */
  double solve_ms;           /* device time of the Krylov loop (CUDA events)                       */
  double setup_ms;           /* device time of lifting + preconditioner setup                      */
  double apply_ms;           /* device time spent in operator applies inside the solve             */
  int64_t apply_count;
} dpp_result;

typedef struct {
  int32_t kernel_family;     /* dpp_kernel_family */
  int32_t dim, degree;
  int32_t grid_nodes[3];     /* structured: nodes per axis (x slowest); else 0 */
  int64_t n_nodes, n_cells, n_owned_nodes;
  int32_t rank, world;
  int32_t sm_count;
  int32_t peer_memory;       /* bit 0: mailbox all-reduce, bit 1: fused-CG halo push, bit 2: halo inboxes for generic
                                vectors (all over CUDA IPC peer memory, dpp_comm_ipc_import); 0: NCCL only */
  int64_t device_bytes;      /* device memory held by the handle */
} dpp_info;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Upload one mesh + scalar function space V (W = V x V).  Replaces what Firedrake derives from
 * MixedFunctionSpace((V, V)) in solver.py:64-69: V.cell_node_map().values, mesh.coordinates.
 *   cell_node_map        [n_cells * nodes_per_cell] int32, local order tensor-lexicographic
 *                        (x slowest); nodes_per_cell = (degree+1)^dim
 *   coords               [n_coord_nodes * dim] vertex coordinates (the Q1 coordinate field)
 *   coord_cell_node_map  [n_cells * 2^dim] (may alias cell_node_map when degree == 1)
 * The library inspects the data: a rectilinear tensor grid numbered lexicographically selects the
 * DPP_KERNEL_STRUCTURED family, anything else DPP_KERNEL_GENERAL. */
int dpp_create(dpp_handle* h, int device, int dim, int degree, int64_t n_nodes, int64_t n_cells,
               int nodes_per_cell, const int32_t* cell_node_map_host, int64_t n_coord_nodes,
               const double* coords_host, const int32_t* coord_cell_node_map_host);
void dpp_destroy(dpp_handle h);
const char* dpp_last_error(dpp_handle h); /* h may be NULL: error of the last failed dpp_create */
int dpp_get_info(dpp_handle h, dpp_info* info);
/* force a kernel family (testing: run the general kernels on a structured mesh). Call before
 * set_params/set_dirichlet. */
int dpp_force_kernel_family(dpp_handle h, int family);

/* Optional numbering map.  The kernels are fastest on lexicographically numbered tensor grids; Firedrake
 * numbers the same lattice arbitrarily (DMPlex order).  The host layer may therefore
 * create the handle on a lexicographically RE-numbered copy of the mesh and register the map here: from
 * then on every HOST-facing node list and vector (dpp_set_dirichlet, dpp_apply_host,
 * dpp_get_diagonal_host, dpp_solve's u_host) is in the caller's numbering, node u being stored internally
 * at user_to_internal[u].  Device-pointer calls and dpp_get_csr_host stay in internal numbering.
 * Single-GPU handles only. */
int dpp_set_numbering(dpp_handle h, const int32_t* user_to_internal_host);

/* DPPParameters (models/dpp/parameters.py:5-53): float(k1), float(k2), float(beta), float(mu). */
int dpp_set_params(dpp_handle h, double k1, double k2, double beta, double mu);

/* fd.DirichletBC(W.sub(field), g, ...) (README.md:79-82): bc.nodes and g at those nodes.
 * Replaces the previous set for that field; n == 0 clears it. */
int dpp_set_dirichlet(dpp_handle h, int field, int64_t n, const int32_t* nodes_host, const double* values_host);

/* ---- distributed (slab partition; one handle per rank/GPU) ---------------------------------- */

/* The local mesh holds owned + ghost nodes; rows are computed for local ids [owned_begin,
 * owned_end).  nccl_unique_id = the 128-byte ncclUniqueId created by rank 0 and broadcast by the
 * host layer (torch.distributed). */
int dpp_comm_init(dpp_handle h, int rank, int world, const void* nccl_unique_id, int64_t owned_begin,
                  int64_t owned_end);
/* ncclGetUniqueId into a caller-provided 128-byte buffer (rank 0 calls it, the host layer broadcasts) */
int dpp_nccl_unique_id(void* out128);
/* one neighbour: local node ids whose values are sent to / received from `peer` before an apply */
int dpp_comm_add_neighbor(dpp_handle h, int peer, int64_t n_send, const int32_t* send_nodes_host,
                          int64_t n_recv, const int32_t* recv_nodes_host);

/* Peer-memory fast path of the fused Jacobi-CG iteration (optional; one box, NVLink/NVSwitch): every
 * rank exports CUDA IPC handles of its residual vector and of a small mailbox; after the import the
 * r-update kernel stores its boundary planes straight into the neighbours' ghost planes and the Krylov
 * scalars are summed through the mailboxes inside the reduction kernel (no NCCL call per iteration).
 * Protocol: dpp_comm_init + dpp_comm_add_neighbor on every rank, then dpp_comm_ipc_export on every rank,
 * all-gather the blobs (host layer), dpp_comm_ipc_import on every rank.  Without it the NCCL path runs. */
int dpp_comm_ipc_blob_size(void);
int dpp_comm_ipc_export(dpp_handle h, void* blob_out);
int dpp_comm_ipc_import(dpp_handle h, const void* blobs_all_ranks);
/* back to the NCCL path (the host layer calls it on every rank when any rank's import did not succeed:
 * dpp_info.peer_memory must agree across ranks; bit 0 = mailbox all-reduce, bit 1 = halo push) */
int dpp_comm_ipc_disable(dpp_handle h);

/* Slab runs: whether the fused two-kernel Jacobi-CG iteration (uniform tensor grid, TMA available) serves this
 * handle is a rank-local fact, but the number and kind of reductions per iteration depends on it, so all ranks must
 * run the same protocol.  The host layer all-gathers dpp_fused_cg_supported() (1 / 0) and, when the ranks disagree,
 * calls dpp_set_fused_cg(h, 0) on every rank (the unfused kernel sequence then runs everywhere). */
int dpp_fused_cg_supported(dpp_handle h);
int dpp_set_fused_cg(dpp_handle h, int enable);

/* ---- operator ------------------------------------------------------------------------------- */

/* y = A_bc x with A_bc = P A P + (I - P) (Firedrake DirichletBC semantics), A the dpp_form matrix
 * (forms/dpp.py:27,57,89).  x, y: [2*n_nodes] doubles.  This is PETSc MatMult on the path
 * (solver.py:71).  *_dev variant: no copies, asynchronous on the handle's stream. */
int dpp_apply_host(dpp_handle h, const double* x_host, double* y_host, int operator_mode);
int dpp_apply_dev(dpp_handle h, const double* x_dev, double* y_dev, int operator_mode);
/* diag(A_bc) -> [2*n_nodes] */
int dpp_get_diagonal_host(dpp_handle h, double* diag_host);

/* ---- assembly: fd.assemble(a, bcs=bcs, mat_type="aij") (conditioning.py:51-63) --------------- */

/* Builds the 2x2-block CSR (full element pattern, sorted columns, explicit zeros where Dirichlet
 * rows/columns were eliminated -- conditioning.py:86 then calls eliminate_zeros()). */
int dpp_assemble_csr(dpp_handle h, int64_t* nnz);
/* petsc_matrix.getValuesCSR() (conditioning.py:85): indptr [2*n_nodes+1], indices/data [nnz] */
int dpp_get_csr_host(dpp_handle h, int64_t* indptr_host, int32_t* indices_host, double* data_host);

/* One block of that matrix as a scalar-space CSR matrix (row field, column field in {0, 1}): what
 * get_matrix_data_from_form(a_macro, [bc]) / (a_micro, [bc]) of the dpp_delayed_form split (forms/dpp.py:135-205,
 * used at notebooks/conforming-galerkin-fem-operator-splitting-2D-perphil.py:463-480) returns, and the slices
 * csr[:n0, :n0] / csr[n0:, n0:] of iterative_bench.py:323-324, gathered on the device.
 * indptr [n_nodes+1], indices / data [nnz / 4]; call dpp_assemble_csr first. */
int dpp_get_csr_block_host(dpp_handle h, int row_field, int col_field, int64_t* indptr_host, int32_t* indices_host,
                           double* data_host);

/* ---- solve: LinearVariationalSolver.solve() + ksp getters (solver.py:66-74) ------------------ */

void dpp_default_options(dpp_options* opt);
/* Lifts the Dirichlet data (u0 = g on Gamma; b = -(A u0) on interior rows), solves A_bc d = b
 * from d = 0 with the configured Krylov method, writes u = u0 + d.  u_host: [2*n_nodes] (may be
 * NULL: keep the solution on the device, see dpp_solution_dev).  residual_history (optional):
 * one entry per KSP iteration starting with iteration 0, like -ksp_monitor. */
int dpp_solve(dpp_handle h, const dpp_options* opt, double* u_host, dpp_result* result,
              double* residual_history_host, int32_t history_capacity);
const double* dpp_solution_dev(dpp_handle h);

/* ---- post-processing: error norms (utils/postprocessing.py:89-124) ---------- */

/* out = { ||p1_h - p1||^2_L2, ||p2_h - p2||^2_L2, |p1_h - p1|^2_H1, |p2_h - p2|^2_H1 } by nq^dim-point Gauss
 * quadrature over the cells.  u_host [2*n_nodes] (NULL: the solution of the last dpp_solve);
 * exact_host [2*n_nodes] nodal exact field of the same space, or NULL for the manufactured closed form of
 * utils/manufactured_solutions.py:39-51 (2-D) / :82-88 (3-D) with the handle's parameters.  Single-GPU. */
int dpp_error_norms(dpp_handle h, const double* u_host, const double* exact_host, int nq, double* out4);

/* ---- post-processing: Darcy velocity (utils/postprocessing.py:34-63) ---------------- */

/* calculate_darcy_velocity_from_pressure(pressure_field, conductivity): fd.project(-k grad(p_h), V^dim) on the
 * pressure's own Lagrange space = one nodal-mass solve per component (Jacobi-CG to rtol, Firedrake's projection
 * default is 1e-8).  p_host [n_nodes] scalar nodal pressure, or NULL for field `field` (0|1) of the last dpp_solve;
 * velocity_host [dim * n_nodes], component-blocked (u_x of all nodes, then u_y, ...); iterations [dim] or NULL.
 * Single-GPU. */
int dpp_darcy_velocity(dpp_handle h, const double* p_host, int field, double conductivity, double rtol, int32_t max_it,
                       double* velocity_host, int32_t* iterations);

/* ---- spectrum estimates for the conditioning study (solvers/conditioning.py:105-218) ---- */

/* `steps` Lanczos steps on the symmetric BC'd operator (which = 0: monolithic A_bc, 1: block A00, 2: block A11,
 * the matrices get_matrix_data_from_form / iterative_bench.py:323-324 hand to calculate_condition_number), matrix-
 * free, from a pseudo-random start vector (seed).  alpha_host, beta_host [steps]: diagonal and sub-diagonal of the
 * Lanczos tridiagonal matrix, whose extreme eigenvalues converge to the extreme singular values of the (symmetric)
 * operator; steps_done < steps when an invariant subspace was exhausted.  Replaces the dense SVD that caps the
 * reference's 3-D study at N = 16.  Single-GPU. */
int dpp_lanczos(dpp_handle h, int which, int32_t steps, uint64_t seed, double* alpha_host, double* beta_host,
                int32_t* steps_done);

/* ---- page-locked host buffers for results (full-rate D2H of the solution vector) -------------- */
int dpp_host_alloc(void** ptr, int64_t bytes);   /* cudaMallocHost */
int dpp_host_free(void* ptr);

/* ---- measurement helpers (bench.py; timed with CUDA events on the handle's stream) ----------- */

/* mean device milliseconds of `reps` back-to-back applies of the monolithic operator on internal
 * vectors (after `warmup` untimed ones); with_dot fuses the (x, Ax) reduction as CG uses it. */
int dpp_time_apply(dpp_handle h, int operator_mode, int warmup, int reps, int with_dot, double* mean_ms);
/* mean device milliseconds of the two kernels of one fused Jacobi-CG iteration (uniform-grid path,
 * csrc/cg_fused_uniform.cu), each timed alone over `reps` launches on the solver's work vectors:
 * apply_ms = p/x update + matrix-free apply + <p,Ap> (+ row fix-up), update_ms = r update + <r,z>,<z,z>,
 * matvec_ms = the same TMA kernel in plain mode (w = A p, fused <p,Ap>: PETSc MatMult on the path).
 * Returns DPP_ERR_INVALID when the handle does not run the fused path. */
int dpp_time_cg_kernels(dpp_handle h, int warmup, int reps, double* apply_ms, double* update_ms, double* matvec_ms);
/* the same two kernels on the one-field diagonal block of `field` (0 or 1): the iteration of the block solves inside
 * PCFIELDSPLIT and the block Picard solver (solvers/solver.py:167-172 sub-KSPs, :198-262 solve_dpp_nonlinear). */
int dpp_time_cg_block_kernels(dpp_handle h, int field, int warmup, int reps, double* apply_ms, double* update_ms);
/* device milliseconds of the two assembly phases: symbolic (node graph, row pointers, scatter permutation,
 * block pattern; rebuilt from scratch) and numeric (mean of `reps` value fills through the permutation; the
 * bandwidth model is 12 B per stored entry: 8 B value written + 4 B column index of the pattern). */
int dpp_time_assembly(dpp_handle h, int reps, double* symbolic_ms, double* numeric_ms, int64_t* nnz);
int dpp_kernel_launch_count(dpp_handle h, int64_t* launches); /* kernels launched so far by this handle */
/* The launch plan of the plane-streaming apply kernels: number of x-segments chosen for `tiles` in-plane tiles,
 * `planes` owned x-planes, `resident_ctas` co-resident CTAs (2 per SM) and at most `max_ctas` CTAs (host-only,
 * no GPU needed; the cost model and the measurements behind it: DESIGN.md 4.2, csrc/dpp_internal.cuh). */
int dpp_plan_x_segments(int tiles, int planes, int resident_ctas, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* DPP_B200_H */
