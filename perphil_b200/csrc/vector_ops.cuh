// Fused BLAS-1 kernels of the Krylov loops (PETSc VecAXPY/VecAYPX/VecDot/VecNorm/VecMDot/VecMAXPY
// on the path).  All reductions are two-stage and deterministic: per-block
// partial sums in a fixed order (warp shuffle tree + fixed smem order), then one block sums the
// partials in index order.  Scalars (alpha, beta, norms, iteration counter, converged reason) live
// in a device scalar block so that a whole CG iteration is launched without host round trips.
#pragma once

#include "dpp_internal.cuh"

namespace dpp {

// layout of one solver slot in ctx->d_scalars (doubles)
enum {
  S_RZ = 0, S_RZ_OLD, S_PAP, S_ZZ, S_TTOL, S_RNORM0, S_ITS, S_REASON, S_RNORM, S_DTOL, S_MAXIT, S_ATOL,
  S_RTOL, S_HISTCAP, S_ALPHA /* rz / pAp of the current iteration */, S_XPEND /* x += alpha p not applied yet */, S_TMP /* 40 doubles of reduction output */, S_SLOT_SIZE = 64
};

enum PostOp { POST_NONE = 0, POST_CG_INIT = 1, POST_CG_PAP = 2, POST_CG_RZ = 3 };

// vectors are field-blocked with stride `stride` (= local n_nodes); only rows [ob, oe) of each of
// the nf fields are touched.
struct VecLayout {
  int nf;
  int64_t stride, ob, oe;
};

int vec_launch_blocks(const dpp_context* ctx, const VecLayout& L);

// y = a*x + b*y over the layout (a, b host scalars)
int vec_axpby(dpp_context* ctx, const VecLayout& L, double a, const double* x, double b, double* y);
// z = dinv .* r (dinv may be null -> copy)
int vec_pointwise_mult(dpp_context* ctx, const VecLayout& L, const double* dinv, const double* r, double* z);
// z = B^-1 r with 2x2 nodal blocks (inv00, inv01, inv11) ; nf must be 2
int vec_pbjacobi(dpp_context* ctx, const VecLayout& L, const double* inv00, const double* inv01,
                 const double* inv11, const double* r, double* z);
// partial sums of up to two dots: out0 = <a0,b0>, out1 = <a1,b1>; reduced into S[S_TMP..] of `slot`
int vec_dot2(dpp_context* ctx, const VecLayout& L, const double* a0, const double* b0, const double* a1,
             const double* b1, int slot, PostOp post);
// reduce ctx->d_partials[nblocks*width] -> S[S_TMP + w] (+ allreduce when distributed) then post op
int reduce_partials(dpp_context* ctx, int nblocks, int width, int slot, PostOp post, int out_offset = 0);
constexpr int kGmresNormOffset = 36;  // S[S_TMP + 36] holds ||w||^2 after gmres_maxpy_norm

// CG fused kernels (device scalars of `slot`)
int cg_p_update(dpp_context* ctx, const VecLayout& L, double* p, const double* r, const double* dinv,
                const double* z, int slot);
int cg_xr_update(dpp_context* ctx, const VecLayout& L, double* x, double* r, const double* p, const double* w,
                 const double* dinv, bool fused_pc, int slot, PostOp post);

// fused CG iteration on uniform grids (cg_fused_uniform.cu): padded private layout + TMA loads
bool cg_fused_available(dpp_context* ctx, int nf, int operator_mode, int pc_type);
// (the reductions + PETSc bookkeeping after each kernel are folded into the kernels, or run as
//  reduce_partials + NCCL when neither single-GPU nor peer-memory)
int cg_fused_variant(dpp_context* ctx);   // 0: x updated in the apply kernel; 1: deferred x update (direction ring); 2: + residual update that recomputes A p
int cg_fused_table(dpp_context* ctx, const Coef& c, int nf, int pc_type, const int* fld, double* d_tab);
int cg_fused_begin(dpp_context* ctx, int nf, const double* b);
int cg_fused_rz_init(dpp_context* ctx, int nf, const int* fld, int slot, const double* dtab);
int cg_fused_apply(dpp_context* ctx, int nf, const Coef& c, long long it, const int* fld, int slot, const double* dtab);
int cg_fused_r_update(dpp_context* ctx, int nf, const Coef& c, long long it, const int* fld, int slot, const double* dtab);
int cg_fused_x_finalize(dpp_context* ctx, int nf, long long its, int slot, double* x);
int cg_fused_halo_r(dpp_context* ctx, int nf, bool after_update, int slot);
int cg_fused_plain_apply(dpp_context* ctx, int nf, const Coef& c, bool want_dot, int* n_partial_blocks);
int cg_fused_pad_from(dpp_context* ctx, int which /*0 r, 1 p0, 2 p1, 3 w, 4 x*/, const double* src);

// GMRES kernels
int gmres_mdot(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv, const double* w, int slot,
               const double* skip = nullptr);
int gmres_maxpy_norm(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv, double* w, int slot,
                     const double* skip = nullptr);
int vec_maxpy_dev(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv_max, const double* coef_dev,
                  const double* nv_dev, double* x);
int vec_scale_dev(dpp_context* ctx, const VecLayout& L, double* v, const double* factor_dev, const double* skip0,
                  const double* skip1);
// x += sum_j coef[j] V_j  (coef: host array, nv <= 32)
int vec_maxpy_host(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv, const double* coef,
                   double* x);
int vec_scale_into(dpp_context* ctx, const VecLayout& L, double a, const double* x, double* y);  // y = a x

// host <-> device scalar block
int scalars_fetch(dpp_context* ctx, int slot);  // sync: ctx->h_scalars[slot*S_SLOT_SIZE ..] <- device
int scalars_init(dpp_context* ctx, int slot, double rtol, double atol, double dtol, int max_it, int hist_cap);
double* hist_device(dpp_context* ctx, int slot);

}  // namespace dpp
