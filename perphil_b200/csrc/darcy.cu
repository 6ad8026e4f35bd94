// Darcy velocity recovery u = -k grad(p_h), L2-projected into the vector space V^dim of the pressure's own
// Lagrange space (perphil.utils.postprocessing.calculate_darcy_velocity_from_pressure,
// utils/postprocessing.py:34-63: `fd.project(-conductivity * fd.grad(pressure_field), velocity_space)`).
// SURVEY 8(f) item 3.  A Galerkin projection is one mass solve per component,
//     M u_c = b_c ,   b_c[i] = int phi_i (-k d p_h / d x_c) dx ,
// so the work is (1) the load vectors b_c, assembled here cell by cell with (P+1)^dim Gauss points on the
// multilinear cell geometry (any numbering, any cell shape; exact on parallelepiped cells), and (2) dim
// Jacobi-CG solves with the nodal mass matrix (krylov_mass_solve: the handle's matrix-free operator with
// coefficient block cK = 0, cM = 1 and no Dirichlet rows).
// No atomics: the cell kernel stores its NPC x dim element load vector into a per-cell scratch array and a
// row-owner gather (one thread per node, incident cells in ascending order through the node -> (cell, local)
// adjacency the CSR assembly and the general apply use) adds them in a fixed order: bitwise reproducible
// like every other kernel of the library.
#include <algorithm>
#include <vector>

#include "fe_common.cuh"

namespace dpp {

int general_setup(dpp_context* ctx, const int32_t* cnm);
int krylov_mass_solve(dpp_context* ctx, const double* d_rhs, double* d_out, double rtol, int max_it, int* its,
                      double* rnorm, int* reason);

namespace {

struct GradArgs {
  long long n_cells, n_nodes;
  const int32_t* cnm;
  const int32_t* ccnm;
  const double* coords;
  const double* p;     // [n_nodes] nodal pressure
  double scale;        // -k
  double* cell_out;    // [n_cells][NPC][dim] element load vectors
};

template <int DIM, int P>
__global__ void __launch_bounds__(128) k_grad_load(const GradArgs a) {
  constexpr int P1 = P + 1, NQ = P + 1;
  constexpr int NPC = DIM == 2 ? P1 * P1 : P1 * P1 * P1;
  constexpr int NV = 1 << DIM;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= a.n_cells) return;
  const int32_t* verts = a.ccnm + cell * NV;
  double pe[NPC], be[NPC][DIM];
  for (int b = 0; b < NPC; ++b) {
    pe[b] = a.p[a.cnm[cell * NPC + b]];
    for (int d = 0; d < DIM; ++d) be[b][d] = 0.0;
  }
  constexpr int NQ0 = DIM == 3 ? NQ : 1;
  for (int q0 = 0; q0 < NQ0; ++q0)
    for (int q1 = 0; q1 < NQ; ++q1)
      for (int q2 = 0; q2 < NQ; ++q2) {
        int q[3];   // quadrature index per axis (x slowest)
        if (DIM == 3) { q[0] = q0; q[1] = q1; q[2] = q2; }
        else { q[0] = q1; q[1] = q2; q[2] = 0; }
        double J[3][3];
        jacobian_at<DIM, P>(a.coords, verts, q[0], q[1], q[2], J);
        double det, inv[3][3];
        if (DIM == 2) {
          det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          inv[0][0] = J[1][1] / det; inv[0][1] = -J[0][1] / det;
          inv[1][0] = -J[1][0] / det; inv[1][1] = J[0][0] / det;
        } else {
          inv[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
          inv[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
          inv[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
          inv[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
          inv[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
          inv[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
          inv[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
          inv[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
          inv[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          det = J[0][0] * inv[0][0] + J[0][1] * inv[1][0] + J[0][2] * inv[2][0];
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) inv[r][c] /= det;
        }
        double w = a.scale * fabs(det);
        for (int ax = 0; ax < DIM; ++ax) w *= cW[P - 1][q[ax]];
        // reference gradient of p_h, then the physical one: g_d = sum_ax (d xi_ax / d x_d) g_xi[ax]
        double gxi[3] = {0, 0, 0};
        for (int b = 0; b < NPC; ++b) {
          int l[3];
          if (DIM == 3) { l[0] = b / (P1 * P1); l[1] = (b / P1) % P1; l[2] = b % P1; }
          else { l[0] = b / P1; l[1] = b % P1; l[2] = 0; }
          for (int ax = 0; ax < DIM; ++ax) {
            double dn = 1.0;
            for (int ax2 = 0; ax2 < DIM; ++ax2) dn *= (ax2 == ax) ? cD[P - 1][l[ax2]][q[ax2]] : cB[P - 1][l[ax2]][q[ax2]];
            gxi[ax] += dn * pe[b];
          }
        }
        double gd[3] = {0, 0, 0};
        for (int d = 0; d < DIM; ++d)
          for (int ax = 0; ax < DIM; ++ax) gd[d] += inv[ax][d] * gxi[ax];
        for (int b = 0; b < NPC; ++b) {
          int l[3];
          if (DIM == 3) { l[0] = b / (P1 * P1); l[1] = (b / P1) % P1; l[2] = b % P1; }
          else { l[0] = b / P1; l[1] = b % P1; l[2] = 0; }
          double N = w;
          for (int ax = 0; ax < DIM; ++ax) N *= cB[P - 1][l[ax]][q[ax]];
          for (int d = 0; d < DIM; ++d) be[b][d] = fma(N, gd[d], be[b][d]);
        }
      }
  double* o = a.cell_out + cell * (NPC * DIM);
  for (int b = 0; b < NPC; ++b)
    for (int d = 0; d < DIM; ++d) o[b * DIM + d] = be[b][d];
}

// b_c[node] = sum over the incident cells (ascending cell order) of their element load vector entry
template <int DIM>
__global__ void __launch_bounds__(256) k_grad_gather(long long n_nodes, int npc, const int64_t* __restrict__ adj_ptr,
                                                     const int32_t* __restrict__ adj_cell,
                                                     const uint8_t* __restrict__ adj_loc,
                                                     const double* __restrict__ cell_out, double* __restrict__ out) {
  for (long long node = blockIdx.x * (long long)blockDim.x + threadIdx.x; node < n_nodes;
       node += (long long)gridDim.x * blockDim.x) {
    double acc[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc[d] = 0.0;
    for (long long e = adj_ptr[node]; e < adj_ptr[node + 1]; ++e) {
      const double* o = cell_out + ((long long)adj_cell[e] * npc + adj_loc[e]) * DIM;
#pragma unroll
      for (int d = 0; d < DIM; ++d) acc[d] += o[d];
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) out[d * n_nodes + node] = acc[d];
  }
}

}  // namespace

// d_p: [n_nodes] nodal pressure (device, internal numbering); d_vel: [dim][n_nodes] (device);
// d_rhs: [dim][n_nodes] scratch
int darcy_velocity(dpp_context* ctx, const double* d_p, double conductivity, double rtol, int max_it, double* d_rhs,
                   double* d_vel, int32_t* iterations, double* residuals) {
  DPP_CHECK(fe_upload_tables(ctx));   // this translation unit's copy of the constant tabulations
  const long long n = ctx->n_nodes;
  if (!ctx->general_ready) {  // node -> (cell, local) adjacency lives with the general family
    std::vector<int32_t> cnm((size_t)ctx->n_cells * ctx->npc);
    DPP_CUDA(cudaMemcpy(cnm.data(), ctx->d_cnm, sizeof(int32_t) * cnm.size(), cudaMemcpyDeviceToHost));
    DPP_CHECK(general_setup(ctx, cnm.data()));
  }
  double* cell_out = nullptr;
  DPP_CHECK(dev_alloc(ctx, &cell_out, (int64_t)ctx->n_cells * ctx->npc * ctx->dim));
  GradArgs a{ctx->n_cells, n, ctx->d_cnm, ctx->d_ccnm, ctx->d_coords, d_p, -conductivity, cell_out};
  const int threads = 128;
  const unsigned blocks = (unsigned)((ctx->n_cells + threads - 1) / threads);
  if (ctx->dim == 2 && ctx->degree == 1) k_grad_load<2, 1><<<blocks, threads, 0, ctx->stream>>>(a);
  else if (ctx->dim == 2) k_grad_load<2, 2><<<blocks, threads, 0, ctx->stream>>>(a);
  else if (ctx->degree == 1) k_grad_load<3, 1><<<blocks, threads, 0, ctx->stream>>>(a);
  else k_grad_load<3, 2><<<blocks, threads, 0, ctx->stream>>>(a);
  ctx->launches++;
  const unsigned gblocks = (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16));
  if (ctx->dim == 2)
    k_grad_gather<2><<<gblocks, 256, 0, ctx->stream>>>(n, ctx->npc, ctx->d_adj_ptr, ctx->d_adj_cell, ctx->d_adj_loc, cell_out, d_rhs);
  else
    k_grad_gather<3><<<gblocks, 256, 0, ctx->stream>>>(n, ctx->npc, ctx->d_adj_ptr, ctx->d_adj_cell, ctx->d_adj_loc, cell_out, d_rhs);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(cell_out);
  ctx->device_bytes -= (int64_t)sizeof(double) * ctx->n_cells * ctx->npc * ctx->dim;
  for (int c = 0; c < ctx->dim; ++c) {
    int its = 0, reason = 0;
    double rn = 0.0;
    DPP_CHECK(krylov_mass_solve(ctx, d_rhs + c * n, d_vel + c * n, rtol, max_it, &its, &rn, &reason));
    if (iterations) iterations[c] = its;
    if (residuals) residuals[c] = rn;
    if (reason < 0) {
      ctx->set_error("dpp_darcy_velocity: mass solve of component " + std::to_string(c) + " did not converge (reason " +
                     std::to_string(reason) + ")");
      return DPP_ERR_STATE;
    }
  }
  return DPP_OK;
}

}  // namespace dpp
