// Matrix-free DPP operator for arbitrary node/cell numbering (DPP_KERNEL_GENERAL).
//
// Data model = what the host layer pulls from a Firedrake mesh: vertex
// coordinates + cell->node maps; nothing is assumed about the numbering.  Replaces the TSFC
// element kernels + PyOP2 cell loop + MatMult of the reference path (solver.py:66-71).
//
// Design: row-owner gather.  A precomputed node -> (cell, local index) adjacency lets one thread
// own one output row for both fields; it walks its incident cells in a fixed order, gathers the
// cell's nodal values, and evaluates row `a` of the element operator by sum-factorised Gauss
// quadrature ((P+1)^dim points; affine cells use a per-cell constant metric computed once at
// setup, non-affine cells evaluate the multilinear Jacobian per point).  No atomics, results are
// bitwise reproducible.  The element work is redundant across the rows of a cell; the structured
// family (apply_structured.cu) is the fast path for tensor grids, this family is the general one.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "dpp_internal.cuh"
#include "fe_common.cuh"

namespace dpp {

namespace {

struct GenArgs {
  const int64_t* adj_ptr;
  const int32_t* adj_cell;
  const uint8_t* adj_loc;
  const int32_t* cnm;
  const int32_t* ccnm;
  const double* coords;
  const double* geom;
  const double* x[2];
  double* y[2];
  const uint8_t* in_mask[2];
  const uint8_t* out_mask[2];
  int identity_on_masked;
  Coef c;
  double* dot_partials;
  long long ob, oe;
  const double* skip_flag;
  // diagonal mode
  double* diag;
  const uint8_t* diag_mask;
  long long n_nodes;
};

// per-cell setup: affine test + constant metric (unit weight)
template <int DIM>
__global__ void k_cell_geometry(long long n_cells, const int32_t* __restrict__ ccnm, const double* __restrict__ coords,
                                double* __restrict__ geom) {
  constexpr int NV = 1 << DIM;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_cells;
       c += (long long)gridDim.x * blockDim.x) {
    double X[NV][3];
    for (int v = 0; v < NV; ++v) {
      const long long id = ccnm[c * NV + v];
      for (int d = 0; d < 3; ++d) X[v][d] = d < DIM ? coords[id * DIM + d] : 0.0;
    }
    // edges from vertex 0: local index bit (DIM-1-axis)
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double scale = 0.0;
    for (int ax = 0; ax < DIM; ++ax) {
      const int v = 1 << (DIM - 1 - ax);
      for (int d = 0; d < DIM; ++d) {
        J[d][ax] = X[v][d] - X[0][d];
        scale = fmax(scale, fabs(J[d][ax]));
      }
    }
    bool affine = true;
    for (int v = 0; v < NV; ++v) {
      for (int d = 0; d < DIM; ++d) {
        double e = X[0][d];
        for (int ax = 0; ax < DIM; ++ax)
          if (v & (1 << (DIM - 1 - ax))) e += J[d][ax];
        if (fabs(e - X[v][d]) > 1e-12 * scale) affine = false;
      }
    }
    double G[3][3], dm;
    metric_from_J<DIM>(J, 1.0, G, dm);
    double* g = geom + c * 8;
    g[0] = G[0][0]; g[1] = G[0][1]; g[2] = G[0][2]; g[3] = G[1][1]; g[4] = G[1][2]; g[5] = G[2][2];
    g[6] = dm;
    g[7] = affine ? 1.0 : 0.0;
  }
}

template <int DIM, int P, int NF, bool DIAG>
__global__ void __launch_bounds__(128) k_general(const GenArgs g) {
  constexpr int P1 = P + 1;
  constexpr int NPC = DIM == 2 ? P1 * P1 : P1 * P1 * P1;
  constexpr int NQ = P1;
  constexpr int NV = 1 << DIM;
  constexpr int Q0N = DIM == 2 ? 1 : NQ;  // 2-D: axis 0 of the loops is a dummy
  __shared__ double red[4];
  if (!DIAG && g.skip_flag != nullptr && *g.skip_flag != 0.0) return;
  const long long node = g.ob + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double dot = 0.0;
  if (node < g.oe) {
    double Kx[NF], Mx[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) Kx[f] = Mx[f] = 0.0;
    double Kd = 0.0, Md = 0.0;
    for (long long e = g.adj_ptr[node]; e < g.adj_ptr[node + 1]; ++e) {
      const long long cell = g.adj_cell[e];
      const int a = g.adj_loc[e];
      int a0, a1, a2;
      if (DIM == 2) { a0 = 0; a1 = a / P1; a2 = a % P1; }
      else { a0 = a / (P1 * P1); a1 = (a / P1) % P1; a2 = a % P1; }
      double xe[NF][NPC];
      if (!DIAG) {
#pragma unroll
        for (int b = 0; b < NPC; ++b) {
          const long long nb = g.cnm[cell * NPC + b];
#pragma unroll
          for (int f = 0; f < NF; ++f) {
            double v = g.x[f][nb];
            if (g.in_mask[f] != nullptr && g.in_mask[f][nb]) v = 0.0;
            xe[f][b] = v;
          }
        }
      }
      const double* gm = g.geom + cell * 8;
      const bool affine = gm[7] != 0.0;
      double G[3][3], dm0 = gm[6];
      G[0][0] = gm[0]; G[0][1] = G[1][0] = gm[1]; G[0][2] = G[2][0] = gm[2];
      G[1][1] = gm[3]; G[1][2] = G[2][1] = gm[4]; G[2][2] = gm[5];
      for (int q0 = 0; q0 < Q0N; ++q0)
        for (int q1 = 0; q1 < NQ; ++q1)
          for (int q2 = 0; q2 < NQ; ++q2) {
            const double wq = (DIM == 2 ? 1.0 : cW[P - 1][q0]) * cW[P - 1][q1] * cW[P - 1][q2];
            double Gq[3][3], dm;
            if (affine) {
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Gq[r][c] = G[r][c] * wq;
              dm = dm0 * wq;
            } else {
              double J[3][3];
              jacobian_at<DIM, P>(g.coords, g.ccnm + cell * NV, DIM == 2 ? q1 : q0, DIM == 2 ? q2 : q1, q2, J);
              metric_from_J<DIM>(J, wq, Gq, dm);
            }
            // test function a at q (reference gradient; xi-axis order = (0,1,2); 2-D uses axes (1,2)->(0,1))
            const double Ba0 = DIM == 2 ? 1.0 : cB[P - 1][a0][q0], Da0 = DIM == 2 ? 0.0 : cD[P - 1][a0][q0];
            const double Ba1 = cB[P - 1][a1][q1], Da1 = cD[P - 1][a1][q1];
            const double Ba2 = cB[P - 1][a2][q2], Da2 = cD[P - 1][a2][q2];
            double ta = Ba0 * Ba1 * Ba2;
            double gta[3];
            if (DIM == 2) { gta[0] = Da1 * Ba2; gta[1] = Ba1 * Da2; gta[2] = 0.0; }
            else { gta[0] = Da0 * Ba1 * Ba2; gta[1] = Ba0 * Da1 * Ba2; gta[2] = Ba0 * Ba1 * Da2; }
            // G gta (symmetric)
            double Gt[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) Gt[r] = Gq[r][0] * gta[0] + Gq[r][1] * gta[1] + Gq[r][2] * gta[2];
            if (DIAG) {
              Kd += Gt[0] * gta[0] + Gt[1] * gta[1] + Gt[2] * gta[2];
              Md += ta * ta * dm;
            } else {
#pragma unroll
              for (int f = 0; f < NF; ++f) {
                double u = 0.0, gu0 = 0.0, gu1 = 0.0, gu2 = 0.0;
#pragma unroll
                for (int b0 = 0; b0 < (DIM == 2 ? 1 : P1); ++b0) {
                  const double B0 = DIM == 2 ? 1.0 : cB[P - 1][b0][q0], D0 = DIM == 2 ? 0.0 : cD[P - 1][b0][q0];
#pragma unroll
                  for (int b1 = 0; b1 < P1; ++b1) {
                    const double B1 = cB[P - 1][b1][q1], D1 = cD[P - 1][b1][q1];
#pragma unroll
                    for (int b2 = 0; b2 < P1; ++b2) {
                      const double B2 = cB[P - 1][b2][q2], D2 = cD[P - 1][b2][q2];
                      const double xv = xe[f][(b0 * P1 + b1) * P1 + b2];
                      u = fma(B0 * B1 * B2, xv, u);
                      if (DIM == 2) {
                        gu0 = fma(D1 * B2, xv, gu0);
                        gu1 = fma(B1 * D2, xv, gu1);
                      } else {
                        gu0 = fma(D0 * B1 * B2, xv, gu0);
                        gu1 = fma(B0 * D1 * B2, xv, gu1);
                        gu2 = fma(B0 * B1 * D2, xv, gu2);
                      }
                    }
                  }
                }
                Kx[f] += Gt[0] * gu0 + Gt[1] * gu1 + Gt[2] * gu2;
                Mx[f] += ta * dm * u;
              }
            }
          }
    }
    if (DIAG) {
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        double d = g.c.cK[f][f] * Kd + g.c.cM[f][f] * Md;
        if (g.diag_mask != nullptr && g.diag_mask[f * g.n_nodes + node]) d = 1.0;
        g.diag[f * g.n_nodes + node] = d;
      }
    } else {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        double yv = 0.0;
#pragma unroll
        for (int h = 0; h < NF; ++h) {
          yv = fma(g.c.cK[f][h], Kx[h], yv);
          yv = fma(g.c.cM[f][h], Mx[h], yv);
        }
        double xc = g.x[f][node];
        if (g.out_mask[f] != nullptr && g.out_mask[f][node]) yv = g.identity_on_masked ? xc : 0.0;
        g.y[f][node] = yv;
        dot = fma(xc, yv, dot);
      }
    }
  }
  if (!DIAG && g.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) g.dot_partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
  }
}

template <bool DIAG>
int launch_general(dpp_context* ctx, const GenArgs& g, int nf, int blocks) {
  const int dim = ctx->dim, p = ctx->degree;
#define GEN_CASE(D, PP)                                                                  \
  if (dim == D && p == PP) {                                                             \
    if (DIAG || nf == 2) k_general<D, PP, 2, DIAG><<<blocks, 128, 0, ctx->stream>>>(g);  \
    else k_general<D, PP, 1, DIAG><<<blocks, 128, 0, ctx->stream>>>(g);                  \
  }
  GEN_CASE(2, 1) GEN_CASE(2, 2) GEN_CASE(3, 1) GEN_CASE(3, 2)
#undef GEN_CASE
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

void fill_common(const dpp_context* ctx, GenArgs& g) {
  g.adj_ptr = ctx->d_adj_ptr;
  g.adj_cell = ctx->d_adj_cell;
  g.adj_loc = ctx->d_adj_loc;
  g.cnm = ctx->d_cnm;
  g.ccnm = ctx->d_ccnm;
  g.coords = ctx->d_coords;
  g.geom = ctx->d_cell_geom;
  g.n_nodes = ctx->n_nodes;
}

}  // namespace

int general_setup(dpp_context* ctx, const int32_t* cnm) {
  if (ctx->general_ready) return DPP_OK;
  const int64_t n = ctx->n_nodes, nc = ctx->n_cells;
  const int npc = ctx->npc;
  // node -> (cell, local) adjacency, cells in ascending order per node (deterministic)
  std::vector<int64_t> ptr((size_t)n + 1, 0);
  for (int64_t c = 0; c < nc; ++c)
    for (int a = 0; a < npc; ++a) {
      const int32_t v = cnm[c * npc + a];
      if (v < 0 || v >= n) {
        ctx->set_error("cell_node_map entry out of range");
        return DPP_ERR_INVALID;
      }
      ptr[(size_t)v + 1]++;
    }
  for (int64_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  std::vector<int32_t> acell((size_t)ptr[n]);
  std::vector<uint8_t> aloc((size_t)ptr[n]);
  {
    std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int a = 0; a < npc; ++a) {
        const int64_t pos = cur[cnm[c * npc + a]]++;
        acell[pos] = (int32_t)c;
        aloc[pos] = (uint8_t)a;
      }
  }
  DPP_CHECK(dev_alloc(ctx, &ctx->d_adj_ptr, n + 1));
  DPP_CHECK(dev_alloc(ctx, &ctx->d_adj_cell, ptr[n]));
  DPP_CHECK(dev_alloc(ctx, &ctx->d_adj_loc, ptr[n]));
  DPP_CUDA(cudaMemcpy(ctx->d_adj_ptr, ptr.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice));
  DPP_CUDA(cudaMemcpy(ctx->d_adj_cell, acell.data(), sizeof(int32_t) * ptr[n], cudaMemcpyHostToDevice));
  DPP_CUDA(cudaMemcpy(ctx->d_adj_loc, aloc.data(), sizeof(uint8_t) * ptr[n], cudaMemcpyHostToDevice));
  DPP_CHECK(fe_upload_tables(ctx));
  // per-cell geometry
  DPP_CHECK(dev_alloc(ctx, &ctx->d_cell_geom, nc * 8));
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 127) / 128, (int64_t)ctx->sm_count * 32));
  if (ctx->dim == 2)
    k_cell_geometry<2><<<blocks, 128, 0, ctx->stream>>>(nc, ctx->d_ccnm, ctx->d_coords, ctx->d_cell_geom);
  else
    k_cell_geometry<3><<<blocks, 128, 0, ctx->stream>>>(nc, ctx->d_ccnm, ctx->d_coords, ctx->d_cell_geom);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->general_ready = true;
  return DPP_OK;
}

int general_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks) {
  if (!ctx->general_ready) {
    ctx->set_error("general kernel family not set up");
    return DPP_ERR_STATE;
  }
  // Q1 hexahedra on a whole (unpartitioned) mesh: the element-based cell-block kernel (apply_cells.cu); the
  // row-owner gather below serves 2-D / Q2 meshes (DPP_GENERAL_ROW_OWNER=1 forces it: parity tests compare both)
  if (cells_supported(ctx) && a.owned_begin == 0 && a.owned_end == ctx->n_nodes && getenv("DPP_GENERAL_ROW_OWNER") == nullptr) {
    DPP_CHECK(cells_setup(ctx));
    return cells_apply(ctx, a, n_partial_blocks);
  }
  GenArgs g{};
  fill_common(ctx, g);
  for (int f = 0; f < 2; ++f) {
    g.x[f] = a.x[f]; g.y[f] = a.y[f]; g.in_mask[f] = a.in_mask[f]; g.out_mask[f] = a.out_mask[f];
  }
  g.identity_on_masked = a.identity_on_masked;
  g.c = a.c;
  g.dot_partials = a.dot_partials;
  g.ob = a.owned_begin;
  g.oe = a.owned_end;
  g.skip_flag = a.skip_flag;
  const long long nown = g.oe - g.ob;
  const int blocks = (int)((nown + 127) / 128);
  if (a.dot_partials != nullptr && blocks > kMaxPartialBlocks * kMaxDotWidth) {
    ctx->set_error("general apply: reduction scratch too small");
    return DPP_ERR_INVALID;
  }
  if (blocks > 0) DPP_CHECK(launch_general<false>(ctx, g, a.nf, blocks));
  if (n_partial_blocks) *n_partial_blocks = blocks;
  return DPP_OK;
}

int general_diagonal(dpp_context* ctx, const Coef& c, double* d_diag) {
  if (!ctx->general_ready) {
    ctx->set_error("general kernel family not set up");
    return DPP_ERR_STATE;
  }
  GenArgs g{};
  fill_common(ctx, g);
  g.c = c;
  g.diag = d_diag;
  g.diag_mask = ctx->d_mask;
  g.ob = 0;
  g.oe = ctx->n_nodes;
  const int blocks = (int)((ctx->n_nodes + 127) / 128);
  return launch_general<true>(ctx, g, 2, blocks);
}

}  // namespace dpp
