// Error norms of a DPP solution: ||p_h - p||_L2 and |p_h - p|_H1 per field, by Gauss quadrature over the
// cells (perphil.utils.postprocessing.l2_error / h1_seminorm_error, utils/postprocessing.py:89-124 --
// there `fd.assemble(inner(diff, diff) * dx)`; the numbers stored in
// notebooks/results-conforming-2d/convergence.csv are produced by it).  SURVEY 8(f) item 1.
//
// The exact solution is either the manufactured closed form (utils/manufactured_solutions.py:39-51, 82-88,
// evaluated on the device together with its gradient) or a nodal field of the same space.  One thread
// per cell, multilinear geometry from the vertex coordinates (any numbering, any cell shape), nq^dim
// Gauss points; per-block partial sums in a fixed order, summed on the host in block order.
#include <cmath>
#include <vector>

#include "dpp_internal.cuh"

namespace dpp {

namespace {

constexpr int kMaxQ = 8;

struct ErrArgs {
  int dim, degree, nq;
  long long n_cells, n_nodes;
  const int32_t* cnm;
  const int32_t* ccnm;
  const double* coords;
  const double* u;       // [2n]
  const double* exact;   // [2n] nodal exact field (kind 0) or null (kind 1: manufactured closed form)
  double k1, k2, beta, mu, eta;
  double xq[kMaxQ], wq[kMaxQ];
  double B[3][kMaxQ], D[3][kMaxQ];   // pressure basis (degree P) at the Gauss points of [0,1]
  double* partials;      // [nblocks][4]
};

// manufactured pressures and gradients at X
__device__ __forceinline__ void manufactured(const ErrArgs& a, const double* X, double* p, double (*g)[3]) {
  const double pi = 3.14159265358979323846;
  const double ex = exp(pi * X[0]);
  double s, ey[2] = {0, 0}, c[2] = {0, 0};
  if (a.dim == 2) {
    s = sin(pi * X[1]);
    c[0] = cos(pi * X[1]);
    ey[0] = exp(a.eta * X[1]);
  } else {
    s = sin(pi * X[1]) + sin(pi * X[2]);
    c[0] = cos(pi * X[1]);
    c[1] = cos(pi * X[2]);
    ey[0] = exp(a.eta * X[1]);
    ey[1] = exp(a.eta * X[2]);
  }
  const double common = (a.mu / pi) * ex * s;
  const double f1 = -a.mu / (a.beta * a.k1), f2 = a.mu / (a.beta * a.k2);
  p[0] = common + f1 * (ey[0] + ey[1]);
  p[1] = common + f2 * (ey[0] + ey[1]);
  for (int f = 0; f < 2; ++f) {
    const double ff = f == 0 ? f1 : f2;
    g[f][0] = a.mu * ex * s;
    g[f][1] = a.mu * ex * c[0] + ff * a.eta * ey[0];
    g[f][2] = a.dim == 3 ? a.mu * ex * c[1] + ff * a.eta * ey[1] : 0.0;
  }
}

template <int DIM, int P>
__global__ void __launch_bounds__(128) k_error_norms(const ErrArgs a) {
  constexpr int P1 = P + 1;
  constexpr int NPC = DIM == 2 ? P1 * P1 : P1 * P1 * P1;
  constexpr int NV = 1 << DIM;
  __shared__ double red[4][4];
  double acc[4] = {0, 0, 0, 0};
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell < a.n_cells) {
    double XV[NV][DIM];
    for (int v = 0; v < NV; ++v)
      for (int d = 0; d < DIM; ++d) XV[v][d] = a.coords[(long long)a.ccnm[cell * NV + v] * DIM + d];
    double ue[2][NPC], ee[2][NPC];
    for (int b = 0; b < NPC; ++b) {
      const long long nb = a.cnm[cell * NPC + b];
      for (int f = 0; f < 2; ++f) {
        ue[f][b] = a.u[f * a.n_nodes + nb];
        ee[f][b] = a.exact ? a.exact[f * a.n_nodes + nb] : 0.0;
      }
    }
    const int nq = a.nq;
    const int nq0 = DIM == 3 ? nq : 1;
    for (int q0 = 0; q0 < nq0; ++q0)
      for (int q1 = 0; q1 < nq; ++q1)
        for (int q2 = 0; q2 < nq; ++q2) {
          // reference point: axes (x, y[, z]) <-> local tensor index, x slowest
          int q[3];
          if (DIM == 3) { q[0] = q0; q[1] = q1; q[2] = q2; }
          else { q[0] = q1; q[1] = q2; q[2] = 0; }
          double w = 1.0;
          for (int d = 0; d < DIM; ++d) w *= a.wq[q[d]];
          // geometry: multilinear map, J[d][ax] = d x_d / d xi_ax
          double X[3] = {0, 0, 0}, J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
          for (int v = 0; v < NV; ++v) {
            double N = 1.0, dN[3] = {1.0, 1.0, 1.0};
            for (int ax = 0; ax < DIM; ++ax) {
              const int bit = (v >> (DIM - 1 - ax)) & 1;
              const double xi = a.xq[q[ax]];
              const double n = bit ? xi : 1.0 - xi, dn = bit ? 1.0 : -1.0;
              N *= n;
              for (int ax2 = 0; ax2 < DIM; ++ax2) dN[ax2] *= (ax2 == ax) ? dn : n;
            }
            for (int d = 0; d < DIM; ++d) {
              X[d] += N * XV[v][d];
              for (int ax = 0; ax < DIM; ++ax) J[d][ax] += dN[ax] * XV[v][d];
            }
          }
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
          // discrete fields and reference gradients
          double uh[2] = {0, 0}, eh[2] = {0, 0}, gu[2][3] = {{0, 0, 0}, {0, 0, 0}}, ge[2][3] = {{0, 0, 0}, {0, 0, 0}};
          for (int b = 0; b < NPC; ++b) {
            int l[3];
            if (DIM == 3) { l[0] = b / (P1 * P1); l[1] = (b / P1) % P1; l[2] = b % P1; }
            else { l[0] = b / P1; l[1] = b % P1; l[2] = 0; }
            double N = 1.0, dN[3] = {1.0, 1.0, 1.0};
            for (int ax = 0; ax < DIM; ++ax) {
              const double n = a.B[l[ax]][q[ax]], dn = a.D[l[ax]][q[ax]];
              N *= n;
              for (int ax2 = 0; ax2 < DIM; ++ax2) dN[ax2] *= (ax2 == ax) ? dn : n;
            }
            for (int f = 0; f < 2; ++f) {
              uh[f] += N * ue[f][b];
              eh[f] += N * ee[f][b];
              for (int ax = 0; ax < DIM; ++ax) {
                gu[f][ax] += dN[ax] * ue[f][b];
                ge[f][ax] += dN[ax] * ee[f][b];
              }
            }
          }
          double pe[2], gpe[2][3];
          if (a.exact == nullptr) manufactured(a, X, pe, gpe);
          const double jw = fabs(det) * w;
          for (int f = 0; f < 2; ++f) {
            // physical gradient = J^-T grad_xi : g_d = sum_ax inv[ax][d] * gxi[ax]
            double gd[3] = {0, 0, 0}, gx[3] = {0, 0, 0};
            for (int d = 0; d < DIM; ++d)
              for (int ax = 0; ax < DIM; ++ax) {
                gd[d] += inv[ax][d] * gu[f][ax];
                gx[d] += inv[ax][d] * ge[f][ax];
              }
            const double dv = uh[f] - (a.exact ? eh[f] : pe[f]);
            acc[f] += jw * dv * dv;
            double s = 0.0;
            for (int d = 0; d < DIM; ++d) {
              const double dg = gd[d] - (a.exact ? gx[d] : gpe[f][d]);
              s += dg * dg;
            }
            acc[2 + f] += jw * s;
          }
        }
  }
  for (int v = 0; v < 4; ++v) {
    double t = acc[v];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][v] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] =
      red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

void gauss_legendre01(int n, double* x, double* w) {  // Newton on P_n, mapped to [0,1]
  for (int i = 0; i < n; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-15) break;
    }
    x[n - 1 - i] = 0.5 * (z + 1.0);
    w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}

}  // namespace

int error_norms(dpp_context* ctx, const double* d_u, const double* d_exact, int nq, double out[4]) {
  if (nq < 1 || nq > kMaxQ) {
    ctx->set_error("dpp_error_norms: 1 <= nq <= 8");
    return DPP_ERR_INVALID;
  }
  ErrArgs a{};
  a.dim = ctx->dim; a.degree = ctx->degree; a.nq = nq;
  a.n_cells = ctx->n_cells; a.n_nodes = ctx->n_nodes;
  a.cnm = ctx->d_cnm; a.ccnm = ctx->d_ccnm; a.coords = ctx->d_coords;
  a.u = d_u; a.exact = d_exact;
  a.k1 = ctx->k1; a.k2 = ctx->k2; a.beta = ctx->beta; a.mu = ctx->mu;
  a.eta = std::sqrt(ctx->beta * (ctx->k1 + ctx->k2) / (ctx->k1 * ctx->k2));  // models/dpp/parameters.py:52
  gauss_legendre01(nq, a.xq, a.wq);
  const int p = ctx->degree;
  for (int q = 0; q < nq; ++q) {
    const double x = a.xq[q];
    if (p == 1) {
      a.B[0][q] = 1 - x; a.B[1][q] = x;
      a.D[0][q] = -1; a.D[1][q] = 1;
    } else {  // equispaced quadratic Lagrange basis on nodes 0, 1/2, 1
      a.B[0][q] = 2 * (x - 0.5) * (x - 1); a.B[1][q] = -4 * x * (x - 1); a.B[2][q] = 2 * x * (x - 0.5);
      a.D[0][q] = 4 * x - 3; a.D[1][q] = -8 * x + 4; a.D[2][q] = 4 * x - 1;
    }
  }
  const int threads = 128;
  const long long nblocks = (ctx->n_cells + threads - 1) / threads;
  double* d_part = nullptr;
  DPP_CUDA(cudaMalloc((void**)&d_part, sizeof(double) * 4 * nblocks));
  a.partials = d_part;
  if (ctx->dim == 2 && p == 1) k_error_norms<2, 1><<<(unsigned)nblocks, threads, 0, ctx->stream>>>(a);
  else if (ctx->dim == 2) k_error_norms<2, 2><<<(unsigned)nblocks, threads, 0, ctx->stream>>>(a);
  else if (p == 1) k_error_norms<3, 1><<<(unsigned)nblocks, threads, 0, ctx->stream>>>(a);
  else k_error_norms<3, 2><<<(unsigned)nblocks, threads, 0, ctx->stream>>>(a);
  ctx->launches++;
  std::vector<double> h((size_t)4 * nblocks);
  cudaError_t e = cudaMemcpyAsync(h.data(), d_part, sizeof(double) * 4 * nblocks, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_part);
  if (e != cudaSuccess) {
    ctx->set_error(std::string("dpp_error_norms: ") + cudaGetErrorString(e));
    return DPP_ERR_CUDA;
  }
  for (int v = 0; v < 4; ++v) out[v] = 0.0;
  for (long long b = 0; b < nblocks; ++b)
    for (int v = 0; v < 4; ++v) out[v] += h[(size_t)b * 4 + v];
  return DPP_OK;
}

}  // namespace dpp
