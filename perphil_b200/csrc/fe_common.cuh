// Finite-element building blocks shared by the general matrix-free apply (apply_general.cu) and
// the CSR assembly (assemble_csr.cu): 1-D Lagrange tabulations at Gauss points in constant memory,
// the Jacobian of the multilinear cell map and the metric G = |J| J^-1 J^-T w_q.
// Every translation unit that includes this header owns a private copy of the constant tables and
// must call fe_upload_tables() once before launching kernels that use them.
#pragma once

#include <cmath>

#include "dpp_internal.cuh"

namespace dpp {
namespace {

// 1-D tabulations at the NQ = P+1 Gauss points of [0,1]: pressure basis (degree P) and the
// linear geometry basis.  Index [P-1][a][q].
static __constant__ double cB[2][3][3];
static __constant__ double cD[2][3][3];
static __constant__ double cBg[2][2][3];
static __constant__ double cDg[2][2][3];
static __constant__ double cW[2][3];

template <int DIM>
__device__ __forceinline__ void metric_from_J(const double (&J)[3][3], double wq, double (&G)[3][3], double& dm) {
  if (DIM == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    // Jinv
    const double a = J[1][1] * id, b = -J[0][1] * id, c = -J[1][0] * id, d = J[0][0] * id;
    const double s = fabs(det) * wq;
    // G = s * Jinv Jinv^T   (Jinv[xi][x])
    G[0][0] = s * (a * a + b * b);
    G[0][1] = G[1][0] = s * (a * c + b * d);
    G[1][1] = s * (c * c + d * d);
    G[0][2] = G[2][0] = G[1][2] = G[2][1] = G[2][2] = 0.0;
    dm = s;
  } else {
    double inv[3][3];
    inv[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    inv[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    inv[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    inv[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    inv[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    inv[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    inv[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    inv[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    inv[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * inv[0][0] + J[0][1] * inv[1][0] + J[0][2] * inv[2][0];
    const double id = 1.0 / det;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) inv[r][c] *= id;
    const double s = fabs(det) * wq;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        G[r][c] = s * (inv[r][0] * inv[c][0] + inv[r][1] * inv[c][1] + inv[r][2] * inv[c][2]);
    dm = s;
  }
}

// J[x][xi] at quadrature point (q0,q1,q2) of the multilinear map through the cell's vertices
template <int DIM, int P>
__device__ __forceinline__ void jacobian_at(const double* __restrict__ coords, const int32_t* __restrict__ verts,
                                            int q0, int q1, int q2, double (&J)[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) J[r][c] = 0.0;
  if (DIM == 2) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const long long v = verts[a * 2 + b];
        const double X = coords[v * 2], Y = coords[v * 2 + 1];
        const double d0 = cDg[P - 1][a][q0] * cBg[P - 1][b][q1];
        const double d1 = cBg[P - 1][a][q0] * cDg[P - 1][b][q1];
        J[0][0] += X * d0; J[0][1] += X * d1;
        J[1][0] += Y * d0; J[1][1] += Y * d1;
      }
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const long long v = verts[a * 4 + b * 2 + c];
          const double X = coords[v * 3], Y = coords[v * 3 + 1], Z = coords[v * 3 + 2];
          const double d0 = cDg[P - 1][a][q0] * cBg[P - 1][b][q1] * cBg[P - 1][c][q2];
          const double d1 = cBg[P - 1][a][q0] * cDg[P - 1][b][q1] * cBg[P - 1][c][q2];
          const double d2 = cBg[P - 1][a][q0] * cBg[P - 1][b][q1] * cDg[P - 1][c][q2];
          J[0][0] += X * d0; J[0][1] += X * d1; J[0][2] += X * d2;
          J[1][0] += Y * d0; J[1][1] += Y * d1; J[1][2] += Y * d2;
          J[2][0] += Z * d0; J[2][1] += Z * d1; J[2][2] += Z * d2;
        }
  }
}

inline void tabulate(int p, double B[3][3], double D[3][3], double Bg[2][3], double Dg[2][3], double W[3]) {
  const int nq = p + 1;
  double xq[3], wq[3];
  if (nq == 2) {
    const double s = 1.0 / std::sqrt(3.0);
    xq[0] = 0.5 * (1 - s); xq[1] = 0.5 * (1 + s); wq[0] = wq[1] = 0.5;
  } else {
    const double s = std::sqrt(0.6);
    xq[0] = 0.5 * (1 - s); xq[1] = 0.5; xq[2] = 0.5 * (1 + s);
    wq[0] = wq[2] = 5.0 / 18.0; wq[1] = 8.0 / 18.0;
  }
  for (int a = 0; a < 3; ++a)
    for (int q = 0; q < 3; ++q) B[a][q] = D[a][q] = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double x = xq[q];
    W[q] = wq[q];
    Bg[0][q] = 1 - x; Bg[1][q] = x; Dg[0][q] = -1; Dg[1][q] = 1;
    if (p == 1) {
      B[0][q] = 1 - x; B[1][q] = x; D[0][q] = -1; D[1][q] = 1;
    } else {
      B[0][q] = 2 * (x - 0.5) * (x - 1); B[1][q] = -4 * x * (x - 1); B[2][q] = 2 * x * (x - 0.5);
      D[0][q] = 4 * x - 3; D[1][q] = -8 * x + 4; D[2][q] = 4 * x - 1;
    }
  }
}


inline int fe_upload_tables(dpp_context* ctx) {
  double B[2][3][3] = {}, D[2][3][3] = {}, Bg[2][2][3] = {}, Dg[2][2][3] = {}, W[2][3] = {};
  for (int p = 1; p <= 2; ++p) tabulate(p, B[p - 1], D[p - 1], Bg[p - 1], Dg[p - 1], W[p - 1]);
  DPP_CUDA(cudaMemcpyToSymbol(cB, B, sizeof(B)));
  DPP_CUDA(cudaMemcpyToSymbol(cD, D, sizeof(D)));
  DPP_CUDA(cudaMemcpyToSymbol(cBg, Bg, sizeof(Bg)));
  DPP_CUDA(cudaMemcpyToSymbol(cDg, Dg, sizeof(Dg)));
  DPP_CUDA(cudaMemcpyToSymbol(cW, W, sizeof(W)));
  return DPP_OK;
}

}  // namespace
}  // namespace dpp
