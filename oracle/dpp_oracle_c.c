/*
 * dpp_oracle_c.c -- CPU restatement of perphil's DPP hot path in plain C + OpenMP.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library (through oracle/c_oracle.py), and
 * only as the checker or as the timed CPU baseline.  perphil_b200 never links or loads it.
 *
 * What it restates (file:line relative to /root/reference/src/perphil; the arithmetic itself
 * runs in Firedrake 2025.10.2 / PETSc 3.23-3.24, which are not vendored and not installable here):
 *   forms/dpp.py:27,57,89         A = (1/mu) [[k1 K + beta M, -beta M], [-beta M, k2 K + beta M]]
 *   solvers/conditioning.py:62    fd.assemble(a, bcs=bcs, mat_type="aij"): AIJ (CSR) matrix, Dirichlet
 *                                 rows and columns zeroed, unit diagonal
 *   solvers/solver.py:66-74       LinearVariationalSolver.solve(): lifting b = -(A u0) on interior rows,
 *                                 KSP from a zero initial guess, its / residual norm read back
 *   solvers/parameters.py:1,12-18 rtol 1e-8, atol 1e-12, max_it 50000
 * PETSc semantics restated: MatMult on SeqAIJ (row-wise CSR SpMV), KSPCG with the preconditioned
 * norm, PCJACOBI, KSPConvergedDefault.
 *
 * Pinning: this file is checked against oracle/dpp_oracle.py (itself pinned to the reference's
 * stored CSV / notebook numbers, tests/test_oracle_golden.py) in tests/test_oracle_c.py: CSR
 * pattern bit-exact, values to 1e-13 relative, identical CG iteration counts and histories.
 *
 * Meshes: uniform tensor grids of Q1/Q2 quads/hexes with lexicographic node numbering (x slowest),
 * the layout of every BASELINE.json configuration.  Element matrices are the Kronecker products of
 * the 1-D matrices of SURVEY A.2.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int dim, degree;
  int n[3];          /* nodes per axis (axis 0 = x slowest; 2-D: n[0] = 1) */
  int64_t n_nodes, n_dof, nnz;
  int64_t* indptr;   /* [n_dof + 1] */
  int32_t* indices;  /* [nnz] sorted per row */
  double* data;      /* [nnz]  A with Dirichlet rows/cols eliminated (explicit zeros kept) */
  double* b;         /* lifted right-hand side */
  double* u0;        /* Dirichlet lift */
  double* diag;
  uint8_t* mask;     /* [n_dof] 1 = constrained */
  double* m1d[3];    /* assembled 1-D mass, band storage [n][2p+1] */
  double* k1d[3];
} orc_system;

static void tables_1d(int ncell, int p, double h, double** m_out, double** k_out) {
  const int n = p * ncell + 1, w = 2 * p + 1;
  double* m = (double*)calloc((size_t)n * w, sizeof(double));
  double* k = (double*)calloc((size_t)n * w, sizeof(double));
  if (ncell == 0) {
    m[p] = 1.0; /* dummy axis of a 2-D mesh */
  } else {
    double Ke[3][3], Me[3][3];
    if (p == 1) {
      const double a = 1.0 / h, b = h / 6.0;
      Ke[0][0] = a; Ke[0][1] = -a; Ke[1][0] = -a; Ke[1][1] = a;
      Me[0][0] = 2 * b; Me[0][1] = b; Me[1][0] = b; Me[1][1] = 2 * b;
    } else {
      static const double K2[3][3] = {{7, -8, 1}, {-8, 16, -8}, {1, -8, 7}};
      static const double M2[3][3] = {{4, 2, -1}, {2, 16, 2}, {-1, 2, 4}};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          Ke[r][c] = K2[r][c] / (3.0 * h);
          Me[r][c] = M2[r][c] * (h / 30.0);
        }
    }
    for (int e = 0; e < ncell; ++e)
      for (int r = 0; r <= p; ++r)
        for (int c = 0; c <= p; ++c) {
          const int row = p * e + r;
          m[(size_t)row * w + p + (c - r)] += Me[r][c];
          k[(size_t)row * w + p + (c - r)] += Ke[r][c];
        }
  }
  *m_out = m;
  *k_out = k;
}

/* coupled 1-D neighbours of node a on an axis with n nodes, degree p: [lo, hi] */
static inline void nbr_range(int a, int n, int p, int* lo, int* hi) {
  if (n == 1) { *lo = *hi = 0; return; }
  if (a % p == 0) { *lo = a - p < 0 ? 0 : a - p; *hi = a + p > n - 1 ? n - 1 : a + p; }
  else { const int e = a / p; *lo = e * p; *hi = e * p + p; }
}

void orc_destroy(orc_system* s) {
  if (!s) return;
  free(s->indptr); free(s->indices); free(s->data); free(s->b); free(s->u0); free(s->diag); free(s->mask);
  for (int a = 0; a < 3; ++a) { free(s->m1d[a]); free(s->k1d[a]); }
  free(s);
}

/* cells: (nx, ny, nz) with nx = 0 for a 2-D (ny x nz) mesh; unit square / unit cube. */
orc_system* orc_build(int dim, int degree, const int* cells, double k1, double k2, double beta, double mu,
                      int64_t nbc0, const int32_t* bc_nodes0, const double* bc_vals0, int64_t nbc1,
                      const int32_t* bc_nodes1, const double* bc_vals1) {
  orc_system* s = (orc_system*)calloc(1, sizeof(orc_system));
  const int p = degree, w = 2 * p + 1;
  s->dim = dim; s->degree = p;
  int nc[3] = {dim == 3 ? cells[0] : 0, dim == 3 ? cells[1] : cells[0], dim == 3 ? cells[2] : cells[1]};
  for (int a = 0; a < 3; ++a) {
    s->n[a] = nc[a] > 0 ? p * nc[a] + 1 : 1;
    tables_1d(nc[a], p, nc[a] > 0 ? 1.0 / nc[a] : 1.0, &s->m1d[a], &s->k1d[a]);
  }
  const int n0 = s->n[0], n1 = s->n[1], n2 = s->n[2];
  const int64_t nn = (int64_t)n0 * n1 * n2;
  s->n_nodes = nn; s->n_dof = 2 * nn;
  s->mask = (uint8_t*)calloc((size_t)s->n_dof, 1);
  s->u0 = (double*)calloc((size_t)s->n_dof, sizeof(double));
  for (int64_t t = 0; t < nbc0; ++t) { s->mask[bc_nodes0[t]] = 1; s->u0[bc_nodes0[t]] = bc_vals0[t]; }
  for (int64_t t = 0; t < nbc1; ++t) { s->mask[nn + bc_nodes1[t]] = 1; s->u0[nn + bc_nodes1[t]] = bc_vals1[t]; }
  /* row widths: tensor product of the 1-D neighbour ranges, times two column fields */
  s->indptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(s->n_dof + 1));
  s->indptr[0] = 0;
  for (int f = 0; f < 2; ++f)
    for (int64_t node = 0; node < nn; ++node) {
      const int k = (int)(node % n2), j = (int)((node / n2) % n1), i = (int)(node / ((int64_t)n1 * n2));
      int lo, hi, cnt = 1;
      nbr_range(i, n0, p, &lo, &hi); cnt *= hi - lo + 1;
      nbr_range(j, n1, p, &lo, &hi); cnt *= hi - lo + 1;
      nbr_range(k, n2, p, &lo, &hi); cnt *= hi - lo + 1;
      s->indptr[f * nn + node + 1] = 2 * cnt;
    }
  for (int64_t r = 0; r < s->n_dof; ++r) s->indptr[r + 1] += s->indptr[r];
  s->nnz = s->indptr[s->n_dof];
  s->indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)s->nnz);
  s->data = (double*)malloc(sizeof(double) * (size_t)s->nnz);
  s->b = (double*)calloc((size_t)s->n_dof, sizeof(double));
  s->diag = (double*)malloc(sizeof(double) * (size_t)s->n_dof);
  const double cK[2] = {k1 / mu, k2 / mu}, bm = beta / mu;
  /* fill: entry (f,node ; g,col) = [f==g] cK_f K + (f==g ? bm : -bm) M ; lifted RHS from the
   * un-eliminated row; then symmetric elimination (conditioning.py:62 / SURVEY A.3) */
#pragma omp parallel for schedule(static)
  for (int64_t row = 0; row < s->n_dof; ++row) {
    const int f = row >= nn;
    const int64_t node = row - (int64_t)f * nn;
    const int k = (int)(node % n2), j = (int)((node / n2) % n1), i = (int)(node / ((int64_t)n1 * n2));
    int ilo, ihi, jlo, jhi, klo, khi;
    nbr_range(i, n0, p, &ilo, &ihi);
    nbr_range(j, n1, p, &jlo, &jhi);
    nbr_range(k, n2, p, &klo, &khi);
    int64_t q = s->indptr[row];
    double lift = 0.0;
    for (int g = 0; g < 2; ++g)
      for (int ii = ilo; ii <= ihi; ++ii) {
        const double mx = s->m1d[0][(size_t)i * w + p + (ii - i)], kx = s->k1d[0][(size_t)i * w + p + (ii - i)];
        for (int jj = jlo; jj <= jhi; ++jj) {
          const double my = s->m1d[1][(size_t)j * w + p + (jj - j)], ky = s->k1d[1][(size_t)j * w + p + (jj - j)];
          for (int kk = klo; kk <= khi; ++kk) {
            const double mz = s->m1d[2][(size_t)k * w + p + (kk - k)], kz = s->k1d[2][(size_t)k * w + p + (kk - k)];
            const double K = kx * my * mz + mx * ky * mz + mx * my * kz, M = mx * my * mz;
            const int64_t cnode = ((int64_t)ii * n1 + jj) * n2 + kk;
            const int64_t col = (int64_t)g * nn + cnode;
            double v = (f == g) ? cK[f] * K + bm * M : -bm * M;
            lift += v * s->u0[col];
            if (s->mask[row] || s->mask[col]) v = (row == col) ? 1.0 : 0.0;
            s->indices[q] = (int32_t)col;
            s->data[q] = v;
            if (col == row) s->diag[row] = v;
            ++q;
          }
        }
      }
    s->b[row] = s->mask[row] ? 0.0 : -lift;
  }
  return s;
}

int64_t orc_n_dof(const orc_system* s) { return s->n_dof; }
int64_t orc_nnz(const orc_system* s) { return s->nnz; }
void orc_export(const orc_system* s, int64_t* indptr, int32_t* indices, double* data, double* b, double* u0) {
  memcpy(indptr, s->indptr, sizeof(int64_t) * (size_t)(s->n_dof + 1));
  memcpy(indices, s->indices, sizeof(int32_t) * (size_t)s->nnz);
  memcpy(data, s->data, sizeof(double) * (size_t)s->nnz);
  if (b) memcpy(b, s->b, sizeof(double) * (size_t)s->n_dof);
  if (u0) memcpy(u0, s->u0, sizeof(double) * (size_t)s->n_dof);
}

/* PETSc MatMult_SeqAIJ: y = A x, one row per inner loop */
void orc_spmv(const orc_system* s, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < s->n_dof; ++r) {
    double acc = 0.0;
    for (int64_t q = s->indptr[r]; q < s->indptr[r + 1]; ++q) acc += s->data[q] * x[s->indices[q]];
    y[r] = acc;
  }
}

static double dot(int64_t n, const double* a, const double* b) {
  double t = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : t)
  for (int64_t i = 0; i < n; ++i) t += a[i] * b[i];
  return t;
}

/* KSPConvergedDefault */
static int conv_test(int its, double rnorm, double* rnorm0, double* ttol, double rtol, double atol, double dtol) {
  if (its == 0) { *rnorm0 = rnorm; *ttol = fmax(rtol * rnorm, atol); }
  if (!isfinite(rnorm)) return -9;
  if (rnorm <= *ttol) return rnorm < atol ? 3 : 2;
  if (rnorm >= dtol * *rnorm0) return -4;
  return 0;
}

/* KSPCG, KSP_NORM_PRECONDITIONED, zero initial guess; pc: 0 none, 1 jacobi.  Returns its.
 * u (optional) = u0 + x.  hist (optional, capacity hist_cap) = ||z|| per iteration from 0. */
int orc_cg(const orc_system* s, int pc, double rtol, double atol, double dtol, int max_it, double* u,
           double* rnorm_out, int* reason_out, double* hist, int hist_cap, double* spmv_seconds) {
  const int64_t n = s->n_dof;
  double* x = (double*)calloc((size_t)n, sizeof(double));
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  double* z = (double*)malloc(sizeof(double) * (size_t)n);
  double* pv = (double*)calloc((size_t)n, sizeof(double));
  double* wv = (double*)malloc(sizeof(double) * (size_t)n);
  double* dinv = (double*)malloc(sizeof(double) * (size_t)n); /* PCJACOBI stores the reciprocal diagonal */
  double t_mv = 0.0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) { dinv[i] = pc ? 1.0 / s->diag[i] : 1.0; r[i] = s->b[i]; z[i] = dinv[i] * r[i]; }
  double dp = sqrt(dot(n, z, z)), rnorm0 = 0, ttol = 0;
  int reason = conv_test(0, dp, &rnorm0, &ttol, rtol, atol, dtol);
  if (hist && hist_cap > 0) hist[0] = dp;
  double beta = dot(n, z, r), betaold = 1.0;
  int its = 0;
  while (!reason) {
    if (its >= max_it) { reason = -3; break; }
    if (beta == 0.0) { reason = 3; break; }
    const double bb = its == 0 ? 0.0 : beta / betaold;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) pv[i] = z[i] + bb * pv[i];
    betaold = beta;
#ifdef _OPENMP
    const double t0 = omp_get_wtime();
#endif
    orc_spmv(s, pv, wv);
#ifdef _OPENMP
    t_mv += omp_get_wtime() - t0;
#endif
    const double dpi = dot(n, pv, wv);
    if (!(dpi > 0.0) || !isfinite(dpi)) { reason = -5; break; }
    const double a = beta / dpi;
    double zz = 0.0, rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : zz, rz)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += a * pv[i];
      const double rn = r[i] - a * wv[i];
      r[i] = rn;
      const double zv = dinv[i] * rn;
      z[i] = zv;
      zz += zv * zv;
      rz += zv * rn;
    }
    dp = sqrt(zz);
    beta = rz;
    ++its;
    if (hist && its < hist_cap) hist[its] = dp;
    reason = conv_test(its, dp, &rnorm0, &ttol, rtol, atol, dtol);
  }
  if (u)
    for (int64_t i = 0; i < n; ++i) u[i] = s->u0[i] + x[i];
  if (rnorm_out) *rnorm_out = dp;
  if (reason_out) *reason_out = reason;
  if (spmv_seconds) *spmv_seconds = t_mv;
  free(x); free(r); free(z); free(pv); free(wv); free(dinv);
  return its;
}

void orc_get_b(const orc_system* s, double* b) { memcpy(b, s->b, sizeof(double) * (size_t)s->n_dof); }

/* ------------------------------------------------------------------------------------------------
 * KSPGMRES(restart) with left preconditioning and PCFIELDSPLIT (solvers/parameters.py:21-57;
 * BASELINE config 5).  Same arithmetic as oracle/dpp_oracle.py: ksp_gmres / pc_fieldsplit, to which
 * tests/test_oracle_c.py pins it: classical Gram-Schmidt without refinement (VecMDot on the
 * un-modified w, then VecMAXPY), Givens residual recurrence, true preconditioned residual
 * recomputed at each restart, zero initial guess, KSPConvergedDefault on the preconditioned norm.
 * ------------------------------------------------------------------------------------------------ */

/* y = A[fr][fc] x on one field block; every row stores its field-0 columns first, then field 1 */
static void spmv_block(const orc_system* s, int fr, int fc, const double* x, double* y) {
  const int64_t nn = s->n_nodes;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nn; ++r) {
    const int64_t row = (int64_t)fr * nn + r;
    const int64_t q0 = s->indptr[row], half = (s->indptr[row + 1] - q0) / 2;
    const int64_t qa = q0 + (fc ? half : 0);
    double acc = 0.0;
    for (int64_t q = qa; q < qa + half; ++q) acc += s->data[q] * x[s->indices[q] - (int64_t)fc * nn];
    y[r] = acc;
  }
}

typedef struct {
  int ksp;         /* 0 = preonly (one Jacobi application), 1 = cg */
  double rtol, atol;
  int max_it;
} orc_inner;

/* inner block solve  y = A_ff^-1 r  by Jacobi-CG (KSPCG semantics as orc_cg), zero initial guess */
static int block_cg(const orc_system* s, int f, const orc_inner* in, const double* rhs, double* x, double* wk /*[4*nn]*/) {
  const int64_t nn = s->n_nodes;
  const double* diag = s->diag + (int64_t)f * nn;
  double *r = wk, *z = wk + nn, *pv = wk + 2 * nn, *wv = wk + 3 * nn;
  if (in->ksp == 0) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) x[i] = rhs[i] / diag[i];
    return 1;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nn; ++i) { x[i] = 0.0; r[i] = rhs[i]; z[i] = (1.0 / diag[i]) * r[i]; pv[i] = 0.0; }
  double dp = sqrt(dot(nn, z, z)), rnorm0 = 0, ttol = 0;
  int reason = conv_test(0, dp, &rnorm0, &ttol, in->rtol, in->atol, 1e4);
  double beta = dot(nn, z, r), betaold = 1.0;
  int its = 0;
  while (!reason) {
    if (its >= in->max_it || beta == 0.0) break;
    const double bb = its == 0 ? 0.0 : beta / betaold;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) pv[i] = z[i] + bb * pv[i];
    betaold = beta;
    spmv_block(s, f, f, pv, wv);
    const double dpi = dot(nn, pv, wv);
    if (!(dpi > 0.0) || !isfinite(dpi)) break;
    const double a = beta / dpi;
    double zz = 0.0, rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : zz, rz)
    for (int64_t i = 0; i < nn; ++i) {
      x[i] += a * pv[i];
      const double rn = r[i] - a * wv[i];
      r[i] = rn;
      const double zv = (1.0 / diag[i]) * rn;
      z[i] = zv;
      zz += zv * zv;
      rz += zv * rn;
    }
    dp = sqrt(zz);
    beta = rz;
    ++its;
    reason = conv_test(its, dp, &rnorm0, &ttol, in->rtol, in->atol, 1e4);
  }
  return its;
}

/* pc: 0 none, 1 jacobi, 2 fieldsplit multiplicative, 3 fieldsplit additive.  out = B in */
static void pc_apply(const orc_system* s, int pc, const orc_inner* in, const double* vin, double* vout, double* wk,
                     int64_t* inner_its) {
  const int64_t n = s->n_dof, nn = s->n_nodes;
  if (pc == 0) { memcpy(vout, vin, sizeof(double) * (size_t)n); return; }
  if (pc == 1) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) vout[i] = (1.0 / s->diag[i]) * vin[i];
    return;
  }
  double* r1 = wk + 4 * nn;
  *inner_its += block_cg(s, 0, in, vin, vout, wk);
  if (pc == 2) {
    spmv_block(s, 1, 0, vout, r1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nn; ++i) r1[i] = vin[nn + i] - r1[i];
  } else {
    memcpy(r1, vin + nn, sizeof(double) * (size_t)nn);
  }
  *inner_its += block_cg(s, 1, in, r1, vout + nn, wk);
}

int orc_gmres(const orc_system* s, int pc, int restart, double rtol, double atol, double dtol, int max_it,
              int inner_ksp, double inner_rtol, double inner_atol, int inner_max_it, double* u, double* rnorm_out,
              int* reason_out, double* hist, int hist_cap, int64_t* inner_its_out) {
  const int64_t n = s->n_dof;
  const int m = restart;
  orc_inner in = {inner_ksp, inner_rtol, inner_atol, inner_max_it};
  double* x = (double*)calloc((size_t)n, sizeof(double));
  double* V = (double*)malloc(sizeof(double) * (size_t)n * (size_t)(m + 1));
  double* w = (double*)malloc(sizeof(double) * (size_t)n);
  double* t = (double*)malloc(sizeof(double) * (size_t)n);
  double* wk = (double*)malloc(sizeof(double) * (size_t)(5 * s->n_nodes));
  double* H = (double*)calloc((size_t)(m + 2) * (size_t)(m + 1), sizeof(double));
  double* cc = (double*)calloc((size_t)m + 1, sizeof(double));
  double* ss = (double*)calloc((size_t)m + 1, sizeof(double));
  double* grs = (double*)calloc((size_t)m + 2, sizeof(double));
  double* hcol = (double*)calloc((size_t)m + 2, sizeof(double));
#define HH(a, b) H[(size_t)(a) * (size_t)(m + 1) + (size_t)(b)]
  int64_t inner_its = 0;
  int its = 0, reason = 0, first = 1, nh = 0;
  double res = 0.0, rnorm0 = 0, ttol = 0;
  for (;;) {
    if (first) {
      pc_apply(s, pc, &in, s->b, w, wk, &inner_its);
    } else {
      orc_spmv(s, x, t);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) t[i] = s->b[i] - t[i];
      pc_apply(s, pc, &in, t, w, wk, &inner_its);
    }
    first = 0;
    res = sqrt(dot(n, w, w));
    if (hist && nh < hist_cap) hist[nh] = res;
    ++nh;
    if (res == 0.0) { reason = 3; break; }
    reason = conv_test(its, res, &rnorm0, &ttol, rtol, atol, dtol);
    if (reason) break;
    if (its >= max_it) { reason = -3; break; }
    memset(H, 0, sizeof(double) * (size_t)(m + 2) * (size_t)(m + 1));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) V[i] = w[i] / res;
    grs[0] = res;
    int it = 0;
    while (!reason && it < m && its < max_it) {
      if (it) { if (hist && nh < hist_cap) hist[nh] = res; ++nh; }
      orc_spmv(s, V + (size_t)it * n, t);
      pc_apply(s, pc, &in, t, w, wk, &inner_its);
      for (int j = 0; j <= it; ++j) hcol[j] = dot(n, V + (size_t)j * n, w);   /* VecMDot */
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {                                        /* VecMAXPY */
        double acc = w[i];
        for (int j = 0; j <= it; ++j) acc -= hcol[j] * V[(size_t)j * n + i];
        w[i] = acc;
      }
      const double tt = sqrt(dot(n, w, w));
      int hapend = 0;
      double hapbnd = 1e-30;
      if (grs[it] != 0.0) { hapbnd = fabs(tt / grs[it]); if (hapbnd > 1e-30) hapbnd = 1e-30; }
      if (tt < hapbnd) hapend = 1;
      else {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) V[(size_t)(it + 1) * n + i] = w[i] / tt;
      }
      for (int j = 0; j <= it; ++j) HH(j, it) = hcol[j];
      HH(it + 1, it) = tt;
      for (int j = 0; j < it; ++j) {
        const double a = HH(j, it);
        HH(j, it) = cc[j] * a + ss[j] * HH(j + 1, it);
        HH(j + 1, it) = cc[j] * HH(j + 1, it) - ss[j] * a;
      }
      if (!hapend) {
        const double d = sqrt(HH(it, it) * HH(it, it) + HH(it + 1, it) * HH(it + 1, it));
        if (d == 0.0) { reason = -5; break; }
        cc[it] = HH(it, it) / d;
        ss[it] = HH(it + 1, it) / d;
        grs[it + 1] = -ss[it] * grs[it];
        grs[it] = cc[it] * grs[it];
        HH(it, it) = cc[it] * HH(it, it) + ss[it] * HH(it + 1, it);
        res = fabs(grs[it + 1]);
      } else {
        res = 0.0;
      }
      ++it;
      ++its;
      reason = conv_test(its, res, &rnorm0, &ttol, rtol, atol, dtol);
      if (hapend && !reason) reason = -5;
    }
    if (it && (reason || its >= max_it)) { if (hist && nh < hist_cap) hist[nh] = res; ++nh; }
    if (it) {   /* KSPGMRESBuildSoln */
      double* y = hcol;
      for (int k = it - 1; k >= 0; --k) {
        double acc = grs[k];
        for (int j = k + 1; j < it; ++j) acc -= HH(k, j) * y[j];
        y[k] = acc / HH(k, k);
      }
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        double acc = x[i];
        for (int j = 0; j < it; ++j) acc += y[j] * V[(size_t)j * n + i];
        x[i] = acc;
      }
    }
    if (reason) break;
    if (its >= max_it) { reason = -3; break; }
  }
#undef HH
  if (u)
    for (int64_t i = 0; i < n; ++i) u[i] = s->u0[i] + x[i];
  if (rnorm_out) *rnorm_out = res;
  if (reason_out) *reason_out = reason;
  if (inner_its_out) *inner_its_out = inner_its;
  free(x); free(V); free(w); free(t); free(wk); free(H); free(cc); free(ss); free(grs); free(hcol);
  return its;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
