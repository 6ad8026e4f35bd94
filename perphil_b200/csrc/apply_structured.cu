// Matrix-free DPP operator on rectilinear tensor grids (DPP_KERNEL_STRUCTURED).
//
// Replaces PETSc MatMult on the assembled dpp_form matrix (solver.py:71; forms/dpp.py:27,57,89)
// for meshes whose nodes are numbered lexicographically (x slowest, z contiguous) -- every
// BASELINE.json configuration.  With assembled 1-D matrices Kx,Mx,Ky,My,Kz,Mz
//     K = Kx(x)My(x)Mz + Mx(x)Ky(x)Mz + Mx(x)My(x)Kz,     M = Mx(x)My(x)Mz
// so the apply is a sum-factorised sweep.  The kernel streams planes i = const through shared
// memory (one (j,k) tile per CTA, halo 1), evaluates the in-plane parts
//     c = (My(x)Mz) x_i,   d = (Ky(x)Mz + My(x)Kz) x_i
// with loop-invariant per-thread coefficient products, and carries the x-direction 3-point sweep
// in a register queue:  K x = Kx c + Mx d,  M x = Mx c.  Dirichlet column elimination is applied
// on load, row elimination on store, and the CG reduction <x, A x> is fused into the store.
//
// HBM traffic per apply (algorithmic): read x (8 B/dof) + write y (8 B/dof) + 1 B/dof mask
// = 34 B/node for the two-field operator.  No index arrays, no coordinates.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "dpp_internal.cuh"

namespace dpp {

namespace {

struct StructArgs {
  int n[3];
  const double* m1d[3];
  const double* k1d[3];
  const double* x[2];
  double* y[2];
  const uint8_t* in_mask[2];
  const uint8_t* out_mask[2];
  int identity_on_masked;
  Coef c;
  double* dot_partials;
  int i_begin, i_end;  // owned planes
  int ntj, ntk, nseg;
  const double* skip_flag;  // device scalar: != 0 -> kernel is a no-op (solver already finished)
};

__device__ __forceinline__ int balanced_start(int t, int n, int nt) { return (int)(((long long)t * n) / nt); }

constexpr int TKW = 32;

template <int NF, int TJ>
__global__ void __launch_bounds__(TKW* TJ) k_apply_q1(const StructArgs s) {
  if (s.skip_flag != nullptr && *s.skip_flag != 0.0) return;
  __shared__ double xs[2][NF][TJ + 2][TKW + 2];
  __shared__ double red[TJ];

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2];
  const int tile = blockIdx.x;
  const int tkid = tile % s.ntk, tjid = tile / s.ntk;
  const int k0 = balanced_start(tkid, nk, s.ntk), k1 = balanced_start(tkid + 1, nk, s.ntk);
  const int j0 = balanced_start(tjid, nj, s.ntj), j1 = balanced_start(tjid + 1, nj, s.ntj);
  const int nown = s.i_end - s.i_begin;
  const int i_lo = s.i_begin + balanced_start(blockIdx.y, nown, s.nseg);
  const int i_hi = s.i_begin + balanced_start(blockIdx.y + 1, nown, s.nseg);

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TKW + tx;
  const int j = j0 + ty, k = k0 + tx;
  const bool active = (j < j1) && (k < k1);
  const long long plane = (long long)nj * nk;

  // loop-invariant in-plane coefficient products for this thread's (j,k)
  double cm[3][3], ck[3][3];
#pragma unroll
  for (int dj = 0; dj < 3; ++dj) {
    const double my = active ? __ldg(&s.m1d[1][j * 3 + dj]) : 0.0;
    const double ky = active ? __ldg(&s.k1d[1][j * 3 + dj]) : 0.0;
#pragma unroll
    for (int dk = 0; dk < 3; ++dk) {
      const double mz = active ? __ldg(&s.m1d[2][k * 3 + dk]) : 0.0;
      const double kz = active ? __ldg(&s.k1d[2][k * 3 + dk]) : 0.0;
      cm[dj][dk] = my * mz;
      ck[dj][dk] = ky * mz + my * kz;
    }
  }

  double qc[NF][3], qd[NF][3];
  double xcen[NF][2];  // this thread's (masked) input value at planes i-1, i
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int d = 0; d < 3; ++d) qc[f][d] = qd[f][d] = 0.0;
    xcen[f][0] = xcen[f][1] = 0.0;
  }

  double dot = 0.0;
  const int tile_elems = (TJ + 2) * (TKW + 2);

  for (int i = i_lo - 1; i <= i_hi; ++i) {
    const int b = (i - (i_lo - 1)) & 1;
    const bool in_dom = (i >= 0) && (i < ni);
    if (in_dom) {
      for (int e = tid; e < tile_elems; e += TKW * TJ) {
        const int r = e / (TKW + 2), cidx = e - r * (TKW + 2);
        const int jj = j0 - 1 + r, kk = k0 - 1 + cidx;
        const bool ok = (jj >= 0) && (jj < nj) && (kk >= 0) && (kk < nk) && (jj <= j1) && (kk <= k1);
        const long long node = (long long)i * plane + (long long)jj * nk + kk;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          double v = 0.0;
          if (ok) {
            v = __ldg(&s.x[f][node]);
            if (s.in_mask[f] != nullptr && s.in_mask[f][node]) v = 0.0;
          }
          xs[b][f][r][cidx] = v;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      qc[f][0] = qc[f][1];
      qc[f][1] = qc[f][2];
      qd[f][0] = qd[f][1];
      qd[f][1] = qd[f][2];
      double c = 0.0, d = 0.0;
      xcen[f][0] = xcen[f][1];
      xcen[f][1] = 0.0;
      if (in_dom && active) {
        xcen[f][1] = xs[b][f][ty + 1][tx + 1];
#pragma unroll
        for (int dj = 0; dj < 3; ++dj)
#pragma unroll
          for (int dk = 0; dk < 3; ++dk) {
            const double v = xs[b][f][ty + dj][tx + dk];
            c = fma(cm[dj][dk], v, c);
            d = fma(ck[dj][dk], v, d);
          }
      }
      qc[f][2] = c;
      qd[f][2] = d;
    }
    const int io = i - 1;
    if (active && io >= i_lo && io < i_hi) {
      double mx[3], kx[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mx[d] = __ldg(&s.m1d[0][io * 3 + d]);
        kx[d] = __ldg(&s.k1d[0][io * 3 + d]);
      }
      double Kx[NF], Mx[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        double kk_ = 0.0, mm_ = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          kk_ = fma(kx[d], qc[f][d], kk_);
          kk_ = fma(mx[d], qd[f][d], kk_);
          mm_ = fma(mx[d], qc[f][d], mm_);
        }
        Kx[f] = kk_;
        Mx[f] = mm_;
      }
      const long long node = (long long)io * plane + (long long)j * nk + k;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        double yv = 0.0;
#pragma unroll
        for (int g = 0; g < NF; ++g) {
          yv = fma(s.c.cK[f][g], Kx[g], yv);
          yv = fma(s.c.cM[f][g], Mx[g], yv);
        }
        double xc;
        if (s.out_mask[f] != nullptr && s.out_mask[f][node]) {
          xc = s.x[f][node];
          yv = s.identity_on_masked ? xc : 0.0;
        } else {
          xc = xcen[f][0];  // (the other smem buffer may already be refilled by a faster warp)
        }
        s.y[f][node] = yv;
        dot = fma(xc, yv, dot);
      }
    }
  }

  if (s.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (tx == 0) red[ty] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < TJ; ++w) t += red[w];
      s.dot_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

__global__ void k_diag_structured(int n0, int n1, int n2, int band, const double* __restrict__ mx,
                                  const double* __restrict__ kx, const double* __restrict__ my,
                                  const double* __restrict__ ky, const double* __restrict__ mz,
                                  const double* __restrict__ kz, Coef c, const uint8_t* __restrict__ mask,
                                  long long n_nodes, double* __restrict__ diag) {
  const int w = 2 * band + 1;
  for (long long node = blockIdx.x * (long long)blockDim.x + threadIdx.x; node < n_nodes;
       node += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(node % n2);
    const long long t = node / n2;
    const int j = (int)(t % n1);
    const int i = (int)(t / n1);
    const double mxc = mx[i * w + band], kxc = kx[i * w + band];
    const double myc = my[j * w + band], kyc = ky[j * w + band];
    const double mzc = mz[k * w + band], kzc = kz[k * w + band];
    const double K = kxc * myc * mzc + mxc * kyc * mzc + mxc * myc * kzc;
    const double M = mxc * myc * mzc;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      double d = c.cK[f][f] * K + c.cM[f][f] * M;
      if (mask != nullptr && mask[f * n_nodes + node]) d = 1.0;
      diag[f * n_nodes + node] = d;
    }
  }
}

// assembled 1-D rows in band storage [n][2P+1]
void build_1d_tables(const std::vector<double>& v, int p, std::vector<double>& m, std::vector<double>& k) {
  const int nc = (int)v.size() - 1;
  const int n = p * nc + 1, w = 2 * p + 1;
  m.assign((size_t)n * w, 0.0);
  k.assign((size_t)n * w, 0.0);
  if (nc == 0) {  // dummy axis of a 2-D mesh: M = [1], K = [0]
    m[p] = 1.0;
    return;
  }
  for (int e = 0; e < nc; ++e) {
    const double h = v[e + 1] - v[e];
    double Ke[3][3], Me[3][3];
    if (p == 1) {
      const double a = 1.0 / h, b = h / 6.0;
      Ke[0][0] = a; Ke[0][1] = -a; Ke[1][0] = -a; Ke[1][1] = a;
      Me[0][0] = 2 * b; Me[0][1] = b; Me[1][0] = b; Me[1][1] = 2 * b;
    } else {
      const double a = 1.0 / (3.0 * h), b = h / 30.0;
      const double K2[3][3] = {{7, -8, 1}, {-8, 16, -8}, {1, -8, 7}};
      const double M2[3][3] = {{4, 2, -1}, {2, 16, 2}, {-1, 2, 4}};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { Ke[r][c] = a * K2[r][c]; Me[r][c] = b * M2[r][c]; }
    }
    for (int r = 0; r <= p; ++r)
      for (int c = 0; c <= p; ++c) {
        const int row = p * e + r, d = c - r;
        m[(size_t)row * w + p + d] += Me[r][c];
        k[(size_t)row * w + p + d] += Ke[r][c];
      }
  }
}

}  // namespace

// Decide whether (coords, connectivity) describe a lexicographically numbered rectilinear tensor
// grid; if so build the 1-D tables.  Host-side, O(n_cells) with early exit.
int structured_detect_and_setup(dpp_context* ctx, const int32_t* cnm, const double* X, const int32_t* ccnm) {
  ctx->structured_ok = false;
  const int dim = ctx->dim, p = ctx->degree;
  const int64_t nv = ctx->n_coord_nodes;
  if (nv < (1 << dim)) return DPP_OK;
  // infer vertex counts per axis from where the fastest coordinates wrap
  int64_t nvz = nv, nvy = 1, nvx = 1;
  const int last = dim - 1;
  for (int64_t v = 1; v < nv; ++v)
    if (X[v * dim + last] <= X[(v - 1) * dim + last]) { nvz = v; break; }
  if (nv % nvz) return DPP_OK;
  int64_t rest = nv / nvz;
  if (dim == 3) {
    nvy = rest;
    for (int64_t r = 1; r < rest; ++r)
      if (X[(r * nvz) * dim + 1] <= X[((r - 1) * nvz) * dim + 1]) { nvy = r; break; }
    if (rest % nvy) return DPP_OK;
    nvx = rest / nvy;
  } else {
    nvy = rest;
    nvx = 1;
  }
  // axes (dim==2: axis0 dummy, axis1 = x, axis2 = y)
  std::vector<double> ax[3];
  if (dim == 3) {
    ax[0].resize(nvx); ax[1].resize(nvy); ax[2].resize(nvz);
    for (int64_t i = 0; i < nvx; ++i) ax[0][i] = X[(i * nvy * nvz) * 3 + 0];
    for (int64_t j = 0; j < nvy; ++j) ax[1][j] = X[(j * nvz) * 3 + 1];
    for (int64_t k = 0; k < nvz; ++k) ax[2][k] = X[k * 3 + 2];
  } else {
    ax[0].assign(1, 0.0); ax[1].resize(nvy); ax[2].resize(nvz);
    for (int64_t j = 0; j < nvy; ++j) ax[1][j] = X[(j * nvz) * 2 + 0];
    for (int64_t k = 0; k < nvz; ++k) ax[2][k] = X[k * 2 + 1];
  }
  for (int a = (dim == 3 ? 0 : 1); a < 3; ++a) {
    if (ax[a].size() < 2) return DPP_OK;
    for (size_t t = 1; t < ax[a].size(); ++t)
      if (!(ax[a][t] > ax[a][t - 1])) return DPP_OK;
  }
  const int64_t ncx = (dim == 3 ? nvx - 1 : 1), ncy = nvy - 1, ncz = nvz - 1;
  if (ncx * ncy * ncz != ctx->n_cells) return DPP_OK;
  const int64_t NX = (dim == 3 ? p * ncx + 1 : 1), NY = p * ncy + 1, NZ = p * ncz + 1;
  if (NX * NY * NZ != ctx->n_nodes) return DPP_OK;
  // every vertex sits on the tensor grid
  double ext = 0.0;
  for (int a = 0; a < 3; ++a) ext = std::max(ext, ax[a].back() - ax[a].front());
  const double tol = 1e-12 * ext;
  bool ok = true;
#pragma omp parallel for reduction(&& : ok) schedule(static)
  for (int64_t v = 0; v < nv; ++v) {
    if (!ok) continue;
    const int64_t k = v % nvz, t = v / nvz, j = t % nvy, i = t / nvy;
    if (dim == 3) {
      ok = ok && std::fabs(X[v * 3] - ax[0][i]) <= tol && std::fabs(X[v * 3 + 1] - ax[1][j]) <= tol &&
           std::fabs(X[v * 3 + 2] - ax[2][k]) <= tol;
    } else {
      ok = ok && std::fabs(X[v * 2] - ax[1][j]) <= tol && std::fabs(X[v * 2 + 1] - ax[2][k]) <= tol;
    }
  }
  if (!ok) return DPP_OK;
  // every cell is a grid cell with tensor-lexicographic local numbering, each grid cell once
  std::vector<uint8_t> seen((size_t)ctx->n_cells, 0);
  const int nvc = ctx->nvc, npc = ctx->npc, p1 = p + 1;
  int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (int64_t c = 0; c < ctx->n_cells; ++c) {
    if (bad) continue;
    const int64_t v0 = ccnm[c * nvc];
    if (v0 < 0 || v0 >= nv) { bad++; continue; }
    const int64_t ck = v0 % nvz, t = v0 / nvz, cj = t % nvy, ci = t / nvy;
    if (ck >= ncz || cj >= ncy || ci >= ncx) { bad++; continue; }
    int l = 0;
    for (int a = 0; a < (dim == 3 ? 2 : 1); ++a)
      for (int b = 0; b < 2; ++b)
        for (int cc = 0; cc < 2; ++cc, ++l) {
          const int64_t expect = dim == 3 ? ((ci + a) * nvy + cj + b) * nvz + ck + cc : (cj + b) * nvz + ck + cc;
          if (ccnm[c * nvc + l] != expect) bad++;
        }
    l = 0;
    for (int a = 0; a < (dim == 3 ? p1 : 1); ++a)
      for (int b = 0; b < p1; ++b)
        for (int cc = 0; cc < p1; ++cc, ++l) {
          const int64_t expect = dim == 3 ? ((p * ci + a) * NY + p * cj + b) * NZ + p * ck + cc
                                          : (p * cj + b) * NZ + p * ck + cc;
          if (cnm[c * npc + l] != expect) bad++;
        }
    const int64_t lex = (ci * ncy + cj) * ncz + ck;
    seen[lex] = 1;  // benign race: same value
  }
  if (bad) return DPP_OK;
  for (int64_t c = 0; c < ctx->n_cells; ++c)
    if (!seen[c]) return DPP_OK;

  // tables
  const int w = 2 * p + 1;
  const int nn[3] = {(int)NX, (int)NY, (int)NZ};
  size_t total = 0;
  for (int a = 0; a < 3; ++a) total += (size_t)nn[a] * w * 2;
  std::vector<double> host(total);
  size_t off = 0;
  size_t offs_m[3], offs_k[3];
  for (int a = 0; a < 3; ++a) {
    std::vector<double> m, k;
    build_1d_tables(ax[a], p, m, k);
    offs_m[a] = off;
    std::memcpy(&host[off], m.data(), m.size() * sizeof(double));
    off += m.size();
    offs_k[a] = off;
    std::memcpy(&host[off], k.data(), k.size() * sizeof(double));
    off += k.size();
    ctx->h_axis[a] = ax[a];
    if (a == 0) {  // centre entries of axis 0: [interior, domain boundary] (uniform kernel)
      const int n0 = nn[0], w0 = 2 * p + 1;
      ctx->uni_mxc[1] = m[p];
      ctx->uni_kxc[1] = k[p];
      ctx->uni_mxc[0] = n0 > 2 ? m[(size_t)w0 + p] : m[p];
      ctx->uni_kxc[0] = n0 > 2 ? k[(size_t)w0 + p] : k[p];
    }
  }
  DPP_CHECK(dev_alloc(ctx, &ctx->d_tables, (int64_t)total));
  DPP_CUDA(cudaMemcpy(ctx->d_tables, host.data(), total * sizeof(double), cudaMemcpyHostToDevice));
  for (int a = 0; a < 3; ++a) {
    ctx->grid.n[a] = nn[a];
    ctx->grid.m1d[a] = ctx->d_tables + offs_m[a];
    ctx->grid.k1d[a] = ctx->d_tables + offs_k[a];
  }
  ctx->grid.band = p;
  ctx->structured_ok = true;
  // uniform spacing per axis?  (then all off-diagonals of the 1-D matrices coincide)
  bool uni = (p == 1);
  for (int a = 0; a < 3 && uni; ++a) {
    const std::vector<double>& v = ax[a];
    if (v.size() < 2) { ctx->uni_m_off[a] = ctx->uni_k_off[a] = 0.0; continue; }
    const double h0 = v[1] - v[0];
    for (size_t t = 2; t < v.size(); ++t)
      if (std::fabs((v[t] - v[t - 1]) - h0) > 1e-12 * std::fabs(h0)) uni = false;
    ctx->uni_m_off[a] = h0 / 6.0;
    ctx->uni_k_off[a] = -1.0 / h0;
  }
  ctx->grid_uniform = uni;
  bool uni2 = (p == 2);
  for (int a = 0; a < 3 && uni2; ++a) {
    const std::vector<double>& v = ax[a];
    if (v.size() < 2) continue;   // dummy axis
    ctx->uni_h[a] = v[1] - v[0];
    for (size_t t = 2; t < v.size(); ++t)
      if (std::fabs((v[t] - v[t - 1]) - ctx->uni_h[a]) > 1e-12 * std::fabs(ctx->uni_h[a])) uni2 = false;
  }
  ctx->q2_uniform = uni2;
  if (const char* e = getenv("DPP_FORCE_TABLE_KERNEL")) ctx->force_table_kernel = (e[0] == '1');
  return DPP_OK;
}

int structured_apply_q2(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks);  // apply_structured_q2.cu

int structured_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks) {
  const GridDesc& g = ctx->grid;
  if (g.band == 2) return structured_apply_q2(ctx, a, n_partial_blocks);
  if (ctx->grid_uniform && !ctx->force_table_kernel) return structured_apply_uniform(ctx, a, n_partial_blocks);
  const long long plane = (long long)g.n[1] * g.n[2];
  if (a.owned_begin % plane || a.owned_end % plane) {
    ctx->set_error("structured apply: owned range must consist of whole x-planes");
    return DPP_ERR_INVALID;
  }
  StructArgs s{};
  for (int d = 0; d < 3; ++d) { s.n[d] = g.n[d]; s.m1d[d] = g.m1d[d]; s.k1d[d] = g.k1d[d]; }
  for (int f = 0; f < 2; ++f) {
    s.x[f] = a.x[f]; s.y[f] = a.y[f]; s.in_mask[f] = a.in_mask[f]; s.out_mask[f] = a.out_mask[f];
  }
  s.identity_on_masked = a.identity_on_masked;
  s.c = a.c;
  s.dot_partials = a.dot_partials;
  s.i_begin = (int)(a.owned_begin / plane);
  s.i_end = (int)(a.owned_end / plane);
  s.skip_flag = a.skip_flag;
  constexpr int TJ = 8;
  s.ntk = (g.n[2] + TKW - 1) / TKW;
  s.ntj = (g.n[1] + TJ - 1) / TJ;
  const int tiles = s.ntk * s.ntj;
  const int nown = s.i_end - s.i_begin;
  if (nown <= 0) { if (n_partial_blocks) *n_partial_blocks = 0; return DPP_OK; }
  // enough x-segments for >= ~6 CTAs per SM, but at least 8 planes per segment (2 redundant planes each)
  int nseg = (ctx->sm_count * 6 + tiles - 1) / tiles;
  nseg = std::max(1, std::min(nseg, std::max(1, nown / 8)));
  while ((long long)tiles * nseg > kMaxPartialBlocks && nseg > 1) --nseg;
  if ((long long)tiles * nseg > kMaxPartialBlocks && a.dot_partials != nullptr) {
    ctx->set_error("structured apply: too many tiles for the reduction scratch");
    return DPP_ERR_INVALID;
  }
  s.nseg = nseg;
  dim3 grid(tiles, nseg), block(TKW, TJ);
  if (a.nf == 2)
    k_apply_q1<2, TJ><<<grid, block, 0, ctx->stream>>>(s);
  else
    k_apply_q1<1, TJ><<<grid, block, 0, ctx->stream>>>(s);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (n_partial_blocks) *n_partial_blocks = tiles * nseg;
  return DPP_OK;
}

int structured_diagonal(dpp_context* ctx, const Coef& c, double* d_diag) {
  const GridDesc& g = ctx->grid;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((ctx->n_nodes + threads - 1) / threads, (long long)ctx->sm_count * 16);
  k_diag_structured<<<blocks, threads, 0, ctx->stream>>>(g.n[0], g.n[1], g.n[2], g.band, g.m1d[0], g.k1d[0],
                                                         g.m1d[1], g.k1d[1], g.m1d[2], g.k1d[2], c, ctx->d_mask,
                                                         ctx->n_nodes, d_diag);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

}  // namespace dpp
