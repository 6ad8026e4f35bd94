// Element-by-element assembly of the 2x2-block DPP matrix into CSR + CSR SpMV (K1/K2/K3a).
//
// Replaces fd.assemble(a, bcs=bcs, mat_type="aij") (solvers/conditioning.py:51-63; the Jacobian
// assembly inside LinearVariationalSolver, solvers/solver.py:66-71) and PETSc SeqAIJ MatMult.
//
// Symbolic phase (once per mesh, integer work only, no atomics):
//   node graph rows = sorted union of the nodes of the incident cells (row-owner gather over the
//   node->cell adjacency), exclusive scan -> row pointers, and the SCATTER PERMUTATION
//   pos[cell][a][b] = position of column node(cell,b) inside row node(cell,a)  (one byte each).
//   The monolithic pattern is the node graph replicated in 2x2 blocks: row r of field f holds
//   [cols(r), n + cols(r)], columns sorted -- the full element pattern PETSc preallocates.
// Numeric phase (per parameter / Dirichlet change, fp64, no atomics, bitwise reproducible):
//   a group of LANES >= nodes_per_cell lanes owns one node-graph row for both fields.  The incident
//   cells are visited in ascending order; lane b adds K_e[a][b], M_e[a][b] into slot pos[cell][a][b]
//   of a shared-memory row buffer (the slots of one cell are distinct: no conflicts, fixed order);
//   non-affine cells first tabulate the metric at the (P+1)^dim Gauss points, one point per lane.
//   The group then writes the four block segments of the row with consecutive lanes on consecutive
//   entries (coalesced), with Firedrake's Dirichlet semantics (constrained rows/columns zeroed, unit
//   diagonal; explicit zeros stay in the pattern, conditioning.py:86 removes them on the host).
//   Bytes: writes 8 B per entry; reads the permutation (1 B per (cell, a, b)), the node graph and
//   the column masks.
#include <algorithm>
#include <vector>

#include "dpp_internal.cuh"
#include "fe_common.cuh"

namespace dpp {

struct CsrMatrix {
  int64_t n_nodes = 0, nnz_g = 0, nnz = 0;
  int64_t* g_ptr = nullptr;   // [n+1] node graph
  int32_t* g_cols = nullptr;  // [nnz_g]
  uint8_t* pos = nullptr;     // [n_cells*npc*npc] scatter permutation
  int64_t* indptr = nullptr;  // [2n+1]
  int32_t* indices = nullptr; // [4 nnz_g]
  double* data = nullptr;     // [4 nnz_g]
  double* tabs = nullptr;     // [7][729] reference-cell integrals (6 stiffness tables + mass) for the numeric kernel
  bool numeric_valid = false;
};

namespace {

// reference-cell integrals, indexed like the per-cell metric: 0:G00 1:G01 2:G02 3:G11 4:G12 5:G22
// (off-diagonal tables hold the symmetrised sum T_rs + T_sr); mass table separately.
static __constant__ double cTK[6 * 729];
static __constant__ double cTM[729];

template <int NPC, int MAXC>
__global__ void k_graph_rows(long long n, const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_cell,
                             const int32_t* __restrict__ cnm, int64_t* __restrict__ counts,
                             const int64_t* __restrict__ g_ptr, int32_t* __restrict__ g_cols) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    int32_t cand[MAXC];
    int m = 0;
    for (long long e = adj_ptr[r]; e < adj_ptr[r + 1]; ++e) {
      const long long cell = adj_cell[e];
      for (int b = 0; b < NPC; ++b) {
        const int32_t v = cnm[cell * NPC + b];
        // insertion into the sorted unique list
        int lo = 0, hi = m;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (cand[mid] < v) lo = mid + 1; else hi = mid;
        }
        if (lo < m && cand[lo] == v) continue;
        if (m < MAXC) {
          for (int t = m; t > lo; --t) cand[t] = cand[t - 1];
          cand[lo] = v;
          ++m;
        }
      }
    }
    if (g_cols == nullptr) {
      counts[r] = m;
    } else {
      const long long base = g_ptr[r];
      for (int t = 0; t < m; ++t) g_cols[base + t] = cand[t];
    }
  }
}

// exclusive scan in three phases (deterministic integer work): per-block sums, scan of the block sums by one
// block, per-block scan with the block offset added
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                      // items per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total, long long* sm /*[SCAN_THREADS/32 + 1]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) sm[wid] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) { const long long t = sm[w]; sm[w] = acc; acc += t; }
    sm[SCAN_THREADS / 32] = acc;
  }
  __syncthreads();
  *total = sm[SCAN_THREADS / 32];
  return sm[wid] + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(long long n, const int64_t* __restrict__ counts,
                                                                  int64_t* __restrict__ block_sums) {
  __shared__ long long sm[SCAN_THREADS / 32 + 1];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  long long s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += counts[base + i];
  long long total;
  block_exclusive_scan(s, &total, sm);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// in-place exclusive scan of the block sums by one block (chunked over its threads); sums[nb] = grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(long long nb, int64_t* __restrict__ sums) {
  __shared__ long long sm[SCAN_THREADS / 32 + 1];
  const long long per = (nb + SCAN_THREADS - 1) / SCAN_THREADS;
  const long long b = (long long)threadIdx.x * per, e = b + per < nb ? b + per : nb;
  long long s = 0;
  for (long long i = b; i < e; ++i) s += sums[i];
  long long total;
  long long acc = block_exclusive_scan(s, &total, sm);
  for (long long i = b; i < e; ++i) { const long long v = sums[i]; sums[i] = acc; acc += v; }
  if (threadIdx.x == 0) sums[nb] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_finish(long long n, const int64_t* __restrict__ counts,
                                                              const int64_t* __restrict__ block_offs,
                                                              int64_t* __restrict__ ptr) {
  __shared__ long long sm[SCAN_THREADS / 32 + 1];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  long long v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? counts[base + i] : 0;
    s += v[i];
  }
  long long total;
  long long acc = block_exclusive_scan(s, &total, sm) + block_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) ptr[base + i] = acc;
    acc += v[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ptr[n] = block_offs[gridDim.x];
}

template <int NPC>
__global__ void k_positions(long long n_cells, const int32_t* __restrict__ cnm, const int64_t* __restrict__ g_ptr,
                            const int32_t* __restrict__ g_cols, uint8_t* __restrict__ pos) {
  const long long total = n_cells * NPC;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long cell = t / NPC;
    const int a = (int)(t - cell * NPC);
    const long long row = cnm[cell * NPC + a];
    const long long base = g_ptr[row];
    const int len = (int)(g_ptr[row + 1] - base);
    for (int b = 0; b < NPC; ++b) {
      const int32_t v = cnm[cell * NPC + b];
      int lo = 0, hi = len;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g_cols[base + mid] < v) lo = mid + 1; else hi = mid;
      }
      pos[(cell * NPC + a) * NPC + b] = (uint8_t)lo;
    }
  }
}

__global__ void k_block_pattern(long long n, long long nnz_g, const int64_t* __restrict__ g_ptr,
                                const int32_t* __restrict__ g_cols, int64_t* __restrict__ indptr,
                                int32_t* __restrict__ indices) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const long long gb = g_ptr[r];
    const int len = (int)(g_ptr[r + 1] - gb);
    const long long p0 = 2 * gb, p1 = 2 * nnz_g + 2 * gb;
    indptr[r] = p0;
    indptr[n + r] = p1;
    for (int t = 0; t < len; ++t) {
      const int32_t c = g_cols[gb + t];
      indices[p0 + t] = c;
      indices[p0 + len + t] = (int32_t)(n + c);
      indices[p1 + t] = c;
      indices[p1 + len + t] = (int32_t)(n + c);
    }
    if (r == n - 1) indptr[2 * n] = 4 * nnz_g;
  }
}

struct NumArgs {
  const int64_t* adj_ptr;
  const int32_t* adj_cell;
  const uint8_t* adj_loc;
  const int32_t* cnm;
  const int32_t* ccnm;
  const double* coords;
  const double* geom;
  const uint8_t* pos;
  const int64_t* g_ptr;
  const int32_t* g_cols;
  const uint8_t* mask;  // [2n]
  const double* tabs;   // [7][729] reference-cell integrals in global memory
  long long n, nnz_g;
  Coef c;
  double* data;
};

constexpr int NUM_THREADS = 128;

// LANES lanes per node-graph row (LANES = nodes per cell rounded up to a power of two, <= 32)
template <int DIM, int P>
__global__ void __launch_bounds__(NUM_THREADS) k_numeric(const NumArgs g) {
  constexpr int P1 = P + 1;
  constexpr int NPC = DIM == 2 ? P1 * P1 : P1 * P1 * P1;
  constexpr int LANES = NPC <= 4 ? 4 : NPC <= 8 ? 8 : NPC <= 16 ? 16 : 32;
  constexpr int MAXROW = DIM == 2 ? (2 * P + 1) * (2 * P + 1) : (2 * P + 1) * (2 * P + 1) * (2 * P + 1);
  constexpr int ROWS = NUM_THREADS / LANES;   // rows per block
  constexpr int NQ = P1, NV = 1 << DIM, NQP = NPC;   // (P+1)^dim Gauss points = one per lane b < NPC
  __shared__ double sK[ROWS][MAXROW], sM[ROWS][MAXROW];
  __shared__ double sG[ROWS][NQP][7];        // metric at the Gauss points of the current (non-affine) cell
  // reference-cell integrals: constant memory serialises the per-lane addresses (a * NPC + lane), so the tables
  // the affine path reads are staged in shared memory (NPC <= 9) or read through the read-only path (Q2 hex)
  constexpr bool TAB_SMEM = NPC <= 9;
  __shared__ double sT[TAB_SMEM ? 7 * NPC * NPC : 1];
  if (TAB_SMEM) {
    for (int i = threadIdx.x; i < 7 * NPC * NPC; i += NUM_THREADS) sT[i] = g.tabs[(i / (NPC * NPC)) * 729 + i % (NPC * NPC)];
    __syncthreads();
  }
  const int grp = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const long long r = (long long)blockIdx.x * ROWS + grp;
  // the lanes of one group live in one warp: group-wide synchronisation = __syncwarp on the group's mask
  const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
  const bool live = r < g.n;
  for (int t = lane; t < MAXROW; t += LANES) sK[grp][t] = sM[grp][t] = 0.0;
  __syncwarp(gmask);
  const long long e0 = live ? g.adj_ptr[r] : 0, e1 = live ? g.adj_ptr[r + 1] : 0;
  int b0 = 0, b1 = 0, b2 = 0;   // lane as trial node b / as Gauss point q
  if (DIM == 2) { b1 = lane / P1; b2 = lane % P1; }
  else { b0 = lane / (P1 * P1); b1 = (lane / P1) % P1; b2 = lane % P1; }
  // Fast path (every incident cell affine, <= 2^DIM of them: the common case): all loads of the row's cells are
  // issued before any is used -- two memory round trips per row instead of two per incident cell (the sequential
  // loop below ran at 0.22 of the nnz * 12 B model, bound by that latency chain).  The shared-memory accumulation
  // keeps the ascending cell order: same bits as the loop.
  constexpr int NCMAX = 1 << DIM;
  long long e_start = e0;
  if (e1 - e0 <= NCMAX) {
    long long cell[NCMAX];
    int aa[NCMAX];
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) {
      cell[c] = e0 + c < e1 ? (long long)g.adj_cell[e0 + c] : -1;
      aa[c] = e0 + c < e1 ? (int)g.adj_loc[e0 + c] : 0;
    }
    double kabv[NCMAX], mabv[NCMAX];
    int pp[NCMAX];
    bool all_affine = true;
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) {
      kabv[c] = mabv[c] = 0.0;
      pp[c] = 0;
      if (cell[c] >= 0) {
        const double* gm = g.geom + cell[c] * 8;
        all_affine = all_affine && gm[7] != 0.0;
        if (lane < NPC) {
          const int ab = aa[c] * NPC + lane;
          double kab = 0.0;
#pragma unroll
          for (int s = 0; s < 6; ++s) kab = fma(gm[s], TAB_SMEM ? sT[s * NPC * NPC + ab] : __ldg(g.tabs + s * 729 + ab), kab);
          kabv[c] = kab;
          mabv[c] = gm[6] * (TAB_SMEM ? sT[6 * NPC * NPC + ab] : __ldg(g.tabs + 6 * 729 + ab));
          pp[c] = (int)g.pos[(cell[c] * NPC + aa[c]) * NPC + lane];
        }
      }
    }
    if (all_affine) {   // uniform over the group: every lane tests the same cells
#pragma unroll
      for (int c = 0; c < NCMAX; ++c) {
        if (cell[c] >= 0 && lane < NPC) {
          sK[grp][pp[c]] += kabv[c];
          sM[grp][pp[c]] += mabv[c];
        }
        __syncwarp(gmask);
      }
      e_start = e1;   // done
    }
  }
  for (long long e = e_start; e < e1; ++e) {
    const long long cell = g.adj_cell[e];
    const int a = g.adj_loc[e];
    const double* gm = g.geom + cell * 8;
    double kab = 0.0, mab = 0.0;
    if (gm[7] != 0.0) {  // affine cell: constant metric x reference integrals
      if (lane < NPC) {
        const int ab = a * NPC + lane;
#pragma unroll
        for (int s = 0; s < 6; ++s) kab = fma(gm[s], TAB_SMEM ? sT[s * NPC * NPC + ab] : __ldg(g.tabs + s * 729 + ab), kab);
        mab = gm[6] * (TAB_SMEM ? sT[6 * NPC * NPC + ab] : __ldg(g.tabs + 6 * 729 + ab));
      }
    } else {             // general cell: Gauss quadrature of the multilinear map
      if (lane < NQP) {  // lane = Gauss point (q0, q1, q2) = (b0, b1, b2)
        const double wq = (DIM == 2 ? 1.0 : cW[P - 1][b0]) * cW[P - 1][b1] * cW[P - 1][b2];
        double J[3][3], Gq[3][3], dm;
        jacobian_at<DIM, P>(g.coords, g.ccnm + cell * NV, DIM == 2 ? b1 : b0, DIM == 2 ? b2 : b1, b2, J);
        metric_from_J<DIM>(J, wq, Gq, dm);
        double* o = sG[grp][lane];
        o[0] = Gq[0][0]; o[1] = Gq[0][1]; o[2] = Gq[0][2]; o[3] = Gq[1][1]; o[4] = Gq[1][2]; o[5] = Gq[2][2]; o[6] = dm;
      }
      __syncwarp(gmask);
      if (lane < NPC) {
        int a0, a1, a2;
        if (DIM == 2) { a0 = 0; a1 = a / P1; a2 = a % P1; }
        else { a0 = a / (P1 * P1); a1 = (a / P1) % P1; a2 = a % P1; }
        for (int q0 = 0; q0 < (DIM == 2 ? 1 : NQ); ++q0)
          for (int q1 = 0; q1 < NQ; ++q1)
            for (int q2 = 0; q2 < NQ; ++q2) {
              const double* o = sG[grp][DIM == 2 ? q1 * NQ + q2 : (q0 * NQ + q1) * NQ + q2];
              const double Ba0 = DIM == 2 ? 1.0 : cB[P - 1][a0][q0], Da0 = DIM == 2 ? 0.0 : cD[P - 1][a0][q0];
              const double Ba1 = cB[P - 1][a1][q1], Da1 = cD[P - 1][a1][q1];
              const double Ba2 = cB[P - 1][a2][q2], Da2 = cD[P - 1][a2][q2];
              const double Bb0 = DIM == 2 ? 1.0 : cB[P - 1][b0][q0], Db0 = DIM == 2 ? 0.0 : cD[P - 1][b0][q0];
              const double Bb1 = cB[P - 1][b1][q1], Db1 = cD[P - 1][b1][q1];
              const double Bb2 = cB[P - 1][b2][q2], Db2 = cD[P - 1][b2][q2];
              double gta[3], gtb[3];
              if (DIM == 2) {
                gta[0] = Da1 * Ba2; gta[1] = Ba1 * Da2; gta[2] = 0.0;
                gtb[0] = Db1 * Bb2; gtb[1] = Bb1 * Db2; gtb[2] = 0.0;
              } else {
                gta[0] = Da0 * Ba1 * Ba2; gta[1] = Ba0 * Da1 * Ba2; gta[2] = Ba0 * Ba1 * Da2;
                gtb[0] = Db0 * Bb1 * Bb2; gtb[1] = Bb0 * Db1 * Bb2; gtb[2] = Bb0 * Bb1 * Db2;
              }
              const double Gt0 = o[0] * gta[0] + o[1] * gta[1] + o[2] * gta[2];
              const double Gt1 = o[1] * gta[0] + o[3] * gta[1] + o[4] * gta[2];
              const double Gt2 = o[2] * gta[0] + o[4] * gta[1] + o[5] * gta[2];
              kab += Gt0 * gtb[0] + Gt1 * gtb[1] + Gt2 * gtb[2];
              mab += (Ba0 * Ba1 * Ba2) * (Bb0 * Bb1 * Bb2) * o[6];
            }
      }
    }
    if (lane < NPC) {   // the NPC slots of one (cell, a) are distinct: conflict-free, cells in ascending order
      const int p = g.pos[(cell * NPC + a) * NPC + lane];
      sK[grp][p] += kab;
      sM[grp][p] += mab;
    }
    __syncwarp(gmask);
  }
  if (!live) return;
  // write the four block segments of rows r (field 0) and n + r (field 1): consecutive lanes, consecutive entries
  const long long gb = g.g_ptr[r];
  const int len = (int)(g.g_ptr[r + 1] - gb);
  const long long n = g.n;
  const bool m0 = g.mask[r] != 0, m1 = g.mask[n + r] != 0;
  const long long p0 = 2 * gb, p1 = 2 * g.nnz_g + 2 * gb;
  constexpr int NW = (MAXROW + LANES - 1) / LANES;   // entries per lane: column loads first, then the dependent masks
  long long cc[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) cc[w] = lane + w * LANES < len ? (long long)g.g_cols[gb + lane + w * LANES] : -1;
  bool c0v[NW], c1v[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    c0v[w] = cc[w] >= 0 && g.mask[cc[w]] != 0;
    c1v[w] = cc[w] >= 0 && g.mask[n + cc[w]] != 0;
  }
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const int t = lane + w * LANES;
    if (t >= len) continue;
    const long long c = cc[w];
    const bool c0 = c0v[w], c1 = c1v[w];
    const double K = sK[grp][t], M = sM[grp][t];
    const bool diag = (c == r);
    double v00 = g.c.cK[0][0] * K + g.c.cM[0][0] * M, v01 = g.c.cK[0][1] * K + g.c.cM[0][1] * M;
    double v10 = g.c.cK[1][0] * K + g.c.cM[1][0] * M, v11 = g.c.cK[1][1] * K + g.c.cM[1][1] * M;
    if (m0 || c0) v00 = (m0 && diag) ? 1.0 : 0.0;
    if (m0 || c1) v01 = 0.0;
    if (m1 || c0) v10 = 0.0;
    if (m1 || c1) v11 = (m1 && diag) ? 1.0 : 0.0;
    g.data[p0 + t] = v00;
    g.data[p0 + len + t] = v01;
    g.data[p1 + t] = v10;
    g.data[p1 + len + t] = v11;
  }
}

// one diagonal-or-coupling block (row field fr, column field fc) of the assembled matrix as a scalar-space CSR
// matrix: pattern = the node graph, values gathered from the monolithic storage
__global__ void k_extract_block(long long n, long long nnz_g, const int64_t* __restrict__ g_ptr,
                                const double* __restrict__ data, int fr, int fc, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const long long gb = g_ptr[r];
    const int len = (int)(g_ptr[r + 1] - gb);
    const double* src = data + (fr ? 2 * nnz_g : 0) + 2 * gb + (fc ? len : 0);
    for (int t = 0; t < len; ++t) out[gb + t] = src[t];
  }
}

// y = A x, 8 lanes per row (rows hold <= 54 / 250 entries); optional fused <x, y> partials
constexpr int SPMV_LANES = 8;
constexpr int SPMV_THREADS = 256;

__global__ void __launch_bounds__(SPMV_THREADS) k_spmv(long long n_rows, const int64_t* __restrict__ indptr,
                                                        const int32_t* __restrict__ indices,
                                                        const double* __restrict__ data, const double* __restrict__ x,
                                                        double* __restrict__ y, double* __restrict__ dot_partials,
                                                        const double* __restrict__ skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0.0) return;
  __shared__ double red[SPMV_THREADS / 32];
  const int lane = threadIdx.x % SPMV_LANES;
  const long long row = (blockIdx.x * (long long)SPMV_THREADS + threadIdx.x) / SPMV_LANES;
  double s = 0.0, xr = 0.0;
  if (row < n_rows) {
    const long long b = indptr[row], e = indptr[row + 1];
    for (long long t = b + lane; t < e; t += SPMV_LANES) s = fma(data[t], __ldg(&x[indices[t]]), s);
  }
#pragma unroll
  for (int o = SPMV_LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  double d = 0.0;
  if (row < n_rows && lane == 0) {
    y[row] = s;
    xr = x[row];
    d = xr * s;
  }
  if (dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < SPMV_THREADS / 32; ++w) t += red[w];
      dot_partials[blockIdx.x] = t;
    }
  }
}

// host: reference-cell integral tables from 1-D Gauss quadrature (exact for the degrees involved)
int upload_reference_tables(dpp_context* ctx) {
  const int dim = ctx->dim, p = ctx->degree, p1 = p + 1;
  const int npc = ctx->npc;
  // 1-D integrals with 3 Gauss points on [0,1] (exact to degree 5 >= 2p)
  const double s = std::sqrt(0.6);
  const double xq[3] = {0.5 * (1 - s), 0.5, 0.5 * (1 + s)}, wq[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  double M1[3][3] = {}, K1[3][3] = {}, C1[3][3] = {};
  for (int q = 0; q < 3; ++q) {
    const double x = xq[q];
    double B[3] = {0, 0, 0}, D[3] = {0, 0, 0};
    if (p == 1) { B[0] = 1 - x; B[1] = x; D[0] = -1; D[1] = 1; }
    else {
      B[0] = 2 * (x - 0.5) * (x - 1); B[1] = -4 * x * (x - 1); B[2] = 2 * x * (x - 0.5);
      D[0] = 4 * x - 3; D[1] = -8 * x + 4; D[2] = 4 * x - 1;
    }
    for (int a = 0; a < p1; ++a)
      for (int b = 0; b < p1; ++b) {
        M1[a][b] += wq[q] * B[a] * B[b];
        K1[a][b] += wq[q] * D[a] * D[b];
        C1[a][b] += wq[q] * D[a] * B[b];  // int phi'_a phi_b
      }
  }
  std::vector<double> TK(6 * 729, 0.0), TM(729, 0.0);
  for (int a = 0; a < npc; ++a)
    for (int b = 0; b < npc; ++b) {
      const int ab = a * npc + b;
      if (dim == 3) {
        const int a0 = a / (p1 * p1), a1 = (a / p1) % p1, a2 = a % p1;
        const int b0 = b / (p1 * p1), b1 = (b / p1) % p1, b2 = b % p1;
        TM[ab] = M1[a0][b0] * M1[a1][b1] * M1[a2][b2];
        TK[0 * 729 + ab] = K1[a0][b0] * M1[a1][b1] * M1[a2][b2];
        TK[3 * 729 + ab] = M1[a0][b0] * K1[a1][b1] * M1[a2][b2];
        TK[5 * 729 + ab] = M1[a0][b0] * M1[a1][b1] * K1[a2][b2];
        TK[1 * 729 + ab] = (C1[a0][b0] * C1[b1][a1] + C1[b0][a0] * C1[a1][b1]) * M1[a2][b2];
        TK[2 * 729 + ab] = (C1[a0][b0] * C1[b2][a2] + C1[b0][a0] * C1[a2][b2]) * M1[a1][b1];
        TK[4 * 729 + ab] = (C1[a1][b1] * C1[b2][a2] + C1[b1][a1] * C1[a2][b2]) * M1[a0][b0];
      } else {
        const int a0 = a / p1, a1 = a % p1, b0 = b / p1, b1 = b % p1;
        TM[ab] = M1[a0][b0] * M1[a1][b1];
        TK[0 * 729 + ab] = K1[a0][b0] * M1[a1][b1];
        TK[3 * 729 + ab] = M1[a0][b0] * K1[a1][b1];
        TK[1 * 729 + ab] = C1[a0][b0] * C1[b1][a1] + C1[b0][a0] * C1[a1][b1];
      }
    }
  DPP_CUDA(cudaMemcpyToSymbol(cTK, TK.data(), sizeof(double) * 6 * 729));
  DPP_CUDA(cudaMemcpyToSymbol(cTM, TM.data(), sizeof(double) * 729));
  if (ctx->csr) {
    if (!ctx->csr->tabs) DPP_CHECK(dev_alloc(ctx, &ctx->csr->tabs, 7 * 729));
    DPP_CUDA(cudaMemcpy(ctx->csr->tabs, TK.data(), sizeof(double) * 6 * 729, cudaMemcpyHostToDevice));
    DPP_CUDA(cudaMemcpy(ctx->csr->tabs + 6 * 729, TM.data(), sizeof(double) * 729, cudaMemcpyHostToDevice));
  }
  return fe_upload_tables(ctx);
}

int blocks_for(const dpp_context* ctx, long long n, int threads) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)ctx->sm_count * 32));
}

int symbolic(dpp_context* ctx, CsrMatrix* A) {
  const long long n = ctx->n_nodes, nc = ctx->n_cells;
  const int npc = ctx->npc;
  A->n_nodes = n;
  int64_t* counts = nullptr;
  DPP_CHECK(dev_alloc(ctx, &counts, n));
  DPP_CHECK(dev_alloc(ctx, &A->g_ptr, n + 1));
  const int blocks = blocks_for(ctx, n, 128);
#define GRAPH_CASE(NPC, MAXC, COUNTS, PTR, COLS)                                                        \
  k_graph_rows<NPC, MAXC><<<blocks, 128, 0, ctx->stream>>>(n, ctx->d_adj_ptr, ctx->d_adj_cell, ctx->d_cnm, COUNTS, \
                                                           PTR, COLS)
  auto graph = [&](int64_t* cnt, const int64_t* ptr, int32_t* cols) {
    if (npc == 4) GRAPH_CASE(4, 16, cnt, ptr, cols);
    else if (npc == 9) GRAPH_CASE(9, 36, cnt, ptr, cols);
    else if (npc == 8) GRAPH_CASE(8, 64, cnt, ptr, cols);
    else GRAPH_CASE(27, 216, cnt, ptr, cols);
    ctx->launches++;
  };
#undef GRAPH_CASE
  graph(counts, nullptr, nullptr);
  DPP_CUDA(cudaGetLastError());
  {
    const long long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* sums = nullptr;
    DPP_CHECK(dev_alloc(ctx, &sums, nb + 1));
    k_scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(n, counts, sums);
    k_scan_sums<<<1, SCAN_THREADS, 0, ctx->stream>>>(nb, sums);
    k_scan_finish<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(n, counts, sums, A->g_ptr);
    ctx->launches += 3;
    DPP_CUDA(cudaGetLastError());
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(sums);
    ctx->device_bytes -= (int64_t)sizeof(int64_t) * (nb + 1);
  }
  int64_t nnz_g = 0;
  DPP_CUDA(cudaMemcpyAsync(&nnz_g, A->g_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(counts);
  ctx->device_bytes -= (int64_t)sizeof(int64_t) * n;
  A->nnz_g = nnz_g;
  A->nnz = 4 * nnz_g;
  if (4 * nnz_g + 2 * n >= (1LL << 40)) {
    ctx->set_error("csr: matrix too large");
    return DPP_ERR_INVALID;
  }
  DPP_CHECK(dev_alloc(ctx, &A->g_cols, nnz_g));
  graph(nullptr, A->g_ptr, A->g_cols);
  DPP_CUDA(cudaGetLastError());
  DPP_CHECK(dev_alloc(ctx, &A->pos, nc * npc * npc));
  {
    const int b2 = blocks_for(ctx, nc * npc, 128);
    if (npc == 4) k_positions<4><<<b2, 128, 0, ctx->stream>>>(nc, ctx->d_cnm, A->g_ptr, A->g_cols, A->pos);
    else if (npc == 9) k_positions<9><<<b2, 128, 0, ctx->stream>>>(nc, ctx->d_cnm, A->g_ptr, A->g_cols, A->pos);
    else if (npc == 8) k_positions<8><<<b2, 128, 0, ctx->stream>>>(nc, ctx->d_cnm, A->g_ptr, A->g_cols, A->pos);
    else k_positions<27><<<b2, 128, 0, ctx->stream>>>(nc, ctx->d_cnm, A->g_ptr, A->g_cols, A->pos);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
  }
  DPP_CHECK(dev_alloc(ctx, &A->indptr, 2 * n + 1));
  DPP_CHECK(dev_alloc(ctx, &A->indices, 4 * nnz_g));
  DPP_CHECK(dev_alloc(ctx, &A->data, 4 * nnz_g));
  k_block_pattern<<<blocks, 128, 0, ctx->stream>>>(n, nnz_g, A->g_ptr, A->g_cols, A->indptr, A->indices);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

int numeric(dpp_context* ctx, CsrMatrix* A) {
  NumArgs g{};
  g.adj_ptr = ctx->d_adj_ptr; g.adj_cell = ctx->d_adj_cell; g.adj_loc = ctx->d_adj_loc;
  g.cnm = ctx->d_cnm; g.ccnm = ctx->d_ccnm; g.coords = ctx->d_coords; g.geom = ctx->d_cell_geom;
  g.pos = A->pos; g.g_ptr = A->g_ptr; g.g_cols = A->g_cols; g.mask = ctx->d_mask; g.tabs = A->tabs;
  g.n = ctx->n_nodes; g.nnz_g = A->nnz_g; g.c = dpp_coef(ctx); g.data = A->data;
  const int dim = ctx->dim, p = ctx->degree;
  const int npc = ctx->npc;
  const int lanes = npc <= 4 ? 4 : npc <= 8 ? 8 : npc <= 16 ? 16 : 32;
  const int rows = NUM_THREADS / lanes;
  const unsigned blocks = (unsigned)((ctx->n_nodes + rows - 1) / rows);
  if (dim == 2 && p == 1) k_numeric<2, 1><<<blocks, NUM_THREADS, 0, ctx->stream>>>(g);
  else if (dim == 2 && p == 2) k_numeric<2, 2><<<blocks, NUM_THREADS, 0, ctx->stream>>>(g);
  else if (dim == 3 && p == 1) k_numeric<3, 1><<<blocks, NUM_THREADS, 0, ctx->stream>>>(g);
  else k_numeric<3, 2><<<blocks, NUM_THREADS, 0, ctx->stream>>>(g);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  A->numeric_valid = true;
  return DPP_OK;
}

}  // namespace

int csr_assemble(dpp_context* ctx, int64_t* nnz) {
  if (!ctx->general_ready) {  // adjacency + per-cell geometry live with the general family
    std::vector<int32_t> cnm((size_t)ctx->n_cells * ctx->npc);
    DPP_CUDA(cudaMemcpy(cnm.data(), ctx->d_cnm, sizeof(int32_t) * cnm.size(), cudaMemcpyDeviceToHost));
    DPP_CHECK(general_setup(ctx, cnm.data()));
  }
  if (!ctx->csr) {
    ctx->csr = new CsrMatrix();
    DPP_CHECK(upload_reference_tables(ctx));
    DPP_CHECK(symbolic(ctx, ctx->csr));
  }
  DPP_CHECK(numeric(ctx, ctx->csr));
  if (nnz) *nnz = ctx->csr->nnz;
  return DPP_OK;
}

int csr_export(dpp_context* ctx, int64_t* indptr, int32_t* indices, double* data) {
  CsrMatrix* A = ctx->csr;
  if (!A || !A->numeric_valid) {
    ctx->set_error("dpp_get_csr: call dpp_assemble_csr first");
    return DPP_ERR_STATE;
  }
  const long long n2 = 2 * A->n_nodes;
  if (indptr) DPP_CUDA(cudaMemcpyAsync(indptr, A->indptr, sizeof(int64_t) * (n2 + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (indices) DPP_CUDA(cudaMemcpyAsync(indices, A->indices, sizeof(int32_t) * A->nnz, cudaMemcpyDeviceToHost, ctx->stream));
  if (data) DPP_CUDA(cudaMemcpyAsync(data, A->data, sizeof(double) * A->nnz, cudaMemcpyDeviceToHost, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DPP_OK;
}

int csr_spmv(dpp_context* ctx, const double* x, double* y, double* dot_partials, int* n_partial_blocks,
             const double* skip_flag) {
  CsrMatrix* A = ctx->csr;
  if (!A || !A->numeric_valid) {
    ctx->set_error("csr_spmv: matrix not assembled");
    return DPP_ERR_STATE;
  }
  if (ctx->world > 1) {
    ctx->set_error("assembled operator mode is single-GPU (use the matrix-free operator for slab runs)");
    return DPP_ERR_INVALID;
  }
  const long long n_rows = 2 * A->n_nodes;
  const long long threads = n_rows * SPMV_LANES;
  const int blocks = (int)((threads + SPMV_THREADS - 1) / SPMV_THREADS);
  if (dot_partials != nullptr && blocks > kMaxPartialBlocks * kMaxDotWidth) {
    ctx->set_error("csr_spmv: reduction scratch too small for the fused dot");
    return DPP_ERR_INVALID;
  }
  k_spmv<<<blocks, SPMV_THREADS, 0, ctx->stream>>>(n_rows, A->indptr, A->indices, A->data, x, y, dot_partials, skip_flag);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (n_partial_blocks) *n_partial_blocks = blocks;
  return DPP_OK;
}

// block (fr, fc) as a scalar-space CSR triplet: indptr [n+1], indices / data [nnz / 4]
int csr_export_block(dpp_context* ctx, int fr, int fc, int64_t* indptr, int32_t* indices, double* data) {
  CsrMatrix* A = ctx->csr;
  if (!A || !A->numeric_valid) {
    ctx->set_error("dpp_get_csr_block_host: call dpp_assemble_csr first");
    return DPP_ERR_STATE;
  }
  const long long n = A->n_nodes;
  if (indptr) DPP_CUDA(cudaMemcpyAsync(indptr, A->g_ptr, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (indices) DPP_CUDA(cudaMemcpyAsync(indices, A->g_cols, sizeof(int32_t) * A->nnz_g, cudaMemcpyDeviceToHost, ctx->stream));
  if (data) {
    double* tmp = nullptr;
    DPP_CHECK(dev_alloc(ctx, &tmp, A->nnz_g));
    k_extract_block<<<blocks_for(ctx, n, 128), 128, 0, ctx->stream>>>(n, A->nnz_g, A->g_ptr, A->data, fr, fc, tmp);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
    DPP_CUDA(cudaMemcpyAsync(data, tmp, sizeof(double) * A->nnz_g, cudaMemcpyDeviceToHost, ctx->stream));
    DPP_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    ctx->device_bytes -= (int64_t)sizeof(double) * A->nnz_g;
  }
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DPP_OK;
}

// measurement: device milliseconds of the symbolic phase (pattern + scatter permutation, rebuilt from scratch)
// and of the numeric phase (mean of `reps` fills)
int csr_time_phases(dpp_context* ctx, int reps, double* symbolic_ms, double* numeric_ms, int64_t* nnz) {
  if (!ctx->general_ready) {
    std::vector<int32_t> cnm((size_t)ctx->n_cells * ctx->npc);
    DPP_CUDA(cudaMemcpy(cnm.data(), ctx->d_cnm, sizeof(int32_t) * cnm.size(), cudaMemcpyDeviceToHost));
    DPP_CHECK(general_setup(ctx, cnm.data()));
  }
  csr_destroy(ctx);
  ctx->csr = new CsrMatrix();
  DPP_CHECK(upload_reference_tables(ctx));
  cudaEvent_t ev[4];
  for (auto& e : ev) DPP_CUDA(cudaEventCreate(&e));
  DPP_CUDA(cudaEventRecord(ev[0], ctx->stream));
  DPP_CHECK(symbolic(ctx, ctx->csr));
  DPP_CUDA(cudaEventRecord(ev[1], ctx->stream));
  DPP_CHECK(numeric(ctx, ctx->csr));   // warm-up
  DPP_CUDA(cudaEventRecord(ev[2], ctx->stream));
  for (int i = 0; i < std::max(reps, 1); ++i) DPP_CHECK(numeric(ctx, ctx->csr));
  DPP_CUDA(cudaEventRecord(ev[3], ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  float a = 0, b = 0;
  cudaEventElapsedTime(&a, ev[0], ev[1]);
  cudaEventElapsedTime(&b, ev[2], ev[3]);
  for (auto& e : ev) cudaEventDestroy(e);
  if (symbolic_ms) *symbolic_ms = a;
  if (numeric_ms) *numeric_ms = b / std::max(reps, 1);
  if (nnz) *nnz = ctx->csr->nnz;
  return DPP_OK;
}

// parameters / Dirichlet data changed: values are stale, the pattern is not
void csr_invalidate(dpp_context* ctx) {
  if (ctx->csr) ctx->csr->numeric_valid = false;
}

void csr_destroy(dpp_context* ctx) {
  CsrMatrix* A = ctx->csr;
  if (!A) return;
  void* p[] = {A->g_ptr, A->g_cols, A->pos, A->indptr, A->indices, A->data, A->tabs};
  for (void* q : p)
    if (q) cudaFree(q);
  delete A;
  ctx->csr = nullptr;
}

}  // namespace dpp

namespace dpp {
bool csr_valid(const dpp_context* ctx) { return ctx->csr != nullptr && ctx->csr->numeric_valid; }
}  // namespace dpp
