// Matrix-free DPP operator on GENERAL hexahedral Q1 meshes (arbitrary numbering, distorted cells): the
// element-based kernel of the north star -- "sum-factorised tensor-product quadrature, shared-memory staged
// element gathers, coalesced loads, fused Dirichlet masking".  Replaces the TSFC element kernels + PyOP2 cell
// loop + MatMult below solver.py:66-71 for meshes that are not tensor grids (those run apply_structured*.cu).
//
// Data model (what the host layer pulls from Firedrake): vertex coordinates + cell -> node maps.
//
// Setup, once per mesh (host, integer work): cells are sorted along a Morton curve of their centroids and cut
// into CELL BLOCKS of <= CB consecutive cells.  Every block owns a fixed-stride record in every per-block array
// (NLMAX node slots, CB cell slots), so no address in the hot kernel depends on a loaded value:
//   blk_nodes [nb][NLMAX]      the block's distinct nodes, ascending (-1 = padding)
//   cblk      [nb][3][NLMAX]   their vertex coordinates, staged ONCE at setup (constant per mesh)
//   cell_loc  [nb][CB][8]      16-bit local connectivity (0xFFFF = no cell)
//   ladj_ptr / ladj            the transposed connectivity: local node -> incident (cell, corner) pairs
//   nd_ptr / nd_slot           node -> the (block, local) slots it appears in
//
// Apply = three kernels, no atomics, bitwise reproducible:
//  0. k_cells_stage: xblk[f][b][l] = x_f[blk_nodes[b][l]] (masked) -- a streaming permutation that takes the
//     one indexed gather of the apply out of the compute kernel (latency-tolerant there, exposed here).
//  1. k_cells_q1: one CTA per cell block, one thread per cell.  All loads of a CTA are issued up front from
//     block-contiguous, index-free addresses (one memory round trip) into shared memory.  Each thread evaluates
//     its cell: the metric G = w |J| J^-1 J^-T at the 2x2x2 Gauss points from sum-factorised derivatives of the
//     trilinear map (kept in shared memory, shared by both fields), then per output field y_f = K a_f + M b_f
//     with a_f = sum_g cK[f][g] x_g, b_f = sum_g cM[f][g] x_g: sum-factorised forward interpolation, G applied
//     at the Gauss points, transposed sum-factorised contraction.  The element vectors go to shared memory;
//     one thread per LOCAL NODE then adds the entries of its incident cells in a fixed order (row-owner gather
//     inside the block: no conflicts, no colouring, three barriers per block) and stores the block's partial
//     result into its own slot range (coalesced).
//  2. k_cells_gather: one thread per node adds the node's <= 8 block partials in ascending slot order, applies
//     the Dirichlet row semantics and the fused <x, y> partial sums.
//
// Bound: fp64 FMA pipe, not HBM.  A distorted trilinear cell needs ~1260 fp64 instructions (geometry 610,
// fields 2 x 326), an affine one ~720; at 64 fp64 lanes/clk/SM that is 1.17 ms (0.67 ms affine) for the 256^3
// mesh against 0.23 ms for its 1.52 GB of algorithmic traffic (58 B/node + 32 B/cell): DESIGN.md 4.3.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "cg_device.cuh"
#include "dpp_internal.cuh"

namespace dpp {

constexpr int CB = 128;       // cells per block = threads per CTA
constexpr int NLMAX = 256;    // node slots per block (8 x 4 x 4 cells have 225; a block that needs more is halved)
constexpr int ROUNDS = NLMAX / CB;

struct CellBlocks {
  int n_blocks = 0;
  long long total_slots = 0;          // n_blocks * NLMAX
  bool same_numbering = true;         // coordinate nodes == pressure nodes
  int32_t* blk_nodes = nullptr;       // [nb][NLMAX]
  double* cblk = nullptr;             // [nb][3][NLMAX]
  uint16_t* cell_loc = nullptr;       // [nb][CB][8] local pressure-node ids
  uint16_t* cell_cloc = nullptr;      // [nb][CB][8] local coordinate-node ids (aliases cell_loc when same_numbering)
  uint16_t* ladj_ptr = nullptr;       // [nb][NLMAX + 8]: nl + 1 offsets, [NLMAX + 1] = nl
  uint16_t* ladj = nullptr;           // [nb][CB * 8]: (cell_in_block * 8 + corner), grouped by local node
  uint8_t* blk_affine = nullptr;      // [nb]
  int64_t* nd_ptr = nullptr;          // [n_nodes + 1]
  int32_t* nd_slot = nullptr;         // [sum of node counts] ascending per node
  double* xblk = nullptr;             // [2][total_slots] staged input
  double* ypart = nullptr;            // [2][total_slots] block partial results
  long long affine_cells = 0, used_slots = 0;
};

namespace {

constexpr int LPTR = NLMAX + 8;   // uint16 entries of one block's ladj_ptr record

// Gauss points of [0,1] (2-point rule); the weights are 1/2 each, w_q = 1/8 per 3-D point
__device__ constexpr double kG0 = 0.21132486540518711775;   // (1 - 1/sqrt(3)) / 2
__device__ constexpr double kG1 = 0.78867513459481288225;

// sum-factorised derivatives of a trilinear field at the 2x2x2 Gauss points.
//   D0[q1][q2] = d/dxi0 (independent of q0), D1[q0][q2] = d/dxi1, D2[q0][q1] = d/dxi2
template <bool VAL>
__device__ __forceinline__ void forward_q1(const double (&u)[8], double (&D0)[4], double (&D1)[4], double (&D2)[4],
                                           double (&val)[8]) {
  const double g[2] = {kG0, kG1};
  double d2[4], v[4][2];   // [a0*2+a1][q2]
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    d2[p] = u[2 * p + 1] - u[2 * p];
    v[p][0] = fma(d2[p], g[0], u[2 * p]);
    v[p][1] = fma(d2[p], g[1], u[2 * p]);
  }
  double dv1[2][2], w[2][2][2];   // dv1[a0][q2], w[a0][q1][q2]
#pragma unroll
  for (int a0 = 0; a0 < 2; ++a0)
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2) {
      dv1[a0][q2] = v[2 * a0 + 1][q2] - v[2 * a0][q2];
      w[a0][0][q2] = fma(dv1[a0][q2], g[0], v[2 * a0][q2]);
      w[a0][1][q2] = fma(dv1[a0][q2], g[1], v[2 * a0][q2]);
    }
  double e2[2][2];   // [a0][q1]
#pragma unroll
  for (int a0 = 0; a0 < 2; ++a0) {
    const double de = d2[2 * a0 + 1] - d2[2 * a0];
    e2[a0][0] = fma(de, g[0], d2[2 * a0]);
    e2[a0][1] = fma(de, g[1], d2[2 * a0]);
  }
#pragma unroll
  for (int q1 = 0; q1 < 2; ++q1)
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2) {
      D0[q1 * 2 + q2] = w[1][q1][q2] - w[0][q1][q2];
      if (VAL) {
        val[0 * 4 + q1 * 2 + q2] = fma(D0[q1 * 2 + q2], g[0], w[0][q1][q2]);
        val[1 * 4 + q1 * 2 + q2] = fma(D0[q1 * 2 + q2], g[1], w[0][q1][q2]);
      }
    }
#pragma unroll
  for (int q2 = 0; q2 < 2; ++q2) {
    const double dd = dv1[1][q2] - dv1[0][q2];
    D1[0 * 2 + q2] = fma(dd, g[0], dv1[0][q2]);
    D1[1 * 2 + q2] = fma(dd, g[1], dv1[0][q2]);
  }
#pragma unroll
  for (int q1 = 0; q1 < 2; ++q1) {
    const double dd = e2[1][q1] - e2[0][q1];
    D2[0 * 2 + q1] = fma(dd, g[0], e2[0][q1]);
    D2[1 * 2 + q1] = fma(dd, g[1], e2[0][q1]);
  }
}

// transpose of forward_q1: y_a += sum_q [ mh_q N_a(q) + sum_d F[d]_q dN_a/dxi_d(q) ],  q = q0*4 + q1*2 + q2
__device__ __forceinline__ void backward_q1(const double (&F0)[8], const double (&F1)[8], const double (&F2)[8],
                                            const double (&mh)[8], double (&y)[8]) {
  double P[2][4], Q1[2][4], Q2[2][4];   // [a0][q1*2+q2]
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double ma = mh[r], mb = mh[4 + r];
    const double T = fma(mb, kG1, ma * kG0), S = ma + mb, f0 = F0[r] + F0[4 + r];
    P[1][r] = T + f0;
    P[0][r] = (S - T) - f0;
    const double T1 = fma(F1[4 + r], kG1, F1[r] * kG0);
    Q1[1][r] = T1;
    Q1[0][r] = (F1[r] + F1[4 + r]) - T1;
    const double T2 = fma(F2[4 + r], kG1, F2[r] * kG0);
    Q2[1][r] = T2;
    Q2[0][r] = (F2[r] + F2[4 + r]) - T2;
  }
  double R[2][2][2], S2[2][2][2];   // [a0][a1][q2]
#pragma unroll
  for (int a0 = 0; a0 < 2; ++a0)
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2) {
      const double pa = P[a0][0 * 2 + q2], pb = P[a0][1 * 2 + q2];
      const double T = fma(pb, kG1, pa * kG0), qs = Q1[a0][q2] + Q1[a0][2 + q2];
      R[a0][1][q2] = T + qs;
      R[a0][0][q2] = ((pa + pb) - T) - qs;
      const double sa = Q2[a0][q2], sb = Q2[a0][2 + q2];
      const double Ts = fma(sb, kG1, sa * kG0);
      S2[a0][1][q2] = Ts;
      S2[a0][0][q2] = (sa + sb) - Ts;
    }
#pragma unroll
  for (int a0 = 0; a0 < 2; ++a0)
#pragma unroll
    for (int a1 = 0; a1 < 2; ++a1) {
      const double ra = R[a0][a1][0], rb = R[a0][a1][1];
      const double T = fma(rb, kG1, ra * kG0), ss = S2[a0][a1][0] + S2[a0][a1][1];
      y[a0 * 4 + a1 * 2 + 1] = T + ss;
      y[a0 * 4 + a1 * 2 + 0] = ((ra + rb) - T) - ss;
    }
}

// G = (w / |det|) adj adj^T (6 unique entries) and dm = w |det| from the Jacobian columns j0, j1, j2
__device__ __forceinline__ void metric_q1(const double (&j0)[3], const double (&j1)[3], const double (&j2)[3], double w,
                                          double (&G)[7]) {
  double A[3][3];
  A[0][0] = j1[1] * j2[2] - j1[2] * j2[1]; A[0][1] = j1[2] * j2[0] - j1[0] * j2[2]; A[0][2] = j1[0] * j2[1] - j1[1] * j2[0];
  A[1][0] = j2[1] * j0[2] - j2[2] * j0[1]; A[1][1] = j2[2] * j0[0] - j2[0] * j0[2]; A[1][2] = j2[0] * j0[1] - j2[1] * j0[0];
  A[2][0] = j0[1] * j1[2] - j0[2] * j1[1]; A[2][1] = j0[2] * j1[0] - j0[0] * j1[2]; A[2][2] = j0[0] * j1[1] - j0[1] * j1[0];
  const double det = j0[0] * A[0][0] + j0[1] * A[0][1] + j0[2] * A[0][2];
  const double ad = fabs(det);
  const double s = w * __drcp_rn(ad);
  G[0] = s * (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]);
  G[1] = s * (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]);
  G[2] = s * (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]);
  G[3] = s * (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
  G[4] = s * (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]);
  G[5] = s * (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
  G[6] = w * ad;
}

struct StageArgs {
  const int32_t* blk_nodes;
  long long total_slots;
  const double* x[2];
  const uint8_t* in_mask[2];
  double* xblk;
  const double* skip_flag;
};

// xblk[f][slot] = x_f[node(slot)] with eliminated columns zeroed; padding slots get 0
template <int NF>
__global__ void __launch_bounds__(VT) k_cells_stage(const StageArgs g) {
  if (g.skip_flag != nullptr && *g.skip_flag != 0.0) return;
  constexpr int U = 4;   // slots per thread and trip: all index loads first, then the dependent gathers
  const long long stride = (long long)gridDim.x * VT;
  for (long long s0 = (long long)blockIdx.x * VT + threadIdx.x; s0 < g.total_slots; s0 += U * stride) {
    int node[U];
#pragma unroll
    for (int u = 0; u < U; ++u) node[u] = s0 + u * stride < g.total_slots ? g.blk_nodes[s0 + u * stride] : -1;
    double v[NF][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        v[f][u] = 0.0;
        if (node[u] >= 0) {
          v[f][u] = g.x[f][node[u]];
          if (g.in_mask[f] != nullptr && g.in_mask[f][node[u]]) v[f][u] = 0.0;
        }
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (s0 + u * stride < g.total_slots)
#pragma unroll
        for (int f = 0; f < NF; ++f) g.xblk[f * g.total_slots + s0 + u * stride] = v[f][u];
  }
}

struct CellArgs {
  const double* cblk;
  const uint16_t* cell_loc;
  const uint16_t* cell_cloc;
  const uint16_t* ladj_ptr;
  const uint16_t* ladj;
  const uint8_t* blk_affine;
  const double* xblk;
  Coef c;
  double* ypart;
  long long total_slots;
  const double* skip_flag;
};

__device__ __forceinline__ void unpack8(const uint4& pk, int (&v)[8]) {
  v[0] = pk.x & 0xffff; v[1] = pk.x >> 16; v[2] = pk.y & 0xffff; v[3] = pk.y >> 16;
  v[4] = pk.z & 0xffff; v[5] = pk.z >> 16; v[6] = pk.w & 0xffff; v[7] = pk.w >> 16;
}

template <int NF>
__global__ void __launch_bounds__(CB, 3) k_cells_q1(const CellArgs g) {
  extern __shared__ __align__(16) double smem[];
  if (g.skip_flag != nullptr && *g.skip_flag != 0.0) return;
  double* xs = smem;                       // [NF][NLMAX]
  double* cs = xs + NF * NLMAX;            // [3][NLMAX]
  double* sG = cs + 3 * NLMAX;             // [8 qp][7][CB] metric; reused as [NF][8][CB] element vectors
  uint16_t* sla = reinterpret_cast<uint16_t*>(sG + 8 * 7 * CB);   // [CB * 8]
  uint16_t* slp = sla + CB * 8;                                   // [LPTR]
  const int b = blockIdx.x, t = threadIdx.x;
  const long long slot0 = (long long)b * NLMAX;
  // ---- every load of the CTA, issued back to back from index-free addresses (one memory round trip)
  const uint4 pk_loc = *reinterpret_cast<const uint4*>(g.cell_loc + ((size_t)b * CB + t) * 8);
  const uint4 pk_cl = *reinterpret_cast<const uint4*>(g.cell_cloc + ((size_t)b * CB + t) * 8);
  const uint4 pk_la = *reinterpret_cast<const uint4*>(g.ladj + ((size_t)b * CB + t) * 8);
  double xv[NF][ROUNDS], cv[3][ROUNDS];
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
#pragma unroll
    for (int f = 0; f < NF; ++f) xv[f][k] = g.xblk[f * g.total_slots + slot0 + t + k * CB];
#pragma unroll
    for (int d = 0; d < 3; ++d) cv[d][k] = g.cblk[((size_t)b * 3 + d) * NLMAX + t + k * CB];
  }
  uint16_t lpv[(LPTR + CB - 1) / CB];
#pragma unroll
  for (int k = 0; k < (LPTR + CB - 1) / CB; ++k) lpv[k] = t + k * CB < LPTR ? g.ladj_ptr[(size_t)b * LPTR + t + k * CB] : 0;
  const bool affine = g.blk_affine[b] != 0;
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
#pragma unroll
    for (int f = 0; f < NF; ++f) xs[f * NLMAX + t + k * CB] = xv[f][k];
#pragma unroll
    for (int d = 0; d < 3; ++d) cs[d * NLMAX + t + k * CB] = cv[d][k];
  }
  *reinterpret_cast<uint4*>(sla + t * 8) = pk_la;
#pragma unroll
  for (int k = 0; k < (LPTR + CB - 1) / CB; ++k)
    if (t + k * CB < LPTR) slp[t + k * CB] = lpv[k];
  __syncthreads();
  const bool live = (pk_loc.x & 0xffff) != 0xffff;
  double out[NF][8];
  if (live) {
    int loc[8];
    unpack8(pk_loc, loc);
    // ---- geometry: metric at the Gauss points -> shared memory, column t
    {
      int cl[8];
      unpack8(pk_cl, cl);
      if (affine) {
        double j0[3], j1[3], j2[3], G[7];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double o = cs[d * NLMAX + cl[0]];
          j0[d] = cs[d * NLMAX + cl[4]] - o;
          j1[d] = cs[d * NLMAX + cl[2]] - o;
          j2[d] = cs[d * NLMAX + cl[1]] - o;
        }
        metric_q1(j0, j1, j2, 0.125, G);
#pragma unroll
        for (int i = 0; i < 7; ++i) sG[i * CB + t] = G[i];
      } else {
        double D0[3][4], D1[3][4], D2[3][4], dummy[8];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double X[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) X[a] = cs[d * NLMAX + cl[a]];
          forward_q1<false>(X, D0[d], D1[d], D2[d], dummy);
        }
#pragma unroll
        for (int q0 = 0; q0 < 2; ++q0)
#pragma unroll
          for (int q1 = 0; q1 < 2; ++q1)
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
              double j0[3], j1[3], j2[3], G[7];
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                j0[d] = D0[d][q1 * 2 + q2];
                j1[d] = D1[d][q0 * 2 + q2];
                j2[d] = D2[d][q0 * 2 + q1];
              }
              metric_q1(j0, j1, j2, 0.125, G);
              const int q = q0 * 4 + q1 * 2 + q2;
#pragma unroll
              for (int i = 0; i < 7; ++i) sG[(q * 7 + i) * CB + t] = G[i];
            }
      }
    }
    // ---- fields: y_f = K a_f + M b_f
    double xin[NF][8];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
      for (int a = 0; a < 8; ++a) xin[f][a] = xs[f * NLMAX + loc[a]];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      double ak[8], bm[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int h = 0; h < NF; ++h) {
          sa = fma(g.c.cK[f][h], xin[h][a], sa);
          sb = fma(g.c.cM[f][h], xin[h][a], sb);
        }
        ak[a] = sa;
        bm[a] = sb;
      }
      double D0[4], D1[4], D2[4], val[8], dummy[8], E0[4], E1[4], E2[4];
      forward_q1<false>(ak, D0, D1, D2, dummy);
      forward_q1<true>(bm, E0, E1, E2, val);
      double F0[8], F1[8], F2[8], mh[8];
#pragma unroll
      for (int q0 = 0; q0 < 2; ++q0)
#pragma unroll
        for (int q1 = 0; q1 < 2; ++q1)
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            const int q = q0 * 4 + q1 * 2 + q2;
            const double* G = sG + (affine ? 0 : q * 7 * CB) + t;
            const double g0 = D0[q1 * 2 + q2], g1 = D1[q0 * 2 + q2], g2 = D2[q0 * 2 + q1];
            const double G00 = G[0], G01 = G[CB], G02 = G[2 * CB], G11 = G[3 * CB], G12 = G[4 * CB], G22 = G[5 * CB];
            F0[q] = fma(G02, g2, fma(G01, g1, G00 * g0));
            F1[q] = fma(G12, g2, fma(G11, g1, G01 * g0));
            F2[q] = fma(G22, g2, fma(G12, g1, G02 * g0));
            mh[q] = G[6 * CB] * val[q];
          }
      backward_q1(F0, F1, F2, mh, out[f]);
    }
  }
  __syncthreads();   // every thread is done with the metric: its storage now takes the element vectors
  if (live) {
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
      for (int a = 0; a < 8; ++a) sG[(f * 8 + a) * CB + t] = out[f][a];
  }
  __syncthreads();
  // ---- row-owner gather inside the block: local node l adds the entries of its incident (cell, corner) pairs
  const int nl = slp[NLMAX + 1];
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
    const int l = t + k * CB;
    if (l < nl) {
      double acc[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[f] = 0.0;
      const int e1 = slp[l + 1];
      for (int e = slp[l]; e < e1; ++e) {
        const int ca = sla[e];                  // cell_in_block * 8 + corner
        const int idx = (ca & 7) * CB + (ca >> 3);
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[f] += sG[f * 8 * CB + idx];
      }
#pragma unroll
      for (int f = 0; f < NF; ++f) g.ypart[f * g.total_slots + slot0 + l] = acc[f];
    }
  }
}

struct GatherArgs {
  const int64_t* nd_ptr;
  const int32_t* nd_slot;
  const double* ypart;
  long long total_slots, n_nodes;
  const double* x[2];
  double* y[2];
  const uint8_t* out_mask[2];
  int identity_on_masked;
  double* dot_partials;
  const double* skip_flag;
};

template <int NF, int U>
__global__ void __launch_bounds__(VT) k_cells_gather(const GatherArgs g) {
  __shared__ double sm[VT / 32];
  if (g.skip_flag != nullptr && *g.skip_flag != 0.0) return;
  double dot = 0.0;
  for (long long node = (long long)blockIdx.x * VT + threadIdx.x; node < g.n_nodes; node += (long long)gridDim.x * VT) {
    double acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.0;
    const long long e0 = g.nd_ptr[node], e1 = g.nd_ptr[node + 1];
    // a hexahedral node sits in at most 8 blocks: all slot loads first, then the dependent partials, added in
    // ascending slot order; meshes with higher node valence take the tail loop
    long long sl[U];
#pragma unroll
    for (int j = 0; j < U; ++j) sl[j] = e0 + j < e1 ? (long long)g.nd_slot[e0 + j] : -1;
    double pv[NF][U];
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int f = 0; f < NF; ++f) pv[f][j] = sl[j] >= 0 ? g.ypart[f * g.total_slots + sl[j]] : 0.0;
#pragma unroll
    for (int j = 0; j < U; ++j)
      if (sl[j] >= 0)
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[f] += pv[f][j];
    for (long long e = e0 + U; e < e1; ++e) {
      const long long s = g.nd_slot[e];
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[f] += g.ypart[f * g.total_slots + s];
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double xc = g.x[f][node];
      double yv = acc[f];
      if (g.out_mask[f] != nullptr && g.out_mask[f][node]) yv = g.identity_on_masked ? xc : 0.0;
      g.y[f][node] = yv;
      dot = fma(xc, yv, dot);
    }
  }
  if (g.dot_partials != nullptr) {
    const double t = block_sum(dot, sm);
    if (threadIdx.x == 0) g.dot_partials[blockIdx.x] = t;
  }
}

inline uint64_t spread21(uint64_t v) {   // 21 bits -> every third bit
  v &= 0x1fffff;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

template <typename T>
int upload(dpp_context* ctx, T** dst, const std::vector<T>& src) {
  DPP_CHECK(dev_alloc(ctx, dst, (int64_t)src.size()));
  if (!src.empty()) DPP_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
  return DPP_OK;
}

size_t cells_smem(int nf) {
  return sizeof(double) * ((size_t)(nf + 3) * NLMAX + (size_t)8 * 7 * CB) + sizeof(uint16_t) * ((size_t)CB * 8 + LPTR);
}

}  // namespace

bool cells_supported(const dpp_context* ctx) { return ctx->dim == 3 && ctx->degree == 1; }
bool cells_ready(const dpp_context* ctx) { return ctx->cells != nullptr; }

void cells_destroy(dpp_context* ctx) {
  CellBlocks* B = ctx->cells;
  if (!B) return;
  void* p[] = {B->blk_nodes, B->cblk, B->cell_loc, B->same_numbering ? nullptr : B->cell_cloc, B->ladj_ptr, B->ladj,
               B->blk_affine, B->nd_ptr, B->nd_slot, B->xblk, B->ypart};
  for (void* q : p)
    if (q) cudaFree(q);
  delete B;
  ctx->cells = nullptr;
}

// Build the cell blocks from the device copies of the mesh (host integer work, once per mesh).
int cells_setup(dpp_context* ctx) {
  if (ctx->cells) return DPP_OK;
  if (!cells_supported(ctx)) {
    ctx->set_error("cell-block kernel: 3-D hexahedral Q1 meshes only");
    return DPP_ERR_INVALID;
  }
  const int64_t nc = ctx->n_cells, nn = ctx->n_nodes, ncn = ctx->n_coord_nodes;
  std::vector<int32_t> cnm((size_t)nc * 8), ccnm_store;
  std::vector<double> xyz((size_t)ncn * 3);
  DPP_CUDA(cudaMemcpy(cnm.data(), ctx->d_cnm, sizeof(int32_t) * cnm.size(), cudaMemcpyDeviceToHost));
  DPP_CUDA(cudaMemcpy(xyz.data(), ctx->d_coords, sizeof(double) * xyz.size(), cudaMemcpyDeviceToHost));
  const int32_t* ccnm = cnm.data();
  bool same = ctx->ccnm_alias;
  if (!same) {
    ccnm_store.resize((size_t)nc * 8);
    DPP_CUDA(cudaMemcpy(ccnm_store.data(), ctx->d_ccnm, sizeof(int32_t) * ccnm_store.size(), cudaMemcpyDeviceToHost));
    same = nn == ncn && std::memcmp(ccnm_store.data(), cnm.data(), sizeof(int32_t) * cnm.size()) == 0;
    if (!same) ccnm = ccnm_store.data();
  }
  for (int64_t i = 0; i < nc * 8; ++i)
    if (cnm[i] < 0 || cnm[i] >= nn || ccnm[i] < 0 || ccnm[i] >= ncn) {
      ctx->set_error("cell_node_map entry out of range");
      return DPP_ERR_INVALID;
    }
  // ---- Morton order of the cell centroids
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t v = 0; v < ncn; ++v)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], xyz[v * 3 + d]);
      hi[d] = std::max(hi[d], xyz[v * 3 + d]);
    }
  double inv[3];
  for (int d = 0; d < 3; ++d) inv[d] = hi[d] > lo[d] ? 2097151.0 / (hi[d] - lo[d]) : 0.0;
  std::vector<std::pair<uint64_t, int32_t>> key((size_t)nc);
  std::vector<uint8_t> cell_affine((size_t)nc);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double X[8][3], ctr[3] = {0, 0, 0};
    for (int a = 0; a < 8; ++a)
      for (int d = 0; d < 3; ++d) {
        X[a][d] = xyz[(size_t)ccnm[c * 8 + a] * 3 + d];
        ctr[d] += 0.125 * X[a][d];
      }
    uint64_t k = 0;
    for (int d = 0; d < 3; ++d) k |= spread21((uint64_t)((ctr[d] - lo[d]) * inv[d])) << (2 - d);
    key[c] = {k, (int32_t)c};
    // affine test (same criterion as k_cell_geometry): every vertex = v0 + the edge vectors of its set bits
    double e[3][3], scale = 0.0;   // e[axis][d]; local index bit (2 - axis)
    for (int ax = 0; ax < 3; ++ax)
      for (int d = 0; d < 3; ++d) {
        e[ax][d] = X[1 << (2 - ax)][d] - X[0][d];
        scale = std::max(scale, std::fabs(e[ax][d]));
      }
    bool aff = true;
    for (int a = 0; a < 8; ++a)
      for (int d = 0; d < 3; ++d) {
        double v = X[0][d];
        for (int ax = 0; ax < 3; ++ax)
          if (a & (1 << (2 - ax))) v += e[ax][d];
        if (std::fabs(v - X[a][d]) > 1e-12 * scale) aff = false;
      }
    cell_affine[c] = aff ? 1 : 0;
  }
  std::sort(key.begin(), key.end());
  // ---- CB consecutive cells per block; a block whose node list would overflow NLMAX is halved (recursively)
  std::vector<int32_t> blk_cell_ptr{0};
  {
    const int64_t nb0 = (nc + CB - 1) / CB;
    std::vector<std::vector<int32_t>> cuts((size_t)nb0);   // end positions of the pieces of each initial block
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t b = 0; b < nb0; ++b) {
      std::vector<std::pair<int64_t, int64_t>> work{{b * CB, std::min<int64_t>(nc, (b + 1) * CB)}};
      std::vector<int32_t> L;
      std::vector<int32_t>& out = cuts[(size_t)b];
      while (!work.empty()) {
        const auto [c0, c1] = work.back();
        work.pop_back();
        size_t uniq = 0;
        for (int pass = 0; pass < (same ? 1 : 2); ++pass) {
          const int32_t* map = pass ? ccnm : cnm.data();
          L.clear();
          for (int64_t c = c0; c < c1; ++c)
            for (int a = 0; a < 8; ++a) L.push_back(map[(size_t)key[c].second * 8 + a]);
          std::sort(L.begin(), L.end());
          uniq = std::max<size_t>(uniq, std::unique(L.begin(), L.end()) - L.begin());
        }
        if (uniq > (size_t)NLMAX && c1 - c0 > 1) {   // later half first on the stack: pieces come out in order
          const int64_t mid = (c0 + c1) / 2;
          work.push_back({mid, c1});
          work.push_back({c0, mid});
        } else {
          out.push_back((int32_t)c1);
        }
      }
    }
    for (const auto& v : cuts)
      for (int32_t e : v) blk_cell_ptr.push_back(e);
  }
  const int nb = (int)blk_cell_ptr.size() - 1;
  if ((long long)nb * NLMAX >= (1LL << 31)) {
    ctx->set_error("cell-block kernel: slot index exceeds int32");
    return DPP_ERR_INVALID;
  }
  CellBlocks* B = new CellBlocks();
  ctx->cells = B;
  B->same_numbering = same;
  B->n_blocks = nb;
  const long long total = (long long)nb * NLMAX;
  B->total_slots = total;
  // ---- per block: node lists (+ staged coordinates), local connectivity and its transpose
  std::vector<int32_t> blk_nodes((size_t)total, -1);
  std::vector<double> cblk((size_t)nb * 3 * NLMAX, 0.0);
  std::vector<uint16_t> cell_loc((size_t)nb * CB * 8, 0xffff), cell_cloc(same ? 0 : (size_t)nb * CB * 8, 0xffff);
  std::vector<uint16_t> ladj((size_t)nb * CB * 8, 0), ladj_ptr((size_t)nb * LPTR, 0);
  std::vector<uint8_t> blk_affine(nb, 1);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 64)
  for (int b = 0; b < nb; ++b) {
    const int c0 = blk_cell_ptr[b], cn = blk_cell_ptr[b + 1] - c0;
    std::vector<int32_t> L, C;
    L.reserve((size_t)cn * 8);
    for (int i = 0; i < cn; ++i)
      for (int a = 0; a < 8; ++a) L.push_back(cnm[(size_t)key[c0 + i].second * 8 + a]);
    std::sort(L.begin(), L.end());
    L.erase(std::unique(L.begin(), L.end()), L.end());
    if (!same) {
      for (int i = 0; i < cn; ++i)
        for (int a = 0; a < 8; ++a) C.push_back(ccnm[(size_t)key[c0 + i].second * 8 + a]);
      std::sort(C.begin(), C.end());
      C.erase(std::unique(C.begin(), C.end()), C.end());
    }
    const std::vector<int32_t>& CL = same ? L : C;
    if (L.size() > (size_t)NLMAX || CL.size() > (size_t)NLMAX || cn > CB) {
      bad = 1;
      continue;
    }
    std::copy(L.begin(), L.end(), blk_nodes.begin() + (size_t)b * NLMAX);
    for (size_t l = 0; l < CL.size(); ++l)
      for (int d = 0; d < 3; ++d) cblk[((size_t)b * 3 + d) * NLMAX + l] = xyz[(size_t)CL[l] * 3 + d];
    uint16_t* P = ladj_ptr.data() + (size_t)b * LPTR;   // counts first, offsets after the prefix sum
    for (int i = 0; i < cn; ++i) {
      const int32_t cell = key[c0 + i].second;
      for (int a = 0; a < 8; ++a) {
        const int l = (int)(std::lower_bound(L.begin(), L.end(), cnm[(size_t)cell * 8 + a]) - L.begin());
        cell_loc[((size_t)b * CB + i) * 8 + a] = (uint16_t)l;
        P[l + 1]++;
        if (!same)
          cell_cloc[((size_t)b * CB + i) * 8 + a] =
              (uint16_t)(std::lower_bound(C.begin(), C.end(), ccnm[(size_t)cell * 8 + a]) - C.begin());
      }
      if (!cell_affine[cell]) blk_affine[b] = 0;
    }
    for (size_t l = 0; l < L.size(); ++l) P[l + 1] = (uint16_t)(P[l + 1] + P[l]);
    std::vector<uint16_t> cur(P, P + L.size());
    uint16_t* A = ladj.data() + (size_t)b * CB * 8;     // ascending (cell, corner) order per node
    for (int i = 0; i < cn; ++i)
      for (int a = 0; a < 8; ++a) A[cur[cell_loc[((size_t)b * CB + i) * 8 + a]]++] = (uint16_t)(i * 8 + a);
    P[NLMAX + 1] = (uint16_t)L.size();
  }
  if (bad) {
    ctx->set_error("cell-block kernel: node list of a block exceeds the shared-memory staging area");
    return DPP_ERR_INVALID;
  }
  // ---- node -> slots (ascending: blocks ascending, one slot per block)
  std::vector<int64_t> nd_ptr((size_t)nn + 1, 0);
  for (long long s = 0; s < total; ++s)
    if (blk_nodes[s] >= 0) nd_ptr[(size_t)blk_nodes[s] + 1]++;
  for (int64_t i = 0; i < nn; ++i) nd_ptr[i + 1] += nd_ptr[i];
  B->used_slots = nd_ptr[nn];
  std::vector<int32_t> nd_slot((size_t)nd_ptr[nn]);
  {
    std::vector<int64_t> cur(nd_ptr.begin(), nd_ptr.end() - 1);
    for (long long s = 0; s < total; ++s)
      if (blk_nodes[s] >= 0) nd_slot[(size_t)cur[blk_nodes[s]]++] = (int32_t)s;
  }
  for (int64_t c = 0; c < nc; ++c) B->affine_cells += cell_affine[c];
  DPP_CHECK(upload(ctx, &B->blk_nodes, blk_nodes));
  DPP_CHECK(upload(ctx, &B->cblk, cblk));
  DPP_CHECK(upload(ctx, &B->cell_loc, cell_loc));
  if (same) B->cell_cloc = B->cell_loc;
  else DPP_CHECK(upload(ctx, &B->cell_cloc, cell_cloc));
  DPP_CHECK(upload(ctx, &B->ladj_ptr, ladj_ptr));
  DPP_CHECK(upload(ctx, &B->ladj, ladj));
  DPP_CHECK(upload(ctx, &B->blk_affine, blk_affine));
  DPP_CHECK(upload(ctx, &B->nd_ptr, nd_ptr));
  DPP_CHECK(upload(ctx, &B->nd_slot, nd_slot));
  DPP_CHECK(dev_alloc(ctx, &B->xblk, 2 * total));
  DPP_CHECK(dev_alloc(ctx, &B->ypart, 2 * total));
  DPP_CUDA(cudaFuncSetAttribute(k_cells_q1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cells_smem(1)));
  DPP_CUDA(cudaFuncSetAttribute(k_cells_q1<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cells_smem(2)));
  return DPP_OK;
}

// y = A_bc x on every node (single-GPU handles; slab runs use the structured family)
int cells_apply(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks) {
  CellBlocks* B = ctx->cells;
  if (!B) {
    ctx->set_error("cell-block kernel not set up");
    return DPP_ERR_STATE;
  }
  const int sblocks = (int)std::max<long long>(1, std::min<long long>((B->total_slots + VT - 1) / VT, (long long)ctx->sm_count * 16));
  StageArgs st{};
  st.blk_nodes = B->blk_nodes; st.total_slots = B->total_slots; st.xblk = B->xblk; st.skip_flag = a.skip_flag;
  for (int f = 0; f < 2; ++f) {
    st.x[f] = a.x[f];
    st.in_mask[f] = a.input_premasked ? nullptr : a.in_mask[f];
  }
  if (a.nf == 2) k_cells_stage<2><<<sblocks, VT, 0, ctx->stream>>>(st);
  else k_cells_stage<1><<<sblocks, VT, 0, ctx->stream>>>(st);
  CellArgs g{};
  g.cblk = B->cblk; g.cell_loc = B->cell_loc; g.cell_cloc = B->cell_cloc;
  g.ladj_ptr = B->ladj_ptr; g.ladj = B->ladj; g.blk_affine = B->blk_affine;
  g.xblk = B->xblk;
  g.c = a.c;
  g.ypart = B->ypart;
  g.total_slots = B->total_slots;
  g.skip_flag = a.skip_flag;
  if (a.nf == 2) k_cells_q1<2><<<B->n_blocks, CB, cells_smem(2), ctx->stream>>>(g);
  else k_cells_q1<1><<<B->n_blocks, CB, cells_smem(1), ctx->stream>>>(g);
  ctx->launches += 2;
  DPP_CUDA(cudaGetLastError());
  GatherArgs h{};
  h.nd_ptr = B->nd_ptr; h.nd_slot = B->nd_slot; h.ypart = B->ypart;
  h.total_slots = B->total_slots; h.n_nodes = ctx->n_nodes;
  for (int f = 0; f < 2; ++f) {
    h.x[f] = a.x[f]; h.y[f] = a.y[f]; h.out_mask[f] = a.out_mask[f];
  }
  h.identity_on_masked = a.identity_on_masked;
  h.dot_partials = a.dot_partials;
  h.skip_flag = a.skip_flag;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((ctx->n_nodes + VT - 1) / VT, (long long)ctx->sm_count * 8));
  if (a.nf == 2) k_cells_gather<2, 8><<<blocks, VT, 0, ctx->stream>>>(h);
  else k_cells_gather<1, 8><<<blocks, VT, 0, ctx->stream>>>(h);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (n_partial_blocks) *n_partial_blocks = blocks;
  return DPP_OK;
}

int cells_stats(const dpp_context* ctx, int64_t* n_blocks, int64_t* total_slots, int64_t* affine_cells) {
  const CellBlocks* B = ctx->cells;
  if (!B) return DPP_ERR_STATE;
  if (n_blocks) *n_blocks = B->n_blocks;
  if (total_slots) *total_slots = B->used_slots;
  if (affine_cells) *affine_cells = B->affine_cells;
  return DPP_OK;
}

}  // namespace dpp
