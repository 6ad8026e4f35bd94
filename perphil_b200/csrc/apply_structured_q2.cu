// Matrix-free DPP operator for degree-2 (Q2) spaces on rectilinear tensor grids numbered
// lexicographically (BASELINE.json configs[3]: 3-D hex Q2 192^3, block Picard).  Same factorisation as
// apply_structured.cu,
//     K = Kx(x)My(x)Mz + Mx(x)Ky(x)Mz + Mx(x)My(x)Kz,     M = Mx(x)My(x)Mz,
// with assembled 1-D matrices of half bandwidth 2 in band storage [n][5] (vertex rows have 5 entries,
// mid-node rows 3, zeros elsewhere and outside the domain), so vertex/mid/boundary node types need no
// special cases: the row of the table IS the stencil.  Planes x = const stream through a 3-slot
// shared-memory ring (cp.async, zero-fill outside the domain, halo 2); per plane each thread forms the
// in-plane parts of its node
//     c = (My(x)Mz) x_i,   d = (Ky(x)Mz + My(x)Kz) x_i
// separably (5 rows x 5 columns: 25 shared loads, 65 FMA per field) and keeps the last five planes of
// (c, d) in a register queue for the x-direction sweep  K x = Kx c + Mx d,  M x = Mx c.
// Input vectors must be zero on eliminated columns (true for every Krylov vector; otherwise a pre-mask
// pass); eliminated rows are rewritten by the list-driven fix-up kernel of apply_structured_uniform.cu.
//
// Algorithmic HBM traffic: read x + write y (+1 B Dirichlet) = 34 B/node for the two-field operator.
// The general (unstructured) kernel this replaces on such meshes ran at 0.5 GDoF/s.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "dpp_internal.cuh"

namespace dpp {

namespace {

constexpr int TK = 32;
constexpr int TJ = 8;
constexpr int NT = TK * TJ;
constexpr int H = 2;                    // halo width = half bandwidth of the 1-D matrices
constexpr int SROW = TK + 2 * H;        // 36
constexpr int SLOT = (TJ + 2 * H) * SROW;  // 432 doubles per field per ring slot
constexpr int RING = 3;

struct Q2Args {
  int n[3];
  const double* m1d[3];
  const double* k1d[3];
  const double* x[2];
  double* y[2];
  Coef c;
  double* dot_partials;
  int i_begin, i_end;
  int ntj, ntk, nseg;
  const double* skip_flag;
};

__device__ __forceinline__ int bstart(int t, int n, int nt) { return (int)(((long long)t * n) / nt); }

__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gptr, bool valid) {
  asm volatile(
      "{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n cp.async.ca.shared.global [%0], [%1], 8, p;\n}\n" ::"r"(smem_addr),
      "l"(gptr), "r"((unsigned)valid)
      : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int NF>
__global__ void __launch_bounds__(NT, 2) k_apply_q2(const Q2Args s) {
  if (s.skip_flag != nullptr && *s.skip_flag != 0.0) return;
  __shared__ __align__(16) double xs[RING][NF][SLOT];
  __shared__ double red[NT / 32];

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2];
  const int ntiles = s.ntj * s.ntk;
  const int tile = blockIdx.x % ntiles, seg = blockIdx.x / ntiles;
  const int tkid = tile % s.ntk, tjid = tile / s.ntk;
  const int k0 = bstart(tkid, nk, s.ntk), k1 = bstart(tkid + 1, nk, s.ntk);
  const int j0 = bstart(tjid, nj, s.ntj), j1 = bstart(tjid + 1, nj, s.ntj);
  const int nown = s.i_end - s.i_begin;
  const int i_lo = s.i_begin + bstart(seg, nown, s.nseg);
  const int i_hi = s.i_begin + bstart(seg + 1, nown, s.nseg);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TK + tx;
  const int j = j0 + ty, k = k0 + tx;
  const bool act = (j < j1) && (k < k1);
  const long long plane = (long long)nj * nk;

  // copy duties (fixed across planes): slot elements tid and tid + NT
  long long coff[2];
  bool cok[2];
  unsigned cs[2];
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(&xs[0][0][0]);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int e = tid + q * NT;
    const int r = e / SROW, c = e - r * SROW;
    const int jj = j0 - H + r, kk = k0 - H + c;
    cok[q] = (e < SLOT) && (jj >= 0) && (jj < nj) && (kk >= 0) && (kk < nk);
    coff[q] = cok[q] ? (long long)jj * nk + kk : 0;
    cs[q] = smem_base + (unsigned)(e * 8);
  }

  // in-plane 1-D rows of this thread's node (zero for inactive threads)
  double my[5], ky[5], mz[5], kz[5];
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    my[d] = act ? __ldg(&s.m1d[1][j * 5 + d]) : 0.0;
    ky[d] = act ? __ldg(&s.k1d[1][j * 5 + d]) : 0.0;
    mz[d] = act ? __ldg(&s.m1d[2][k * 5 + d]) : 0.0;
    kz[d] = act ? __ldg(&s.k1d[2][k * 5 + d]) : 0.0;
  }

  double qc[NF][5], qd[NF][5], cen[NF][3];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int d = 0; d < 5; ++d) qc[f][d] = qd[f][d] = 0.0;
    cen[f][0] = cen[f][1] = cen[f][2] = 0.0;
  }
  double dot = 0.0;
  const int i_first = i_lo - H;
  const long long own = (long long)j * nk + k;

  auto issue = [&](int pl, int slot) {
    const bool in = (unsigned)pl < (unsigned)ni;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double* base = s.x[f] + (long long)pl * plane;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (q == 0 || tid + NT < SLOT) cp_async8(cs[q] + (unsigned)((slot * NF + f) * SLOT * 8), base + coff[q], in && cok[q]);
    }
    cp_async_commit();
  };

  issue(i_first, 0);
  issue(i_first + 1, 1);
  int slot = 0;     // ring slot of plane ip
  int ip = i_first;

  // one plane step; queue slot V receives plane ip, the output plane io = ip - 2 reads band entry dd from
  // queue slot (V + 1 + dd) % 5
#define DPP_Q2_STEP(V)                                                                               \
  {                                                                                                  \
    cp_async_wait<1>();                                                                              \
    __syncthreads();                                                                                 \
    {                                                                                                \
      int nslot = slot + 2;                                                                          \
      if (nslot >= RING) nslot -= RING;                                                              \
      issue(ip + 2, nslot);                                                                          \
    }                                                                                                \
    const bool in = (unsigned)ip < (unsigned)ni;                                                     \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                 \
      double c = 0.0, d = 0.0, xc = 0.0;                                                             \
      if (in && act) {                                                                               \
        const double* t = &xs[slot][f][ty * SROW + tx];                                              \
        _Pragma("unroll") for (int dj = 0; dj < 5; ++dj) {                                           \
          const double v0 = t[dj * SROW], v1 = t[dj * SROW + 1], v2 = t[dj * SROW + 2],              \
                       v3 = t[dj * SROW + 3], v4 = t[dj * SROW + 4];                                 \
          if (dj == 2) xc = v2;                                                                      \
          const double tz = fma(mz[0], v0, fma(mz[1], v1, fma(mz[2], v2, fma(mz[3], v3, mz[4] * v4)))); \
          const double uz = fma(kz[0], v0, fma(kz[1], v1, fma(kz[2], v2, fma(kz[3], v3, kz[4] * v4)))); \
          c = fma(my[dj], tz, c);                                                                    \
          d = fma(ky[dj], tz, fma(my[dj], uz, d));                                                   \
        }                                                                                            \
      }                                                                                              \
      qc[f][V] = c;                                                                                  \
      qd[f][V] = d;                                                                                  \
      cen[f][0] = cen[f][1];                                                                         \
      cen[f][1] = cen[f][2];                                                                         \
      cen[f][2] = xc;                                                                                \
    }                                                                                                \
    const int io = ip - H;                                                                           \
    if (act && io >= i_lo && io < i_hi) {                                                            \
      double Kx[NF], Mx[NF];                                                                         \
      _Pragma("unroll") for (int f = 0; f < NF; ++f) Kx[f] = Mx[f] = 0.0;                            \
      _Pragma("unroll") for (int dd = 0; dd < 5; ++dd) {                                             \
        const double mx = __ldg(&s.m1d[0][io * 5 + dd]), kx = __ldg(&s.k1d[0][io * 5 + dd]);         \
        _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                             \
          Mx[f] = fma(mx, qc[f][((V) + 1 + dd) % 5], Mx[f]);                                         \
          Kx[f] = fma(kx, qc[f][((V) + 1 + dd) % 5], fma(mx, qd[f][((V) + 1 + dd) % 5], Kx[f]));     \
        }                                                                                            \
      }                                                                                              \
      const long long node = (long long)io * plane + own;                                            \
      _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                               \
        double yv = 0.0;                                                                             \
        _Pragma("unroll") for (int g = 0; g < NF; ++g) {                                             \
          yv = fma(s.c.cK[f][g], Kx[g], yv);                                                         \
          yv = fma(s.c.cM[f][g], Mx[g], yv);                                                         \
        }                                                                                            \
        s.y[f][node] = yv;                                                                           \
        dot = fma(cen[f][0], yv, dot);                                                               \
      }                                                                                              \
    }                                                                                                \
    if (++slot == RING) slot = 0;                                                                    \
  }

  // planes i_first .. i_hi + 1 ; outputs i_lo .. i_hi - 1
  const int i_last = i_hi + H - 1;
  while (true) {
    DPP_Q2_STEP(0)
    if (++ip > i_last) break;
    DPP_Q2_STEP(1)
    if (++ip > i_last) break;
    DPP_Q2_STEP(2)
    if (++ip > i_last) break;
    DPP_Q2_STEP(3)
    if (++ip > i_last) break;
    DPP_Q2_STEP(4)
    if (++ip > i_last) break;
  }
#undef DPP_Q2_STEP
  cp_async_wait<0>();

  if (s.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (tx == 0) red[ty] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) t += red[w];
      s.dot_partials[blockIdx.x] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Uniform-grid specialisation (BASELINE configs[3] is a uniform 192^3 cube).  On an equally spaced axis the
// assembled Q2 rows take two shapes only -- a vertex row (offsets -2..2: m = h/30 {-1, 2, 8, 2, -1},
// k = 1/(3h) {1, -8, 14, -8, 1}; centre halved on the domain boundary) and a mid-node row (offsets -1..1:
// m = h/30 {2, 16, 2}, k = 1/(3h) {-8, 16, -8}) -- so the coefficients are kernel constants instead of 20 table
// registers per thread, zero entries are never multiplied, and the symmetric rows need one product per PAIR of
// neighbours.  Each thread owns a (vertex, mid) PAIR of nodes along z and reads its 6-column window with three
// 16-byte shared-memory loads per row (the table kernel: 25 8-byte loads per node); rows of one parity share a
// warp, planes of one parity the whole CTA, so no branch diverges.
// Per node and field: ~36 fp64 operations in-plane + ~10 along x (table kernel: 65 + 15) and 0.75 shared-memory
// wavefronts (1.56).  Same plane streaming (cp.async ring, zero-filled halo) and register queue along x.
// ---------------------------------------------------------------------------------------------------------------
// The assembled 1-D rows on an equally spaced axis are (h/30) x {-1, 2, 8|4, 2, -1} / {2, 16, 2} (mass: vertex / mid
// row; 4 = one cell only, on the domain boundary) and 1/(3h) x {1, -8, 14|7, -8, 1} / {-8, 16, -8} (stiffness).  The
// kernel evaluates the INTEGER stencils -- immediate operands of the fp64 instructions: no constant loads, no uniform
// registers (the 30 coefficients of the first version spilled them: 7 % of the issued instructions were UR moves) --
// and the scale factors ride on three coefficients per field pair:
//   y_f = sum_g a1[f][g] (Kx' c') + a2[f][g] (Mx' d') + a3[f][g] (Mx' c'),
//   c' = (My' x Mz') x,  d' = rho_y (Ky' x Mz') x + rho_z (My' x Kz') x,  rho = (1/(3h)) / (h/30) = 10 / h^2.
struct Q2UArgs {
  int n[3];
  double rho_y, rho_z;
  double a1[2][2], a2[2][2], a3[2][2];
  double cxM, cxK;           // centre entries of the boundary vertex rows along x: 4, 7 (dummy axis of a 2-D mesh: 1, 0)
  const double* x[2];
  double* y[2];
  double* dot_partials;
  int i_begin, i_end;
  int ntj, ntk, nseg;
  int dom_lo, dom_hi;
  const double* skip_flag;
};

constexpr int UPT = 16;                 // pair-threads per tile row: 32 columns
constexpr int UNT = UPT * TJ;           // 128 threads

template <int NF>
__global__ void __launch_bounds__(UNT, 3) k_apply_q2u(const Q2UArgs s) {
  if (s.skip_flag != nullptr && *s.skip_flag != 0.0) return;
  __shared__ __align__(16) double xs[RING][NF][SLOT];
  __shared__ double red[UNT / 32];

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2];
  const int ntiles = s.ntj * s.ntk;
  const int tile = blockIdx.x % ntiles, seg = blockIdx.x / ntiles;
  const int tkid = tile % s.ntk, tjid = tile / s.ntk;
  const int k0 = tkid * TK, j0 = tjid * TJ;            // even: node parity = local parity
  const int nown = s.i_end - s.i_begin;
  const int i_lo = s.i_begin + bstart(seg, nown, s.nseg);
  const int i_hi = s.i_begin + bstart(seg + 1, nown, s.nseg);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int jr = warp + 4 * (lane >> 4);               // rows {w, w + 4} of a warp have one parity
  const int kp = 2 * (lane & 15);                      // even column of the pair
  const int j = j0 + jr, k = k0 + kp;
  const bool rowV = (jr & 1) == 0;
  const bool actV = (j < nj) && (k < nk), actM = (j < nj) && (k + 1 < nk);
  const long long plane = (long long)nj * nk;

  // copy duties (fixed across planes): slot elements tid + q * UNT
  constexpr int NCOPY = (SLOT + UNT - 1) / UNT;
  long long coff[NCOPY];
  bool cok[NCOPY];
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(&xs[0][0][0]);
#pragma unroll
  for (int q = 0; q < NCOPY; ++q) {
    const int e = tid + q * UNT;
    const int r = e / SROW, c = e - r * SROW;
    const int jj = j0 - H + r, kk = k0 - H + c;
    cok[q] = (e < SLOT) && (jj >= 0) && (jj < nj) && (kk >= 0) && (kk < nk);
    coff[q] = cok[q] ? (long long)jj * nk + kk : 0;
  }
  // per-thread centre coefficients (boundary rows / columns have one cell instead of two)
  const bool jb = (j == 0 || j == nj - 1), kb = (k == 0 || k == nk - 1);
  const double myVc = jb ? 4.0 : 8.0, kyVc = jb ? 7.0 : 14.0, mzVc = kb ? 4.0 : 8.0, kzVc = kb ? 7.0 : 14.0;

  double qcV[NF][5], qdV[NF][5], qcM[NF][5], qdM[NF][5], cenV[NF][3], cenM[NF][3];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int d = 0; d < 5; ++d) qcV[f][d] = qdV[f][d] = qcM[f][d] = qdM[f][d] = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) cenV[f][d] = cenM[f][d] = 0.0;
  }
  double dot = 0.0;
  const int i_first = i_lo - H;
  const long long own = (long long)j * nk + k;

  auto issue = [&](int pl, int slot) {
    const bool in = (unsigned)pl < (unsigned)ni;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double* base = s.x[f] + (long long)pl * plane;
#pragma unroll
      for (int q = 0; q < NCOPY; ++q)
        if (tid + q * UNT < SLOT)
          cp_async8(smem_base + (unsigned)(((slot * NF + f) * SLOT + tid + q * UNT) * 8), base + coff[q], in && cok[q]);
    }
    cp_async_commit();
  };

  // one row of the in-plane window: both nodes of the pair
#define DPP_Q2U_ROW(T, MYR, KYR)                                                                      \
  {                                                                                                   \
    const double2 p0 = *reinterpret_cast<const double2*>(T);                                          \
    const double2 p1 = *reinterpret_cast<const double2*>((T) + 2);                                    \
    const double2 p2 = *reinterpret_cast<const double2*>((T) + 4);                                    \
    const double s2 = p0.x + p2.x, s1 = p0.y + p1.y, sm = p1.x + p2.x, e16 = 16.0 * p1.y;             \
    const double tzV = fma(2.0, s1, fma(mzVc, p1.x, -s2));                                            \
    const double uzV = fma(-8.0, s1, fma(kzVc, p1.x, s2));                                            \
    const double tzM = fma(2.0, sm, e16);                                                             \
    const double uzM = fma(-8.0, sm, e16);                                                            \
    cV = fma(MYR, tzV, cV);                                                                           \
    d1V = fma(KYR, tzV, d1V);                                                                         \
    d2V = fma(MYR, uzV, d2V);                                                                         \
    cM = fma(MYR, tzM, cM);                                                                           \
    d1M = fma(KYR, tzM, d1M);                                                                         \
    d2M = fma(MYR, uzM, d2M);                                                                         \
  }

  issue(i_first, 0);
  issue(i_first + 1, 1);
  int slot = 0;     // ring slot of plane ip
  int ip = i_first;

  // one plane step: the queues shift by one plane (entry 4 = plane ip just computed, entry 2 = the output plane
  // io = ip - 2).  The shift costs 16 register moves per field and step but keeps ONE copy of the step in the
  // instruction stream: the 5-fold unrolled rotation of the table kernel did not fit the instruction cache here
  // (ncu: 24 % of the stall samples were instruction fetches).
  const int i_last = i_hi + H - 1;
#pragma unroll 1
  for (; ip <= i_last; ++ip) {
    cp_async_wait<1>();
    __syncthreads();
    {
      int nslot = slot + 2;
      if (nslot >= RING) nslot -= RING;
      issue(ip + 2, nslot);
    }
    const bool in = (unsigned)ip < (unsigned)ni;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      double cV = 0.0, dV = 0.0, cM = 0.0, dM = 0.0, xcV = 0.0, xcM = 0.0;
      if (in && actV) {
        double d1V = 0.0, d2V = 0.0, d1M = 0.0, d2M = 0.0;
        const double* t = &xs[slot][f][jr * SROW + kp];
        if (rowV) {
          DPP_Q2U_ROW(t, -1.0, 1.0)
          DPP_Q2U_ROW(t + SROW, 2.0, -8.0)
          DPP_Q2U_ROW(t + 2 * SROW, myVc, kyVc)
          DPP_Q2U_ROW(t + 3 * SROW, 2.0, -8.0)
          DPP_Q2U_ROW(t + 4 * SROW, -1.0, 1.0)
        } else {
          DPP_Q2U_ROW(t + SROW, 2.0, -8.0)
          DPP_Q2U_ROW(t + 2 * SROW, 16.0, 16.0)
          DPP_Q2U_ROW(t + 3 * SROW, 2.0, -8.0)
        }
        dV = fma(s.rho_y, d1V, s.rho_z * d2V);
        dM = fma(s.rho_y, d1M, s.rho_z * d2M);
        xcV = t[2 * SROW + 2];
        xcM = t[2 * SROW + 3];
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        qcV[f][d] = qcV[f][d + 1]; qdV[f][d] = qdV[f][d + 1];
        qcM[f][d] = qcM[f][d + 1]; qdM[f][d] = qdM[f][d + 1];
      }
      qcV[f][4] = cV; qdV[f][4] = dV; qcM[f][4] = cM; qdM[f][4] = dM;
      cenV[f][0] = cenV[f][1]; cenV[f][1] = cenV[f][2]; cenV[f][2] = xcV;
      cenM[f][0] = cenM[f][1]; cenM[f][1] = cenM[f][2]; cenM[f][2] = xcM;
    }
    const int io = ip - H;
    if (actV && io >= i_lo && io < i_hi) {
      // Kc = Kx' c', Md = Mx' d', Mc = Mx' c' (integer rows along x)
      double KcV[NF], MdV[NF], McV[NF], KcM[NF], MdM[NF], McM[NF];
      if ((io & 1) == 0) {
        const bool xb = (io == 0 && s.dom_lo) || (io == ni - 1 && s.dom_hi);
        const double mc = xb ? s.cxM : 8.0, kc = xb ? s.cxK : 14.0;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const double c2 = qcV[f][0] + qcV[f][4], c1 = qcV[f][1] + qcV[f][3], c0 = qcV[f][2];
          const double d2 = qdV[f][0] + qdV[f][4], d1 = qdV[f][1] + qdV[f][3], d0 = qdV[f][2];
          McV[f] = fma(2.0, c1, fma(mc, c0, -c2));
          KcV[f] = fma(-8.0, c1, fma(kc, c0, c2));
          MdV[f] = fma(2.0, d1, fma(mc, d0, -d2));
          const double e2 = qcM[f][0] + qcM[f][4], e1 = qcM[f][1] + qcM[f][3], e0 = qcM[f][2];
          const double g2 = qdM[f][0] + qdM[f][4], g1 = qdM[f][1] + qdM[f][3], g0 = qdM[f][2];
          McM[f] = fma(2.0, e1, fma(mc, e0, -e2));
          KcM[f] = fma(-8.0, e1, fma(kc, e0, e2));
          MdM[f] = fma(2.0, g1, fma(mc, g0, -g2));
        }
      } else {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const double c1 = qcV[f][1] + qcV[f][3], c16 = 16.0 * qcV[f][2], d1 = qdV[f][1] + qdV[f][3];
          McV[f] = fma(2.0, c1, c16);
          KcV[f] = fma(-8.0, c1, c16);
          MdV[f] = fma(2.0, d1, 16.0 * qdV[f][2]);
          const double e1 = qcM[f][1] + qcM[f][3], e16 = 16.0 * qcM[f][2], g1 = qdM[f][1] + qdM[f][3];
          McM[f] = fma(2.0, e1, e16);
          KcM[f] = fma(-8.0, e1, e16);
          MdM[f] = fma(2.0, g1, 16.0 * qdM[f][2]);
        }
      }
      const long long node = (long long)io * plane + own;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        double yV = 0.0, yM = 0.0;
#pragma unroll
        for (int g = 0; g < NF; ++g) {
          yV = fma(s.a1[f][g], KcV[g], fma(s.a2[f][g], MdV[g], fma(s.a3[f][g], McV[g], yV)));
          yM = fma(s.a1[f][g], KcM[g], fma(s.a2[f][g], MdM[g], fma(s.a3[f][g], McM[g], yM)));
        }
        s.y[f][node] = yV;
        dot = fma(cenV[f][0], yV, dot);
        if (actM) {
          s.y[f][node + 1] = yM;
          dot = fma(cenM[f][0], yM, dot);
        }
      }
    }
    if (++slot == RING) slot = 0;
  }
#undef DPP_Q2U_ROW
  cp_async_wait<0>();

  if (s.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < UNT / 32; ++w) t += red[w];
      s.dot_partials[blockIdx.x] = t;
    }
  }
}

__global__ void k_premask_q2(long long n, const double* __restrict__ x, const uint8_t* __restrict__ m,
                             double* __restrict__ xm) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    xm[i] = m[i] ? 0.0 : x[i];
}

}  // namespace

int structured_apply_q2(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks) {
  const GridDesc& g = ctx->grid;
  const long long plane = (long long)g.n[1] * g.n[2];
  if (a.owned_begin % plane || a.owned_end % plane) {
    ctx->set_error("structured Q2 apply: owned range must consist of whole x-planes");
    return DPP_ERR_INVALID;
  }
  Q2Args s{};
  for (int d = 0; d < 3; ++d) { s.n[d] = g.n[d]; s.m1d[d] = g.m1d[d]; s.k1d[d] = g.k1d[d]; }
  int fld[2] = {0, 0};
  double* ys[2] = {nullptr, nullptr};
  const double* xid[2] = {nullptr, nullptr};
  bool need_fix = false;
  for (int f = 0; f < a.nf; ++f) {
    s.x[f] = a.x[f];
    s.y[f] = a.y[f];
    if (a.in_mask[f] != nullptr && !a.input_premasked) {
      if (!ctx->d_premask) DPP_CHECK(dev_alloc(ctx, &ctx->d_premask, 2 * ctx->n_nodes));
      double* xm = ctx->d_premask + (size_t)f * ctx->n_nodes;
      const int blocks = (int)std::min<long long>((ctx->n_nodes + 255) / 256, (long long)ctx->sm_count * 16);
      k_premask_q2<<<blocks, 256, 0, ctx->stream>>>(ctx->n_nodes, a.x[f], a.in_mask[f], xm);
      ctx->launches++;
      s.x[f] = xm;
    }
    if (a.out_mask[f] != nullptr) {
      const long long fl = (a.out_mask[f] - ctx->d_mask) / ctx->n_nodes;
      if (fl < 0 || fl > 1 || a.out_mask[f] != ctx->d_mask + fl * ctx->n_nodes) {
        ctx->set_error("structured Q2 apply: out_mask must be a field of the handle's Dirichlet mask");
        return DPP_ERR_INVALID;
      }
      fld[f] = (int)fl;
      need_fix = true;
    } else {
      fld[f] = -1;
    }
    ys[f] = a.y[f];
    xid[f] = a.x[f];
  }
  s.c = a.c;
  s.dot_partials = a.dot_partials;
  s.i_begin = (int)(a.owned_begin / plane);
  s.i_end = (int)(a.owned_end / plane);
  s.skip_flag = a.skip_flag;
  s.ntk = (g.n[2] + TK - 1) / TK;
  s.ntj = (g.n[1] + TJ - 1) / TJ;
  const int tiles = s.ntk * s.ntj;
  const int nown = s.i_end - s.i_begin;
  if (nown <= 0) {
    if (n_partial_blocks) *n_partial_blocks = 0;
    return DPP_OK;
  }
  int nseg = 1;
  if (tiles <= kMaxPartialBlocks / 2) {
    nseg = choose_x_segments(tiles, nown, ctx->sm_count * 2, kMaxPartialBlocks, 2 * H);
  } else if (a.dot_partials != nullptr && tiles > kMaxPartialBlocks * kMaxDotWidth) {
    ctx->set_error("structured Q2 apply: too many tiles for the reduction scratch");
    return DPP_ERR_INVALID;
  }
  // equally spaced axes: the specialised kernel (DPP_Q2_TABLE_KERNEL=1 forces the table-driven one)
  bool uniform = getenv("DPP_Q2_TABLE_KERNEL") == nullptr;
  double hh[3] = {1.0, 1.0, 1.0};
  for (int d = 0; d < 3 && uniform; ++d) {
    const std::vector<double>& v = ctx->h_axis[d];
    if (v.size() < 2) continue;   // dummy axis
    hh[d] = v[1] - v[0];
    for (size_t t = 2; t < v.size(); ++t)
      if (std::fabs((v[t] - v[t - 1]) - hh[d]) > 1e-12 * std::fabs(hh[d])) uniform = false;
  }
  if (uniform) {
    Q2UArgs u{};
    for (int d = 0; d < 3; ++d) u.n[d] = g.n[d];
    {
      // scale factors of the 1-D rows: mass h/30, stiffness 1/(3h); dummy axis of a 2-D mesh: M = [1], K = [0]
      const bool dummy = g.n[0] == 1;
      const double bmx = dummy ? 1.0 : hh[0] / 30.0, bkx = dummy ? 0.0 : 1.0 / (3.0 * hh[0]);
      const double B = (hh[1] / 30.0) * (hh[2] / 30.0);
      u.rho_y = 10.0 / (hh[1] * hh[1]);
      u.rho_z = 10.0 / (hh[2] * hh[2]);
      u.cxM = dummy ? 1.0 : 4.0;
      u.cxK = dummy ? 0.0 : 7.0;
      for (int f = 0; f < 2; ++f)
        for (int q = 0; q < 2; ++q) {
          u.a1[f][q] = a.c.cK[f][q] * bkx * B;
          u.a2[f][q] = a.c.cK[f][q] * bmx * B;
          u.a3[f][q] = a.c.cM[f][q] * bmx * B;
        }
    }
    for (int f = 0; f < a.nf; ++f) { u.x[f] = s.x[f]; u.y[f] = s.y[f]; }
    u.dot_partials = a.dot_partials;
    u.i_begin = s.i_begin; u.i_end = s.i_end;
    u.ntj = s.ntj; u.ntk = s.ntk;
    u.dom_lo = ctx->dom_lo; u.dom_hi = ctx->dom_hi;
    u.skip_flag = a.skip_flag;
    if (tiles <= kMaxPartialBlocks / 2) nseg = choose_x_segments(tiles, nown, ctx->sm_count * 3, kMaxPartialBlocks, 2 * H);
    u.nseg = nseg;
    if (a.nf == 2) k_apply_q2u<2><<<tiles * nseg, UNT, 0, ctx->stream>>>(u);
    else k_apply_q2u<1><<<tiles * nseg, UNT, 0, ctx->stream>>>(u);
  } else {
    s.nseg = nseg;
    dim3 grid(tiles * nseg), block(TK, TJ);
    if (a.nf == 2)
      k_apply_q2<2><<<grid, block, 0, ctx->stream>>>(s);
    else
      k_apply_q2<1><<<grid, block, 0, ctx->stream>>>(s);
  }
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (need_fix) DPP_CHECK(structured_fix_rows(ctx, a.nf, fld, ys, xid, a.identity_on_masked, a.skip_flag));
  if (n_partial_blocks) *n_partial_blocks = tiles * nseg;
  return DPP_OK;
}

}  // namespace dpp
