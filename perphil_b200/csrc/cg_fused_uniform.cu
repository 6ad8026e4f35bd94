// One Jacobi/unpreconditioned CG iteration in two kernels on uniform tensor grids (the headline
// path of BASELINE.json: 3-D hex Q1, matrix-free, Jacobi-CG).  PETSc's KSPCG loop (solver.py:71)
//     p = z + beta p ; w = A p ; <p,w> ; x += alpha p ; r -= alpha w ; z = D^-1 r ; <r,z>, <z,z>
// needs two global reductions per iteration, so two kernels are the minimum without a grid-wide
// barrier.  HBM passes (one pass = one field-blocked vector, 8 B/dof):
//   k_cg_fused_apply  reads r, p_old, x (3)   writes p, w = A p, x (3)      [x update of the previous
//                     iteration is deferred into it: p_old is on chip anyway]
//   k_cg_r_update     reads r, w (2)          writes r (1)                  [z = D^-1 r only in registers]
// = 9 passes per iteration instead of 13 for the unfused sequence (apply 2 + xr-update 7 + p-update 4).
// The reciprocal diagonal is never read from memory: on a uniform grid diag(A) takes one of 8 values
// per field (node on the domain boundary of axis x/y/z or not), kept in a 16-entry table.
//
// The apply part is the kernel of apply_structured_uniform.cu (cp.async ring of planes, two nodes
// per thread, register queue along x); what changes is that the stencil input p is formed on chip:
// each thread combines exactly the elements it copied itself (its two nodes + at most one halo
// node), p = dinv*r + beta*p_old, in place in the shared ring before the per-plane barrier.
#include <algorithm>

#include "vector_ops.cuh"

namespace dpp {

namespace {

constexpr int TK = 32;
constexpr int TY = 8;
constexpr int TJ = 2 * TY;
constexpr int NT = TK * TY;
constexpr int SROW = TK + 2;
constexpr int SLOT = (TJ + 2) * SROW;
constexpr int XSLOT = TJ * TK;
constexpr int HALO = 2 * SROW + 2 * TJ;
constexpr int RING = 3;

struct FArgs {
  int n[3];
  const double* m1d[3];
  const double* k1d[3];
  double mo[3], ko[3];
  double mxc_i, mxc_b, kxc_i, kxc_b;
  const double* r[2];
  const double* pin[2];   // p of the previous iteration (ghost planes valid)
  double* pout[2];        // p of this iteration (own + ghost planes written)
  double* x[2];
  double* w[2];
  Coef c;
  double* dot_partials;
  int i_begin, i_end;
  int ntj, ntk, nseg;
  const double* S;        // scalar slot of the solver
  const double* dtab;     // [2][8] reciprocal diagonal per boundary class bx*4 + by*2 + bz
  int dom_lo, dom_hi;     // local plane 0 / n[0]-1 lies on the domain boundary (else it is a ghost plane)
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
struct Smem {
  double r[RING][NF][SLOT];
  double p[RING][NF][SLOT];
  double x[RING][NF][XSLOT];
  double red[TY];
  double dtab[16];
};

template <int NF>
__global__ void __launch_bounds__(NT, 2) k_cg_fused_apply(const FArgs s) {
  if (s.S[S_REASON] != 0.0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<NF>& sm = *reinterpret_cast<Smem<NF>*>(smem_raw);

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2];
  const int tile = blockIdx.x;
  const int tkid = tile % s.ntk, tjid = tile / s.ntk;
  const int k0 = bstart(tkid, nk, s.ntk), k1 = bstart(tkid + 1, nk, s.ntk);
  const int j0 = bstart(tjid, nj, s.ntj), j1 = bstart(tjid + 1, nj, s.ntj);
  const int nown = s.i_end - s.i_begin;
  const int i_lo = s.i_begin + bstart(blockIdx.y, nown, s.nseg);
  const int i_hi = s.i_begin + bstart(blockIdx.y + 1, nown, s.nseg);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TK + tx;
  const int jA = j0 + 2 * ty, jB = jA + 1, k = k0 + tx;
  const bool actA = (jA < j1) && (k < k1), actB = (jB < j1) && (k < k1);
  const long long plane = (long long)nj * nk;
  const int i_first = i_lo - 1;
  // planes whose p this CTA stores: its output planes, plus a ghost plane next to the first/last segment
  const int pw_lo = (blockIdx.y == 0 && s.i_begin > 0) ? i_lo - 1 : i_lo;
  const int pw_hi = (blockIdx.y == gridDim.y - 1 && s.i_end < ni) ? i_hi + 1 : i_hi;

  if (tid < 16) sm.dtab[tid] = (tid < 8 * NF) ? s.dtab[tid] : 1.0;

  const bool ownA_ok = (jA < nj) && (k < nk), ownB_ok = (jB < nj) && (k < nk);
  const long long own_off = (long long)jA * nk + k;
  int hr = 0, hc = 0;
  bool halo_ok = false;
  const bool is_halo = tid < HALO;
  if (is_halo) {
    if (tid < SROW) { hr = 0; hc = tid; }
    else if (tid < 2 * SROW) { hr = TJ + 1; hc = tid - SROW; }
    else if (tid < 2 * SROW + TJ) { hr = tid - 2 * SROW + 1; hc = 0; }
    else { hr = tid - 2 * SROW - TJ + 1; hc = TK + 1; }
    const int jj = j0 - 1 + hr, kk = k0 - 1 + hc;
    halo_ok = (jj >= 0) && (jj < nj) && (kk >= 0) && (kk < nk);
  }
  const int jH = j0 - 1 + hr, kH = k0 - 1 + hc;
  const long long halo_delta = halo_ok ? ((long long)jH * nk + kH) - own_off : 0;
  const int own_e = (2 * ty + 1) * SROW + tx + 1;   // element index of node A inside a slot
  const int halo_e = hr * SROW + hc;
  const int xown_e = 2 * ty * TK + tx;
  const unsigned sr_base = (unsigned)__cvta_generic_to_shared(&sm.r[0][0][0]);
  const unsigned sp_base = (unsigned)__cvta_generic_to_shared(&sm.p[0][0][0]);
  const unsigned sx_base = (unsigned)__cvta_generic_to_shared(&sm.x[0][0][0]);

  // boundary class (y,z part) of the three nodes this thread combines
  const int clsA = ((jA == 0 || jA == nj - 1) ? 2 : 0) + ((k == 0 || k == nk - 1) ? 1 : 0);
  const int clsB = ((jB == 0 || jB == nj - 1) ? 2 : 0) + ((k == 0 || k == nk - 1) ? 1 : 0);
  const int clsH = ((jH == 0 || jH == nj - 1) ? 2 : 0) + ((kH == 0 || kH == nk - 1) ? 1 : 0);

  // scalars of the iteration (device resident; written by the reduction epilogues)
  const double beta = (s.S[S_ITS] == 0.0) ? 0.0 : s.S[S_RZ] / s.S[S_RZ_OLD];
  const bool xpend = s.S[S_XPEND] != 0.0;
  const double alpha_prev = xpend ? s.S[S_ALPHA] : 0.0;

  long long off_c = (long long)i_first * plane + own_off;  // own node A in the plane being combined
  int ipl = i_first;

#define DPP_ISSUE(SL)                                                                                     \
  {                                                                                                       \
    const bool in = (unsigned)ipl < (unsigned)ni;                                                         \
    const bool xin = in && ipl >= i_lo && ipl < i_hi && xpend;                                            \
    const long long o = off_c + (long long)(ipl - ip) * plane;                                            \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                      \
      const unsigned so = (unsigned)((((SL)*NF + f) * SLOT + own_e) * 8);                                 \
      cp_async8(sr_base + so, s.r[f] + o, in && ownA_ok);                                                 \
      cp_async8(sr_base + so + SROW * 8, s.r[f] + o + nk, in && ownB_ok);                                 \
      cp_async8(sp_base + so, s.pin[f] + o, in && ownA_ok);                                               \
      cp_async8(sp_base + so + SROW * 8, s.pin[f] + o + nk, in && ownB_ok);                               \
      if (is_halo) {                                                                                      \
        const unsigned sh = (unsigned)((((SL)*NF + f) * SLOT + halo_e) * 8);                              \
        cp_async8(sr_base + sh, s.r[f] + o + halo_delta, in && halo_ok);                                  \
        cp_async8(sp_base + sh, s.pin[f] + o + halo_delta, in && halo_ok);                                \
      }                                                                                                   \
      const unsigned sxo = (unsigned)((((SL)*NF + f) * XSLOT + xown_e) * 8);                              \
      cp_async8(sx_base + sxo, s.x[f] + o, xin && actA);                                                  \
      cp_async8(sx_base + sxo + TK * 8, s.x[f] + o + nk, xin && actB);                                    \
    }                                                                                                     \
    cp_async_commit();                                                                                    \
    ++ipl;                                                                                                \
  }

  const double myo = s.mo[1], mzo = s.mo[2], kyo = s.ko[1], kzo = s.ko[2];
  const double mCor = myo * mzo, kCor = kyo * mzo + myo * kzo;
  double mEJ = 0, kEJ = 0;
  double mEK[2] = {0, 0}, kEK[2] = {0, 0};
  double mC[2] = {0, 0}, kC[2] = {0, 0};
  if (k < nk) {
    const double mzc = __ldg(&s.m1d[2][k * 3 + 1]), kzc = __ldg(&s.k1d[2][k * 3 + 1]);
    mEJ = myo * mzc;
    kEJ = kyo * mzc + myo * kzc;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int jr = jA + r;
      if (jr < nj) {
        const double myc = __ldg(&s.m1d[1][jr * 3 + 1]), kyc = __ldg(&s.k1d[1][jr * 3 + 1]);
        mEK[r] = myc * mzo;
        kEK[r] = kyc * mzo + myc * kzo;
        mC[r] = myc * mzc;
        kC[r] = kyc * mzc + myc * kzc;
      }
    }
  }
  const double mxo = s.mo[0], kxo = s.ko[0];

  double qc[NF][2][3], qd[NF][2][3], prev_cen[NF][2];
#pragma unroll
  for (int f = 0; f < NF; ++f)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int d = 0; d < 3; ++d) qc[f][r][d] = qd[f][r][d] = 0.0;
      prev_cen[f][r] = 0.0;
    }
  double dot = 0.0;
  const double* tbase = &sm.p[0][0][2 * ty * SROW + tx];

  int ip = i_first;
  __syncthreads();  // dtab visible
  // interior-plane reciprocal diagonals of the three nodes (boundary planes re-read the table)
  double dIA[NF], dIB[NF], dIH[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    dIA[f] = sm.dtab[f * 8 + clsA];
    dIB[f] = sm.dtab[f * 8 + clsB];
    dIH[f] = sm.dtab[f * 8 + clsH];
  }

  DPP_ISSUE(2)
  DPP_ISSUE(0)

#define DPP_STEP(A, B, C)                                                                             \
  {                                                                                                   \
    cp_async_wait<RING - 2>();                                                                        \
    const bool pbnd = (ip == 0 && s.dom_lo) || (ip == ni - 1 && s.dom_hi);                            \
    const bool pwr = ip >= pw_lo && ip < pw_hi;                                                       \
    const bool xwr = xpend && ip >= i_lo && ip < i_hi;                                                \
    double cen[NF][2];                                                                                \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      double* sp = &sm.p[C][f][0];                                                                    \
      const double* sr = &sm.r[C][f][0];                                                              \
      const double dA = pbnd ? sm.dtab[f * 8 + 4 + clsA] : dIA[f];                                    \
      const double dB = pbnd ? sm.dtab[f * 8 + 4 + clsB] : dIB[f];                                    \
      const double poA = sp[own_e], poB = sp[own_e + SROW];                                           \
      const double pnA = fma(beta, poA, dA * sr[own_e]);                                              \
      const double pnB = fma(beta, poB, dB * sr[own_e + SROW]);                                       \
      sp[own_e] = pnA;                                                                                \
      sp[own_e + SROW] = pnB;                                                                         \
      if (is_halo) {                                                                                  \
        const double dH = pbnd ? sm.dtab[f * 8 + 4 + clsH] : dIH[f];                                  \
        sp[halo_e] = fma(beta, sp[halo_e], dH * sr[halo_e]);                                          \
      }                                                                                               \
      if (pwr) {                                                                                      \
        if (actA) s.pout[f][off_c] = pnA;                                                             \
        if (actB) s.pout[f][off_c + nk] = pnB;                                                        \
      }                                                                                               \
      if (xwr) {                                                                                      \
        const double* sx = &sm.x[C][f][0];                                                            \
        if (actA) s.x[f][off_c] = fma(alpha_prev, poA, sx[xown_e]);                                   \
        if (actB) s.x[f][off_c + nk] = fma(alpha_prev, poB, sx[xown_e + TK]);                         \
      }                                                                                               \
      cen[f][0] = pnA;                                                                                \
      cen[f][1] = pnB;                                                                                \
    }                                                                                                 \
    __syncthreads();                                                                                  \
    DPP_ISSUE(B)                                                                                      \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      const double* t = tbase + ((C)*NF + f) * SLOT;                                                  \
      const double e0 = t[0] + t[2], c0 = t[1];                                                       \
      const double e1 = t[SROW] + t[SROW + 2], c1 = t[SROW + 1];                                      \
      const double e2 = t[2 * SROW] + t[2 * SROW + 2], c2 = t[2 * SROW + 1];                          \
      const double e3 = t[3 * SROW] + t[3 * SROW + 2], c3 = t[3 * SROW + 1];                          \
      const double corA = e0 + e2, ejA = c0 + c2, corB = e1 + e3, ejB = c1 + c3;                      \
      qc[f][0][C] = fma(mCor, corA, fma(mEK[0], e1, fma(mEJ, ejA, mC[0] * c1)));                      \
      qd[f][0][C] = fma(kCor, corA, fma(kEK[0], e1, fma(kEJ, ejA, kC[0] * c1)));                      \
      qc[f][1][C] = fma(mCor, corB, fma(mEK[1], e2, fma(mEJ, ejB, mC[1] * c2)));                      \
      qd[f][1][C] = fma(kCor, corB, fma(kEK[1], e2, fma(kEJ, ejB, kC[1] * c2)));                      \
    }                                                                                                 \
    if (ip > i_lo) { /* output plane io = ip-1 in [i_lo, i_hi) */                                     \
      const bool bnd = (ip == 1) || (ip == ni);                                                       \
      const double mxc = bnd ? s.mxc_b : s.mxc_i, kxc = bnd ? s.kxc_b : s.kxc_i;                      \
      double Kx[NF][2], Mx[NF][2];                                                                    \
      _Pragma("unroll") for (int f = 0; f < NF; ++f) _Pragma("unroll") for (int r = 0; r < 2; ++r) {  \
        const double sc = qc[f][r][A] + qc[f][r][C], sd = qd[f][r][A] + qd[f][r][C];                  \
        Mx[f][r] = fma(mxo, sc, mxc * qc[f][r][B]);                                                   \
        Kx[f][r] = fma(kxo, sc, fma(kxc, qc[f][r][B], fma(mxo, sd, mxc * qd[f][r][B])));              \
      }                                                                                               \
      _Pragma("unroll") for (int f = 0; f < NF; ++f) _Pragma("unroll") for (int r = 0; r < 2; ++r) {  \
        double yv = 0.0;                                                                              \
        _Pragma("unroll") for (int g = 0; g < NF; ++g) {                                              \
          yv = fma(s.c.cK[f][g], Kx[g][r], yv);                                                       \
          yv = fma(s.c.cM[f][g], Mx[g][r], yv);                                                       \
        }                                                                                             \
        if (r == 0 ? actA : actB) {                                                                   \
          s.w[f][off_c - plane + r * nk] = yv;                                                        \
          dot = fma(prev_cen[f][r], yv, dot);                                                         \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      prev_cen[f][0] = cen[f][0];                                                                     \
      prev_cen[f][1] = cen[f][1];                                                                     \
    }                                                                                                 \
    off_c += plane;                                                                                   \
  }

  while (true) {
    DPP_STEP(0, 1, 2)
    if (++ip > i_hi) break;
    DPP_STEP(1, 2, 0)
    if (++ip > i_hi) break;
    DPP_STEP(2, 0, 1)
    if (++ip > i_hi) break;
  }
#undef DPP_STEP
#undef DPP_ISSUE
  cp_async_wait<0>();

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (tx == 0) sm.red[ty] = dot;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TY; ++w) t += sm.red[w];
    s.dot_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// r -= alpha w ; z = dinv .* r (registers only) ; partial <r,z>, <z,z>.   dinv from the class table.
// ---------------------------------------------------------------------------------------------
struct RArgs {
  VecLayout L;
  double* r;
  const double* w;
  const double* S;
  const double* dtab;
  double* partials;
  unsigned nj, nk, ni;
  unsigned long long mag_k, mag_j;   // magic multipliers: q / nk = (q * mag_k) >> sh_k   for q < 2^31
  unsigned sh_k, sh_j;
  int dom_lo, dom_hi;
};

constexpr int VT = 256;
constexpr int UNROLL = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < VT / 32; ++w) t += sm[w];
  }
  return t;
}

__device__ __forceinline__ int node_class(const RArgs& a, unsigned q) {
  const unsigned t = (unsigned)(((unsigned long long)q * a.mag_k) >> a.sh_k);   // q / nk
  const unsigned k = q - t * a.nk;
  const unsigned i = (unsigned)(((unsigned long long)t * a.mag_j) >> a.sh_j);   // t / nj
  const unsigned j = t - i * a.nj;
  const int bx = ((i == 0 && a.dom_lo) || (i == a.ni - 1 && a.dom_hi)) ? 4 : 0;
  const int by = (j == 0 || j == a.nj - 1) ? 2 : 0;
  const int bz = (k == 0 || k == a.nk - 1) ? 1 : 0;
  return bx + by + bz;
}

__global__ void __launch_bounds__(VT) k_cg_r_update(const RArgs a) {
  __shared__ double sm[VT / 32];
  __shared__ double tab[16];
  if (a.S[S_REASON] != 0.0) return;
  if (threadIdx.x < 16) tab[threadIdx.x] = (threadIdx.x < 8 * a.L.nf) ? a.dtab[threadIdx.x] : 1.0;
  __syncthreads();
  const double alpha = a.S[S_ALPHA];
  const long long nown = a.L.oe - a.L.ob;
  const long long per = (nown + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * per;
  const long long e = b + per < nown ? b + per : nown;
  const int f = blockIdx.y;
  const long long base = (long long)f * a.L.stride;
  const double* tf = tab + f * 8;
  double srz = 0.0, szz = 0.0;
  long long q = a.L.ob + b + threadIdx.x;
  const long long qe = a.L.ob + (e > b ? e : b);
  for (; q + (UNROLL - 1) * VT < qe; q += UNROLL * VT) {
    double rv[UNROLL], wv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      rv[u] = a.r[base + q + u * VT];
      wv[u] = a.w[base + q + u * VT];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const double rn = fma(-alpha, wv[u], rv[u]);
      a.r[base + q + u * VT] = rn;
      const double zv = tf[node_class(a, (unsigned)(q + u * VT))] * rn;
      srz = fma(rn, zv, srz);
      szz = fma(zv, zv, szz);
    }
  }
  for (; q < qe; q += VT) {
    const double rn = fma(-alpha, a.w[base + q], a.r[base + q]);
    a.r[base + q] = rn;
    const double zv = tf[node_class(a, (unsigned)q)] * rn;
    srz = fma(rn, zv, srz);
    szz = fma(zv, zv, szz);
  }
  const double t0 = block_sum(srz, sm);
  const double t1 = block_sum(szz, sm);
  if (threadIdx.x == 0) {
    const size_t bb = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[bb * 2] = t0;
    a.partials[bb * 2 + 1] = t1;
  }
}

// z0 = dinv .* r with the class table (start of the solve: <r,z>, <z,z> of iteration 0)
__global__ void __launch_bounds__(VT) k_rz_init(const RArgs a) {
  __shared__ double sm[VT / 32];
  __shared__ double tab[16];
  if (threadIdx.x < 16) tab[threadIdx.x] = (threadIdx.x < 8 * a.L.nf) ? a.dtab[threadIdx.x] : 1.0;
  __syncthreads();
  const long long nown = a.L.oe - a.L.ob;
  const long long per = (nown + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * per;
  const long long e = b + per < nown ? b + per : nown;
  const int f = blockIdx.y;
  const long long base = (long long)f * a.L.stride;
  double srz = 0.0, szz = 0.0;
  for (long long q = a.L.ob + b + threadIdx.x; q < a.L.ob + e; q += VT) {
    const double rn = a.r[base + q];
    const double zv = tab[f * 8 + node_class(a, (unsigned)q)] * rn;
    srz = fma(rn, zv, srz);
    szz = fma(zv, zv, szz);
  }
  const double t0 = block_sum(srz, sm);
  const double t1 = block_sum(szz, sm);
  if (threadIdx.x == 0) {
    const size_t bb = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[bb * 2] = t0;
    a.partials[bb * 2 + 1] = t1;
  }
}

// x += alpha p for the update still pending when the loop stops
__global__ void __launch_bounds__(VT) k_cg_x_finalize(VecLayout L, double* __restrict__ x, const double* __restrict__ p,
                                                       const double* __restrict__ S) {
  if (S[S_XPEND] == 0.0) return;
  const double alpha = S[S_ALPHA];
  const long long nown = L.oe - L.ob;
  const long long base = (long long)blockIdx.y * L.stride + L.ob;
  for (long long q = (long long)blockIdx.x * VT + threadIdx.x; q < nown; q += (long long)gridDim.x * VT)
    x[base + q] = fma(alpha, p[base + q], x[base + q]);
}

// reciprocal diagonal per boundary class; same expression as k_diag_structured (apply_structured.cu)
__global__ void k_dinv_table(GridDesc g, Coef c, int nf, int jacobi, double* __restrict__ tab) {
  const int t = threadIdx.x;
  if (t >= 8 * nf) return;
  if (!jacobi) { tab[t] = 1.0; return; }
  const int f = t >> 3, bx = (t >> 2) & 1, by = (t >> 1) & 1, bz = t & 1;
  // centre entries: row 0 of the assembled 1-D matrices is a domain-boundary row; on a uniform axis an
  // interior row's centre entry is exactly twice that (two cells instead of one); dummy axis: M=[1], K=[0]
  const int b[3] = {bx, by, bz};
  double m[3], k[3];
  for (int a = 0; a < 3; ++a) {
    const double mb = g.m1d[a][1], kb = g.k1d[a][1];
    const bool dummy = g.n[a] == 1;
    m[a] = (b[a] || dummy) ? mb : 2.0 * mb;
    k[a] = (b[a] || dummy) ? kb : 2.0 * kb;
  }
  const double mxc = m[0], kxc = k[0], myc = m[1], kyc = k[1], mzc = m[2], kzc = k[2];
  const double K = kxc * myc * mzc + mxc * kyc * mzc + mxc * myc * kzc;
  const double M = mxc * myc * mzc;
  const double d = c.cK[f][f] * K + c.cM[f][f] * M;
  tab[t] = 1.0 / d;
}

void magic_div(unsigned d, unsigned long long* mag, unsigned* sh) {
  // exact floor(q / d) for all q < 2^31 (proof in DESIGN.md): k = 32 + ceil(log2 d), mag = ceil(2^k / d)
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned k = 32 + l;
  *sh = k;
  *mag = (unsigned long long)((((unsigned __int128)1 << k) + d - 1) / d);
}

}  // namespace

bool cg_fused_available(const dpp_context* ctx, int nf, int operator_mode, int pc_type) {
  return ctx->family == DPP_KERNEL_STRUCTURED && ctx->grid.band == 1 && ctx->grid_uniform && !ctx->force_table_kernel &&
         operator_mode == DPP_OP_MATRIX_FREE && (pc_type == DPP_PC_NONE || pc_type == DPP_PC_JACOBI) &&
         (nf == 1 || nf == 2) && getenv("DPP_NO_FUSED_CG") == nullptr;
}

static int make_rargs(dpp_context* ctx, const VecLayout& L, double* r, const double* w, int slot, const double* dtab,
                      RArgs* out) {
  const GridDesc& g = ctx->grid;
  RArgs a{};
  a.L = L;
  a.r = r;
  a.w = w;
  a.S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  a.dtab = dtab;
  a.partials = ctx->d_partials;
  a.ni = (unsigned)g.n[0]; a.nj = (unsigned)g.n[1]; a.nk = (unsigned)g.n[2];
  magic_div(a.nk, &a.mag_k, &a.sh_k);
  magic_div(a.nj, &a.mag_j, &a.sh_j);
  a.dom_lo = ctx->dom_lo;
  a.dom_hi = ctx->dom_hi;
  *out = a;
  return DPP_OK;
}

int cg_fused_table(dpp_context* ctx, const Coef& c, int nf, int pc_type, double* d_tab) {
  k_dinv_table<<<1, 16, 0, ctx->stream>>>(ctx->grid, c, nf, pc_type == DPP_PC_JACOBI ? 1 : 0, d_tab);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

int cg_fused_rz_init(dpp_context* ctx, const VecLayout& L, const double* r, int slot, const double* dtab, int* nblocks) {
  RArgs a{};
  DPP_CHECK(make_rargs(ctx, L, const_cast<double*>(r), nullptr, slot, dtab, &a));
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_rz_init<<<grid, VT, 0, ctx->stream>>>(a);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  *nblocks = grid.x * grid.y;
  return DPP_OK;
}

int cg_fused_r_update(dpp_context* ctx, const VecLayout& L, double* r, const double* w, int slot, const double* dtab,
                      int* nblocks) {
  RArgs a{};
  DPP_CHECK(make_rargs(ctx, L, r, w, slot, dtab, &a));
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_cg_r_update<<<grid, VT, 0, ctx->stream>>>(a);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  *nblocks = grid.x * grid.y;
  return DPP_OK;
}

int cg_fused_x_finalize(dpp_context* ctx, const VecLayout& L, double* x, const double* p, int slot) {
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_cg_x_finalize<<<grid, VT, 0, ctx->stream>>>(L, x, p, ctx->d_scalars + (size_t)slot * S_SLOT_SIZE);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

// w = A p with p = dinv.*r + beta*pin formed on chip; pout = p; x += alpha_prev * pin; partial <p,w>
int cg_fused_apply(dpp_context* ctx, int nf, const Coef& c, const double* const* r, const double* const* pin,
                   double* const* pout, double* const* x, double* const* w, int slot, const double* dtab,
                   int* n_partial_blocks) {
  const GridDesc& g = ctx->grid;
  const long long plane = (long long)g.n[1] * g.n[2];
  if (ctx->owned_begin % plane || ctx->owned_end % plane) {
    ctx->set_error("fused CG: owned range must consist of whole x-planes");
    return DPP_ERR_INVALID;
  }
  FArgs s{};
  for (int d = 0; d < 3; ++d) {
    s.n[d] = g.n[d]; s.m1d[d] = g.m1d[d]; s.k1d[d] = g.k1d[d];
    s.mo[d] = ctx->uni_m_off[d]; s.ko[d] = ctx->uni_k_off[d];
  }
  s.mxc_i = ctx->uni_mxc[0]; s.mxc_b = ctx->uni_mxc[1];
  s.kxc_i = ctx->uni_kxc[0]; s.kxc_b = ctx->uni_kxc[1];
  for (int f = 0; f < nf; ++f) { s.r[f] = r[f]; s.pin[f] = pin[f]; s.pout[f] = pout[f]; s.x[f] = x[f]; s.w[f] = w[f]; }
  s.c = c;
  s.dot_partials = ctx->d_partials;
  s.i_begin = (int)(ctx->owned_begin / plane);
  s.i_end = (int)(ctx->owned_end / plane);
  s.S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  s.dtab = dtab;
  s.dom_lo = ctx->dom_lo;
  s.dom_hi = ctx->dom_hi;
  s.ntk = (g.n[2] + TK - 1) / TK;
  s.ntj = (g.n[1] + TJ - 1) / TJ;
  const int tiles = s.ntk * s.ntj;
  const int nown = s.i_end - s.i_begin;
  if (nown <= 0) { *n_partial_blocks = 0; return DPP_OK; }
  const int capacity = ctx->sm_count * 2;
  int nseg = (2 * capacity + tiles / 2) / tiles;
  nseg = std::max(1, std::min(nseg, std::max(1, nown / 8)));
  while ((long long)tiles * nseg > kMaxPartialBlocks && nseg > 1) --nseg;
  if ((long long)tiles * nseg > kMaxPartialBlocks * (long long)kMaxDotWidth) {
    ctx->set_error("fused CG: too many tiles for the reduction scratch");
    return DPP_ERR_INVALID;
  }
  s.nseg = nseg;
  dim3 grid(tiles, nseg), block(TK, TY);
  static bool attr_set[2] = {false, false};
  if (nf == 2) {
    if (!attr_set[1]) {
      DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<2>)));
      attr_set[1] = true;
    }
    k_cg_fused_apply<2><<<grid, block, sizeof(Smem<2>), ctx->stream>>>(s);
  } else {
    if (!attr_set[0]) {
      DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<1>)));
      attr_set[0] = true;
    }
    k_cg_fused_apply<1><<<grid, block, sizeof(Smem<1>), ctx->stream>>>(s);
  }
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  *n_partial_blocks = tiles * nseg;
  return DPP_OK;
}

}  // namespace dpp
