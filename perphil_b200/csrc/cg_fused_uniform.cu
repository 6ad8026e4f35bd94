// One Jacobi/unpreconditioned CG iteration in two kernels on uniform tensor grids (the headline
// path of BASELINE.json: 3-D hex Q1, matrix-free, Jacobi-CG).  PETSc's KSPCG loop (solver.py:71)
//     p = z + beta p ; w = A p ; <p,w> ; x += alpha p ; r -= alpha w ; z = D^-1 r ; <r,z>, <z,z>
// needs two global reductions per iteration, so two kernels are the minimum without a grid-wide
// barrier.  HBM passes (one pass = one field-blocked vector, 8 B/dof):
//   k_cg_fused_apply  reads r, p_old, x (3)   writes p, w = A p, x (3)      [x update of the previous
//                     iteration is deferred into it: p_old is on chip anyway]
//   k_cg_r_update     reads r, w (2)          writes r (1)                  [z = D^-1 r only in registers]
// = 9 passes per iteration instead of 13 for the unfused sequence (apply 2 + xr-update 7 + p-update 4).
// Deferred x update (default; DPP_NO_DEFER_X=1 restores the scheme above): x is not touched every iteration.
// The directions p_k go round a ring of kXRing = 16 buffers, their step lengths alpha_k into a small device
// ring; every 15th iteration the r-update kernel adds the last fifteen alpha_k p_k to x, in iteration order
// (bitwise the x of KSPCG).  The apply kernel then moves r, p_old in and p, w out (4 passes), x costs
// (15 + 2) / 15 = 1.13 passes per iteration instead of 2: 8.13 passes per iteration; the ring takes 16 x 575 MB
// of the 180 GB of HBM at 256^3.
// Residual update with the stencil in it (default for full-boundary Dirichlet sets; DPP_NO_STENCIL_RUPD=1 restores the
// scheme above): w = A p is not stored.  k_cg_fused_apply<NF, 2> -- the same plane-streaming kernel in its third mode
// -- reads the direction the iteration kernel stored (with halo) and the r tiles, recomputes w with the identical
// arithmetic and forms r, z, <r,z>, <z,z> (+ halo push, + the x accumulation): 3 + 3 + 1.13 = 7.13 passes.
// Degree 2 (second half of this file): k_cg_fused_apply_q2, cp.async plane ring on the same padded layout.
// The reciprocal diagonal is never read from memory: on a uniform grid diag(A) takes one of 8 values
// per field (node on the domain boundary of axis x/y/z or not), kept in a 16-entry table.
//
// Data layout.  The vectors of this path live in a PADDED layout private to the library: row pitch
// (z direction) rounded up to 16 doubles, so that every row starts on a 128-byte boundary and the
// arrays are legal TMA tensors (16-byte strides).  The pad columns are zero and stay zero.
//
// Data movement.  Plane tiles (18 x 34 nodes incl. halo, both fields) are fetched with ONE TMA tensor
// load per array (cp.async.bulk.tensor.4d, SASS UTMALDG) into a 3-slot shared-memory ring, completion
// on an mbarrier per slot; out-of-domain halo nodes are zero-filled by the TMA unit.  No per-thread
// address arithmetic, no LSU traffic for the loads (the cp.async/LDGSTS predecessor of this kernel was
// bound by the LSU pipe and by 8-byte copy granularity at ~50 % of HBM peak; profiles/).
// The stencil part is the one of apply_structured_uniform.cu: two nodes per thread, in-plane
// neighbour-class sums, register queue along x, fused <p, A p>.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "cg_device.cuh"

namespace dpp {

namespace {

constexpr int kClsPerField = 64;   // reciprocal-diagonal classes per field: (px py pz) * 8 + (bx by bz)
constexpr int TK = 32;
constexpr int TY = 8;
constexpr int TJ = 2 * TY;
constexpr int NT = TK * TY;
// TMA needs the innermost box coordinate on a 16-byte boundary (probed on B200: an odd fp64 coordinate
// raises 'illegal instruction'), so tiles start at even k and the halo'd box starts at k0-2: 36 columns,
// column c of a slot row holds node k0 - 2 + c (columns 0 and 35 are fetched but never used).
constexpr int SROW = TK + 4;
constexpr int SLOT = (TJ + 2) * SROW;   // doubles per field per ring slot (halo'd tile)
constexpr int XSLOT = TJ * TK;
constexpr int RING = 3;

struct FArgs {
  int n[3];
  int pitch;                 // padded row length (doubles)
  long long plane, field;    // padded plane / field strides (doubles)
  const double* m1d[3];
  const double* k1d[3];
  double mo[3], ko[3];
  double mxc_i, mxc_b, kxc_i, kxc_b;
  double* pout;              // p of this iteration (own + ghost planes written)   [padded, field-blocked]
  double* x;
  double* w;
  Coef c;
  double* dot_partials;
  int i_begin, i_end;
  int ntj, ntk, nseg;
  int j_lo, j_hi, k_lo, k_hi;  // nodes of a plane that are computed (class-mask mode: interior only)
  const double* S;           // scalar slot of the solver (null in plain-apply mode)
  const double* dtab;        // [2][64] reciprocal diagonal per class: (node parity px py pz) * 8 + (boundary bx by bz); Q1: parity 0
  int dom_lo, dom_hi;        // local plane 0 / n[0]-1 lies on the domain boundary (else it is a ghost plane)
  int defer_x;               // x is updated by the r-update kernel from the direction ring, not here
  FoldArgs fold;             // <p,Ap> reduction epilogue run by the last CTA
  // MODE 1 with no_w: w = A p is not stored (the MODE 2 kernel that follows recomputes it).
  // MODE 2 (r-update with the stencil in it): r <- r - alpha A p on the computed nodes, z = D^-1 r, <r,z>, <z,z>
  int no_w;
  double* r;                 // residual, padded (MODE 2: read through tm_x tiles, written here)
  double* partials2;         // [blocks][2]
  IpcHalo halo;              // peer residual vectors: the first / last owned plane is stored there as well
  int own_first, own_last;   // first / last owned plane of the slab (halo push), -1: no neighbour on that side
  long long ob, oe;          // owned padded range (deferred x accumulation)
  const double* pr[kXRing];  // direction ring
  const double* xring;       // step lengths of the ring
};

__device__ __forceinline__ int bstart(int t, int n, int nt) { return (int)(((long long)t * n) / nt); }

__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n .reg .pred P1;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned mbar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// programmatic dependent launch (PDL): let the next kernel of the stream start launching while this one
// drains, and wait for the previous kernel's results only where they are first needed
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <int NF>
struct __align__(128) Smem {
  static constexpr int RS = ((NF * SLOT * 8 + 127) / 128) * 16;   // doubles per slot, 128-byte multiple
  double r[RING][RS];
  double p[RING][RS];
  double x[RING][NF * XSLOT];
  unsigned long long mbar[RING];
  double red[TY];
  double dtab[16];
  double fin[kFinishSmem];
  double spre[kPreScalars];          // scalar slot as of the kernel start (epilogue reads it: cg_device.cuh)
  unsigned long long seq_pre;
  int last_flag;
};

// MODE 1: CG iteration kernel (p = dinv r + beta p_old formed on chip, deferred x update, p stored)
// MODE 0: plain apply w = A p on padded vectors (tm_p = input), optional <p, w>
// MODE 2: residual update with the stencil in it: w = A p is RECOMPUTED from the direction the MODE 1 kernel stored
//         (tm_p) instead of being written there and read here -- one vector pass less per iteration, the same bits --
//         r (tiles through tm_x) <- r - alpha w, z = D^-1 r in registers, <r,z>, <z,z> + KSPCG bookkeeping, halo push,
//         and every 15th launch the deferred x accumulation.  Class-mask handles only (computed nodes = free nodes).
template <int NF, int MODE>
__global__ void __launch_bounds__(NT, 2) k_cg_fused_apply(const FArgs s, const __grid_constant__ CUtensorMap tm_r,
                                                          const __grid_constant__ CUtensorMap tm_p,
                                                          const __grid_constant__ CUtensorMap tm_x) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<NF>& sm = *reinterpret_cast<Smem<NF>*>(smem_raw);
  constexpr int RS = Smem<NF>::RS;
  constexpr bool FUSED = MODE == 1, RUPD = MODE == 2;

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2];
  const int nown = s.i_end - s.i_begin;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TK + tx;
  const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(&sm.mbar[0]);
  const unsigned sr_base = (unsigned)__cvta_generic_to_shared(&sm.r[0][0]);
  const unsigned sp_base = (unsigned)__cvta_generic_to_shared(&sm.p[0][0]);
  const unsigned sx_base = (unsigned)__cvta_generic_to_shared(&sm.x[0][0]);
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < RING; ++q) mbar_init(mbar0 + 8 * q, 1);
    fence_mbar_init();
  }
  // the class table is written once per solve, before the first kernel of the iteration: safe to fetch while
  // the previous kernel of the stream is still in its reduction epilogue
  if (tid < 16) sm.dtab[tid] = (FUSED && tid < 8 * NF) ? s.dtab[(tid >> 3) * kClsPerField + (tid & 7)] : 1.0;
  pdl_wait();  // everything below reads what the previous kernel of the stream produced
  if ((FUSED || RUPD) && s.S[S_REASON] != 0.0) return;
  if ((FUSED || RUPD) && s.fold.enabled) {
    if (tid < kPreScalars) sm.spre[tid] = s.fold.S[tid];
    if (tid == kPreScalars && s.fold.ipc.world > 1) sm.seq_pre = *s.fold.ipc.seq_dev;
  }
  // scalars of the iteration (device resident; written by the reduction epilogues)
  double beta = 0.0, alpha_prev = 0.0;
  bool xpend = false;
  if (FUSED) {
    beta = (s.S[S_ITS] == 0.0) ? 0.0 : s.S[S_RZ] / s.S[S_RZ_OLD];
    xpend = !s.defer_x && s.S[S_XPEND] != 0.0;
    alpha_prev = xpend ? s.S[S_ALPHA] : 0.0;
  }
  const double alpha_r = RUPD ? s.S[S_ALPHA] : 0.0;
  double dot = 0.0, srz = 0.0, szz = 0.0;
  unsigned phase = 0;  // bit q = parity of the next completion of ring slot q
  const long long plane = s.plane;
  const int pitch = s.pitch;
  __syncthreads();

  // Balanced persistent partition: the (tile, plane) steps of the whole grid are cut into gridDim.x equal
  // contiguous ranges (one CTA each, one wave of 2 CTAs per SM); a range is processed as runs of
  // consecutive planes of one tile (2 redundant planes per run).
  const long long total = (long long)s.ntj * s.ntk * nown;
  long long wbeg, wend;
  if (s.nseg > 0) {  // one (tile, x-segment) item per CTA, tile index fastest
    const int ntiles = s.ntj * s.ntk;
    const int tile_ = blockIdx.x % ntiles, seg_ = blockIdx.x / ntiles;
    wbeg = (long long)tile_ * nown + ((long long)seg_ * nown) / s.nseg;
    wend = (long long)tile_ * nown + ((long long)(seg_ + 1) * nown) / s.nseg;
  } else {           // persistent: equal contiguous shares of the (tile, plane) steps
    wbeg = (total * blockIdx.x) / gridDim.x;
    wend = (total * (blockIdx.x + 1)) / gridDim.x;
  }
  for (; wbeg < wend;) {
    const int tile = (int)(wbeg / nown);
    const int run_a = (int)(wbeg - (long long)tile * nown);
    const int run_len = (int)((wend - wbeg) < (long long)(nown - run_a) ? (wend - wbeg) : (long long)(nown - run_a));
    wbeg += run_len;
    const int i_lo = s.i_begin + run_a, i_hi = i_lo + run_len;
    const int tkid = tile % s.ntk, tjid = tile / s.ntk;
    // tiles cover the computed node range [j_lo, j_hi) x [k_lo, k_hi); they are cut between even columns
    const int kp_lo = s.k_lo >> 1, nkp = ((s.k_hi + 1) >> 1) - kp_lo;
    const int k0 = 2 * (kp_lo + bstart(tkid, nkp, s.ntk)), k1 = min(s.k_hi, 2 * (kp_lo + bstart(tkid + 1, nkp, s.ntk)));
    const int j0 = s.j_lo + bstart(tjid, s.j_hi - s.j_lo, s.ntj), j1 = s.j_lo + bstart(tjid + 1, s.j_hi - s.j_lo, s.ntj);
    const int jA = j0 + 2 * ty, jB = jA + 1, k = k0 + tx;
    const bool actA = (jA < j1) && (k >= s.k_lo) && (k < k1), actB = (jB < j1) && (k >= s.k_lo) && (k < k1);
    const int i_first = i_lo - 1;
    // planes whose p this run stores: its output planes, plus the ghost plane next to the first/last owned one
    const int pw_lo = (i_lo == s.i_begin && s.i_begin > 0) ? i_lo - 1 : i_lo;
    const int pw_hi = (i_hi == s.i_end && s.i_end < ni) ? i_hi + 1 : i_hi;

    // producer (one thread): fetch plane PL into ring slot SL
#define DPP_ISSUE(SL, PL)                                                                              \
  if (tid == 0) {                                                                                      \
    const bool xin = ((FUSED && xpend) || RUPD) && (PL) >= i_lo && (PL) < i_hi;                        \
    const unsigned mb = mbar0 + 8 * (SL);                                                              \
    fence_proxy_async();                                                                               \
    mbar_expect_tx(mb, (unsigned)((FUSED ? 2 : 1) * NF * SLOT * 8 + (xin ? NF * XSLOT * 8 : 0)));      \
    if (FUSED) tma_load_4d(sr_base + (SL)*RS * 8, &tm_r, mb, k0 - 2, j0 - 1, (PL), 0);                 \
    tma_load_4d(sp_base + (SL)*RS * 8, &tm_p, mb, k0 - 2, j0 - 1, (PL), 0);                            \
    if (xin) tma_load_4d(sx_base + (SL)*NF * XSLOT * 8, &tm_x, mb, k0, j0, (PL), 0);                   \
  }

    // first two planes of the run are requested before the coefficient set-up below (which then overlaps the
    // load latency instead of preceding it)
    __syncthreads();  // the ring is free: every thread finished the previous run
    DPP_ISSUE(2, i_first)
    if (i_first + 1 <= i_hi) DPP_ISSUE(0, i_first + 1)


    // the elements this thread combines: its two nodes + at most one node of the halo ring
    constexpr int HW = TK + 2;               // width of the halo ring (nodes k0-1 .. k0+TK)
    constexpr int HALO = 2 * HW + 2 * TJ;
    int hr = 0, hc = 0;
    const bool is_halo = FUSED && tid < HALO;
    if (is_halo) {
      if (tid < HW) { hr = 0; hc = tid; }
      else if (tid < 2 * HW) { hr = TJ + 1; hc = tid - HW; }
      else if (tid < 2 * HW + TJ) { hr = tid - 2 * HW + 1; hc = 0; }
      else { hr = tid - 2 * HW - TJ + 1; hc = TK + 1; }
    }
    const int jH = j0 - 1 + hr, kH = k0 - 1 + hc;
    const int own_e = (2 * ty + 1) * SROW + tx + 2;
    const int halo_e = hr * SROW + hc + 1;
    const int xown_e = 2 * ty * TK + tx;
    const int clsA = ((jA == 0 || jA == nj - 1) ? 2 : 0) + ((k == 0 || k == nk - 1) ? 1 : 0);
    const int clsB = ((jB == 0 || jB == nj - 1) ? 2 : 0) + ((k == 0 || k == nk - 1) ? 1 : 0);
    const int clsH = ((jH == 0 || jH == nj - 1) ? 2 : 0) + ((kH == 0 || kH == nk - 1) ? 1 : 0);

    long long off_c = (long long)i_first * plane + (long long)jA * pitch + k;  // node A in the plane being combined

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
    double rprev[NF][2], d0[NF];   // MODE 2: r of the output plane (tile read one step earlier); free-node D^-1
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      rprev[f][0] = rprev[f][1] = 0.0;
      d0[f] = RUPD ? s.dtab[f * kClsPerField] : 0.0;
    }
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int d = 0; d < 3; ++d) qc[f][r][d] = qd[f][r][d] = 0.0;
        prev_cen[f][r] = 0.0;
      }
    const double* tbase = &sm.p[0][2 * ty * SROW + tx + 1];
    // interior-plane reciprocal diagonals of the three nodes (boundary planes re-read the table)
    double dIA[NF], dIB[NF], dIH[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      dIA[f] = sm.dtab[f * 8 + clsA];
      dIB[f] = sm.dtab[f * 8 + clsB];
      dIH[f] = sm.dtab[f * 8 + clsH];
    }

    int ip = i_first;

  // one plane step: plane ip is in ring slot C; plane ip+2 is fetched into slot B (= slot of plane ip-1)
#define DPP_STEP(A, B, C)                                                                             \
  {                                                                                                   \
    mbar_wait(mbar0 + 8 * (C), (phase >> (C)) & 1u);                                                  \
    phase ^= 1u << (C);                                                                               \
    double cen[NF][2], rcur[NF][2];                                                                   \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      rcur[f][0] = rcur[f][1] = 0.0;                                                                  \
      if (RUPD && ip >= i_lo && ip < i_hi) {                                                          \
        rcur[f][0] = sm.x[C][f * XSLOT + xown_e];                                                     \
        rcur[f][1] = sm.x[C][f * XSLOT + xown_e + TK];                                                \
      }                                                                                               \
    }                                                                                                 \
    if (FUSED) {                                                                                      \
      const bool pbnd = ni > 1 && ((ip == 0 && s.dom_lo) || (ip == ni - 1 && s.dom_hi));              \
      const bool pwr = ip >= pw_lo && ip < pw_hi;                                                     \
      const bool xwr = xpend && ip >= i_lo && ip < i_hi;                                              \
      _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                \
        double* sp = &sm.p[C][f * SLOT];                                                              \
        const double* sr = &sm.r[C][f * SLOT];                                                        \
        const double dA = pbnd ? sm.dtab[f * 8 + 4 + clsA] : dIA[f];                                  \
        const double dB = pbnd ? sm.dtab[f * 8 + 4 + clsB] : dIB[f];                                  \
        const double poA = sp[own_e], poB = sp[own_e + SROW];                                         \
        const double pnA = fma(beta, poA, dA * sr[own_e]);                                            \
        const double pnB = fma(beta, poB, dB * sr[own_e + SROW]);                                     \
        sp[own_e] = pnA;                                                                              \
        sp[own_e + SROW] = pnB;                                                                       \
        if (is_halo) {                                                                                \
          const double dH = pbnd ? sm.dtab[f * 8 + 4 + clsH] : dIH[f];                                \
          sp[halo_e] = fma(beta, sp[halo_e], dH * sr[halo_e]);                                        \
        }                                                                                             \
        double* po = s.pout + f * s.field + off_c;                                                    \
        if (pwr) {                                                                                    \
          if (actA) po[0] = pnA;                                                                      \
          if (actB) po[pitch] = pnB;                                                                  \
        }                                                                                             \
        if (xwr) {                                                                                    \
          const double* sx = &sm.x[C][f * XSLOT];                                                     \
          double* xo = s.x + f * s.field + off_c;                                                     \
          if (actA) xo[0] = fma(alpha_prev, poA, sx[xown_e]);                                         \
          if (actB) xo[pitch] = fma(alpha_prev, poB, sx[xown_e + TK]);                                \
        }                                                                                             \
        cen[f][0] = pnA;                                                                              \
        cen[f][1] = pnB;                                                                              \
      }                                                                                               \
    }                                                                                                 \
    __syncthreads();                                                                                  \
    if (ip + 2 <= i_hi) DPP_ISSUE(B, ip + 2)                                                          \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      const double* t = tbase + (C)*RS + f * SLOT;                                                    \
      const double e0 = t[0] + t[2], c0 = t[1];                                                       \
      const double e1 = t[SROW] + t[SROW + 2], c1 = t[SROW + 1];                                      \
      const double e2 = t[2 * SROW] + t[2 * SROW + 2], c2 = t[2 * SROW + 1];                          \
      const double e3 = t[3 * SROW] + t[3 * SROW + 2], c3 = t[3 * SROW + 1];                          \
      const double corA = e0 + e2, ejA = c0 + c2, corB = e1 + e3, ejB = c1 + c3;                      \
      if (!FUSED) {                                                                                   \
        cen[f][0] = c1;                                                                               \
        cen[f][1] = c2;                                                                               \
      }                                                                                               \
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
          if (RUPD) {                                                                                 \
            const long long q = off_c - plane + r * pitch;                                            \
            const double rn = fma(-alpha_r, yv, rprev[f][r]);                                         \
            s.r[f * s.field + q] = rn;                                                                \
            const int io = ip - 1;                                                                    \
            const long long inpl = q - (long long)io * plane;                                         \
            if (io == s.own_first) s.halo.peer_r[0][f * s.halo.peer_field[0] + s.halo.peer_ghost_off[0] + inpl] = rn; \
            if (io == s.own_last) s.halo.peer_r[1][f * s.halo.peer_field[1] + s.halo.peer_ghost_off[1] + inpl] = rn;  \
            const double z = d0[f] * rn;                                                              \
            srz = fma(rn, z, srz);                                                                    \
            szz = fma(z, z, szz);                                                                     \
          } else {                                                                                    \
            if (!(FUSED && s.no_w)) s.w[f * s.field + off_c - plane + r * pitch] = yv;                \
            dot = fma(prev_cen[f][r], yv, dot);                                                       \
          }                                                                                           \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      prev_cen[f][0] = cen[f][0];                                                                     \
      prev_cen[f][1] = cen[f][1];                                                                     \
      rprev[f][0] = rcur[f][0];                                                                       \
      rprev[f][1] = rcur[f][1];                                                                       \
    }                                                                                                 \
    off_c += plane;                                                                                   \
  }

    // planes i_first .. i_hi ; outputs i_lo .. i_hi-1
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
    // nothing is in flight here: every fetched plane was consumed, so the per-slot parity bits stay valid
  }  // runs

  if (RUPD) {
    if (s.defer_x) {
      // deferred x update (same rule and arithmetic as in k_cg_r_update): after every 15th iteration the fifteen
      // pending directions are added in iteration order; flat partition of the owned range over the CTAs
      const int kit = (int)s.S[S_ITS];
      if ((kit + 1) % (kXRing - 1) == 0) {
        constexpr int NDIR = kXRing - 1, CH = 8;
        const long long nown2 = (s.oe - s.ob) / 2;
        const long long b2 = (nown2 * blockIdx.x) / gridDim.x, e2 = (nown2 * (blockIdx.x + 1)) / gridDim.x;
#pragma unroll 1
        for (int f = 0; f < NF; ++f) {
          double* xf = s.x + (long long)f * s.field;
          const long long foff = (long long)f * s.field;
          for (long long t = b2 + tid; t < e2; t += NT) {
            const long long qq = s.ob + 2 * t;
            double2 xv = *reinterpret_cast<const double2*>(xf + qq);
#pragma unroll
            for (int c0 = 0; c0 < NDIR; c0 += CH) {
              double2 pv[CH];
              double al[CH];
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (c0 + j < NDIR) {
                  const int slot = (kit - (NDIR - 1) + c0 + j + 1) % kXRing;
                  al[j] = s.xring[slot];
                  pv[j] = *reinterpret_cast<const double2*>(s.pr[slot] + foff + qq);
                }
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (c0 + j < NDIR) {
                  xv.x = fma(al[j], pv[j].x, xv.x);
                  xv.y = fma(al[j], pv[j].y, xv.y);
                }
            }
            *reinterpret_cast<double2*>(xf + qq) = xv;
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      srz += __shfl_xor_sync(0xffffffffu, srz, o);
      szz += __shfl_xor_sync(0xffffffffu, szz, o);
    }
    __syncthreads();   // sm.red / sm.fin are free
    if (tx == 0) { sm.red[ty] = srz; sm.fin[ty] = szz; }
    __syncthreads();
    if (tid == 0) {
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int w = 0; w < TY; ++w) { t0 += sm.red[w]; t1 += sm.fin[w]; }
      s.partials2[(size_t)blockIdx.x * 2] = t0;
      s.partials2[(size_t)blockIdx.x * 2 + 1] = t1;
    }
    const bool remote = s.halo.peer_r[0] != nullptr || s.halo.peer_r[1] != nullptr;
    if (s.fold.enabled && last_block_arrives(s.fold.counter, gridDim.x, &sm.last_flag))
      finish_reduction(s.partials2, (int)gridDim.x, 2, s.fold.S, s.fold.hist, s.fold.post, 0, s.fold.ipc, sm.fin, remote,
                       FoldPre{sm.spre, &sm.seq_pre});
    return;
  }
  if (s.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (tx == 0) sm.red[ty] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < TY; ++w) t += sm.red[w];
      s.dot_partials[blockIdx.x] = t;
    }
    if (FUSED && s.fold.enabled && last_block_arrives(s.fold.counter, gridDim.x, &sm.last_flag))
      finish_reduction(s.dot_partials, (int)gridDim.x, 1, s.fold.S, s.fold.hist, s.fold.post, 0, s.fold.ipc, sm.fin,
                       s.fold.ipc.ll == 3, FoldPre{sm.spre, &sm.seq_pre}, s.fold.xring);
  }
}

// ---------------------------------------------------------------------------------------------
// r -= alpha w ; z = dinv .* r (registers only) ; partial <r,z>, <z,z>.   Padded layout, dinv from
// the class table.  Pad columns hold zeros in r and w and therefore stay zero.
// Slab-partitioned runs with peer memory: the first / last owned plane of the new r is ALSO stored into
// the ghost plane of the lower / upper neighbour (fused halo push over NVLink; the mailbox reduction
// that follows publishes its flag after a system fence, which is what tells the neighbour the plane
// has landed).
// ---------------------------------------------------------------------------------------------
struct RArgs {
  int nf;
  long long field;           // padded field stride
  long long ob, oe;          // owned padded range [i_begin*plane, i_end*plane)
  double* r;
  const double* w;
  const double* S;
  const double* dtab;
  double* partials;
  unsigned nj, nk, ni, pitch;
  unsigned long long mag_k, mag_j;   // magic multipliers: q / pitch = (q * mag_k) >> sh_k   for q < 2^31
  unsigned sh_k, sh_j;
  int dom_lo, dom_hi;
  long long plane;           // padded plane (doubles)
  int zero_class[2];         // field f: every domain-boundary node is a Dirichlet node (and no other)
  FoldArgs fold;
  IpcHalo halo;              // peer residual vectors: boundary planes are stored there as well
  int q2;                    // degree-2 lattice: odd indices are mid nodes (their own diagonal class)
  int defer_x;               // deferred x update: every (kXRing - 1)-th iteration x += sum alpha_k p_k from the ring
  double* x;
  const double* pr[kXRing];  // direction ring (padded, field-blocked)
  const double* xring;       // [kXRing] alpha, [kXRing] iteration tag
  long long up_win;          // doubles at the end of the owned range the upper neighbour mirrors (degree planes)
};

constexpr int UNROLL = 4;   // pairs of nodes per thread and loop trip: 2 arrays x 4 x 16 B = 128 B in flight per thread

__device__ __forceinline__ void push_halo(const RArgs& a, int f, long long q, double v) {
  if (a.halo.peer_r[0] != nullptr && q < a.ob + a.plane)
    a.halo.peer_r[0][f * a.halo.peer_field[0] + a.halo.peer_ghost_off[0] + (q - a.ob)] = v;
  if (a.halo.peer_r[1] != nullptr && q >= a.oe - a.up_win)
    a.halo.peer_r[1][f * a.halo.peer_field[1] + a.halo.peer_ghost_off[1] + (q - (a.oe - a.up_win))] = v;
}
// the same for an aligned pair (q even; a pair never straddles a plane: planes hold an even number of doubles)
__device__ __forceinline__ void push_halo2(const RArgs& a, int f, long long q, double2 v) {
  if (a.halo.peer_r[0] != nullptr && q < a.ob + a.plane)
    *reinterpret_cast<double2*>(a.halo.peer_r[0] + f * a.halo.peer_field[0] + a.halo.peer_ghost_off[0] + (q - a.ob)) = v;
  if (a.halo.peer_r[1] != nullptr && q >= a.oe - a.up_win)
    *reinterpret_cast<double2*>(a.halo.peer_r[1] + f * a.halo.peer_field[1] + a.halo.peer_ghost_off[1] +
                                (q - (a.oe - a.up_win))) = v;
}

__device__ __forceinline__ int node_class(const RArgs& a, unsigned q) {
  const unsigned t = (unsigned)(((unsigned long long)q * a.mag_k) >> a.sh_k);   // q / pitch
  const unsigned k = q - t * a.pitch;
  const unsigned i = (unsigned)(((unsigned long long)t * a.mag_j) >> a.sh_j);   // t / nj
  const unsigned j = t - i * a.nj;
  const int bx = (a.ni > 1 && ((i == 0 && a.dom_lo) || (i == a.ni - 1 && a.dom_hi))) ? 4 : 0;  // dummy axis: no class
  const int by = (j == 0 || j == a.nj - 1) ? 2 : 0;
  const int bz = (k == 0 || k == a.nk - 1) ? 1 : 0;
  const int par = a.q2 ? (int)(((i & 1u) << 2) | ((j & 1u) << 1) | (k & 1u)) : 0;
  return par * 8 + bx + by + bz;
}

// Streaming kernel: every thread walks its block's range in aligned PAIRS of nodes (16-byte loads and stores;
// the padded pitch is a multiple of 16 doubles, so a pair never leaves its row) and keeps the lattice
// position (i, j, k) of its pair incrementally -- one magic division per thread, not per node.
template <bool INIT>
__global__ void __launch_bounds__(VT, 3) k_cg_r_update(const RArgs a) {
  __shared__ double sm[kFinishSmem];
  __shared__ double tab[2 * kClsPerField];
  __shared__ double spre[kPreScalars];
  __shared__ unsigned long long seq_pre;
  __shared__ int last_flag;
  pdl_launch_dependents();
  if (threadIdx.x < 2 * kClsPerField) tab[threadIdx.x] = (threadIdx.x < kClsPerField * a.nf) ? a.dtab[threadIdx.x] : 1.0;   // per-solve constant
  pdl_wait();
  if (!INIT && a.S[S_REASON] != 0.0) return;
  if (a.fold.enabled) {   // snapshot for the epilogue (cg_device.cuh: apply_post / finish_reduction)
    if (threadIdx.x < kPreScalars) spre[threadIdx.x] = a.fold.S[threadIdx.x];
    if (threadIdx.x == kPreScalars && a.fold.ipc.world > 1) seq_pre = *a.fold.ipc.seq_dev;
  }
  __syncthreads();
  const double alpha = INIT ? 0.0 : a.S[S_ALPHA];
  const long long nown = a.oe - a.ob;                                      // even (whole padded planes)
  const long long per = 2 * ((nown / 2 + gridDim.x - 1) / gridDim.x);      // even share per block
  const long long b = (long long)blockIdx.x * per;
  const long long e = b + per < nown ? b + per : nown;
  const int f = blockIdx.y;
  double* rf = a.r + (long long)f * a.field;
  const double* wf = INIT ? nullptr : a.w + (long long)f * a.field;
  const double* tf = tab + f * kClsPerField;
  const bool zc = a.zero_class[f] != 0;
  const bool xdim = a.ni > 1;
  double srz = 0.0, szz = 0.0;
  long long q = a.ob + b + 2 * threadIdx.x;
  const long long qe = a.ob + (e > b ? e : b);
  // lattice position of the pair at q
  unsigned k, j, i;
  {
    const unsigned qq = (unsigned)(q < qe ? q : a.ob);
    const unsigned t = (unsigned)(((unsigned long long)qq * a.mag_k) >> a.sh_k);   // q / pitch
    k = qq - t * a.pitch;
    i = (unsigned)(((unsigned long long)t * a.mag_j) >> a.sh_j);                   // t / nj
    j = t - i * a.nj;
  }
  auto advance = [&]() {   // position of the pair 2*VT doubles further on
    k += 2 * VT;
    while (k >= a.pitch) {
      k -= a.pitch;
      if (++j == a.nj) { j = 0; ++i; }
    }
  };
  auto classes = [&](int& c0, int& c1) {
    const int bx = (xdim && ((i == 0 && a.dom_lo) || (i == a.ni - 1 && a.dom_hi))) ? 4 : 0;
    int bxy = bx + ((j == 0 || j == a.nj - 1) ? 2 : 0);
    if (a.q2) bxy += (int)((((i & 1u) << 2) | ((j & 1u) << 1)) << 3);   // mid-node parities of the row
    c0 = bxy + ((k == 0 || k == a.nk - 1) ? 1 : 0) + (a.q2 ? (int)((k & 1u) << 3) : 0);
    c1 = bxy + ((k + 1 == a.nk - 1) ? 1 : 0) + (a.q2 ? (int)(((k + 1) & 1u) << 3) : 0);
  };
  auto one = [&](long long qq, double2 rv, double2 wv) {
    int c0, c1;
    classes(c0, c1);
    double2 rn = rv;
    if (!INIT) {
      rn.x = (zc && (c0 & 7)) ? 0.0 : fma(-alpha, wv.x, rv.x);
      rn.y = (zc && (c1 & 7)) ? 0.0 : fma(-alpha, wv.y, rv.y);
      *reinterpret_cast<double2*>(rf + qq) = rn;
      push_halo2(a, f, qq, rn);
    }
    const double z0 = tf[c0] * rn.x, z1 = tf[c1] * rn.y;
    srz = fma(rn.x, z0, srz);
    szz = fma(z0, z0, szz);
    srz = fma(rn.y, z1, srz);
    szz = fma(z1, z1, szz);
  };
  for (; q + (UNROLL - 1) * 2 * VT < qe; q += UNROLL * 2 * VT) {
    double2 rv[UNROLL], wv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      rv[u] = *reinterpret_cast<const double2*>(rf + q + u * 2 * VT);
      wv[u] = INIT ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(wf + q + u * 2 * VT);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      one(q + u * 2 * VT, rv[u], wv[u]);
      advance();
    }
  }
  for (; q < qe; q += 2 * VT) {
    one(q, *reinterpret_cast<const double2*>(rf + q),
        INIT ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(wf + q));
    advance();
  }
  if (!INIT && a.defer_x) {
    // iteration k = its (this kernel's epilogue makes it k + 1): after every (kXRing - 1)-th iteration the fifteen
    // directions p_{k-14} .. p_k are added to x in iteration order.  S[S_ITS] is written only by the last block's
    // epilogue, after every block has arrived, so all blocks read the same k here.
    const int k = (int)a.S[S_ITS];
    if ((k + 1) % (kXRing - 1) == 0) {
      constexpr int NDIR = kXRing - 1, CH = 8;     // directions per accumulate; loads in flight per chunk
      double* xf = a.x + (long long)f * a.field;
      const long long foff = (long long)f * a.field;
      for (long long qq = a.ob + b + 2 * threadIdx.x; qq < qe; qq += 2 * VT) {
        double2 xv = *reinterpret_cast<const double2*>(xf + qq);
#pragma unroll
        for (int c0 = 0; c0 < NDIR; c0 += CH) {
          double2 pv[CH];
          double al[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (c0 + j < NDIR) {
              const int slot = (k - (NDIR - 1) + c0 + j + 1) % kXRing;
              al[j] = a.xring[slot];
              pv[j] = *reinterpret_cast<const double2*>(a.pr[slot] + foff + qq);
            }
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (c0 + j < NDIR) {
              xv.x = fma(al[j], pv[j].x, xv.x);
              xv.y = fma(al[j], pv[j].y, xv.y);
            }
        }
        *reinterpret_cast<double2*>(xf + qq) = xv;
      }
    }
  }
  const double t0 = block_sum(srz, sm);
  const double t1 = block_sum(szz, sm);
  if (threadIdx.x == 0) {
    const size_t bb = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[bb * 2] = t0;
    a.partials[bb * 2 + 1] = t1;
  }
  // Halo planes stored into peer memory: every block reports in through a gpu-scope fence + the arrival
  // counter (last_block_arrives); the last block's single system-scope fence before the mailbox words is
  // cumulative over all of them (finish_reduction, sys_release).  A system fence in every pushing block
  // cost 2 us per iteration: MEMBAR.SYS stalls the whole SM, and the pushing blocks sit on every SM.
  const bool remote = !INIT && (a.halo.peer_r[0] != nullptr || a.halo.peer_r[1] != nullptr);
  if (remote && a.halo.debug_fence_all) __threadfence_system();
  if (a.fold.enabled && last_block_arrives(a.fold.counter, gridDim.x * gridDim.y, &last_flag))
    finish_reduction(a.partials, (int)(gridDim.x * gridDim.y), 2, a.fold.S, a.fold.hist, a.fold.post, 0, a.fold.ipc, sm,
                     remote || INIT, FoldPre{spre, &seq_pre});
}

// boundary planes of r -> the neighbours' ghost planes (start of a solve)
__global__ void __launch_bounds__(VT) k_push_planes(const RArgs a) {
  const int f = blockIdx.y;
  const double* rf = a.r + (long long)f * a.field;
  // first owned plane (lower neighbour) and the last up_win doubles (upper neighbour); a range may be both
  for (long long t = (long long)blockIdx.x * VT + threadIdx.x; t < a.plane + a.up_win; t += (long long)gridDim.x * VT) {
    const long long q = t < a.plane ? a.ob + t : a.oe - a.up_win + (t - a.plane);
    if (t >= a.plane && q < a.ob + a.plane) continue;   // already pushed (to both sides) by the first part
    push_halo(a, f, q, rf[q]);
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversion (private padded layout <-> the field-blocked lexicographic layout of the C ABI)
// ---------------------------------------------------------------------------------------------
struct PadGeom {
  int nf;
  long long n_nodes;     // unpadded field stride
  long long field;       // padded field stride
  unsigned nj, nk, pitch;
  unsigned long long mag_k;
  unsigned sh_k;
};

// dst_padded[f][row][k] = src[f][row*nk + k]  (pads untouched: they are zero from allocation)
__global__ void __launch_bounds__(VT) k_pad_copy(PadGeom g, const double* __restrict__ src, double* __restrict__ dst) {
  const int f = blockIdx.y;
  for (long long n = (long long)blockIdx.x * VT + threadIdx.x; n < g.n_nodes; n += (long long)gridDim.x * VT) {
    const unsigned row = (unsigned)(((unsigned long long)(unsigned)n * g.mag_k) >> g.sh_k);  // n / nk
    const unsigned k = (unsigned)n - row * g.nk;
    dst[f * g.field + (long long)row * g.pitch + k] = src[f * g.n_nodes + n];
  }
}

// x_unpadded = x_padded (+ alpha p_padded when an update is still pending)
__global__ void __launch_bounds__(VT) k_cg_x_finalize(PadGeom g, const double* __restrict__ xp, const double* __restrict__ pp,
                                                       const double* __restrict__ S, long long ob, long long oe,
                                                       double* __restrict__ x) {
  const double alpha = (S[S_XPEND] != 0.0) ? S[S_ALPHA] : 0.0;
  const int f = blockIdx.y;
  for (long long n = ob + (long long)blockIdx.x * VT + threadIdx.x; n < oe; n += (long long)gridDim.x * VT) {
    const unsigned row = (unsigned)(((unsigned long long)(unsigned)n * g.mag_k) >> g.sh_k);
    const unsigned k = (unsigned)n - row * g.nk;
    const long long q = f * g.field + (long long)row * g.pitch + k;
    x[f * g.n_nodes + n] = fma(alpha, pp[q], xp[q]);
  }
}

// deferred x update: x_unpadded = x_padded + the step lengths times directions not yet added (iterations
// floor(its / 15) * 15 .. its - 1, in iteration order)
struct RingPtrs {
  const double* p[kXRing];
};
__global__ void __launch_bounds__(VT) k_cg_x_finalize_ring(PadGeom g, const double* __restrict__ xp, RingPtrs pr,
                                                            const double* __restrict__ S, const double* __restrict__ xring,
                                                            long long ob, long long oe, double* __restrict__ x) {
  const int its = (int)S[S_ITS];
  const int lo = (its / (kXRing - 1)) * (kXRing - 1);
  const int f = blockIdx.y;
  for (long long n = ob + (long long)blockIdx.x * VT + threadIdx.x; n < oe; n += (long long)gridDim.x * VT) {
    const unsigned row = (unsigned)(((unsigned long long)(unsigned)n * g.mag_k) >> g.sh_k);
    const unsigned k = (unsigned)n - row * g.nk;
    const long long q = f * g.field + (long long)row * g.pitch + k;
    double v = xp[q];
    for (int kk = lo; kk < its; ++kk) {
      const int slot = (kk + 1) % kXRing;
      v = fma(xring[slot], pr.p[slot][q], v);
    }
    x[f * g.n_nodes + n] = v;
  }
}

// node id -> padded offset (row elimination list)
__global__ void k_pad_ids(long long n, const int32_t* __restrict__ nodes, PadGeom g, int32_t* __restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const unsigned node = (unsigned)nodes[t];
    const unsigned row = (unsigned)(((unsigned long long)node * g.mag_k) >> g.sh_k);
    out[t] = (int32_t)((long long)row * g.pitch + (node - row * g.nk));
  }
}

struct FixPArgs {
  const int32_t* ids[2];
  long long count[2];
  double* w;
  const double* p;
  long long field, ob, oe;
  int identity;
  const double* skip_flag;
};

__global__ void k_fix_rows_padded(const FixPArgs a) {
  if (a.skip_flag != nullptr && *a.skip_flag != 0.0) return;
  const long long total = a.count[0] + a.count[1];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int f = t < a.count[0] ? 0 : 1;
    const long long q = a.ids[f][f ? t - a.count[0] : t];
    if (q >= a.ob && q < a.oe) a.w[f * a.field + q] = a.identity ? a.p[f * a.field + q] : 0.0;
  }
}

// count[0] = constrained nodes in the interior, count[1] = unconstrained nodes on the domain boundary
__global__ void __launch_bounds__(VT) k_classify_mask(RArgs a, const uint8_t* __restrict__ mask, long long n_nodes,
                                                       unsigned nk_u, unsigned long long mag_nk, unsigned sh_nk,
                                                       unsigned long long* __restrict__ count) {
  unsigned long long c0 = 0, c1 = 0;
  for (long long n = (long long)blockIdx.x * VT + threadIdx.x; n < n_nodes; n += (long long)gridDim.x * VT) {
    const unsigned row = (unsigned)(((unsigned long long)(unsigned)n * mag_nk) >> sh_nk);
    const unsigned k = (unsigned)n - row * nk_u;
    const int cls = node_class(a, row * a.pitch + k) & 7;
    const bool m = mask[n] != 0;
    c0 += (m && cls == 0);
    c1 += (!m && cls != 0);
  }
  if (c0) atomicAdd(&count[0], c0);
  if (c1) atomicAdd(&count[1], c1);
}

// reciprocal diagonal per boundary class; same expression as k_diag_structured (apply_structured.cu)
__global__ void k_dinv_table(GridDesc g, Coef c, int nf, int jacobi, int zc0, int zc1, double* __restrict__ tab) {
  const int t = threadIdx.x;
  if (t >= kClsPerField * nf) return;
  const int f = t / kClsPerField, cls = t % kClsPerField;
  const int bnd = cls & 7, par = cls >> 3;
  // full-boundary Dirichlet field: every boundary class is a constrained row -> z = 0, p = 0 there
  if ((f ? zc1 : zc0) && bnd) { tab[t] = 0.0; return; }
  if (!jacobi) { tab[t] = 1.0; return; }
  // centre entries of the assembled 1-D matrices on a uniform axis.  Degree 1: row 0 is a domain-boundary row,
  // an interior row's centre entry is exactly twice that (two cells instead of one).  Degree 2 (band 2): row 0 =
  // boundary vertex, row 1 = mid node, an interior vertex again twice row 0.  Dummy axis: M = [1], K = [0].
  const int w = 2 * g.band + 1;
  double m[3], k[3];
  for (int a = 0; a < 3; ++a) {
    const int b = (bnd >> (2 - a)) & 1, odd = (par >> (2 - a)) & 1;
    const bool dummy = g.n[a] == 1;
    const double mb = g.m1d[a][g.band], kb = g.k1d[a][g.band];
    if (dummy) { m[a] = mb; k[a] = kb; }
    else if (g.band == 2 && odd) { m[a] = g.m1d[a][w + g.band]; k[a] = g.k1d[a][w + g.band]; }
    else { m[a] = b ? mb : 2.0 * mb; k[a] = b ? kb : 2.0 * kb; }
  }
  const double mxc = m[0], kxc = k[0], myc = m[1], kyc = k[1], mzc = m[2], kzc = k[2];
  const double K = kxc * myc * mzc + mxc * kyc * mzc + mxc * myc * kzc;
  const double M = mxc * myc * mzc;
  const double d = c.cK[f][f] * K + c.cM[f][f] * M;
  tab[t] = 1.0 / d;
}

// ---------------------------------------------------------------------------------------------------------------
// Degree 2 (BASELINE configs[3]: Q2 192^3 block Picard): the same two-kernel iteration on the padded layout.
// k_cg_fused_apply_q2 = the uniform-grid Q2 stencil of apply_structured_q2.cu (node pairs along z, 16-byte shared
// loads, constant vertex / mid rows, shifting 5-plane queue along x) with the CG prologue fused in: planes of r and
// p_old stream through a cp.async ring (16-byte copies: the padded rows are 16-byte aligned; zero-filled outside
// the domain), p = D^-1 r + beta p_old is formed in shared memory over the whole halo'd tile (the reciprocal
// diagonal comes from the 64-entry class table: node parity x boundary per axis), p is stored into the direction
// ring (ghost planes included) and w = A p, <p, w> follow from the tile.  x is always deferred (section 4.6): the
// r-update kernel of the Q1 path serves Q2 unchanged (class index extended by the parity bits).
// Passes per iteration: r, p_old in; p, w out (4) + r-update 3 + x 17/15 = 8.13, against 13 + reciprocal diagonal
// for the unfused sequence this replaces.
// ---------------------------------------------------------------------------------------------------------------
constexpr int Q2K = 32, Q2H = 2, Q2ROW = Q2K + 2 * Q2H, Q2RING = 3;
constexpr int Q2PT = 16;     // 16 pair-threads per row; TJ rows per tile -> 16 * TJ threads, (TJ + 4) x 36 halo'd tile

struct Q2FArgs {
  int n[3];
  int pitch;
  long long plane, field;
  // The assembled 1-D rows on an equally spaced axis are (h/30) x {-1, 2, 8|4, 2, -1} / {2, 16, 2} (mass, vertex / mid
  // row; 4 on the domain boundary) and 1/(3h) x {1, -8, 14|7, -8, 1} / {-8, 16, -8} (stiffness): the kernel evaluates
  // the integer stencils (immediate operands) and the scale factors ride on three coefficients per field pair:
  // y_f = sum_g a1[f][g] (Kx' c') + a2[f][g] (Mx' d') + a3[f][g] (Mx' c'),  d' = rho_y (Ky' x Mz') p + rho_z (My' x Kz') p
  double rho_y, rho_z;       // (1/(3h)) / (h/30) = 10 / h^2 of the y and z axes
  double a1[2][2], a2[2][2], a3[2][2];
  double cxM, cxK;           // centre entries of the boundary vertex rows along x: 4, 7 (dummy axis of a 2-D mesh: 1, 0)
  const double* r;
  const double* pin;
  double* pout;
  double* w;
  double* dot_partials;
  int i_begin, i_end;
  int ntj, ntk, nseg;
  int dom_lo, dom_hi;
  const double* S;
  const double* dtab;
  FoldArgs fold;
};

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr, bool valid) {
  asm volatile(
      "{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n cp.async.cg.shared.global [%0], [%1], 16, p;\n}\n" ::"r"(smem_addr),
      "l"(gptr), "r"((unsigned)valid)
      : "memory");
}

template <int NF, int TJ>
struct __align__(16) SmemQ2 {
  static constexpr int SLOT = (TJ + 2 * Q2H) * Q2ROW;
  double r[Q2RING][NF][SLOT];
  double p[Q2RING][NF][SLOT];
  double dtab[2 * kClsPerField];
  double red[Q2PT * TJ / 32];
  double fin[kFinishSmem];
  double spre[kPreScalars];
  unsigned long long seq_pre;
  int last_flag;
};

// OCC = resident CTAs per SM the register budget is cut for: 3 for two fields (168 registers), 4 for the one-field
// blocks of the Picard / fieldsplit solves (122 registers; measured at 128^3: 4 -> 1466 ms, 3 -> 1525 ms, 5 -> 1490 ms)
template <int NF, int OCC, int TJ, bool PERSIST>
__global__ void __launch_bounds__(Q2PT * TJ, OCC) k_cg_fused_apply_q2(const Q2FArgs s) {
  constexpr int Q2J = TJ, Q2NT = Q2PT * TJ, Q2SLOT = SmemQ2<NF, TJ>::SLOT;
  extern __shared__ __align__(16) unsigned char smem_raw_q2[];
  SmemQ2<NF, TJ>& sm = *reinterpret_cast<SmemQ2<NF, TJ>*>(smem_raw_q2);
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  if (tid < 2 * kClsPerField) sm.dtab[tid] = tid < kClsPerField * NF ? s.dtab[tid] : 1.0;   // per-solve constant
  pdl_wait();   // everything below reads what the previous kernel of the stream produced
  if (s.S[S_REASON] != 0.0) return;
  if (s.fold.enabled) {
    if (tid < kPreScalars) sm.spre[tid] = s.fold.S[tid];
    if (tid == kPreScalars && s.fold.ipc.world > 1) sm.seq_pre = *s.fold.ipc.seq_dev;
  }
  const double beta = (s.S[S_ITS] == 0.0) ? 0.0 : s.S[S_RZ] / s.S[S_RZ_OLD];

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2], pitch = s.pitch;
  const int ntiles = s.ntj * s.ntk;
  const int nown = s.i_end - s.i_begin;
  const int warp = tid >> 5, lane = tid & 31;
  double dot = 0.0;
  // Work = (tile, plane) steps.  nseg > 0: one (tile, x-segment) item per CTA, tile index fastest (neighbouring tiles
  // run side by side: their halo rows hit in L2).  nseg == 0 (thin slabs: few planes, CTA count not a multiple of the
  // resident slots): equal contiguous shares of the steps per CTA, processed as runs of consecutive planes of one tile.
  // (PERSIST is a template parameter: the one-run variant compiles without the loop -- it is 6 % faster that way)
  long long wbeg, wend;
  if (!PERSIST) {
    const int tile_ = blockIdx.x % ntiles, seg_ = blockIdx.x / ntiles;
    wbeg = (long long)tile_ * nown + bstart(seg_, nown, s.nseg);
    wend = (long long)tile_ * nown + bstart(seg_ + 1, nown, s.nseg);
  } else {
    const long long total = (long long)ntiles * nown;
    wbeg = (total * blockIdx.x) / gridDim.x;
    wend = (total * (blockIdx.x + 1)) / gridDim.x;
  }
#pragma unroll 1
  do {
  const int tile = PERSIST ? (int)(wbeg / nown) : (int)(blockIdx.x % ntiles);
  const int run_a = (int)(wbeg - (long long)tile * nown);
  const int run_len = (int)((wend - wbeg) < (long long)(nown - run_a) ? (wend - wbeg) : (long long)(nown - run_a));
  wbeg += run_len;
  const int tkid = tile % s.ntk, tjid = tile / s.ntk;
  const int k0 = tkid * Q2K, j0 = tjid * Q2J;            // even: node parity = local parity
  const int i_lo = s.i_begin + run_a, i_hi = i_lo + run_len;
  const int jr = warp + (TJ / 2) * (lane >> 4);        // rows {w, w + TJ / 2} of a warp have one parity
  const int kp = 2 * (lane & 15);
  const int j = j0 + jr, k = k0 + kp;
  const bool rowV = (jr & 1) == 0;
  const bool actV = (j < nj) && (k < nk), actM = (j < nj) && (k + 1 < nk);
  const long long plane = s.plane;
  // planes whose p this run stores: its own planes plus the ghost planes next to the first / last owned one
  const int pw_lo = (i_lo == s.i_begin && s.i_begin > 0) ? i_lo - Q2H : i_lo;
  const int pw_hi = (i_hi == s.i_end && s.i_end < ni) ? i_hi + 1 : i_hi;

  // copy / combine duties (fixed across planes): 16-byte pairs tid + q * Q2NT of a slot
  constexpr int NPAIR = Q2SLOT / 2, NDUTY = (NPAIR + Q2NT - 1) / Q2NT;
  long long coff[NDUTY];
  bool cok[NDUTY];
  int ccls[NDUTY][2];      // class bits of the two elements that do not depend on the plane
#pragma unroll
  for (int q = 0; q < NDUTY; ++q) {
    const int e = 2 * (tid + q * Q2NT);
    const int rr = e / Q2ROW, cc = e - rr * Q2ROW;
    const int jj = j0 - Q2H + rr, kk = k0 - Q2H + cc;
    cok[q] = (e < Q2SLOT) && (jj >= 0) && (jj < nj) && (kk >= 0) && (kk < pitch);
    coff[q] = cok[q] ? (long long)jj * pitch + kk : 0;
    const int by = (jj == 0 || jj == nj - 1) ? 2 : 0, py = (jj & 1) << 1;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int kq = kk + u;
      const int bz = (kq == 0 || kq == nk - 1) ? 1 : 0, pz = kq & 1;
      ccls[q][u] = ((py | pz) << 3) | by | bz;
    }
  }
  const unsigned sr_base = (unsigned)__cvta_generic_to_shared(&sm.r[0][0][0]);
  const unsigned sp_base = (unsigned)__cvta_generic_to_shared(&sm.p[0][0][0]);

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
  const int i_first = i_lo - Q2H;
  const long long own = (long long)j * pitch + k;

  auto issue = [&](int pl, int slot) {
    const bool in = (unsigned)pl < (unsigned)ni;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const long long base = (long long)f * s.field + (long long)(in ? pl : 0) * plane;
#pragma unroll
      for (int q = 0; q < NDUTY; ++q)
        if (2 * (tid + q * Q2NT) < Q2SLOT) {
          const unsigned off = (unsigned)(((slot * NF + f) * Q2SLOT + 2 * (tid + q * Q2NT)) * 8);
          cp_async16(sr_base + off, s.r + base + coff[q], in && cok[q]);
          cp_async16(sp_base + off, s.pin + base + coff[q], in && cok[q]);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

#define DPP_Q2F_ROW(T, MYR, KYR)                                                                      \
  {                                                                                                   \
    const double2 p0 = *reinterpret_cast<const double2*>(T);                                          \
    const double2 p1 = *reinterpret_cast<const double2*>((T) + 2);                                    \
    const double2 p2 = *reinterpret_cast<const double2*>((T) + 4);                                    \
    const double s2 = p0.x + p2.x, s1 = p0.y + p1.y, sm_ = p1.x + p2.x, e16 = 16.0 * p1.y;            \
    const double tzV = fma(2.0, s1, fma(mzVc, p1.x, -s2));                                            \
    const double uzV = fma(-8.0, s1, fma(kzVc, p1.x, s2));                                            \
    const double tzM = fma(2.0, sm_, e16);                                                            \
    const double uzM = fma(-8.0, sm_, e16);                                                           \
    cV = fma(MYR, tzV, cV);                                                                           \
    d1V = fma(KYR, tzV, d1V);                                                                         \
    d2V = fma(MYR, uzV, d2V);                                                                         \
    cM = fma(MYR, tzM, cM);                                                                           \
    d1M = fma(KYR, tzM, d1M);                                                                         \
    d2M = fma(MYR, uzM, d2M);                                                                         \
  }

  __syncthreads();   // dtab, spre; later runs: the plane ring is free (every thread finished the previous run)
  issue(i_first, 0);
  issue(i_first + 1, 1);
  int slot = 0;
  const int i_last = i_hi + Q2H - 1;
#pragma unroll 1
  for (int ip = i_first; ip <= i_last; ++ip) {
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    {
      int nslot = slot + 2;
      if (nslot >= Q2RING) nslot -= Q2RING;
      issue(ip + 2, nslot);
    }
    const bool in = (unsigned)ip < (unsigned)ni;
    // ---- p = D^-1 r + beta p_old over the halo'd tile, in place (zero-filled elements stay zero)
    if (in) {
      const int cx = ((ip & 1) << 5) | (((ni > 1) && ((ip == 0 && s.dom_lo) || (ip == ni - 1 && s.dom_hi))) ? 4 : 0);
#pragma unroll
      for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int q = 0; q < NDUTY; ++q)
          if (2 * (tid + q * Q2NT) < Q2SLOT) {
            const int e = 2 * (tid + q * Q2NT);
            double2 pv = *reinterpret_cast<double2*>(&sm.p[slot][f][e]);
            const double2 rv = *reinterpret_cast<const double2*>(&sm.r[slot][f][e]);
            pv.x = fma(beta, pv.x, sm.dtab[f * kClsPerField + (cx | ccls[q][0])] * rv.x);
            pv.y = fma(beta, pv.y, sm.dtab[f * kClsPerField + (cx | ccls[q][1])] * rv.y);
            *reinterpret_cast<double2*>(&sm.p[slot][f][e]) = pv;
          }
    }
    __syncthreads();
    const bool pwr = in && ip >= pw_lo && ip < pw_hi;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      double cV = 0.0, dV = 0.0, cM = 0.0, dM = 0.0, xcV = 0.0, xcM = 0.0;
      if (in && actV) {
        double d1V = 0.0, d2V = 0.0, d1M = 0.0, d2M = 0.0;
        const double* t = &sm.p[slot][f][jr * Q2ROW + kp];
        if (rowV) {
          DPP_Q2F_ROW(t, -1.0, 1.0)
          DPP_Q2F_ROW(t + Q2ROW, 2.0, -8.0)
          DPP_Q2F_ROW(t + 2 * Q2ROW, myVc, kyVc)
          DPP_Q2F_ROW(t + 3 * Q2ROW, 2.0, -8.0)
          DPP_Q2F_ROW(t + 4 * Q2ROW, -1.0, 1.0)
        } else {
          DPP_Q2F_ROW(t + Q2ROW, 2.0, -8.0)
          DPP_Q2F_ROW(t + 2 * Q2ROW, 16.0, 16.0)
          DPP_Q2F_ROW(t + 3 * Q2ROW, 2.0, -8.0)
        }
        dV = fma(s.rho_y, d1V, s.rho_z * d2V);
        dM = fma(s.rho_y, d1M, s.rho_z * d2M);
        xcV = t[2 * Q2ROW + 2];
        xcM = t[2 * Q2ROW + 3];
        if (pwr) {   // the direction of this iteration, for the next one and for the deferred x update
          double* po = s.pout + (long long)f * s.field + (long long)ip * plane + own;
          if (actM) *reinterpret_cast<double2*>(po) = make_double2(xcV, xcM);
          else po[0] = xcV;
        }
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
    const int io = ip - Q2H;
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
        double* wo = s.w + (long long)f * s.field + node;
        dot = fma(cenV[f][0], yV, dot);
        if (actM) {
          *reinterpret_cast<double2*>(wo) = make_double2(yV, yM);
          dot = fma(cenM[f][0], yM, dot);
        } else {
          wo[0] = yV;
        }
      }
    }
    if (++slot == Q2RING) slot = 0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  } while (PERSIST && wbeg < wend);  // runs
#undef DPP_Q2F_ROW

  if (s.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) sm.red[warp] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < Q2NT / 32; ++w) t += sm.red[w];
      s.dot_partials[blockIdx.x] = t;
    }
    if (s.fold.enabled && last_block_arrives(s.fold.counter, gridDim.x, &sm.last_flag))
      finish_reduction<Q2NT>(s.dot_partials, (int)gridDim.x, 1, s.fold.S, s.fold.hist, s.fold.post, 0, s.fold.ipc, sm.fin,
                             s.fold.ipc.ll == 3, FoldPre{sm.spre, &sm.seq_pre}, s.fold.xring);
  }
}

void magic_div(unsigned d, unsigned long long* mag, unsigned* sh) {
  // exact floor(q / d) for all q < 2^31: k = 32 + ceil(log2 d), mag = ceil(2^k / d)  
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned k = 32 + l;
  *sh = k;
  *mag = (unsigned long long)((((unsigned __int128)1 << k) + d - 1) / d);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// state of the fused path kept per handle
struct FusedState {
  int ni = 0, nj = 0, nk = 0, pitch = 0;
  long long plane = 0, field = 0;
  double* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // r, p0, p1, w, x  (2 * field doubles each)
  CUtensorMap tm[2][4];   // [nf-1][r, p0, p1, x]
  CUtensorMap tm_rx[2];   // [nf-1] r without halo (the tiles the stencil r-update reads)
  int32_t* bc_pad[2] = {nullptr, nullptr};
  long long bc_count[2] = {0, 0};
  long long bc_cap[2] = {0, 0};
  int64_t bc_gen[2] = {-1, -1};   // ctx->bc_gen the padded ids were derived from
  int bc_full[2] = {0, 0};        // Dirichlet set of the field == all domain-boundary nodes (class-mask mode)
  int64_t bc_full_gen[2] = {-1, -1};
  int bc_full_dom[2] = {-1, -1};  // dom_lo + 2*dom_hi the classification was made for
  unsigned long long* d_count = nullptr;
  bool attr_set = false;
  // deferred x update: ring of direction buffers (pring[0], pring[1] alias buf[1], buf[2]) + their tensor maps
  double* pring[kXRing] = {};
  CUtensorMap tmr[2][kXRing];
  double* d_xring = nullptr;      // [2 * kXRing]: alpha, iteration tag
  bool ring_ready = false, ring_failed = false;
  bool defer = false;             // the current solve defers the x update (decide_defer_x)
  unsigned long long ring_clean_gen = ~0ull;   // ctx->state_gen the ring buffers were last zeroed for
};

void cg_fused_destroy(dpp_context* ctx) {
  FusedState* F = ctx->fused;
  if (!F) return;
  for (double* b : F->buf)
    if (b) cudaFree(b);
  for (int32_t* b : F->bc_pad)
    if (b) cudaFree(b);
  for (int j = 2; j < kXRing; ++j)
    if (F->pring[j]) cudaFree(F->pring[j]);
  if (F->d_xring) cudaFree(F->d_xring);
  if (F->d_count) cudaFree(F->d_count);
  delete F;
  ctx->fused = nullptr;
}

static int fused_state(dpp_context* ctx, FusedState** out);
static int ring_state(dpp_context* ctx, FusedState* F);

// degree 2: single GPU, or slabs whose residual halo goes through peer memory (the r-update kernel stores its boundary
// planes into the neighbours: one plane down, two up; without peer memory the unfused sequence with its NCCL halos
// runs), and only with the direction ring (the kernel has no in-kernel x update)
static bool q2_fused_usable(dpp_context* ctx) {
  if (!(ctx->world == 1 || comm_ipc_halo_ready(ctx)) || getenv("DPP_NO_FUSED_Q2") != nullptr) return false;
  FusedState* F = nullptr;
  if (fused_state(ctx, &F) != DPP_OK) { ctx->err.clear(); cudaGetLastError(); return false; }
  return ring_state(ctx, F) == DPP_OK;
}

bool cg_fused_available(dpp_context* ctx, int nf, int operator_mode, int pc_type) {
  const bool common = ctx->family == DPP_KERNEL_STRUCTURED && !ctx->force_table_kernel && !ctx->fused_cg_disabled && operator_mode == DPP_OP_MATRIX_FREE &&
                      (pc_type == DPP_PC_NONE || pc_type == DPP_PC_JACOBI) && (nf == 1 || nf == 2) &&
                      getenv("DPP_NO_FUSED_CG") == nullptr;
  if (!common) return false;
  if (ctx->grid.band == 1) return ctx->grid_uniform && encode_fn() != nullptr;
  return ctx->grid.band == 2 && ctx->q2_uniform && q2_fused_usable(ctx);
}

static PadGeom pad_geom(const dpp_context* ctx, const FusedState* F, int nf) {
  PadGeom g{};
  g.nf = nf;
  g.n_nodes = ctx->n_nodes;
  g.field = F->field;
  g.nj = (unsigned)F->nj; g.nk = (unsigned)F->nk; g.pitch = (unsigned)F->pitch;
  magic_div(g.nk, &g.mag_k, &g.sh_k);
  return g;
}

static int make_map(dpp_context* ctx, FusedState* F, CUtensorMap* m, double* base, int nf, bool halo) {
  const cuuint64_t dims[4] = {(cuuint64_t)F->nk, (cuuint64_t)F->nj, (cuuint64_t)F->ni, (cuuint64_t)nf};
  const cuuint64_t strides[3] = {(cuuint64_t)F->pitch * 8, (cuuint64_t)F->plane * 8, (cuuint64_t)F->field * 8};
  const cuuint32_t box[4] = {(cuuint32_t)(halo ? SROW : TK), (cuuint32_t)(halo ? TJ + 2 : TJ), 1u, (cuuint32_t)nf};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult rc = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
    return DPP_ERR_CUDA;
  }
  return DPP_OK;
}

// allocate the padded vectors and tensor maps (once per handle / mesh)
static int fused_state(dpp_context* ctx, FusedState** out) {
  if (!ctx->fused) {
    FusedState* F = new FusedState();
    ctx->fused = F;
    const GridDesc& g = ctx->grid;
    F->ni = g.n[0]; F->nj = g.n[1]; F->nk = g.n[2];
    F->pitch = ((F->nk + 15) / 16) * 16;
    F->plane = (long long)F->nj * F->pitch;
    F->field = (long long)F->ni * F->plane;
    if (F->field >= (1LL << 31)) {
      ctx->set_error("fused CG: padded field exceeds int32 indexing");
      return DPP_ERR_INVALID;
    }
    for (double*& b : F->buf) {
      DPP_CHECK(dev_alloc(ctx, &b, 2 * F->field));
      DPP_CUDA(cudaMemsetAsync(b, 0, sizeof(double) * 2 * F->field, ctx->stream));
    }
    for (int nf = 1; nf <= 2 && g.band == 1; ++nf) {
      DPP_CHECK(make_map(ctx, F, &F->tm[nf - 1][0], F->buf[0], nf, true));
      DPP_CHECK(make_map(ctx, F, &F->tm[nf - 1][1], F->buf[1], nf, true));
      DPP_CHECK(make_map(ctx, F, &F->tm[nf - 1][2], F->buf[2], nf, true));
      DPP_CHECK(make_map(ctx, F, &F->tm[nf - 1][3], F->buf[4], nf, false));
      DPP_CHECK(make_map(ctx, F, &F->tm_rx[nf - 1], F->buf[0], nf, false));
    }
  }
  *out = ctx->fused;
  return DPP_OK;
}

// deferred x update is used whenever the reduction epilogues run inside the kernels (single GPU or peer-memory
// mailboxes): they are what records the step lengths.  DPP_NO_DEFER_X=1: x updated in the apply kernel.  If the
// direction ring cannot be allocated (16 vectors; 9.2 GB at 256^3) the handle silently keeps the two-buffer scheme:
// same arithmetic, two more vector passes per iteration.
static int ring_state(dpp_context* ctx, FusedState* F) {
  if (F->ring_ready) return DPP_OK;
  if (F->ring_failed) return DPP_ERR_CUDA;
  F->pring[0] = F->buf[1];
  F->pring[1] = F->buf[2];
  for (int j = 2; j < kXRing; ++j) {
    if (dev_alloc(ctx, &F->pring[j], 2 * F->field) != DPP_OK) {
      cudaGetLastError();
      for (int i = 2; i < j; ++i) {
        cudaFree(F->pring[i]);
        ctx->device_bytes -= (int64_t)sizeof(double) * 2 * F->field;
        F->pring[i] = nullptr;
      }
      F->pring[j] = nullptr;
      F->ring_failed = true;
      ctx->err.clear();
      return DPP_ERR_CUDA;
    }
    DPP_CUDA(cudaMemsetAsync(F->pring[j], 0, sizeof(double) * 2 * F->field, ctx->stream));
  }
  for (int nf = 1; nf <= 2 && ctx->grid.band == 1; ++nf)
    for (int j = 0; j < kXRing; ++j) DPP_CHECK(make_map(ctx, F, &F->tmr[nf - 1][j], F->pring[j], nf, true));
  DPP_CHECK(dev_alloc(ctx, &F->d_xring, 2 * kXRing));
  DPP_CUDA(cudaMemsetAsync(F->d_xring, 0, sizeof(double) * 2 * kXRing, ctx->stream));
  F->ring_ready = true;
  return DPP_OK;
}

// decided at the start of every solve / timing run (cg_fused_table) and read by all launches of that solve
static void decide_defer_x(dpp_context* ctx, FusedState* F) {
  const bool want = (ctx->world == 1 || comm_ipc_ready(ctx)) && (getenv("DPP_NO_DEFER_X") == nullptr || ctx->grid.band == 2);
  F->defer = want && ring_state(ctx, F) == DPP_OK;
}

// variant of the fused iteration the launches of the current solve use (part of the CUDA-graph key of a CG batch)
int cg_fused_variant(dpp_context* ctx) {
  if (ctx->fused == nullptr || !ctx->fused->defer) return 0;
  return getenv("DPP_NO_STENCIL_RUPD") != nullptr ? 1 : 2;
}

// launch with programmatic stream serialization allowed (the kernel calls griddepcontrol.wait itself)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = getenv("DPP_NO_PDL") ? 0 : 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static FoldArgs fold_args(dpp_context* ctx, int slot, int post, int counter) {
  FoldArgs f{};
  f.enabled = (ctx->world == 1 || comm_ipc_ready(ctx)) ? 1 : 0;
  f.counter = ctx->d_counters + counter;
  f.S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  f.hist = hist_device(ctx, slot);
  f.post = post;
  f.ipc = comm_ipc_reduce_args(ctx);
  return f;
}

// mode: 0 plain apply, 1 fused CG iteration kernel, 2 residual update with the stencil in it (see the kernel)
static int launch_apply(dpp_context* ctx, FusedState* F, int nf, int mode, const Coef& c, const CUtensorMap& tm_pin,
                        double* pout, int slot, const double* dtab, bool want_dot, int* n_partial_blocks,
                        bool interior_only = false, bool defer_x = false, bool no_w = false) {
  const bool fused = mode != 0;
  const GridDesc& g = ctx->grid;
  const long long uplane = (long long)g.n[1] * g.n[2];
  if (ctx->owned_begin % uplane || ctx->owned_end % uplane) {
    ctx->set_error("fused CG: owned range must consist of whole x-planes");
    return DPP_ERR_INVALID;
  }
  FArgs s{};
  for (int d = 0; d < 3; ++d) {
    s.n[d] = g.n[d]; s.m1d[d] = g.m1d[d]; s.k1d[d] = g.k1d[d];
    s.mo[d] = ctx->uni_m_off[d]; s.ko[d] = ctx->uni_k_off[d];
  }
  s.pitch = F->pitch; s.plane = F->plane; s.field = F->field;
  s.mxc_i = ctx->uni_mxc[0]; s.mxc_b = ctx->uni_mxc[1];
  s.kxc_i = ctx->uni_kxc[0]; s.kxc_b = ctx->uni_kxc[1];
  s.pout = pout; s.x = F->buf[4]; s.w = F->buf[3];
  s.c = c;
  s.dot_partials = want_dot ? ctx->d_partials : nullptr;
  s.i_begin = (int)(ctx->owned_begin / uplane);
  s.i_end = (int)(ctx->owned_end / uplane);
  s.S = fused ? ctx->d_scalars + (size_t)slot * S_SLOT_SIZE : nullptr;
  s.dtab = dtab;
  s.dom_lo = ctx->dom_lo; s.dom_hi = ctx->dom_hi;
  if (mode == 1) s.fold = fold_args(ctx, slot, POST_CG_PAP, 0);
  s.defer_x = defer_x ? 1 : 0;
  if (defer_x && mode == 1) s.fold.xring = F->d_xring;
  s.no_w = no_w ? 1 : 0;
  s.own_first = s.own_last = -1;
  if (mode == 2) {
    s.fold = fold_args(ctx, slot, POST_CG_RZ, 1);
    s.r = F->buf[0];
    s.partials2 = ctx->d_partials;
    s.halo = comm_ipc_halo(ctx);
    if (s.halo.peer_r[0] != nullptr) s.own_first = s.i_begin;
    if (s.halo.peer_r[1] != nullptr) s.own_last = s.i_end - 1;
    s.ob = (long long)s.i_begin * F->plane;
    s.oe = (long long)s.i_end * F->plane;
    for (int j = 0; j < kXRing; ++j) s.pr[j] = F->pring[j];
    s.xring = F->d_xring;
  }
  s.j_lo = 0; s.j_hi = g.n[1]; s.k_lo = 0; s.k_hi = g.n[2];
  if (fused && interior_only && g.n[1] >= 3 && g.n[2] >= 3 && (g.n[0] >= 3 || g.n[0] == 1)) {
    // class-mask mode: every domain-boundary node is a constrained row whose p, w, x stay zero and whose
    // w is never read -- compute the interior only (at 257^2 planes: 16 x 8 = 128 tiles instead of 153)
    s.j_lo = 1; s.j_hi = g.n[1] - 1; s.k_lo = 1; s.k_hi = g.n[2] - 1;
    if (g.n[0] > 1) {  // domain-boundary planes too, as long as the rank keeps at least one plane to compute
      if (ctx->dom_lo && s.i_begin == 0 && s.i_end - s.i_begin > 1) s.i_begin = 1;
      if (ctx->dom_hi && s.i_end == g.n[0] && s.i_end - s.i_begin > 1) s.i_end = g.n[0] - 1;
    }
  }
  s.ntk = (((s.k_hi + 1) >> 1) - (s.k_lo >> 1) + TK / 2 - 1) / (TK / 2);   // tiles hold <= TK/2 column pairs
  s.ntj = (s.j_hi - s.j_lo + TJ - 1) / TJ;
  if (const char* e = getenv("DPP_FUSED_TILES")) {   // measurement override "ntj,ntk" (>= the minimal counts)
    int tj = 0, tk = 0;
    if (sscanf(e, "%d,%d", &tj, &tk) == 2) {
      s.ntj = std::max(s.ntj, tj);
      s.ntk = std::max(s.ntk, tk);
    }
  }
  const int tiles = s.ntk * s.ntj;
  const int nown = s.i_end - s.i_begin;
  if (nown <= 0) { *n_partial_blocks = 0; return DPP_OK; }
  // (tile, x-segment) items in hardware dispatch order; see choose_x_segments()
  const long long total = (long long)tiles * nown;
  int nctas;
  if (tiles <= kMaxPartialBlocks / 2) {
    s.nseg = choose_x_segments(tiles, nown, ctx->sm_count * 2, kMaxPartialBlocks);
    nctas = tiles * s.nseg;
  } else {  // very wide planes: persistent partition (partials scratch is bounded)
    s.nseg = 0;
    nctas = (int)std::max<long long>(1, std::min<long long>(ctx->sm_count * 2, total / 8));
  }
  if (const char* e = getenv("DPP_FUSED_SCHED")) {  // measurement override: "p" persistent, "<n>" segments
    if (e[0] == 'p') {
      s.nseg = 0;
      nctas = (int)std::max<long long>(1, std::min<long long>(ctx->sm_count * 2, total / 4));
    } else if (atoi(e) > 0 && (long long)tiles * atoi(e) <= kMaxPartialBlocks) {
      s.nseg = std::min(atoi(e), nown);
      nctas = tiles * s.nseg;
    }
  }
  if (!F->attr_set) {
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<2>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<1>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<2>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<1>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<2>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<1>)));
    F->attr_set = true;
  }
  dim3 grid(nctas), block(TK, TY);
  const CUtensorMap* T = F->tm[nf - 1];
  const CUtensorMap& tm_rx = F->tm_rx[nf - 1];
  if (nf == 2) {
    if (mode == 2) DPP_CUDA(launch_pdl(k_cg_fused_apply<2, 2>, grid, block, sizeof(Smem<2>), ctx->stream, s, T[0], tm_pin, tm_rx));
    else if (mode == 1) DPP_CUDA(launch_pdl(k_cg_fused_apply<2, 1>, grid, block, sizeof(Smem<2>), ctx->stream, s, T[0], tm_pin, T[3]));
    else k_cg_fused_apply<2, 0><<<grid, block, sizeof(Smem<2>), ctx->stream>>>(s, T[0], tm_pin, T[3]);
  } else {
    if (mode == 2) DPP_CUDA(launch_pdl(k_cg_fused_apply<1, 2>, grid, block, sizeof(Smem<1>), ctx->stream, s, T[0], tm_pin, tm_rx));
    else if (mode == 1) DPP_CUDA(launch_pdl(k_cg_fused_apply<1, 1>, grid, block, sizeof(Smem<1>), ctx->stream, s, T[0], tm_pin, T[3]));
    else k_cg_fused_apply<1, 0><<<grid, block, sizeof(Smem<1>), ctx->stream>>>(s, T[0], tm_pin, T[3]);
  }
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  *n_partial_blocks = nctas;
  return DPP_OK;
}

// degree 2: iteration kernel on the padded layout (always with the direction ring)
static int launch_apply_q2(dpp_context* ctx, FusedState* F, int nf, const Coef& c, const double* pin, double* pout,
                           int slot, const double* dtab, int* n_partial_blocks) {
  const GridDesc& g = ctx->grid;
  const long long uplane = (long long)g.n[1] * g.n[2];
  if (ctx->owned_begin % uplane || ctx->owned_end % uplane) {
    ctx->set_error("fused CG: owned range must consist of whole x-planes");
    return DPP_ERR_INVALID;
  }
  Q2FArgs s{};
  for (int d = 0; d < 3; ++d) s.n[d] = g.n[d];
  {
    // scale factors of the 1-D rows: mass h/30, stiffness 1/(3h); dummy axis of a 2-D mesh: M = [1], K = [0]
    const bool dummy = g.n[0] == 1;
    const double hx = ctx->uni_h[0], hy = ctx->uni_h[1], hz = ctx->uni_h[2];
    const double bmx = dummy ? 1.0 : hx / 30.0, bkx = dummy ? 0.0 : 1.0 / (3.0 * hx);
    const double B = (hy / 30.0) * (hz / 30.0);
    s.rho_y = 10.0 / (hy * hy);
    s.rho_z = 10.0 / (hz * hz);
    s.cxM = dummy ? 1.0 : 4.0;
    s.cxK = dummy ? 0.0 : 7.0;
    for (int f = 0; f < 2; ++f)
      for (int q = 0; q < 2; ++q) {
        s.a1[f][q] = c.cK[f][q] * bkx * B;
        s.a2[f][q] = c.cK[f][q] * bmx * B;
        s.a3[f][q] = c.cM[f][q] * bmx * B;
      }
  }
  s.pitch = F->pitch; s.plane = F->plane; s.field = F->field;
  s.r = F->buf[0]; s.pin = pin; s.pout = pout; s.w = F->buf[3];
  s.dot_partials = ctx->d_partials;
  s.i_begin = (int)(ctx->owned_begin / uplane);
  s.i_end = (int)(ctx->owned_end / uplane);
  s.S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  s.dtab = dtab;
  s.dom_lo = ctx->dom_lo; s.dom_hi = ctx->dom_hi;
  s.fold = fold_args(ctx, slot, POST_CG_PAP, 0);
  s.fold.xring = F->d_xring;
  // 8-row tiles, 128 threads; 3 CTAs per SM for two fields (168 registers), 4 for one.  16-row tiles (256 threads,
  // halo'd tile 1.41x instead of 1.69x the output tile) measured the same within noise (config 4 at 128^3: 1516 / 1573
  // against 1523 / 1495 ms): the stencil phase, not the combine, carries the kernel.
  constexpr int TJ1 = 8, OCC1 = 4;
  const int tj = nf == 2 ? 8 : TJ1;
  s.ntk = (g.n[2] + Q2K - 1) / Q2K;
  s.ntj = (g.n[1] + tj - 1) / tj;
  const int tiles = s.ntk * s.ntj;
  const int nown = s.i_end - s.i_begin;
  if (nown <= 0) { *n_partial_blocks = 0; return DPP_OK; }
  if (tiles > kMaxPartialBlocks) {
    ctx->set_error("fused CG (degree 2): plane too wide for the partials scratch");
    return DPP_ERR_INVALID;
  }
  const int capacity = ctx->sm_count * (nf == 1 ? OCC1 : 3);
  s.nseg = choose_x_segments(tiles, nown, capacity, kMaxPartialBlocks, 2 * Q2H);
  {
    // Thin slabs (N > 1: a few dozen planes per rank): tiles x segments is a small non-integer number of waves of the
    // resident slots and every segment pays 4 redundant planes.  Equal shares of the (tile, plane) steps per resident
    // CTA fill the last wave; taken when the step model says >= 10 % fewer plane steps and neighbouring tiles still
    // run within a few planes of each other (their halo rows then hit in L2: the in-flight footprint between the two
    // visits, offset x resident CTAs x tile bytes, stays well inside the 126 MB).
    const int len = (nown + s.nseg - 1) / s.nseg;
    const long long rounds = ((long long)tiles * s.nseg + capacity - 1) / capacity;
    const double cost_seg = (double)rounds * (len + 2 * Q2H + 2 + (len > 32 ? (len - 32) / 4 : 0));
    const double spc = std::ceil((double)tiles * nown / capacity);
    const double cost_per = spc + (1.0 + spc / nown) * (2 * Q2H + 2);
    const double off = std::fabs(nown - spc * std::max(1.0, std::floor(nown / spc + 0.5)));
    const double between = off * capacity * ((8.0 + 2 * Q2H) * Q2ROW * 16.0 * nf);
    if (cost_per < 0.9 * cost_seg && between < 32e6 && (long long)tiles * nown >= 4LL * capacity) s.nseg = 0;
  }
  if (const char* e = getenv("DPP_FUSED_SCHED")) {   // measurement override: "p" persistent, "<n>" segments
    if (e[0] == 'p') s.nseg = 0;
    else if (atoi(e) > 0 && (long long)tiles * atoi(e) <= kMaxPartialBlocks) s.nseg = std::min(atoi(e), nown);
  }
  static bool attr = false;
  if (!attr) {
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply_q2<2, 3, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemQ2<2, 8>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply_q2<2, 3, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemQ2<2, 8>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply_q2<1, OCC1, TJ1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemQ2<1, TJ1>)));
    DPP_CUDA(cudaFuncSetAttribute(k_cg_fused_apply_q2<1, OCC1, TJ1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemQ2<1, TJ1>)));
    attr = true;
  }
  const dim3 grid(s.nseg > 0 ? tiles * s.nseg : (unsigned)std::min<long long>(capacity, (long long)tiles * nown)), block(Q2PT * tj);
  if (nf == 2 && s.nseg > 0) DPP_CUDA(launch_pdl(k_cg_fused_apply_q2<2, 3, 8, false>, grid, block, sizeof(SmemQ2<2, 8>), ctx->stream, s));
  else if (nf == 2) DPP_CUDA(launch_pdl(k_cg_fused_apply_q2<2, 3, 8, true>, grid, block, sizeof(SmemQ2<2, 8>), ctx->stream, s));
  else if (s.nseg > 0) DPP_CUDA(launch_pdl(k_cg_fused_apply_q2<1, OCC1, TJ1, false>, grid, block, sizeof(SmemQ2<1, TJ1>), ctx->stream, s));
  else DPP_CUDA(launch_pdl(k_cg_fused_apply_q2<1, OCC1, TJ1, true>, grid, block, sizeof(SmemQ2<1, TJ1>), ctx->stream, s));
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  *n_partial_blocks = (int)grid.x;
  return DPP_OK;
}

// is the Dirichlet set of mask field `fl` exactly the set of domain-boundary nodes?  (decided once per BC
// change with a device pass over the mask; then boundary classes double as the row/column mask)
static int classify_bcs(dpp_context* ctx, FusedState* F, int fl, const RArgs& geom) {
  const int dom = ctx->dom_lo + 2 * ctx->dom_hi;
  if (F->bc_full_gen[fl] == ctx->bc_gen[fl] && F->bc_full_dom[fl] == dom) return DPP_OK;
  if (!F->d_count) DPP_CHECK(dev_alloc(ctx, &F->d_count, 2));
  DPP_CUDA(cudaMemsetAsync(F->d_count, 0, 2 * sizeof(unsigned long long), ctx->stream));
  unsigned long long mag; unsigned sh;
  magic_div((unsigned)F->nk, &mag, &sh);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((ctx->n_nodes + VT - 1) / VT, (long long)ctx->sm_count * 8));
  k_classify_mask<<<blocks, VT, 0, ctx->stream>>>(geom, ctx->d_mask + (size_t)fl * ctx->n_nodes, ctx->n_nodes,
                                                  (unsigned)F->nk, mag, sh, F->d_count);
  ctx->launches++;
  unsigned long long h[2] = {1, 1};
  DPP_CUDA(cudaMemcpyAsync(h, F->d_count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  F->bc_full[fl] = (h[0] == 0 && h[1] == 0 && getenv("DPP_NO_CLASS_MASK") == nullptr) ? 1 : 0;
  F->bc_full_gen[fl] = ctx->bc_gen[fl];
  F->bc_full_dom[fl] = dom;
  return DPP_OK;
}

static int make_rargs(dpp_context* ctx, FusedState* F, int nf, int slot, const double* dtab, RArgs* out,
                      const int* fld = nullptr, int post = POST_NONE) {
  const GridDesc& g = ctx->grid;
  const long long uplane = (long long)g.n[1] * g.n[2];
  RArgs a{};
  a.nf = nf;
  a.field = F->field;
  a.ob = (ctx->owned_begin / uplane) * F->plane;
  a.oe = (ctx->owned_end / uplane) * F->plane;
  a.r = F->buf[0];
  a.w = F->buf[3];
  a.S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  a.dtab = dtab;
  a.partials = ctx->d_partials;
  a.ni = (unsigned)g.n[0]; a.nj = (unsigned)g.n[1]; a.nk = (unsigned)g.n[2]; a.pitch = (unsigned)F->pitch;
  magic_div(a.pitch, &a.mag_k, &a.sh_k);
  magic_div(a.nj, &a.mag_j, &a.sh_j);
  a.dom_lo = ctx->dom_lo;
  a.dom_hi = ctx->dom_hi;
  a.q2 = ctx->grid.band == 2 ? 1 : 0;
  a.plane = F->plane;
  a.up_win = (long long)ctx->grid.band * F->plane;
  a.halo = comm_ipc_halo(ctx);
  if (fld != nullptr) {
    for (int f = 0; f < nf; ++f) {
      DPP_CHECK(classify_bcs(ctx, F, fld[f], a));
      a.zero_class[f] = F->bc_full[fld[f]];
    }
    a.fold = fold_args(ctx, slot, post, 1);
    if (post == POST_CG_RZ && F->defer) {
      a.defer_x = 1;
      a.x = F->buf[4];
      for (int j = 0; j < kXRing; ++j) a.pr[j] = F->pring[j];
      a.xring = F->d_xring;
    }
  }
  *out = a;
  return DPP_OK;
}

static int r_blocks(const dpp_context* ctx, const RArgs& a) {
  // one wave of resident blocks (occupancy from the driver: 3 blocks of 256 threads per SM at 78 registers),
  // at least two loop trips of UNROLL pairs per thread; DPP_RUPD_WAVES: measurement override
  static int occ = 0;
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cg_r_update<false>, VT, 0) != cudaSuccess || occ < 1) occ = 3;
  }
  int waves = 1;
  if (const char* e = getenv("DPP_RUPD_WAVES")) waves = std::max(1, atoi(e));
  const long long nown = a.oe - a.ob;
  const long long trip = 2LL * VT * UNROLL;
  long long want = (nown + 2 * trip - 1) / (2 * trip);
  const long long cap = std::max(1, (ctx->sm_count * occ * waves) / std::max(1, a.nf));
  return (int)std::max(1LL, std::min(std::min(want, cap), (long long)kMaxPartialBlocks / 2));
}

int cg_fused_table(dpp_context* ctx, const Coef& c, int nf, int pc_type, const int* fld, double* d_tab) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  decide_defer_x(ctx, F);
  RArgs geom{};
  DPP_CHECK(make_rargs(ctx, F, nf, 0, d_tab, &geom, fld));
  k_dinv_table<<<1, 2 * kClsPerField, 0, ctx->stream>>>(ctx->grid, c, nf, pc_type == DPP_PC_JACOBI ? 1 : 0, geom.zero_class[0],
                                          nf == 2 ? geom.zero_class[1] : 0, d_tab);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

// start of a solve: r_padded = b, p = 0, x = 0; partial <r,z>, <z,z>
int cg_fused_begin(dpp_context* ctx, int nf, const double* b) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  const PadGeom g = pad_geom(ctx, F, nf);
  for (int v : {1, 2, 4})
    for (int f = 0; f < nf; ++f)
      DPP_CUDA(cudaMemsetAsync(F->buf[v] + f * F->field, 0, sizeof(double) * F->field, ctx->stream));
  if (F->defer) {   // ring buffer 0 is p_old of the first iteration (beta = 0 times it: must be finite)
    DPP_CUDA(cudaMemsetAsync(F->d_xring, 0, sizeof(double) * 2 * kXRing, ctx->stream));
    // rows a solve never writes (Dirichlet rows in class-mask mode, ghost planes) must read as zero: after a
    // change of parameters / BCs / partition the buffers of earlier solves are wiped once
    if (F->ring_clean_gen != ctx->state_gen) {
      for (int j = 2; j < kXRing; ++j)
        DPP_CUDA(cudaMemsetAsync(F->pring[j], 0, sizeof(double) * 2 * F->field, ctx->stream));
      F->ring_clean_gen = ctx->state_gen;
    }
  }
  dim3 grid((unsigned)std::min<long long>((ctx->n_nodes + VT - 1) / VT, (long long)ctx->sm_count * 16), nf);
  k_pad_copy<<<grid, VT, 0, ctx->stream>>>(g, b, F->buf[0]);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

// The residual update recomputes w = A p (k_cg_fused_apply<NF, 2>) instead of reading a stored w when: degree 1, the
// direction ring is in use, every field of the solve is in class-mask mode (Dirichlet set = domain boundary: the
// computed nodes are exactly the free nodes) and the rank keeps interior planes to compute; slabs need the
// peer-memory halo (the kernel pushes its boundary planes like k_cg_r_update).  DPP_NO_STENCIL_RUPD=1: stored w.
static bool stencil_rupd(dpp_context* ctx, FusedState* F, int nf, const int* fld) {
  const GridDesc& g = ctx->grid;
  if (g.band != 1 || !F->defer || getenv("DPP_NO_STENCIL_RUPD") != nullptr) return false;
  if (!(ctx->world == 1 || comm_ipc_halo_ready(ctx))) return false;
  for (int f = 0; f < nf; ++f)
    if (!(F->bc_full_gen[fld[f]] == ctx->bc_gen[fld[f]] && F->bc_full[fld[f]])) return false;
  if (!(g.n[1] >= 3 && g.n[2] >= 3 && (g.n[0] >= 3 || g.n[0] == 1))) return false;
  if (g.n[0] > 1) {
    const long long uplane = (long long)g.n[1] * g.n[2];
    int ib = (int)(ctx->owned_begin / uplane), ie = (int)(ctx->owned_end / uplane);
    if (ie - ib < 3) return false;   // (the trimming rule of launch_apply then always removes the boundary planes)
  }
  return true;
}

// <r,z>, <z,z> of the initial residual + POST_CG_INIT
int cg_fused_rz_init(dpp_context* ctx, int nf, const int* fld, int slot, const double* dtab) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  RArgs a{};
  DPP_CHECK(make_rargs(ctx, F, nf, slot, dtab, &a, fld, POST_CG_INIT));
  dim3 grid(r_blocks(ctx, a), nf);
  k_cg_r_update<true><<<grid, VT, 0, ctx->stream>>>(a);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (!a.fold.enabled) DPP_CHECK(reduce_partials(ctx, grid.x * grid.y, 2, slot, POST_CG_INIT));
  return DPP_OK;
}

// r -= alpha w (+ halo push), <r,z>, <z,z> + POST_CG_RZ
int cg_fused_r_update(dpp_context* ctx, int nf, const Coef& c, long long it, const int* fld, int slot, const double* dtab) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  if (stencil_rupd(ctx, F, nf, fld)) {   // iteration `it`: its direction p_it sits in ring buffer (it + 1) % 16
    int nb = 0;
    return launch_apply(ctx, F, nf, 2, c, F->tmr[nf - 1][(it + 1) % kXRing], nullptr, slot, dtab, false, &nb, true, true);
  }
  RArgs a{};
  DPP_CHECK(make_rargs(ctx, F, nf, slot, dtab, &a, fld, POST_CG_RZ));
  dim3 grid(r_blocks(ctx, a), nf);
  DPP_CUDA(launch_pdl(k_cg_r_update<false>, grid, dim3(VT), 0, ctx->stream, a));
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (!a.fold.enabled) DPP_CHECK(reduce_partials(ctx, grid.x * grid.y, 2, slot, POST_CG_RZ));
  return DPP_OK;
}

// iteration `it` (0-based count of applies so far): w = A p, p = dinv.*r + beta*p_prev, x += alpha_prev p_prev
int cg_fused_apply(dpp_context* ctx, int nf, const Coef& c, long long it, const int* fld, int slot, const double* dtab) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  const bool defer = F->defer;
  // classic: p ping-pongs between two buffers; deferred x: iteration `it` reads ring buffer it % 16 (p_{it-1}) and
  // writes (it + 1) % 16 (p_it), so the last fifteen directions are still there when x is brought up to date
  double* pout_buf = defer ? F->pring[(it + 1) % kXRing] : F->buf[1 + (int)((it + 1) & 1)];
  bool all_class_masked = true;
  for (int f = 0; f < nf; ++f) all_class_masked = all_class_masked && F->bc_full_gen[fld[f]] == ctx->bc_gen[fld[f]] && F->bc_full[fld[f]];
  int nb = 0;
  if (ctx->grid.band == 2) {
    if (!defer) {
      ctx->set_error("fused CG (degree 2) needs the direction ring");
      return DPP_ERR_INVALID;
    }
    DPP_CHECK(launch_apply_q2(ctx, F, nf, c, F->pring[it % kXRing], pout_buf, slot, dtab, &nb));
  } else {
    const CUtensorMap& tm_pin = defer ? F->tmr[nf - 1][it % kXRing] : F->tm[nf - 1][1 + (int)(it & 1)];
    DPP_CHECK(launch_apply(ctx, F, nf, 1, c, tm_pin, pout_buf, slot, dtab, true, &nb, all_class_masked, defer,
                           stencil_rupd(ctx, F, nf, fld)));
  }
  if (!(ctx->world == 1 || comm_ipc_ready(ctx))) DPP_CHECK(reduce_partials(ctx, nb, 1, slot, POST_CG_PAP));
  if (all_class_masked) return DPP_OK;   // constrained rows never leave zero: no row fix-up needed
  // row elimination: w = p on constrained rows (identity rows of A_bc)
  const PadGeom g = pad_geom(ctx, F, nf);
  FixPArgs fx{};
  for (int f = 0; f < nf; ++f) {
    const int fl = fld[f];
    if (F->bc_gen[fl] != ctx->bc_gen[fl]) {
      if (ctx->n_bc[fl] > F->bc_cap[fl]) {
        if (F->bc_pad[fl]) cudaFree(F->bc_pad[fl]);
        F->bc_pad[fl] = nullptr;
        F->bc_cap[fl] = 0;
      }
      F->bc_count[fl] = ctx->n_bc[fl];
      F->bc_gen[fl] = ctx->bc_gen[fl];
      if (ctx->n_bc[fl] > 0) {
        if (!F->bc_pad[fl]) {
          DPP_CHECK(dev_alloc(ctx, &F->bc_pad[fl], ctx->n_bc[fl]));
          F->bc_cap[fl] = ctx->n_bc[fl];
        }
        const int blocks = (int)std::min<long long>((ctx->n_bc[fl] + 255) / 256, 4096);
        k_pad_ids<<<blocks, 256, 0, ctx->stream>>>(ctx->n_bc[fl], ctx->d_bc_nodes[fl], g, F->bc_pad[fl]);
        ctx->launches++;
      }
    }
    fx.ids[f] = F->bc_pad[fl];
    fx.count[f] = F->bc_count[fl];
  }
  const long long total = fx.count[0] + fx.count[1];
  if (total > 0) {
    const long long uplane = (long long)ctx->grid.n[1] * ctx->grid.n[2];
    fx.w = F->buf[3];
    fx.p = pout_buf;
    fx.field = F->field;
    fx.ob = (ctx->owned_begin / uplane) * F->plane;
    fx.oe = (ctx->owned_end / uplane) * F->plane;
    fx.identity = 1;
    fx.skip_flag = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE + S_REASON;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8));
    k_fix_rows_padded<<<blocks, 256, 0, ctx->stream>>>(fx);
    ctx->launches++;
  }
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

// x (unpadded, owned rows) = x_padded + pending alpha * p, with p the direction of the last executed apply
int cg_fused_x_finalize(dpp_context* ctx, int nf, long long its, int slot, double* x) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  const PadGeom g = pad_geom(ctx, F, nf);
  const long long nown = ctx->owned_end - ctx->owned_begin;
  dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((nown + VT - 1) / VT, (long long)ctx->sm_count * 16)), nf);
  if (F->defer) {
    RingPtrs pr{};
    for (int j = 0; j < kXRing; ++j) pr.p[j] = F->pring[j];
    k_cg_x_finalize_ring<<<grid, VT, 0, ctx->stream>>>(g, F->buf[4], pr, ctx->d_scalars + (size_t)slot * S_SLOT_SIZE,
                                                       F->d_xring, ctx->owned_begin, ctx->owned_end, x);
  } else {
    k_cg_x_finalize<<<grid, VT, 0, ctx->stream>>>(g, F->buf[4], F->buf[1 + (int)(its & 1)],
                                                  ctx->d_scalars + (size_t)slot * S_SLOT_SIZE, ctx->owned_begin,
                                                  ctx->owned_end, x);
  }
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

double* cg_fused_r_buffer(dpp_context* ctx, long long* field, long long* plane) {
  const bool q1 = ctx->grid.band == 1 && ctx->grid_uniform && encode_fn() != nullptr;
  const bool q2 = ctx->grid.band == 2 && ctx->q2_uniform;
  if (!(ctx->family == DPP_KERNEL_STRUCTURED && (q1 || q2))) return nullptr;
  FusedState* F = nullptr;
  if (fused_state(ctx, &F) != DPP_OK) return nullptr;
  *field = F->field;
  *plane = F->plane;
  return F->buf[0];
}

// ghost planes of r (padded layout) for slab-partitioned runs.  `after_update`: the r-update kernel has
// already pushed them (peer-memory path); otherwise (start of a solve) exchange explicitly.
int cg_fused_halo_r(dpp_context* ctx, int nf, bool after_update, int slot) {
  if (ctx->world <= 1) return DPP_OK;
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  if (comm_ipc_halo_ready(ctx)) {
    if (after_update) return DPP_OK;
    // every rank must have finished writing its own r (ghost rows included) before neighbours store into
    // it: a zero-width mailbox reduction is the barrier; the push is then fenced by the next reduction
    DPP_CHECK(reduce_partials(ctx, 0, 1, slot, POST_NONE, 30));
    RArgs a{};
    DPP_CHECK(make_rargs(ctx, F, nf, slot, ctx->d_dtab, &a));
    dim3 grid((unsigned)std::min<long long>((F->plane + a.up_win + VT - 1) / VT, (long long)ctx->sm_count * 4), nf);
    k_push_planes<<<grid, VT, 0, ctx->stream>>>(a);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
    return DPP_OK;
  }
  const long long uplane = (long long)ctx->grid.n[1] * ctx->grid.n[2];
  return comm_halo_planes(ctx, F->buf[0], nf, F->field, F->plane, (int)(ctx->owned_begin / uplane),
                          (int)(ctx->owned_end / uplane));
}

// measurement: fill the padded work vectors with pseudo-random data (zero on constrained rows and pads)
int cg_fused_plain_apply(dpp_context* ctx, int nf, const Coef& c, bool want_dot, int* n_partial_blocks) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  return launch_apply(ctx, F, nf, 0, c, F->tm[nf - 1][1], nullptr, 0, ctx->d_dtab, want_dot, n_partial_blocks);
}

double* cg_fused_buffer(dpp_context* ctx, int which) {
  FusedState* F = nullptr;
  if (fused_state(ctx, &F) != DPP_OK) return nullptr;
  return F->buf[which];
}

int cg_fused_pad_from(dpp_context* ctx, int which, const double* src) {
  FusedState* F = nullptr;
  DPP_CHECK(fused_state(ctx, &F));
  const PadGeom g = pad_geom(ctx, F, 2);
  dim3 grid((unsigned)std::min<long long>((ctx->n_nodes + VT - 1) / VT, (long long)ctx->sm_count * 16), 2);
  k_pad_copy<<<grid, VT, 0, ctx->stream>>>(g, src, F->buf[which]);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

}  // namespace dpp
