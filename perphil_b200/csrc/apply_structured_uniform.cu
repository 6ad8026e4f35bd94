// Matrix-free DPP operator, uniform tensor grids (equal spacing per axis): the fast member of the
// DPP_KERNEL_STRUCTURED family and the kernel every BASELINE.json configuration runs.
//
// Same mathematics as apply_structured.cu (K = Kx(x)My(x)Mz + Mx(x)Ky(x)Mz + Mx(x)My(x)Kz,
// M = Mx(x)My(x)Mz) but it exploits that on a uniform axis all off-diagonal entries of the 1-D
// matrices are equal and only the centre entry changes at the two domain-boundary nodes.  With
// zero padding outside the domain the in-plane 9-point parts collapse to four neighbour-class sums
//     corner = x[j-1][k-1]+x[j-1][k+1]+x[j+1][k-1]+x[j+1][k+1],  edgeK = x[j][k-1]+x[j][k+1],
//     edgeJ  = x[j-1][k]+x[j+1][k],                              centre = x[j][k]
// shared by the mass-like part c and the stiffness-like part d, and the x-direction sweep to
// off*(q[i-1]+q[i+1]) + centre_i*q[i]: ~50 fp64 issue slots per node for the two-field operator.
//
// Organisation (driven by the ncu captures under profiles/):
//  * the input must already be zero on eliminated (Dirichlet) columns -- true for every Krylov
//    vector, arranged by a pre-mask pass for arbitrary input -- so plane tiles go global->shared
//    with cp.async (LDGSTS, zero-fill outside the domain) into a 3-slot ring: two planes in flight
//    per CTA while one is computed, one __syncthreads per plane, no staging registers;
//  * each thread owns two vertically adjacent nodes (rows j, j+1): 12 instead of 18 shared loads
//    per field, and pointer/loop/barrier overhead amortised over two nodes;
//  * ring slot and x-direction queue rotate together through a 3x unrolled loop; all global
//    addresses are running pointers;
//  * row elimination is NOT done here (a mask byte per node on the critical path costs a DRAM
//    latency per plane): a tiny follow-up kernel rewrites the constrained rows from the node list.
#include <algorithm>
#include <cstdlib>

#include "dpp_internal.cuh"

namespace dpp {

namespace {

constexpr int TK = 32;          // tile width (k), one lane per column
constexpr int TY = 8;           // thread rows
constexpr int TJ = 2 * TY;      // tile height (j): two rows per thread
constexpr int NT = TK * TY;
constexpr int SROW = TK + 2;
constexpr int SLOT = (TJ + 2) * SROW;  // doubles per field per ring slot
constexpr int HALO = 2 * SROW + 2 * TJ;
constexpr int RING = 3;

struct UArgs {
  int n[3];
  const double* m1d[3];
  const double* k1d[3];
  double mo[3], ko[3];      // uniform off-diagonal entries of the 1-D mass / stiffness matrices
  double mxc_i, mxc_b, kxc_i, kxc_b;  // axis-0 centre entries: interior / domain-boundary plane
  const double* x[2];       // stencil input, zero on eliminated columns
  double* y[2];
  Coef c;
  double* dot_partials;
  int i_begin, i_end;
  int ntj, ntk, nseg;
  const double* skip_flag;
};

__device__ __forceinline__ int bstart(int t, int n, int nt) { return (int)(((long long)t * n) / nt); }

// 8-byte global->shared async copy; when `valid` is false the source is ignored and zeros are written
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
__global__ void __launch_bounds__(NT, 2) k_apply_uniform(const UArgs s) {
  if (s.skip_flag != nullptr && *s.skip_flag != 0.0) return;
  __shared__ __align__(16) double xs[RING][NF][SLOT];
  __shared__ double red[TY];

  const int ni = s.n[0], nj = s.n[1], nk = s.n[2];
  const int nown = s.i_end - s.i_begin;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TK + tx;
  double dot = 0.0;
  // Balanced persistent partition: the (tile, plane) steps of the whole grid are cut into gridDim.x equal
  // contiguous ranges (one CTA each, a single wave of 2 CTAs per SM); a range is processed as runs of
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
  const int k0 = bstart(tkid, nk, s.ntk), k1 = bstart(tkid + 1, nk, s.ntk);
  const int j0 = bstart(tjid, nj, s.ntj), j1 = bstart(tjid + 1, nj, s.ntj);
  const int jA = j0 + 2 * ty, jB = jA + 1, k = k0 + tx;
  const bool actA = (jA < j1) && (k < k1), actB = (jB < j1) && (k < k1);
  const long long plane = (long long)nj * nk;
  const int i_first = i_lo - 1;

  // copy duties of this thread (fixed across planes): its two own elements + at most one ring element
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
  const long long halo_off = halo_ok ? (long long)(j0 - 1 + hr) * nk + (k0 - 1 + hc) : 0;
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(&xs[0][0][0]);
  const unsigned own_s = smem_base + (unsigned)(((2 * ty + 1) * SROW + tx + 1) * 8);
  const unsigned halo_s = smem_base + (unsigned)((hr * SROW + hc) * 8);

  // running pointers: copy sources (plane being issued), outputs (plane ip-1).  They may point
  // outside the arrays for planes outside the domain; those are never dereferenced.
  const double* px[NF];
  const double* ph[NF];
  double* py[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    px[f] = s.x[f] + (long long)i_first * plane + own_off;
    ph[f] = s.x[f] + (long long)i_first * plane + halo_off;
    py[f] = s.y[f] + (long long)(i_first - 2) * plane + own_off;  // advanced before use: plane ip-1
  }
  int ipl = i_first;  // plane index the copy pointers refer to

  // copy plane `ipl` into ring slot SL (zero-fill when outside the domain), then advance
#define DPP_ISSUE(SL)                                                                           \
  {                                                                                             \
    const bool in = (unsigned)ipl < (unsigned)ni;                                               \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                            \
      cp_async8(own_s + ((SL)*NF + f) * SLOT * 8, px[f], in && ownA_ok);                        \
      cp_async8(own_s + ((SL)*NF + f) * SLOT * 8 + SROW * 8, px[f] + nk, in && ownB_ok);        \
      if (is_halo) cp_async8(halo_s + ((SL)*NF + f) * SLOT * 8, ph[f], in && halo_ok);          \
      px[f] += plane;                                                                           \
      ph[f] += plane;                                                                           \
    }                                                                                           \
    cp_async_commit();                                                                          \
    ++ipl;                                                                                      \
  }

  // in-plane coefficients (centre entries come from the tables: domain-boundary nodes differ)
  const double myo = s.mo[1], mzo = s.mo[2], kyo = s.ko[1], kzo = s.ko[2];
  const double mCor = myo * mzo, kCor = kyo * mzo + myo * kzo;
  double mEJ = 0, kEJ = 0;                       // neighbours (j+-1, k): depend on k only
  double mEK[2] = {0, 0}, kEK[2] = {0, 0};       // neighbours (j, k+-1): depend on j only
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
  const double* tbase = &xs[0][0][2 * ty * SROW + tx];

  __syncthreads();  // the ring is free (previous run of this CTA finished)
  // prologue: planes i_first (slot 2) and i_first+1 (slot 0) in flight
  DPP_ISSUE(2)
  DPP_ISSUE(0)

  // one plane step.  Plane ip lives in ring slot C, plane ip+1 in slot A (in flight), plane ip+2
  // is issued into slot B; A,B,C are also the queue slots of planes ip-2, ip-1, ip.
#define DPP_STEP(A, B, C)                                                                             \
  {                                                                                                   \
    cp_async_wait<RING - 2>();                                                                        \
    __syncthreads();                                                                                  \
    DPP_ISSUE(B)                                                                                      \
    double cen[NF][2];                                                                                \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      const double* t = tbase + ((C)*NF + f) * SLOT;                                                  \
      const double e0 = t[0] + t[2], c0 = t[1];                                                       \
      const double e1 = t[SROW] + t[SROW + 2], c1 = t[SROW + 1];                                      \
      const double e2 = t[2 * SROW] + t[2 * SROW + 2], c2 = t[2 * SROW + 1];                          \
      const double e3 = t[3 * SROW] + t[3 * SROW + 2], c3 = t[3 * SROW + 1];                          \
      const double corA = e0 + e2, ejA = c0 + c2, corB = e1 + e3, ejB = c1 + c3;                      \
      cen[f][0] = c1;                                                                                 \
      cen[f][1] = c2;                                                                                 \
      qc[f][0][C] = fma(mCor, corA, fma(mEK[0], e1, fma(mEJ, ejA, mC[0] * c1)));                      \
      qd[f][0][C] = fma(kCor, corA, fma(kEK[0], e1, fma(kEJ, ejA, kC[0] * c1)));                      \
      qc[f][1][C] = fma(mCor, corB, fma(mEK[1], e2, fma(mEJ, ejB, mC[1] * c2)));                      \
      qd[f][1][C] = fma(kCor, corB, fma(kEK[1], e2, fma(kEJ, ejB, kC[1] * c2)));                      \
    }                                                                                                 \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) py[f] += plane;                                    \
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
          py[f][r * nk] = yv;                                                                         \
          dot = fma(prev_cen[f][r], yv, dot);                                                         \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
    _Pragma("unroll") for (int f = 0; f < NF; ++f) {                                                  \
      prev_cen[f][0] = cen[f][0];                                                                     \
      prev_cen[f][1] = cen[f][1];                                                                     \
    }                                                                                                 \
  }

  // planes i_first .. i_hi ; outputs i_lo .. i_hi-1
  int ip = i_first;
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
  }  // runs

  if (s.dot_partials != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (tx == 0) red[ty] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < TY; ++w) t += red[w];
      s.dot_partials[blockIdx.x] = t;
    }
  }
}

// xm = mask ? 0 : x   (arbitrary input of the public apply; Krylov vectors never need it)
__global__ void k_premask(long long n, const double* __restrict__ x, const uint8_t* __restrict__ m,
                          double* __restrict__ xm) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    xm[i] = m[i] ? 0.0 : x[i];
}

// Row elimination: y[node] = identity ? xid[node] : 0 for the constrained nodes of up to two fields
struct FixArgs {
  const int32_t* nodes[2];
  long long count[2];
  double* y[2];
  const double* xid[2];
  int identity;
  long long ob, oe;
  const double* skip_flag;
};

__global__ void k_fix_rows(const FixArgs a) {
  if (a.skip_flag != nullptr && *a.skip_flag != 0.0) return;
  const long long total = a.count[0] + a.count[1];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int f = t < a.count[0] ? 0 : 1;
    const long long node = a.nodes[f][f ? t - a.count[0] : t];
    if (node >= a.ob && node < a.oe) a.y[f][node] = a.identity ? a.xid[f][node] : 0.0;
  }
}

}  // namespace

int structured_fix_rows(dpp_context* ctx, int nf, const int* fld, double* const* y, const double* const* xid,
                        int identity, const double* skip_flag) {
  FixArgs fx{};
  for (int f = 0; f < nf; ++f) {
    if (fld[f] < 0) continue;  // this field has no row elimination
    fx.nodes[f] = ctx->d_bc_nodes[fld[f]];
    fx.count[f] = ctx->n_bc[fld[f]];
    fx.y[f] = y[f];
    fx.xid[f] = xid[f];
  }
  const long long total = fx.count[0] + fx.count[1];
  if (total <= 0) return DPP_OK;
  fx.identity = identity;
  fx.ob = ctx->owned_begin;
  fx.oe = ctx->owned_end;
  fx.skip_flag = skip_flag;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8));
  k_fix_rows<<<blocks, 256, 0, ctx->stream>>>(fx);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

int structured_apply_uniform(dpp_context* ctx, const OpArgs& a, int* n_partial_blocks) {
  const GridDesc& g = ctx->grid;
  const long long plane = (long long)g.n[1] * g.n[2];
  if (a.owned_begin % plane || a.owned_end % plane) {
    ctx->set_error("structured apply: owned range must consist of whole x-planes");
    return DPP_ERR_INVALID;
  }
  UArgs s{};
  for (int d = 0; d < 3; ++d) {
    s.n[d] = g.n[d]; s.m1d[d] = g.m1d[d]; s.k1d[d] = g.k1d[d];
    s.mo[d] = ctx->uni_m_off[d]; s.ko[d] = ctx->uni_k_off[d];
  }
  s.mxc_i = ctx->uni_mxc[0]; s.mxc_b = ctx->uni_mxc[1];
  s.kxc_i = ctx->uni_kxc[0]; s.kxc_b = ctx->uni_kxc[1];
  FixArgs fx{};
  for (int f = 0; f < a.nf; ++f) {
    s.x[f] = a.x[f]; s.y[f] = a.y[f];
    if (a.in_mask[f] != nullptr && !a.input_premasked) {
      // arbitrary input: zero the eliminated columns into scratch first
      if (!ctx->d_premask) DPP_CHECK(dev_alloc(ctx, &ctx->d_premask, 2 * ctx->n_nodes));
      double* xm = ctx->d_premask + (size_t)f * ctx->n_nodes;
      const int blocks = (int)std::min<long long>((ctx->n_nodes + 255) / 256, (long long)ctx->sm_count * 16);
      k_premask<<<blocks, 256, 0, ctx->stream>>>(ctx->n_nodes, a.x[f], a.in_mask[f], xm);
      ctx->launches++;
      s.x[f] = xm;
    }
    if (a.out_mask[f] != nullptr) {
      const long long fld = (a.out_mask[f] - ctx->d_mask) / ctx->n_nodes;
      if (fld < 0 || fld > 1 || a.out_mask[f] != ctx->d_mask + fld * ctx->n_nodes) {
        ctx->set_error("structured apply: out_mask must be a field of the handle's Dirichlet mask");
        return DPP_ERR_INVALID;
      }
      fx.nodes[f] = ctx->d_bc_nodes[fld];
      fx.count[f] = ctx->n_bc[fld];
      fx.y[f] = a.y[f];
      fx.xid[f] = a.x[f];
    }
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
  // one wave of persistent CTAs (2 per SM) with equal shares of the (tile, plane) steps; small grids
  // get fewer CTAs so that a run stays >= ~8 planes
  const long long total = (long long)tiles * nown;
  int nctas;
  if (tiles <= kMaxPartialBlocks / 2) {
    s.nseg = choose_x_segments(tiles, nown, ctx->sm_count * 2, kMaxPartialBlocks);
    nctas = tiles * s.nseg;
  } else {
    s.nseg = 0;
    nctas = (int)std::max<long long>(1, std::min<long long>(ctx->sm_count * 2, total / 8));
  }
  dim3 grid(nctas), block(TK, TY);
  if (a.nf == 2)
    k_apply_uniform<2><<<grid, block, 0, ctx->stream>>>(s);
  else
    k_apply_uniform<1><<<grid, block, 0, ctx->stream>>>(s);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (fx.count[0] + fx.count[1] > 0) {
    fx.identity = a.identity_on_masked;
    fx.ob = a.owned_begin;
    fx.oe = a.owned_end;
    fx.skip_flag = a.skip_flag;
    const long long total = fx.count[0] + fx.count[1];
    const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8));
    k_fix_rows<<<blocks, 256, 0, ctx->stream>>>(fx);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
  }
  if (n_partial_blocks) *n_partial_blocks = nctas;
  return DPP_OK;
}

}  // namespace dpp
