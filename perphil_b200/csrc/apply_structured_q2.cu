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
  s.nseg = nseg;
  dim3 grid(tiles * nseg), block(TK, TJ);
  if (a.nf == 2)
    k_apply_q2<2><<<grid, block, 0, ctx->stream>>>(s);
  else
    k_apply_q2<1><<<grid, block, 0, ctx->stream>>>(s);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (need_fix) DPP_CHECK(structured_fix_rows(ctx, a.nf, fld, ys, xid, a.identity_on_masked, a.skip_flag));
  if (n_partial_blocks) *n_partial_blocks = tiles * nseg;
  return DPP_OK;
}

}  // namespace dpp
