// See vector_ops.cuh.  HBM-bound streaming kernels: grid = (blocks, nf), each block walks a
// contiguous chunk of one field with 4 independent 8-byte loads in flight per thread per array.
#include <algorithm>
#include <cmath>

#include "cg_device.cuh"

namespace dpp {

namespace {

constexpr int UNROLL = 4;

struct Chunk {
  long long begin, end;  // element range inside the field (absolute index into the array)
};

__device__ __forceinline__ Chunk my_chunk(const VecLayout& L) {
  const long long nown = L.oe - L.ob;
  const long long per = (nown + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * per;
  const long long e = b + per < nown ? b + per : nown;
  const long long base = (long long)blockIdx.y * L.stride + L.ob;
  return Chunk{base + b, base + (e > b ? e : b)};
}

__global__ void __launch_bounds__(VT) k_reduce_partials(const double* __restrict__ partials, int nblocks, int width,
                                                         double* S, double* hist, int post, int do_post,
                                                         int out_offset) {
  __shared__ double sm[VT / 32];
  for (int w = 0; w < width; ++w) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += VT) v += partials[(size_t)b * width + w];
    const double t = block_sum(v, sm);
    if (threadIdx.x == 0) S[S_TMP + out_offset + w] = t;
  }
  if (threadIdx.x == 0 && do_post) apply_post(S, hist, post);
}

__global__ void k_post_only(double* S, double* hist, int post) { apply_post(S, hist, post); }

// Local reduction + all-reduce over the ranks' mailboxes + post-op, one block (cg_device.cuh)
__global__ void __launch_bounds__(VT) k_reduce_partials_ipc(const double* __restrict__ partials, int nblocks, int width,
                                                             double* S, double* hist, int post, int out_offset,
                                                             IpcReduce ipc) {
  __shared__ double sm[kFinishSmem];
  finish_reduction(partials, nblocks, width, S, hist, post, out_offset, ipc, sm);
}

// The same all-reduce for up to kMboxWideVals values (GMRES(30): the j + 1 <= 31 dot products of one classical
// Gram-Schmidt pass).  Tagged words as in finish_reduction (value = flag: two 8-byte words per value, the upper halves
// carry the sequence number), but one THREAD per (peer, value): the 2 x 31 x world remote stores and the polls of the
// own mailbox run side by side, so a wide exchange costs about what a one-value exchange costs.  Shares the sequence
// counter with the narrow mailboxes (every rank runs the same sequence of exchanges, narrow and wide alike).
__global__ void __launch_bounds__(VT) k_reduce_partials_wide(const double* __restrict__ partials, int nblocks, int width,
                                                              double* S, double* hist, int post, int out_offset,
                                                              IpcReduce ipc) {
  __shared__ double sm[VT / 32];
  __shared__ double vals[kMboxWideVals];
  __shared__ double recv[kMaxIpcRanks][kMboxWideVals];
  __shared__ unsigned long long seq_sm;
  __shared__ int timed_out;
  const int tid = threadIdx.x;
  for (int w = 0; w < width; ++w) {
    double v = 0.0;
    for (int b = tid; b < nblocks; b += VT) v += partials[(size_t)b * width + w];
    const double t = block_sum(v, sm);
    if (tid == 0) vals[w] = t;
  }
  if (tid == 0) {
    timed_out = 0;
    seq_sm = *ipc.seq_dev + 1;
    *ipc.seq_dev = seq_sm;
  }
  __syncthreads();
  const unsigned long long seq = seq_sm;
  const int slot = (int)(seq & 1ull);
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const int items = ipc.world * width;
  const long long t0 = clock64();
  // every storing thread orders what this rank wrote before (and what it observed) ahead of its mailbox words
  if (tid < items) asm volatile("fence.acq_rel.sys;" ::: "memory");
  for (int it = tid; it < items; it += VT) {
    const int peer = it % ipc.world, w = it / ipc.world;
    const size_t mine = kMboxWideOffset + ((size_t)slot * ipc.world + ipc.rank) * kMboxWideWords;
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(ipc.peer[peer]) + mine;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[w]);
    dst[2 * w] = tag | (bits & 0xffffffffull);
    dst[2 * w + 1] = tag | (bits >> 32);
  }
  for (int it = tid; it < items; it += VT) {
    const int peer = it % ipc.world, w = it / ipc.world;
    const size_t theirs = kMboxWideOffset + ((size_t)slot * ipc.world + peer) * kMboxWideWords;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(ipc.local) + theirs;
    unsigned long long lo, hi;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(lo) : "l"(src + 2 * w) : "memory");
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(hi) : "l"(src + 2 * w + 1) : "memory");
      if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
      if (clock64() - t0 > 60000000000LL) {  // ~30 s: a peer died; report instead of hanging the GPU
        timed_out = 1;
        break;
      }
    }
    recv[peer][w] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
  }
  __syncthreads();
  if (timed_out) {
    if (tid == 0) S[S_REASON] = (double)DPP_DIVERGED_COMM_TIMEOUT;
    return;
  }
  if (tid < width) {   // rank order: every rank adds the same numbers in the same order
    double t = 0.0;
    for (int r = 0; r < ipc.world; ++r) t += recv[r][tid];
    S[S_TMP + out_offset + tid] = t;
  }
  __syncthreads();
  if (tid == 0 && post != POST_NONE) {
    __threadfence();
    apply_post(S, hist, post);
  }
}

__global__ void __launch_bounds__(VT) k_axpby(VecLayout L, double a, const double* x, double b, double* y) {
  const Chunk c = my_chunk(L);
  for (long long i = c.begin + threadIdx.x; i < c.end; i += VT) {
    const double yv = (b == 0.0) ? 0.0 : b * y[i];
    y[i] = fma(a, x[i], yv);
  }
}

__global__ void __launch_bounds__(VT) k_pointwise(VecLayout L, const double* __restrict__ d, const double* __restrict__ r,
                                                   double* __restrict__ z) {
  const Chunk c = my_chunk(L);
  for (long long i = c.begin + threadIdx.x; i < c.end; i += VT) z[i] = d ? d[i] * r[i] : r[i];
}

__global__ void __launch_bounds__(VT) k_pbjacobi(VecLayout L, const double* __restrict__ i00, const double* __restrict__ i01,
                                                  const double* __restrict__ i11, const double* __restrict__ r,
                                                  double* __restrict__ z) {
  // blockIdx.y ignored: node loop handles both fields
  const long long nown = L.oe - L.ob;
  for (long long t = (long long)blockIdx.x * VT + threadIdx.x; t < nown; t += (long long)gridDim.x * VT) {
    const long long n = L.ob + t;
    const double r0 = r[n], r1 = r[L.stride + n];
    z[n] = i00[n] * r0 + i01[n] * r1;
    z[L.stride + n] = i01[n] * r0 + i11[n] * r1;
  }
}

__global__ void __launch_bounds__(VT) k_dot2(VecLayout L, const double* __restrict__ a0, const double* __restrict__ b0,
                                              const double* __restrict__ a1, const double* __restrict__ b1,
                                              double* __restrict__ partials) {
  __shared__ double sm[VT / 32];
  const Chunk c = my_chunk(L);
  double s0 = 0.0, s1 = 0.0;
  for (long long i = c.begin + threadIdx.x; i < c.end; i += VT) {
    s0 = fma(a0[i], b0[i], s0);
    if (a1 != nullptr) s1 = fma(a1[i], b1[i], s1);
  }
  const double t0 = block_sum(s0, sm);
  const double t1 = block_sum(s1, sm);
  if (threadIdx.x == 0) {
    const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    partials[b * 2] = t0;
    partials[b * 2 + 1] = t1;
  }
}

// p = z + (rz/rz_old) p   with z = dinv.*r (Jacobi), z = r (none) or an explicit z vector
__global__ void __launch_bounds__(VT) k_cg_p_update(VecLayout L, double* __restrict__ p, const double* __restrict__ r,
                                                     const double* __restrict__ dinv, const double* __restrict__ z,
                                                     const double* __restrict__ S) {
  if (S[S_REASON] != 0.0) return;
  const double beta = (S[S_ITS] == 0.0) ? 0.0 : S[S_RZ] / S[S_RZ_OLD];
  const Chunk c = my_chunk(L);
  long long i = c.begin + threadIdx.x;
  for (; i + (UNROLL - 1) * VT < c.end; i += UNROLL * VT) {
    double zv[UNROLL], pv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long q = i + u * VT;
      zv[u] = z ? z[q] : (dinv ? dinv[q] * r[q] : r[q]);
      pv[u] = (beta == 0.0) ? 0.0 : p[q];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) p[i + u * VT] = fma(beta, pv[u], zv[u]);
  }
  for (; i < c.end; i += VT) {
    const double zv = z ? z[i] : (dinv ? dinv[i] * r[i] : r[i]);
    const double pv = (beta == 0.0) ? 0.0 : p[i];
    p[i] = fma(beta, pv, zv);
  }
}

// x += a p ; r -= a w ; [z = dinv.*r ; partial <r,z>, <z,z>]    a = rz / pAp
template <bool FUSED_PC>
__global__ void __launch_bounds__(VT) k_cg_xr_update(VecLayout L, double* __restrict__ x, double* __restrict__ r,
                                                      const double* __restrict__ p, const double* __restrict__ w,
                                                      const double* __restrict__ dinv, const double* __restrict__ S,
                                                      double* __restrict__ partials) {
  __shared__ double sm[VT / 32];
  if (S[S_REASON] != 0.0) return;
  const double a = S[S_RZ] / S[S_PAP];
  const Chunk c = my_chunk(L);
  double srz = 0.0, szz = 0.0;
  long long i = c.begin + threadIdx.x;
  for (; i + (UNROLL - 1) * VT < c.end; i += UNROLL * VT) {
    double xv[UNROLL], rv[UNROLL], pv[UNROLL], wv[UNROLL], dv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long q = i + u * VT;
      xv[u] = x[q]; rv[u] = r[q]; pv[u] = p[q]; wv[u] = w[q];
      dv[u] = (FUSED_PC && dinv) ? dinv[q] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long q = i + u * VT;
      x[q] = fma(a, pv[u], xv[u]);
      const double rn = fma(-a, wv[u], rv[u]);
      r[q] = rn;
      if (FUSED_PC) {
        const double zv = dv[u] * rn;
        srz = fma(rn, zv, srz);
        szz = fma(zv, zv, szz);
      }
    }
  }
  for (; i < c.end; i += VT) {
    x[i] = fma(a, p[i], x[i]);
    const double rn = fma(-a, w[i], r[i]);
    r[i] = rn;
    if (FUSED_PC) {
      const double zv = (dinv ? dinv[i] : 1.0) * rn;
      srz = fma(rn, zv, srz);
      szz = fma(zv, zv, szz);
    }
  }
  if (FUSED_PC) {
    const double t0 = block_sum(srz, sm);
    const double t1 = block_sum(szz, sm);
    if (threadIdx.x == 0) {
      const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      partials[b * 2] = t0;
      partials[b * 2 + 1] = t1;
    }
  }
}

struct VecPtrs {
  const double* v[32];
};

// partials[b*width + j] = <V_j, w> over the block's chunk, j < nv (classical Gram-Schmidt VecMDot)
__global__ void __launch_bounds__(VT) k_mdot(VecLayout L, VecPtrs V, int nv, const double* __restrict__ w,
                                              double* __restrict__ partials, int width, const double* skip) {
  __shared__ double sm[VT / 32];
  if (skip != nullptr && *skip != 0.0) return;
  const Chunk c = my_chunk(L);
  const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int j0 = 0; j0 < nv; j0 += 8) {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    for (long long i = c.begin + threadIdx.x; i < c.end; i += VT) {
      const double wv = w[i];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (j0 + q < nv) acc[q] = fma(V.v[j0 + q][i], wv, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (j0 + q < nv) {
        const double t = block_sum(acc[q], sm);
        if (threadIdx.x == 0) partials[b * width + j0 + q] = t;
      }
    }
  }
}

// w -= sum_j h[j] V_j ; partial ||w||^2      (h = S[S_TMP .. S_TMP+nv))
// nv_dev (optional): the number of vectors is read from device memory (GMRES cycle length)
__global__ void __launch_bounds__(VT) k_maxpy_norm(VecLayout L, VecPtrs V, int nv, double* __restrict__ w,
                                                    const double* __restrict__ h, double sign,
                                                    double* __restrict__ partials, const double* skip,
                                                    const double* nv_dev) {
  __shared__ double sm[VT / 32];
  __shared__ double hs[32];
  if (skip != nullptr && *skip != 0.0) return;
  if (nv_dev != nullptr) nv = (int)*nv_dev;
  if (threadIdx.x < nv) hs[threadIdx.x] = h[threadIdx.x];
  __syncthreads();
  const Chunk c = my_chunk(L);
  double s = 0.0;
  for (long long i = c.begin + threadIdx.x; i < c.end; i += VT) {
    double wv = w[i];
    for (int j = 0; j < nv; ++j) wv = fma(sign * hs[j], V.v[j][i], wv);
    w[i] = wv;
    s = fma(wv, wv, s);
  }
  if (partials != nullptr) {
    const double t = block_sum(s, sm);
    if (threadIdx.x == 0) {
      const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      partials[b] = t;
    }
  }
}

}  // namespace

int vec_launch_blocks(const dpp_context* ctx, const VecLayout& L) {
  const long long nown = L.oe - L.ob;
  long long want = (nown + (long long)VT * UNROLL * 4 - 1) / ((long long)VT * UNROLL * 4);
  const long long cap = std::max(1, (ctx->sm_count * 8) / std::max(1, L.nf));
  want = std::max(1LL, std::min(want, cap));
  return (int)std::min<long long>(want, kMaxPartialBlocks / 2);
}

#define VLAUNCH(kernel, ...)                                       \
  do {                                                             \
    dim3 grid__(vec_launch_blocks(ctx, L), L.nf);                  \
    kernel<<<grid__, VT, 0, ctx->stream>>>(__VA_ARGS__);           \
    ctx->launches++;                                               \
    DPP_CUDA(cudaGetLastError());                                  \
  } while (0)

int vec_axpby(dpp_context* ctx, const VecLayout& L, double a, const double* x, double b, double* y) {
  VLAUNCH(k_axpby, L, a, x, b, y);
  return DPP_OK;
}

int vec_scale_into(dpp_context* ctx, const VecLayout& L, double a, const double* x, double* y) {
  VLAUNCH(k_axpby, L, a, x, 0.0, y);
  return DPP_OK;
}

int vec_pointwise_mult(dpp_context* ctx, const VecLayout& L, const double* dinv, const double* r, double* z) {
  VLAUNCH(k_pointwise, L, dinv, r, z);
  return DPP_OK;
}

int vec_pbjacobi(dpp_context* ctx, const VecLayout& L, const double* i00, const double* i01, const double* i11,
                 const double* r, double* z) {
  dim3 grid(vec_launch_blocks(ctx, L) * 2, 1);
  k_pbjacobi<<<grid, VT, 0, ctx->stream>>>(L, i00, i01, i11, r, z);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

double* hist_device(dpp_context* ctx, int slot) { return ctx->hist_cap[slot] > 0 ? ctx->d_hist[slot] : nullptr; }

int reduce_partials(dpp_context* ctx, int nblocks, int width, int slot, PostOp post, int out_offset) {
  double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  const bool dist = ctx->world > 1;
  if (dist && comm_ipc_ready(ctx) && width < kMboxEntry) {
    k_reduce_partials_ipc<<<1, VT, 0, ctx->stream>>>(ctx->d_partials, nblocks, width, S, hist_device(ctx, slot), (int)post,
                                                     out_offset, comm_ipc_reduce_args(ctx));
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
    return DPP_OK;
  }
  if (dist && comm_ipc_ready(ctx) && width <= kMboxWideVals) {
    const IpcReduce ipc = comm_ipc_reduce_args(ctx);
    if (ipc.world > 1 && (ipc.ll == 1 || ipc.ll == 3)) {   // tagged-word protocol only; else NCCL below
      k_reduce_partials_wide<<<1, VT, 0, ctx->stream>>>(ctx->d_partials, nblocks, width, S, hist_device(ctx, slot),
                                                        (int)post, out_offset, ipc);
      ctx->launches++;
      DPP_CUDA(cudaGetLastError());
      return DPP_OK;
    }
  }
  k_reduce_partials<<<1, VT, 0, ctx->stream>>>(ctx->d_partials, nblocks, width, S, hist_device(ctx, slot), (int)post,
                                               dist ? 0 : 1, out_offset);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  if (dist) {
    DPP_CHECK(comm_allreduce_sum(ctx, S + S_TMP + out_offset, width));
    if (post != POST_NONE) {
      k_post_only<<<1, 1, 0, ctx->stream>>>(S, hist_device(ctx, slot), (int)post);
      ctx->launches++;
      DPP_CUDA(cudaGetLastError());
    }
  }
  return DPP_OK;
}

int vec_dot2(dpp_context* ctx, const VecLayout& L, const double* a0, const double* b0, const double* a1,
             const double* b1, int slot, PostOp post) {
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_dot2<<<grid, VT, 0, ctx->stream>>>(L, a0, b0, a1, b1, ctx->d_partials);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return reduce_partials(ctx, grid.x * grid.y, 2, slot, post);
}

int cg_p_update(dpp_context* ctx, const VecLayout& L, double* p, const double* r, const double* dinv, const double* z,
                int slot) {
  const double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  VLAUNCH(k_cg_p_update, L, p, r, dinv, z, S);
  return DPP_OK;
}

int cg_xr_update(dpp_context* ctx, const VecLayout& L, double* x, double* r, const double* p, const double* w,
                 const double* dinv, bool fused_pc, int slot, PostOp post) {
  const double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  if (fused_pc) {
    k_cg_xr_update<true><<<grid, VT, 0, ctx->stream>>>(L, x, r, p, w, dinv, S, ctx->d_partials);
    ctx->launches++;
    DPP_CUDA(cudaGetLastError());
    return reduce_partials(ctx, grid.x * grid.y, 2, slot, post);
  }
  k_cg_xr_update<false><<<grid, VT, 0, ctx->stream>>>(L, x, r, p, w, dinv, S, ctx->d_partials);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

int gmres_mdot(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv, const double* w, int slot,
               const double* skip) {
  VecPtrs P{};
  for (int j = 0; j < nv; ++j) P.v[j] = V[j];
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_mdot<<<grid, VT, 0, ctx->stream>>>(L, P, nv, w, ctx->d_partials, nv, skip);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return reduce_partials(ctx, grid.x * grid.y, nv, slot, POST_NONE);
}

int gmres_maxpy_norm(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv, double* w, int slot,
                     const double* skip) {
  VecPtrs P{};
  for (int j = 0; j < nv; ++j) P.v[j] = V[j];
  const double* S = ctx->d_scalars + (size_t)slot * S_SLOT_SIZE;
  // h lives in S[S_TMP..]; copy to spare area because the norm reduction overwrites S_TMP
  double* hcopy = ctx->d_scalars + (size_t)kNumScalars - 40;
  DPP_CUDA(cudaMemcpyAsync(hcopy, S + S_TMP, sizeof(double) * nv, cudaMemcpyDeviceToDevice, ctx->stream));
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_maxpy_norm<<<grid, VT, 0, ctx->stream>>>(L, P, nv, w, hcopy, -1.0, ctx->d_partials, skip, nullptr);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  // h stays in S[S_TMP .. S_TMP+nv) for the host; ||w||^2 goes to S[S_TMP + kGmresNormOffset]
  return reduce_partials(ctx, grid.x * grid.y, 1, slot, POST_NONE, kGmresNormOffset);
}

int vec_maxpy_host(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv, const double* coef,
                   double* x) {
  VecPtrs P{};
  for (int j = 0; j < nv; ++j) P.v[j] = V[j];
  double* hcopy = ctx->d_scalars + (size_t)kNumScalars - 40;
  DPP_CUDA(cudaMemcpyAsync(hcopy, coef, sizeof(double) * nv, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_maxpy_norm<<<grid, VT, 0, ctx->stream>>>(L, P, nv, x, hcopy, 1.0, nullptr, nullptr, nullptr);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

// x += sum_{j < *nv_dev} coef_dev[j] V_j   (coefficients and count live on the device)
int vec_maxpy_dev(dpp_context* ctx, const VecLayout& L, const double* const* V, int nv_max, const double* coef_dev,
                  const double* nv_dev, double* x) {
  VecPtrs P{};
  for (int j = 0; j < nv_max; ++j) P.v[j] = V[j];
  dim3 grid(vec_launch_blocks(ctx, L), L.nf);
  k_maxpy_norm<<<grid, VT, 0, ctx->stream>>>(L, P, nv_max, x, coef_dev, 1.0, nullptr, nullptr, nv_dev);
  ctx->launches++;
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

__global__ void __launch_bounds__(VT) k_scale_dev(VecLayout L, double* __restrict__ v, const double* __restrict__ factor,
                                                   const double* skip0, const double* skip1) {
  if ((skip0 != nullptr && *skip0 != 0.0) || (skip1 != nullptr && *skip1 != 0.0)) return;
  const double a = *factor;
  const Chunk c = my_chunk(L);
  for (long long i = c.begin + threadIdx.x; i < c.end; i += VT) v[i] *= a;
}

// v *= *factor_dev unless *skip0 or *skip1 is non-zero
int vec_scale_dev(dpp_context* ctx, const VecLayout& L, double* v, const double* factor_dev, const double* skip0,
                  const double* skip1) {
  VLAUNCH(k_scale_dev, L, v, factor_dev, skip0, skip1);
  return DPP_OK;
}

int scalars_fetch(dpp_context* ctx, int slot) {
  DPP_CUDA(cudaMemcpyAsync(ctx->h_scalars + (size_t)slot * S_SLOT_SIZE, ctx->d_scalars + (size_t)slot * S_SLOT_SIZE,
                           sizeof(double) * S_SLOT_SIZE, cudaMemcpyDeviceToHost, ctx->stream));
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DPP_OK;
}

int scalars_init(dpp_context* ctx, int slot, double rtol, double atol, double dtol, int max_it, int hist_cap) {
  double* h = ctx->h_scalars + (size_t)slot * S_SLOT_SIZE;
  for (int i = 0; i < S_SLOT_SIZE; ++i) h[i] = 0.0;
  h[S_RTOL] = rtol; h[S_ATOL] = atol; h[S_DTOL] = dtol; h[S_MAXIT] = (double)max_it;
  h[S_HISTCAP] = (double)hist_cap; h[S_RZ_OLD] = 1.0;
  DPP_CUDA(cudaMemcpyAsync(ctx->d_scalars + (size_t)slot * S_SLOT_SIZE, h, sizeof(double) * S_SLOT_SIZE,
                           cudaMemcpyHostToDevice, ctx->stream));
  // the pinned mirror is reused by scalars_fetch: make sure the upload has been consumed
  DPP_CUDA(cudaStreamSynchronize(ctx->stream));
  return DPP_OK;
}

int vec_zero(dpp_context* ctx, double* x, int64_t n) {
  DPP_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, ctx->stream));
  return DPP_OK;
}

int vec_copy(dpp_context* ctx, double* dst, const double* src, int64_t n) {
  DPP_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  return DPP_OK;
}

}  // namespace dpp
