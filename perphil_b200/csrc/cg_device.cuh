// Device-side pieces shared by the reduction kernels (vector_ops.cu) and the fused CG kernels
// (cg_fused_uniform.cu): deterministic block sums, the PETSc KSPCG bookkeeping that runs on the device
// after each global reduction, and the reduction epilogue itself (local sum in block order, optional
// mailbox all-reduce over peer memory, post-op).
#pragma once

#include "vector_ops.cuh"

namespace dpp {
namespace {

constexpr int VT = 256;   // threads per block of every kernel that reduces
constexpr int kFinishSmem = VT / 32 + kMboxEntry + 2 + kMaxIpcRanks * (kMboxEntry - 1);

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum in a fixed order; result valid in thread 0 (linear thread id; NTH threads per block)
template <int NTH = VT>
__device__ __forceinline__ double block_sum(double v, double* sm /*[VT/32]*/) {
  v = warp_sum(v);
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  const int lane = tid & 31, wid = tid >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < NTH / 32; ++w) t += sm[w];
  }
  return t;
}

// KSPCG bookkeeping after a reduction, executed by one thread.  t0, t1: the reduced values; S: the scalar slot
// (written); R: where the slot's PREVIOUS state is read from -- the slot itself, or a shared-memory snapshot the
// kernel took in its prologue (S changes only in these epilogues, stream-ordered, so the snapshot is current;
// reading it saves two dependent L2 round trips on the critical path of every iteration).
// Deferred x update (cg_fused_uniform.cu): the search directions of the last kXRing = 16 iterations stay in a ring of
// buffers, their step lengths in `xring` ([kXRing] alpha, [kXRing] iteration tag); x += alpha_k p_k is applied for
// kXRing - 1 iterations at a time by the r-update kernel (in iteration order: bitwise the x of KSPCG).
constexpr int kXRing = 16;

__device__ __forceinline__ void apply_post(double* S, double* hist, int post, const double* R, double t0, double t1,
                                           double* xring = nullptr) {
  if (post == POST_CG_PAP) {
    if (R[S_REASON] != 0.0) return;  // the apply was a no-op: partials are stale
    const double pap = t0;
    S[S_PAP] = pap;
    if (!(pap > 0.0)) {
      S[S_REASON] = (pap == pap) ? DPP_DIVERGED_INDEFINITE_MAT : DPP_DIVERGED_NANORINF;
      S[S_XPEND] = 0.0;
    } else {
      const double alpha = R[S_RZ] / pap;
      S[S_ALPHA] = alpha;
      S[S_XPEND] = 1.0;
      if (xring != nullptr) {   // direction p_k of iteration k = its lives in ring buffer (k + 1) % kXRing
        const int k = (int)R[S_ITS], j = (k + 1) % kXRing;
        xring[j] = alpha;
        xring[kXRing + j] = (double)k;
      }
    }
    return;
  }
  if (post == POST_CG_INIT || post == POST_CG_RZ) {
    if (R[S_REASON] != 0.0) return;
    const double rz = t0, zz = t1;
    const double rnorm = sqrt(zz);
    const double atol = R[S_ATOL];
    double ttol = R[S_TTOL], rnorm0 = R[S_RNORM0];
    int its;
    if (post == POST_CG_INIT) {
      its = 0;
      S[S_RZ_OLD] = 1.0;
      S[S_RNORM0] = rnorm;
      rnorm0 = rnorm;
      const double t = R[S_RTOL] * rnorm;
      ttol = t > atol ? t : atol;
      S[S_TTOL] = ttol;
    } else {
      its = (int)R[S_ITS] + 1;
      S[S_RZ_OLD] = R[S_RZ];
    }
    S[S_RZ] = rz;
    S[S_ZZ] = zz;
    S[S_RNORM] = rnorm;
    S[S_ITS] = (double)its;
    if (hist != nullptr && its < (int)R[S_HISTCAP]) hist[its] = rnorm;
    // KSPConvergedDefault
    double reason = 0.0;
    if (!(rnorm == rnorm) || isinf(rnorm)) reason = DPP_DIVERGED_NANORINF;
    else if (rnorm <= ttol) reason = (rnorm < atol) ? DPP_CONVERGED_ATOL : DPP_CONVERGED_RTOL;
    else if (rnorm >= R[S_DTOL] * rnorm0) reason = DPP_DIVERGED_DTOL;
    else if (rz == 0.0) reason = DPP_CONVERGED_ATOL;
    else if (its >= (int)R[S_MAXIT]) reason = DPP_DIVERGED_ITS;
    S[S_REASON] = reason;
  }
}
__device__ __forceinline__ void apply_post(double* S, double* hist, int post) {
  apply_post(S, hist, post, S, S[S_TMP], S[S_TMP + 1]);
}

constexpr int kPreScalars = 16;   // S[0 .. S_TMP): what a kernel prologue snapshots for its epilogue
struct FoldPre {                  // shared-memory snapshot taken after griddepcontrol.wait (null members: read global)
  const double* S;                // [kPreScalars]
  const unsigned long long* seq;  // mailbox sequence counter
};

// Reduction epilogue run by ONE block of VT threads: S[S_TMP + out_offset + w] = sum over blocks (fixed
// order) of partials[b*width + w], all-reduced over the ranks' mailboxes when ipc.world > 1, then the
// post-op.  Mailbox protocol: every rank writes its sums into slot (seq & 1) of EVERY rank's mailbox, then
// waits until all ranks' entries of this sequence number arrived in its own mailbox and adds them in rank
// order -- the same order everywhere, so all ranks hold bit-identical sums.  Two slots are enough: a rank
// can run at most one reduction ahead of the slowest one.
//   ipc.ll = 1 (default): every 8-byte word carries 32 bits of payload and the 32-bit sequence tag, so a
//     value IS its own flag (two words per double) and no fence separates data from flag: one NVLink
//     traversal per reduction.  `sys_release`: the producing kernel stored into peer memory (halo push);
//     the sending threads then execute ONE fence.acq_rel.sys before the words -- cumulative over the
//     stores of all blocks, which reported in through gpu-scope fence + arrival counter -- and the
//     receivers poll with ld.acquire.sys (LDG.STRONG.SYS + CCTL.IVALL, no MEMBAR).  A MEMBAR.SYS costs
//     ~1.2 us even with nothing outstanding (measured, profiles/r01_exchange_cost.md), so reductions that
//     order no remote data (<p,Ap>: only the write-after-read hazard on the ghost planes, whose loads have
//     completed before the block reported in) skip it.
//   ipc.ll = 0 (DPP_MBOX_LL=0): values, system fence, separate flag word (two traversals).
// sm: kFinishSmem doubles of shared memory.
template <int NTH = VT>   // threads of the calling block (<= VT: the shared-memory layout is sized for VT)
__device__ __forceinline__ void finish_reduction(const double* __restrict__ partials, int nblocks, int width, double* S,
                                                 double* hist, int post, int out_offset, const IpcReduce& ipc,
                                                 double* sm, bool sys_release = true, FoldPre pre = FoldPre{nullptr, nullptr},
                                                 double* xring = nullptr) {
  double* vals = sm + VT / 32;
  volatile int* timed_out = reinterpret_cast<volatile int*>(sm + VT / 32 + kMboxEntry);
  double* recv = sm + VT / 32 + kMboxEntry + 2;   // [world][kMboxEntry - 1]
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  if (tid == 0) *timed_out = 0;
  if (width == 2) {   // both sums in one pass (the r-update epilogue sits on the critical path of every iteration)
    double v0 = 0.0, v1 = 0.0;
    for (int b = tid; b < nblocks; b += NTH) {
      const double2 pv = *reinterpret_cast<const double2*>(partials + (size_t)b * 2);
      v0 += pv.x;
      v1 += pv.y;
    }
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    const int lane = tid & 31, wid = tid >> 5;
    __syncthreads();
    if (lane == 0) {
      sm[wid] = v0;
      recv[wid] = v1;
    }
    __syncthreads();
    if (tid == 0) {
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int w = 0; w < NTH / 32; ++w) {
        t0 += sm[w];
        t1 += recv[w];
      }
      vals[0] = t0;
      vals[1] = t1;
    }
  } else {
    for (int w = 0; w < width; ++w) {
      double v = 0.0;
      for (int b = tid; b < nblocks; b += NTH) v += partials[(size_t)b * width + w];
      const double t = block_sum<NTH>(v, sm);
      if (tid == 0) vals[w] = t;
    }
  }
  const bool dist = ipc.world > 1;
  unsigned long long* seq_sm = reinterpret_cast<unsigned long long*>(sm + VT / 32 + kMboxEntry + 1);
  if (dist && tid == 0) {
    const unsigned long long sq = (pre.seq != nullptr ? *pre.seq : *ipc.seq_dev) + 1;
    *ipc.seq_dev = sq;
    *seq_sm = sq;
  }
  __syncthreads();
  const unsigned long long seq = dist ? *seq_sm : 0ull;
  const int slot = (int)(seq & 1ull);
  if (dist && tid < ipc.world) {
    const size_t mine = ((size_t)slot * ipc.world + ipc.rank) * kMboxWords;
    const size_t theirs = ((size_t)slot * ipc.world + tid) * kMboxWords;
    const long long t0 = clock64();
    if (ipc.ll) {
      const unsigned long long tag = (seq & 0xffffffffull) << 32;
      volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(ipc.peer[tid]) + mine;
      if (sys_release && ipc.ll != 2) asm volatile("fence.acq_rel.sys;" ::: "memory");
      for (int w = 0; w < width; ++w) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[w]);
        dst[2 * w] = tag | (bits & 0xffffffffull);
        dst[2 * w + 1] = tag | (bits >> 32);
      }
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(ipc.local) + theirs;
      for (int w = 0; w < width; ++w) {
        unsigned long long lo, hi;
        while (true) {
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(lo) : "l"(src + 2 * w) : "memory");
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(hi) : "l"(src + 2 * w + 1) : "memory");
          if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
          if (clock64() - t0 > 60000000000LL) {  // ~30 s: a peer died; report instead of hanging the GPU
            *timed_out = 1;
            break;
          }
        }
        recv[tid * (kMboxEntry - 1) + w] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
      }
    } else {
      const double tag = (double)seq;
      double* dst = ipc.peer[tid] + mine;
      for (int w = 0; w < width; ++w) dst[w] = vals[w];
      __threadfence_system();
      *reinterpret_cast<volatile double*>(dst + kMboxEntry - 1) = tag;
      const volatile double* src = ipc.local + theirs;
      while (src[kMboxEntry - 1] != tag) {
        if (clock64() - t0 > 60000000000LL) {
          *timed_out = 1;
          break;
        }
      }
      for (int w = 0; w < width; ++w) recv[tid * (kMboxEntry - 1) + w] = src[w];
      __threadfence_system();
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (*timed_out) {
      S[S_REASON] = (double)DPP_DIVERGED_COMM_TIMEOUT;
    } else {
      double tw[2] = {0.0, 0.0};
      for (int w = 0; w < width; ++w) {
        double t = vals[w];
        if (dist) {
          t = 0.0;
          for (int r = 0; r < ipc.world; ++r) t += recv[r * (kMboxEntry - 1) + w];
        }
        S[S_TMP + out_offset + w] = t;
        if (w < 2) tw[w] = t;
      }
      if (out_offset == 0) apply_post(S, hist, post, pre.S != nullptr ? pre.S : S, tw[0], tw[1], xring);
      else apply_post(S, hist, post);
    }
  }
}

// "last block done" hand-over: every block calls this after writing its partial sums; returns true in
// exactly one block (the last to arrive), whose threads then see all partials.  The counter resets itself.
__device__ __forceinline__ bool last_block_arrives(unsigned* counter, unsigned nblocks, int* flag_smem) {
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  __syncthreads();                       // this block's partial has been written by its thread 0
  if (tid == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(counter, 1u);
    const int last = ticket == nblocks - 1;
    if (last) {
      *counter = 0;
      __threadfence();
    }
    *flag_smem = last;
  }
  __syncthreads();
  return *flag_smem != 0;
}

}  // namespace
}  // namespace dpp
