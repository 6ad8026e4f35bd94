// Slab-partition communication: NCCL over NVLink 5 / NVSwitch.
//   - halo exchange before an apply: each rank sends the node planes its neighbours' boundary cells
//     touch (grouped ncclSend/ncclRecv, <= 2 peers for slabs; SURVEY 8e "forward halo");
//   - ncclAllReduce(sum, fp64) of the 1..31 Krylov scalars.
// Replaces PETSc VecScatter + MPI_Allreduce of an mpiexec run of the reference.
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dpp_internal.cuh"

namespace dpp {

// NCCL is resolved with dlopen on first use instead of a link-time dependency: a single-GPU
// process never loads it, and a process that already holds PyTorch's bundled libnccl.so.2 reuses
// that copy (RTLD_NOLOAD) instead of pulling a second, older one into the namespace.
struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  if (api.ok || !api.error.empty()) return api;
  const char* env = getenv("DPP_NCCL_LIBRARY");
  if (env && *env) api.so = dlopen(env, RTLD_NOW | RTLD_LOCAL);
  if (!api.so) api.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!api.so) api.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!api.so) {
    api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
    return api;
  }
#define DPP_SYM(field, name)                                              \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.so, name)); \
  if (!api.field) { api.error = std::string("libnccl.so.2 lacks ") + name; return api; }
  DPP_SYM(GetUniqueId, "ncclGetUniqueId")
  DPP_SYM(CommInitRank, "ncclCommInitRank")
  DPP_SYM(CommDestroy, "ncclCommDestroy")
  DPP_SYM(Send, "ncclSend")
  DPP_SYM(Recv, "ncclRecv")
  DPP_SYM(AllReduce, "ncclAllReduce")
  DPP_SYM(GroupStart, "ncclGroupStart")
  DPP_SYM(GroupEnd, "ncclGroupEnd")
  DPP_SYM(GetErrorString, "ncclGetErrorString")
#undef DPP_SYM
  api.ok = true;
  return api;
}

struct Neighbor {
  int peer = -1;
  int64_t n_send = 0, n_recv = 0;
  int32_t *d_send_idx = nullptr, *d_recv_idx = nullptr;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;  // [2 * n]
  bool send_contig = false, recv_contig = false;
  int64_t send0 = 0, recv0 = 0;
};

struct IpcBlob {  // what every rank publishes (POD; all-gathered by the host layer)
  cudaIpcMemHandle_t r_handle;
  cudaIpcMemHandle_t mbox_handle;
  cudaIpcMemHandle_t box_handle;   // halo inbox (generic vectors; see comm_halo_exchange)
  int64_t field, plane;      // padded strides of the residual vector (doubles)
  int64_t box_cap;           // doubles per (parity, side) section of the inbox
  int32_t i_begin, i_end;    // owned local planes
  int32_t rank, valid;       // valid: bit 0 mailbox handle, bit 1 residual-vector handle, bit 2 halo inbox
};

struct Comm {
  ncclComm_t comm = nullptr;
  std::vector<Neighbor> nbrs;
  // peer-memory fast path
  bool ipc = false;                       // mailbox all-reduce active
  bool ipc_halo = false;                  // halo push into the neighbours' residual vectors active
  double* mbox = nullptr;                 // own mailbox [2][world][kMboxWords]
  double* mbox_peer[kMaxIpcRanks] = {};   // mapped mailboxes of the other ranks
  double* r_peer[2] = {nullptr, nullptr}; // mapped residual vectors of rank-1 / rank+1
  long long r_peer_field[2] = {0, 0}, r_peer_ghost_off[2] = {0, 0};
  // halo inbox: neighbours store their boundary planes of ANY vector here (peer writes over NVLink), then a flag
  bool ipc_box = false;
  double* box = nullptr;                  // own inbox: [2 parities][2 sides][box_cap] + flags
  long long box_cap = 0;
  double* box_peer[2] = {nullptr, nullptr};   // mapped inbox of rank-1 / rank+1
  long long box_peer_cap[2] = {0, 0};
  unsigned long long* d_hseq = nullptr;   // device: number of halo exchanges executed, arrival counter, error flag
  std::vector<void*> mapped;              // everything opened with cudaIpcOpenMemHandle
};

namespace {

#define DPP_NCCL(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r__ = (call);                                                           \
    if (r__ != ncclSuccess) {                                                            \
      ctx->set_error(std::string(#call) + ": " + nccl().GetErrorString(r__));            \
      return DPP_ERR_NCCL;                                                               \
    }                                                                                    \
  } while (0)

__global__ void k_pack(long long n, const int32_t* __restrict__ idx, const double* __restrict__ f0,
                       const double* __restrict__ f1, double* __restrict__ buf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long s = idx[i];
    buf[i] = f0[s];
    if (f1 != nullptr) buf[n + i] = f1[s];
  }
}

__global__ void k_unpack(long long n, const int32_t* __restrict__ idx, double* __restrict__ f0, double* __restrict__ f1,
                         const double* __restrict__ buf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long s = idx[i];
    f0[s] = buf[i];
    if (f1 != nullptr) f1[s] = buf[n + i];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Peer-memory halo exchange for ANY field-blocked vector (the unfused Krylov paths, Q2, GMRES, Picard).  Every
// rank owns an inbox [2 parities][2 sides][cap]; exchange number s (device counter, the same on all ranks because
// all ranks execute the same sequence of exchanges) works on parity s & 1:
//   k_halo_put   stores this rank's boundary planes into the neighbours' inboxes (NVLink peer writes); every block
//                reports in through a gpu-scope fence + arrival counter, the last block executes ONE
//                fence.acq_rel.sys (cumulative over all blocks' stores) and then writes s into the neighbours'
//                flag words;
//   k_halo_get   polls its own flag words (ld.acquire.sys) until both neighbours' planes of exchange s have landed,
//                then copies them from the inbox into the ghost planes of the vector.
// Two parities suffice: a neighbour can start exchange s + 1 before this rank has emptied section s (other parity),
// but exchange s + 2 needs this rank's data of s + 1, which this rank sends after it emptied s (stream order).
// Replaces a grouped ncclSend/ncclRecv (~35 us per exchange on two B200s) by two small kernels.
// ------------------------------------------------------------------------------------------------------------
struct HaloSide {
  double* peer_data;                 // neighbour's inbox section base for my side (parity 0), null: no neighbour
  unsigned long long* peer_flag;     // neighbour's flag word for my side
  const double* own_data;            // my inbox section base for that neighbour's side (parity 0)
  const unsigned long long* own_flag;
  long long peer_par_stride, own_par_stride;   // doubles between the two parities
  long long n_send, n_recv, send0, recv0;
  const int32_t* send_idx;           // null: contiguous from send0
  const int32_t* recv_idx;
};
struct HaloArgs {
  HaloSide side[2];
  double* f[2];
  int nf;
  unsigned long long* hseq;          // [0] exchanges done, [1] arrival counter, [2] error flag
};

__global__ void __launch_bounds__(256) k_halo_put(const HaloArgs a) {
  const unsigned long long seq = a.hseq[0] + 1;
  const int par = (int)(seq & 1ull);
  for (int s = 0; s < 2; ++s) {
    const HaloSide& h = a.side[s];
    if (h.peer_data == nullptr) continue;
    double* dst = h.peer_data + par * h.peer_par_stride;
    for (int f = 0; f < a.nf; ++f)
      for (long long i = blockIdx.x * 256ll + threadIdx.x; i < h.n_send; i += gridDim.x * 256ll)
        dst[f * h.n_send + i] = a.f[f][h.send_idx != nullptr ? (long long)h.send_idx[i] : h.send0 + i];
  }
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(a.hseq + 1), 1u);
    last = t == gridDim.x - 1;
    if (last) {
      *reinterpret_cast<unsigned*>(a.hseq + 1) = 0u;
      __threadfence();
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      for (int s = 0; s < 2; ++s)
        if (a.side[s].peer_flag != nullptr) *reinterpret_cast<volatile unsigned long long*>(a.side[s].peer_flag) = seq;
      a.hseq[0] = seq;
    }
  }
}

__global__ void __launch_bounds__(256) k_halo_get(const HaloArgs a) {
  const unsigned long long seq = a.hseq[0];   // written by k_halo_put of this exchange (stream order)
  const int par = (int)(seq & 1ull);
  __shared__ int bad;
  if (threadIdx.x == 0) {
    bad = 0;
    const long long t0 = clock64();
    for (int s = 0; s < 2; ++s) {
      if (a.side[s].own_flag == nullptr) continue;
      unsigned long long v;
      while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.side[s].own_flag) : "memory");
        if (v >= seq) break;
        if (clock64() - t0 > 60000000000LL) {   // ~30 s: a peer died; report instead of hanging the GPU
          bad = 1;
          a.hseq[2] = 1ull;
          break;
        }
      }
    }
  }
  __syncthreads();
  if (bad) return;
  for (int s = 0; s < 2; ++s) {
    const HaloSide& h = a.side[s];
    if (h.own_data == nullptr) continue;
    const double* src = h.own_data + par * h.own_par_stride;
    for (int f = 0; f < a.nf; ++f)
      for (long long i = blockIdx.x * 256ll + threadIdx.x; i < h.n_recv; i += gridDim.x * 256ll) {
        double v;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + f * h.n_recv + i) : "memory");
        a.f[f][h.recv_idx != nullptr ? (long long)h.recv_idx[i] : h.recv0 + i] = v;
      }
  }
}

bool contiguous(const int32_t* v, int64_t n) {
  for (int64_t i = 1; i < n; ++i)
    if (v[i] != v[0] + i) return false;
  return n > 0;
}

}  // namespace

int comm_halo_exchange(dpp_context* ctx, double* const* fields, int nf) {
  Comm* C = ctx->comm;
  if (!C || ctx->world <= 1) return DPP_OK;
  if (C->ipc && C->ipc_box) {
    HaloArgs a{};
    a.nf = nf;
    a.f[0] = fields[0];
    a.f[1] = nf == 2 ? fields[1] : nullptr;
    a.hseq = C->d_hseq;
    long long work = 1;
    for (Neighbor& nb : C->nbrs) {
      const int s = nb.peer < ctx->rank ? 0 : 1;        // the neighbour below / above me
      HaloSide& h = a.side[s];
      // in the neighbour's inbox I am its upper (s == 0) / lower (s == 1) side
      const int my_side_there = s == 0 ? 1 : 0;
      const long long pcap = C->box_peer_cap[s];
      if ((long long)nf * nb.n_send > pcap || (long long)nf * nb.n_recv > C->box_cap) {
        ctx->set_error("halo inbox too small for this exchange");
        return DPP_ERR_INVALID;
      }
      h.peer_data = C->box_peer[s] + my_side_there * pcap;
      h.peer_par_stride = 2 * pcap;
      h.peer_flag = reinterpret_cast<unsigned long long*>(C->box_peer[s] + 4 * pcap) + my_side_there;
      h.own_data = C->box + s * C->box_cap;
      h.own_par_stride = 2 * C->box_cap;
      h.own_flag = reinterpret_cast<unsigned long long*>(C->box + 4 * C->box_cap) + s;
      h.n_send = nb.n_send; h.n_recv = nb.n_recv; h.send0 = nb.send0; h.recv0 = nb.recv0;
      h.send_idx = nb.send_contig ? nullptr : nb.d_send_idx;
      h.recv_idx = nb.recv_contig ? nullptr : nb.d_recv_idx;
      work = std::max<long long>(work, std::max(nb.n_send, nb.n_recv));
    }
    const int blocks = (int)std::max<long long>(1, std::min<long long>((work + 255) / 256, 2LL * ctx->sm_count));
    k_halo_put<<<blocks, 256, 0, ctx->stream>>>(a);
    k_halo_get<<<blocks, 256, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    DPP_CUDA(cudaGetLastError());
    return DPP_OK;
  }
  for (Neighbor& nb : C->nbrs) {
    if (nb.n_send > 0 && !nb.send_contig) {
      const int blocks = (int)std::min<int64_t>((nb.n_send + 255) / 256, 1024);
      k_pack<<<blocks, 256, 0, ctx->stream>>>(nb.n_send, nb.d_send_idx, fields[0], nf == 2 ? fields[1] : nullptr, nb.d_sendbuf);
      ctx->launches++;
    }
  }
  DPP_NCCL(nccl().GroupStart());
  for (Neighbor& nb : C->nbrs) {
    for (int f = 0; f < nf; ++f) {
      if (nb.n_send > 0) {
        const double* src = nb.send_contig ? fields[f] + nb.send0 : nb.d_sendbuf + f * nb.n_send;
        DPP_NCCL(nccl().Send(src, (size_t)nb.n_send, ncclDouble, nb.peer, C->comm, ctx->stream));
      }
      if (nb.n_recv > 0) {
        double* dst = nb.recv_contig ? fields[f] + nb.recv0 : nb.d_recvbuf + f * nb.n_recv;
        DPP_NCCL(nccl().Recv(dst, (size_t)nb.n_recv, ncclDouble, nb.peer, C->comm, ctx->stream));
      }
    }
  }
  DPP_NCCL(nccl().GroupEnd());
  for (Neighbor& nb : C->nbrs) {
    if (nb.n_recv > 0 && !nb.recv_contig) {
      const int blocks = (int)std::min<int64_t>((nb.n_recv + 255) / 256, 1024);
      k_unpack<<<blocks, 256, 0, ctx->stream>>>(nb.n_recv, nb.d_recv_idx, fields[0], nf == 2 ? fields[1] : nullptr, nb.d_recvbuf);
      ctx->launches++;
    }
  }
  DPP_CUDA(cudaGetLastError());
  return DPP_OK;
}

int comm_halo_planes(dpp_context* ctx, double* base, int nf, long long field_stride, long long plane_elems, int i_begin,
                     int i_end) {
  Comm* C = ctx->comm;
  if (!C || ctx->world <= 1) return DPP_OK;
  DPP_NCCL(nccl().GroupStart());
  for (Neighbor& nb : C->nbrs) {
    const bool lower = nb.peer < ctx->rank;
    if (lower ? i_begin <= 0 : false) continue;
    // lower neighbour: send my first owned plane, receive my lower ghost plane; upper: last owned / upper ghost
    const long long send_pl = lower ? i_begin : i_end - 1;
    const long long recv_pl = lower ? i_begin - 1 : i_end;
    for (int f = 0; f < nf; ++f) {
      double* fb = base + f * field_stride;
      DPP_NCCL(nccl().Send(fb + send_pl * plane_elems, (size_t)plane_elems, ncclDouble, nb.peer, C->comm, ctx->stream));
      DPP_NCCL(nccl().Recv(fb + recv_pl * plane_elems, (size_t)plane_elems, ncclDouble, nb.peer, C->comm, ctx->stream));
    }
  }
  DPP_NCCL(nccl().GroupEnd());
  return DPP_OK;
}

int comm_allreduce_sum(dpp_context* ctx, double* d_vals, int n) {
  Comm* C = ctx->comm;
  if (!C || ctx->world <= 1) return DPP_OK;
  DPP_NCCL(nccl().AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclSum, C->comm, ctx->stream));
  return DPP_OK;
}

bool comm_ipc_ready(const dpp_context* ctx) { return ctx->comm != nullptr && ctx->comm->ipc; }
bool comm_ipc_halo_ready(const dpp_context* ctx) { return ctx->comm != nullptr && ctx->comm->ipc && ctx->comm->ipc_halo; }
bool comm_ipc_box_ready(const dpp_context* ctx) { return ctx->comm != nullptr && ctx->comm->ipc && ctx->comm->ipc_box; }
// 1 when a peer never delivered a halo plane (k_halo_get timed out); synchronises the stream
int comm_halo_failed(dpp_context* ctx) {
  Comm* C = ctx->comm;
  if (!C || !C->d_hseq || !C->ipc_box) return 0;
  unsigned long long v = 0;
  if (cudaMemcpyAsync(&v, C->d_hseq + 2, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
  cudaStreamSynchronize(ctx->stream);
  return v != 0 ? 1 : 0;
}

// Measurement switches that deliberately break the protocol (profiles/r01_exchange_cost.md) exist only in
// builds made with -DDPP_MEASUREMENT (make MEASUREMENT=1); the release library ignores the variables, so
// no benchmark number can be taken with wrong numerics by accident.
#ifdef DPP_MEASUREMENT
static void warn_measurement_switches() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* ll = getenv("DPP_MBOX_LL");
  if (getenv("DPP_DEBUG_NO_MBOX") || getenv("DPP_DEBUG_NO_PUSH") || (ll != nullptr && atoi(ll) == 2))
    fprintf(stderr, "libdppb200 (DPP_MEASUREMENT build): DPP_DEBUG_NO_MBOX / DPP_DEBUG_NO_PUSH / DPP_MBOX_LL=2 are "
                    "timing experiments: the multi-GPU RESULTS OF THIS RUN ARE INVALID\n");
}
static bool meas_env(const char* name) { return getenv(name) != nullptr; }
#else
static void warn_measurement_switches() {}
static bool meas_env(const char*) { return false; }
#endif

IpcReduce comm_ipc_reduce_args(dpp_context* ctx) {
  warn_measurement_switches();
  Comm* C = ctx->comm;
  IpcReduce a{};
  a.world = 1;
  if (!C || !C->ipc) return a;
  a.local = C->mbox;
  for (int r = 0; r < ctx->world; ++r) a.peer[r] = (r == ctx->rank) ? C->mbox : C->mbox_peer[r];
  a.rank = ctx->rank;
  a.world = meas_env("DPP_DEBUG_NO_MBOX") ? 1 : ctx->world;   // timing experiments only: local sums
  // the sequence counter lives behind the mailbox entries; it advances only when an exchange really runs
  a.seq_dev = reinterpret_cast<unsigned long long*>(C->mbox + 2 * kMaxIpcRanks * kMboxWords);
  a.ll = 1;   // tagged words (cg_device.cuh); 0 = values + fence + flag word (valid, slower), 3 = fence for <p,Ap> too
#ifdef DPP_MEASUREMENT
  if (const char* ll = getenv("DPP_MBOX_LL")) a.ll = atoi(ll);
#else
  if (const char* ll = getenv("DPP_MBOX_LL")) {   // only the protocol-correct variants are selectable
    const int v = atoi(ll);
    if (v == 0 || v == 1 || v == 3) a.ll = v;
  }
#endif
  return a;
}

IpcHalo comm_ipc_halo(const dpp_context* ctx) {
  IpcHalo h{};
  const Comm* C = ctx->comm;
  // DPP_DEBUG_NO_PUSH (measurement builds only): the solve is wrong without the ghost planes
  if (C && C->ipc && C->ipc_halo && !meas_env("DPP_DEBUG_NO_PUSH"))
    for (int s = 0; s < 2; ++s) {
      h.peer_r[s] = C->r_peer[s];
      h.peer_field[s] = C->r_peer_field[s];
      h.peer_ghost_off[s] = C->r_peer_ghost_off[s];
    }
  h.debug_fence_all = meas_env("DPP_DEBUG_FENCE_ALL");
  return h;
}

void comm_destroy(dpp_context* ctx) {
  Comm* C = ctx->comm;
  if (!C) return;
  for (void* m : C->mapped) cudaIpcCloseMemHandle(m);
  if (C->mbox) cudaFree(C->mbox);
  if (C->box) cudaFree(C->box);
  if (C->d_hseq) cudaFree(C->d_hseq);
  for (Neighbor& nb : C->nbrs) {
    void* p[] = {nb.d_send_idx, nb.d_recv_idx, nb.d_sendbuf, nb.d_recvbuf};
    for (void* q : p)
      if (q) cudaFree(q);
  }
  if (C->comm && nccl().ok) nccl().CommDestroy(C->comm);
  delete C;
  ctx->comm = nullptr;
}

}  // namespace dpp

extern "C" {

int dpp_nccl_unique_id(void* out128) {
  if (!out128) return DPP_ERR_INVALID;
  ncclUniqueId id;
  if (!dpp::nccl().ok || dpp::nccl().GetUniqueId(&id) != ncclSuccess) return DPP_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return DPP_OK;
}

int dpp_comm_init(dpp_handle ctx, int rank, int world, const void* unique_id, int64_t owned_begin, int64_t owned_end) {
  if (!ctx || world < 1 || rank < 0 || rank >= world || owned_begin < 0 || owned_end > ctx->n_nodes || owned_begin > owned_end) {
    if (ctx) ctx->set_error("dpp_comm_init: invalid argument");
    return DPP_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  dpp::comm_destroy(ctx);
  ctx->rank = rank;
  ctx->world = world;
  ctx->owned_begin = owned_begin;
  ctx->owned_end = owned_end;
  ctx->dom_lo = owned_begin == 0;            // a stored plane that is not owned is a ghost plane
  ctx->dom_hi = owned_end == ctx->n_nodes;
  ctx->invalidate();
  if (world == 1) return DPP_OK;
  if (!unique_id) {
    ctx->set_error("dpp_comm_init: nccl_unique_id required for world > 1");
    return DPP_ERR_INVALID;
  }
  if (!dpp::nccl().ok) {
    ctx->set_error(dpp::nccl().error);
    return DPP_ERR_NCCL;
  }
  using dpp::nccl;
  ctx->comm = new dpp::Comm();
  ncclUniqueId id;
  memcpy(&id, unique_id, 128);
  DPP_NCCL(nccl().CommInitRank(&ctx->comm->comm, world, id, rank));
  return DPP_OK;
}

int dpp_comm_ipc_blob_size(void) { return (int)sizeof(dpp::IpcBlob); }

int dpp_comm_ipc_export(dpp_handle ctx, void* blob_out) {
  if (!ctx || !blob_out) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  dpp::IpcBlob b;
  memset(&b, 0, sizeof(b));
  b.rank = ctx->rank;
  dpp::Comm* C = ctx->comm;
  if (C && ctx->world > 1 && ctx->world <= dpp::kMaxIpcRanks && !getenv("DPP_NO_IPC")) {
    if (!C->mbox) {
      DPP_CHECK(dpp::dev_alloc(ctx, &C->mbox, (int64_t)dpp::kMboxDoubles));
      DPP_CUDA(cudaMemset(C->mbox, 0, sizeof(double) * dpp::kMboxDoubles));
      DPP_CUDA(cudaDeviceSynchronize());   // the zeros are in place before any peer can learn the handle
    }
    if (cudaIpcGetMemHandle(&b.mbox_handle, C->mbox) == cudaSuccess) b.valid |= 1;
    else cudaGetLastError();
    if ((b.valid & 1) && !C->nbrs.empty() && !getenv("DPP_NO_IPC_BOX")) {   // halo inbox for generic vectors
      long long cap = 0;
      for (const dpp::Neighbor& nb : C->nbrs) cap = std::max<long long>(cap, 2 * std::max(nb.n_send, nb.n_recv));
      cap = (cap + 1) & ~1ll;
      if (!C->box || C->box_cap != cap) {
        if (C->box) cudaFree(C->box);
        C->box = nullptr;
        C->box_cap = cap;
        DPP_CHECK(dpp::dev_alloc(ctx, &C->box, 4 * cap + 2));
        DPP_CUDA(cudaMemset(C->box, 0, sizeof(double) * (4 * cap + 2)));
        if (!C->d_hseq) DPP_CHECK(dpp::dev_alloc(ctx, &C->d_hseq, 4));
        DPP_CUDA(cudaMemset(C->d_hseq, 0, sizeof(unsigned long long) * 4));
        DPP_CUDA(cudaDeviceSynchronize());
      }
      if (cudaIpcGetMemHandle(&b.box_handle, C->box) == cudaSuccess) {
        b.box_cap = cap;
        b.valid |= 4;
      } else {
        cudaGetLastError();
      }
    }
    long long field = 0, plane = 0;
    double* r = dpp::cg_fused_r_buffer(ctx, &field, &plane);   // only the fused (uniform-grid) CG paths have one
    if (r != nullptr && (b.valid & 1)) {
      const long long uplane = (long long)ctx->grid.n[1] * ctx->grid.n[2];
      if (cudaIpcGetMemHandle(&b.r_handle, r) == cudaSuccess) {
        b.field = field;
        b.plane = plane;
        b.i_begin = (int32_t)(ctx->owned_begin / uplane);
        b.i_end = (int32_t)(ctx->owned_end / uplane);
        b.valid |= 2;
      } else {
        cudaGetLastError();
      }
    }
  }
  memcpy(blob_out, &b, sizeof(b));
  return DPP_OK;
}

int dpp_comm_ipc_import(dpp_handle ctx, const void* blobs) {
  if (!ctx || !blobs || !ctx->comm) return DPP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  dpp::Comm* C = ctx->comm;
  C->ipc = false;
  C->ipc_halo = false;
  const dpp::IpcBlob* B = static_cast<const dpp::IpcBlob*>(blobs);
  if (ctx->world > dpp::kMaxIpcRanks) return DPP_OK;
  bool all_r = true;
  for (int r = 0; r < ctx->world; ++r) {
    if (!(B[r].valid & 1) || B[r].rank != r) return DPP_OK;  // some rank cannot take part: keep the NCCL path
    all_r = all_r && (B[r].valid & 2);
  }
  for (int r = 0; r < ctx->world; ++r) {
    if (r == ctx->rank) continue;
    void* m = nullptr;
    if (cudaIpcOpenMemHandle(&m, B[r].mbox_handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return DPP_OK;  // no peer access between these GPUs: NCCL path
    }
    C->mapped.push_back(m);
    C->mbox_peer[r] = static_cast<double*>(m);
  }
  C->ipc = true;
  // halo inboxes of the two neighbours (all ranks must have one: the protocol is the same everywhere)
  bool all_box = true;
  for (int r = 0; r < ctx->world; ++r) all_box = all_box && (B[r].valid & 4);
  C->ipc_box = false;
  if (all_box) {
    bool ok = true;
    for (int s = 0; s < 2 && ok; ++s) {
      const int peer = s == 0 ? ctx->rank - 1 : ctx->rank + 1;
      C->box_peer[s] = nullptr;
      if (peer < 0 || peer >= ctx->world) continue;
      void* m = nullptr;
      if (cudaIpcOpenMemHandle(&m, B[peer].box_handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      C->mapped.push_back(m);
      C->box_peer[s] = static_cast<double*>(m);
      C->box_peer_cap[s] = B[peer].box_cap;
    }
    C->ipc_box = ok;
  }
  for (int s = 0; s < 2 && all_r; ++s) {
    const int peer = s == 0 ? ctx->rank - 1 : ctx->rank + 1;
    if (peer < 0 || peer >= ctx->world) continue;
    void* m = nullptr;
    if (cudaIpcOpenMemHandle(&m, B[peer].r_handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return DPP_OK;
    }
    C->mapped.push_back(m);
    C->r_peer[s] = static_cast<double*>(m);
    C->r_peer_field[s] = B[peer].field;
    // my first owned plane is the lower neighbour's upper ghost (its local plane i_end);
    // my last `degree` owned planes are the upper neighbour's lower ghosts (its local planes i_begin - degree ..)
    const int band = ctx->family == DPP_KERNEL_STRUCTURED ? std::max(1, ctx->grid.band) : 1;
    C->r_peer_ghost_off[s] = (s == 0 ? (long long)B[peer].i_end : (long long)B[peer].i_begin - band) * B[peer].plane;
    if (C->r_peer_ghost_off[s] < 0) return DPP_OK;
  }
  C->ipc_halo = all_r;
  return DPP_OK;
}

int dpp_comm_ipc_disable(dpp_handle ctx) {
  if (!ctx) return DPP_ERR_INVALID;
  if (ctx->comm) ctx->comm->ipc = ctx->comm->ipc_halo = ctx->comm->ipc_box = false;
  return DPP_OK;
}

int dpp_comm_add_neighbor(dpp_handle ctx, int peer, int64_t n_send, const int32_t* send_nodes, int64_t n_recv,
                          const int32_t* recv_nodes) {
  if (!ctx || !ctx->comm || peer < 0 || peer >= ctx->world || peer == ctx->rank) {
    if (ctx) ctx->set_error("dpp_comm_add_neighbor: invalid argument or dpp_comm_init not called");
    return DPP_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  dpp::Neighbor nb;
  nb.peer = peer;
  nb.n_send = n_send;
  nb.n_recv = n_recv;
  if (n_send > 0) {
    nb.send_contig = dpp::contiguous(send_nodes, n_send);
    nb.send0 = send_nodes[0];
    DPP_CHECK(dpp::dev_alloc(ctx, &nb.d_send_idx, n_send));
    DPP_CHECK(dpp::dev_alloc(ctx, &nb.d_sendbuf, 2 * n_send));
    DPP_CUDA(cudaMemcpy(nb.d_send_idx, send_nodes, sizeof(int32_t) * n_send, cudaMemcpyHostToDevice));
  }
  if (n_recv > 0) {
    nb.recv_contig = dpp::contiguous(recv_nodes, n_recv);
    nb.recv0 = recv_nodes[0];
    DPP_CHECK(dpp::dev_alloc(ctx, &nb.d_recv_idx, n_recv));
    DPP_CHECK(dpp::dev_alloc(ctx, &nb.d_recvbuf, 2 * n_recv));
    DPP_CUDA(cudaMemcpy(nb.d_recv_idx, recv_nodes, sizeof(int32_t) * n_recv, cudaMemcpyHostToDevice));
  }
  ctx->comm->nbrs.push_back(nb);
  return DPP_OK;
}

}  // extern "C"
