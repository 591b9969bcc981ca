// Multi-GPU plumbing: one rank per GPU, NCCL send/recv of the skirt (halo) rows over NVLink, overlapped with
// interior compute on the compute stream.  NCCL is resolved with dlopen at ibx_comm_init so that the
// single-GPU library has no hard dependency on it (and shares torch's already-loaded libnccl.so.2 when the
// host program is a torch.distributed process).
#include "device.cuh"

#include <dlfcn.h>

using namespace ibx;

namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_t_;
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_PROD = 1, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t_*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t_) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle) return IBX_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(IBX_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
  auto sym = [&](const char* s) { return dlsym(h, s); };
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId_t*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t_*, int, ncclUniqueId_t, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(ncclComm_t_))sym("ncclCommDestroy");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, ncclComm_t_, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, ncclComm_t_, cudaStream_t))sym("ncclRecv");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t))sym("ncclAllReduce");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
      !g_nccl.GroupEnd || !g_nccl.AllReduce)
    return fail(IBX_ERR_NCCL, "libnccl is missing required symbols");
  g_nccl.handle = h;
  return IBX_OK;
}

#define NC(call)                                                                                          \
  do {                                                                                                    \
    int r_ = (call);                                                                                      \
    if (r_ != 0)                                                                                          \
      return fail(IBX_ERR_NCCL, std::string("NCCL error in " #call ": ") +                                \
                                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "unknown"));    \
  } while (0)

// pack rows idx[] of every column into a contiguous buffer (column-major, n rows)
__global__ void k_pack(const float* __restrict__ a, int64_t rows, const int32_t* __restrict__ idx, float* __restrict__ buf,
                       int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    buf[t] = a[col * rows + idx[i]];
  }
}

__global__ void k_unpack(float* __restrict__ a, int64_t rows, const int32_t* __restrict__ idx, const float* __restrict__ buf,
                         int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    a[col * rows + idx[i]] = buf[t];
  }
}

// all peers in ONE launch (blockIdx.y = peer slot): an exchange is then pack -> grouped send/recv -> unpack, three
// dependent launches instead of 2 x peers + 1 -- the exchange is latency-bound, not bandwidth-bound (8 MB at 8 ranks)
constexpr int MAXPEER = 32;
struct PeerLists {
  const int32_t* idx[MAXPEER];
  float* buf[MAXPEER];
  int64_t n[MAXPEER];
};
template <bool PACK>
__global__ void k_pack_all(float* __restrict__ a, int64_t rows, PeerLists L, int cols) {
  const int p = blockIdx.y;
  const int64_t n = L.n[p], tot = n * cols;
  const int32_t* __restrict__ idx = L.idx[p];
  float* __restrict__ buf = L.buf[p];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t % n, col = t / n;
    if (PACK) buf[t] = a[col * rows + idx[i]];
    else a[col * rows + idx[i]] = buf[t];
  }
}

}  // namespace

extern "C" {

int ibx_comm_unique_id(char id[128]) {
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t u;
  NC(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return IBX_OK;
}

int ibx_comm_init(ibx_ctx* c, int rank, int nranks, const char id[128]) {
  CHECK_CTX(c);
  int rc = load_nccl();
  if (rc) return rc;
  if (rank < 0 || rank >= nranks) return fail(IBX_ERR_ARG, "ibx_comm_init: rank out of range");
  ncclUniqueId_t u;
  memcpy(u.internal, id, 128);
  ncclComm_t_ comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, nranks, u, rank));
  c->nccl_comm = comm;
  c->rank = rank;
  c->nranks = nranks;
  return IBX_OK;
}

int ibx_comm_finalize(ibx_ctx* c) {
  if (!c) return IBX_OK;
  if (c->nccl_comm && g_nccl.CommDestroy) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    g_nccl.CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  return IBX_OK;
}

// Post the halo exchange of array `a` (rows = n_owned + n_halo of the rank-local domain): pack the rows every
// peer needs on the comm stream, grouped ncclSend/ncclRecv, unpack into the halo rows.  Compute issued on the
// compute stream after ibx_halo_begin overlaps with the transfer; ibx_halo_end makes the compute stream wait.
int ibx_halo_begin(ibx_ctx* c, const ibx_domain* d, ibx_array ah) { return ibx::halo_begin_impl(c, d, ah, true); }

}  // extern "C"

// wait_compute = false: the data to send was produced on the halo stream itself (the ghost update of the overlapped
// sharded step), so the exchange does not wait for what the compute stream is doing
int ibx::halo_begin_impl(ibx_ctx* c, const ibx_domain* d, ibx_array ah, bool wait_compute) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  if (!D.shard.active) return fail(IBX_ERR_STATE, "ibx_halo_begin: domain is not a rank-local shard (ibx_domain_shard)");
  if (!c->nccl_comm) return fail(IBX_ERR_STATE, "ibx_halo_begin: ibx_comm_init has not been called");
  GET_ARR(A, ah);
  Shard& S = D.shard;
  if (A.rows != S.n_owned + S.n_halo) return fail(IBX_ERR_ARG, "ibx_halo_begin: array must have n_owned + n_halo rows");
  int cols = (int)A.cols;
  // the comm stream must see everything the compute stream wrote into `a` so far
  if (wait_compute) {
    CU(cudaEventRecord(c->ev_ready, c->stream));
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
  }
  for (int peer = 0; peer < S.nranks; ++peer) {
    int64_t ns = (int64_t)S.send_local[peer].size(), nr = (int64_t)S.recv_local[peer].size();
    int64_t need = std::max(ns, nr) * cols;
    if (need > S.buf_cap[peer]) {
      if (S.d_sendbuf[peer]) cudaFree(S.d_sendbuf[peer]);
      if (S.d_recvbuf[peer]) cudaFree(S.d_recvbuf[peer]);
      CU(cudaMalloc((void**)&S.d_sendbuf[peer], (size_t)need * sizeof(float)));
      CU(cudaMalloc((void**)&S.d_recvbuf[peer], (size_t)need * sizeof(float)));
      S.buf_cap[peer] = need;
    }
    if (ns && S.nranks > MAXPEER) {
      k_pack<<<grid_for(ns * cols, 256, c->sm_count, 4), 256, 0, c->comm_stream>>>(A.p, A.rows, S.d_send[peer], S.d_sendbuf[peer], ns, cols);
      LAUNCH_CHECK();
    }
  }
  if (S.nranks <= MAXPEER) {
    PeerLists L;
    int np = 0;
    int64_t nmax = 0;
    for (int peer = 0; peer < S.nranks; ++peer) {
      int64_t ns = (int64_t)S.send_local[peer].size();
      if (!ns) continue;
      L.idx[np] = S.d_send[peer]; L.buf[np] = S.d_sendbuf[peer]; L.n[np] = ns;
      nmax = std::max(nmax, ns);
      ++np;
    }
    if (np) {
      dim3 grid(grid_for(nmax * cols, 256, c->sm_count, 2), np);
      k_pack_all<true><<<grid, 256, 0, c->comm_stream>>>(A.p, A.rows, L, cols);
      LAUNCH_CHECK();
    }
  }
  NC(g_nccl.GroupStart());
  {
    // an error between GroupStart and GroupEnd must not leave the group open (the next NCCL call would hang): close it,
    // poison the context, then report
    int bad = 0;
    for (int peer = 0; peer < S.nranks && !bad; ++peer) {
      int64_t ns = (int64_t)S.send_local[peer].size(), nr = (int64_t)S.recv_local[peer].size();
      if (ns) bad = g_nccl.Send(S.d_sendbuf[peer], (size_t)ns * cols, NCCL_FLOAT32, peer, c->nccl_comm, c->comm_stream);
      if (nr && !bad) bad = g_nccl.Recv(S.d_recvbuf[peer], (size_t)nr * cols, NCCL_FLOAT32, peer, c->nccl_comm, c->comm_stream);
    }
    int end = g_nccl.GroupEnd();
    if (bad || end) {
      c->poisoned = true;
      return fail(IBX_ERR_NCCL, std::string("NCCL error in the grouped halo send/recv: ") +
                                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(bad ? bad : end) : "unknown"));
    }
  }
  for (int peer = 0; peer < S.nranks; ++peer) {
    int64_t nr = (int64_t)S.recv_local[peer].size();
    if (nr && S.nranks > MAXPEER) {
      k_unpack<<<grid_for(nr * cols, 256, c->sm_count, 4), 256, 0, c->comm_stream>>>(A.p, A.rows, S.d_recv[peer], S.d_recvbuf[peer], nr, cols);
      LAUNCH_CHECK();
    }
  }
  if (S.nranks <= MAXPEER) {
    PeerLists L;
    int np = 0;
    int64_t nmax = 0;
    for (int peer = 0; peer < S.nranks; ++peer) {
      int64_t nr = (int64_t)S.recv_local[peer].size();
      if (!nr) continue;
      L.idx[np] = S.d_recv[peer]; L.buf[np] = S.d_recvbuf[peer]; L.n[np] = nr;
      nmax = std::max(nmax, nr);
      ++np;
    }
    if (np) {
      dim3 grid(grid_for(nmax * cols, 256, c->sm_count, 2), np);
      k_pack_all<false><<<grid, 256, 0, c->comm_stream>>>(A.p, A.rows, L, cols);
      LAUNCH_CHECK();
    }
  }
  CU(cudaEventRecord(c->ev_halo, c->comm_stream));
  c->halo_pending = ah;
  return IBX_OK;
}

extern "C" {

int ibx_halo_end(ibx_ctx* c, const ibx_domain* d, ibx_array a) {
  CHECK_CTX(c);
  (void)d;
  (void)a;
  CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  c->halo_pending = 0;
  return IBX_OK;
}

int ibx_allreduce(ibx_ctx* c, int op, double* inout, int n) {
  CHECK_CTX(c);
  if (c->nranks == 1 || !c->nccl_comm) return IBX_OK;
  if (n < 1 || n > c->red_cap) return fail(IBX_ERR_ARG, "ibx_allreduce: n out of range");
  if (op < 0 || op > 2) return fail(IBX_ERR_ARG, "ibx_allreduce: op must be 0 sum, 1 max, 2 min");
  CU(cudaMemcpyAsync(c->d_red, inout, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int nop = op == 0 ? NCCL_SUM : (op == 1 ? NCCL_MAX : NCCL_MIN);
  NC(g_nccl.AllReduce(c->d_red, c->d_red, (size_t)n, NCCL_FLOAT64, nop, c->nccl_comm, c->stream));
  CU(cudaMemcpyAsync(inout, c->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IBX_OK;
}

}  // extern "C"
