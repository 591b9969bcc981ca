// Device runtime of libibx.so: context, opaque float32 arrays, one-time upload of the Domain tables.
// The upload replaces the reference's per-call `to_backend(part, conv_to_backend)`
// (src/ImmersedBoundary.jl:846-849, src/arraybends.jl:14-77).
#include "device.cuh"

namespace ibx {

int cuda_fail(ibx_ctx* c, cudaError_t e, const char* what, const char* file, int line) {
  if (c && e != cudaErrorMemoryAllocation && e != cudaErrorInvalidValue) c->poisoned = true;
  return fail(IBX_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
                                std::to_string(line) + ")");
}

bool get_array(ibx_ctx* c, ibx_array h, ibx_ctx::Arr& out) {
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->arrays.find(h);
  if (it == c->arrays.end()) return false;
  out = it->second;
  return true;
}

float* ensure_scratch(ibx_ctx* c, int64_t nfloats) {
  if (nfloats <= c->scratch_cap) return c->d_scratch;
  if (c->d_scratch) cudaFree(c->d_scratch);
  c->d_scratch = nullptr;
  c->scratch_cap = 0;
  if (cudaMalloc((void**)&c->d_scratch, (size_t)nfloats * sizeof(float)) != cudaSuccess) return nullptr;
  c->scratch_cap = nfloats;
  return c->d_scratch;
}

static void free_tables(ibx_domain& D) {
  auto fr = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
  for (auto& P : D.parts) {
    fr(P.d_domain); fr(P.d_image_in_domain); fr(P.d_spacing); fr(P.d_centers);
    for (auto& T : P.dims) { fr(T.d_owners); fr(T.d_neighbors); fr(T.d_lptr); fr(T.d_lidx); fr(T.d_rptr); fr(T.d_ridx); }
  }
  for (auto& F : D.boundaries)
    for (auto& B : F.parts) { fr(B.d_ghost); fr(B.d_ptr); fr(B.d_idx_global); fr(B.d_image_domain); fr(B.d_idx); fr(B.d_w); fr(B.d_normals); fr(B.d_eta); }
  for (auto& S : D.surfaces) fr(S.d_areas);
  fr(D.d_block_faces); fr(D.d_block_h);
  for (auto& L : D.phase)
    for (auto& q : L.d) fr(q);
  D.phased = false;
  fr(D.d_blk_noghost); fr(D.d_blk_ghost);
  fr(D.d_blk_all_plain); fr(D.d_blk_all_finer); fr(D.d_blk_own_plain); fr(D.d_blk_own_finer); fr(D.d_blk_own_regular); fr(D.d_blk_all_regular);
  for (auto& p : D.shard.d_send) fr(p);
  for (auto& p : D.shard.d_recv) fr(p);
  for (auto& p : D.shard.d_sendbuf) fr(p);
  for (auto& p : D.shard.d_recvbuf) fr(p);
  D.uploaded = false;
}

}  // namespace ibx

ibx_domain::~ibx_domain() {
  if (uploaded) { cudaSetDevice(device); ibx::free_tables(*this); }
}

ibx_accum::~ibx_accum() {
  if (uploaded) { cudaFree(d_ptr); cudaFree(d_idx); if (d_w) cudaFree(d_w); }
}

using namespace ibx;

extern "C" {

int ibx_init(int device, ibx_ctx** out) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(IBX_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                  "); libibx has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(IBX_ERR_ARG, "ibx_init: device index out of range");
  cudaDeviceProp prop;
  ibx_ctx* c = new ibx_ctx();
  c->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete c;
    return fail(IBX_ERR_CUDA, std::string("ibx_init: ") + cudaGetErrorString(e));
  }
  if (prop.major < 10) {
    delete c;
    return fail(IBX_ERR_UNSUPPORTED, "ibx_init: libibx is built for sm_100a only; device is sm_" +
                                         std::to_string(prop.major) + std::to_string(prop.minor));
  }
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->total_mem = prop.totalGlobalMem;
  c->red_cap = 4096;
  // a failure below must not leak the half-built context: ibx_finalize releases whatever exists
  auto setup = [&]() -> int {
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
      // halo exchange, and the ghost update of the overlapped sharded step: highest priority, so that their short kernels
      // get SM slots while the long flux kernels of the compute stream are resident
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    }
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&c->d_red, c->red_cap * sizeof(double)));
    CU(cudaMallocHost((void**)&c->h_red, c->red_cap * sizeof(double)));
    return IBX_OK;
  };
  if (int rc = setup()) {
    ibx_finalize(c);
    return rc;
  }
  *out = c;
  return IBX_OK;
}

int ibx_finalize(ibx_ctx* c) {
  if (!c) return IBX_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->arrays) cudaFree(kv.second.p);
  if (c->d_red) cudaFree(c->d_red);
  if (c->h_red) cudaFreeHost(c->h_red);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->d_scratch2) cudaFree(c->d_scratch2);
  if (c->d_scratch3) cudaFree(c->d_scratch3);
  if (c->d_scratch4) cudaFree(c->d_scratch4);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  for (int k = 0; k < 2; ++k) {
    if (c->aux_stream[k]) cudaStreamDestroy(c->aux_stream[k]);
    if (c->aux_join[k]) cudaEventDestroy(c->aux_join[k]);
  }
  if (c->aux_fork) cudaEventDestroy(c->aux_fork);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (auto& S : c->e2e) {
    if (S.up) cudaEventDestroy(S.up);
    if (S.done) cudaEventDestroy(S.done);
    if (S.down) cudaEventDestroy(S.down);
  }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  for (auto& e : c->copy_ev)
    if (e) cudaEventDestroy(e);
  delete c;
  return IBX_OK;
}

int ibx_sync(ibx_ctx* c) {
  CHECK_CTX(c);
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->comm_stream));
  return IBX_OK;
}

int ibx_device_info(ibx_ctx* c, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem) {
  CHECK_CTX(c);
  *sm_count = c->sm_count;
  *cc_major = c->cc_major;
  *cc_minor = c->cc_minor;
  *total_mem = (int64_t)c->total_mem;
  return IBX_OK;
}

static int* option_slot(ibx_ctx* c, const char* name, int* lo, int* hi) {
  std::string n = name ? name : "";
  if (n == "path") { *lo = 0; *hi = 2; return &c->opt_path; }
  if (n == "arithmetic") { *lo = 0; *hi = 1; return &c->opt_arith; }
  if (n == "sensor") { *lo = 0; *hi = 1; return &c->opt_sensor; }
  return nullptr;
}

int ibx_set_option(ibx_ctx* c, const char* name, int value) {
  CHECK_CTX(c);
  int lo, hi;
  int* slot = option_slot(c, name, &lo, &hi);
  if (!slot) return fail(IBX_ERR_ARG, std::string("ibx_set_option: unknown option '") + (name ? name : "") + "' (path, arithmetic, sensor)");
  if (value < lo || value > hi) return fail(IBX_ERR_ARG, std::string("ibx_set_option: value out of range for '") + name + "'");
  *slot = value;
  return IBX_OK;
}

int ibx_get_option(ibx_ctx* c, const char* name, int* value) {
  CHECK_CTX(c);
  int lo, hi;
  int* slot = option_slot(c, name, &lo, &hi);
  if (!slot) return fail(IBX_ERR_ARG, std::string("ibx_get_option: unknown option '") + (name ? name : "") + "'");
  *value = *slot;
  return IBX_OK;
}

int ibx_launch_count(ibx_ctx* c, int64_t* out) {
  CHECK_CTX(c);
  *out = c->launches;
  return IBX_OK;
}

int ibx_timer_start(ibx_ctx* c) {
  CHECK_CTX(c);
  CU(cudaEventRecord(c->ev0, c->stream));
  return IBX_OK;
}

int ibx_timer_stop(ibx_ctx* c, float* ms) {
  CHECK_CTX(c);
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return IBX_OK;
}

static int alloc_impl(ibx_ctx* c, int64_t rows, int64_t cols, bool f64, ibx_array* out) {
  CHECK_CTX(c);
  if (rows < 0 || cols < 1) return fail(IBX_ERR_ARG, "ibx_array_alloc: rows >= 0 and cols >= 1 required");
  float* p = nullptr;
  size_t bytes = std::max<size_t>((size_t)rows * cols, 1) * (f64 ? sizeof(double) : sizeof(float));
  CU(cudaMalloc((void**)&p, bytes));
  CU(cudaMemsetAsync(p, 0, bytes, c->stream));
  std::lock_guard<std::mutex> lk(c->mu);
  int64_t h = c->next_handle++;
  c->arrays[h] = {p, rows, cols, f64};
  *out = h;
  return IBX_OK;
}

int ibx_array_alloc(ibx_ctx* c, int64_t rows, int64_t cols, ibx_array* out) { return alloc_impl(c, rows, cols, false, out); }
int ibx_array_alloc_f64(ibx_ctx* c, int64_t rows, int64_t cols, ibx_array* out) { return alloc_impl(c, rows, cols, true, out); }

int ibx_array_is_f64(ibx_ctx* c, ibx_array a, int* out) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, a);
  *out = A.f64;
  return IBX_OK;
}

int ibx_array_download_f64(ibx_ctx* c, ibx_array a, double* host) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, a);
  if (!A.f64) return fail(IBX_ERR_ARG, "ibx_array_download_f64: array is float32");
  CU(cudaMemcpyAsync(host, A.p, (size_t)A.rows * A.cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IBX_OK;
}

int ibx_array_free(ibx_ctx* c, ibx_array a) {
  if (!c) return fail(IBX_ERR_ARG, "ibx_array_free: null context");
  cudaSetDevice(c->device);
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->arrays.find(a);
  if (it == c->arrays.end()) return fail(IBX_ERR_ARG, "ibx_array_free: invalid handle");
  cudaStreamSynchronize(c->stream);
  cudaFree(it->second.p);
  c->arrays.erase(it);
  return IBX_OK;
}

int ibx_array_shape(ibx_ctx* c, ibx_array a, int64_t* rows, int64_t* cols) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, a);
  *rows = A.rows;
  *cols = A.cols;
  return IBX_OK;
}

int ibx_array_upload(ibx_ctx* c, ibx_array a, const float* host) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  CU(cudaMemcpyAsync(A.p, host, (size_t)A.rows * A.cols * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IBX_OK;
}

int ibx_array_download(ibx_ctx* c, ibx_array a, float* host) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  CU(cudaMemcpyAsync(host, A.p, (size_t)A.rows * A.cols * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return IBX_OK;
}

static int ensure_copy_streams(ibx_ctx* c) {
  if (!c->h2d_stream) {
    CU(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  }
  for (auto& e : c->copy_ev)
    if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return IBX_OK;
}

int ibx_array_upload_async(ibx_ctx* c, ibx_array a, const float* host) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  int rc;
  if ((rc = ensure_copy_streams(c))) return rc;
  CU(cudaMemcpyAsync(A.p, host, (size_t)A.rows * A.cols * sizeof(float), cudaMemcpyHostToDevice, c->h2d_stream));
  CU(cudaEventRecord(c->copy_ev[0], c->h2d_stream));
  CU(cudaStreamWaitEvent(c->stream, c->copy_ev[0], 0));
  CU(cudaStreamWaitEvent(c->comm_stream, c->copy_ev[0], 0));
  return IBX_OK;
}

int ibx_array_download_async(ibx_ctx* c, ibx_array a, float* host) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  int rc;
  if ((rc = ensure_copy_streams(c))) return rc;
  CU(cudaEventRecord(c->copy_ev[1], c->stream));
  CU(cudaStreamWaitEvent(c->d2h_stream, c->copy_ev[1], 0));
  CU(cudaMemcpyAsync(host, A.p, (size_t)A.rows * A.cols * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream));
  return IBX_OK;
}

int ibx_download_fence(ibx_ctx* c, int slot) {
  CHECK_CTX(c);
  if (slot < 0 || slot > 1) return fail(IBX_ERR_ARG, "ibx_download_fence: slot must be 0 or 1");
  int rc;
  if ((rc = ensure_copy_streams(c))) return rc;
  CU(cudaEventRecord(c->copy_ev[2 + slot], c->d2h_stream));
  c->copy_fence_set[slot] = true;
  return IBX_OK;
}

int ibx_download_wait(ibx_ctx* c, int slot) {
  CHECK_CTX(c);
  if (slot < 0 || slot > 1) return fail(IBX_ERR_ARG, "ibx_download_wait: slot must be 0 or 1");
  if (!c->copy_fence_set[slot]) return IBX_OK;
  CU(cudaEventSynchronize(c->copy_ev[2 + slot]));
  c->copy_fence_set[slot] = false;
  return IBX_OK;
}

int ibx_array_copy(ibx_ctx* c, ibx_array dst, ibx_array src) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, dst);
  GET_ARR_ANY(B, src);
  if (A.rows != B.rows || A.cols != B.cols || A.f64 != B.f64) return fail(IBX_ERR_ARG, "ibx_array_copy: shape or element type mismatch");
  CU(cudaMemcpyAsync(A.p, B.p, (size_t)A.rows * A.cols * (A.f64 ? sizeof(double) : sizeof(float)), cudaMemcpyDeviceToDevice, c->stream));
  return IBX_OK;
}

int ibx_array_devptr(ibx_ctx* c, ibx_array a, void** ptr) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  *ptr = A.p;
  return IBX_OK;
}

int ibx_host_alloc(int64_t bytes, void** out) {
  cudaError_t e = cudaMallocHost(out, (size_t)bytes);
  if (e != cudaSuccess) return fail(IBX_ERR_CUDA, std::string("ibx_host_alloc: ") + cudaGetErrorString(e));
  return IBX_OK;
}

int ibx_host_free(void* p) {
  cudaFreeHost(p);
  return IBX_OK;
}

int ibx_domain_upload(ibx_ctx* c, ibx_domain* d) {
  CHECK_CTX(c);
  ibx_domain* Dp = find_domain(d);
  if (!Dp) return fail(IBX_ERR_ARG, "ibx_domain_upload: unknown domain handle");
  ibx_domain& D = *Dp;
  if (D.uploaded) return IBX_OK;
  int nd = D.nd;
  D.device = c->device;
  D.uploaded = true;  // so that a partial upload is released by the destructor
  const int64_t* l2g = nullptr;
  (void)l2g;
  for (auto& P : D.parts) {
    int rc;
    if ((rc = upload_vec(c, P.domain, &P.d_domain))) return rc;
    if ((rc = upload_vec(c, P.image_in_domain, &P.d_image_in_domain))) return rc;
    int64_t n = (int64_t)P.domain.size();
    std::vector<float> sp((size_t)n * nd), ce((size_t)n * nd);
    for (int64_t i = 0; i < n; ++i)
      for (int k = 0; k < nd; ++k) {
        sp[(size_t)k * n + i] = D.widths[(size_t)P.domain[i] * nd + k];
        ce[(size_t)k * n + i] = D.centers[(size_t)P.domain[i] * nd + k];
      }
    if ((rc = upload_vec(c, sp, &P.d_spacing))) return rc;
    if ((rc = upload_vec(c, ce, &P.d_centers))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    for (auto& T : P.dims) {
      if ((rc = upload_vec(c, T.owners, &T.d_owners))) return rc;
      if ((rc = upload_vec(c, T.neighbors, &T.d_neighbors))) return rc;
      if ((rc = upload_vec(c, T.lptr, &T.d_lptr))) return rc;
      if ((rc = upload_vec(c, T.lidx, &T.d_lidx))) return rc;
      if ((rc = upload_vec(c, T.rptr, &T.d_rptr))) return rc;
      if ((rc = upload_vec(c, T.ridx, &T.d_ridx))) return rc;
    }
  }
  for (auto& F : D.boundaries)
    for (auto& B : F.parts) {
      int rc;
      int64_t G = (int64_t)B.ghost.size();
      std::vector<int32_t> gidx(B.idx.size());
      for (size_t q = 0; q < B.idx.size(); ++q) gidx[q] = B.image_domain[B.idx[q]];
      std::vector<float> nrm((size_t)G * nd), eta(G);
      for (int64_t g = 0; g < G; ++g) {
        eta[g] = B.ghost_dist[g] / B.image_dist[g];  // eta, src/ImmersedBoundary.jl:1220
        for (int k = 0; k < nd; ++k) nrm[(size_t)k * G + g] = B.normals[(size_t)g * nd + k];
      }
      if ((rc = upload_vec(c, B.ghost, &B.d_ghost))) return rc;
      if ((rc = upload_vec(c, B.ptr, &B.d_ptr))) return rc;
      if ((rc = upload_vec(c, gidx, &B.d_idx_global))) return rc;
      if ((rc = upload_vec(c, B.image_domain, &B.d_image_domain))) return rc;
      if ((rc = upload_vec(c, B.idx, &B.d_idx))) return rc;
      if ((rc = upload_vec(c, B.w, &B.d_w))) return rc;
      if ((rc = upload_vec(c, nrm, &B.d_normals))) return rc;
      if ((rc = upload_vec(c, eta, &B.d_eta))) return rc;
      CU(cudaStreamSynchronize(c->stream));
    }
  for (auto& S : D.surfaces) {
    int rc;
    if ((rc = upload_vec(c, S.areas, &S.d_areas))) return rc;
    for (ibx_accum* A : {&S.interp, &S.offset_interp}) {
      if ((rc = upload_vec(c, A->ptr, &A->d_ptr))) return rc;
      if ((rc = upload_vec(c, A->idx, &A->d_idx))) return rc;
      if ((rc = upload_vec(c, A->w, &A->d_w))) return rc;
      A->uploaded = true;
    }
  }
  if (D.shard.active) {
    int rc;
    for (int peer = 0; peer < D.shard.nranks; ++peer) {
      if ((rc = upload_vec(c, D.shard.send_local[peer], &D.shard.d_send[peer]))) return rc;
      if ((rc = upload_vec(c, D.shard.recv_local[peer], &D.shard.d_recv[peer]))) return rc;
    }
  }
  {
    int rc;
    if ((rc = upload_vec(c, D.block_faces, &D.d_block_faces))) return rc;
    if ((rc = upload_vec(c, D.block_h, &D.d_block_h))) return rc;
    // work lists of the tile kernels
    int64_t cpb = 1;
    for (int k = 0; k < nd; ++k) cpb *= D.block_size;
    int64_t nblk = D.ncells / cpb, nown = D.shard.active ? D.shard.n_owned / cpb : nblk;
    std::vector<int32_t> ap, af, op, of, oreg, areg;
    for (int64_t b = 0; b < nblk; ++b) {
      bool finer = false, regular = true;
      for (int f = 0; f < 2 * nd; ++f) {
        finer |= D.block_faces[(size_t)b * 2 * nd + f].kind == 3;
        regular &= D.block_faces[(size_t)b * 2 * nd + f].kind == 1;
      }
      if (regular) areg.push_back((int32_t)b);
      else (finer ? af : ap).push_back((int32_t)b);
      if (b < nown) {
        if (regular) oreg.push_back((int32_t)b);
        else (finer ? of : op).push_back((int32_t)b);
      }
    }
    D.n_own_regular = (int)oreg.size();
    D.n_all_regular = (int)areg.size();
    if ((rc = upload_vec(c, oreg, &D.d_blk_own_regular))) return rc;
    if ((rc = upload_vec(c, areg, &D.d_blk_all_regular))) return rc;
    D.all_pow2 = true;
    for (float hv : D.block_h) {
      uint32_t u;
      memcpy(&u, &hv, 4);
      uint32_t e = (u >> 23) & 0xffu;
      if ((u & 0x007fffffu) != 0u || e <= 32u || e >= 222u || (u >> 31)) D.all_pow2 = false;
    }
    D.n_all_plain = (int)ap.size(); D.n_all_finer = (int)af.size(); D.n_own_plain = (int)op.size(); D.n_own_finer = (int)of.size();
    if ((rc = upload_vec(c, ap, &D.d_blk_all_plain))) return rc;
    if ((rc = upload_vec(c, af, &D.d_blk_all_finer))) return rc;
    if ((rc = upload_vec(c, op, &D.d_blk_own_plain))) return rc;
    if ((rc = upload_vec(c, of, &D.d_blk_own_finer))) return rc;
    if (nd == 3 && D.block_size == 8) {
      // ---- phase lists of the overlapped step (whole domains: only the ghost update is hidden; shards: the exchanges too).  final0: owned block without ghost cells (its state is final when the step
      // starts); E = blocks whose two rings of face neighbours are final0.
      using PL = ibx_domain::PhaseLists;
      std::vector<uint8_t> final0(nblk, 0), ok1(nblk, 0), inE(nblk, 0), inS(nblk, 0), inP(nblk, 0);
      for (int64_t b = 0; b < nown; ++b) final0[b] = 1;
      for (auto& F : D.boundaries)
        for (auto& B : F.parts)
          for (int32_t g : B.ghost) final0[(int64_t)g / cpb] = 0;
      auto for_nbrs = [&](int64_t b, auto&& fn) {
        for (int f = 0; f < 2 * nd; ++f) {
          const BlockFace& bf = D.block_faces[(size_t)b * 2 * nd + f];
          const int cnt = bf.kind == 3 ? (1 << (nd - 1)) : (bf.kind == 1 || bf.kind == 2 ? 1 : 0);
          for (int q = 0; q < cnt; ++q)
            if (bf.nb[q] >= 0 && bf.nb[q] < nblk) fn((int64_t)bf.nb[q]);
        }
      };
      for (int64_t b = 0; b < nown; ++b) {
        bool ok = final0[b];
        for_nbrs(b, [&](int64_t n) { ok = ok && final0[n]; });
        ok1[b] = ok;
      }
      for (int64_t b = 0; b < nown; ++b) {
        bool ok = ok1[b];
        for_nbrs(b, [&](int64_t n) { ok = ok && ok1[n]; });
        inE[b] = ok;
      }
      for (int64_t b = 0; b < nown; ++b)
        if (inE[b]) {
          inS[b] = 1;
          for_nbrs(b, [&](int64_t n) { inS[n] = 1; });
        }
      for (int64_t b = 0; b < nblk; ++b)
        if (inS[b]) {
          inP[b] = 1;
          for_nbrs(b, [&](int64_t n) { inP[n] = 1; });
        }
      std::vector<int32_t> L[2][PL::NLISTS];
      for (int64_t b = 0; b < nblk; ++b) {
        bool finer = false, regular = true;
        for (int f = 0; f < 2 * nd; ++f) {
          finer |= D.block_faces[(size_t)b * 2 * nd + f].kind == 3;
          regular &= D.block_faces[(size_t)b * 2 * nd + f].kind == 1;
        }
        const int cls = regular ? 0 : (finer ? 2 : 1);
        L[inP[b] ? 0 : 1][PL::PRIM].push_back((int32_t)b);
        L[inS[b] ? 0 : 1][PL::S_REG + cls].push_back((int32_t)b);
        if (b < nown) L[inE[b] ? 0 : 1][PL::F_REG + cls].push_back((int32_t)b);
      }
      {
        std::vector<int32_t> ng, gh;
        for (int64_t b = 0; b < nblk; ++b) (final0[b] || b >= nown ? ng : gh).push_back((int32_t)b);
        D.n_blk_noghost = (int)ng.size();
        D.n_blk_ghost = (int)gh.size();
        if ((rc = upload_vec(c, ng, &D.d_blk_noghost))) return rc;
        if ((rc = upload_vec(c, gh, &D.d_blk_ghost))) return rc;
      }
      for (int ph = 0; ph < 2; ++ph)
        for (int k = 0; k < PL::NLISTS; ++k) {
          D.phase[ph].n[k] = (int)L[ph][k].size();
          if ((rc = upload_vec(c, L[ph][k], &D.phase[ph].d[k]))) return rc;
        }
      D.phased = true;
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  return IBX_OK;
}

int ibx_shard_phase_info(const ibx_domain* d, int64_t* n_flux_early, int64_t* n_flux_late) {
  ibx_domain* Dp = find_domain(d);
  if (!Dp) return fail(IBX_ERR_ARG, "ibx_shard_phase_info: unknown domain handle");
  if (!Dp->uploaded) return fail(IBX_ERR_STATE, "ibx_shard_phase_info: domain tables not uploaded (ibx_domain_upload)");
  using PL = ibx_domain::PhaseLists;
  int64_t n[2] = {0, 0};
  for (int ph = 0; ph < 2; ++ph)
    for (int k = PL::F_REG; k <= PL::F_FINER; ++k) n[ph] += Dp->phase[ph].n[k];
  *n_flux_early = Dp->phased ? n[0] : 0;
  *n_flux_late = Dp->phased ? n[1] : 0;
  return IBX_OK;
}

int ibx_accum_upload(ibx_ctx* c, ibx_accum* a) {
  CHECK_CTX(c);
  ibx_accum* A = find_accum(a);
  if (!A) return fail(IBX_ERR_ARG, "ibx_accum_upload: unknown accumulator handle");
  if (A->uploaded) return IBX_OK;
  int rc;
  if ((rc = upload_vec(c, A->ptr, &A->d_ptr))) return rc;
  if ((rc = upload_vec(c, A->idx, &A->d_idx))) return rc;
  if (A->weighted && (rc = upload_vec(c, A->w, &A->d_w))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  A->uploaded = true;
  return IBX_OK;
}

}  // extern "C"
