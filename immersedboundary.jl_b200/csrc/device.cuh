// Device-side internals shared by the CUDA translation units of libibx.so.
#pragma once
#include <cuda_runtime.h>
#include <unordered_map>
#include <unordered_set>
#include <mutex>
#include "ibx_internal.h"

struct ibx_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;       // compute stream
  cudaStream_t comm_stream = nullptr;  // halo exchange / copies
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_halo = nullptr, ev_ready = nullptr;
  int64_t launches = 0;
  int64_t halo_pending = 0;   // array handle of a posted, not yet completed halo exchange (ibx_halo_begin / _end)
  bool poisoned = false;
  // run-time options of the fused residual (ibx_set_option): code path, arithmetic, sensor blend
  int opt_path = 0, opt_arith = 0, opt_sensor = 1;
  struct Arr { float* p; int64_t rows, cols; bool f64; };  // f64: the buffer holds doubles (HLL fluxes, src/cfd.jl:504-507)
  std::unordered_map<int64_t, Arr> arrays;
  int64_t next_handle = 1;
  std::mutex mu;
  std::unordered_set<const void*> smem_attr_done;   // kernels whose dynamic shared-memory limit is set on THIS device
  // scratch for reductions
  double* d_red = nullptr;
  double* h_red = nullptr;  // pinned
  int64_t red_cap = 0;
  // scratch for fused kernels (ghost staging etc.)
  float* d_scratch = nullptr;
  int64_t scratch_cap = 0;
  float* d_scratch2 = nullptr;  // fluxes of the general faces of irregular blocks (two-pass hybrid kernel)
  int64_t scratch2_cap = 0;
  float* d_scratch3 = nullptr;  // C5: primitives + R, cell gradients, shear rate / nu_eff, source (rans.cu)
  int64_t scratch3_cap = 0;
  float* d_scratch4 = nullptr;  // C5: ghost staging of the transported variable
  int64_t scratch4_cap = 0;
  // end-to-end staging arrays
  // two slots: the copies of one call overlap the compute / opposite-direction copies of the other (PCIe is full duplex)
  struct E2ESlot {
    ibx_array Q = 0, R = 0, cfl = 0;
    cudaEvent_t up = nullptr, done = nullptr, down = nullptr;
    bool busy = false;
  } e2e[2];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  // asynchronous array copies (ibx_array_upload_async / _download_async): [0] upload done, [1] compute done, [2 + slot] fences
  cudaEvent_t copy_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool copy_fence_set[2] = {false, false};
  // side streams of the fused residual: the irregular-block passes run beside the regular-block kernel
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
  cudaEvent_t aux_fork = nullptr, aux_join[2] = {nullptr, nullptr};
  // NCCL
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
};

namespace ibx {

int cuda_fail(ibx_ctx* c, cudaError_t e, const char* what, const char* file, int line);
ibx_domain* find_domain(const ibx_domain* d);
ibx_accum* find_accum(const ibx_accum* a);
bool get_array(ibx_ctx* c, ibx_array h, ibx_ctx::Arr& out);
float* ensure_scratch(ibx_ctx* c, int64_t nfloats);
// general faces of irregular blocks (gen.cu)
int general_faces(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, bool finer, ibx_fluid f, int flux_kind,
                  const float* P, const float* S, double* GF, float* GC, cudaStream_t st);
int sensor_direct(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, const float* p, float* S);
int sensor_regular(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, const float* p, float* S);
// halo exchange with / without waiting for the compute stream first (comm.cu)
int halo_begin_impl(ibx_ctx* c, const ibx_domain* d, ibx_array a, bool wait_compute);
// pencil-marching flux pass (march.cu)
bool march_supported(const ibx_domain& D);
int march_flux(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, int hyb, ibx_fluid f, int flux_kind, const float* P,
               const float* S, float* R, float* cfl, const double* GF, const float* GC, cudaStream_t st);
// the same pass with option "arithmetic" = 1 (march_fast.cu)
int march_flux_fast(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, int hyb, ibx_fluid f, int flux_kind, const float* P,
                    const float* S, float* R, float* cfl, const double* GF, const float* GC, cudaStream_t st);

#define CU(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) return ibx::cuda_fail(c, e_, #call, __FILE__, __LINE__);   \
  } while (0)

#define CHECK_CTX(c)                                                                        \
  do {                                                                                      \
    if (!(c)) return ibx::fail(IBX_ERR_ARG, std::string(__func__) + ": null context");      \
    if ((c)->poisoned) return ibx::fail(IBX_ERR_CUDA, "context poisoned by an earlier CUDA error"); \
    cudaSetDevice((c)->device);                                                             \
  } while (0)

#define GET_ARR_ANY(var, h)                                                                          \
  ibx_ctx::Arr var;                                                                                  \
  if (!ibx::get_array(c, (h), var)) return ibx::fail(IBX_ERR_ARG, std::string(__func__) + ": invalid array handle " #h)

#define GET_ARR(var, h)                                                                              \
  GET_ARR_ANY(var, h);                                                                               \
  if (var.f64) return ibx::fail(IBX_ERR_UNSUPPORTED, std::string(__func__) + ": float64 array passed as " #h " (only the HLL flux, green_gauss and the elementwise entry points take float64)")

#define GET_DOM(D, d)                                                                             \
  ibx_domain* D##_p = ibx::find_domain(d);                                                        \
  if (!D##_p) return ibx::fail(IBX_ERR_ARG, std::string(__func__) + ": unknown domain handle");   \
  if (!D##_p->uploaded) return ibx::fail(IBX_ERR_STATE, std::string(__func__) + ": domain tables not uploaded (ibx_domain_upload)"); \
  ibx_domain& D = *D##_p

#define LAUNCH_CHECK()                                                           \
  do {                                                                           \
    c->launches++;                                                               \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) return ibx::cuda_fail(c, e_, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: remember it per context (one context
// per device), not per process, and under the context lock (entry points may be called from several host threads).
inline int ensure_dyn_smem(ibx_ctx* c, const void* fn, size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->smem_attr_done.count(fn)) return IBX_OK;
  }
  if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  std::lock_guard<std::mutex> lk(c->mu);
  c->smem_attr_done.insert(fn);
  return IBX_OK;
}

inline int grid_for(int64_t n, int block, int sm_count, int per_sm = 16) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <class T>
int upload_vec(ibx_ctx* c, const std::vector<T>& v, T** dptr) {
  if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CU(cudaMalloc((void**)dptr, bytes));
  if (!v.empty()) CU(cudaMemcpyAsync(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return IBX_OK;
}

}  // namespace ibx
