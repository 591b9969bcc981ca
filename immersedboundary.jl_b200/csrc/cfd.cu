// Pointwise cfd.jl kernels (src/cfd.jl:62-64, :106-151, :243-300, :459-554) and the point-implicit block
// operations (src/point_implicit.jl:125-161).  Compiled with -fmad=false like ops.cu: same operation order
// and roundings as the Float32 reference; the HLL combination runs in double because the reference's `0.0`
// literals promote it to Float64 (src/cfd.jl:504-507) -- the result is rounded to float32 on store.
#include "device.cuh"

using namespace ibx;

namespace {

constexpr int TB = 256;
#define GRID(n) grid_for((n), TB, c->sm_count, 32), TB, 0, c->stream

__device__ __forceinline__ float clampT(float T) { return fmaxf(T, 10.0f); }

template <int ND>
__device__ __forceinline__ void prim2state(ibx_fluid f, const float* P, float* Q) {
  float T = clampT(P[1]);
  float k = P[2] * P[2];
#pragma unroll
  for (int d = 1; d < ND; ++d) k = k + P[2 + d] * P[2 + d];
  k = k / 2.0f;
  float rho = P[0] / (f.R * T);
  Q[0] = rho;
  Q[1] = rho * (f.R / (f.gamma - 1.0f) * T + k);
#pragma unroll
  for (int d = 0; d < ND; ++d) Q[2 + d] = rho * P[2 + d];
}

template <int ND>
__device__ __forceinline__ void state2prim(ibx_fluid f, const float* Q, float* P) {
  float rho = Q[0];
  float u[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) u[d] = Q[2 + d] / rho;
  float k = u[0] * u[0];
#pragma unroll
  for (int d = 1; d < ND; ++d) k = k + u[d] * u[d];
  k = k / 2.0f;
  float p = (f.gamma - 1.0f) * (Q[1] - rho * k);
  P[0] = p;
  P[1] = clampT(p / (rho * f.R));
#pragma unroll
  for (int d = 0; d < ND; ++d) P[2 + d] = u[d];
}

template <int ND, bool TO_PRIM>
__global__ void k_convert(ibx_fluid f, const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a[2 + ND], b[2 + ND];
#pragma unroll
    for (int v = 0; v < 2 + ND; ++v) a[v] = in[(int64_t)v * n + i];
    if (TO_PRIM) state2prim<ND>(f, a, b); else prim2state<ND>(f, a, b);
#pragma unroll
    for (int v = 0; v < 2 + ND; ++v) out[(int64_t)v * n + i] = b[v];
  }
}

__global__ void k_sound(ibx_fluid f, const float* __restrict__ T, float* __restrict__ a, int64_t n) {
  float gr = f.gamma * f.R;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    a[i] = sqrtf(gr * clampT(T[i]));
}

// HLL (src/cfd.jl:459-508) along Cartesian dim
template <int ND, class TO>
__global__ void k_hll(ibx_fluid f, const float* __restrict__ PL, const float* __restrict__ PR, int dim,
                      TO* __restrict__ F, int64_t n) {
  constexpr int NV = 2 + ND;
  float gr = f.gamma * f.R;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pl[NV], pr[NV], ql[NV], qr[NV], fl[NV], fr[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { pl[v] = PL[(int64_t)v * n + i]; pr[v] = PR[(int64_t)v * n + i]; }
    prim2state<ND>(f, pl, ql);
    prim2state<ND>(f, pr, qr);
    float uL = pl[2 + dim], uR = pr[2 + dim];
    float aL = sqrtf(gr * clampT(pl[1])), aR = sqrtf(gr * clampT(pr[1]));
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float l = ql[v], r = qr[v];
      if (v == 1) { l = l + pl[0]; r = r + pr[0]; }
      l = l * uL;
      r = r * uR;
      if (v == 2 + dim) { l = l + pl[0]; r = r + pr[0]; }
      fl[v] = l;
      fr[v] = r;
    }
    double SR = fmin((double)(uR - aR), 0.0);
    double SL = fmax((double)(uL + aL), 0.0);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      F[(int64_t)v * n + i] = (TO)((SL * (double)fl[v] - SR * (double)fr[v] + SR * SL * (double)(qr[v] - ql[v])) / (SL - SR));
  }
}

// sensor-Rusanov (src/cfd.jl:516-554)
template <int ND>
__global__ void k_sensor_flux(ibx_fluid f, const float* __restrict__ PL, const float* __restrict__ PR,
                              const float* __restrict__ nuL, const float* __restrict__ nuR, int nucols, int dim,
                              float* __restrict__ F, int64_t n) {
  constexpr int NV = 2 + ND;
  float gr = f.gamma * f.R;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pl[NV], pr[NV], ul[NV], ur[NV], pm[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      pl[v] = PL[(int64_t)v * n + i];
      pr[v] = PR[(int64_t)v * n + i];
      pm[v] = (pl[v] + pr[v]) / 2.0f;
    }
    prim2state<ND>(f, pl, ul);
    prim2state<ND>(f, pr, ur);
    ul[1] = ul[1] + pl[0];
    ur[1] = ur[1] + pr[0];
    float u = pm[2 + dim];
    float a = sqrtf(gr * clampT(pm[1]));
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float fv = (ul[v] + ur[v]) * u / 2.0f;
      if (v == 2 + dim) fv = fv + pm[0];
      float nu = fmaxf(nuL[(nucols == 1 ? 0 : (int64_t)v * n) + i], nuR[(nucols == 1 ? 0 : (int64_t)v * n) + i]);
      fv = fv + (ul[v] - ur[v]) * (nu * (a + fabsf(u)) / 2.0f);
      F[(int64_t)v * n + i] = fv;
    }
  }
}

struct BCParams {
  float p_inf, T_inf, u_inf[3];
  int normal_flow;
};

// FlowBC call (src/cfd.jl:243-300); the optional wall-shear scaling (:290-296) is applied by k_flowbc_ex
template <int ND>
__device__ __forceinline__ void flowbc_point(ibx_fluid f, const BCParams& bc, const float* P, const float* nrm, float* out,
                                             float transpiration = 0.0f) {
  float un;
  if (bc.normal_flow) {
    un = bc.u_inf[0];
  } else {
    un = nrm[0] * bc.u_inf[0];
#pragma unroll
    for (int d = 1; d < ND; ++d) un = un + nrm[d] * bc.u_inf[d];
  }
  float cur = P[2] * nrm[0];
#pragma unroll
  for (int d = 1; d < ND; ++d) cur = cur + P[2 + d] * nrm[d];
  float a = sqrtf(f.gamma * f.R * clampT(P[1]));
  float M = fabsf(un) / a;
  float sup = M > 1.0f ? 1.0f : 0.0f, sub = M <= 1.0f ? 1.0f : 0.0f;
  float ge = un >= 0.0f ? 1.0f : 0.0f, lt = un < 0.0f ? 1.0f : 0.0f;
  out[0] = ge * (sup * bc.p_inf + sub * P[0]) + lt * (sup * P[0] + sub * bc.p_inf);
  out[1] = (un > 0.0f ? 1.0f : 0.0f) * bc.T_inf + (un <= 0.0f ? 1.0f : 0.0f) * P[1];
  if (bc.normal_flow) {
    float corr = un - cur + transpiration;
#pragma unroll
    for (int d = 0; d < ND; ++d) out[2 + d] = P[2 + d] + nrm[d] * corr;
  } else {
#pragma unroll
    for (int d = 0; d < ND; ++d) out[2 + d] = lt * P[2 + d] + ge * bc.u_inf[d];
  }
}

template <int ND>
__global__ void k_flowbc(ibx_fluid f, BCParams bc, const float* __restrict__ P, const float* __restrict__ nrm,
                         float* __restrict__ out, int64_t n) {
  constexpr int NV = 2 + ND;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float p[NV], nn[ND], o[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) p[v] = P[(int64_t)v * n + i];
#pragma unroll
    for (int d = 0; d < ND; ++d) nn[d] = nrm[(int64_t)d * n + i];
    flowbc_point<ND>(f, bc, p, nn, o);
#pragma unroll
    for (int v = 0; v < NV; ++v) out[(int64_t)v * n + i] = o[v];
  }
}

// FlowBC with the keyword arguments of src/cfd.jl:245-249: transpiration (scalar or per point) and, when du!dn and
// image_distances are given, the wall-shear scaling ub *= (V - du!dn * d) / V, V = |ub| + eps (:290-296)
template <int ND>
__global__ void k_flowbc_ex(ibx_fluid f, BCParams bc, const float* __restrict__ P, const float* __restrict__ nrm,
                            const float* __restrict__ dist, const float* __restrict__ dudn, const float* __restrict__ transp,
                            float transp_scalar, float* __restrict__ out, int64_t n) {
  constexpr int NV = 2 + ND;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float p[NV], nn[ND], o[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) p[v] = P[(int64_t)v * n + i];
#pragma unroll
    for (int d = 0; d < ND; ++d) nn[d] = nrm[(int64_t)d * n + i];
    flowbc_point<ND>(f, bc, p, nn, o, transp ? transp[i] : transp_scalar);
    if (dudn) {
      float s = o[2] * o[2];
#pragma unroll
      for (int d = 1; d < ND; ++d) s = s + o[2 + d] * o[2 + d];
      const float V = sqrtf(s) + 1.1920929e-7f;
      const float k = (V - dudn[i] * dist[i]) / V;
#pragma unroll
      for (int d = 0; d < ND; ++d) o[2 + d] = o[2 + d] * k;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) out[(int64_t)v * n + i] = o[v];
  }
}

// per-cell pinv of an nv x nv block through one-sided Jacobi in double; D[p, j, i] at column j + nv * i
__global__ void k_block_pinv(const float* __restrict__ D, float* __restrict__ Dinv, int64_t n, int nv, double rtol) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    double U[49], V[49];
    for (int j = 0; j < nv; ++j)
      for (int i = 0; i < nv; ++i) {
        U[j * nv + i] = (double)D[(int64_t)(j + nv * i) * n + p];
        V[j * nv + i] = (i == j) ? 1.0 : 0.0;
      }
    for (int sweep = 0; sweep < 40; ++sweep) {
      double off = 0.0;
      for (int a = 0; a < nv - 1; ++a)
        for (int b = a + 1; b < nv; ++b) {
          double aa = 0, bb = 0, ab = 0;
          for (int r = 0; r < nv; ++r) { aa += U[r * nv + a] * U[r * nv + a]; bb += U[r * nv + b] * U[r * nv + b]; ab += U[r * nv + a] * U[r * nv + b]; }
          if (ab == 0.0) continue;
          off = fmax(off, fabs(ab) / sqrt(fmax(aa * bb, 1e-300)));
          double zeta = (bb - aa) / (2.0 * ab);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          for (int r = 0; r < nv; ++r) {
            double x = U[r * nv + a], y = U[r * nv + b];
            U[r * nv + a] = cs * x - sn * y;
            U[r * nv + b] = sn * x + cs * y;
            x = V[r * nv + a]; y = V[r * nv + b];
            V[r * nv + a] = cs * x - sn * y;
            V[r * nv + b] = sn * x + cs * y;
          }
        }
      if (off < 1e-14) break;
    }
    double s2[7], smax = 0.0;
    for (int a = 0; a < nv; ++a) {
      double aa = 0;
      for (int r = 0; r < nv; ++r) aa += U[r * nv + a] * U[r * nv + a];
      s2[a] = aa;
      smax = fmax(smax, sqrt(aa));
    }
    for (int r = 0; r < nv; ++r)
      for (int q = 0; q < nv; ++q) {
        double acc = 0.0;
        for (int a = 0; a < nv; ++a)
          if (sqrt(s2[a]) > rtol * smax && s2[a] > 0.0) acc += V[r * nv + a] * U[q * nv + a] / s2[a];
        Dinv[(int64_t)(r + nv * q) * n + p] = (float)acc;
      }
  }
}

// out[p, j] = sum_i Dinv[p, j, i] * v[p, i], left to right (sum(...; dims = 3), src/point_implicit.jl:153-161)
__global__ void k_block_apply(const float* __restrict__ Dinv, const float* __restrict__ v, float* __restrict__ out,
                              int64_t n, int nv) {
  int64_t tot = n * nv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = t % n;
    int j = (int)(t / n);
    float acc = v[p] * Dinv[(int64_t)j * n + p];
    for (int i = 1; i < nv; ++i) acc = acc + v[(int64_t)i * n + p] * Dinv[(int64_t)(j + nv * i) * n + p];
    out[t] = acc;
  }
}

}  // namespace

#define SHAPE(cond, msg) \
  if (!(cond)) return fail(IBX_ERR_ARG, std::string(__func__) + ": shape mismatch: " + (msg))

extern "C" {

int ibx_state2primitive(ibx_ctx* c, ibx_fluid f, ibx_array Q, ibx_array P) {
  CHECK_CTX(c);
  GET_ARR(A, Q);
  GET_ARR(B, P);
  SHAPE(A.rows == B.rows && A.cols == B.cols && (A.cols == 4 || A.cols == 5), "Q, P must be N x (2 + nd), nd = 2 or 3");
  if (A.cols == 4) k_convert<2, true><<<GRID(A.rows)>>>(f, A.p, B.p, A.rows);
  else k_convert<3, true><<<GRID(A.rows)>>>(f, A.p, B.p, A.rows);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_primitive2state(ibx_ctx* c, ibx_fluid f, ibx_array P, ibx_array Q) {
  CHECK_CTX(c);
  GET_ARR(A, P);
  GET_ARR(B, Q);
  SHAPE(A.rows == B.rows && A.cols == B.cols && (A.cols == 4 || A.cols == 5), "P, Q must be N x (2 + nd), nd = 2 or 3");
  if (A.cols == 4) k_convert<2, false><<<GRID(A.rows)>>>(f, A.p, B.p, A.rows);
  else k_convert<3, false><<<GRID(A.rows)>>>(f, A.p, B.p, A.rows);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_speed_of_sound(ibx_ctx* c, ibx_fluid f, ibx_array T, ibx_array a) {
  CHECK_CTX(c);
  GET_ARR(A, T);
  GET_ARR(B, a);
  SHAPE(A.rows == B.rows && A.cols == B.cols, "T and a must match");
  k_sound<<<GRID(A.rows * A.cols)>>>(f, A.p, B.p, A.rows * A.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_inviscid_fluxes_hll(ibx_ctx* c, ibx_fluid f, ibx_array PL, ibx_array PR, int dim, ibx_array F) {
  CHECK_CTX(c);
  GET_ARR(L, PL);
  GET_ARR(R, PR);
  GET_ARR_ANY(O, F);  // float64 output keeps the reference's promotion (src/cfd.jl:504-507); float32 rounds once
  SHAPE(L.rows == R.rows && L.rows == O.rows && L.cols == R.cols && L.cols == O.cols && (L.cols == 4 || L.cols == 5),
        "PL, PR, F must be nfaces x (2 + nd)");
  if (dim < 0 || dim >= L.cols - 2) return fail(IBX_ERR_ARG, "ibx_inviscid_fluxes_hll: dim out of range");
  if (O.f64) {
    if (L.cols == 4) k_hll<2, double><<<GRID(L.rows)>>>(f, L.p, R.p, dim, (double*)O.p, L.rows);
    else k_hll<3, double><<<GRID(L.rows)>>>(f, L.p, R.p, dim, (double*)O.p, L.rows);
  } else if (L.cols == 4) k_hll<2, float><<<GRID(L.rows)>>>(f, L.p, R.p, dim, O.p, L.rows);
  else k_hll<3, float><<<GRID(L.rows)>>>(f, L.p, R.p, dim, O.p, L.rows);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_inviscid_fluxes_sensor(ibx_ctx* c, ibx_fluid f, ibx_array PL, ibx_array PR, ibx_array nuL, ibx_array nuR, int dim,
                               ibx_array F) {
  CHECK_CTX(c);
  GET_ARR(L, PL);
  GET_ARR(R, PR);
  GET_ARR(NL, nuL);
  GET_ARR(NR, nuR);
  GET_ARR(O, F);
  SHAPE(L.rows == R.rows && L.rows == O.rows && L.cols == R.cols && L.cols == O.cols && (L.cols == 4 || L.cols == 5),
        "PL, PR, F must be nfaces x (2 + nd)");
  SHAPE(NL.rows == L.rows && NR.rows == L.rows && NL.cols == NR.cols && (NL.cols == 1 || NL.cols == L.cols),
        "nuL, nuR must be nfaces x (1 | nv)");
  if (dim < 0 || dim >= L.cols - 2) return fail(IBX_ERR_ARG, "ibx_inviscid_fluxes_sensor: dim out of range");
  if (L.cols == 4) k_sensor_flux<2><<<GRID(L.rows)>>>(f, L.p, R.p, NL.p, NR.p, (int)NL.cols, dim, O.p, L.rows);
  else k_sensor_flux<3><<<GRID(L.rows)>>>(f, L.p, R.p, NL.p, NR.p, (int)NL.cols, dim, O.p, L.rows);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_flowbc(ibx_ctx* c, ibx_fluid f, const float* Pinf, int n_pinf, int normal_flow, ibx_array P, ibx_array normals,
               ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(A, P);
  GET_ARR(N, normals);
  GET_ARR(O, out);
  SHAPE(A.rows == N.rows && A.rows == O.rows && A.cols == O.cols && (A.cols == 4 || A.cols == 5) && N.cols == A.cols - 2,
        "P, out N x (2 + nd); normals N x nd");
  int nd = (int)A.cols - 2;
  if (normal_flow) {
    if (n_pinf != 3)  // the reference's @assert (src/cfd.jl:254)
      return fail(IBX_ERR_ARG, "Only 3 parcels in P (p, T and normal flow) allowed for normal_flow = true BC");
  } else if (n_pinf != 2 + nd) {
    return fail(IBX_ERR_ARG, "ibx_flowbc: Pinf must hold p, T and nd velocity components");
  }
  BCParams bc{};
  bc.p_inf = Pinf[0];
  bc.T_inf = Pinf[1];
  for (int k = 0; k < n_pinf - 2; ++k) bc.u_inf[k] = Pinf[2 + k];
  bc.normal_flow = normal_flow;
  if (nd == 2) k_flowbc<2><<<GRID(A.rows)>>>(f, bc, A.p, N.p, O.p, A.rows);
  else k_flowbc<3><<<GRID(A.rows)>>>(f, bc, A.p, N.p, O.p, A.rows);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_flowbc_ex(ibx_ctx* c, ibx_fluid f, const float* Pinf, int n_pinf, int normal_flow, ibx_array P, ibx_array normals,
                  ibx_array image_distances, ibx_array du_dn, ibx_array transpiration, float transpiration_scalar, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(A, P);
  GET_ARR(N, normals);
  GET_ARR(O, out);
  SHAPE(A.rows == N.rows && A.rows == O.rows && A.cols == O.cols && (A.cols == 4 || A.cols == 5) && N.cols == A.cols - 2,
        "P, out N x (2 + nd); normals N x nd");
  int nd = (int)A.cols - 2;
  if (normal_flow) {
    if (n_pinf != 3)  // the reference's @assert (src/cfd.jl:254)
      return fail(IBX_ERR_ARG, "Only 3 parcels in P (p, T and normal flow) allowed for normal_flow = true BC");
  } else if (n_pinf != 2 + nd) {
    return fail(IBX_ERR_ARG, "ibx_flowbc_ex: Pinf must hold p, T and nd velocity components");
  }
  if ((du_dn == 0) != (image_distances == 0))  // the reference's error (src/cfd.jl:286-288)
    return fail(IBX_ERR_ARG, "du!dn and image_distances must be passed together for BC imposition");
  const float *dist = nullptr, *dn = nullptr, *tr = nullptr;
  if (du_dn) {
    GET_ARR(Dd, image_distances);
    GET_ARR(Dn, du_dn);
    SHAPE(Dd.rows * Dd.cols == A.rows && Dn.rows * Dn.cols == A.rows, "image_distances and du!dn must have one entry per point");
    dist = Dd.p;
    dn = Dn.p;
  }
  if (transpiration) {
    GET_ARR(Tt, transpiration);
    SHAPE(Tt.rows * Tt.cols == A.rows, "transpiration must have one entry per point");
    tr = Tt.p;
  }
  BCParams bc{};
  bc.p_inf = Pinf[0];
  bc.T_inf = Pinf[1];
  for (int k = 0; k < n_pinf - 2; ++k) bc.u_inf[k] = Pinf[2 + k];
  bc.normal_flow = normal_flow;
  if (nd == 2) k_flowbc_ex<2><<<GRID(A.rows)>>>(f, bc, A.p, N.p, dist, dn, tr, transpiration_scalar, O.p, A.rows);
  else k_flowbc_ex<3><<<GRID(A.rows)>>>(f, bc, A.p, N.p, dist, dn, tr, transpiration_scalar, O.p, A.rows);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_block_pinv(ibx_ctx* c, ibx_array D, int nv, ibx_array Dinv) {
  CHECK_CTX(c);
  GET_ARR(A, D);
  GET_ARR(B, Dinv);
  if (nv < 1 || nv > 7) return fail(IBX_ERR_ARG, "ibx_block_pinv: nv must be in 1..7");
  SHAPE(A.cols == (int64_t)nv * nv && B.rows == A.rows && B.cols == A.cols, "D, Dinv must be N x nv^2");
  k_block_pinv<<<grid_for(A.rows, 64, c->sm_count, 32), 64, 0, c->stream>>>(A.p, B.p, A.rows, nv, 1.1920928955078125e-07 * nv);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_block_apply(ibx_ctx* c, ibx_array Dinv, int nv, ibx_array v, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(A, Dinv);
  GET_ARR(V, v);
  GET_ARR(O, out);
  if (nv < 1 || nv > 7) return fail(IBX_ERR_ARG, "ibx_block_apply: nv must be in 1..7");
  SHAPE(A.cols == (int64_t)nv * nv && V.rows == A.rows && V.cols == nv && O.rows == A.rows && O.cols == nv, "Dinv N x nv^2; v, out N x nv");
  k_block_apply<<<GRID(A.rows * nv)>>>(A.p, V.p, O.p, A.rows, nv);
  LAUNCH_CHECK();
  return IBX_OK;
}

}  // extern "C"
