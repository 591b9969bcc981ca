// Configuration C5: viscous + Wray-Agarwal part of the canonical RANS residual (oracle/euler.py: rans_residual; the inviscid
// mean-flow part is ibx_residual_euler's fused kernels, unchanged).  Table-free per-cell kernels over the block-structured
// topology (topo.cuh), so the path runs on rank-local shards like the Euler residual: every quantity a cell needs comes
// from cells within its 2-deep face skirt (src/ImmersedBoundary.jl:610-619) -- gradients of the first skirt layer are
// recomputed locally, nothing beyond the state has to be exchanged.
//
//   k_rans_state   W = [state2primitive(Q) | R = qR / rho]                         (src/cfd.jl:137-151)
//   k_rans_grad    G_d = cell_gradient(part, [T u v w R], d), d = 1..3                 (src/ImmersedBoundary.jl:965-987)
//                  S = shear_rate(grad u) (src/turbulence.jl:110-124), nu_eff = mu(T) / rho + sigma_R R
//   k_rans_source  rho * Wray_Agarwal(R, S, grad R, grad S).S                          (src/turbulence.jl:222-241)
//   k_rans_flux    R5 += green_gauss(viscous_fluxes(at_faces(P), face_gradient(P, d, grad P), d; mu_t = at_faces(rho R)))
//                  RR  = sum_d green_gauss(rho_f nu_eff_f face_gradient(R) - rho_f (u_f (R_L + R_R) / 2 - |u_f| (R_R - R_L) / 2))
//                        + source        (src/cfd.jl:664-736, src/ImmersedBoundary.jl:899-926, 1039-1069, 1113-1157)
// Float32 in the reference's operation order, no FMA contraction (built with -fmad=false): products by the 1/len face
// weights first, sums in list order.
#include "device.cuh"
#include "physics.cuh"
#include "topo.cuh"

namespace {

constexpr int TB = 256;
constexpr int ND = 3, NV = 5, NW = 6, NG = 5;   // W: p T u v w R;  gradients of T u v w R
constexpr float EPS32 = 1.1920928955078125e-07f;

struct WA { float sigma_R, C1, kappa; };

__device__ __forceinline__ float ipw(float x, int n) {
  double r = 1.0;
  for (int k = 0; k < n; ++k) r *= (double)x;
  return (float)r;
}
// src/cfd.jl:71-77.  (T / T_ref) ^ (2/3) as cbrt(x)^2 in Float32 (~2 ulp) instead of Julia's Float64 exp2(log2 x * y): the
// viscosity is evaluated at every face of every cell (12 times per cell), and the Float64 transcendental pair dominated the
// kernel; 2e-7 relative on mu is far inside the 1e-5 tolerance of the viscous terms (the pointwise ibx_dynamic_viscosity keeps
// the exact form).
__device__ __forceinline__ float viscosity(const ibx_transport& t, float T) {
  T = fmaxf(T, 10.0f);
  const float cr = cbrtf(T / t.T_ref);
  return t.mu_ref * (cr * cr) * (t.T_ref + t.S) / (T + t.S);
}
__device__ __forceinline__ float conductivity(const ibx_transport& t, float T) {   // src/cfd.jl:84-90
  float k = 0.0f * T;
  for (int i = 0; i < t.nk; ++i) k = k + t.k[i] * ipw(T, i);
  return k;
}

__global__ void k_rans_state(ibx_fluid f, const float* __restrict__ Q, const float* __restrict__ qR, float* __restrict__ W, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float q[NV], p[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) q[v] = Q[(int64_t)v * n + i];
    s2p<ND>(f, q, p);
#pragma unroll
    for (int v = 0; v < NV; ++v) W[(int64_t)v * n + i] = p[v];
    W[(int64_t)NV * n + i] = qR[i] / q[0];
  }
}

// G[(d * NG + k) * N + cell], k = 0..4 <-> T u v w R;  AUX[0] = shear rate, AUX[1] = nu_eff
__global__ void __launch_bounds__(TB) k_rans_grad(Topo T, ibx_transport tr, WA wa, const float* __restrict__ Q, const float* __restrict__ W,
                                                  float* __restrict__ G, float* __restrict__ AUX) {
  const int64_t N = T.ncells;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < N; cell += (int64_t)gridDim.x * blockDim.x) {
    float vg[ND][ND];   // vg[i][j] = d u_i / d x_j
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float g[NW];
      cell_grad<ND, NW>(T, W, cell, d, g);
#pragma unroll
      for (int k = 0; k < NG; ++k) G[((int64_t)d * NG + k) * N + cell] = g[1 + k];
#pragma unroll
      for (int i = 0; i < ND; ++i) vg[i][d] = g[2 + i];
    }
    float s = 0.0f;   // shear_rate: sqrt(2 S_ij S_ij), i outer, j inner
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        const float e = (vg[i][j] + vg[j][i]) / 2.0f;
        s = s + e * e;
      }
    AUX[cell] = sqrtf(2.0f * s);
    AUX[N + cell] = viscosity(tr, W[N + cell]) / Q[cell] + W[(int64_t)NV * N + cell] * wa.sigma_R;
  }
}

__global__ void __launch_bounds__(TB) k_rans_source(Topo T, WA wa, const float* __restrict__ Q, const float* __restrict__ W,
                                                    const float* __restrict__ G, const float* __restrict__ AUX, float* __restrict__ SRC) {
  const int64_t N = T.ncells;
  const float C2 = wa.sigma_R + wa.C1 / (wa.kappa * wa.kappa);
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < T.n_compute; cell += (int64_t)gridDim.x * blockDim.x) {
    float dot = 0.0f;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float gS;
      cell_grad<ND, 1>(T, AUX, cell, d, &gS);
      const float t = G[((int64_t)d * NG + 4) * N + cell] * gS;
      dot = d == 0 ? t : dot + t;
    }
    const float R = W[(int64_t)NV * N + cell], S = AUX[cell];
    const float src = wa.C1 * R * S + C2 * dot * (R / (S + EPS32));
    SRC[cell] = Q[cell] * fminf(src, 10.0f * R);
  }
}

// latency-bound gathers (about 50 scattered loads per face): occupancy matters more than a few spills -> 128 threads, 6 CTAs / SM
__global__ void __launch_bounds__(128, 6) k_rans_flux(Topo T, ibx_transport tr, const float* __restrict__ Q, const float* __restrict__ qR,
                                                  const float* __restrict__ W, const float* __restrict__ G, const float* __restrict__ AUX,
                                                  const float* __restrict__ SRC, float* __restrict__ R5, float* __restrict__ RR) {
  const int64_t N = T.ncells;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < T.n_compute; cell += (int64_t)gridDim.x * blockDim.x) {
    int64_t b;
    int ii[ND];
    decode<ND>(T, cell, b, ii);
    float res[NV], rr = 0.0f;
#pragma unroll
    for (int v = 1; v < NV; ++v) res[v] = R5[(int64_t)v * N + cell];
#pragma unroll   // d is a compile-time index of the small per-face arrays below (no local memory)
    for (int d = 0; d < ND; ++d) {
      const float hc = T.h[b * ND + d];
      float msum[2][NV], m6[2];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        Nbr<ND> nb = neighbors<ND>(T, cell, b, ii, d, side);
        const float w = 1.0f / (float)nb.cnt;
        for (int k = 0; k < nb.cnt; ++k) {
          // owner = the low-side cell of the face; a box face has owner == neighbour == this cell
          const int64_t o = side ? cell : nb.cell[k], n = side ? nb.cell[k] : cell;
          const float ho = side ? hc : nb.h, hn = side ? nb.h : hc;
          const float fd = (ho + hn) / 2.0f;                                    // face_distance (:995-1002)
          float Pf[NV], gf[ND][NG];   // face values of T u v w (Pf[1..4]); gradients of T u v w R along each axis
#pragma unroll
          for (int v = 1; v < NV; ++v) Pf[v] = face_interp(W[(int64_t)v * N + o], W[(int64_t)v * N + n], ho, hn);
#pragma unroll
          for (int j = 0; j < ND; ++j)
#pragma unroll
            for (int q = 0; q < NG; ++q) {
              if (j == d) gf[j][q] = (W[(int64_t)(1 + q) * N + n] - W[(int64_t)(1 + q) * N + o]) / fd;
              else gf[j][q] = face_interp(G[((int64_t)j * NG + q) * N + o], G[((int64_t)j * NG + q) * N + n], ho, hn);
            }
          const float mu = viscosity(tr, Pf[1]) + face_interp(qR[o], qR[n], ho, hn);
          const float kc = conductivity(tr, Pf[1]);
          float divu = 0.0f;
#pragma unroll
          for (int i = 0; i < ND; ++i) divu = divu + gf[i][1 + i];
          float F[NV];
          F[0] = 0.0f;
          float fe = 0.0f + gf[d][0] * kc;
#pragma unroll
          for (int j = 0; j < ND; ++j) {
            // tau(d, j) = ((du_d/dx_j + du_j/dx_d) - (d == j ? 2/3 : 0) divu) mu
            const float tau = ((gf[j][1 + d] + gf[d][1 + j]) - (d == j ? 2.0f / 3 : 0.0f) * divu) * mu;
            fe = fe + tau * Pf[2 + j];
            F[2 + j] = tau;
          }
          F[1] = fe;
          // transported R: upwind convection on MUSCL states (no sensor) + diffusion
          float RL, RRt;
          {
            const float uo = W[(int64_t)NV * N + o], un = W[(int64_t)NV * N + n];
            const float go = G[((int64_t)d * NG + 4) * N + o], gn = G[((int64_t)d * NG + 4) * N + n];
            muscl_face<1>(&uo, &un, &go, &gn, ho, hn, 0.0f, 0.0f, false, false, &RL, &RRt);
          }
          const float uf = Pf[2 + d];
          const float rf = face_interp(Q[o], Q[n], ho, hn);
          const float Fc = rf * (uf * (RL + RRt) / 2.0f - fabsf(uf) * (RRt - RL) / 2.0f);
          const float Fd = rf * face_interp(AUX[N + o], AUX[N + n], ho, hn) * gf[d][4];
          const float F6 = Fd - Fc;
#pragma unroll
          for (int v = 1; v < NV; ++v) msum[side][v] = k == 0 ? F[v] * w : msum[side][v] + F[v] * w;
          m6[side] = k == 0 ? F6 * w : m6[side] + F6 * w;
        }
      }
#pragma unroll
      for (int v = 1; v < NV; ++v) res[v] = res[v] + (msum[1][v] - msum[0][v]) / hc;
      rr = rr + (m6[1] - m6[0]) / hc;
    }
#pragma unroll
    for (int v = 1; v < NV; ++v) R5[(int64_t)v * N + cell] = res[v];
    RR[cell] = rr + SRC[cell];
  }
}

// IB ghost update of the transported variable: R = qR / rho at the image points (interpolated like impose_bc!,
// src/ImmersedBoundary.jl:1228: products by the weights first, list order), blended with the prescribed boundary value
// (`R = 0` at walls, `R_inf = 3 nu` in the far field, src/turbulence.jl:203), written back as rho_ghost * R_ghost.
__global__ void k_rans_ghost_stage(const float* __restrict__ rho, const float* __restrict__ qR, const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ eta,
                                   float R_bc, float* __restrict__ stage, int64_t G) {
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b0 = ptr[g], e0 = ptr[g + 1];
    float ia = 0.0f;
    for (int32_t k = b0; k < e0; ++k) {
      const int64_t c = idx[k];
      const float t = (qR[c] / rho[c]) * w[k];
      ia = k == b0 ? t : ia + t;
    }
    const float e = eta[g];
    stage[g] = e * ia + (1.0f - e) * R_bc;
  }
}

__global__ void k_rans_ghost_commit(const float* __restrict__ stage, const int32_t* __restrict__ ghost, const float* __restrict__ rho,
                                    float* __restrict__ qR, int64_t G) {
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x)
    qR[ghost[g]] = rho[ghost[g]] * stage[g];
}

}  // namespace

#define SHAPE(cond, msg) \
  if (!(cond)) return fail(IBX_ERR_ARG, std::string(__func__) + ": shape mismatch: " + (msg))

extern "C" {

int ibx_residual_rans(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, ibx_transport tr, float sigma_R, float C1, float kappa,
                      ibx_array Qh, ibx_array qRh, ibx_array Rh, ibx_array RRh, ibx_array cflh) {
  int rc;
  // 1. the Euler part: fused kernels (checks the context, the domain and the shapes of Q, R, cfl)
  if ((rc = ibx_residual_euler(c, d, f, 0, Qh, Rh, cflh))) return rc;
  GET_DOM(D, d);
  if (D.nd != 3) return fail(IBX_ERR_UNSUPPORTED, "ibx_residual_rans: 3-D meshes only (configuration C5)");
  if (tr.nk < 0 || tr.nk > 4) return fail(IBX_ERR_ARG, "ibx_residual_rans: 0 to 4 conductivity coefficients");
  GET_ARR(Q, Qh);
  GET_ARR(QR, qRh);
  GET_ARR(R, Rh);
  GET_ARR(RR, RRh);
  const int64_t N = D.ncells;
  SHAPE(QR.rows == N && QR.cols == 1 && RR.rows == N && RR.cols == 1, "qR, RR must be ncells x 1");
  // scratch: W (6 N) | G (15 N) | AUX (2 N) | SRC (N), in its own buffer (the Euler scratch holds P and the sensor)
  const int64_t need = (int64_t)(NW + ND * NG + 2 + 1) * N;
  if (need > c->scratch3_cap) {
    if (c->d_scratch3) cudaFree(c->d_scratch3);
    c->d_scratch3 = nullptr;
    c->scratch3_cap = 0;
    CU(cudaMalloc((void**)&c->d_scratch3, (size_t)need * sizeof(float)));
    c->scratch3_cap = need;
  }
  float* W = c->d_scratch3;
  float* G = W + (int64_t)NW * N;
  float* AUX = G + (int64_t)ND * NG * N;
  float* SRC = AUX + 2 * N;
  Topo T = make_topo(D);
  const WA wa{sigma_R, C1, kappa};
  const int g = grid_for(N, TB, c->sm_count, 16);
  k_rans_state<<<g, TB, 0, c->stream>>>(f, Q.p, QR.p, W, N);
  LAUNCH_CHECK();
  k_rans_grad<<<g, TB, 0, c->stream>>>(T, tr, wa, Q.p, W, G, AUX);
  LAUNCH_CHECK();
  k_rans_source<<<g, TB, 0, c->stream>>>(T, wa, Q.p, W, G, AUX, SRC);
  LAUNCH_CHECK();
  k_rans_flux<<<grid_for(N, 128, c->sm_count, 32), 128, 0, c->stream>>>(T, tr, Q.p, QR.p, W, G, AUX, SRC, R.p, RR.p);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_ghost_update_rans(ibx_ctx* c, const ibx_domain* d, int b, ibx_array Qh, ibx_array qRh, float R_bc) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  if (b < 0 || b >= (int)D.boundaries.size()) return fail(IBX_ERR_ARG, "ibx_ghost_update_rans: boundary index out of range");
  GET_ARR(Q, Qh);
  GET_ARR(QR, qRh);
  const int64_t N = D.ncells;
  SHAPE(Q.rows == N && Q.cols == D.nd + 2 && QR.rows == N && QR.cols == 1, "Q must be ncells x (nd + 2), qR ncells x 1");
  int64_t Gtot = 0;
  for (auto& B : D.boundaries[b].parts) Gtot += (int64_t)B.ghost.size();
  if (Gtot == 0) return IBX_OK;
  if (Gtot > c->scratch4_cap) {
    if (c->d_scratch4) cudaFree(c->d_scratch4);
    c->d_scratch4 = nullptr;
    c->scratch4_cap = 0;
    CU(cudaMalloc((void**)&c->d_scratch4, (size_t)Gtot * sizeof(float)));
    c->scratch4_cap = Gtot;
  }
  // Jacobi across the chunks of the family: stage all, then commit all
  int64_t off = 0;
  for (auto& B : D.boundaries[b].parts) {
    const int64_t G = (int64_t)B.ghost.size();
    k_rans_ghost_stage<<<grid_for(G, 128, c->sm_count, 16), 128, 0, c->stream>>>(Q.p, QR.p, B.d_ptr, B.d_idx_global, B.d_w, B.d_eta, R_bc,
                                                                                  c->d_scratch4 + off, G);
    LAUNCH_CHECK();
    off += G;
  }
  off = 0;
  for (auto& B : D.boundaries[b].parts) {
    const int64_t G = (int64_t)B.ghost.size();
    k_rans_ghost_commit<<<grid_for(G, 128, c->sm_count, 16), 128, 0, c->stream>>>(c->d_scratch4 + off, B.d_ghost, Q.p, QR.p, G);
    LAUNCH_CHECK();
    off += G;
  }
  return IBX_OK;
}

}  // extern "C"
