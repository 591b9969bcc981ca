// Pointwise closures on either side of the residual: transport properties and viscous fluxes (src/cfd.jl:71-90,
// :664-736), the pointwise sensors (:563-617), the pressure coefficient (:420-426) and the turbulence closures of
// src/turbulence.jl.  One thread per point, structure-of-arrays operands, compiled with -fmad=false: the Float32
// operation order of the reference is kept.  Where the reference itself leaves Float32 it is followed:
//   * `x ^ y` with Float32 operands is evaluated by Julia in Float64 and rounded once (Base.Math.pow_body) -> pw();
//   * log / exp are taken in double and rounded (Julia's Float32 kernels are < 1 ulp; the tests allow 1e-5);
//   * WALE: `g2 .* (δ / 3)` has a Float64 scalar, so each S^d_ij term is Float64 and the running sum is rounded to
//     Float32 once per term (src/turbulence.jl:325-331).
#include "device.cuh"

using namespace ibx;

namespace {

constexpr int TB = 256;
#define GRID(n) grid_for((n), TB, c->sm_count, 32), TB, 0, c->stream

template <typename F>
__global__ void k_points(int64_t n, F f) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f(i);
}

__device__ __forceinline__ float pw(float x, float y) { return (float)exp2(log2(fabs((double)x)) * (double)y); }
__device__ __forceinline__ float ipw(float x, int n) {
  double r = 1.0;
  for (int k = 0; k < n; ++k) r *= (double)x;
  return (float)r;
}
__device__ __forceinline__ float logf_r(float x) { return (float)log((double)x); }
__device__ __forceinline__ float expf_r(float x) { return (float)exp((double)x); }

__device__ __forceinline__ float viscosity(const ibx_transport& t, float T) {
  T = fmaxf(T, 10.0f);
  return t.mu_ref * pw(T / t.T_ref, 2.0f / 3) * (t.T_ref + t.S) / (T + t.S);
}
__device__ __forceinline__ float conductivity(const ibx_transport& t, float T) {
  float k = 0.0f * T;
  for (int i = 0; i < t.nk; ++i) k = k + t.k[i] * ipw(T, i);
  return k;
}

struct Ptrs9 { const float* p[9]; };

// velocity-gradient table g[i * nd + j] = d u_i / d x_j, one vector per entry
int get_grad_table(ibx_ctx* c, int nd, const ibx_array* g, int64_t n, Ptrs9& T, const char* who) {
  if (nd != 2 && nd != 3) return fail(IBX_ERR_ARG, std::string(who) + ": nd must be 2 or 3");
  for (int k = 0; k < nd * nd; ++k) {
    ibx_ctx::Arr a;
    if (!get_array(c, g[k], a) || a.f64) return fail(IBX_ERR_ARG, std::string(who) + ": invalid gradient array handle");
    if (a.rows * a.cols != n) return fail(IBX_ERR_ARG, std::string(who) + ": gradient vectors must have the output's length");
    T.p[k] = a.p;
  }
  return IBX_OK;
}

struct WallOut { float y, u, mu, k, dudy; };

__device__ __forceinline__ WallOut wall_point(float rey, const ibx_wall_params& w) {
  const float eps = 1.1920929e-7f;
  rey = fmaxf(fabsf(rey), eps);
  float y = sqrtf(rey), u = 0.0f;
  for (int it = 0; it < w.n_iter; ++it) {
    u = fminf(logf_r(fmaxf(y, 1.0f)) / w.kappa + w.C, y);
    y = w.omega * (rey / u) + (1.0f - w.omega) * y;
  }
  u = rey / y;
  WallOut o;
  o.y = y;
  o.u = u;
  float t = 1.0f - expf_r(-y / w.A);
  o.mu = w.kappa * y * (t * t);
  o.dudy = 1.0f / (1.0f + o.mu);
  o.k = fminf(y * y / (6.0f * w.beta_star / w.beta - 2.0f), w.D * expf_r(-y / w.A_plus));
  return o;
}

}  // namespace

#define SHAPE(cond, msg) \
  if (!(cond)) return fail(IBX_ERR_ARG, std::string(__func__) + ": shape mismatch: " + (msg))
#define VEC(var, h, n_expected, what)                                                                          \
  GET_ARR(var, h);                                                                                             \
  if (var.rows * var.cols != (n_expected)) return fail(IBX_ERR_ARG, std::string(__func__) + ": " what " has the wrong length")

extern "C" {

int ibx_dynamic_viscosity(ibx_ctx* c, ibx_transport t, ibx_array T, ibx_array mu) {
  CHECK_CTX(c);
  GET_ARR(A, T);
  GET_ARR(B, mu);
  SHAPE(A.rows == B.rows && A.cols == B.cols, "T and mu must match");
  const float* a = A.p;
  float* b = B.p;
  k_points<<<GRID(A.rows * A.cols)>>>(A.rows * A.cols, [=] __device__(int64_t i) { b[i] = viscosity(t, a[i]); });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_heat_conductivity(ibx_ctx* c, ibx_transport t, ibx_array T, ibx_array k) {
  CHECK_CTX(c);
  GET_ARR(A, T);
  GET_ARR(B, k);
  SHAPE(A.rows == B.rows && A.cols == B.cols, "T and k must match");
  if (t.nk < 0 || t.nk > 4) return fail(IBX_ERR_ARG, "ibx_heat_conductivity: 0 to 4 polynomial coefficients");
  const float* a = A.p;
  float* b = B.p;
  k_points<<<GRID(A.rows * A.cols)>>>(A.rows * A.cols, [=] __device__(int64_t i) { b[i] = conductivity(t, a[i]); });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_viscous_fluxes(ibx_ctx* c, ibx_transport t, ibx_array P, const ibx_array* Pgrad, int dim, ibx_array normals,
                       ibx_array mu_t, float mu_t_scalar, ibx_array F) {
  CHECK_CTX(c);
  GET_ARR(A, P);
  GET_ARR(O, F);
  const int nv = (int)A.cols, nd = nv - 2;
  SHAPE((nv == 4 || nv == 5) && O.rows == A.rows && O.cols == A.cols, "P and F must be N x (2 + nd), nd = 2 or 3");
  if (t.nk < 0 || t.nk > 4) return fail(IBX_ERR_ARG, "ibx_viscous_fluxes: 0 to 4 conductivity coefficients");
  const int64_t n = A.rows;
  const float* G[3] = {nullptr, nullptr, nullptr};
  for (int j = 0; j < nd; ++j) {
    ibx_ctx::Arr g;
    if (!get_array(c, Pgrad[j], g) || g.f64) return fail(IBX_ERR_ARG, "ibx_viscous_fluxes: invalid gradient array handle");
    if (g.rows != n || g.cols != nv) return fail(IBX_ERR_ARG, "ibx_viscous_fluxes: every Pgrad[j] must have the shape of P");
    G[j] = g.p;
  }
  const float* nrm = nullptr;
  if (normals) {
    GET_ARR(Nn, normals);
    SHAPE(Nn.rows == n && Nn.cols == nd, "normals must be N x nd");
    nrm = Nn.p;
  } else if (dim < 0 || dim >= nd) {
    return fail(IBX_ERR_ARG, "ibx_viscous_fluxes: dim out of range (and no normal matrix given)");
  }
  const float* mt = nullptr;
  if (mu_t) {
    VEC(M, mu_t, n, "mu_t");
    mt = M.p;
  }
  const float* p = A.p;
  float* f = O.p;
  const float *g0 = G[0], *g1 = G[1], *g2 = G[2];
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    const float* g[3] = {g0, g1, g2};
    const float T = p[n + i];
    const float mu = viscosity(t, T) + (mt ? mt[i] : mu_t_scalar);
    const float k = conductivity(t, T);
    auto vg = [&](int a, int b) { return g[b][(int64_t)(2 + a) * n + i]; };   // d u_a / d x_b
    float divu = 0.0f;
    for (int a = 0; a < nd; ++a) divu = divu + vg(a, a);
    auto tau = [&](int a, int b) { return ((vg(a, b) + vg(b, a)) - (a == b ? 2.0f / 3 : 0.0f) * divu) * mu; };
    float out[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (!nrm) {
      out[1] = out[1] + g[dim][n + i] * k;
      for (int b = 0; b < nd; ++b) out[1] = out[1] + tau(dim, b) * p[(int64_t)(2 + b) * n + i];
      for (int b = 0; b < nd; ++b) out[2 + b] = out[2 + b] + tau(dim, b);
    } else {
      float td[3];
      for (int a = 0; a < nd; ++a) {
        float s = 0.0f;
        for (int b = 0; b < nd; ++b) s = s + tau(a, b) * nrm[(int64_t)b * n + i];
        td[a] = s;
      }
      for (int b = 0; b < nd; ++b) {
        out[1] = out[1] + (g[b][n + i] * k) * nrm[(int64_t)b * n + i];
        out[1] = out[1] + td[b] * p[(int64_t)(2 + b) * n + i];
      }
      for (int b = 0; b < nd; ++b) out[2 + b] = out[2 + b] + td[b];
    }
    for (int v = 0; v < nv; ++v) f[(int64_t)v * n + i] = out[v];
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_jst_sensor3(ibx_ctx* c, ibx_array Pim1, ibx_array Pi, ibx_array Pip1, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(A, Pim1);
  GET_ARR(B, Pi);
  GET_ARR(D, Pip1);
  GET_ARR(O, out);
  SHAPE(A.rows == B.rows && A.cols == B.cols && A.rows == D.rows && A.cols == D.cols && A.rows == O.rows && A.cols == O.cols,
        "Pim1, Pi, Pip1 and the output must match");
  const float *a = A.p, *b = B.p, *d = D.p;
  float* o = O.p;
  k_points<<<GRID(A.rows * A.cols)>>>(A.rows * A.cols, [=] __device__(int64_t i) {
    const float e = 1e-14f;
    o[i] = (fabsf(a[i] + d[i] - 2.0f * b[i]) + e) / (fabsf(a[i] - b[i]) + fabsf(d[i] - b[i]) + e);
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_shock_sensor(ibx_ctx* c, int nd, const ibx_array* g, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(O, out);
  const int64_t n = O.rows * O.cols;
  Ptrs9 T;
  int rc = get_grad_table(c, nd, g, n, T, "ibx_shock_sensor");
  if (rc) return rc;
  float* o = O.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    float vort = 0.0f, divu = 0.0f;
    for (int a = 0; a < nd; ++a) {                 // 0-based i, in = (i + 1) % nd, inn = (in + 1) % nd
      const int b = (a + 1) % nd, cc = (b + 1) % nd;
      divu = divu + T.p[a * nd + a][i];
      const float w = T.p[cc * nd + b][i] - T.p[b * nd + cc][i];
      vort = vort + w * w;
    }
    divu = divu * divu;
    const float e = 1e-14f;
    o[i] = (divu + e) / (divu + vort + e);
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_pressure_coefficient(ibx_ctx* c, float gamma, ibx_array p, float p_inf, float M_inf, ibx_array Cp) {
  CHECK_CTX(c);
  GET_ARR(A, p);
  GET_ARR(B, Cp);
  SHAPE(A.rows == B.rows && A.cols == B.cols, "p and Cp must match");
  const float* a = A.p;
  float* b = B.p;
  k_points<<<GRID(A.rows * A.cols)>>>(A.rows * A.cols, [=] __device__(int64_t i) {
    b[i] = 2.0f * (a[i] / p_inf - 1.0f) / (M_inf * M_inf * gamma);
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

// ------------------------------------------------------------------------------------------ turbulence.jl
int ibx_wall_function_rey(ibx_ctx* c, ibx_wall_params w, ibx_array Rey, ibx_array y_plus, ibx_array u_plus, ibx_array mu_plus,
                          ibx_array k_plus, ibx_array dudy_plus) {
  CHECK_CTX(c);
  GET_ARR(A, Rey);
  const int64_t n = A.rows * A.cols;
  VEC(Y, y_plus, n, "y+");
  VEC(U, u_plus, n, "u+");
  VEC(M, mu_plus, n, "mu+");
  VEC(K, k_plus, n, "k+");
  VEC(D, dudy_plus, n, "du+/dy+");
  const float* a = A.p;
  float *y = Y.p, *u = U.p, *m = M.p, *k = K.p, *d = D.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    WallOut o = wall_point(a[i], w);
    y[i] = o.y; u[i] = o.u; m[i] = o.mu; k[i] = o.k; d[i] = o.dudy;
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_wall_function(ibx_ctx* c, ibx_wall_params w, ibx_array y, ibx_array u, ibx_array nu, ibx_array u_tau, ibx_array nu_t,
                      ibx_array k, ibx_array omega, ibx_array eps, ibx_array dudn) {
  CHECK_CTX(c);
  GET_ARR(Yd, y);
  const int64_t n = Yd.rows * Yd.cols;
  VEC(Ud, u, n, "u");
  VEC(Nd, nu, n, "nu");
  VEC(O1, u_tau, n, "u_tau");
  VEC(O2, nu_t, n, "nu_t");
  VEC(O3, k, n, "k");
  VEC(O4, omega, n, "omega");
  VEC(O5, eps, n, "eps");
  VEC(O6, dudn, n, "du/dn");
  const float *yy = Yd.p, *uu = Ud.p, *nn = Nd.p;
  float *o1 = O1.p, *o2 = O2.p, *o3 = O3.p, *o4 = O4.p, *o5 = O5.p, *o6 = O6.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    WallOut o = wall_point(uu[i] * yy[i] / nn[i], w);
    const float ut = uu[i] / o.u;
    const float nt = o.mu * nn[i];
    const float kk = o.k * (ut * ut);
    const float om = kk / nt;
    o1[i] = ut; o2[i] = nt; o3[i] = kk; o4[i] = om;
    o5[i] = w.beta_star * om * kk;
    o6[i] = o.dudy * (ut * ut) / nn[i];
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_shear_rate(ibx_ctx* c, int nd, const ibx_array* g, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(O, out);
  const int64_t n = O.rows * O.cols;
  Ptrs9 T;
  int rc = get_grad_table(c, nd, g, n, T, "ibx_shear_rate");
  if (rc) return rc;
  float* o = O.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    float s = 0.0f;
    for (int a = 0; a < nd; ++a)
      for (int b = 0; b < nd; ++b) {
        const float e = (T.p[a * nd + b][i] + T.p[b * nd + a][i]) / 2.0f;
        s = s + e * e;
      }
    o[i] = sqrtf(2.0f * s);
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_smagorinsky(ibx_ctx* c, ibx_array Delta, ibx_array S, float Cs, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(A, Delta);
  const int64_t n = A.rows * A.cols;
  VEC(B, S, n, "S");
  VEC(O, out, n, "nu_SGS");
  const float *a = A.p, *b = B.p;
  float* o = O.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) { const float t = Cs * a[i]; o[i] = t * t * b[i]; });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_standard_keps(ibx_ctx* c, ibx_array k, ibx_array eps, ibx_array S, float Cmu, float sigma_k, float sigma_eps, float C1,
                      float C2, ibx_array nu_k, ibx_array nu_eps, ibx_array Sk, ibx_array Seps, ibx_array nu_t) {
  CHECK_CTX(c);
  GET_ARR(Kd, k);
  const int64_t n = Kd.rows * Kd.cols;
  VEC(Ed, eps, n, "eps");
  VEC(Sd, S, n, "S");
  VEC(O1, nu_k, n, "nu_k");
  VEC(O2, nu_eps, n, "nu_eps");
  VEC(O3, Sk, n, "Sk");
  VEC(O4, Seps, n, "Seps");
  VEC(O5, nu_t, n, "nu_t");
  const float *kk = Kd.p, *ee = Ed.p, *ss = Sd.p;
  float *o1 = O1.p, *o2 = O2.p, *o3 = O3.p, *o4 = O4.p, *o5 = O5.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    const float kv = kk[i], ev = ee[i], sv = ss[i];
    const float nt = Cmu * (kv * kv) / ev;
    const float Pk = nt * (sv * sv);
    o3[i] = Pk - ev;
    o4[i] = C1 * Pk * ev / kv - C2 * (ev * ev) / kv;
    o1[i] = nt / sigma_k;
    o2[i] = nt / sigma_eps;
    o5[i] = nt;
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_wray_agarwal(ibx_ctx* c, ibx_array R, ibx_array S, ibx_array gradR, ibx_array gradS, float sigma_R, float C1, float kappa,
                     ibx_array nu_R, ibx_array S_out) {
  CHECK_CTX(c);
  GET_ARR(Rd, R);
  const int64_t n = Rd.rows * Rd.cols;
  VEC(Sd, S, n, "S");
  GET_ARR(GR, gradR);
  GET_ARR(GS, gradS);
  SHAPE(GR.rows == n && GS.rows == n && GR.cols == GS.cols && GR.cols >= 1 && GR.cols <= 3, "gradR, gradS must be N x nd");
  VEC(O1, nu_R, n, "nu_R");
  VEC(O2, S_out, n, "S_out");
  const int nd = (int)GR.cols;
  const float C2 = sigma_R + C1 / (kappa * kappa);
  const float *rr = Rd.p, *ss = Sd.p, *gr = GR.p, *gs = GS.p;
  float *o1 = O1.p, *o2 = O2.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    const float eps = 1.1920929e-7f;
    float dot = gr[i] * gs[i];
    for (int d = 1; d < nd; ++d) dot = dot + gr[(int64_t)d * n + i] * gs[(int64_t)d * n + i];
    float s = C1 * rr[i] * ss[i] + C2 * dot * (rr[i] / (ss[i] + eps));
    o2[i] = fminf(s, 10.0f * rr[i]);
    o1[i] = rr[i] * sigma_R;
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_ducros_sensor(ibx_ctx* c, int nd, const ibx_array* g, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(O, out);
  const int64_t n = O.rows * O.cols;
  Ptrs9 T;
  int rc = get_grad_table(c, nd, g, n, T, "ibx_ducros_sensor");
  if (rc) return rc;
  float* o = O.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    const float eps = 1.1920929e-7f;
    float div2 = 0.0f;
    for (int a = 0; a < nd; ++a) div2 = div2 + T.p[a * nd + a][i];
    div2 = div2 * div2;
    float curl2;
    auto G = [&](int a, int b) { return T.p[a * nd + b][i]; };
    if (nd == 2) {
      const float w = G(1, 0) - G(0, 1);
      curl2 = w * w;
    } else {
      const float w0 = G(2, 1) - G(1, 2), w1 = G(0, 2) - G(2, 0), w2 = G(1, 0) - G(0, 1);
      curl2 = w0 * w0 + w1 * w1 + w2 * w2;
    }
    o[i] = (div2 + eps) / (div2 + curl2 + eps);
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_wale(ibx_ctx* c, ibx_array Delta, const ibx_array* g, float Cw, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(O, out);
  const int64_t n = O.rows * O.cols;
  VEC(Dd, Delta, n, "Delta");
  Ptrs9 T;
  int rc = get_grad_table(c, 3, g, n, T, "ibx_wale");
  if (rc) return rc;
  const float* dl = Dd.p;
  float* o = O.p;
  k_points<<<GRID(n)>>>(n, [=] __device__(int64_t i) {
    const float eps = 1.1920929e-7f;
    float G[3][3], G2[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) G[a][b] = T.p[a * 3 + b][i];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        float s = 0.0f;
        for (int k = 0; k < 3; ++k) s = s + G[a][k] * G[k][b];
        G2[a][b] = s;
      }
    float SS = 0.0f, SD = 0.0f;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        const float e = (G[a][b] + G[b][a]) / 2.0f;
        SS = SS + e * e;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        // Float64 term: the reference's `δ / 3` is a Float64 scalar
        const double term = (double)((G2[a][b] + G2[b][a]) / 2.0f) - (double)G2[a][b] * ((a == b ? 1.0 : 0.0) / 3.0);
        SD = (float)((double)SD + term * term);
      }
    const float d = dl[i];
    o[i] = Cw * (d * d) * pw(SD, 1.5f) / (pw(SS, 2.5f) + pw(SD, 1.25f) + eps);
  });
  LAUNCH_CHECK();
  return IBX_OK;
}

}  // extern "C"
