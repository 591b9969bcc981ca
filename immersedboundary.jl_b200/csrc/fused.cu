// Fused, block-structured residual kernels: the throughput path.
//
// Connectivity is implicit.  Inside a bs^nd block neighbours are index arithmetic; across a block face the
// per-block BlockFace table (6 entries of 28 B per 512 cells = 0.33 B/cell) names the same-level, coarser
// (2:1) or finer (1:2) neighbour block(s), or the domain box.  No per-cell/per-face index table is read,
// which is what lets the path approach the algorithmic traffic of SURVEY.md section 8(d).
//
// Semantics are those of the reference operators evaluated on ONE partition covering the whole domain
// (src/ImmersedBoundary.jl:605-698, :879-1157): a box face is a face whose owner and neighbour are the
// same cell; a coarse cell's list towards finer cells holds 2^(nd-1) faces in ascending cell order,
// averaged with weight 1/len.  Image-cell results do not depend on the partitioning because the skirt
// is 2 deep, so this equals the reference's per-partition evaluation followed by its scatter.
#include "device.cuh"
#include "physics.cuh"
#include "topo.cuh"

using namespace ibx;
using namespace ibxk;

namespace ibx {
int prim_blocks(ibx_ctx* c, const ibx_domain& D, bool early, ibx_fluid f, const float* Q, float* P);
int residual_euler_after_prim(ibx_ctx* c, const ibx_domain& D, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S, float* R,
                              float* cfl);
int residual_euler_phase(ibx_ctx* c, const ibx_domain& D, int phase, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S,
                         float* R, float* cfl);
bool tile_supported(const ibx_domain& D);
int residual_euler_tiles(ibx_ctx* c, const ibx_domain& D, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S,
                         float* R, float* cfl);
}

namespace {

constexpr int TB = 256;

// ------------------------------------------------------------------ pass 1a: Q -> P
template <int ND>
__global__ void k_prim(ibx_fluid f, const float* __restrict__ Q, float* __restrict__ P, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float rho = Q[i], E = Q[n + i];
    float u[ND], k = 0.0f;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      u[d] = Q[(int64_t)(2 + d) * n + i] / rho;
      k = d == 0 ? u[d] * u[d] : k + u[d] * u[d];
    }
    k = k / 2.0f;
    float p = (f.gamma - 1.0f) * (E - rho * k);
    P[i] = p;
    P[n + i] = clampT(p / (rho * f.R));
#pragma unroll
    for (int d = 0; d < ND; ++d) P[(int64_t)(2 + d) * n + i] = u[d];
  }
}

// ------------------------------------------------------------------ pass 1b: JST sensor of a scalar, max over dims
template <int ND>
__device__ __forceinline__ float sensor_cell(const Topo& T, const float* __restrict__ p, int64_t cell, int64_t b,
                                             const int (&ii)[ND]) {
  float pc = p[cell];
  float nu = 1e-7f;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    float g[2], a[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      Nbr<ND> nb = neighbors<ND>(T, cell, b, ii, d, side);
      float w = 1.0f / (float)nb.cnt;
      float accg = 0.0f, acca = 0.0f;
      for (int k = 0; k < nb.cnt; ++k) {
        float pn = p[nb.cell[k]];
        float fd = side ? pn - pc : pc - pn;  // p_neighbour - p_owner
        accg = k == 0 ? fd * w : accg + fd * w;
        acca = k == 0 ? fabsf(fd) * w : acca + fabsf(fd) * w;
      }
      g[side] = accg;
      a[side] = acca;
    }
    float hc = T.h[b * ND + d];
    float gg = (g[1] - g[0]) / hc, ugg = (a[1] + a[0]) / hc;
    nu = fmaxf(nu, (1e-7f + fabsf(gg)) / (1e-7f + ugg));
  }
  return nu;
}

template <int ND>
__global__ void k_sensor(Topo T, const float* __restrict__ p, float* __restrict__ D) {
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < T.ncells; cell += (int64_t)gridDim.x * blockDim.x) {
    int64_t b;
    int ii[ND];
    decode<ND>(T, cell, b, ii);
    D[cell] = sensor_cell<ND>(T, p, cell, b, ii);
  }
}

// ------------------------------------------------------------------ pass 2: gradients, MUSCL, flux, divergence
template <int ND>
__global__ void __launch_bounds__(TB) k_euler_flux(Topo T, ibx_fluid f, int flux_kind, const float* __restrict__ P,
                                                   const float* __restrict__ D, float* __restrict__ R, float* __restrict__ cfl) {
  constexpr int NV = ND + 2;
  const int64_t N = T.ncells;
  const float gr = f.gamma * f.R;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < T.n_compute; cell += (int64_t)gridDim.x * blockDim.x) {
    int64_t b;
    int ii[ND];
    decode<ND>(T, cell, b, ii);
    float pc[NV], res[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { pc[v] = P[(int64_t)v * N + cell]; res[v] = 0.0f; }
    float Dc = D[cell];
    float ac = sqrtf(gr * clampT(pc[1]));
    float cf = 0.0f;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float hc = T.h[b * ND + d];
      float gc[NV];
      cell_grad<ND, NV>(T, P, cell, d, gc);
      double msum[2][NV];
      float csum[2];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        Nbr<ND> nb = neighbors<ND>(T, cell, b, ii, d, side);
        float w = 1.0f / (float)nb.cnt;
        for (int k = 0; k < nb.cnt; ++k) {
          int64_t n = nb.cell[k];
          float pn[NV], gn[NV], pl[NV], pr[NV];
          double F[NV];
          float Dn;
          if (n == cell) {  // box face
#pragma unroll
            for (int v = 0; v < NV; ++v) { pn[v] = pc[v]; gn[v] = gc[v]; }
            Dn = Dc;
          } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) pn[v] = P[(int64_t)v * N + n];
            cell_grad<ND, NV>(T, P, n, d, gn);
            Dn = D[n];
          }
          float an = sqrtf(gr * clampT(pn[1]));
          float uf, af;
          if (side) {  // this cell is the owner
            muscl_face<NV>(pc, pn, gc, gn, hc, nb.h, Dc, Dn, true, false, pl, pr);
            uf = face_interp(pc[2 + d], pn[2 + d], hc, nb.h);
            af = face_interp(ac, an, hc, nb.h);
          } else {     // the neighbour entry is the owner
            muscl_face<NV>(pn, pc, gn, gc, nb.h, hc, Dn, Dc, true, false, pl, pr);
            uf = face_interp(pn[2 + d], pc[2 + d], nb.h, hc);
            af = face_interp(an, ac, nb.h, hc);
          }
          if (flux_kind == 0) {
            hll_flux<ND>(f, pl, pr, d, F);
          } else {
            float Ff[NV];
            float nu = side ? face_interp(Dc, Dn, hc, nb.h) : face_interp(Dn, Dc, nb.h, hc);
            rusanov_flux<ND>(f, pl, pr, nu, d, Ff);
#pragma unroll
            for (int v = 0; v < NV; ++v) F[v] = (double)Ff[v];
          }
          float ct = (fabsf(uf) + af) * w;
          csum[side] = k == 0 ? ct : csum[side] + ct;
          if (flux_kind == 0) {   // Float64 Green-Gauss sums, like the reference after its HLL promotion
#pragma unroll
            for (int v = 0; v < NV; ++v) msum[side][v] = k == 0 ? F[v] * (double)w : msum[side][v] + F[v] * (double)w;
          } else {                // the sensor flux stays Float32 end to end
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              float t = (float)F[v] * w;
              msum[side][v] = k == 0 ? (double)t : (double)((float)msum[side][v] + t);
            }
          }
        }
      }
      if (flux_kind == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) res[v] = (float)((double)res[v] - (msum[1][v] - msum[0][v]) / (double)hc);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) res[v] = res[v] - ((float)msum[1][v] - (float)msum[0][v]) / hc;
      }
      cf = cf + (csum[1] + csum[0]) / hc;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) R[(int64_t)v * N + cell] = res[v];
    cfl[cell] = cf;
  }
}

// linear advection of test/advection.jl:67-83: ud = -sum_d GG((uL+uR) Cf / 2 + |Cf| (uL-uR) / 2),
// spec = max_d UGG(at_faces(C_d))
template <int ND>
__global__ void __launch_bounds__(TB) k_advection(Topo T, const float* __restrict__ u, const float* __restrict__ C,
                                                  const float* __restrict__ D, float* __restrict__ ud, float* __restrict__ spec) {
  const int64_t N = T.ncells;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < T.n_compute; cell += (int64_t)gridDim.x * blockDim.x) {
    int64_t b;
    int ii[ND];
    decode<ND>(T, cell, b, ii);
    float uc = u[cell], Dc = D[cell];
    float res = 0.0f, sp = 0.0f;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float hc = T.h[b * ND + d];
      float cc = C[(int64_t)d * N + cell];
      float gc;
      cell_grad<ND, 1>(T, u, cell, d, &gc);
      float msum[2], csum[2];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        Nbr<ND> nb = neighbors<ND>(T, cell, b, ii, d, side);
        float w = 1.0f / (float)nb.cnt;
        for (int k = 0; k < nb.cnt; ++k) {
          int64_t n = nb.cell[k];
          float un = uc, gn = gc, Dn = Dc, cn = cc;
          if (n != cell) {
            un = u[n];
            cell_grad<ND, 1>(T, u, n, d, &gn);
            Dn = D[n];
            cn = C[(int64_t)d * N + n];
          }
          float uL, uR, Cf;
          if (side) {
            muscl_face<1>(&uc, &un, &gc, &gn, hc, nb.h, Dc, Dn, true, true, &uL, &uR);
            Cf = face_interp(cc, cn, hc, nb.h);
          } else {
            muscl_face<1>(&un, &uc, &gn, &gc, nb.h, hc, Dn, Dc, true, true, &uL, &uR);
            Cf = face_interp(cn, cc, nb.h, hc);
          }
          float F = (uL + uR) * Cf / 2.0f + fabsf(Cf) * (uL - uR) / 2.0f;
          msum[side] = k == 0 ? F * w : msum[side] + F * w;
          csum[side] = k == 0 ? Cf * w : csum[side] + Cf * w;
        }
      }
      res = res - (msum[1] - msum[0]) / hc;
      float s = (csum[1] + csum[0]) / hc;
      sp = d == 0 ? s : fmaxf(sp, s);
    }
    ud[cell] = res;
    spec[cell] = sp;
  }
}

// ------------------------------------------------------------------ IB ghost update on the conservative state
struct BCParams {
  float p_inf, T_inf, u_inf[3];
  int normal_flow;
};

template <int ND>
__global__ void k_ghost_stage(ibx_fluid f, BCParams bc, const float* __restrict__ Q, int64_t N,
                              const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ w,
                              const float* __restrict__ nrm, const float* __restrict__ eta, float* __restrict__ stage, int64_t G) {
  constexpr int NV = ND + 2;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    float ia[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) ia[v] = 0.0f;
    int32_t b0 = ptr[g], e0 = ptr[g + 1];
    for (int32_t k = b0; k < e0; ++k) {
      int64_t c = idx[k];
      float q[NV], p[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) q[v] = Q[(int64_t)v * N + c];
      s2p<ND>(f, q, p);
      float wk = w[k];
#pragma unroll
      for (int v = 0; v < NV; ++v) ia[v] = k == b0 ? p[v] * wk : ia[v] + p[v] * wk;
    }
    float n_[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) n_[d] = nrm[(int64_t)d * G + g];
    // FlowBC (src/cfd.jl:243-300)
    float un;
    if (bc.normal_flow) {
      un = bc.u_inf[0];
    } else {
      un = n_[0] * bc.u_inf[0];
#pragma unroll
      for (int d = 1; d < ND; ++d) un = un + n_[d] * bc.u_inf[d];
    }
    float cur = ia[2] * n_[0];
#pragma unroll
    for (int d = 1; d < ND; ++d) cur = cur + ia[2 + d] * n_[d];
    float a = sqrtf(f.gamma * f.R * clampT(ia[1]));
    float M = fabsf(un) / a;
    float ba[NV];
    bool sup = M > 1.0f;
    ba[0] = un >= 0.0f ? (sup ? bc.p_inf : ia[0]) : (sup ? ia[0] : bc.p_inf);
    ba[1] = un > 0.0f ? bc.T_inf : ia[1];
    if (bc.normal_flow) {
      float corr = un - cur;
#pragma unroll
      for (int d = 0; d < ND; ++d) ba[2 + d] = ia[2 + d] + n_[d] * corr;
    } else {
#pragma unroll
      for (int d = 0; d < ND; ++d) ba[2 + d] = un < 0.0f ? ia[2 + d] : bc.u_inf[d];
    }
    float e = eta[g];
    float pg[NV], qg[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) pg[v] = e * ia[v] + (1.0f - e) * ba[v];
    p2s<ND>(f, pg, qg);
#pragma unroll
    for (int v = 0; v < NV; ++v) stage[(int64_t)v * G + g] = qg[v];
  }
}

__global__ void k_ghost_commit(const float* __restrict__ stage, const int32_t* __restrict__ ghost, float* __restrict__ Q,
                               int64_t N, int64_t G, int nv) {
  int64_t tot = G * nv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = t % G, v = t / G;
    Q[v * N + ghost[g]] = stage[t];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------ host entry points
#define SHAPE(cond, msg) \
  if (!(cond)) return fail(IBX_ERR_ARG, std::string(__func__) + ": shape mismatch: " + (msg))

static int fused_grid(ibx_ctx* c, int64_t n) {
  // whole waves of resident CTAs: multiples of the SM count (148 on B200)
  int64_t g = (n + TB - 1) / TB;
  int64_t wave = (int64_t)c->sm_count * 8;
  if (g > wave) g = ((g + wave - 1) / wave > 4 ? 4 : (g + wave - 1) / wave) * wave;
  return (int)std::max<int64_t>(g, 1);
}

static int check_fused(const ibx_domain& D, const char* fn) {
  if (!D.two_to_one)
    return fail(IBX_ERR_UNSUPPORTED, std::string(fn) + ": mesh has block contacts that are neither same-level nor 2:1; "
                                                        "use the per-operator path");
  if (D.block_size < 2 || (D.block_size & 1))
    return fail(IBX_ERR_UNSUPPORTED, std::string(fn) + ": fused kernels need an even block_size >= 2");
  return IBX_OK;
}

extern "C" {

int ibx_residual_euler(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, ibx_array Qh, ibx_array Rh, ibx_array cflh) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  int rc = check_fused(D, __func__);
  if (rc) return rc;
  GET_ARR(Q, Qh);
  GET_ARR(R, Rh);
  GET_ARR(CF, cflh);
  int nv = D.nd + 2;
  SHAPE(Q.rows == D.ncells && Q.cols == nv && R.rows == D.ncells && R.cols == nv && CF.rows == D.ncells && CF.cols == 1,
        "Q, R must be ncells x (nd + 2); cfl ncells x 1");
  if (flux_kind != 0 && flux_kind != 1) return fail(IBX_ERR_ARG, "ibx_residual_euler: flux_kind must be 0 (HLL) or 1 (sensor-Rusanov)");
  int64_t N = D.ncells;
  float* scratch = ensure_scratch(c, N * (nv + 1));
  if (!scratch) return fail(IBX_ERR_CUDA, "ibx_residual_euler: out of device memory for the primitive/sensor scratch");
  float* P = scratch;
  float* S = scratch + N * nv;
  // tile / marching kernels (one CTA per block, shared-memory staging) for block sizes 8, 4, 2; option "path" = 2 forces
  // the per-cell gather kernels below (the fallback for other block sizes, and an independent cross-check)
  const bool tiles = tile_supported(D) && c->opt_path != 2;
  if (!tiles && (c->opt_arith != 0 || c->opt_sensor != 1))
    return fail(IBX_ERR_UNSUPPORTED, "ibx_residual_euler: the gather kernels implement arithmetic = 0 and sensor = 1 only");
  // a posted halo exchange (ibx_halo_begin without ibx_halo_end): the tile path overlaps it with the conversion of the
  // owned rows when it is Q's; anything else waits for it here
  if (c->halo_pending && (!tiles || c->halo_pending != Qh)) {
    CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    c->halo_pending = 0;
  }
  if (tiles) return residual_euler_tiles(c, D, f, flux_kind, Q.p, P, S, R.p, CF.p);
  Topo T = make_topo(D);
  int g = fused_grid(c, N);
  if (D.nd == 2) {
    k_prim<2><<<g, TB, 0, c->stream>>>(f, Q.p, P, N);
    LAUNCH_CHECK();
    k_sensor<2><<<g, TB, 0, c->stream>>>(T, P, S);
    LAUNCH_CHECK();
    k_euler_flux<2><<<g, TB, 0, c->stream>>>(T, f, flux_kind, P, S, R.p, CF.p);
    LAUNCH_CHECK();
  } else {
    k_prim<3><<<g, TB, 0, c->stream>>>(f, Q.p, P, N);
    LAUNCH_CHECK();
    k_sensor<3><<<g, TB, 0, c->stream>>>(T, P, S);
    LAUNCH_CHECK();
    k_euler_flux<3><<<g, TB, 0, c->stream>>>(T, f, flux_kind, P, S, R.p, CF.p);
    LAUNCH_CHECK();
  }
  return IBX_OK;
}

int ibx_residual_advection(ibx_ctx* c, const ibx_domain* d, ibx_array uh, ibx_array Ch, ibx_array udh, ibx_array spech) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  int rc = check_fused(D, __func__);
  if (rc) return rc;
  GET_ARR(U, uh);
  GET_ARR(C, Ch);
  GET_ARR(UD, udh);
  GET_ARR(SP, spech);
  int64_t N = D.ncells;
  SHAPE(U.rows == N && U.cols == 1 && C.rows == N && C.cols == D.nd && UD.rows == N && UD.cols == 1 && SP.rows == N && SP.cols == 1,
        "u, ud, spec ncells x 1; C ncells x nd");
  float* S = ensure_scratch(c, N);
  if (!S) return fail(IBX_ERR_CUDA, "ibx_residual_advection: out of device memory");
  Topo T = make_topo(D);
  int g = fused_grid(c, N);
  if (D.nd == 2) {
    k_sensor<2><<<g, TB, 0, c->stream>>>(T, U.p, S);
    LAUNCH_CHECK();
    k_advection<2><<<g, TB, 0, c->stream>>>(T, U.p, C.p, S, UD.p, SP.p);
    LAUNCH_CHECK();
  } else {
    k_sensor<3><<<g, TB, 0, c->stream>>>(T, U.p, S);
    LAUNCH_CHECK();
    k_advection<3><<<g, TB, 0, c->stream>>>(T, U.p, C.p, S, UD.p, SP.p);
    LAUNCH_CHECK();
  }
  return IBX_OK;
}

static int ghost_update_on(ibx_ctx* c, const ibx_domain* d, int b, ibx_fluid f, const float* Pinf, int n_pinf, int normal_flow,
                           ibx_array Qh, cudaStream_t st) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  if (b < 0 || b >= (int)D.boundaries.size()) return fail(IBX_ERR_ARG, "ibx_ghost_update_euler: boundary index out of range");
  GET_ARR(Q, Qh);
  int nv = D.nd + 2;
  SHAPE(Q.rows == D.ncells && Q.cols == nv, "Q must be ncells x (nd + 2)");
  if (normal_flow) {
    if (n_pinf != 3) return fail(IBX_ERR_ARG, "Only 3 parcels in P (p, T and normal flow) allowed for normal_flow = true BC");
  } else if (n_pinf != nv) {
    return fail(IBX_ERR_ARG, "ibx_ghost_update_euler: Pinf must hold p, T and nd velocity components");
  }
  BCParams bc{};
  bc.p_inf = Pinf[0];
  bc.T_inf = Pinf[1];
  for (int k = 0; k < n_pinf - 2; ++k) bc.u_inf[k] = Pinf[2 + k];
  bc.normal_flow = normal_flow;
  int64_t Gtot = 0;
  for (auto& B : D.boundaries[b].parts) Gtot += (int64_t)B.ghost.size();
  if (Gtot == 0) return IBX_OK;
  // the staging area lives after the residual scratch so both can coexist
  int64_t N = D.ncells;
  float* scratch = ensure_scratch(c, N * (nv + 1) + Gtot * nv);
  if (!scratch) return fail(IBX_ERR_CUDA, "ibx_ghost_update_euler: out of device memory");
  float* stage = scratch + N * (nv + 1);
  int64_t off = 0;
  // Jacobi across the chunks of one boundary: stage all, then commit all
  for (auto& B : D.boundaries[b].parts) {
    int64_t G = (int64_t)B.ghost.size();
    int g = grid_for(G, 128, c->sm_count, 16);
    if (D.nd == 2) k_ghost_stage<2><<<g, 128, 0, st>>>(f, bc, Q.p, N, B.d_ptr, B.d_idx_global, B.d_w, B.d_normals, B.d_eta, stage + off * nv, G);
    else k_ghost_stage<3><<<g, 128, 0, st>>>(f, bc, Q.p, N, B.d_ptr, B.d_idx_global, B.d_w, B.d_normals, B.d_eta, stage + off * nv, G);
    LAUNCH_CHECK();
    off += G;
  }
  off = 0;
  for (auto& B : D.boundaries[b].parts) {
    int64_t G = (int64_t)B.ghost.size();
    k_ghost_commit<<<grid_for(G * nv, TB, c->sm_count, 16), TB, 0, st>>>(stage + off * nv, B.d_ghost, Q.p, N, G, nv);
    LAUNCH_CHECK();
    off += G;
  }
  return IBX_OK;
}

int ibx_ghost_update_euler(ibx_ctx* c, const ibx_domain* d, int b, ibx_fluid f, const float* Pinf, int n_pinf, int normal_flow,
                           ibx_array Qh) {
  if (!c) return fail(IBX_ERR_ARG, "ibx_ghost_update_euler: null context");
  return ghost_update_on(c, d, b, f, Pinf, n_pinf, normal_flow, Qh, c->stream);
}

// One step of the solver loop -- ghost updates of `bcs`, then the residual -- with everything that is not flux work hidden
// behind compute ("launch interior cells first, boundary cells after halo_end"):
//   halo stream (highest priority):  [exchange(Q) ->] ghost updates of the owned ghosts [-> exchange(Q)]
//   compute stream:                  phase 0 (blocks that read neither a ghost nor a halo cell) ... wait ... phase 1 (the rest)
// The exchanges exist on rank-local shards only.  Results are those of `[halo(Q);] ghost updates; [halo(Q);] ibx_residual_euler`
// bit for bit (tools/mgpu_check.py, tests/test_fused_gpu.py).
static int step_euler_impl(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                           int exchange_between_families, ibx_array Qh, ibx_array Rh, ibx_array cflh) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  int rc;
  if (flux_kind != 0 && flux_kind != 1) return fail(IBX_ERR_ARG, "ibx_step_euler: flux_kind must be 0 (HLL) or 1 (sensor-Rusanov)");
  const bool sharded = D.shard.active && D.shard.nranks > 1;
  const bool overlap = D.phased && march_supported(D) && c->opt_path == 0;
  if (!overlap) {   // same sequence without the phase split (2-D, other block sizes / paths)
    if (sharded) {
      if ((rc = ibx_halo_begin(c, d, Qh))) return rc;
      if ((rc = ibx_halo_end(c, d, Qh))) return rc;
    }
    for (int k = 0; k < nbc; ++k) {
      if (k > 0 && exchange_between_families && sharded) {
        if ((rc = ibx_halo_begin(c, d, Qh))) return rc;
        if ((rc = ibx_halo_end(c, d, Qh))) return rc;
      }
      if ((rc = ibx_ghost_update_euler(c, d, bcs[k].boundary, f, bcs[k].Pinf, bcs[k].n_pinf, bcs[k].normal_flow, Qh))) return rc;
    }
    if (sharded && (rc = ibx_halo_begin(c, d, Qh))) return rc;
    return ibx_residual_euler(c, d, f, flux_kind, Qh, Rh, cflh);
  }
  GET_ARR(Q, Qh);
  GET_ARR(R, Rh);
  GET_ARR(CF, cflh);
  const int nv = D.nd + 2;
  const int64_t N = D.ncells;
  SHAPE(Q.rows == N && Q.cols == nv && R.rows == N && R.cols == nv && CF.rows == N && CF.cols == 1,
        "Q, R must be ncells x (nd + 2); cfl ncells x 1");
  // one scratch allocation for both phases and the ghost staging (no reallocation while kernels are in flight)
  int64_t Gmax = 0;
  for (int k = 0; k < nbc; ++k) {
    if (bcs[k].boundary < 0 || bcs[k].boundary >= (int)D.boundaries.size())
      return fail(IBX_ERR_ARG, "ibx_step_euler: boundary index out of range");
    int64_t G = 0;
    for (auto& B : D.boundaries[bcs[k].boundary].parts) G += (int64_t)B.ghost.size();
    Gmax = std::max(Gmax, G);
  }
  float* scratch = ensure_scratch(c, N * (nv + 1) + Gmax * nv);
  if (!scratch) return fail(IBX_ERR_CUDA, "ibx_step_euler: out of device memory for the scratch arrays");
  float* P = scratch;
  float* S = scratch + N * nv;
  if (c->halo_pending) {   // an exchange posted by the caller: complete it first
    CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    c->halo_pending = 0;
  }
  if (!sharded) {
    // Whole domain: only the ghost update (0.1 ms on C4) is there to hide, and splitting the flux kernels costs about as
    // much in kernel tails.  So only the Q -> P conversion is split: the ghost-free blocks are converted (HBM-bound, 0.3 ms)
    // while the ghost update runs on the second stream, then the blocks holding ghost cells; sensors and fluxes unsplit.
    CU(cudaEventRecord(c->ev_ready, c->stream));                                     // the ghost update sees the writers of Q
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
    for (int k = 0; k < nbc; ++k)
      if ((rc = ghost_update_on(c, d, bcs[k].boundary, f, bcs[k].Pinf, bcs[k].n_pinf, bcs[k].normal_flow, Qh, c->comm_stream))) return rc;
    CU(cudaEventRecord(c->ev_halo, c->comm_stream));
    if ((rc = prim_blocks(c, D, true, f, Q.p, P))) return rc;
    CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    if ((rc = prim_blocks(c, D, false, f, Q.p, P))) return rc;
    return residual_euler_after_prim(c, D, f, flux_kind, Q.p, P, S, R.p, CF.p);
  }
  if ((rc = halo_begin_impl(c, d, Qh, true))) return rc;                           // exchange 1 (waits for the writers of Q)
  if ((rc = residual_euler_phase(c, D, 0, f, flux_kind, Q.p, P, S, R.p, CF.p))) return rc;   // phase 0 under it
  for (int k = 0; k < nbc; ++k) {                                                    // ghost updates on the halo stream
    // a ghost of this family may read, on another rank, a ghost of a family applied above (Domain.shard reports it)
    if (k > 0 && exchange_between_families && (rc = halo_begin_impl(c, d, Qh, false))) return rc;
    if ((rc = ghost_update_on(c, d, bcs[k].boundary, f, bcs[k].Pinf, bcs[k].n_pinf, bcs[k].normal_flow, Qh, c->comm_stream))) return rc;
  }
  if ((rc = halo_begin_impl(c, d, Qh, false))) return rc;                          // exchange 2 behind them
  CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  c->halo_pending = 0;
  return residual_euler_phase(c, D, 1, f, flux_kind, Q.p, P, S, R.p, CF.p);
}

int ibx_step_euler(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs, ibx_array Qh, ibx_array Rh,
                   ibx_array cflh) {
  return step_euler_impl(c, d, f, flux_kind, nbc, bcs, 0, Qh, Rh, cflh);
}

// Pseudo-time march with local time steps, whole loop on the device (the driver a user writes around `FAS!` or by hand,
// test/advection.jl:28-46): per step  ghost updates in place;  Q0 = Q;  for every stage coefficient a:
// ibx_step_euler (ghost updates + residual), Q = Q0 + ((a CFL / cfl) R) live.  On small meshes the loop is launch-bound
// (C3: ~40 launches of a few microseconds per stage), so one step is captured into a CUDA graph after a warm-up step
// has sized every scratch array, and replayed; the kernels and their order are those of the plain loop: same bits.
int ibx_march_euler(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs, ibx_array Qh,
                    ibx_array liveh, int64_t n_steps, float CFL, int nstages, const float* alphas, int use_graph, int* graph_used) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  GET_ARR(Q, Qh);
  if (graph_used) *graph_used = 0;
  if (nstages < 1 || nstages > 8 || !alphas) return fail(IBX_ERR_ARG, "ibx_march_euler: 1 .. 8 stage coefficients");
  if (n_steps < 0) return fail(IBX_ERR_ARG, "ibx_march_euler: negative step count");
  if (D.shard.active && D.shard.nranks > 1) return fail(IBX_ERR_UNSUPPORTED, "ibx_march_euler: whole domains only (use ibx_step_euler_sharded in a host loop)");
  const int nv = D.nd + 2;
  SHAPE(Q.rows == D.ncells && Q.cols == nv, "Q must be ncells x (nd + 2)");
  if (n_steps == 0) return IBX_OK;
  int rc;
  ibx_array Q0h = 0, Rh = 0, cfh = 0;
  if ((rc = ibx_array_alloc(c, Q.rows, nv, &Q0h)) || (rc = ibx_array_alloc(c, Q.rows, nv, &Rh)) || (rc = ibx_array_alloc(c, Q.rows, 1, &cfh))) {
    if (Q0h) ibx_array_free(c, Q0h);
    if (Rh) ibx_array_free(c, Rh);
    return rc;
  }
  auto one_step = [&]() -> int {
    int r;
    for (int k = 0; k < nbc; ++k)
      if ((r = ibx_ghost_update_euler(c, d, bcs[k].boundary, f, bcs[k].Pinf, bcs[k].n_pinf, bcs[k].normal_flow, Qh))) return r;
    if ((r = ibx_array_copy(c, Q0h, Qh))) return r;
    for (int s = 0; s < nstages; ++s) {
      if ((r = step_euler_impl(c, d, f, flux_kind, nbc, bcs, 0, Qh, Rh, cfh))) return r;
      if ((r = ibx_local_step_update(c, Q0h, Rh, cfh, liveh, alphas[s] * CFL, Qh))) return r;
    }
    return IBX_OK;
  };
  auto done = [&](int code) {
    ibx_array_free(c, Q0h);
    ibx_array_free(c, Rh);
    ibx_array_free(c, cfh);
    return code;
  };
  int64_t left = n_steps;
  if ((rc = one_step())) return done(rc);        // also the warm-up that sizes the scratch arrays and sets kernel attributes
  --left;
  if (use_graph && left > 0) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const bool was_poisoned = c->poisoned;   // an error raised while capturing is not a sticky device error
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      const int rcap = one_step();
      const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
      if (rcap == IBX_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        for (; left > 0; --left)
          if (cudaGraphLaunch(exec, c->stream) != cudaSuccess) break;
        if (graph_used) *graph_used = left == 0;
      }
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
      if (left > 0) {
        cudaGetLastError();
        c->poisoned = was_poisoned;
        return done(fail(IBX_ERR_CUDA, "ibx_march_euler: CUDA graph capture / replay of one step failed (call again with use_graph = 0)"));
      }
    } else {
      cudaGetLastError();
      return done(fail(IBX_ERR_CUDA, "ibx_march_euler: cudaStreamBeginCapture failed (call again with use_graph = 0)"));
    }
  }
  for (; left > 0; --left)
    if ((rc = one_step())) return done(rc);
  return done(IBX_OK);
}

int ibx_step_euler_sharded(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                           int exchange_between_families, ibx_array Qh, ibx_array Rh, ibx_array cflh) {
  if (c) {
    ibx_domain* Dp = find_domain(d);
    if (Dp && !Dp->shard.active) return fail(IBX_ERR_STATE, "ibx_step_euler_sharded: domain is not a rank-local shard (ibx_domain_shard)");
  }
  return step_euler_impl(c, d, f, flux_kind, nbc, bcs, exchange_between_families, Qh, Rh, cflh);
}

static int e2e_slot_prepare(ibx_ctx* c, ibx_ctx::E2ESlot& S, int64_t N, int nv) {
  int rc;
  if (!c->h2d_stream) {
    CU(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  }
  if (!S.up) {
    CU(cudaEventCreateWithFlags(&S.up, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&S.down, cudaEventDisableTiming));
  }
  if (S.Q) {
    ibx_ctx::Arr probe;
    if (get_array(c, S.Q, probe) && (probe.rows != N || probe.cols != nv)) {
      ibx_array_free(c, S.Q);
      ibx_array_free(c, S.R);
      ibx_array_free(c, S.cfl);
      S.Q = 0;
    }
  }
  if (!S.Q) {
    if ((rc = ibx_array_alloc(c, N, nv, &S.Q))) return rc;
    if ((rc = ibx_array_alloc(c, N, nv, &S.R))) return rc;
    if ((rc = ibx_array_alloc(c, N, 1, &S.cfl))) return rc;
    // ibx_array_alloc clears the new arrays on the compute stream; the upload below runs on h2d_stream and must not be
    // overtaken by that memset (it may sit behind the other slot's residual): finish it before the slot is used.
    CU(cudaStreamSynchronize(c->stream));
  }
  return IBX_OK;
}

int ibx_euler_step_host_begin(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                              const float* Q_host, float* R_host, float* cfl_host, int slot) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  if (slot < 0 || slot > 1) return fail(IBX_ERR_ARG, "ibx_euler_step_host_begin: slot must be 0 or 1");
  ibx_ctx::E2ESlot& S = c->e2e[slot];
  if (S.busy) return fail(IBX_ERR_STATE, "ibx_euler_step_host_begin: slot still in flight (call ibx_euler_step_host_end first)");
  int nv = D.nd + 2;
  int64_t N = D.ncells;
  int rc;
  if ((rc = e2e_slot_prepare(c, S, N, nv))) return rc;
  GET_ARR(Q, S.Q);
  GET_ARR(R, S.R);
  GET_ARR(CF, S.cfl);
  // H2D on its own stream -> compute stream -> D2H on its own stream, chained by events
  CU(cudaMemcpyAsync(Q.p, Q_host, (size_t)N * nv * sizeof(float), cudaMemcpyHostToDevice, c->h2d_stream));
  CU(cudaEventRecord(S.up, c->h2d_stream));
  CU(cudaStreamWaitEvent(c->stream, S.up, 0));
  // a failure below must not leave the upload in flight behind a slot the caller believes idle (the host buffer is only
  // borrowed until the matching _end): drain the copy stream before reporting
  auto bail = [&](int code) { cudaStreamSynchronize(c->h2d_stream); return code; };
  if ((rc = step_euler_impl(c, d, f, flux_kind, nbc, bcs, 0, S.Q, S.R, S.cfl))) return bail(rc);
  CU(cudaEventRecord(S.done, c->stream));
  CU(cudaStreamWaitEvent(c->d2h_stream, S.done, 0));
  CU(cudaMemcpyAsync(R_host, R.p, (size_t)N * nv * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream));
  CU(cudaMemcpyAsync(cfl_host, CF.p, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream));
  CU(cudaEventRecord(S.down, c->d2h_stream));
  S.busy = true;
  return IBX_OK;
}

int ibx_euler_step_host_end(ibx_ctx* c, int slot) {
  CHECK_CTX(c);
  if (slot < 0 || slot > 1) return fail(IBX_ERR_ARG, "ibx_euler_step_host_end: slot must be 0 or 1");
  ibx_ctx::E2ESlot& S = c->e2e[slot];
  if (!S.busy) return IBX_OK;
  CU(cudaEventSynchronize(S.down));
  S.busy = false;
  return IBX_OK;
}

int ibx_euler_step_host(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                        const float* Q_host, float* R_host, float* cfl_host) {
  int rc;
  if ((rc = ibx_euler_step_host_end(c, 0))) return rc;
  if ((rc = ibx_euler_step_host_begin(c, d, f, flux_kind, nbc, bcs, Q_host, R_host, cfl_host, 0))) return rc;
  return ibx_euler_step_host_end(c, 0);
}

}  // extern "C"
