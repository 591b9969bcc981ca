// Tile kernels of the fused Euler residual: one CTA per octree block, everything staged in shared memory.
//
// Layout in HBM: every field is column-major N x nv, i.e. one contiguous plane per variable, and cell ids are
// block-major with the first dimension fastest inside a block (src/mesher.jl:1079-1092).  A block's BS^ND cells of
// one variable are therefore ONE contiguous run (2 KB for 8^3 floats): own cells load fully coalesced; halo cells
// come in runs of BS along x.
//
// Per block the CTA stages its own cells plus, for each of the 2*ND block faces, the first TWO cell layers of the
// neighbouring block(s) at THEIR resolution (same level: BS^(ND-1) x 2, coarser: (BS/2)^(ND-1) x 2, finer:
// (2 BS)^(ND-1) x 2).  That is the reference's 2-deep skirt (src/ImmersedBoundary.jl:610-619) of a partition whose
// image is the block, restricted to what a dimension-split residual reads: the gradient of a neighbour cell along
// the face normal needs (a) this block's cells and (b) the neighbour's second layer -- both staged.  The JST sensor
// also needs lateral neighbours of the neighbour cells, so it is produced by its own tile pass (k_tile_sensor) and
// staged here as a sixth field.
// Work per dimension: gradients -> face fluxes (each face ONCE: MUSCL + HLL) -> divergence; three barriers.
#include "device.cuh"
#include "physics.cuh"
#include "tile_common.cuh"

using namespace ibx;
using namespace ibxk;

namespace {

template <int ND, int BS, bool FINER>
struct Cfg {
  static constexpr int NV = ND + 2;
  static constexpr int CPB = ND == 3 ? BS * BS * BS : BS * BS;
  static constexpr int FACE = ND == 3 ? BS * BS : BS;
  static constexpr int MAXL1 = FINER ? FACE * (ND == 3 ? 4 : 2) : FACE;  // cells per halo layer == faces per block face
  static constexpr int NFACES = 2 * ND;
  static constexpr int NS = CPB + NFACES * 2 * MAXL1;                    // staged cells (own + 2-layer halos)
  static constexpr int NG = CPB + 2 * MAXL1;                             // gradient / flux slots per dimension
  static constexpr int NT = CPB >= 512 ? 256 : (CPB >= 64 ? 64 : 32);
  static constexpr int CPT = (CPB + NT - 1) / NT;
  // doubles first (8-byte alignment): face fluxes [NV][NG]; then floats: P [NV][NS], D [NS], CFL term [NG]
  static constexpr size_t SMEM_SENSOR = sizeof(float) * ((size_t)CPB + NFACES * MAXL1);
};


// ------------------------------------------------------------------------------------------ sensor pass
// D = JST_sensor(part, p) with dim = 0 (src/ImmersedBoundary.jl:1077-1097) for the cells of one block
template <int ND, int BS, bool FINER, bool P2 = false>
__global__ void __launch_bounds__(Cfg<ND, BS, FINER>::NT)
k_tile_sensor(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh,
              const float* __restrict__ p, float* __restrict__ D) {
  using C = Cfg<ND, BS, FINER>;
  extern __shared__ float smem[];
  __shared__ FaceInfo fi[C::NFACES];
  float* sp = smem;  // [CPB + NFACES * MAXL1]
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid < C::NFACES) fill_face_info<ND, BS>(fi[tid], faces[b * C::NFACES + tid], C::CPB + tid * C::MAXL1, 1.0f);
  __syncthreads();
  const int64_t cell0 = b * C::CPB;
  for (int l = tid; l < C::CPB; l += C::NT) sp[l] = p[cell0 + l];
#pragma unroll
  for (int f = 0; f < C::NFACES; ++f) {
    const FaceInfo& F = fi[f];
    if (F.kind == 0) continue;
    int d = f >> 1, side = f & 1, n = F.n1 * F.n2;
    for (int k = tid; k < n; k += C::NT) sp[F.base + k] = p[halo_cell<ND, BS>(F, d, side, k % F.n1, k / F.n1, 0, C::CPB)];
  }
  __syncthreads();
  float h[ND], ih[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) { h[d] = bh[b * ND + d]; ih[d] = 1.0f / h[d]; }
  for (int l = tid; l < C::CPB; l += C::NT) {
    int ii[3];
    split<ND, BS>(l, ii);
    float pc = sp[l], nu = 1e-7f;
    int stride = 1;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float g[2], a[2];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        float accg = 0.0f, acca = 0.0f;
        bool inner = side ? ii[d] < BS - 1 : ii[d] > 0;
        if (inner) {
          float fd = side ? sp[l + stride] - pc : pc - sp[l - stride];  // p_neighbour - p_owner
          accg = fd;
          acca = fabsf(fd);
        } else {
          const FaceInfo& F = fi[2 * d + side];
          if (F.kind != 0) {  // box face: owner == neighbour, difference 0
            int slot[4];
            int cnt = own_to_halo<ND, BS>(F, ii[T1(d)], ND == 3 ? ii[T2(d)] : 0, slot);
            float w = 1.0f / (float)cnt;
            for (int q = 0; q < cnt; ++q) {
              float pn = sp[F.base + slot[q]];
              float fd = side ? pn - pc : pc - pn;
              accg = q == 0 ? fd * w : accg + fd * w;
              acca = q == 0 ? fabsf(fd) * w : acca + fabsf(fd) * w;
            }
          }
        }
        g[side] = accg;
        a[side] = acca;
      }
      // P2: h is a power of two, x / h == x * (1 / h) bit for bit (physics.cuh)
      float gg = P2 ? (g[1] - g[0]) * ih[d] : (g[1] - g[0]) / h[d], ugg = P2 ? (a[1] + a[0]) * ih[d] : (a[1] + a[0]) / h[d];
      nu = fmaxf(nu, (1e-7f + fabsf(gg)) / (1e-7f + ugg));
      stride *= BS;
    }
    D[cell0 + l] = nu;
  }
}

// ------------------------------------------------------------------------------------------ regular blocks
// Lean flux kernel for blocks whose 2*ND neighbours are all same-level blocks (80 % of the C4 mesh): the staged tile
// is a plain (BS+4)^ND padded array, every face of a dimension is one item of one uniform loop, no per-face
// descriptors, no branches.  Same arithmetic, same order, same bits as k_tile_flux; compact enough for the 32 KB
// instruction cache, which the general kernel is not.
template <int ND, int BS>
struct RegCfg {
  static constexpr int NV = ND + 2;
  static constexpr int PAD = BS + 4;
  static constexpr int TS = ND == 3 ? PAD * PAD * PAD : PAD * PAD;           // padded tile slots
  static constexpr int CPB = ND == 3 ? BS * BS * BS : BS * BS;
  static constexpr int FACE = ND == 3 ? BS * BS : BS;
  static constexpr int NFD = (BS + 1) * FACE;                                // faces per dimension
  static constexpr int NT = (ND == 3 && BS == 8) ? 192 : (CPB >= 64 ? 64 : 32);
  static constexpr int CPT = (CPB + NT - 1) / NT;
  static constexpr size_t SMEM = sizeof(double) * (size_t)NV * NFD + sizeof(float) * ((size_t)(NV + 1) * TS + NFD);
};

// it -> (c0, c1, c2) in the face index space of dimension d (extent BS + 1 along d, BS elsewhere); written with
// compile-time divisors per d (a runtime divisor costs ~20 instructions per division)
template <int ND, int BS>
__device__ __forceinline__ void face_decode(int it, int d, int (&cc)[3]) {
  if (d == 0) {
    cc[0] = it % (BS + 1); cc[1] = (it / (BS + 1)) % BS; cc[2] = ND == 3 ? it / ((BS + 1) * BS) : 0;
  } else if (d == 1) {
    cc[0] = it % BS; cc[1] = (it / BS) % (BS + 1); cc[2] = ND == 3 ? it / (BS * (BS + 1)) : 0;
  } else {
    cc[0] = it % BS; cc[1] = (it / BS) % BS; cc[2] = it / (BS * BS);
  }
}

template <int ND, int BS, bool P2, int FLUX>
__global__ void __launch_bounds__(RegCfg<ND, BS>::NT, (ND == 3 && BS == 8) ? 3 : 1)
k_reg_flux(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh, int64_t N,
           ibx_fluid fl, const float* __restrict__ P, const float* __restrict__ Dg, float* __restrict__ R,
           float* __restrict__ cfl) {
  using C = RegCfg<ND, BS>;
  constexpr int NV = C::NV, PAD = C::PAD, TS = C::TS, CPB = C::CPB, NT = C::NT, NFD = C::NFD, FACE = C::FACE;
  extern __shared__ double smem_d[];
  double* sF = smem_d;                  // [NV][NFD] fluxes of the current dimension
  float* sP = (float*)(sF + NV * NFD);  // [NV][TS]
  float* sD = sP + NV * TS;             // [TS]
  float* sC = sD + TS;                  // [NFD]
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  const int64_t cell0 = b * CPB;
  // ---- stage: own cells, then two layers of each same-level neighbour
  for (int l = tid; l < CPB; l += NT) {
    int ii[3];
    split<ND, BS>(l, ii);
    int s = (ii[0] + 2) + PAD * ((ii[1] + 2) + (ND == 3 ? PAD * (ii[2] + 2) : 0));
#pragma unroll
    for (int v = 0; v < NV; ++v) sP[v * TS + s] = P[(int64_t)v * N + cell0 + l];
    sD[s] = Dg[cell0 + l];
  }
  for (int k = tid; k < 2 * ND * 2 * FACE; k += NT) {
    int f = k / (2 * FACE), r = k - f * 2 * FACE;
    int layer = r / FACE, q = r - layer * FACE;
    int d = f >> 1, side = f & 1;
    int j1 = q % BS, j2 = q / BS;
    int64_t nb = faces[b * (2 * ND) + f].nb[0];
    int64_t c = nb * CPB + compose<ND, BS>(d, side ? layer : BS - 1 - layer, j1, j2);
    int cc[3] = {0, 0, 0};
    cc[d] = side ? BS + layer : -1 - layer;
    cc[T1(d)] = j1;
    if (ND == 3) cc[T2(d)] = j2;
    int s = (cc[0] + 2) + PAD * ((cc[1] + 2) + (ND == 3 ? PAD * (cc[2] + 2) : 0));
#pragma unroll
    for (int v = 0; v < NV; ++v) sP[v * TS + s] = P[(int64_t)v * N + c];
    sD[s] = Dg[c];
  }
  __syncthreads();
  float res[C::CPT][NV], cf[C::CPT];
#pragma unroll
  for (int q = 0; q < C::CPT; ++q) {
    cf[q] = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) res[q][v] = 0.0f;
  }
  const float gr = fl.gamma * fl.R;
#pragma unroll 1   // keep the body once: the loop must stay inside the 32 KB instruction cache
  for (int d = 0; d < ND; ++d) {
    const float hd = bh[b * ND + d];
    const float inv_hd = 1.0f / hd;
    const double inv_hd_d = 1.0 / (double)hd;
    constexpr int n0 = BS, n1 = BS;
    const int m0 = d == 0 ? BS + 1 : BS, m1 = d == 1 ? BS + 1 : BS;   // extents of the face index space
    const int ss = d == 0 ? 1 : (d == 1 ? PAD : PAD * PAD);           // tile stride along d
    (void)n0; (void)n1;
    // ---- (1) every face normal to d once: owner o = cell (c_d - 1), neighbour n = cell c_d
    for (int it = tid; it < NFD; it += NT) {
      int cc[3];
      face_decode<ND, BS>(it, d, cc);
      int so = (cc[0] + 2) + PAD * ((cc[1] + 2) + (ND == 3 ? PAD * (cc[2] + 2) : 0)) - ss;  // owner sits one step below along d
      int sn = so + ss;
      float po[NV], pn[NV], g0[NV], g1[NV], pl[NV], pr[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float um = sP[v * TS + so - ss], uo = sP[v * TS + so], un = sP[v * TS + sn], up = sP[v * TS + sn + ss];
        float fm = face_interp_f(um, uo, hd, hd, P2), fc = face_interp_f(uo, un, hd, hd, P2), fp = face_interp_f(un, up, hd, hd, P2);
        g0[v] = P2 ? (fc - fm) * inv_hd : (fc - fm) / hd;
        g1[v] = P2 ? (fp - fc) * inv_hd : (fp - fc) / hd;
        po[v] = uo;
        pn[v] = un;
      }
      float Do = sD[so], Dn = sD[sn];
      muscl_face<NV>(po, pn, g0, g1, hd, hd, Do, Dn, true, false, pl, pr, P2);
      double F_[NV];
      if (FLUX == 0) {
        hll_flux<ND>(fl, pl, pr, d, F_);
      } else {
        float Ff[NV];
        rusanov_flux<ND>(fl, pl, pr, face_interp(Do, Dn, hd, hd), d, Ff);
#pragma unroll
        for (int v = 0; v < NV; ++v) F_[v] = (double)Ff[v];
      }
      float ao = sqrtf(gr * clampT(po[1])), an = sqrtf(gr * clampT(pn[1]));
      float ct = fabsf(face_interp_f(pick<ND>(po + 2, d), pick<ND>(pn + 2, d), hd, hd, P2)) + face_interp_f(ao, an, hd, hd, P2);
#pragma unroll
      for (int v = 0; v < NV; ++v) sF[v * NFD + it] = F_[v];
      sC[it] = ct;
    }
    __syncthreads();
    // ---- (2) divergence
#pragma unroll
    for (int q = 0; q < C::CPT; ++q) {
      int l = tid + q * NT;
      if (l < CPB) {
        int ii[3];
        split<ND, BS>(l, ii);
        int lo = ii[0] + m0 * (ii[1] + (ND == 3 ? m1 * ii[2] : 0));      // low face: c_d = i_d
        int hi = lo + (d == 0 ? 1 : (d == 1 ? m0 : m0 * m1));            // high face: c_d = i_d + 1
        if (FLUX == 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            res[q][v] = (float)((double)res[q][v] - (sF[v * NFD + hi] - sF[v * NFD + lo]) * inv_hd_d);
        } else {
#pragma unroll
          for (int v = 0; v < NV; ++v) res[q][v] = res[q][v] - ((float)sF[v * NFD + hi] - (float)sF[v * NFD + lo]) / hd;
        }
        float cs = sC[hi] + sC[lo];
        cf[q] = cf[q] + (P2 ? cs * inv_hd : cs / hd);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < C::CPT; ++q) {
    int l = tid + q * NT;
    if (l < CPB) {
#pragma unroll
      for (int v = 0; v < NV; ++v) R[(int64_t)v * N + cell0 + l] = res[q][v];
      cfl[cell0 + l] = cf[q];
    }
  }
}

template <int ND, int BS, bool P2>
int launch_reg(ibx_ctx* c, const ibx_domain& D, ibx_fluid f, int flux_kind, const float* P, const float* S, float* R, float* cfl) {
  using C = RegCfg<ND, BS>;
  if (D.n_own_regular == 0) return IBX_OK;
  int rc;
  if ((rc = ensure_dyn_smem(c, (const void*)k_reg_flux<ND, BS, P2, 0>, C::SMEM))) return rc;
  if ((rc = ensure_dyn_smem(c, (const void*)k_reg_flux<ND, BS, P2, 1>, C::SMEM))) return rc;
  if (flux_kind == 0)
    k_reg_flux<ND, BS, P2, 0><<<D.n_own_regular, C::NT, C::SMEM, c->stream>>>(D.d_blk_own_regular, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, R, cfl);
  else
    k_reg_flux<ND, BS, P2, 1><<<D.n_own_regular, C::NT, C::SMEM, c->stream>>>(D.d_blk_own_regular, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, R, cfl);
  LAUNCH_CHECK();
  return IBX_OK;
}

// ------------------------------------------------------------------------------------------ irregular blocks
// Hybrid flux kernel for blocks with at least one face that is not a same-level contact (domain box, coarser or finer
// neighbour).  Same padded tile and the same uniform face loop as k_reg_flux; only the faces whose 4-cell stencil
// touches an irregular block face (the block face itself and the first internal face behind it) take the general
// path, which reads the coarser / finer halo cells from a separate staging area at their own resolution.
template <int ND, int BS, bool FINER>
struct HybCfg {
  using R = RegCfg<ND, BS>;
  static constexpr int NV = ND + 2;
  static constexpr int FACE = R::FACE;
  static constexpr int MAXL = FINER ? FACE * (ND == 3 ? 4 : 2) : (FACE / (ND == 3 ? 4 : 2) > 0 ? FACE / (ND == 3 ? 4 : 2) : 1);  // cells per irregular halo layer
  static constexpr int NI = 2 * ND * 2 * MAXL;                 // irregular halo cells (2 layers per face)
  static constexpr int NX = FINER ? 2 * FACE * (ND == 3 ? 4 : 2) : 0;  // extra flux slots: fine faces of the two block faces of a dimension
  static constexpr int NSLOT = R::TS + NI;                     // staged cells: padded tile, then irregular halos
  static constexpr int NFS = R::NFD + NX;
  static constexpr int NT = R::NT;
  static constexpr size_t SMEM = sizeof(double) * (size_t)NV * NFS + sizeof(float) * ((size_t)(NV + 1) * NSLOT + NFS);
  static constexpr size_t SMEM_GENERAL = sizeof(float) * ((size_t)(NV + 1) * NSLOT);   // MODE 1: fluxes go to global memory
};

template <int ND, int BS>
__device__ __forceinline__ int pslot(const int (&ii)[3]) {
  constexpr int PAD = BS + 4;
  return (ii[0] + 2) + PAD * ((ii[1] + 2) + (ND == 3 ? PAD * (ii[2] + 2) : 0));
}
template <int ND, int BS>
__device__ __forceinline__ int pslot_c(int d, int cn, int c1, int c2) {
  int ii[3] = {0, 0, 0};
  ii[d] = cn;
  ii[T1(d)] = c1;
  if (ND == 3) ii[T2(d)] = c2;
  return pslot<ND, BS>(ii);
}

// gradient along d at own cell ii (general: any kind of block face on either side)
template <int ND, int BS, bool FINER, int NV, int NS>
__device__ __forceinline__ void hyb_own_grad(const float* __restrict__ sP, const FaceInfo& FL, const FaceInfo& FH, const int (&ii)[3],
                                             int d, int ss, float hd, bool p2, float inv_hd, float* g) {
  const int l = pslot<ND, BS>(ii);
  float m[2][NV];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const FaceInfo& F = side ? FH : FL;
    bool inner = side ? ii[d] < BS - 1 : ii[d] > 0;
    if (inner || F.kind == 1) {   // same-level neighbour: own cell or padded halo slab
      int n = side ? l + ss : l - ss;
#pragma unroll
      for (int v = 0; v < NV; ++v) m[side][v] = face_interp_f(sP[v * NS + l], sP[v * NS + n], hd, hd, p2);
    } else if (F.kind == 0) {      // box face: owner == neighbour == this cell
#pragma unroll
      for (int v = 0; v < NV; ++v) m[side][v] = face_interp_f(sP[v * NS + l], sP[v * NS + l], hd, hd, p2);
    } else if (!FINER || F.kind == 2) {
      int a1 = ii[T1(d)], a2 = ND == 3 ? ii[T2(d)] : 0;
      int n = F.base + (a2 >> 1) * F.n1 + (a1 >> 1);
#pragma unroll
      for (int v = 0; v < NV; ++v) m[side][v] = face_interp(sP[v * NS + l], sP[v * NS + n], hd, F.hn);
    } else {
      int slot[4];
      int cnt = own_to_halo<ND, BS>(F, ii[T1(d)], ND == 3 ? ii[T2(d)] : 0, slot);
      float w = 1.0f / (float)cnt;
      for (int q = 0; q < cnt; ++q) {
        int n = F.base + slot[q];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float fv = face_interp(sP[v * NS + l], sP[v * NS + n], hd, F.hn);
          m[side][v] = q == 0 ? fv * w : m[side][v] + fv * w;
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) g[v] = p2 ? (m[1][v] - m[0][v]) * inv_hd : (m[1][v] - m[0][v]) / hd;
}

// gradient along d at coarser / finer halo cell r (layer 0) of face (d, side)
template <int ND, int BS, bool FINER, int NV, int NS>
__device__ __forceinline__ void hyb_halo_grad(const float* __restrict__ sP, const FaceInfo& F, int side, int r, int d, float hd,
                                              bool p2, float* g) {
  int j1 = r % F.n1, j2 = r / F.n1;
  int n = F.base + r, far = n + F.n1 * F.n2;
  float hc = F.hn;
  int bnd = side ? BS - 1 : 0;
  float nearv[NV], farv[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) farv[v] = face_interp_f(sP[v * NS + n], sP[v * NS + far], hc, hc, p2);
  if (FINER && F.kind == 3) {
    int o = pslot_c<ND, BS>(d, bnd, j1 >> 1, j2 >> 1);
#pragma unroll
    for (int v = 0; v < NV; ++v) nearv[v] = face_interp(sP[v * NS + n], sP[v * NS + o], hc, hd);
  } else {
    constexpr int CNT = ND == 3 ? 4 : 2;
    const float w = 1.0f / (float)CNT;
#pragma unroll
    for (int q = 0; q < CNT; ++q) {
      int o = pslot_c<ND, BS>(d, bnd, 2 * j1 + (q & 1), 2 * j2 + (q >> 1));
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float fv = face_interp(sP[v * NS + n], sP[v * NS + o], hc, hd);
        nearv[v] = q == 0 ? fv * w : nearv[v] + fv * w;
      }
    }
  }
  const float invc = 1.0f / hc;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float df = side ? farv[v] - nearv[v] : nearv[v] - farv[v];
    g[v] = p2 ? df * invc : df / hc;
  }
}

// ---- compact general path: the neighbours of a cell on one side as a short slot list, one gradient routine
struct NL {
  int cnt;
  int s[4];
  float h;
};

template <int ND, int BS, bool FINER>
__device__ __forceinline__ NL nl_own(const FaceInfo& F, const int (&ii)[3], int l, int d, int side, int ss, float hd) {
  NL r;
  r.cnt = 1;
  r.h = hd;
  r.s[1] = r.s[2] = r.s[3] = 0;
  bool inner = side ? ii[d] < BS - 1 : ii[d] > 0;
  if (inner || F.kind == 1) {
    r.s[0] = side ? l + ss : l - ss;
  } else if (F.kind == 0) {
    r.s[0] = l;                       // box face: owner == neighbour == this cell
  } else if (!FINER || F.kind == 2) {
    int a1 = ii[T1(d)], a2 = ND == 3 ? ii[T2(d)] : 0;
    r.s[0] = F.base + (a2 >> 1) * F.n1 + (a1 >> 1);
    r.h = F.hn;
  } else {
    int slot[4];
    r.cnt = own_to_halo<ND, BS>(F, ii[T1(d)], ND == 3 ? ii[T2(d)] : 0, slot);
    for (int q = 0; q < r.cnt; ++q) r.s[q] = F.base + slot[q];
    r.h = F.hn;
  }
  return r;
}

// near side of halo cell r (layer 0) of face (d, side): the own cells facing it
template <int ND, int BS, bool FINER>
__device__ __forceinline__ NL nl_halo_near(const FaceInfo& F, int side, int r, int d, float hd) {
  NL o;
  o.h = hd;
  o.s[1] = o.s[2] = o.s[3] = 0;
  int j1 = r % F.n1, j2 = r / F.n1;
  int bnd = side ? BS - 1 : 0;
  if (FINER && F.kind == 3) {
    o.cnt = 1;
    o.s[0] = pslot_c<ND, BS>(d, bnd, j1 >> 1, j2 >> 1);
  } else {
    o.cnt = ND == 3 ? 4 : 2;
    for (int q = 0; q < o.cnt; ++q) o.s[q] = pslot_c<ND, BS>(d, bnd, 2 * j1 + (q & 1), 2 * j2 + (q >> 1));
  }
  return o;
}

// Green-Gauss gradient along one dimension of the NV staged variables at slot c (spacing hc), given the slot lists
// of its low and high sides; weights 1/len, products first, sum in list order (src/accumulator.jl:95-106)
template <int NV, int NS>
__device__ __noinline__ void grad_lists(const float* __restrict__ sP, int c, float hc, NL lo, NL hi, int p2, float* __restrict__ g) {
  float m[2][NV];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const NL& L = side ? hi : lo;
    const float w = 1.0f / (float)L.cnt;
    const bool fast = p2 && L.h == hc;
    for (int q = 0; q < L.cnt; ++q) {
      const int n = L.s[q];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float fv = face_interp_f(sP[v * NS + c], sP[v * NS + n], hc, L.h, fast);
        m[side][v] = q == 0 ? fv * w : m[side][v] + fv * w;
      }
    }
  }
  const float inv = 1.0f / hc;
#pragma unroll
  for (int v = 0; v < NV; ++v) g[v] = p2 ? (m[1][v] - m[0][v]) * inv : (m[1][v] - m[0][v]) / hc;
}

// MODE 0: everything in one kernel.  MODE 1: only the general faces, fluxes written to a global scratch (GF, GC).
// MODE 2: the uniform faces (lean loop), general fluxes read back from the scratch, divergence.  Splitting keeps
// each kernel's working set of instructions near the 32 KB instruction cache (ncu: 32 % of the samples of the
// one-kernel form were instruction-fetch stalls).
template <int ND, int BS, bool FINER, bool P2, int FLUX, int MODE>
__global__ void __launch_bounds__(HybCfg<ND, BS, FINER>::NT, (ND == 3 && BS == 8 && !FINER) ? 3 : 1)
k_hyb_flux(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh, int64_t N,
           ibx_fluid fl, const float* __restrict__ P, const float* __restrict__ Dg, float* __restrict__ R,
           float* __restrict__ cfl, double* __restrict__ GF, float* __restrict__ GC) {
  using C = HybCfg<ND, BS, FINER>;
  using RC = RegCfg<ND, BS>;
  constexpr int NV = C::NV, PAD = RC::PAD, TS = RC::TS, CPB = RC::CPB, NT = C::NT, NFD = RC::NFD, FACE = RC::FACE;
  constexpr int NS = C::NSLOT, NFS = C::NFS, MAXL = C::MAXL;
  constexpr int NSL = 4 * FACE + C::NX;  // scratch slots per (block, dimension): 2 x 2 x FACE general faces + fine faces
  extern __shared__ double smem_d[];
  __shared__ FaceInfo fi[2 * ND];
  double* sF = smem_d;                  // [NV][NFS] (absent in MODE 1)
  float* sP = (float*)(sF + (MODE == 1 ? 0 : NV * NFS));  // [NV][NS]
  float* sD = sP + NV * NS;             // [NS]
  float* sC = sD + NS;                  // [NFS]
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  const int64_t cell0 = b * CPB;
  if (tid < 2 * ND) fill_face_info<ND, BS>(fi[tid], faces[b * (2 * ND) + tid], TS + tid * 2 * MAXL, bh[b * ND + (tid >> 1)]);
  // ---- stage own cells into the padded tile (float4 along x when a row has 4+ cells; all loads before the stores)
  if constexpr (BS % 4 == 0) {
    constexpr int NQ = (NV + 1) * CPB / 4, PERQ = (NQ + NT - 1) / NT;
    float4 buf[PERQ];
#pragma unroll
    for (int k = 0; k < PERQ; ++k) {
      const int it = tid + k * NT;
      if (it < NQ) {
        const int v = it / (CPB / 4), l = 4 * (it - v * (CPB / 4));
        buf[k] = *reinterpret_cast<const float4*>((v < NV ? P + (int64_t)v * N : Dg) + cell0 + l);
      }
    }
#pragma unroll
    for (int k = 0; k < PERQ; ++k) {
      const int it = tid + k * NT;
      if (it < NQ) {
        const int v = it / (CPB / 4), l = 4 * (it - v * (CPB / 4));
        int ii[3];
        split<ND, BS>(l, ii);
        float* t = sP + v * NS + pslot<ND, BS>(ii);
        t[0] = buf[k].x; t[1] = buf[k].y; t[2] = buf[k].z; t[3] = buf[k].w;
      }
    }
  } else {
    for (int l = tid; l < CPB; l += NT) {
      int ii[3];
      split<ND, BS>(l, ii);
      int s = pslot<ND, BS>(ii);
#pragma unroll
      for (int v = 0; v < NV; ++v) sP[v * NS + s] = P[(int64_t)v * N + cell0 + l];
      sD[s] = Dg[cell0 + l];
    }
  }
  __syncthreads();
  // ---- halos: same-level faces into the padded slabs, coarser / finer faces into their own areas
#pragma unroll 1
  for (int f = 0; f < 2 * ND; ++f) {
    const FaceInfo& F = fi[f];
    if (F.kind == 0) continue;
    if (MODE == 2 && F.kind != 1) continue;  // coarser / finer halos are only read by the general faces
    // ... and, with 4 or more cells per block edge, the general faces (block face + first internal face: own cells
    // 0..2 along d and the irregular halo) never read a same-level halo
    if (MODE == 1 && F.kind == 1 && BS >= 4) continue;
    int d = f >> 1, side = f & 1, n1n2 = F.n1 * F.n2;
    for (int k = tid; k < 2 * n1n2; k += NT) {
      int layer = k / n1n2, r = k - layer * n1n2;
      int j1 = r % F.n1, j2 = r / F.n1;
      int64_t c = halo_cell<ND, BS>(F, d, side, j1, j2, layer, CPB);
      int s = F.kind == 1 ? pslot_c<ND, BS>(d, side ? BS + layer : -1 - layer, j1, j2) : F.base + k;
#pragma unroll
      for (int v = 0; v < NV; ++v) sP[v * NS + s] = P[(int64_t)v * N + c];
      sD[s] = Dg[c];
    }
  }
  __syncthreads();
  float res[RC::CPT][NV], cf[RC::CPT];
#pragma unroll
  for (int q = 0; q < RC::CPT; ++q) {
    cf[q] = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) res[q][v] = 0.0f;
  }
  const float gr = fl.gamma * fl.R;
#pragma unroll 1
  for (int d = 0; d < ND; ++d) {
    const float hd = bh[b * ND + d];
    const bool p2 = P2;
    const float inv_hd = 1.0f / hd;
    const double inv_hd_d = 1.0 / (double)hd;
    const int m0 = d == 0 ? BS + 1 : BS, m1 = d == 1 ? BS + 1 : BS;
    const int ss = d == 0 ? 1 : (d == 1 ? PAD : PAD * PAD);
    const FaceInfo& FL = fi[2 * d];
    const FaceInfo& FH = fi[2 * d + 1];
    const bool irr_lo = FL.kind != 1, irr_hi = FH.kind != 1;
    const int nxl = (FINER && FL.kind == 3) ? FL.n1 * FL.n2 : 0, nxh = (FINER && FH.kind == 3) ? FH.n1 * FH.n2 : 0;
    // ---- (1) fluxes.  Items [0, NFD): the uniform index space of k_reg_flux; faces whose stencil touches an
    //      irregular block face are skipped there and re-enumerated, warp-uniformly, as items >= NFD:
    //      [low block face, first internal face] if the low face is irregular, [last internal face, high block face]
    //      if the high one is, then the fine faces of finer neighbours.
    const int nlo = irr_lo ? 2 * FACE : 0, nhi = irr_hi ? 2 * FACE : 0;
    const int64_t gbase = ((int64_t)blockIdx.x * ND + d) * NSL;
    for (int it = (MODE == 1 ? NFD + tid : tid); it < (MODE == 2 ? NFD : NFD + nlo + nhi + nxl + nxh); it += NT) {
      float po[NV], pn[NV], g0[NV], g1[NV];
      float ho = hd, hn = hd, Do, Dn;
      int fslot = it, sid = 0;
      if (it < NFD + nlo + nhi) {
        int cc[3];
        if (it < NFD) {
          face_decode<ND, BS>(it, d, cc);
        } else {
          int k = it - NFD;
          int cdv;
          if (k < nlo) { cdv = k / FACE; sid = k; }                           // 0: low block face, 1: first internal face
          else { k -= nlo; cdv = BS - 1 + k / FACE; sid = 2 * FACE + k; }     // BS-1: last internal face, BS: high block face
          int pq = k % FACE;
          cc[0] = cc[1] = cc[2] = 0;
          cc[d] = cdv;
          cc[T1(d)] = pq % BS;
          if (ND == 3) cc[T2(d)] = pq / BS;
          fslot = cc[0] + m0 * (cc[1] + (ND == 3 ? m1 * cc[2] : 0));
        }
        const int cd = cc[d];
        const int sn = pslot<ND, BS>(cc), so = sn - ss;
        const bool irregular = (irr_lo && cd <= 1) || (irr_hi && cd >= BS - 1);
        if (it < NFD) {
          if (irregular) continue;
          // uniform stencil: identical to k_reg_flux
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            float um = sP[v * NS + so - ss], uo = sP[v * NS + so], un = sP[v * NS + sn], up = sP[v * NS + sn + ss];
            float fm = face_interp_f(um, uo, hd, hd, p2), fc = face_interp_f(uo, un, hd, hd, p2), fp = face_interp_f(un, up, hd, hd, p2);
            g0[v] = p2 ? (fc - fm) * inv_hd : (fc - fm) / hd;
            g1[v] = p2 ? (fp - fc) * inv_hd : (fp - fc) / hd;
            po[v] = uo;
            pn[v] = un;
          }
          Do = sD[so]; Dn = sD[sn];
        } else {
          // general item: describe owner and neighbour as (slot, spacing, low list, high list), one gradient routine
          int co, cn;
          float hco = hd, hcn = hd;
          NL lo_o, hi_o, lo_n, hi_n;
          if (cd >= 1 && cd <= BS - 1) {        // internal face next to an irregular block face: two own cells
            int io[3] = {cc[0], cc[1], cc[2]};
            io[d] -= 1;
            co = so; cn = sn;
            lo_o = nl_own<ND, BS, FINER>(FL, io, co, d, 0, ss, hd); hi_o = nl_own<ND, BS, FINER>(FH, io, co, d, 1, ss, hd);
            lo_n = nl_own<ND, BS, FINER>(FL, cc, cn, d, 0, ss, hd); hi_n = nl_own<ND, BS, FINER>(FH, cc, cn, d, 1, ss, hd);
          } else {                              // block face: cd == 0 (low) or cd == BS (high)
            const int side = cd == 0 ? 0 : 1;
            const FaceInfo& F = side ? FH : FL;
            if (FINER && F.kind == 3) continue;  // its fine faces are the extra items below
            int io[3] = {cc[0], cc[1], cc[2]};
            io[d] = side ? BS - 1 : 0;
            const int own = pslot<ND, BS>(io);
            NL lo_w = nl_own<ND, BS, FINER>(FL, io, own, d, 0, ss, hd), hi_w = nl_own<ND, BS, FINER>(FH, io, own, d, 1, ss, hd);
            int ch = own;
            float hh = hd;
            NL lo_h = lo_w, hi_h = hi_w;         // box face: both sides are the own cell
            if (F.kind == 2) {                   // coarser neighbour
              int a1 = cc[T1(d)], a2 = ND == 3 ? cc[T2(d)] : 0;
              int hr = (a2 >> 1) * F.n1 + (a1 >> 1);
              ch = F.base + hr;
              hh = F.hn;
              NL nearl = nl_halo_near<ND, BS, FINER>(F, side, hr, d, hd);
              NL farl; farl.cnt = 1; farl.s[0] = ch + F.n1 * F.n2; farl.s[1] = farl.s[2] = farl.s[3] = 0; farl.h = hh;
              lo_h = side ? nearl : farl;
              hi_h = side ? farl : nearl;
            }
            if (side) { co = own; hco = hd; lo_o = lo_w; hi_o = hi_w; cn = ch; hcn = hh; lo_n = lo_h; hi_n = hi_h; }
            else      { co = ch; hco = hh; lo_o = lo_h; hi_o = hi_h; cn = own; hcn = hd; lo_n = lo_w; hi_n = hi_w; }
          }
          grad_lists<NV, NS>(sP, co, hco, lo_o, hi_o, p2, g0);
          grad_lists<NV, NS>(sP, cn, hcn, lo_n, hi_n, p2, g1);
#pragma unroll
          for (int v = 0; v < NV; ++v) { po[v] = sP[v * NS + co]; pn[v] = sP[v * NS + cn]; }
          Do = sD[co]; Dn = sD[cn]; ho = hco; hn = hcn;
        }
      } else {
        // fine face k of a finer neighbour: halo fine cell k <-> own coarse cell (j1/2, j2/2)
        const int kx = it - NFD - nlo - nhi;
        const int side = kx >= nxl ? 1 : 0;
        const FaceInfo& F = side ? FH : FL;
        const int k = kx - (side ? nxl : 0);
        int j1 = k % F.n1, j2 = k / F.n1;
        int io[3] = {0, 0, 0};
        io[d] = side ? BS - 1 : 0;
        io[T1(d)] = j1 >> 1;
        if (ND == 3) io[T2(d)] = j2 >> 1;
        const int own = pslot<ND, BS>(io), hs = F.base + k;
        NL lo_w = nl_own<ND, BS, FINER>(FL, io, own, d, 0, ss, hd), hi_w = nl_own<ND, BS, FINER>(FH, io, own, d, 1, ss, hd);
        NL nearl = nl_halo_near<ND, BS, FINER>(F, side, k, d, hd);
        NL farl; farl.cnt = 1; farl.s[0] = hs + F.n1 * F.n2; farl.s[1] = farl.s[2] = farl.s[3] = 0; farl.h = F.hn;
        int co, cn;
        float hco, hcn;
        if (side) {
          co = own; hco = hd; cn = hs; hcn = F.hn;
          grad_lists<NV, NS>(sP, co, hco, lo_w, hi_w, p2, g0);
          grad_lists<NV, NS>(sP, cn, hcn, nearl, farl, p2, g1);
        } else {
          co = hs; hco = F.hn; cn = own; hcn = hd;
          grad_lists<NV, NS>(sP, co, hco, farl, nearl, p2, g0);
          grad_lists<NV, NS>(sP, cn, hcn, lo_w, hi_w, p2, g1);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) { po[v] = sP[v * NS + co]; pn[v] = sP[v * NS + cn]; }
        Do = sD[co]; Dn = sD[cn]; ho = hco; hn = hcn;
        fslot = NFD + side * (C::NX / 2) + k;
        sid = 4 * FACE + side * (C::NX / 2) + k;
      }
      float pl[NV], pr[NV];
      double F_[NV];
      const bool fast = p2 && ho == hn;
      muscl_face<NV>(po, pn, g0, g1, ho, hn, Do, Dn, true, false, pl, pr, fast);
      if (FLUX == 0) {
        hll_flux<ND>(fl, pl, pr, d, F_);
      } else {
        float Ff[NV];
        rusanov_flux<ND>(fl, pl, pr, face_interp(Do, Dn, ho, hn), d, Ff);
#pragma unroll
        for (int v = 0; v < NV; ++v) F_[v] = (double)Ff[v];
      }
      float ao = sqrtf(gr * clampT(po[1])), an = sqrtf(gr * clampT(pn[1]));
      float ct = fabsf(face_interp_f(pick<ND>(po + 2, d), pick<ND>(pn + 2, d), ho, hn, fast)) + face_interp_f(ao, an, ho, hn, fast);
      if (MODE == 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) GF[(gbase + sid) * NV + v] = F_[v];
        GC[gbase + sid] = ct;
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) sF[v * NFS + fslot] = F_[v];
        sC[fslot] = ct;
      }
    }
    if (MODE == 1) continue;  // general-only pass: no divergence here
    if (MODE == 2) {
      // fetch the fluxes of the general faces computed by the MODE 1 pass (same enumeration, same slots)
      for (int k = tid; k < nlo + nhi + nxl + nxh; k += NT) {
        int fslot, sid;
        if (k < nlo + nhi) {
          int kk = k, cdv;
          if (kk < nlo) { cdv = kk / FACE; sid = kk; }
          else { kk -= nlo; cdv = BS - 1 + kk / FACE; sid = 2 * FACE + kk; }
          if (FINER && ((cdv == 0 && FL.kind == 3) || (cdv == BS && FH.kind == 3))) continue;  // replaced by fine faces
          int pq = kk % FACE, cc[3] = {0, 0, 0};
          cc[d] = cdv;
          cc[T1(d)] = pq % BS;
          if (ND == 3) cc[T2(d)] = pq / BS;
          fslot = cc[0] + m0 * (cc[1] + (ND == 3 ? m1 * cc[2] : 0));
        } else {
          int kx = k - nlo - nhi;
          int side = kx >= nxl ? 1 : 0;
          int kf = kx - (side ? nxl : 0);
          fslot = NFD + side * (C::NX / 2) + kf;
          sid = 4 * FACE + side * (C::NX / 2) + kf;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) sF[v * NFS + fslot] = GF[(gbase + sid) * NV + v];
        sC[fslot] = GC[gbase + sid];
      }
    }
    __syncthreads();
    // ---- (2) divergence
#pragma unroll
    for (int q = 0; q < RC::CPT; ++q) {
      int l = tid + q * NT;
      if (l < CPB) {
        int ii[3];
        split<ND, BS>(l, ii);
        int lo = ii[0] + m0 * (ii[1] + (ND == 3 ? m1 * ii[2] : 0));
        int hi = lo + (d == 0 ? 1 : (d == 1 ? m0 : m0 * m1));
        double mh[NV], ml[NV];
        float ch, cl;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          double* m = side ? mh : ml;
          float& cm = side ? ch : cl;
          const FaceInfo& F = side ? FH : FL;
          bool inner = side ? ii[d] < BS - 1 : ii[d] > 0;
          if (inner || !FINER || F.kind != 3) {
            int fs = side ? hi : lo;
#pragma unroll
            for (int v = 0; v < NV; ++v) m[v] = sF[v * NFS + fs];
            cm = sC[fs];
          } else {
            constexpr int CNT = ND == 3 ? 4 : 2;
            const float w = 1.0f / (float)CNT;
            int a1 = ii[T1(d)], a2 = ND == 3 ? ii[T2(d)] : 0;
#pragma unroll
            for (int qq = 0; qq < CNT; ++qq) {
              int fs = NFD + side * (C::NX / 2) + (ND == 3 ? (2 * a2 + (qq >> 1)) * F.n1 : 0) + 2 * a1 + (qq & 1);
              if (FLUX == 0) {
#pragma unroll
                for (int v = 0; v < NV; ++v) m[v] = qq == 0 ? sF[v * NFS + fs] * (double)w : m[v] + sF[v * NFS + fs] * (double)w;
              } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  float t = (float)sF[v * NFS + fs] * w;
                  m[v] = qq == 0 ? (double)t : (double)((float)m[v] + t);
                }
              }
              cm = qq == 0 ? sC[fs] * w : cm + sC[fs] * w;
            }
          }
        }
        if (FLUX == 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) res[q][v] = (float)((double)res[q][v] - (mh[v] - ml[v]) * inv_hd_d);
        } else {
#pragma unroll
          for (int v = 0; v < NV; ++v) res[q][v] = res[q][v] - ((float)mh[v] - (float)ml[v]) / hd;
        }
        cf[q] = cf[q] + (p2 ? (ch + cl) * inv_hd : (ch + cl) / hd);
      }
    }
    __syncthreads();
  }
  if (MODE == 1) return;
#pragma unroll
  for (int q = 0; q < RC::CPT; ++q) {
    int l = tid + q * NT;
    if (l < CPB) {
#pragma unroll
      for (int v = 0; v < NV; ++v) R[(int64_t)v * N + cell0 + l] = res[q][v];
      cfl[cell0 + l] = cf[q];
    }
  }
}

template <int ND, int BS, bool FINER, bool P2, int FLUX, int MODE>
int launch_hyb_mode(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, ibx_fluid f, const float* P, const float* S,
                    float* R, float* cfl, double* GF, float* GC, cudaStream_t st = nullptr) {
  using C = HybCfg<ND, BS, FINER>;
  if (!st) st = c->stream;
  constexpr size_t SM = MODE == 1 ? C::SMEM_GENERAL : C::SMEM;
  int rc;
  if ((rc = ensure_dyn_smem(c, (const void*)k_hyb_flux<ND, BS, FINER, P2, FLUX, MODE>, SM))) return rc;
  k_hyb_flux<ND, BS, FINER, P2, FLUX, MODE><<<n, C::NT, SM, st>>>(blocks, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, R, cfl, GF, GC);
  LAUNCH_CHECK();
  return IBX_OK;
}

template <int ND, int BS, bool FINER, bool P2, int FLUX>
int launch_hyb_flux(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, ibx_fluid f, const float* P, const float* S,
                    float* R, float* cfl) {
  using C = HybCfg<ND, BS, FINER>;
  // blocks with a finer neighbour: few, heavy (1 CTA/SM either way) -- measured faster in one pass
  if (FINER)
    return launch_hyb_mode<ND, BS, FINER, P2, FLUX, 0>(c, D, blocks, n, f, P, S, R, cfl, nullptr, nullptr);
  // two passes through a global scratch holding the fluxes of the general faces
  constexpr int64_t NSL = 4 * C::FACE + C::NX;
  int64_t slots = (int64_t)n * ND * NSL;
  int64_t need = slots * (C::NV * 2 + 1);  // floats: NV doubles + 1 float per slot
  if (need > c->scratch2_cap) {
    if (c->d_scratch2) cudaFree(c->d_scratch2);
    c->d_scratch2 = nullptr;
    c->scratch2_cap = 0;
    CU(cudaMalloc((void**)&c->d_scratch2, (size_t)need * sizeof(float)));
    c->scratch2_cap = need;
  }
  double* GF = (double*)c->d_scratch2;
  float* GC = (float*)(GF + slots * C::NV);
  int rc;
  if ((rc = launch_hyb_mode<ND, BS, FINER, P2, FLUX, 1>(c, D, blocks, n, f, P, S, R, cfl, GF, GC))) return rc;
  return launch_hyb_mode<ND, BS, FINER, P2, FLUX, 2>(c, D, blocks, n, f, P, S, R, cfl, GF, GC);
}

template <int ND, int BS, bool FINER, bool P2>
int launch_hyb(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, ibx_fluid f, int flux_kind, const float* P,
               const float* S, float* R, float* cfl) {
  if (n == 0) return IBX_OK;
  if (flux_kind == 0) return launch_hyb_flux<ND, BS, FINER, P2, 0>(c, D, blocks, n, f, P, S, R, cfl);
  return launch_hyb_flux<ND, BS, FINER, P2, 1>(c, D, blocks, n, f, P, S, R, cfl);
}

// lean sensor kernel for regular blocks: padded (BS+2)^ND tile of p, uniform loop (same bits as k_tile_sensor)
template <int ND, int BS, bool P2>
__global__ void __launch_bounds__(RegCfg<ND, BS>::NT)
k_reg_sensor(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh,
             const float* __restrict__ p, float* __restrict__ D) {
  using C = RegCfg<ND, BS>;
  constexpr int PAD = BS + 2, CPB = C::CPB, NT = C::NT, FACE = C::FACE;
  constexpr int TS = ND == 3 ? PAD * PAD * PAD : PAD * PAD;
  __shared__ float sp[TS];
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  const int64_t cell0 = b * CPB;
  for (int l = tid; l < CPB; l += NT) {
    int ii[3];
    split<ND, BS>(l, ii);
    sp[(ii[0] + 1) + PAD * ((ii[1] + 1) + (ND == 3 ? PAD * (ii[2] + 1) : 0))] = p[cell0 + l];
  }
  for (int k = tid; k < 2 * ND * FACE; k += NT) {
    int f = k / FACE, q = k - f * FACE;
    int d = f >> 1, side = f & 1;
    int j1 = q % BS, j2 = q / BS;
    int64_t nb = faces[b * (2 * ND) + f].nb[0];
    int cc[3] = {0, 0, 0};
    cc[d] = side ? BS : -1;
    cc[T1(d)] = j1;
    if (ND == 3) cc[T2(d)] = j2;
    sp[(cc[0] + 1) + PAD * ((cc[1] + 1) + (ND == 3 ? PAD * (cc[2] + 1) : 0))] = p[nb * CPB + compose<ND, BS>(d, side ? 0 : BS - 1, j1, j2)];
  }
  __syncthreads();
  float h[ND], ih[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) { h[d] = bh[b * ND + d]; ih[d] = 1.0f / h[d]; }
  for (int l = tid; l < CPB; l += NT) {
    int ii[3];
    split<ND, BS>(l, ii);
    int s = (ii[0] + 1) + PAD * ((ii[1] + 1) + (ND == 3 ? PAD * (ii[2] + 1) : 0));
    float pc = sp[s], nu = 1e-7f;
    int ss = 1;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float fh = sp[s + ss] - pc, fl = pc - sp[s - ss];
      float gg = P2 ? (fh - fl) * ih[d] : (fh - fl) / h[d];
      float ug = P2 ? (fabsf(fh) + fabsf(fl)) * ih[d] : (fabsf(fh) + fabsf(fl)) / h[d];
      nu = fmaxf(nu, (1e-7f + fabsf(gg)) / (1e-7f + ug));
      ss *= PAD;
    }
    D[cell0 + l] = nu;
  }
}

__global__ void fill_ones(float* __restrict__ S, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) S[i] = 1.0f;
}

// Q -> P, elementwise (src/cfd.jl:137-151)
template <int ND>
__global__ void k_prim(ibx_fluid f, const float* __restrict__ Q, float* __restrict__ P, int64_t n, int64_t i0, int64_t i1) {
  for (int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    float q[ND + 2], p[ND + 2];
#pragma unroll
    for (int v = 0; v < ND + 2; ++v) q[v] = Q[(int64_t)v * n + i];
    s2p<ND>(f, q, p);
#pragma unroll
    for (int v = 0; v < ND + 2; ++v) P[(int64_t)v * n + i] = p[v];
  }
}

// JST sensor of the listed blocks through the general tile kernel (2-D meshes, block sizes 4 and 2, cross-check in 3-D)
template <int ND, int BS, bool FINER>
int launch_tile_sensor(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, const float* P, float* S) {
  using C = Cfg<ND, BS, FINER>;
  if (n == 0) return IBX_OK;
  int rc;
  if (C::SMEM_SENSOR > 48 * 1024) {
    if ((rc = ensure_dyn_smem(c, (const void*)k_tile_sensor<ND, BS, FINER, true>, C::SMEM_SENSOR))) return rc;
    if ((rc = ensure_dyn_smem(c, (const void*)k_tile_sensor<ND, BS, FINER, false>, C::SMEM_SENSOR))) return rc;
  }
  if (D.all_pow2) k_tile_sensor<ND, BS, FINER, true><<<n, C::NT, C::SMEM_SENSOR, c->stream>>>(blocks, D.d_block_faces, D.d_block_h, P, S);
  else k_tile_sensor<ND, BS, FINER, false><<<n, C::NT, C::SMEM_SENSOR, c->stream>>>(blocks, D.d_block_faces, D.d_block_h, P, S);
  LAUNCH_CHECK();
  return IBX_OK;
}

// general-face pass of the irregular blocks (MODE 1) followed by the marching kernel on all owned blocks
// the owned blocks to process, by class: regular / irregular without / with finer neighbours
struct FluxLists {
  const int32_t *regular, *plain, *finer;
  int n_regular, n_plain, n_finer;
};

template <int ND, int BS>
int run_march(ibx_ctx* c, const ibx_domain& D, const FluxLists& FL, ibx_fluid f, int flux_kind, const float* P, const float* S, float* R,
              float* cfl) {
  using CP = HybCfg<ND, BS, false>;
  using CF = HybCfg<ND, BS, true>;
  constexpr int NV = CP::NV;
  int rc;
  const int64_t sl_p = (int64_t)FL.n_plain * ND * (4 * CP::FACE + CP::NX), sl_f = (int64_t)FL.n_finer * ND * (4 * CF::FACE + CF::NX);
  const int64_t need = (sl_p + sl_f) * (NV * 2 + 1);  // floats: NV doubles + 1 float per slot
  if (need > c->scratch2_cap) {
    if (c->d_scratch2) cudaFree(c->d_scratch2);
    c->d_scratch2 = nullptr;
    c->scratch2_cap = 0;
    CU(cudaMalloc((void**)&c->d_scratch2, (size_t)need * sizeof(float)));
    c->scratch2_cap = need;
  }
  double* GFp = (double*)c->d_scratch2;
  double* GFf = GFp + sl_p * NV;
  float* GCp = (float*)(GFf + sl_f * NV);
  float* GCf = GCp + sl_p;
  // The three block classes are independent of each other: the two irregular ones (general-face pass, then marching)
  // run on side streams beside the regular-block kernel, which fills the SMs their short launches leave idle.
  if (!c->aux_fork) {
    for (int k = 0; k < 2; ++k) {
      CU(cudaStreamCreateWithFlags(&c->aux_stream[k], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&c->aux_join[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming));
  }
  cudaStream_t sp = c->aux_stream[0], sf = c->aux_stream[1];
  CU(cudaEventRecord(c->aux_fork, c->stream));
  CU(cudaStreamWaitEvent(sp, c->aux_fork, 0));
  CU(cudaStreamWaitEvent(sf, c->aux_fork, 0));
  if (FL.n_plain) {
    if ((rc = general_faces(c, D, FL.plain, FL.n_plain, false, f, flux_kind, P, S, GFp, GCp, sp))) return rc;
    if ((rc = march_flux(c, D, FL.plain, FL.n_plain, 1, f, flux_kind, P, S, R, cfl, GFp, GCp, sp))) return rc;
  }
  if (FL.n_finer) {
    if ((rc = general_faces(c, D, FL.finer, FL.n_finer, true, f, flux_kind, P, S, GFf, GCf, sf))) return rc;
    if ((rc = march_flux(c, D, FL.finer, FL.n_finer, 2, f, flux_kind, P, S, R, cfl, GFf, GCf, sf))) return rc;
  }
  if ((rc = march_flux(c, D, FL.regular, FL.n_regular, 0, f, flux_kind, P, S, R, cfl, nullptr, nullptr, c->stream))) return rc;
  CU(cudaEventRecord(c->aux_join[0], sp));
  CU(cudaEventRecord(c->aux_join[1], sf));
  CU(cudaStreamWaitEvent(c->stream, c->aux_join[0], 0));
  CU(cudaStreamWaitEvent(c->stream, c->aux_join[1], 0));
  return IBX_OK;
}

// Q -> P on the listed 8^3 blocks: one CTA of 128 threads per block, four consecutive cells per thread
__global__ void __launch_bounds__(128) k_prim_blocks8(ibx_fluid f, const int32_t* __restrict__ blocks, const float* __restrict__ Q,
                                                       float* __restrict__ P, int64_t n) {
  const int64_t i = (int64_t)blocks[blockIdx.x] * 512 + 4 * threadIdx.x;
  float4 q[5], p[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) q[v] = *reinterpret_cast<const float4*>(Q + (int64_t)v * n + i);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float qi[5], pi[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) qi[v] = k == 0 ? q[v].x : (k == 1 ? q[v].y : (k == 2 ? q[v].z : q[v].w));
    s2p<3>(f, qi, pi);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      if (k == 0) p[v].x = pi[v];
      else if (k == 1) p[v].y = pi[v];
      else if (k == 2) p[v].z = pi[v];
      else p[v].w = pi[v];
    }
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) *reinterpret_cast<float4*>(P + (int64_t)v * n + i) = p[v];
}

// one phase of the overlapped sharded step (3-D, 8^3 blocks, marching kernels): primitives, sensors and fluxes of the
// phase's block lists, on the compute stream
int run_phase(ibx_ctx* c, const ibx_domain& D, const ibx_domain::PhaseLists& L, ibx_fluid f, int flux_kind, const float* Q, float* P,
              float* S, float* R, float* cfl) {
  using PL = ibx_domain::PhaseLists;
  int rc;
  if (L.n[PL::PRIM]) {
    k_prim_blocks8<<<L.n[PL::PRIM], 128, 0, c->stream>>>(f, L.d[PL::PRIM], Q, P, D.ncells);
    LAUNCH_CHECK();
  }
  if (c->opt_sensor == 0) {
    if (L.n[PL::PRIM]) {   // D == 1 everywhere: filled once per phase for the blocks whose primitives are new (superset of the sensor lists)
      fill_ones<<<grid_for(D.ncells, 256, c->sm_count, 16), 256, 0, c->stream>>>(S, D.ncells);
      LAUNCH_CHECK();
    }
  } else {
    if ((rc = sensor_regular(c, D, L.d[PL::S_REG], L.n[PL::S_REG], P, S))) return rc;
    if ((rc = sensor_direct(c, D, L.d[PL::S_PLAIN], L.n[PL::S_PLAIN], P, S))) return rc;
    if ((rc = sensor_direct(c, D, L.d[PL::S_FINER], L.n[PL::S_FINER], P, S))) return rc;
  }
  const FluxLists FL{L.d[PL::F_REG], L.d[PL::F_PLAIN], L.d[PL::F_FINER], L.n[PL::F_REG], L.n[PL::F_PLAIN], L.n[PL::F_FINER]};
  return run_march<3, 8>(c, D, FL, f, flux_kind, P, S, R, cfl);
}

template <int ND, int BS>
int run_tiles(ibx_ctx* c, const ibx_domain& D, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S, float* R, float* cfl,
              bool prim_done = false) {
  int64_t N = D.ncells;
  if (prim_done) {
    // P was filled by the caller (ibx_step_euler: block-list conversions around the ghost update)
  } else if (c->halo_pending && D.shard.active && D.shard.n_owned <= N) {
    // a halo exchange of Q is in flight (ibx_halo_begin without ibx_halo_end): convert the owned rows while it runs,
    // then wait for it and convert the halo rows
    const int64_t no = D.shard.n_owned;   // owned rows come first, then the halo rows
    k_prim<ND><<<grid_for(no, 256, c->sm_count, 16), 256, 0, c->stream>>>(f, Q, P, N, 0, no);
    LAUNCH_CHECK();
    CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    c->halo_pending = 0;
    if (N > no) {
      k_prim<ND><<<grid_for(N - no, 256, c->sm_count, 16), 256, 0, c->stream>>>(f, Q, P, N, no, N);
      LAUNCH_CHECK();
    }
  } else {
    if (c->halo_pending) {
      CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
      c->halo_pending = 0;
    }
    k_prim<ND><<<grid_for(N, 256, c->sm_count, 16), 256, 0, c->stream>>>(f, Q, P, N, 0, N);
    LAUNCH_CHECK();
  }
  int rc = IBX_OK;
  // sensor on every local block (owned + halo blocks of a shard), fluxes on the owned blocks only
  // option "path" = 1 (ibx_set_option): tile kernels for the sensor and the fluxes also where the marching kernels
  // apply -- an independent implementation kept as the fallback (2-D, block sizes 4 / 2, non-power-of-two spacings)
  // and as the cross-check of the marching path (same bits)
  const bool march = ND == 3 && BS == 8 && march_supported(D) && c->opt_path == 0;
  const bool direct = ND == 3 && BS == 8 && c->opt_path == 0;
  if (c->opt_sensor == 0) {
    // MUSCL(...; D = nothing) (src/ImmersedBoundary.jl:1141): the blend with D == 1 returns uL, uR unchanged
    fill_ones<<<grid_for(N, 256, c->sm_count, 16), 256, 0, c->stream>>>(S, N);
    LAUNCH_CHECK();
  } else if (direct) {
    if ((rc = sensor_regular(c, D, D.d_blk_all_regular, D.n_all_regular, P, S))) return rc;
  } else if (D.n_all_regular) {
    using RC = RegCfg<ND, BS>;
    if (D.all_pow2) k_reg_sensor<ND, BS, true><<<D.n_all_regular, RC::NT, 0, c->stream>>>(D.d_blk_all_regular, D.d_block_faces, D.d_block_h, P, S);
    else k_reg_sensor<ND, BS, false><<<D.n_all_regular, RC::NT, 0, c->stream>>>(D.d_blk_all_regular, D.d_block_faces, D.d_block_h, P, S);
    LAUNCH_CHECK();
  }
  // irregular blocks: straight from global memory (gen.cu; 0.10 ms instead of 0.25 ms on C4 -- for regular blocks the
  // lean tile kernel above stays ahead, 0.22 vs 0.38 ms: boundary cells make up 58 % of a block)
  if (c->opt_sensor == 0) {
  } else if (direct) {
    if ((rc = sensor_direct(c, D, D.d_blk_all_plain, D.n_all_plain, P, S))) return rc;
    if ((rc = sensor_direct(c, D, D.d_blk_all_finer, D.n_all_finer, P, S))) return rc;
  } else {
    if ((rc = launch_tile_sensor<ND, BS, false>(c, D, D.d_blk_all_plain, D.n_all_plain, P, S))) return rc;
    if ((rc = launch_tile_sensor<ND, BS, true>(c, D, D.d_blk_all_finer, D.n_all_finer, P, S))) return rc;
  }
  // 3-D 8^3 blocks with power-of-two spacings: pencil-marching kernel (march.cu) on every owned block; the general
  // faces of the irregular blocks are computed first (MODE 1) and handed over through a global scratch
  if constexpr (ND == 3 && BS == 8) {
    if (march) {
      const FluxLists FL{D.d_blk_own_regular, D.d_blk_own_plain, D.d_blk_own_finer, D.n_own_regular, D.n_own_plain, D.n_own_finer};
      return run_march<ND, BS>(c, D, FL, f, flux_kind, P, S, R, cfl);
    }
  }
  if (c->opt_arith != 0)
    return fail(IBX_ERR_UNSUPPORTED, "ibx_residual_euler: arithmetic = 1 (fast) exists for the marching kernels only (3-D, block size 8, "
                                     "power-of-two spacings, path = 0)");
  // fluxes: regular blocks (all neighbours same level) through the lean kernel, the rest through the general one
  if ((rc = D.all_pow2 ? launch_reg<ND, BS, true>(c, D, f, flux_kind, P, S, R, cfl) : launch_reg<ND, BS, false>(c, D, f, flux_kind, P, S, R, cfl))) return rc;
  if (D.all_pow2) {
    if ((rc = launch_hyb<ND, BS, false, true>(c, D, D.d_blk_own_plain, D.n_own_plain, f, flux_kind, P, S, R, cfl))) return rc;
    if ((rc = launch_hyb<ND, BS, true, true>(c, D, D.d_blk_own_finer, D.n_own_finer, f, flux_kind, P, S, R, cfl))) return rc;
  } else {
    if ((rc = launch_hyb<ND, BS, false, false>(c, D, D.d_blk_own_plain, D.n_own_plain, f, flux_kind, P, S, R, cfl))) return rc;
    if ((rc = launch_hyb<ND, BS, true, false>(c, D, D.d_blk_own_finer, D.n_own_finer, f, flux_kind, P, S, R, cfl))) return rc;
  }
  return IBX_OK;
}

}  // namespace

namespace ibx {

bool tile_supported(const ibx_domain& D) { return D.block_size == 8 || D.block_size == 4 || D.block_size == 2; }

// whole-domain step (3-D, 8^3 blocks): Q -> P on the ghost-free blocks (`early`) / on the blocks holding ghost cells;
// then sensors and fluxes of every block with the primitives taken as given
int prim_blocks(ibx_ctx* c, const ibx_domain& D, bool early, ibx_fluid f, const float* Q, float* P) {
  const int32_t* list = early ? D.d_blk_noghost : D.d_blk_ghost;
  const int n = early ? D.n_blk_noghost : D.n_blk_ghost;
  if (n == 0) return IBX_OK;
  k_prim_blocks8<<<n, 128, 0, c->stream>>>(f, list, Q, P, D.ncells);
  LAUNCH_CHECK();
  return IBX_OK;
}
int residual_euler_after_prim(ibx_ctx* c, const ibx_domain& D, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S, float* R,
                              float* cfl) {
  return run_tiles<3, 8>(c, D, f, flux_kind, Q, P, S, R, cfl, true);
}

// phase 0 / 1 of the overlapped sharded step (see ibx_domain::PhaseLists); P (N x nv) and S (N) are scratch
int residual_euler_phase(ibx_ctx* c, const ibx_domain& D, int phase, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S,
                         float* R, float* cfl) {
  return run_phase(c, D, D.phase[phase], f, flux_kind, Q, P, S, R, cfl);
}

// Euler residual through the tile kernels; P (N x nv) and S (N) are scratch
int residual_euler_tiles(ibx_ctx* c, const ibx_domain& D, ibx_fluid f, int flux_kind, const float* Q, float* P, float* S,
                         float* R, float* cfl) {
  if (D.nd == 3) {
    if (D.block_size == 8) return run_tiles<3, 8>(c, D, f, flux_kind, Q, P, S, R, cfl);
    if (D.block_size == 4) return run_tiles<3, 4>(c, D, f, flux_kind, Q, P, S, R, cfl);
    return run_tiles<3, 2>(c, D, f, flux_kind, Q, P, S, R, cfl);
  }
  if (D.block_size == 8) return run_tiles<2, 8>(c, D, f, flux_kind, Q, P, S, R, cfl);
  if (D.block_size == 4) return run_tiles<2, 4>(c, D, f, flux_kind, Q, P, S, R, cfl);
  return run_tiles<2, 2>(c, D, f, flux_kind, Q, P, S, R, cfl);
}

}  // namespace ibx
