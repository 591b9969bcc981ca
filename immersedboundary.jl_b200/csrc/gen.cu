// General faces of irregular blocks (3-D, 8^3 blocks, power-of-two spacings): the block faces towards the domain box,
// a coarser neighbour or finer neighbours, and the first internal face behind each of them -- the only faces of a
// block whose 4-cell stencil leaves the uniform lattice.  They are 2.5 % of the faces of the C4 mesh; the generic
// neighbour-list code that used to evaluate them (k_hyb_flux MODE 1, tile.cu) took 17 % of the step.
//
// Three CTAs of 64 threads per irregular block (one per dimension, one thread per pencil of an irregular face); no shared memory, no barrier: a thread
// reads the three own cells behind the face and the halo cells in front of it straight from global memory (each is
// read by at most four threads; L1 absorbs that) and evaluates, with the arithmetic and the operation order of the
// reference (weights 1/len, products first, sums in list order: src/accumulator.jl:95-106, at_faces :899-910,
// green_gauss :918-926, MUSCL :1113-1157, src/cfd.jl:459-554):
//   kind 0 (domain box):  owner == neighbour == the boundary cell (src/ImmersedBoundary.jl:648-667);
//   kind 2 (coarser):     one coarse halo cell per 2 x 2 pencils, spacing 2 h; its near face is the mean of 4 faces;
//   kind 3 (finer):       4 fine faces per pencil, spacing h / 2; the boundary cell's outer face is their mean.
// Fluxes go to the scratch (GF, GC) in the slot order k_march_flux (march.cu) reads: per (block, dimension)
// [f = 0 | f = 1 | f = 7 | f = 8] x 64 pencils, then the fine faces of the low / high block face (16 x 16 each).
#include "device.cuh"
#include "physics.cuh"
#include "tile_common.cuh"

using namespace ibx;
using namespace ibxk;

namespace {

constexpr int ND = 3, BS = 8, NV = 5, NF = NV + 1, CPB = 512, FACE = 64;

struct CellVals { float u[NF]; };   // the primitives and the sensor of one cell

__device__ __forceinline__ CellVals load_cell(const float* __restrict__ P, const float* __restrict__ Dg, int64_t N, int64_t c) {
  CellVals r;
#pragma unroll
  for (int v = 0; v < NV; ++v) r.u[v] = P[(int64_t)v * N + c];
  r.u[NV] = Dg[c];
  return r;
}

// MUSCL + flux + CFL term of one face and the store into the scratch (identical tail to k_hyb_flux, tile.cu)
template <int FLUX>
__device__ __forceinline__ void face_flux(ibx_fluid fl, int d, const CellVals& O, const CellVals& Nn, const float* go, const float* gn,
                                          float ho, float hn, double* __restrict__ GF, float* __restrict__ GC, int64_t slot) {
  float pl[NV], pr[NV];
  const bool fast = ho == hn;
  muscl_face<NV>(O.u, Nn.u, go, gn, ho, hn, O.u[NV], Nn.u[NV], true, false, pl, pr, fast);
  double F_[NV];
  // the in-range division / square-root sequences (physics.cuh): same bits as the guarded library forms for pressures,
  // R T and gamma R T, without their range-check branches
  if (FLUX == 0) {
    hll_flux<ND, true>(fl, pl, pr, d, F_);
  } else {
    float Ff[NV];
    rusanov_flux<ND, true>(fl, pl, pr, face_interp(O.u[NV], Nn.u[NV], ho, hn), d, Ff);
#pragma unroll
    for (int v = 0; v < NV; ++v) F_[v] = (double)Ff[v];
  }
  const float gr = fl.gamma * fl.R;
  const float ao = sqrt_rn_inrange(gr * clampT(O.u[1])), an = sqrt_rn_inrange(gr * clampT(Nn.u[1]));
  const float ct = fabsf(face_interp_f(pick<ND>(O.u + 2, d), pick<ND>(Nn.u + 2, d), ho, hn, fast)) + face_interp_f(ao, an, ho, hn, fast);
#pragma unroll
  for (int v = 0; v < NV; ++v) GF[slot * NV + v] = F_[v];
  GC[slot] = ct;
}

template <int FLUX, bool FINER>
__global__ void __launch_bounds__(FACE)
k_gen_faces(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh, int64_t N,
            ibx_fluid fl, const float* __restrict__ P, const float* __restrict__ Dg, double* __restrict__ GF, float* __restrict__ GC) {
  constexpr int NX = FINER ? 2 * FACE * 4 : 0, NSL = 4 * FACE + NX;
  // three CTAs of 64 threads per block, one per dimension (one thread per pencil), each walking the low and the high block
  // face: a CTA whose dimension has no irregular face exits at once and frees its registers for the others
  const int blk = blockIdx.x / 3, grp = blockIdx.x - 3 * blk;
  const int64_t b = blocks[blk];
  const int pen = threadIdx.x, t1 = pen & 7, t2 = pen >> 3;
  const int64_t cell0 = b * CPB;
#pragma unroll 1
  for (int f = 2 * grp; f < 2 * grp + 2; ++f) {
    const BlockFace bf = faces[b * (2 * ND) + f];
    if (bf.kind == 1) continue;
    const int d = f >> 1, side = f & 1;
    const float h = bh[b * ND + d], inv_h = 1.0f / h;
    FaceInfo F;
    fill_face_info<ND, BS>(F, bf, 0, h);
    const int64_t gbase = ((int64_t)blk * ND + d) * NSL;
    // the three own cells behind the face: c0 on the block face, c1, c2 inwards
    const int bnd = side ? BS - 1 : 0, in = side ? -1 : 1;
    const CellVals c0 = load_cell(P, Dg, N, cell0 + compose<ND, BS>(d, bnd, t1, t2));
    const CellVals c1 = load_cell(P, Dg, N, cell0 + compose<ND, BS>(d, bnd + in, t1, t2));
    const CellVals c2 = load_cell(P, Dg, N, cell0 + compose<ND, BS>(d, bnd + 2 * in, t1, t2));
    // ---- gradient along d of c1 (uniform lattice) and of c0 (outer face: by kind)
    float g1[NV], g0[NV], mout[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float a = (c1.u[v] + c0.u[v]) * 0.5f, bq = (c1.u[v] + c2.u[v]) * 0.5f;   // towards c0 / towards c2
      g1[v] = side ? (a - bq) * inv_h : (bq - a) * inv_h;
    }
    CellVals hc;                       // the halo cell in front of the pencil (kinds 0 and 2)
    float hh = h;
    CellVals fine[4];                  // kind 3: the four fine halo cells in front of c0
    if (F.kind == 0) {
      hc = c0;                         // box face: owner == neighbour == the boundary cell
#pragma unroll
      for (int v = 0; v < NV; ++v) mout[v] = (c0.u[v] + c0.u[v]) * 0.5f;
    } else if (!FINER || F.kind == 2) {
      hh = F.hn;
      hc = load_cell(P, Dg, N, halo_cell<ND, BS>(F, d, side, t1 >> 1, t2 >> 1, 0, CPB));
#pragma unroll
      for (int v = 0; v < NV; ++v) mout[v] = face_interp(c0.u[v], hc.u[v], h, hh);
    } else {
      hh = F.hn;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        fine[q] = load_cell(P, Dg, N, halo_cell<ND, BS>(F, d, side, 2 * t1 + (q & 1), 2 * t2 + (q >> 1), 0, CPB));
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float fv = face_interp(c0.u[v], fine[q].u[v], h, hh) * 0.25f;
          mout[v] = q == 0 ? fv : mout[v] + fv;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float min_ = (c0.u[v] + c1.u[v]) * 0.5f;
      g0[v] = side ? (mout[v] - min_) * inv_h : (min_ - mout[v]) * inv_h;
    }
    // ---- first internal face: c0 | c1
    {
      const int64_t slot = gbase + (side ? 2 * FACE : FACE) + pen;
      if (side) face_flux<FLUX>(fl, d, c1, c0, g1, g0, h, h, GF, GC, slot);
      else face_flux<FLUX>(fl, d, c0, c1, g0, g1, h, h, GF, GC, slot);
    }
    // ---- the block face
    if (F.kind == 0) {
      face_flux<FLUX>(fl, d, c0, c0, g0, g0, h, h, GF, GC, gbase + (side ? 3 * FACE : 0) + pen);
    } else if (!FINER || F.kind == 2) {
      // coarse halo cell: far face towards its second layer (same spacing), near face = mean of the 4 fine faces
      const int j1 = t1 >> 1, j2 = t2 >> 1;
      const CellVals far = load_cell(P, Dg, N, halo_cell<ND, BS>(F, d, side, j1, j2, 1, CPB));
      const float inv_hh = 1.0f / hh;
      float gh[NV], mnear[NV];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t c = cell0 + compose<ND, BS>(d, bnd, 2 * j1 + (q & 1), 2 * j2 + (q >> 1));
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float fv = face_interp(hc.u[v], P[(int64_t)v * N + c], hh, h) * 0.25f;
          mnear[v] = q == 0 ? fv : mnear[v] + fv;
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float mfar = (hc.u[v] + far.u[v]) * 0.5f;
        gh[v] = side ? (mfar - mnear[v]) * inv_hh : (mnear[v] - mfar) * inv_hh;
      }
      const int64_t slot = gbase + (side ? 3 * FACE : 0) + pen;
      if (side) face_flux<FLUX>(fl, d, c0, hc, g0, gh, h, hh, GF, GC, slot);
      else face_flux<FLUX>(fl, d, hc, c0, gh, g0, hh, h, GF, GC, slot);
    } else {
      const float inv_hh = 1.0f / hh;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int k1 = 2 * t1 + (q & 1), k2 = 2 * t2 + (q >> 1);
        const CellVals far = load_cell(P, Dg, N, halo_cell<ND, BS>(F, d, side, k1, k2, 1, CPB));
        const CellVals fq = q == 0 ? fine[0] : (q == 1 ? fine[1] : (q == 2 ? fine[2] : fine[3]));
        float gh[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float mfar = (fq.u[v] + far.u[v]) * 0.5f;
          const float mnear = face_interp(fq.u[v], c0.u[v], hh, h);
          gh[v] = side ? (mfar - mnear) * inv_hh : (mnear - mfar) * inv_hh;
        }
        const int64_t slot = gbase + 4 * FACE + side * (NX / 2) + k2 * (2 * BS) + k1;
        if (side) face_flux<FLUX>(fl, d, c0, fq, g0, gh, h, hh, GF, GC, slot);
        else face_flux<FLUX>(fl, d, fq, c0, gh, g0, hh, h, GF, GC, slot);
      }
    }
  }
}

// JST sensor D = JST_sensor(part, p) with dim = 0 (src/ImmersedBoundary.jl:1077-1097) for every cell of the listed
// blocks, straight from global memory: a cell reads its 6 face neighbours (1, or 4 quarter-weighted fine cells across a
// finer contact; none across the domain box) through L1 -- p is a single 2 KB run per block, so the shared-memory tile
// and its barrier bought nothing but latency.  Same differences, same order, same bits as k_reg_sensor / k_tile_sensor.
template <bool P2>
__global__ void __launch_bounds__(256)
k_sensor_direct(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh,
                const float* __restrict__ p, float* __restrict__ D) {
  __shared__ FaceInfo fi[2 * ND];
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid < 2 * ND) fill_face_info<ND, BS>(fi[tid], faces[b * (2 * ND) + tid], 0, bh[b * ND + (tid >> 1)]);
  __syncthreads();
  const int64_t cell0 = b * CPB;
  float h[ND], ih[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) { h[d] = bh[b * ND + d]; ih[d] = 1.0f / h[d]; }
#pragma unroll
  for (int rep = 0; rep < CPB / 256; ++rep) {
    const int l = tid + rep * 256;
    int ii[3];
    split<ND, BS>(l, ii);
    const float pc = p[cell0 + l];
    float nu = 1e-7f;
    int stride = 1;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float g[2], a[2];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        float accg = 0.0f, acca = 0.0f;
        const bool inner = side ? ii[d] < BS - 1 : ii[d] > 0;
        if (inner) {
          const float fd = side ? p[cell0 + l + stride] - pc : pc - p[cell0 + l - stride];   // p_neighbour - p_owner
          accg = fd;
          acca = fabsf(fd);
        } else {
          const FaceInfo& F = fi[2 * d + side];
          const int a1 = ii[T1(d)], a2 = ii[T2(d)];
          if (F.kind == 1 || F.kind == 2) {
            const int sh = F.kind == 2 ? 1 : 0;
            const float pn = p[halo_cell<ND, BS>(F, d, side, a1 >> sh, a2 >> sh, 0, CPB)];
            const float fd = side ? pn - pc : pc - pn;
            accg = fd;
            acca = fabsf(fd);
          } else if (F.kind == 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float pn = p[halo_cell<ND, BS>(F, d, side, 2 * a1 + (q & 1), 2 * a2 + (q >> 1), 0, CPB)];
              const float fd = side ? pn - pc : pc - pn;
              accg = q == 0 ? fd * 0.25f : accg + fd * 0.25f;
              acca = q == 0 ? fabsf(fd) * 0.25f : acca + fabsf(fd) * 0.25f;
            }
          }   // kind 0, box face: owner == neighbour, difference 0
        }
        g[side] = accg;
        a[side] = acca;
      }
      const float gg = P2 ? (g[1] - g[0]) * ih[d] : (g[1] - g[0]) / h[d], ugg = P2 ? (a[1] + a[0]) * ih[d] : (a[1] + a[0]) / h[d];
      nu = fmaxf(nu, div_rn_inrange(1e-7f + fabsf(gg), 1e-7f + ugg));   // both operands in [1e-7, ~1e9]: the in-range sequence (physics.cuh)
      stride *= BS;
    }
    D[cell0 + l] = nu;
  }
}

// JST sensor of REGULAR blocks (all six neighbours same level): the (BS + 2)^3 tile of p of k_reg_sensor (tile.cu), staged
// with three vector loads per thread that are all in flight before the first store (own cells as float4, the x-face
// layers as scalars, the y / z-face layers as float4) instead of five dependent load -> store rounds.  Same bits.
template <bool P2>
__global__ void __launch_bounds__(128)
k_reg_sensor8(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh,
              const float* __restrict__ p, float* __restrict__ D) {
  constexpr int PD = BS + 2, PD2 = PD * PD;
  __shared__ float sp[PD * PD * PD];
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  const int64_t cell0 = b * CPB;
  {
    const int l = 4 * tid;
    const float4 own = *reinterpret_cast<const float4*>(p + cell0 + l);
    const int side = tid >> 6, pn = tid & 63;
    const float xf = p[(int64_t)faces[b * 6 + side].nb[0] * CPB + (side ? 0 : BS - 1) + 8 * pn];
    float4 yz = make_float4(0.f, 0.f, 0.f, 0.f);
    const int isz = tid >> 5 & 1, s2 = (tid >> 4) & 1, q = tid & 15, xh = (q & 1) * 4, r = q >> 1;
    if (tid < 64) {
      const int64_t nb = faces[b * 6 + 2 + 2 * isz + s2].nb[0];
      yz = *reinterpret_cast<const float4*>(p + nb * CPB + (isz ? (s2 ? 0 : 64 * (BS - 1)) + 8 * r : (s2 ? 0 : 8 * (BS - 1)) + 64 * r) + xh);
    }
    float* t = sp + ((l & 7) + 1) + PD * (((l >> 3) & 7) + 1) + PD2 * ((l >> 6) + 1);
    t[0] = own.x; t[1] = own.y; t[2] = own.z; t[3] = own.w;
    sp[(side ? BS + 1 : 0) + PD * ((pn & 7) + 1) + PD2 * ((pn >> 3) + 1)] = xf;
    if (tid < 64) {
      float* u = isz ? sp + (xh + 1) + PD * (r + 1) + PD2 * (s2 ? BS + 1 : 0) : sp + (xh + 1) + PD * (s2 ? BS + 1 : 0) + PD2 * (r + 1);
      u[0] = yz.x; u[1] = yz.y; u[2] = yz.z; u[3] = yz.w;
    }
  }
  __syncthreads();
  float h[ND], ih[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) { h[d] = bh[b * ND + d]; ih[d] = 1.0f / h[d]; }
#pragma unroll
  for (int rep = 0; rep < CPB / 128; ++rep) {
    const int l = tid + rep * 128;
    const int s = ((l & 7) + 1) + PD * (((l >> 3) & 7) + 1) + PD2 * ((l >> 6) + 1);
    const float pc = sp[s];
    float nu = 1e-7f;
    int ss = 1;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const float fh = sp[s + ss] - pc, fl = pc - sp[s - ss];
      const float gg = P2 ? (fh - fl) * ih[d] : (fh - fl) / h[d];
      const float ug = P2 ? (fabsf(fh) + fabsf(fl)) * ih[d] : (fabsf(fh) + fabsf(fl)) / h[d];
      nu = fmaxf(nu, div_rn_inrange(1e-7f + fabsf(gg), 1e-7f + ug));   // both operands in [1e-7, ~1e9]: the in-range sequence (physics.cuh)
      ss *= PD;
    }
    D[cell0 + l] = nu;
  }
}

}  // namespace

namespace ibx {

// JST sensor of the listed REGULAR 8^3 blocks (3-D)
int sensor_regular(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, const float* p, float* S) {
  if (n == 0) return IBX_OK;
  if (D.all_pow2) k_reg_sensor8<true><<<n, 128, 0, c->stream>>>(blocks, D.d_block_faces, D.d_block_h, p, S);
  else k_reg_sensor8<false><<<n, 128, 0, c->stream>>>(blocks, D.d_block_faces, D.d_block_h, p, S);
  LAUNCH_CHECK();
  return IBX_OK;
}

// JST sensor of the listed 8^3 blocks (3-D)
int sensor_direct(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, const float* p, float* S) {
  if (n == 0) return IBX_OK;
  if (D.all_pow2) k_sensor_direct<true><<<n, 256, 0, c->stream>>>(blocks, D.d_block_faces, D.d_block_h, p, S);
  else k_sensor_direct<false><<<n, 256, 0, c->stream>>>(blocks, D.d_block_faces, D.d_block_h, p, S);
  LAUNCH_CHECK();
  return IBX_OK;
}

// General-face pass over the listed irregular blocks; finer: the list's blocks have finer neighbours (scratch layout
// with the fine-face slots).  st: the stream to launch on.
int general_faces(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, bool finer, ibx_fluid f, int flux_kind,
                  const float* P, const float* S, double* GF, float* GC, cudaStream_t st) {
  if (n == 0) return IBX_OK;
  if (finer) {
    if (flux_kind == 0) k_gen_faces<0, true><<<3 * n, FACE, 0, st>>>(blocks, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, GF, GC);
    else k_gen_faces<1, true><<<3 * n, FACE, 0, st>>>(blocks, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, GF, GC);
  } else {
    if (flux_kind == 0) k_gen_faces<0, false><<<3 * n, FACE, 0, st>>>(blocks, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, GF, GC);
    else k_gen_faces<1, false><<<3 * n, FACE, 0, st>>>(blocks, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, GF, GC);
  }
  LAUNCH_CHECK();
  return IBX_OK;
}

}  // namespace ibx
