// Pencil-marching flux kernel of the fused Euler residual (3-D, 8^3 blocks, power-of-two spacings) -- kernel template
// shared by march.cu (reference-exact arithmetic, compiled with -fmad=false) and march_fast.cu (option "arithmetic" = 1,
// compiled with FMA contraction).
//
// Same arithmetic, same operation order, same bits as k_reg_flux / k_hyb_flux (tile.cu) -- cell_gradient + MUSCL +
// inviscid_fluxes + green_gauss + the CFL term (src/ImmersedBoundary.jl:918-1157, src/cfd.jl:459-554) -- but the work
// of one dimension is organised as 64 pencils of 9 faces instead of 576 independent faces:
//   * a thread walks along the dimension and keeps the 4-cell stencil of the previous face in registers, so one face
//     costs ONE shared-memory load, one face interpolation and one gradient per variable (instead of 4 / 3 / 2);
//   * the flux of the low face of a cell is still in registers when its high face is done: the Green-Gauss difference
//     is taken in place, no flux buffer in shared memory;
//   * the running residual of a cell moves between dimensions through a small float buffer (the thread that owns a
//     cell changes with the marching direction) and is rounded once per dimension exactly like `R .-= green_gauss(..)`.
// Irregular blocks (box / coarser / finer contacts) run the same loop; the faces whose 4-cell stencil touches an
// irregular block face (the block face and the first internal face behind it) take their flux from the scratch written
// by the general-face pass (k_gen_faces, gen.cu) instead of computing it.
#pragma once
#include "device.cuh"
#include "physics.cuh"
#include "tile_common.cuh"

using namespace ibx;
using namespace ibxk;

namespace {

// Shared-memory layout of one field: the block's own cells at x + OY y + OZ z (pitches 9 / 72: every marching
// direction touches 32 different banks), then for each dimension d and side the two halo layers of the same-level
// neighbour as [layer][pencil] planes, pencil = t1 + 8 t2 in the plane normal to d, layers ordered along +d.  The
// sensor only needs the layer next to the block.  44.5 KB per CTA in all, so that five CTAs share an SM.
struct MarchCfg {
  static constexpr int SEG = 2;
  static constexpr int ND = 3, BS = 8, NV = 5;
  static constexpr int CPB = BS * BS * BS, FACE = BS * BS;
  static constexpr int OY = BS + 1, OZ = OY * BS;  // own-cell pitches (also those of the running-residual buffer)
  static constexpr int OWN = OZ * BS;              // 576 slots
  static constexpr int HB = OWN;                   // halo planes of the primitives: HB + d*4*FACE + side*2*FACE + layer*FACE + pencil
  static constexpr int FS = OWN + ND * 4 * FACE;   // 1344 slots per primitive field
  static constexpr int DS = OWN + ND * 2 * FACE;   // 960 slots for the sensor: HB + d*2*FACE + side*FACE + pencil
  static constexpr int RS = OWN;
  static constexpr int NT = FACE * SEG;            // SEG threads share a pencil, each marching BS / SEG cells
  static constexpr int CS = BS / SEG;
  static constexpr int NFLD = NV + 1;              // staged fields: the primitives and the sensor
  static constexpr size_t SMEM = sizeof(float) * ((size_t)NV * FS + DS + (size_t)NFLD * RS);
};

// HYB 0: regular blocks (all six neighbours same level).  1: irregular, no finer neighbour.  2: with finer neighbours.
// The stencil advance and MUSCL run on packed FP32 pairs (two variables per FADD2 / FMUL2, physics.cuh).  Two threads
// share a pencil; its middle face is evaluated once: the upper thread skips the flux of its first face, hands the flux
// of its second face to the lower thread through the pencil's (by then dead) low-halo slots, and the lower thread
// updates the cell between them; producer / consumer named barriers, no CTA-wide wait.
// FAST (option "arithmetic" = 1, compiled WITH FMA contraction in march_fast.cu): Float32 flux and Green-Gauss
// difference, approximate reciprocals -- not the reference's roundings, see DESIGN.md 4.1 for the measured error.
template <int FLUX, int HYB, bool FAST>
__global__ void __launch_bounds__(MarchCfg::NT, 5)
k_march_flux(const int32_t* __restrict__ blocks, const BlockFace* __restrict__ faces, const float* __restrict__ bh, int64_t N,
             ibx_fluid fl, const float* __restrict__ P, const float* __restrict__ Dg, float* __restrict__ R,
             float* __restrict__ cfl, const double* __restrict__ GF, const float* __restrict__ GC) {
  using C = MarchCfg;
  constexpr bool X2 = true, SHARE = true;
  constexpr int SEG = C::SEG;
  constexpr int ND = 3, BS = 8, NV = 5, CPB = C::CPB, FACE = C::FACE, NT = C::NT;
  constexpr int OY = C::OY, OZ = C::OZ, HB = C::HB, FS = C::FS, RS = C::RS, CS = C::CS, NFLD = C::NFLD;
  constexpr int NX = HYB == 2 ? 2 * FACE * 4 : 0;   // scratch slots of the fine faces (HybCfg::NX)
  constexpr int NSL = 4 * FACE + NX;                // scratch slots per (block, dimension)
  using FT = typename std::conditional<FLUX == 0 && !FAST, double, float>::type;
  using S = typename std::conditional<X2, P2, float>::type;   // one register of the marching state
  constexpr int NR = X2 ? (NV + 1) / 2 : NV;                   // registers per cell
  extern __shared__ float smem_f[];
  float* sP = smem_f;                 // [NV][FS] primitives, then [DS] sensor
  float* sD = sP + NV * FS;
  float* sR = sD + C::DS;             // [NV][RS] running residual, then [RS] running CFL sum
  float* sCf = sR + NV * RS;
  __shared__ int s_kind[2 * ND];
  const int64_t b = blocks[blockIdx.x];
  const int tid = threadIdx.x;
  const int64_t cell0 = b * CPB;
  if (HYB != 0 && tid < 2 * ND) s_kind[tid] = faces[b * (2 * ND) + tid].kind;   // read after the staging barrier
  // same-level neighbour across face fc, or -1 (straight from the block-face table: no barrier before the loads)
  auto neighbour = [&](int fc) {
    const BlockFace* bf = faces + b * (2 * ND) + fc;
    return HYB == 0 || bf->kind == 1 ? bf->nb[0] : -1;
  };
  // ---- stage: own cells and the two nearest layers of every same-level neighbour, vectorised along x.
  //      Items per field: 128 float4 (own) + 2 x 64 float2 (x faces) + 2 x 32 float4 (y faces) + 2 x 32 float4 (z faces).
  //      Which kind of item a thread handles in round `sub` is a compile-time fact (NT divides the range limits), so
  //      the loop unrolls into straight-line code: ALL loads of the thread are issued before the first store
  //      (one exposed memory latency per CTA instead of one per item).
  {
    constexpr int ITEMS = 384, PER = ITEMS / NT, NIT = NFLD * PER;
    int64_t soff[PER];
    int dP[PER], dD[PER];   // destination slots in a primitive field / in the sensor field (-1: not staged)
#pragma unroll
    for (int sub = 0; sub < PER; ++sub) {
      const int r = tid + sub * NT;
      if (sub * NT < 128) {
        const int l = 4 * r;
        soff[sub] = cell0 + l;
        dP[sub] = dD[sub] = (l & 7) + OY * ((l >> 3) & 7) + OZ * (l >> 6);
      } else if (sub * NT < 256) {
        // x faces: float2 = the two layers of (side, y, z); the sensor keeps the one next to the block
        const int k = r - 128, side = k >> 6, pn = k & 63;
        const int nb = neighbour(side);
        soff[sub] = nb >= 0 ? (int64_t)nb * CPB + (side ? 0 : BS - 2) + 8 * pn : cell0;
        dP[sub] = HB + side * 2 * FACE + pn;
        dD[sub] = HB + side * FACE + pn;
      } else {
        const int k = r - 256, isz = k >> 6, side = (k >> 5) & 1, q = k & 31;
        const int nb = neighbour(2 + 2 * isz + side);
        const int e = isz ? 4 * q : 4 * (q & 3), zy = q >> 2;   // y faces: 16 floats per z; z faces: 128 floats
        const int64_t o = isz ? (side ? 0 : 64 * (BS - 2)) + e : (side ? 0 : 8 * (BS - 2)) + 64 * zy + e;
        soff[sub] = nb >= 0 ? (int64_t)nb * CPB + o : cell0;
        const int layer = isz ? e >> 6 : e >> 3;
        const int pn = isz ? (e & 63) : (e & 7) + 8 * zy;
        dP[sub] = HB + (1 + isz) * 4 * FACE + side * 2 * FACE + layer * FACE + pn;
        dD[sub] = layer == (side ? 0 : 1) ? HB + (1 + isz) * 2 * FACE + side * FACE + pn : -1;
      }
    }
    float4 buf[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int v = k / PER, sub = k % PER;
      const float* __restrict__ src = (v < NV ? P + (int64_t)v * N : Dg) + soff[sub];
      if (sub * NT >= 128 && sub * NT < 256) {
        float2 q = *reinterpret_cast<const float2*>(src);
        buf[k] = make_float4(q.x, q.y, 0.0f, 0.0f);
      } else {
        buf[k] = *reinterpret_cast<const float4*>(src);
      }
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int v = k / PER, sub = k % PER;
      const bool xface = sub * NT >= 128 && sub * NT < 256;
      if (v < NV) {
        float* t = sP + v * FS + dP[sub];
        if (xface) { t[0] = buf[k].x; t[FACE] = buf[k].y; }
        else { t[0] = buf[k].x; t[1] = buf[k].y; t[2] = buf[k].z; t[3] = buf[k].w; }
      } else if (xface) {
        const int side = (tid + sub * NT - 128) >> 6;
        sD[dD[sub]] = side ? buf[k].x : buf[k].y;
      } else if (dD[sub] >= 0) {
        float* t = sD + dD[sub];
        t[0] = buf[k].x; t[1] = buf[k].y; t[2] = buf[k].z; t[3] = buf[k].w;
      }
    }
  }
  // running residual / CFL sum start from zero (`R .= 0`): 16-byte stores, visible after the barrier
  for (int k = tid; k < NFLD * RS / 4; k += NT) reinterpret_cast<float4*>(sR)[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  __syncthreads();
  const int seg = tid / FACE, pen = tid - seg * FACE;
  const int t1 = pen & 7, t2 = pen >> 3;
  const int f0 = seg * CS;
  const float gr = fl.gamma * fl.R;
#pragma unroll 1   // one copy of the body: it has to stay inside the instruction cache
  for (int d = 0; d < ND; ++d) {
    const float hd = bh[b * ND + d];
    const float inv_hd = 1.0f / hd;                 // exact: hd is a power of two on this path
    const double inv_hd_d = 1.0 / (double)hd;
    const int ss = d == 0 ? 1 : (d == 1 ? OY : OZ);   // own-cell (and running-residual) stride along d
    const int s1 = d == 0 ? OY : 1, s2 = d == 2 ? OY : OZ;
    const int own0 = t1 * s1 + t2 * s2;               // slot of the pencil's cell 0
    const int hP = HB + d * 4 * FACE + pen, hD = HB + d * 2 * FACE + pen;
    // slot of cell k of the pencil, k in [-2, BS + 1], in a primitive field / in the sensor field
    auto posP = [&](int k) { return k < 0 ? hP + (k + 2) * FACE : (k < BS ? own0 + k * ss : hP + (k - BS + 2) * FACE); };
    auto posD = [&](int k) { return k < 0 ? hD : (k < BS ? own0 + k * ss : hD + FACE); };
    const bool irr_lo = HYB != 0 && s_kind[2 * d] != 1, irr_hi = HYB != 0 && s_kind[2 * d + 1] != 1;
    const bool fine_lo = HYB == 2 && s_kind[2 * d] == 3, fine_hi = HYB == 2 && s_kind[2 * d + 1] == 3;
    const int64_t gbase = ((int64_t)blockIdx.x * ND + d) * NSL;
    // ---- stencil of the first face of this thread's segment: owner = cell f0 - 1, neighbour = cell f0
    const int sm = posP(f0 - 2), so = posP(f0 - 1), sn = posP(f0);
    const S half = vsplat<S>(0.5f);
    S uo[NR], un[NR], fc[NR], dfo[NR];   // dfo = fl(fc - fm): the owner's gradient along d times hd
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const S um = vload<S, NV>(sP, k, FS, sm);
      uo[k] = vload<S, NV>(sP, k, FS, so);
      un[k] = vload<S, NV>(sP, k, FS, sn);
      const S fm = vmul(vadd(um, uo[k]), half);
      fc[k] = vmul(vadd(uo[k], un[k]), half);
      dfo[k] = vsub(fc[k], fm);
    }
    float Do = sD[posD(f0 - 1)], Dn = sD[posD(f0)];
    float ao = sqrt_rn_inrange(gr * clampT(vget(uo, 1))), an = sqrt_rn_inrange(gr * clampT(vget(un, 1)));
    int rl = own0 + (f0 - 1) * ss;                    // running-residual slot of the owner cell
    FT Fa[NV], Fb[NV];
    float ca = 0.0f, cb = 0.0f;
    // one face: advance the stencil, flux (or scratch), and -- when UPDATE -- the Green-Gauss update of the owner cell
    // Green-Gauss update of the cell at (rl, l) from the fluxes of its low (Fl, cl) and high (Fh, ch) faces along d
    auto cell_update = [&](auto& Fl, auto& Fh, float cl, float ch) {
      const float cprev = sCf[rl];                     // zeroed by the staging phase: no `d == 0` branch in the face
      const float cnew = cprev + (ch + cl) * inv_hd;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float prev = sR[v * RS + rl];
        float r;
        if (FLUX == 0 && !FAST) r = (float)((double)prev - ((double)Fh[v] - (double)Fl[v]) * inv_hd_d);
        else r = prev - ((float)Fh[v] - (float)Fl[v]) * inv_hd;
        sR[v * RS + rl] = r;   // also after the last dimension: the epilogue below stores the block coalesced
      }
      sCf[rl] = cnew;
    };
    // mode 0: plain.  SHARE, upper thread: mode 1 = advance only (no flux), mode 2 = hand the flux over instead of updating
    auto face = [&](int f, auto update, const int mode, auto& qo, auto& qn, auto& qp, auto& fcc, auto& fpp, auto& dfc, auto& dfp,
                    auto& Fl, auto& Fh, float& cl, float& ch, float& D0, float& D1, float& D2, float& a0, float& a1) {   // (S[NR] x 7, FT[NV] x 2)
      constexpr bool UPDATE = decltype(update)::value;
      const bool upper = SHARE && seg == 1;
      const int s = posP(f + 1);
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        qp[k] = vload<S, NV>(sP, k, FS, s);
        fpp[k] = vmul(vadd(qn[k], qp[k]), half);
        dfp[k] = vsub(fpp[k], fcc[k]);
      }
      D2 = sD[posD(f + 1)];
      const bool take = HYB != 0 && ((irr_lo && f <= 1) || (irr_hi && f >= BS - 1));
      if (mode == 1 && upper) {
        // the lower thread of this pencil evaluates this face
      } else if (!take) {
        float pl[NV], pr[NV];
        muscl_face_p2v<S, NV, NR>(qo, qn, fcc, dfc, dfp, D0, D1, pl, pr);
        if (FLUX == 0 && FAST) {
          float Ff[NV];
          hll_flux_f32<ND>(fl, pl, pr, d, Ff);
#pragma unroll
          for (int v = 0; v < NV; ++v) Fh[v] = (FT)Ff[v];
        } else if (FLUX == 0) {
          double Fd[NV];
          hll_flux<ND, true>(fl, pl, pr, d, Fd);
#pragma unroll
          for (int v = 0; v < NV; ++v) Fh[v] = (FT)Fd[v];
        } else {
          float Ff[NV];
          rusanov_flux<ND, true>(fl, pl, pr, (D0 + D1) * 0.5f, d, Ff);
#pragma unroll
          for (int v = 0; v < NV; ++v) Fh[v] = (FT)Ff[v];
        }
        const float vo[ND] = {vget(qo, 2), vget(qo, 3), vget(qo, 4)}, vn[ND] = {vget(qn, 2), vget(qn, 3), vget(qn, 4)};
        ch = fabsf((pick<ND>(vo, d) + pick<ND>(vn, d)) * 0.5f) + (a0 + a1) * 0.5f;
      } else if (HYB == 2 && ((f == 0 && fine_lo) || (f == BS && fine_hi))) {
        // block face towards finer neighbours: mean of its four fine faces, products first, in list order
        const int side = f == 0 ? 0 : 1;
        const float w = 0.25f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int64_t sid = gbase + 4 * FACE + side * (NX / 2) + (2 * t2 + (qq >> 1)) * (2 * BS) + 2 * t1 + (qq & 1);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (FLUX == 0 && !FAST) {
              double g = GF[sid * NV + v] * (double)w;
              Fh[v] = (FT)(qq == 0 ? g : (double)Fh[v] + g);
            } else {
              float g = (float)GF[sid * NV + v] * w;
              Fh[v] = (FT)(qq == 0 ? g : (float)Fh[v] + g);
            }
          }
          float gc = GC[sid] * w;
          ch = qq == 0 ? gc : ch + gc;
        }
      } else {
        const int64_t sid = gbase + (f <= 1 ? f * FACE : (2 + f - (BS - 1)) * FACE) + pen;
#pragma unroll
        for (int v = 0; v < NV; ++v) Fh[v] = (FT)GF[sid * NV + v];
        ch = GC[sid];
      }
      if (UPDATE) {
        if (mode == 2 && upper) {
          // hand (Fh, ch) to the lower thread: the pencil's low-halo slots were last read by its initialisation
          asm volatile("bar.sync 2, %0;" ::"n"(NT));
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (FLUX == 0 && !FAST) {
              const unsigned long long b = (unsigned long long)__double_as_longlong((double)Fh[v]);
              sP[v * FS + hP] = __uint_as_float((unsigned)b);
              sP[v * FS + hP + FACE] = __uint_as_float((unsigned)(b >> 32));
            } else {
              sP[v * FS + hP] = (float)Fh[v];
            }
          }
          sD[hD] = ch;
          __threadfence_block();
          asm volatile("bar.arrive 1, %0;" ::"n"(NT));
        } else {
          cell_update(Fl, Fh, cl, ch);   // the owner cell (f - 1) now has both of its faces along d
        }
      }
      a0 = sqrt_rn_inrange(gr * clampT(vget(qp, 1)));   // becomes the neighbour's speed of sound two faces on
      rl += ss;
    };
    // register roles rotate with period 6 (3 for the cell values, 2 for everything else): the loop is unrolled by hand
    // over one period so that no state is ever copied
    S up[NR], fp[NR], dfn[NR];
    float Dp;
    std::true_type U;
    std::false_type NU;
    // f0: flux only
    if (SHARE && seg == 0) asm volatile("bar.arrive 2, %0;" ::"n"(NT));   // this pencil's low-halo slots are free now
    face(f0, NU, SHARE ? 1 : 0, uo, un, up, fc, fp, dfo, dfn, Fb, Fa, cb, ca, Do, Dn, Dp, ao, an);
    //          owner nb   new  fc  fp  dfc  dfp  Flo Fhi
    // after a step: owner <- nb, nb <- new; fc <- fp; dfc <- dfp; Flo <- Fhi; (D0, D1) <- (D1, D2); (a0, a1) <- (a1, a0')
    static_assert(CS == 4, "two threads per pencil");
    face(f0 + 1, U, SHARE ? 2 : 0, un, up, uo, fp, fc, dfn, dfo, Fa, Fb, ca, cb, Dn, Dp, Do, an, ao);
    face(f0 + 2, U, 0, up, uo, un, fc, fp, dfo, dfn, Fb, Fa, cb, ca, Dp, Do, Dn, ao, an);
    face(f0 + 3, U, 0, uo, un, up, fp, fc, dfn, dfo, Fa, Fb, ca, cb, Do, Dn, Dp, an, ao);
    face(f0 + 4, U, 0, un, up, uo, fc, fp, dfo, dfn, Fb, Fa, cb, ca, Dn, Dp, Do, ao, an);
    if (SHARE && seg == 0) {
      // the cell between the two threads: its low face is this thread's last (Fa, ca), its high face the upper thread's
      asm volatile("bar.sync 1, %0;" ::"n"(NT));
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (FLUX == 0 && !FAST) {
          const unsigned long long b = ((unsigned long long)__float_as_uint(sP[v * FS + hP + FACE]) << 32) | __float_as_uint(sP[v * FS + hP]);
          Fb[v] = (FT)__longlong_as_double((long long)b);
        } else {
          Fb[v] = (FT)sP[v * FS + hP];
        }
      }
      cell_update(Fa, Fb, ca, sD[hD]);
    }
    __syncthreads();
  }
  // ---- epilogue: residual and CFL sum of the block, four consecutive cells (one 16-byte store) per thread and field.
  //      Keeping the global stores out of the face loop leaves a face free of branches and of 64-bit address arithmetic.
  {
    const int lq = 4 * tid;
    const int slot = (lq & 7) + OY * ((lq >> 3) & 7) + OZ * (lq >> 6);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float* t = sR + v * RS + slot;
      *reinterpret_cast<float4*>(R + (int64_t)v * N + cell0 + lq) = make_float4(t[0], t[1], t[2], t[3]);
    }
    const float* t = sCf + slot;
    *reinterpret_cast<float4*>(cfl + cell0 + lq) = make_float4(t[0], t[1], t[2], t[3]);
  }
}

template <int FLUX, int HYB, bool FAST>
int launch_march(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, ibx_fluid f, const float* P, const float* S,
                 float* R, float* cfl, const double* GF, const float* GC, cudaStream_t st) {
  using C = MarchCfg;
  int rc;
  if ((rc = ensure_dyn_smem(c, (const void*)k_march_flux<FLUX, HYB, FAST>, C::SMEM))) return rc;
  k_march_flux<FLUX, HYB, FAST><<<n, C::NT, C::SMEM, st>>>(blocks, D.d_block_faces, D.d_block_h, D.ncells, f, P, S, R, cfl, GF, GC);
  LAUNCH_CHECK();
  return IBX_OK;
}

// Marching flux pass over the listed blocks.  hyb: 0 regular, 1 irregular without / 2 with finer neighbours; for
// hyb != 0 the fluxes of the general faces are read from (GF, GC), laid out by k_gen_faces (gen.cu).
template <bool FAST>
int march_flux_impl(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, int hyb, ibx_fluid f, int flux_kind, const float* P,
                    const float* S, float* R, float* cfl, const double* GF, const float* GC, cudaStream_t st) {
  if (n == 0) return IBX_OK;
#define GO(FLUX, HYB) return launch_march<FLUX, HYB, FAST>(c, D, blocks, n, f, P, S, R, cfl, GF, GC, st)
  if (flux_kind == 0) {
    if (hyb == 0) GO(0, 0);
    if (hyb == 1) GO(0, 1);
    GO(0, 2);
  }
  if (hyb == 0) GO(1, 0);
  if (hyb == 1) GO(1, 1);
  GO(1, 2);
#undef GO
}

}  // namespace
