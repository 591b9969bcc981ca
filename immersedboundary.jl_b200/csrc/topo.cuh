// Block-structured topology helpers shared by the per-cell gather kernels (fused.cu, rans.cu): neighbours of a cell across
// each of its faces straight from the per-block BlockFace table, with the reference's face-list semantics
// (src/ImmersedBoundary.jl:605-698): a box face has owner == neighbour == the cell; a coarse cell's list towards finer
// cells holds 2^(nd-1) faces in ascending cell order, averaged with weight 1/len.
#pragma once
#include "device.cuh"
#include "physics.cuh"

namespace {
using namespace ibx;
using namespace ibxk;

struct Topo {
  const BlockFace* __restrict__ faces;
  const float* __restrict__ h;  // nblocks x nd cell widths
  int bs, log_cpb_unused;
  int64_t cpb;
  int64_t ncells;     // rows of every field array (owned + halo cells on a shard)
  int64_t n_compute;  // cells whose residual is wanted (the owned range, stored first)
};

template <int ND>
struct Nbr {
  int cnt;
  int64_t cell[ND == 3 ? 4 : 2];
  float h;  // neighbour spacing along the face normal
};

template <int ND>
__device__ __forceinline__ void decode(const Topo& T, int64_t cell, int64_t& b, int (&ii)[ND]) {
  if (T.bs == 8) {   // the usual block size: shifts instead of 64-bit divisions by a run-time value (uniform branch)
    constexpr int LOG_CPB = 3 * ND;
    b = cell >> LOG_CPB;
    int l = (int)(cell & ((1 << LOG_CPB) - 1));
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      ii[d] = l & 7;
      l >>= 3;
    }
    return;
  }
  b = cell / T.cpb;
  int l = (int)(cell - b * T.cpb);
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    ii[d] = l % T.bs;
    l /= T.bs;
  }
}

template <int ND>
__device__ __forceinline__ int64_t encode(const Topo& T, int64_t b, const int (&ii)[ND]) {
  int64_t l = 0;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) l = l * T.bs + ii[d];
  return b * T.cpb + l;
}

// neighbours of `cell` across its (d, side) face; side 0 = low (left list), 1 = high (right list)
template <int ND>
__device__ __forceinline__ Nbr<ND> neighbors(const Topo& T, int64_t cell, int64_t b, const int (&ii)[ND], int d, int side) {
  Nbr<ND> out;
  const int bs = T.bs;
  float hc = T.h[b * ND + d];
  if ((side == 0 && ii[d] > 0) || (side == 1 && ii[d] < bs - 1)) {
    int64_t stride = 1;
    for (int k = 0; k < d; ++k) stride *= bs;
    out.cnt = 1;
    out.cell[0] = side ? cell + stride : cell - stride;
    out.h = hc;
    return out;
  }
  const BlockFace bf = T.faces[b * (2 * ND) + 2 * d + side];
  int jj[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) jj[k] = ii[k];
  jj[d] = side ? 0 : bs - 1;
  // tangential dims in increasing order
  int t1 = d == 0 ? 1 : 0;
  int t2 = ND == 3 ? (d == 2 ? 1 : 2) : t1;
  switch (bf.kind) {
    case 1:
      out.cnt = 1;
      out.cell[0] = encode<ND>(T, bf.nb[0], jj);
      out.h = hc;
      break;
    case 2:
      jj[t1] = (ii[t1] + bf.sub[0] * bs) >> 1;
      if (ND == 3) jj[t2] = (ii[t2] + bf.sub[1] * bs) >> 1;
      out.cnt = 1;
      out.cell[0] = encode<ND>(T, bf.nb[0], jj);
      out.h = hc * 2.0f;
      break;
    case 3: {
      int half = bs >> 1;
      int s1 = ii[t1] >= half, s2 = ND == 3 ? (ii[t2] >= half) : 0;
      int64_t nb = bf.nb[s1 + 2 * s2];
      int b1 = 2 * (ii[t1] - s1 * half), b2 = ND == 3 ? 2 * (ii[t2] - s2 * half) : 0;
      out.cnt = ND == 3 ? 4 : 2;
      out.h = hc * 0.5f;
#pragma unroll
      for (int q = 0; q < (ND == 3 ? 4 : 2); ++q) {
        jj[t1] = b1 + (q & 1);
        if (ND == 3) jj[t2] = b2 + (q >> 1);
        out.cell[q] = encode<ND>(T, nb, jj);
      }
      break;
    }
    default:  // domain box: the face's owner and neighbour are both this cell
      out.cnt = 1;
      out.cell[0] = cell;
      out.h = hc;
      break;
  }
  return out;
}

// gradient along d of NV variables at an arbitrary cell (Green-Gauss over its two face lists)
template <int ND, int NV>
__device__ __forceinline__ void cell_grad(const Topo& T, const float* __restrict__ U, int64_t cell, int d, float* g) {
  int64_t b;
  int ii[ND];
  decode<ND>(T, cell, b, ii);
  float hc = T.h[b * ND + d];
  float uc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) uc[v] = U[(int64_t)v * T.ncells + cell];
  float m[2][NV];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    Nbr<ND> nb = neighbors<ND>(T, cell, b, ii, d, side);
    float w = 1.0f / (float)nb.cnt;
    for (int k = 0; k < nb.cnt; ++k) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float un = U[(int64_t)v * T.ncells + nb.cell[k]];
        // at_faces: owner is the low-side cell; the formula is symmetric in (value, spacing) pairs
        float fv = side ? face_interp(uc[v], un, hc, nb.h) : face_interp(un, uc[v], nb.h, hc);
        m[side][v] = k == 0 ? fv * w : m[side][v] + fv * w;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) g[v] = (m[1][v] - m[0][v]) / hc;
}

inline Topo make_topo(const ibx_domain& D) {
  Topo T;
  T.faces = D.d_block_faces;
  T.h = D.d_block_h;
  T.bs = D.block_size;
  T.log_cpb_unused = 0;
  T.cpb = 1;
  for (int d = 0; d < D.nd; ++d) T.cpb *= D.block_size;
  T.ncells = D.ncells;
  T.n_compute = D.shard.active ? D.shard.n_owned : D.ncells;
  return T;
}


}  // namespace
