// Helpers shared by the tile kernels (tile.cu) and the pencil-marching flux kernel (march.cu): per-face descriptors of
// an octree block and the index maps between block-local coordinates, halo layers and global cell ids.
#pragma once
#include "device.cuh"

namespace {
using ibx::BlockFace;

struct FaceInfo {
  int kind, n1, n2, base, nfaces;
  float hn;
  int nb[4];
  int sub1, sub2;
};

__device__ __forceinline__ int T1(int d) { return d == 0 ? 1 : 0; }
__device__ __forceinline__ int T2(int d) { return d == 2 ? 1 : 2; }

template <int ND, int BS>
__device__ __forceinline__ int compose(int d, int cn, int c1, int c2) {
  int idx[3] = {0, 0, 0};
  idx[d] = cn;
  idx[T1(d)] = c1;
  if (ND == 3) idx[T2(d)] = c2;
  return idx[0] + BS * (idx[1] + BS * idx[2]);
}

template <int ND, int BS>
__device__ __forceinline__ void split(int l, int (&ii)[3]) {
  ii[0] = l % BS;
  ii[1] = (l / BS) % BS;
  ii[2] = ND == 3 ? l / (BS * BS) : 0;
}

// base: first slot of this face's halo area; layers: 1 (sensor) or 2 (flux)
template <int ND, int BS>
__device__ __forceinline__ void fill_face_info(FaceInfo& fi, const BlockFace& bf, int base, float h) {
  fi.kind = bf.kind >= 1 && bf.kind <= 3 ? bf.kind : 0;
  fi.sub1 = bf.sub[0];
  fi.sub2 = bf.sub[1];
#pragma unroll
  for (int q = 0; q < 4; ++q) fi.nb[q] = bf.nb[q];
  int n = fi.kind == 1 ? BS : (fi.kind == 2 ? BS / 2 : (fi.kind == 3 ? 2 * BS : 0));
  fi.n1 = n;
  fi.n2 = ND == 3 ? n : (n ? 1 : 0);
  fi.hn = fi.kind == 2 ? h * 2.0f : (fi.kind == 3 ? h * 0.5f : h);
  fi.base = base;
  fi.nfaces = fi.kind == 3 ? fi.n1 * fi.n2 : (ND == 3 ? BS * BS : BS);
}

// global cell id of halo cell (j1, j2, layer) of face (d, side)
template <int ND, int BS>
__device__ __forceinline__ int64_t halo_cell(const FaceInfo& fi, int d, int side, int j1, int j2, int layer, int64_t cpb) {
  int jn = side ? layer : BS - 1 - layer;
  int64_t nb;
  int J1 = j1, J2 = j2;
  if (fi.kind == 1) {
    nb = fi.nb[0];
  } else if (fi.kind == 2) {
    nb = fi.nb[0];
    J1 = j1 + fi.sub1 * (BS / 2);
    J2 = ND == 3 ? j2 + fi.sub2 * (BS / 2) : 0;
  } else {
    int q1 = j1 / BS, q2 = ND == 3 ? j2 / BS : 0;
    nb = fi.nb[q1 + 2 * q2];
    J1 = j1 % BS;
    J2 = ND == 3 ? j2 % BS : 0;
  }
  return nb * cpb + compose<ND, BS>(d, jn, J1, J2);
}

// layer-0 halo slots (relative to fi.base) of the cells facing own cell (a1, a2); ascending cell id
template <int ND, int BS>
__device__ __forceinline__ int own_to_halo(const FaceInfo& fi, int a1, int a2, int (&slot)[4]) {
  if (fi.kind == 1) {
    slot[0] = a2 * fi.n1 + a1;
    return 1;
  }
  if (fi.kind == 2) {
    slot[0] = (a2 >> 1) * fi.n1 + (a1 >> 1);
    return 1;
  }
  constexpr int CNT = ND == 3 ? 4 : 2;
#pragma unroll
  for (int q = 0; q < CNT; ++q) slot[q] = (ND == 3 ? (2 * a2 + (q >> 1)) * fi.n1 : 0) + 2 * a1 + (q & 1);
  return CNT;
}

}  // namespace
