// Internal declarations shared by the host builders and the CUDA translation units of libibx.so.
// Nothing here is part of the ABI (see include/ibx.h).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <memory>
#include <map>
#include <algorithm>
#include <stdexcept>
#include "../../include/ibx.h"

namespace ibx {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define IBX_TRY try {
#define IBX_CATCH                                                     \
  }                                                                   \
  catch (const std::exception& e) { return ibx::fail(IBX_ERR_ARG, e.what()); } \
  catch (...) { return ibx::fail(IBX_ERR_ARG, "unknown C++ exception"); }
#define IBX_REQUIRE(cond, msg) \
  do { if (!(cond)) return ibx::fail(IBX_ERR_ARG, std::string(__func__) + ": " + (msg)); } while (0)

// A scalar that remembers whether Julia would hold it as Float32 or Float64, so that the
// refinement comparisons of src/mesher.jl round exactly where the reference rounds.
struct Num {
  double v;
  bool f32;
};
inline Num nsub(Num a, Num b) {
  if (a.f32 && b.f32) return {(double)((float)a.v - (float)b.v), true};
  return {a.v - b.v, false};
}
inline Num nmul(Num a, Num b) {
  if (a.f32 && b.f32) return {(double)((float)a.v * (float)b.v), true};
  return {a.v * b.v, false};
}
inline Num nmax(Num a, Num b) { return {std::max(a.v, b.v), a.f32 && b.f32}; }
inline Num nmin(Num a, Num b) { return {std::min(a.v, b.v), a.f32 && b.f32}; }

// ---------------------------------------------------------------------------------------------
// Exact nearest-neighbour search over an arbitrary point cloud (stand-in for NearestNeighbors.jl).
// Ranking: squared distance accumulated dimension by dimension in float when both the data and the
// query are Float32 (double otherwise), ties broken by the lower index.
struct KDTree {
  int nd = 0;
  int64_t n = 0;
  bool f32 = true;
  std::vector<double> pts;  // n x nd
  struct Node { int dim; double split; int64_t lo, hi, left, right; };
  std::vector<Node> nodes;
  std::vector<double> bbox;  // per node: min[nd], max[nd] of its points (tight lower bounds also for far-away queries)
  std::vector<int64_t> perm;

  void build(int nd_, int64_t n_, const double* p, bool f32_);
  void build_f(int nd_, int64_t n_, const float* p);
  double d2(int64_t i, const double* x, bool xf32) const;
  // k nearest: idx/d2 sorted ascending by (d2, idx); returns number found (min(k, n))
  // max_d2: only points with d2 <= max_d2 are of interest (prunes the search of far-away queries from the start)
  int knn(const double* x, bool xf32, int k, int64_t* idx, double* d2out, double max_d2 = 1e300) const;
  // all points with d2 <= r^2 (inclusive), ascending index
  void inrange(const double* x, bool xf32, double r, std::vector<int64_t>& out) const;
  double box_d2(int64_t node, const double* x) const {
    const double* b = &bbox[(size_t)node * 2 * nd];
    double acc = 0;
    for (int d = 0; d < nd; ++d) {
      double t = x[d] < b[d] ? b[d] - x[d] : (x[d] > b[nd + d] ? x[d] - b[nd + d] : 0.0);
      acc += t * t;
    }
    return acc;
  }
 private:
  int64_t n_nodes_ = 0;
  int64_t build_rec(int64_t lo, int64_t hi);
};

}  // namespace ibx

// ------------------------------------------------------------------------------------- host objects
struct ibx_stl {
  int nd = 0;
  bool f32 = true;
  std::vector<double> points;      // np x nd
  std::vector<int64_t> simplices;  // ns x nd (segments in 2-D, triangles in 3-D)
  int64_t npoints() const { return nd ? (int64_t)points.size() / nd : 0; }
  int64_t nsimp() const { return nd ? (int64_t)simplices.size() / nd : 0; }
};

struct ibx_dfield {
  std::shared_ptr<ibx_stl> stl;
  std::vector<double> centers;  // ns x nd
  double rmax = 0;              // largest centre-to-vertex distance over all simplices (upper bound, see projection)
  ibx::KDTree tree;
  // analytic sphere alternative (stl == nullptr)
  bool sphere = false;
  double sc[3] = {0, 0, 0};
  double sr = 0;
  ibx::Num distance(const double* x, bool xf32) const;
  // distance(x) if it is <= r (evaluated like distance(), compared like `distance(x) <= r`), else false -- without paying
  // for an exact nearest-neighbour search when x is far from the surface
  bool distance_within(const double* x, bool xf32, double r, ibx::Num* out) const;
  // projection (src/mesher.jl:778-801); result in the promoted type
  void projection(const double* x, bool xf32, double R, double* out) const;
};

struct ibx_mesh {
  int nd = 0;
  int block_size = 8;
  float origin[3] = {0, 0, 0}, widths[3] = {0, 0, 0};
  std::vector<float> block_origins, block_widths;  // nb x nd
  std::vector<std::string> surf_names;
  std::vector<std::shared_ptr<ibx_dfield>> surf_fields;
  int64_t nblocks() const { return nd ? (int64_t)block_origins.size() / nd : 0; }
  int64_t cells_per_block() const { int64_t c = 1; for (int d = 0; d < nd; ++d) c *= block_size; return c; }
  int64_t ncells() const { return nblocks() * cells_per_block(); }
};

struct ibx_accum {
  int64_t n_out = 0;
  bool weighted = true;
  std::vector<int32_t> ptr, idx;
  std::vector<float> w;
  // device copies
  int32_t *d_ptr = nullptr, *d_idx = nullptr;
  float* d_w = nullptr;
  bool uploaded = false;
  ~ibx_accum();
};

namespace ibx {

// Centre / width of any cell straight from the block tables, with get_cells' float32 operations
// (src/mesher.jl:1064-1112): centre = fl(fl((i + 1/2) / bs) * block_width) + block_origin, width = block_width / bs.
// mesh_cells() fills its arrays through this, so an array entry and a value computed on the fly are the same bits;
// a rank-restricted Domain build (ibx_domain_build_for_rank) never materialises the global arrays.
struct CellGeom {
  const float *bo = nullptr, *bw = nullptr;
  int nd = 0, bs = 0, bs2 = 0;
  int64_t cpb = 0;
  float ic[64];
  void init(const ibx_mesh& m) {
    bo = m.block_origins.data();
    bw = m.block_widths.data();
    nd = m.nd;
    bs = m.block_size;
    bs2 = bs * bs;
    cpb = m.cells_per_block();
    for (int i = 0; i < bs && i < 64; ++i) ic[i] = ((float)i + 0.5f) / (float)bs;
  }
  // centre along d of the cell with index i along d in block b
  inline float cb(int64_t b, int i, int d) const {
    float prod = ic[i] * bw[b * nd + d];
    return prod + bo[b * nd + d];
  }
  inline float c(int64_t cell, int d) const {
    const int64_t b = cell / cpb;
    const int l = (int)(cell - b * cpb);
    const int i = d == 0 ? l % bs : (d == 1 ? (l / bs) % bs : l / bs2);
    return cb(b, i, d);
  }
  inline float w(int64_t cell, int d) const { return bw[(cell / cpb) * nd + d] / (float)bs; }
  inline void center(int64_t cell, float* out) const { for (int d = 0; d < nd; ++d) out[d] = c(cell, d); }
  inline void width(int64_t cell, float* out) const { for (int d = 0; d < nd; ++d) out[d] = w(cell, d); }
};

struct FaceTable {  // per partition, per dim
  std::vector<int32_t> owners, neighbors;
  std::vector<int32_t> lptr, lidx, rptr, ridx;  // left/right face lists, CSR over domain cells
  int32_t *d_owners = nullptr, *d_neighbors = nullptr, *d_lptr = nullptr, *d_lidx = nullptr, *d_rptr = nullptr,
          *d_ridx = nullptr;
};

struct PartitionT {
  int64_t image_start = 0, n_image = 0;
  std::vector<int32_t> domain, image_in_domain;
  std::vector<FaceTable> dims;
  int32_t *d_domain = nullptr, *d_image_in_domain = nullptr;
  float *d_spacing = nullptr, *d_centers = nullptr;  // n_domain x nd, column-major
};

struct BoundaryT {
  std::vector<int32_t> ghost;
  std::vector<float> proj, normals;  // G x nd row-major
  std::vector<float> image_dist, ghost_dist;
  std::vector<int32_t> image_domain;
  std::vector<int32_t> ptr, idx;  // idx into image_domain
  std::vector<float> w;
  int64_t n_tied = 0;  // ghosts whose k-th and (k+1)-th nearest donor candidates are exactly equidistant (tie rule applies)
  // device copies (idx_global = image_domain[idx])
  int32_t *d_ghost = nullptr, *d_ptr = nullptr, *d_idx_global = nullptr, *d_image_domain = nullptr, *d_idx = nullptr;
  float *d_w = nullptr, *d_normals = nullptr /* G x nd col-major */, *d_eta = nullptr;
};

struct BoundaryFamily {
  std::string name;
  std::vector<BoundaryT> parts;
};

struct SurfaceT {
  std::string name;
  std::vector<float> points, normals;  // np x nd row-major
  std::vector<float> offsets, areas;
  ibx_accum interp, offset_interp;
  float* d_areas = nullptr;
};

// Block-structured connectivity for the fused kernels: per block and per block-face, the
// neighbouring block(s).  kind: 0 boundary (none), 1 same level, 2 coarser, 3 finer.
struct BlockFace {
  int32_t kind;
  int32_t nb[4];   // same/coarser: nb[0]; finer: 2^(nd-1) blocks ordered tangential-dim-1 fastest
  int32_t sub[2];  // coarser: which half of the coarse face along each tangential dim (0/1)
};

struct Shard {  // multi-GPU: rank-local view
  bool active = false;
  int rank = 0, nranks = 1;
  int64_t owned_start = 0, n_owned = 0, n_halo = 0;
  std::vector<int32_t> local_to_global;               // n_owned + n_halo
  std::vector<std::vector<int32_t>> send_local, recv_local;  // per peer
  std::vector<int32_t*> d_send, d_recv;
  std::vector<float*> d_sendbuf, d_recvbuf;
  std::vector<int64_t> buf_cap;
};

}  // namespace ibx

struct ibx_domain {
  int nd = 0;
  int block_size = 8;
  int64_t ncells = 0;
  std::shared_ptr<ibx_mesh> mesh;  // copy of the mesh description
  std::vector<float> centers, widths;  // ncells x nd row-major (host)
  std::vector<int32_t> faces;          // nf x 3
  int64_t n_interior_faces = 0;
  std::vector<ibx::PartitionT> parts;
  std::vector<ibx::BoundaryFamily> boundaries;
  std::vector<ibx::SurfaceT> surfaces;
  // block structure
  std::vector<ibx::BlockFace> block_faces;  // nblocks x 2*nd, face order: dim-major, low then high
  std::vector<float> block_h;               // nblocks x nd cell widths
  bool two_to_one = true;                   // every inter-block face is same level or 2:1
  // device state
  bool uploaded = false;
  int device = -1;
  ibx::BlockFace* d_block_faces = nullptr;
  float* d_block_h = nullptr;
  // block work lists of the tile kernels: blocks without / with a finer neighbour; all local vs owned only
  int32_t *d_blk_all_plain = nullptr, *d_blk_all_finer = nullptr, *d_blk_own_plain = nullptr, *d_blk_own_finer = nullptr;
  int32_t *d_blk_own_regular = nullptr, *d_blk_all_regular = nullptr;  // blocks whose 2*nd neighbours are all same-level blocks (lean kernels)
  int n_all_plain = 0, n_all_finer = 0, n_own_plain = 0, n_own_finer = 0, n_own_regular = 0, n_all_regular = 0;
  // Shards: block lists of the two phases of the overlapped step (ibx_step_euler_sharded).  Phase 0 ("early") holds the
  // work that depends neither on a ghost cell nor on a halo cell -- flux blocks whose two rings of face neighbours are
  // owned and ghost-free, the sensor blocks (one ring) and primitive blocks (two rings) they read -- and runs while the
  // first exchange, the ghost update and the second exchange are in flight; phase 1 ("late") is everything else.
  struct PhaseLists {
    enum { PRIM = 0, S_REG, S_PLAIN, S_FINER, F_REG, F_PLAIN, F_FINER, NLISTS };
    int32_t* d[NLISTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n[NLISTS] = {0, 0, 0, 0, 0, 0, 0};
  } phase[2];
  bool phased = false;
  int32_t *d_blk_noghost = nullptr, *d_blk_ghost = nullptr;   // blocks without / with ghost cells (whole-domain overlapped step)
  int n_blk_noghost = 0, n_blk_ghost = 0;
  bool all_pow2 = false;                 // every cell width is a power of two (exact fast paths, physics.cuh)
  ibx::Shard shard;
  ~ibx_domain();
};
