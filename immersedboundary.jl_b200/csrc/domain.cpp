// Host builder of the Domain tables: faces, partitions with 2-deep skirts, immersed-boundary ghost /
// image stencils, surfaces, block-face connectivity.  Mirrors src/ImmersedBoundary.jl:63-786 and
// src/nninterp.jl:16-138; every step cites the lines it restates.  O(N) and multithreaded: the
// reference's KD-tree `inrange` over all cells is replaced by block-level candidate search followed by
// the reference's own float32 tests on the same float32 values, so the resulting sets are identical.
#include "ibx_internal.h"
#include <cstdio>
#include <cstdlib>

#include <numeric>
#include <unordered_map>
#include <omp.h>

namespace ibx {

std::shared_ptr<ibx_mesh> lookup_mesh(const ibx_mesh* m);
void mesh_cells(const ibx_mesh& m, float* centers, float* widths);
void pinv_small(const double* A, int m, int n, double rtol, double* out);
void stl_centers_normals(const ibx_stl& s, std::vector<double>& c, std::vector<double>& n);

static const float EPS32 = 1.1920928955078125e-07f;

// ---------------------------------------------------------------------------------------------
// Block index: blocks grouped by size class, one KD-tree of block centres per class.
struct BlockIndex {
  int nd, bs;
  int64_t nb, cpb;
  const float *bo, *bw;     // block origins / widths (nb x nd)
  CellGeom geo;             // cell centres / widths, computed from the block tables (same bits as get_cells' arrays)
  struct Class {
    float w[3];
    double R;
    KDTree tree;
    std::vector<int64_t> ids;
  };
  std::vector<Class> classes;

  void build(const ibx_mesh& m) {
    nd = m.nd;
    bs = m.block_size;
    nb = m.nblocks();
    cpb = m.cells_per_block();
    bo = m.block_origins.data();
    bw = m.block_widths.data();
    geo.init(m);
    std::map<std::vector<float>, int> key2class;
    std::vector<std::vector<double>> pts;
    for (int64_t b = 0; b < nb; ++b) {
      std::vector<float> key(bw + b * nd, bw + (b + 1) * nd);
      auto it = key2class.find(key);
      int c;
      if (it == key2class.end()) {
        c = (int)classes.size();
        key2class[key] = c;
        classes.emplace_back();
        pts.emplace_back();
        double ss = 0;
        for (int d = 0; d < nd; ++d) { classes[c].w[d] = key[d]; ss += (double)key[d] * key[d]; }
        classes[c].R = std::sqrt(ss) / 2;
      } else {
        c = it->second;
      }
      classes[c].ids.push_back(b);
      for (int d = 0; d < nd; ++d) pts[c].push_back((double)bo[b * nd + d] + 0.5 * (double)bw[b * nd + d]);
    }
    for (size_t c = 0; c < classes.size(); ++c)
      classes[c].tree.build(nd, (int64_t)classes[c].ids.size(), pts[c].data(), false);
  }

  double aabb_dist2(int64_t b, const double* x) const {
    double s = 0;
    for (int d = 0; d < nd; ++d) {
      double lo = bo[b * nd + d], hi = lo + (double)bw[b * nd + d];
      double g = x[d] < lo ? lo - x[d] : (x[d] > hi ? x[d] - hi : 0.0);
      s += g * g;
    }
    return s;
  }

  // blocks whose box lies within rho of x
  void blocks_near(const double* x, double rho, std::vector<int64_t>& out, std::vector<int64_t>& scratch) const {
    out.clear();
    for (const Class& c : classes) {
      c.tree.inrange(x, false, rho + c.R * (1 + 1e-6), scratch);
      for (int64_t k : scratch) {
        int64_t b = c.ids[k];
        if (aabb_dist2(b, x) <= rho * rho * (1 + 1e-6)) out.push_back(b);
      }
    }
  }

  static inline float d2f(const float* a, const float* b, int nd) {
    float acc = 0.f;
    for (int d = 0; d < nd; ++d) {
      float df = a[d] - b[d];
      float sq = df * df;
      acc = d == 0 ? sq : acc + sq;
    }
    return acc;
  }

  // exact k nearest cell centres to x (float32 point), ranked by (float d2, index)
  int knn(const float* x, int k, double rho0, int64_t* idx, float* d2out) const {
    double xd[3];
    for (int d = 0; d < nd; ++d) xd[d] = x[d];
    std::vector<int64_t> blocks, scratch;
    std::vector<std::pair<float, int64_t>> cand;
    double rho = rho0;
    for (int iter = 0; iter < 60; ++iter, rho *= 2) {
      cand.clear();
      blocks_near(xd, rho, blocks, scratch);
      float r2 = (float)(rho * rho);
      for (int64_t b : blocks) {
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        bool empty = false;
        for (int d = 0; d < nd; ++d) {
          double cw = (double)bw[b * nd + d] / bs, o = bo[b * nd + d];
          double l = std::floor((xd[d] - rho - o) / cw - 0.5) - 1;
          double h = std::ceil((xd[d] + rho - o) / cw - 0.5) + 1;
          lo[d] = (int)std::max(l, 0.0);
          hi[d] = (int)std::min(h, (double)(bs - 1));
          if (l > bs || h < -1) empty = true;
          if (lo[d] > hi[d]) empty = true;
        }
        if (empty) continue;
        float cc[3] = {0.f, 0.f, 0.f};
        for (int i2 = lo[2]; i2 <= (nd > 2 ? hi[2] : 0); ++i2) {
          if (nd > 2) cc[2] = geo.cb(b, i2, 2);
          for (int i1 = lo[1]; i1 <= hi[1]; ++i1) {
            cc[1] = geo.cb(b, i1, 1);
            for (int i0 = lo[0]; i0 <= hi[0]; ++i0) {
              int64_t cell = b * cpb + i0 + (int64_t)bs * (i1 + (int64_t)bs * i2);
              cc[0] = geo.cb(b, i0, 0);
              float dd = d2f(cc, x, nd);
              if (dd <= r2) cand.emplace_back(dd, cell);
            }
          }
        }
      }
      if ((int)cand.size() >= k) {
        std::partial_sort(cand.begin(), cand.begin() + k, cand.end());
        for (int j = 0; j < k; ++j) { d2out[j] = cand[j].first; idx[j] = cand[j].second; }
        return k;
      }
      if ((int64_t)cand.size() == nb * cpb) break;
    }
    std::sort(cand.begin(), cand.end());
    int n = (int)std::min<size_t>(cand.size(), (size_t)k);
    for (int j = 0; j < n; ++j) { d2out[j] = cand[j].first; idx[j] = cand[j].second; }
    return n;
  }

  // cell width (max over dims) of the block containing x, or of the nearest block
  double local_width(const float* x) const {
    double xd[3];
    for (int d = 0; d < nd; ++d) xd[d] = x[d];
    std::vector<int64_t> scratch;
    double best = -1, bestd = 1e300;
    for (const Class& c : classes) {
      c.tree.inrange(xd, false, c.R * (1 + 1e-6), scratch);
      for (int64_t k : scratch) {
        double dd = aabb_dist2(c.ids[k], xd);
        double cw = std::max({(double)c.w[0], (double)c.w[1], nd > 2 ? (double)c.w[2] : 0.0}) / bs;
        if (dd < bestd || (dd == bestd && cw < best)) { bestd = dd; best = cw; }
      }
    }
    if (best < 0) {  // outside every circumsphere: fall back to the globally nearest block centre
      for (const Class& c : classes) {
        int64_t i;
        double dd;
        if (c.tree.knn(xd, false, 1, &i, &dd) && dd < bestd) {
          bestd = dd;
          best = std::max({(double)c.w[0], (double)c.w[1], nd > 2 ? (double)c.w[2] : 0.0}) / bs;
        }
      }
      best = std::max(best, std::sqrt(bestd));
    }
    return best;
  }
};

// ---------------------------------------------------------------------------------------------
// the reference's face test between cells i (candidate owner) and j (src/ImmersedBoundary.jl:83-118)
static inline int face_test(const float* ci, const float* wi, const float* cj, const float* wj, int nd) {
  float ri2 = 0.f;
  for (int d = 0; d < nd; ++d) { float sq = wi[d] * wi[d]; ri2 = d == 0 ? sq : ri2 + sq; }
  float rr = (std::sqrt(ri2) / 2.0f) * 3.1f;
  if (BlockIndex::d2f(cj, ci, nd) > rr * rr) return -1;  // not returned by inrange(tree, c_i, 3.1 r_i)
  float fw[3], mx = -INFINITY;
  float oi[3], oj[3];
  for (int d = 0; d < nd; ++d) {
    oi[d] = ci[d] - wi[d] / 2.0f;
    oj[d] = cj[d] - wj[d] / 2.0f;
    float fo = std::max(oi[d], oj[d]);
    fw[d] = std::min(oi[d] + wi[d], oj[d] + wj[d]) - fo;
    mx = std::max(mx, fw[d]);
  }
  float tol = 0.01f * mx;
  int n = 0, nz = 0, arg = 0;
  for (int d = 0; d < nd; ++d) {
    n += fw[d] < tol;
    nz += fw[d] < -tol;
    if (fw[d] < fw[arg]) arg = d;
  }
  if (n != 1 || nz > 0) return -1;
  if (oj[arg] < oi[arg]) return -1;  // j is on the - side: registered from j's own loop
  return arg;
}

struct TouchPair { int64_t b; int dim; };  // block b touches on the + side along dim

static void build_faces(ibx_domain& D, const BlockIndex& bi, bool want_faces) {
  const ibx_mesh& m = *D.mesh;
  int nd = D.nd, bs = D.block_size;
  int64_t nb = bi.nb, cpb = bi.cpb;
  const float* C = D.centers.data();   // per-cell arrays: only read below when the face lists are wanted
  const float* W = D.widths.data();
  if (want_faces && D.centers.empty()) throw std::runtime_error("face lists need the per-cell arrays (not a rank-restricted build)");
  // --- block-face connectivity + candidate (+ side) neighbours of every block
  D.block_faces.assign((size_t)nb * 2 * nd, BlockFace{0, {-1, -1, -1, -1}, {0, 0}});
  D.block_h.resize((size_t)nb * nd);
  std::vector<std::vector<TouchPair>> plus(nb);
  bool ok21 = true;
#pragma omp parallel
  {
    std::vector<int64_t> near, scratch;
#pragma omp for schedule(dynamic, 64) reduction(&& : ok21)
    for (int64_t a = 0; a < nb; ++a) {
      const float* ao = bi.bo + a * nd;
      const float* aw = bi.bw + a * nd;
      double ctr[3], ss = 0;
      for (int d = 0; d < nd; ++d) {
        ctr[d] = (double)ao[d] + 0.5 * aw[d];
        ss += (double)aw[d] * aw[d];
        D.block_h[a * nd + d] = aw[d] / (float)bs;
      }
      double tolA = 0.02 * std::min({(double)aw[0], (double)aw[1], nd > 2 ? (double)aw[2] : 1e300}) / bs;
      bi.blocks_near(ctr, 0.5 * std::sqrt(ss) + tolA, near, scratch);
      std::sort(near.begin(), near.end());
      int cnt[6] = {0, 0, 0, 0, 0, 0};
      for (int64_t b : near) {
        if (b == a) continue;
        const float* bo_ = bi.bo + b * nd;
        const float* bw_ = bi.bw + b * nd;
        // classify the contact: exactly one dim with ~zero overlap, positive overlap in the others
        int tdim = -1, nflat = 0;
        bool sep = false, plus_side = false;
        double tol = std::min(tolA, 0.02 * (double)std::min({bw_[0], bw_[1], nd > 2 ? bw_[2] : bw_[0]}) / bs);
        for (int d = 0; d < nd; ++d) {
          double lo = std::max((double)ao[d], (double)bo_[d]);
          double hi = std::min((double)ao[d] + aw[d], (double)bo_[d] + bw_[d]);
          double ov = hi - lo;
          if (ov < -tol) sep = true;
          else if (ov < tol) { ++nflat; tdim = d; plus_side = (double)bo_[d] > (double)ao[d]; }
        }
        if (sep || nflat != 1) continue;
        int fidx = 2 * tdim + (plus_side ? 1 : 0);
        BlockFace& bf = D.block_faces[(size_t)a * 2 * nd + fidx];
        double ratio = (double)bw_[tdim] / (double)aw[tdim];
        bool iso = true;
        for (int d = 0; d < nd; ++d)
          if (std::fabs((double)bw_[d] / (double)aw[d] - ratio) > 1e-4) iso = false;
        int t1 = (tdim + 1) % nd, t2 = (tdim + 2) % nd;
        if (nd == 2) t2 = t1;
        if (nd == 3 && t1 > t2) std::swap(t1, t2);  // tangential dims in increasing order
        if (iso && std::fabs(ratio - 1.0) < 1e-4) {
          bf.kind = 1;
          bf.nb[0] = (int32_t)b;
        } else if (iso && std::fabs(ratio - 2.0) < 1e-4) {
          bf.kind = 2;
          bf.nb[0] = (int32_t)b;
          bf.sub[0] = ((double)ao[t1] - (double)bo_[t1]) > 0.25 * bw_[t1] ? 1 : 0;
          bf.sub[1] = (nd == 3 && ((double)ao[t2] - (double)bo_[t2]) > 0.25 * bw_[t2]) ? 1 : 0;
        } else if (iso && std::fabs(ratio - 0.5) < 1e-4) {
          bf.kind = 3;
          int s1 = ((double)bo_[t1] - (double)ao[t1]) > 0.25 * aw[t1] ? 1 : 0;
          int s2 = (nd == 3 && ((double)bo_[t2] - (double)ao[t2]) > 0.25 * aw[t2]) ? 1 : 0;
          bf.nb[s1 + 2 * s2] = (int32_t)b;
        } else {
          ok21 = false;
          bf.kind = 4;  // irregular contact: only the table-driven path can handle it
        }
        ++cnt[fidx];
        if (plus_side) plus[a].push_back({b, tdim});
      }
      for (int f = 0; f < 2 * nd; ++f) {
        const BlockFace& bf = D.block_faces[(size_t)a * 2 * nd + f];
        int expect = bf.kind == 3 ? (nd == 3 ? 4 : 2) : (bf.kind == 0 ? 0 : 1);
        if (bf.kind != 4 && cnt[f] != expect) ok21 = false;
      }
    }
  }
  D.two_to_one = ok21;
  if (!want_faces) return;

  // --- faces, generated per owner cell in (owner, neigh) order: intra-block +e_d, then + side blocks
  std::vector<std::vector<int32_t>> per_block(nb);  // (dim, owner, neigh) triples
  bool intra_ok = true;
#pragma omp parallel for schedule(dynamic, 16) reduction(&& : intra_ok)
  for (int64_t a = 0; a < nb; ++a) {
    std::vector<int32_t>& out = per_block[a];
    out.reserve((size_t)cpb * nd * 3 + 64);
    std::vector<std::pair<int64_t, int>> mine;  // (neigh, dim) of one owner
    for (int64_t l = 0; l < cpb; ++l) {
      int64_t i = a * cpb + l;
      int ii[3] = {(int)(l % bs), (int)((l / bs) % bs), (int)(l / ((int64_t)bs * bs))};
      mine.clear();
      int64_t stride = 1;
      for (int d = 0; d < nd; ++d) {
        if (ii[d] < bs - 1) {
          int64_t j = i + stride;
          int r = face_test(C + i * nd, W + i * nd, C + j * nd, W + j * nd, nd);
          if (r != d) intra_ok = false;
          mine.emplace_back(j, d);
        }
        stride *= bs;
      }
      for (const TouchPair& tp : plus[a]) {
        if (ii[tp.dim] != bs - 1) continue;
        int64_t b = tp.b;
        // window of b's first layer overlapping cell i tangentially
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        bool empty = false;
        for (int d = 0; d < nd; ++d) {
          if (d == tp.dim) { lo[d] = hi[d] = 0; continue; }
          double cw = (double)bi.bw[b * nd + d] / bs, o = bi.bo[b * nd + d];
          double mn = (double)C[i * nd + d] - 0.5 * W[i * nd + d], mxv = mn + W[i * nd + d];
          lo[d] = std::max(0, (int)std::floor((mn - o) / cw) - 1);
          hi[d] = std::min(bs - 1, (int)std::floor((mxv - o) / cw) + 1);
          if (lo[d] > hi[d]) empty = true;
        }
        if (empty) continue;
        for (int j2 = lo[2]; j2 <= (nd > 2 ? hi[2] : 0); ++j2)
          for (int j1 = lo[1]; j1 <= hi[1]; ++j1)
            for (int j0 = lo[0]; j0 <= hi[0]; ++j0) {
              int64_t j = b * cpb + j0 + (int64_t)bs * (j1 + (int64_t)bs * j2);
              int r = face_test(C + i * nd, W + i * nd, C + j * nd, W + j * nd, nd);
              if (r >= 0) mine.emplace_back(j, r);
            }
      }
      std::sort(mine.begin(), mine.end());
      for (auto& f : mine) {
        out.push_back(f.second);
        out.push_back((int32_t)i);
        out.push_back((int32_t)f.first);
      }
    }
  }
  if (!intra_ok) throw std::runtime_error("degenerate block: an intra-block cell pair fails the reference face test");
  std::vector<int64_t> off(nb + 1, 0);
  for (int64_t a = 0; a < nb; ++a) off[a + 1] = off[a] + (int64_t)per_block[a].size();
  D.faces.resize((size_t)off[nb]);
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < nb; ++a) std::copy(per_block[a].begin(), per_block[a].end(), D.faces.begin() + off[a]);
  D.n_interior_faces = off[nb] / 3;
  per_block.clear();
  // --- hcube_faces (src/ImmersedBoundary.jl:150-184)
  int64_t N = D.ncells;
  for (int dim = 0; dim < nd; ++dim) {
    for (int64_t i = 0; i < N; ++i) {
      float o = C[i * nd + dim] - W[i * nd + dim] / 2.0f;
      if (std::fabs(o - m.origin[dim]) < W[i * nd + dim] * 0.01f) {
        D.faces.push_back(dim); D.faces.push_back(-1); D.faces.push_back((int32_t)i);
      }
    }
    for (int64_t i = 0; i < N; ++i) {
      float o = C[i * nd + dim] - W[i * nd + dim] / 2.0f;
      float v = ((o + W[i * nd + dim]) - m.origin[dim]) - m.widths[dim];
      if (std::fabs(v) < W[i * nd + dim] * 0.01f) {
        D.faces.push_back(dim); D.faces.push_back((int32_t)i); D.faces.push_back(-1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
static void build_partitions(ibx_domain& D, int64_t max_part, int skirt_depth) {
  // src/ImmersedBoundary.jl:575-703
  int nd = D.nd;
  int64_t N = D.ncells, nf = (int64_t)D.faces.size() / 3;
  const int32_t* F = D.faces.data();
  // cells2faces (:575-585)
  std::vector<int64_t> cptr(N + 1, 0);
  for (int64_t f = 0; f < nf; ++f) {
    if (F[f * 3 + 1] >= 0) ++cptr[F[f * 3 + 1] + 1];
    if (F[f * 3 + 2] >= 0) ++cptr[F[f * 3 + 2] + 1];
  }
  for (int64_t i = 0; i < N; ++i) cptr[i + 1] += cptr[i];
  std::vector<int32_t> cidx((size_t)cptr[N]);
  {
    std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
    for (int64_t f = 0; f < nf; ++f) {
      if (F[f * 3 + 1] >= 0) cidx[fill[F[f * 3 + 1]]++] = (int32_t)f;
      if (F[f * 3 + 2] >= 0) cidx[fill[F[f * 3 + 2]]++] = (int32_t)f;
    }
  }
  int64_t nparts = (N + max_part - 1) / max_part;
  D.parts.resize((size_t)nparts);
#pragma omp parallel
  {
    std::vector<int32_t> loc(N, -1);
    std::vector<uint8_t> seen(nf, 0);
#pragma omp for schedule(dynamic, 1)
    for (int64_t p = 0; p < nparts; ++p) {
      PartitionT& P = D.parts[p];
      int64_t s = p * max_part, e = std::min(N, s + max_part);
      P.image_start = s;
      P.n_image = e - s;
      std::vector<int32_t> dom;
      dom.reserve((size_t)((e - s) * 1.3) + 64);
      for (int64_t c = s; c < e; ++c) { dom.push_back((int32_t)c); loc[c] = 0; }
      size_t begin = 0;
      for (int round = 0; round < skirt_depth; ++round) {
        size_t end = dom.size();
        // every current domain cell is revisited by the reference; only the newest ring can add cells
        for (size_t q = begin; q < end; ++q) {
          int64_t c = dom[q];
          for (int64_t k = cptr[c]; k < cptr[c + 1]; ++k) {
            int32_t o = F[cidx[k] * 3 + 1], n = F[cidx[k] * 3 + 2];
            if (o >= 0 && loc[o] < 0) { loc[o] = 0; dom.push_back(o); }
            if (n >= 0 && loc[n] < 0) { loc[n] = 0; dom.push_back(n); }
          }
        }
        begin = end;
      }
      std::sort(dom.begin(), dom.end());
      for (size_t k = 0; k < dom.size(); ++k) loc[dom[k]] = (int32_t)k;
      P.domain = dom;
      P.image_in_domain.resize((size_t)(e - s));
      for (int64_t c = s; c < e; ++c) P.image_in_domain[c - s] = loc[c];
      // face_indices: first appearance over the sorted domain (:628)
      std::vector<int32_t> fidx;
      for (int32_t c : dom)
        for (int64_t k = cptr[c]; k < cptr[c + 1]; ++k)
          if (!seen[cidx[k]]) { seen[cidx[k]] = 1; fidx.push_back(cidx[k]); }
      int64_t nloc = (int64_t)dom.size();
      P.dims.assign(nd, FaceTable{});
      for (int dim = 0; dim < nd; ++dim) {
        FaceTable& T = P.dims[dim];
        std::vector<int32_t> lcount(nloc + 1, 0), rcount(nloc + 1, 0);
        std::vector<uint8_t> addl, addr;
        for (int32_t f : fidx) {
          if (F[f * 3] != dim) continue;
          int32_t o = F[f * 3 + 1] >= 0 ? loc[F[f * 3 + 1]] : -1;
          int32_t n = F[f * 3 + 2] >= 0 ? loc[F[f * 3 + 2]] : -1;
          bool al = true, ar = true;
          if (o < 0) { o = n; ar = false; }
          if (n < 0) { n = o; al = false; }
          T.owners.push_back(o);
          T.neighbors.push_back(n);
          addl.push_back(al);
          addr.push_back(ar);
          if (al) ++lcount[n + 1];
          if (ar) ++rcount[o + 1];
        }
        for (int64_t i = 0; i < nloc; ++i) { lcount[i + 1] += lcount[i]; rcount[i + 1] += rcount[i]; }
        T.lptr = lcount;
        T.rptr = rcount;
        T.lidx.resize((size_t)lcount[nloc]);
        T.ridx.resize((size_t)rcount[nloc]);
        std::vector<int32_t> lf(lcount.begin(), lcount.end() - 1), rf(rcount.begin(), rcount.end() - 1);
        for (size_t k = 0; k < T.owners.size(); ++k) {
          if (addl[k]) T.lidx[lf[T.neighbors[k]]++] = (int32_t)k;
          if (addr[k]) T.ridx[rf[T.owners[k]]++] = (int32_t)k;
        }
      }
      for (int32_t f : fidx) seen[f] = 0;
      for (int32_t c : dom) loc[c] = -1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
static inline float diam_of(const float* w, int nd) {
  float acc = 0.f;
  for (int d = 0; d < nd; ++d) { float sq = w[d] * w[d]; acc = d == 0 ? sq : acc + sq; }
  return std::sqrt(acc);
}

// linear_weights (src/nninterp.jl:16-42) / IDW_weights (:47-69) about the point x for donors idx[0..k)
struct CloudPoints {  // donors of an arbitrary point cloud (row-major n x nd)
  const float* X;
  int nd;
  inline float c(int64_t i, int d) const { return X[i * nd + d]; }
};
template <class Donors>   // CellGeom (cells of the mesh) or CloudPoints
static int stencil_weights(const Donors& C, int nd, const int64_t* idx, int k, const float* x, bool linear,
                           int64_t* oidx, float* ow) {
  float w[16], dX[16 * 3];
  for (int j = 0; j < k; ++j) {
    float acc = 0.f;
    for (int d = 0; d < nd; ++d) {
      float df = C.c(idx[j], d) - x[d];
      dX[j * nd + d] = df;
      float sq = df * df;
      acc = d == 0 ? sq : acc + sq;
    }
    w[j] = 1.0f / (std::sqrt(acc) + EPS32);
  }
  int n = 0;
  if (linear) {
    int nc = nd + 1;
    double A[16 * 4], P[4 * 16];
    for (int j = 0; j < k; ++j) {
      for (int d = 0; d < nd; ++d) A[j * nc + d] = (double)(float)(dX[j * nd + d] * w[j]);
      A[j * nc + nd] = (double)w[j];  // 1 * w
    }
    pinv_small(A, k, nc, (double)EPS32 * std::min(k, nc), P);
    for (int j = 0; j < k; ++j) {
      float wt = (float)P[nd * k + j] * w[j];
      if (std::fabs(wt) > EPS32) { oidx[n] = idx[j]; ow[n] = wt; ++n; }
    }
  } else {
    float s = w[0];
    for (int j = 1; j < k; ++j) s = s + w[j];
    float thr = std::sqrt(EPS32);
    for (int j = 0; j < k; ++j) {
      float wt = w[j] / s;
      if (std::fabs(wt) > thr) { oidx[n] = idx[j]; ow[n] = wt; ++n; }
    }
  }
  return n;
}

static void build_boundary(ibx_domain& D, const BlockIndex& bi, const std::vector<int32_t>& ghosts,
                           const std::vector<float>& projs, int64_t max_part, float ghost_ratio, BoundaryFamily& fam) {
  // boundary_partitions (src/ImmersedBoundary.jl:456-476) + Boundary (:422-448)
  int nd = D.nd;
  int k = 1 << nd;
  const CellGeom& C = bi.geo;
  int64_t G = (int64_t)ghosts.size();
  for (int64_t s = 0; s < G; s += max_part) {
    int64_t e = std::min(G, s + max_part), n = e - s;
    BoundaryT B;
    B.ghost.assign(ghosts.begin() + s, ghosts.begin() + e);
    B.proj.assign(projs.begin() + s * nd, projs.begin() + e * nd);
    B.normals.resize((size_t)n * nd);
    B.image_dist.resize(n);
    B.ghost_dist.resize(n);
    std::vector<int64_t> sidx((size_t)n * k);
    std::vector<float> sw((size_t)n * k);
    std::vector<int32_t> cnt(n + 1, 0);
    int64_t n_tied = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_tied)
    for (int64_t g = 0; g < n; ++g) {
      int64_t c = B.ghost[g];
      float nrm[3], acc = 0.f, img[3];
      for (int d = 0; d < nd; ++d) {
        nrm[d] = C.c(c, d) - B.proj[g * nd + d];
        float sq = nrm[d] * nrm[d];
        acc = d == 0 ? sq : acc + sq;
      }
      float gd = std::sqrt(acc);
      float wc[3];
      C.width(c, wc);
      float idist = diam_of(wc, nd) * ghost_ratio + EPS32;
      for (int d = 0; d < nd; ++d) {
        nrm[d] = nrm[d] / (gd + EPS32);
        B.normals[g * nd + d] = nrm[d];
        float t = nrm[d] * idist;
        img[d] = B.proj[g * nd + d] + t;
      }
      B.ghost_dist[g] = gd;
      B.image_dist[g] = idist;
      int64_t idx[17];
      float dd[17];
      double cw = std::max({(double)wc[0], (double)wc[1], nd > 2 ? (double)wc[2] : 0.0});
      // one candidate more than needed: a k-th / (k+1)-th distance tie is the only place where NearestNeighbors.jl's
      // traversal order (not reproducible) could select another donor than the (distance, index) rule
      int found = bi.knn(img, k + 1, 1.5 * cw, idx, dd);
      if (found == k + 1 && dd[k - 1] == dd[k]) ++n_tied;
      found = std::min(found, k);
      cnt[g + 1] = stencil_weights(C, nd, idx, found, img, true, &sidx[g * k], &sw[g * k]);
    }
    B.n_tied = n_tied;
    for (int64_t g = 0; g < n; ++g) cnt[g + 1] += cnt[g];
    B.ptr = cnt;
    std::vector<int32_t> gidx((size_t)cnt[n]);
    B.w.resize((size_t)cnt[n]);
    for (int64_t g = 0; g < n; ++g)
      for (int j = 0; j < cnt[g + 1] - cnt[g]; ++j) {
        gidx[cnt[g] + j] = (int32_t)sidx[g * k + j];
        B.w[cnt[g] + j] = sw[g * k + j];
      }
    // NNInterpolator.domain / re_index! (src/nninterp.jl:147-183)
    B.image_domain = gidx;
    std::sort(B.image_domain.begin(), B.image_domain.end());
    B.image_domain.erase(std::unique(B.image_domain.begin(), B.image_domain.end()), B.image_domain.end());
    B.idx.resize(gidx.size());
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)gidx.size(); ++q)
      B.idx[q] = (int32_t)(std::lower_bound(B.image_domain.begin(), B.image_domain.end(), gidx[q]) - B.image_domain.begin());
    fam.parts.push_back(std::move(B));
  }
}

static void ghosts_hcube(const ibx_domain& D, const CellGeom& C, const std::vector<std::pair<int, int>>& faces, float glr,
                         std::vector<int32_t>& ghosts, std::vector<float>& projs, int64_t r0, int64_t r1) {
  // src/ImmersedBoundary.jl:258-305; only the cells of [r0, r1) are examined (and only they get scratch)
  int nd = D.nd;
  const ibx_mesh& m = *D.mesh;
  const int64_t n = r1 - r0;
  std::vector<uint8_t> mask((size_t)n, 0);
  std::vector<int8_t> which((size_t)n, -1);
#pragma omp parallel for schedule(static)
  for (int64_t i = r0; i < r1; ++i) {
    float best = INFINITY;
    float wi[3];
    C.width(i, wi);
    float lim = diam_of(wi, nd) * glr;
    for (size_t f = 0; f < faces.size(); ++f) {
      int dim = faces[f].first;
      float plane = faces[f].second ? (m.origin[dim] + m.widths[dim]) : m.origin[dim];
      float df = plane - C.c(i, dim);
      float ds = std::sqrt(df * df);
      if (ds < best) { best = ds; which[i - r0] = (int8_t)f; }
      if (ds < lim) mask[i - r0] = 1;
    }
  }
  for (int64_t i = r0; i < r1; ++i)
    if (mask[i - r0]) {
      ghosts.push_back((int32_t)i);
      int dim = faces[which[i - r0]].first;
      float plane = faces[which[i - r0]].second ? (m.origin[dim] + m.widths[dim]) : m.origin[dim];
      for (int d = 0; d < nd; ++d) projs.push_back(d == dim ? plane : C.c(i, d));
    }
}

static void ghosts_surface(const ibx_domain& D, const CellGeom& C, const ibx_dfield& df, float glr, std::vector<int32_t>& ghosts,
                           std::vector<float>& projs, int64_t r0, int64_t r1) {
  // src/ImmersedBoundary.jl:194-230
  int nd = D.nd;
  std::vector<std::vector<float>> tproj(omp_get_max_threads());
  std::vector<std::vector<int32_t>> tghost(omp_get_max_threads());
#pragma omp parallel
  {
    std::vector<float>& lp = tproj[omp_get_thread_num()];
    std::vector<int32_t>& lg = tghost[omp_get_thread_num()];
    // The test of a cell is distance(x) <= lim2 with lim2 = 2 glr diam, the same for all cells of a block.  A block whose
    // centre is farther from the surface than its circumradius + lim2 cannot hold such a cell (triangle inequality), so
    // one bounded query per block prunes the blocks away from the surface -- all but a few per cent of them; the cells
    // of the others go through the reference's own test.  The margin covers the float32 roundings of the cell-level
    // distance (relative 1e-6) a thousand times over.
    const int64_t cpb = C.cpb;
#pragma omp for schedule(dynamic, 16)
    for (int64_t b = r0 / cpb; b < (r1 + cpb - 1) / cpb; ++b) {
      {
        double xc[3], ss = 0;
        for (int d = 0; d < nd; ++d) {
          xc[d] = (double)C.bo[b * nd + d] + 0.5 * (double)C.bw[b * nd + d];
          ss += (double)C.bw[b * nd + d] * (double)C.bw[b * nd + d];
        }
        float wb[3];
        C.width(b * cpb, wb);
        const double reach = (0.5 * std::sqrt(ss) + (double)(diam_of(wb, nd) * glr * 2.0f)) * (1.0 + 1e-3);
        Num far;
        if (!df.distance_within(xc, false, reach, &far)) continue;
      }
      for (int64_t i = std::max(r0, b * cpb); i < std::min(r1, (b + 1) * cpb); ++i) {
        double x[3];
        float ci[3], wi[3];
        C.center(i, ci);
        C.width(i, wi);
        for (int d = 0; d < nd; ++d) x[d] = ci[d];
        float diam = diam_of(wi, nd);
        float lim2 = diam * glr * 2.0f;
        Num dist;
        if (!df.distance_within(x, true, (double)lim2, &dist)) continue;   // == !(df.distance(x) <= lim2), src/ImmersedBoundary.jl:208
        // `dist` is the distance to the nearest simplex CENTRE and every point of a simplex lies within rmax of its centre:
        // the projection cannot be nearer than dist - rmax.  Beyond the ghost band the cell is dropped by the test
        // below whatever the projection is, so the (expensive, far-from-the-surface) projection is not computed.
        if (dist.v - df.rmax > (double)(diam * glr) * (1.0 + 1e-4)) continue;
        double p[3];
        df.projection(x, true, (double)lim2, p);
        float pf[3], acc = 0.f;
        for (int d = 0; d < nd; ++d) {
          pf[d] = (float)p[d];
          float dfv = pf[d] - ci[d];
          float sq = dfv * dfv;
          acc = d == 0 ? sq : acc + sq;
        }
        if (std::sqrt(acc) <= diam * glr) {
          lg.push_back((int32_t)i);
          for (int d = 0; d < nd; ++d) lp.push_back(pf[d]);
        }
      }
    }
  }
  // ghosts in ascending cell order (src/ImmersedBoundary.jl:229 keeps the order of the cell loop)
  std::vector<std::pair<int32_t, std::pair<int, int64_t>>> order;   // (cell, (thread, position))
  for (size_t t = 0; t < tghost.size(); ++t)
    for (size_t j = 0; j < tghost[t].size(); ++j) order.push_back({tghost[t][j], {(int)t, (int64_t)j}});
  std::sort(order.begin(), order.end());
  for (const auto& o : order) {
    ghosts.push_back(o.first);
    const float* q = tproj[o.second.first].data() + o.second.second * nd;
    projs.insert(projs.end(), q, q + nd);
  }
}

static void build_interp(const BlockIndex& bi, int nd, int64_t q, const float* Xc, const float* bias,
                         bool linear, ibx_accum& out) {
  const CellGeom& C = bi.geo;
  // Interpolator over the domain's own cells (src/nninterp.jl:85-138)
  int k = 1 << nd;
  std::vector<int64_t> sidx((size_t)q * k);
  std::vector<float> sw((size_t)q * k);
  std::vector<int32_t> cnt(q + 1, 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t p = 0; p < q; ++p) {
    float xq[3];
    for (int d = 0; d < nd; ++d) xq[d] = bias ? Xc[p * nd + d] + bias[p * nd + d] : Xc[p * nd + d];
    int64_t idx[16];
    float dd[16];
    int found = bi.knn(xq, k, 1.5 * bi.local_width(xq), idx, dd);
    cnt[p + 1] = stencil_weights(C, nd, idx, found, Xc + p * nd, linear, &sidx[p * k], &sw[p * k]);
  }
  for (int64_t p = 0; p < q; ++p) cnt[p + 1] += cnt[p];
  out.n_out = q;
  out.weighted = true;
  out.ptr = cnt;
  out.idx.resize((size_t)cnt[q]);
  out.w.resize((size_t)cnt[q]);
  for (int64_t p = 0; p < q; ++p)
    for (int j = 0; j < cnt[p + 1] - cnt[p]; ++j) {
      out.idx[cnt[p] + j] = (int32_t)sidx[p * k + j];
      out.w[cnt[p] + j] = sw[p * k + j];
    }
}

static std::map<const ibx_domain*, std::shared_ptr<ibx_domain>> g_dom;
static std::map<const ibx_accum*, std::shared_ptr<ibx_accum>> g_acc;

}  // namespace ibx

using namespace ibx;

extern "C" {

static int domain_build_impl(const ibx_mesh* mh, int64_t max_partition_size, int skirt_depth, float ghost_layer_ratio, int nfam,
                             const char* const* fam_names, const int* fam_ptr, const int* fam_dim, const int* fam_front,
                             int build_partitions_flag, int build_surfaces, int rank, int nranks, ibx_domain** out) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  IBX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank out of range");
  // ghosts (and their donor stencils) are only searched among the cells of this rank's block range
  const int64_t g_r0 = (m->nblocks() * rank / nranks) * m->cells_per_block();
  const int64_t g_r1 = (m->nblocks() * (rank + 1) / nranks) * m->cells_per_block();
  IBX_REQUIRE(max_partition_size >= 1, "max_partition_size must be positive");
  IBX_REQUIRE(m->ncells() < (int64_t)2147483647, "more than 2^31-1 cells: Int64 tables are not implemented");
  auto D = std::make_shared<ibx_domain>();
  D->nd = m->nd;
  D->block_size = m->block_size;
  D->mesh = m;
  D->ncells = m->ncells();
  int nd = D->nd;
  // a rank-restricted build (input of ibx_domain_shard) never holds per-cell arrays of the GLOBAL mesh: every centre /
  // width the ghost search needs is computed from the block tables (CellGeom), the shard fills its own local arrays
  if (nranks == 1) {
    D->centers.resize((size_t)D->ncells * nd);
    D->widths.resize((size_t)D->ncells * nd);
    mesh_cells(*m, D->centers.data(), D->widths.data());
  }
  // the reference prints its build phases when `verbose` (src/ImmersedBoundary.jl:589,705,767); here: IBX_BUILD_VERBOSE=1
  const bool verbose = getenv("IBX_BUILD_VERBOSE") != nullptr;
  double t_phase = omp_get_wtime();
  auto phase = [&](const char* what) {
    if (verbose) fprintf(stderr, "[ibx build] %-28s %8.2f s\n", what, omp_get_wtime() - t_phase);
    t_phase = omp_get_wtime();
  };
  BlockIndex bi;
  bi.build(*m);
  phase("cells + block index");
  build_faces(*D, bi, build_partitions_flag != 0);
  phase("faces / block contacts");
  if (build_partitions_flag) build_partitions(*D, max_partition_size, skirt_depth);
  phase("partitions");
  // boundaries: hypercube families first, then one per surface (src/ImmersedBoundary.jl:716-741)
  for (int f = 0; f < nfam; ++f) {
    std::vector<std::pair<int, int>> faces;
    for (int k = fam_ptr[f]; k < fam_ptr[f + 1]; ++k) {
      IBX_REQUIRE(fam_dim[k] >= 0 && fam_dim[k] < nd, "hypercube family dimension out of range");
      faces.emplace_back(fam_dim[k], fam_front[k]);
    }
    std::vector<int32_t> ghosts;
    std::vector<float> projs;
    ghosts_hcube(*D, bi.geo, faces, ghost_layer_ratio, ghosts, projs, g_r0, g_r1);
    phase("hypercube ghosts");
    BoundaryFamily fam;
    fam.name = fam_names[f];
    build_boundary(*D, bi, ghosts, projs, max_partition_size, ghost_layer_ratio, fam);
    phase("hypercube donors + weights");
    D->boundaries.push_back(std::move(fam));
  }
  for (size_t s = 0; s < m->surf_names.size(); ++s) {
    const ibx_dfield& df = *m->surf_fields[s];
    std::vector<int32_t> ghosts;
    std::vector<float> projs;
    ghosts_surface(*D, bi.geo, df, ghost_layer_ratio, ghosts, projs, g_r0, g_r1);
    phase("surface ghosts + projections");
    BoundaryFamily fam;
    fam.name = m->surf_names[s];
    build_boundary(*D, bi, ghosts, projs, max_partition_size, ghost_layer_ratio, fam);
    phase("surface donors + weights");
    D->boundaries.push_back(std::move(fam));
    if (!build_surfaces || !df.stl) continue;
    // Surface (src/ImmersedBoundary.jl:743-763)
    D->surfaces.emplace_back();
    SurfaceT& S = D->surfaces.back();
    S.name = m->surf_names[s];
    std::vector<double> fc, fn;
    stl_centers_normals(*df.stl, fc, fn);
    int64_t np = df.stl->nsimp();
    S.points.resize((size_t)np * nd);
    S.normals.resize((size_t)np * nd);
    S.offsets.resize(np);
    S.areas.resize(np);
    std::vector<float> bias((size_t)np * nd), off_pts((size_t)np * nd);
    for (int64_t p = 0; p < np; ++p) {
      float x[3], acc = 0.f;
      for (int d = 0; d < nd; ++d) {
        x[d] = (float)fc[p * nd + d];
        S.points[p * nd + d] = x[d];
        float nv = (float)fn[p * nd + d];
        float sq = nv * nv;
        acc = d == 0 ? sq : acc + sq;
      }
      int64_t ci;
      float dd;
      bi.knn(x, 1, 1.5 * bi.local_width(x), &ci, &dd);
      float wci[3];
      bi.geo.width(ci, wci);
      float h = diam_of(wci, nd) * 1.01f;
      float A = std::sqrt(acc) + EPS32;
      S.offsets[p] = h;
      S.areas[p] = A;
      for (int d = 0; d < nd; ++d) {
        float nv = (float)fn[p * nd + d] / A;
        S.normals[p * nd + d] = nv;
        bias[p * nd + d] = nv * h;
        off_pts[p * nd + d] = x[d] + bias[p * nd + d] * ghost_layer_ratio;
      }
    }
    build_interp(bi, nd, np, S.points.data(), bias.data(), true, S.interp);
    build_interp(bi, nd, np, off_pts.data(), nullptr, true, S.offset_interp);
  }
  g_dom[D.get()] = D;
  *out = D.get();
  return IBX_OK;
  IBX_CATCH
}

int ibx_domain_build(const ibx_mesh* mh, int64_t max_partition_size, int skirt_depth, float ghost_layer_ratio, int nfam,
                     const char* const* fam_names, const int* fam_ptr, const int* fam_dim, const int* fam_front,
                     int build_partitions_flag, int build_surfaces, ibx_domain** out) {
  return domain_build_impl(mh, max_partition_size, skirt_depth, ghost_layer_ratio, nfam, fam_names, fam_ptr, fam_dim, fam_front,
                           build_partitions_flag, build_surfaces, 0, 1, out);
}

int ibx_domain_build_for_rank(const ibx_mesh* mh, float ghost_layer_ratio, int nfam, const char* const* fam_names,
                              const int* fam_ptr, const int* fam_dim, const int* fam_front, int rank, int nranks,
                              ibx_domain** out) {
  return domain_build_impl(mh, (int64_t)1 << 40, 2, ghost_layer_ratio, nfam, fam_names, fam_ptr, fam_dim, fam_front, 0, 0, rank,
                           nranks, out);
}

int ibx_domain_free(ibx_domain* d) {
  g_dom.erase(d);
  return IBX_OK;
}

#define DOM(d) auto it_ = g_dom.find(d); IBX_REQUIRE(it_ != g_dom.end(), "unknown ibx_domain handle"); ibx_domain& D = *it_->second

int ibx_domain_info(const ibx_domain* d, int* nd, int64_t* ncells, int64_t* nfaces, int* npartitions, int* nboundaries,
                    int* nsurfaces) {
  IBX_TRY
  DOM(d);
  *nd = D.nd;
  *ncells = D.ncells;
  *nfaces = (int64_t)D.faces.size() / 3;
  *npartitions = (int)D.parts.size();
  *nboundaries = (int)D.boundaries.size();
  *nsurfaces = (int)D.surfaces.size();
  return IBX_OK;
  IBX_CATCH
}

int ibx_domain_flags(const ibx_domain* d, int* two_to_one, int64_t* nblocks) {
  IBX_TRY
  DOM(d);
  *two_to_one = D.two_to_one;
  *nblocks = D.mesh->nblocks();
  return IBX_OK;
  IBX_CATCH
}

int ibx_domain_block_faces(const ibx_domain* d, int32_t* out /* nblocks x 2nd x 7 */) {
  IBX_TRY
  DOM(d);
  for (size_t i = 0; i < D.block_faces.size(); ++i) {
    const BlockFace& f = D.block_faces[i];
    int32_t* o = out + i * 7;
    o[0] = f.kind; o[1] = f.nb[0]; o[2] = f.nb[1]; o[3] = f.nb[2]; o[4] = f.nb[3]; o[5] = f.sub[0]; o[6] = f.sub[1];
  }
  return IBX_OK;
  IBX_CATCH
}

int ibx_domain_faces(const ibx_domain* d, int32_t* faces3) {
  IBX_TRY
  DOM(d);
  std::copy(D.faces.begin(), D.faces.end(), faces3);
  return IBX_OK;
  IBX_CATCH
}

int ibx_domain_cells(const ibx_domain* d, float* centers, float* widths) {
  IBX_TRY
  DOM(d);
  if (D.centers.empty() && D.ncells > 0) {   // rank-restricted global build: nothing stored, compute on request
    mesh_cells(*D.mesh, centers, widths);
    return IBX_OK;
  }
  if (centers) std::copy(D.centers.begin(), D.centers.end(), centers);
  if (widths) std::copy(D.widths.begin(), D.widths.end(), widths);
  return IBX_OK;
  IBX_CATCH
}

int ibx_partition_info(const ibx_domain* d, int p, int64_t* n_domain, int64_t* n_image, int64_t* image_start,
                       int64_t* nfaces_per_dim) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(p >= 0 && p < (int)D.parts.size(), "partition index out of range");
  const PartitionT& P = D.parts[p];
  *n_domain = (int64_t)P.domain.size();
  *n_image = P.n_image;
  *image_start = P.image_start;
  for (int k = 0; k < D.nd; ++k) nfaces_per_dim[k] = (int64_t)P.dims[k].owners.size();
  return IBX_OK;
  IBX_CATCH
}

int ibx_partition_tables(const ibx_domain* d, int p, int32_t* domain, int32_t* image_in_domain) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(p >= 0 && p < (int)D.parts.size(), "partition index out of range");
  const PartitionT& P = D.parts[p];
  std::copy(P.domain.begin(), P.domain.end(), domain);
  std::copy(P.image_in_domain.begin(), P.image_in_domain.end(), image_in_domain);
  return IBX_OK;
  IBX_CATCH
}

int ibx_partition_faces(const ibx_domain* d, int p, int dim, int32_t* owners, int32_t* neighbors) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(p >= 0 && p < (int)D.parts.size(), "partition index out of range");
  IBX_REQUIRE(dim >= 0 && dim < D.nd, "dim out of range");
  const FaceTable& T = D.parts[p].dims[dim];
  std::copy(T.owners.begin(), T.owners.end(), owners);
  std::copy(T.neighbors.begin(), T.neighbors.end(), neighbors);
  return IBX_OK;
  IBX_CATCH
}

int ibx_partition_face_lists(const ibx_domain* d, int p, int dim, int side, int32_t* ptr, int32_t* idx) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(p >= 0 && p < (int)D.parts.size(), "partition index out of range");
  IBX_REQUIRE(dim >= 0 && dim < D.nd, "dim out of range");
  const FaceTable& T = D.parts[p].dims[dim];
  const auto& P = side ? T.rptr : T.lptr;
  const auto& I = side ? T.ridx : T.lidx;
  std::copy(P.begin(), P.end(), ptr);
  if (idx) std::copy(I.begin(), I.end(), idx);
  return IBX_OK;
  IBX_CATCH
}

int ibx_boundary_name(const ibx_domain* d, int b, const char** name, int* nparts) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(b >= 0 && b < (int)D.boundaries.size(), "boundary index out of range");
  *name = D.boundaries[b].name.c_str();
  *nparts = (int)D.boundaries[b].parts.size();
  return IBX_OK;
  IBX_CATCH
}

int ibx_boundary_info(const ibx_domain* d, int b, int part, int64_t* nghost, int64_t* n_image_domain, int64_t* nnz) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(b >= 0 && b < (int)D.boundaries.size(), "boundary index out of range");
  IBX_REQUIRE(part >= 0 && part < (int)D.boundaries[b].parts.size(), "boundary partition out of range");
  const BoundaryT& B = D.boundaries[b].parts[part];
  *nghost = (int64_t)B.ghost.size();
  *n_image_domain = (int64_t)B.image_domain.size();
  *nnz = (int64_t)B.idx.size();
  return IBX_OK;
  IBX_CATCH
}

int ibx_boundary_tie_count(const ibx_domain* d, int b, int part, int64_t* n_tied) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(b >= 0 && b < (int)D.boundaries.size(), "boundary index out of range");
  IBX_REQUIRE(part >= 0 && part < (int)D.boundaries[b].parts.size(), "boundary partition out of range");
  *n_tied = D.boundaries[b].parts[part].n_tied;
  return IBX_OK;
  IBX_CATCH
}

int ibx_boundary_tables(const ibx_domain* d, int b, int part, int32_t* ghost_indices, float* projections, float* normals,
                        float* image_distances, float* ghost_distances, int32_t* image_domain, int32_t* interp_ptr,
                        int32_t* interp_idx, float* interp_w) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(b >= 0 && b < (int)D.boundaries.size(), "boundary index out of range");
  IBX_REQUIRE(part >= 0 && part < (int)D.boundaries[b].parts.size(), "boundary partition out of range");
  const BoundaryT& B = D.boundaries[b].parts[part];
  std::copy(B.ghost.begin(), B.ghost.end(), ghost_indices);
  std::copy(B.proj.begin(), B.proj.end(), projections);
  std::copy(B.normals.begin(), B.normals.end(), normals);
  std::copy(B.image_dist.begin(), B.image_dist.end(), image_distances);
  std::copy(B.ghost_dist.begin(), B.ghost_dist.end(), ghost_distances);
  std::copy(B.image_domain.begin(), B.image_domain.end(), image_domain);
  std::copy(B.ptr.begin(), B.ptr.end(), interp_ptr);
  std::copy(B.idx.begin(), B.idx.end(), interp_idx);
  std::copy(B.w.begin(), B.w.end(), interp_w);
  return IBX_OK;
  IBX_CATCH
}

int ibx_surface_name(const ibx_domain* d, int s, const char** name) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(s >= 0 && s < (int)D.surfaces.size(), "surface index out of range");
  *name = D.surfaces[s].name.c_str();
  return IBX_OK;
  IBX_CATCH
}

int ibx_surface_info(const ibx_domain* d, int s, int64_t* npoints, int64_t* nnz, int64_t* nnz_offset) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(s >= 0 && s < (int)D.surfaces.size(), "surface index out of range");
  *npoints = (int64_t)D.surfaces[s].offsets.size();
  *nnz = (int64_t)D.surfaces[s].interp.idx.size();
  *nnz_offset = (int64_t)D.surfaces[s].offset_interp.idx.size();
  return IBX_OK;
  IBX_CATCH
}

int ibx_surface_tables(const ibx_domain* d, int s, float* points, float* offsets, float* normals, float* areas,
                       int32_t* ptr, int32_t* idx, float* w, int32_t* optr, int32_t* oidx, float* ow) {
  IBX_TRY
  DOM(d);
  IBX_REQUIRE(s >= 0 && s < (int)D.surfaces.size(), "surface index out of range");
  const SurfaceT& S = D.surfaces[s];
  std::copy(S.points.begin(), S.points.end(), points);
  std::copy(S.offsets.begin(), S.offsets.end(), offsets);
  std::copy(S.normals.begin(), S.normals.end(), normals);
  std::copy(S.areas.begin(), S.areas.end(), areas);
  std::copy(S.interp.ptr.begin(), S.interp.ptr.end(), ptr);
  std::copy(S.interp.idx.begin(), S.interp.idx.end(), idx);
  std::copy(S.interp.w.begin(), S.interp.w.end(), w);
  std::copy(S.offset_interp.ptr.begin(), S.offset_interp.ptr.end(), optr);
  std::copy(S.offset_interp.idx.begin(), S.offset_interp.idx.end(), oidx);
  std::copy(S.offset_interp.w.begin(), S.offset_interp.w.end(), ow);
  return IBX_OK;
  IBX_CATCH
}

// ------------------------------------------------------------------------------------ generic accumulators
static ibx_accum* reg_acc(std::shared_ptr<ibx_accum> a) {
  g_acc[a.get()] = a;
  return a.get();
}

int ibx_interpolator_build(int nd, int64_t n, const float* X, int64_t q, const float* Xc, const float* bias, int linear,
                           int k, ibx_accum** out) {
  IBX_TRY
  // Interpolator (src/nninterp.jl:85-138) on an arbitrary cloud
  IBX_REQUIRE(nd == 2 || nd == 3, "nd must be 2 or 3");
  if (k == 0) k = 1 << nd;
  IBX_REQUIRE(k >= 1 && k <= 16 && k <= n, "k must be in 1..min(16, n)");
  KDTree tree;
  tree.build_f(nd, n, X);
  auto A = std::make_shared<ibx_accum>();
  std::vector<int64_t> sidx((size_t)q * k);
  std::vector<float> sw((size_t)q * k);
  std::vector<int32_t> cnt(q + 1, 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t p = 0; p < q; ++p) {
    double xq[3];
    for (int d = 0; d < nd; ++d) xq[d] = (double)(bias ? Xc[p * nd + d] + bias[p * nd + d] : Xc[p * nd + d]);
    int64_t idx[16];
    double dd[16];
    int found = tree.knn(xq, true, k, idx, dd);
    cnt[p + 1] = stencil_weights(CloudPoints{X, nd}, nd, idx, found, Xc + p * nd, linear != 0, &sidx[p * k], &sw[p * k]);
  }
  for (int64_t p = 0; p < q; ++p) cnt[p + 1] += cnt[p];
  A->n_out = q;
  A->ptr = cnt;
  A->idx.resize((size_t)cnt[q]);
  A->w.resize((size_t)cnt[q]);
  for (int64_t p = 0; p < q; ++p)
    for (int j = 0; j < cnt[p + 1] - cnt[p]; ++j) {
      A->idx[cnt[p] + j] = (int32_t)sidx[p * k + j];
      A->w[cnt[p] + j] = sw[p * k + j];
    }
  *out = reg_acc(A);
  return IBX_OK;
  IBX_CATCH
}

int ibx_mgrid_build(int nd, int64_t n, const float* X, int level, const float* volumes, ibx_accum** coarsener,
                    ibx_accum** prolongator) {
  IBX_TRY
  // coarsener_and_prolongator (src/mgrid.jl:24-97) without random permutation
  IBX_REQUIRE(nd == 2 || nd == 3, "nd must be 2 or 3");
  IBX_REQUIRE(level >= 1, "level must be >= 1");
  int64_t step = (int64_t)1 << (nd * level);
  std::vector<float> Xc;
  for (int64_t i = 0; i < n; i += step)
    for (int d = 0; d < nd; ++d) Xc.push_back(X[i * nd + d]);
  int64_t nc = (int64_t)Xc.size() / nd;
  KDTree tree;
  tree.build_f(nd, nc, Xc.data());
  std::vector<int32_t> owner(n);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double x[3];
    for (int d = 0; d < nd; ++d) x[d] = X[i * nd + d];
    int64_t j;
    double dd;
    tree.knn(x, true, 1, &j, &dd);
    owner[i] = (int32_t)j;
  }
  auto Cc = std::make_shared<ibx_accum>();
  auto Pp = std::make_shared<ibx_accum>();
  Cc->n_out = nc;
  Cc->ptr.assign(nc + 1, 0);
  for (int64_t i = 0; i < n; ++i) ++Cc->ptr[owner[i] + 1];
  for (int64_t c = 0; c < nc; ++c) Cc->ptr[c + 1] += Cc->ptr[c];
  Cc->idx.resize(n);
  Cc->w.resize(n);
  std::vector<int32_t> fill(Cc->ptr.begin(), Cc->ptr.end() - 1);
  for (int64_t i = 0; i < n; ++i) Cc->idx[fill[owner[i]]++] = (int32_t)i;
  for (int64_t c = 0; c < nc; ++c) {
    float s = 0.f;
    for (int32_t k = Cc->ptr[c]; k < Cc->ptr[c + 1]; ++k) s = s + (volumes ? volumes[Cc->idx[k]] : 1.0f);
    for (int32_t k = Cc->ptr[c]; k < Cc->ptr[c + 1]; ++k) Cc->w[k] = (volumes ? volumes[Cc->idx[k]] : 1.0f) / s;
  }
  Pp->n_out = n;
  Pp->weighted = false;
  Pp->ptr.resize(n + 1);
  std::iota(Pp->ptr.begin(), Pp->ptr.end(), 0);
  Pp->idx.assign(owner.begin(), owner.end());
  *coarsener = reg_acc(Cc);
  *prolongator = reg_acc(Pp);
  return IBX_OK;
  IBX_CATCH
}

int ibx_accum_create(int64_t n_out, const int32_t* ptr, const int32_t* idx, const float* w, ibx_accum** out) {
  IBX_TRY
  auto A = std::make_shared<ibx_accum>();
  A->n_out = n_out;
  A->ptr.assign(ptr, ptr + n_out + 1);
  IBX_REQUIRE(ptr[0] == 0, "ptr[0] must be 0");
  for (int64_t i = 0; i < n_out; ++i) IBX_REQUIRE(ptr[i + 1] >= ptr[i], "ptr must be non-decreasing");
  A->idx.assign(idx, idx + ptr[n_out]);
  A->weighted = w != nullptr;
  if (w) A->w.assign(w, w + ptr[n_out]);
  *out = reg_acc(A);
  return IBX_OK;
  IBX_CATCH
}

int ibx_accum_info(const ibx_accum* a, int64_t* n_out, int64_t* nnz, int* weighted) {
  IBX_TRY
  auto it = g_acc.find(a);
  IBX_REQUIRE(it != g_acc.end(), "unknown ibx_accum handle");
  *n_out = it->second->n_out;
  *nnz = (int64_t)it->second->idx.size();
  *weighted = it->second->weighted;
  return IBX_OK;
  IBX_CATCH
}

int ibx_accum_tables(const ibx_accum* a, int32_t* ptr, int32_t* idx, float* w) {
  IBX_TRY
  auto it = g_acc.find(a);
  IBX_REQUIRE(it != g_acc.end(), "unknown ibx_accum handle");
  const ibx_accum& A = *it->second;
  std::copy(A.ptr.begin(), A.ptr.end(), ptr);
  std::copy(A.idx.begin(), A.idx.end(), idx);
  if (w && A.weighted) std::copy(A.w.begin(), A.w.end(), w);
  return IBX_OK;
  IBX_CATCH
}

int ibx_accum_free(ibx_accum* a) {
  g_acc.erase(a);
  return IBX_OK;
}

}  // extern "C"

namespace ibx {
ibx_domain* find_domain(const ibx_domain* d) {
  auto it = g_dom.find(d);
  return it == g_dom.end() ? nullptr : it->second.get();
}
ibx_accum* find_accum(const ibx_accum* a) {
  auto it = g_acc.find(a);
  return it == g_acc.end() ? nullptr : it->second.get();
}
ibx_domain* register_domain(std::shared_ptr<ibx_domain> d) {
  g_dom[d.get()] = d;
  return d.get();
}
}  // namespace ibx
