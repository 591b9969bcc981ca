// Host geometry: exact nearest-neighbour search, Stereolitography handling, distance fields.
// Mirrors src/mesher.jl:124-801 of the reference (module BlockMesher); each function cites its lines.
#include "ibx_internal.h"

#include <cstdio>
#include <fstream>
#include <sstream>
#include <unordered_map>
#include <numeric>
#include <cstdlib>
#include <omp.h>

namespace ibx {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// ------------------------------------------------------------------------------------ KD-tree
static const int kLeaf = 12;

void KDTree::build(int nd_, int64_t n_, const double* p, bool f32_) {
  nd = nd_;
  n = n_;
  f32 = f32_;
  pts.assign(p, p + n * nd);
  perm.resize(n);
  std::iota(perm.begin(), perm.end(), (int64_t)0);
  // a node that splits has more than kLeaf points and halves, so a leaf holds at least (kLeaf + 1) / 2 of them: the
  // arrays are sized for the worst case up front and node ids are drawn from a counter, which lets the two halves of
  // a large node be built by different threads (ids depend on the schedule, the tree and every query result do not)
  const int64_t cap = 2 * (n / ((kLeaf + 1) / 2)) + 16;
  nodes.assign((size_t)cap, Node{-1, 0.0, 0, 0, -1, -1});
  bbox.assign((size_t)cap * 2 * nd, 0.0);
  n_nodes_ = 0;
  if (n > 0) {
#pragma omp parallel
#pragma omp single
    build_rec(0, n);
  }
  nodes.resize((size_t)n_nodes_);
  bbox.resize((size_t)n_nodes_ * 2 * nd);
}

void KDTree::build_f(int nd_, int64_t n_, const float* p) {
  std::vector<double> tmp((size_t)n_ * nd_);
  for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = p[i];
  build(nd_, n_, tmp.data(), true);
}

int64_t KDTree::build_rec(int64_t lo, int64_t hi) {
  int64_t id;
#pragma omp atomic capture
  id = n_nodes_++;
  nodes[(size_t)id] = Node{-1, 0.0, lo, hi, -1, -1};
  // bounding box of the node's points; split the widest dimension at the median
  int best = 0;
  double bw = -1;
  for (int d = 0; d < nd; ++d) {
    double mn = 1e300, mx = -1e300;
    for (int64_t i = lo; i < hi; ++i) {
      double v = pts[perm[i] * nd + d];
      mn = std::min(mn, v);
      mx = std::max(mx, v);
    }
    bbox[(size_t)id * 2 * nd + d] = mn;
    bbox[(size_t)id * 2 * nd + nd + d] = mx;
    if (mx - mn > bw) { bw = mx - mn; best = d; }
  }
  if (hi - lo <= kLeaf) return id;
  if (bw <= 0) return id;  // all points identical: keep as a (large) leaf
  int64_t mid = (lo + hi) / 2;
  std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi, [&](int64_t a, int64_t b) {
    double va = pts[a * nd + best], vb = pts[b * nd + best];
    return va < vb || (va == vb && a < b);
  });
  double split = pts[perm[mid] * nd + best];
  int64_t l, r;
  if (hi - lo > 16384) {
#pragma omp task shared(l)
    l = build_rec(lo, mid);
    r = build_rec(mid, hi);
#pragma omp taskwait
  } else {
    l = build_rec(lo, mid);
    r = build_rec(mid, hi);
  }
  nodes[id].dim = best;
  nodes[id].split = split;
  nodes[id].left = l;
  nodes[id].right = r;
  return id;
}

double KDTree::d2(int64_t i, const double* x, bool xf32) const {
  const double* p = &pts[i * nd];
  if (f32 && xf32) {
    float acc = 0.f;
    for (int d = 0; d < nd; ++d) {
      float df = (float)p[d] - (float)x[d];
      float sq = df * df;
      acc = d == 0 ? sq : acc + sq;
    }
    return (double)acc;
  }
  double acc = 0;
  for (int d = 0; d < nd; ++d) {
    double df = p[d] - x[d];
    double sq = df * df;
    acc = d == 0 ? sq : acc + sq;
  }
  return acc;
}

namespace {
struct KnnState {
  int k, found = 0;
  int64_t* idx;
  double* dd;
  bool better(double d, int64_t i) const {
    if (found < k) return d <= cap;
    return d < dd[k - 1] || (d == dd[k - 1] && i < idx[k - 1]);
  }
  void insert(double d, int64_t i) {
    int pos = found < k ? found : k - 1;
    while (pos > 0 && (d < dd[pos - 1] || (d == dd[pos - 1] && i < idx[pos - 1]))) {
      dd[pos] = dd[pos - 1];
      idx[pos] = idx[pos - 1];
      --pos;
    }
    dd[pos] = d;
    idx[pos] = i;
    if (found < k) ++found;
  }
  double cap = 1e300;   // points beyond this squared distance are not wanted
  double worst() const { return found < k ? cap : dd[k - 1]; }
};
}  // namespace

int KDTree::knn(const double* x, bool xf32, int k, int64_t* idx, double* d2out, double max_d2) const {
  KnnState st{k, 0, idx, d2out};
  st.cap = max_d2;
  if (n == 0) return 0;
  // explicit stack of (node, lower bound on squared distance)
  struct Item { int64_t node; double bound; };
  std::vector<Item> stack;
  stack.reserve(64);
  stack.push_back({0, 0.0});
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    if (it.bound > st.worst() * (1 + 1e-5) + 1e-300) continue;
    const Node& nd_ = nodes[it.node];
    if (nd_.dim < 0) {
      for (int64_t i = nd_.lo; i < nd_.hi; ++i) {
        int64_t p = perm[i];
        double d = d2(p, x, xf32);
        if (st.better(d, p)) st.insert(d, p);
      }
      continue;
    }
    double diff = x[nd_.dim] - nd_.split;
    int64_t nearc = diff < 0 ? nd_.left : nd_.right;
    int64_t farc = diff < 0 ? nd_.right : nd_.left;
    // lower bounds from the children's bounding boxes: the split-plane distance alone prunes nothing when the query lies
    // far outside the cloud (coarse cells around a finely refined surface)
    stack.push_back({farc, box_d2(farc, x)});
    stack.push_back({nearc, box_d2(nearc, x)});
  }
  return st.found;
}

void KDTree::inrange(const double* x, bool xf32, double r, std::vector<int64_t>& out) const {
  out.clear();
  if (n == 0) return;
  double r2 = (f32 && xf32) ? (double)((float)r * (float)r) : r * r;
  struct Item { int64_t node; double bound; };
  std::vector<Item> stack;
  stack.push_back({0, 0.0});
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    if (it.bound > r2 * (1 + 1e-5) + 1e-300) continue;
    const Node& nd_ = nodes[it.node];
    if (nd_.dim < 0) {
      for (int64_t i = nd_.lo; i < nd_.hi; ++i) {
        int64_t p = perm[i];
        if (d2(p, x, xf32) <= r2) out.push_back(p);
      }
      continue;
    }
    stack.push_back({nd_.right, box_d2(nd_.right, x)});
    stack.push_back({nd_.left, box_d2(nd_.left, x)});
  }
  std::sort(out.begin(), out.end());
}

// ------------------------------------------------------------------------------------ small linear algebra
// Moore-Penrose pseudo-inverse of an m x n (m >= n or m < n, both tiny) row-major matrix through a one-sided
// Jacobi SVD in double; singular values <= rtol * smax are dropped (Julia: rtol = eps(T) * min(m, n)).
void pinv_small(const double* A, int m, int n, double rtol, double* out /* n x m */) {
  // work on columns of A (m x n): orthogonalise columns -> A V = U S
  const int MAXD = 16;
  double U[MAXD * MAXD], V[MAXD * MAXD];
  bool transposed = false;
  int mm = m, nn = n;
  std::vector<double> At;
  const double* src = A;
  if (m < n) {  // use the transpose so that rows >= cols
    At.resize((size_t)m * n);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) At[(size_t)j * m + i] = A[(size_t)i * n + j];
    src = At.data();
    mm = n;
    nn = m;
    transposed = true;
  }
  for (int i = 0; i < mm; ++i)
    for (int j = 0; j < nn; ++j) U[i * nn + j] = src[(size_t)i * nn + j];
  for (int i = 0; i < nn; ++i)
    for (int j = 0; j < nn; ++j) V[i * nn + j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < nn - 1; ++p)
      for (int q = p + 1; q < nn; ++q) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < mm; ++i) {
          a += U[i * nn + p] * U[i * nn + p];
          b += U[i * nn + q] * U[i * nn + q];
          c += U[i * nn + p] * U[i * nn + q];
        }
        if (c == 0) continue;
        off = std::max(off, std::fabs(c) / std::sqrt(std::max(a * b, 1e-300)));
        double zeta = (b - a) / (2 * c);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
        for (int i = 0; i < mm; ++i) {
          double up = U[i * nn + p], uq = U[i * nn + q];
          U[i * nn + p] = cs * up - sn * uq;
          U[i * nn + q] = sn * up + cs * uq;
        }
        for (int i = 0; i < nn; ++i) {
          double vp = V[i * nn + p], vq = V[i * nn + q];
          V[i * nn + p] = cs * vp - sn * vq;
          V[i * nn + q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-15) break;
  }
  double s[MAXD], smax = 0;
  for (int j = 0; j < nn; ++j) {
    double a = 0;
    for (int i = 0; i < mm; ++i) a += U[i * nn + j] * U[i * nn + j];
    s[j] = std::sqrt(a);
    smax = std::max(smax, s[j]);
  }
  // pinv(src) = V S^-1 U^T  (nn x mm)
  std::vector<double> P((size_t)nn * mm, 0.0);
  for (int j = 0; j < nn; ++j) {
    if (!(s[j] > rtol * smax) || s[j] == 0) continue;
    double inv = 1.0 / (s[j] * s[j]);  // U column is unnormalised: u_j = U[:,j]/s_j
    for (int r = 0; r < nn; ++r)
      for (int i = 0; i < mm; ++i) P[(size_t)r * mm + i] += V[r * nn + j] * inv * U[i * nn + j];
  }
  if (!transposed) {
    for (size_t i = 0; i < P.size(); ++i) out[i] = P[i];
  } else {  // pinv(A) = pinv(A^T)^T ; P is m x n -> out n x m
    for (int r = 0; r < nn; ++r)
      for (int i = 0; i < mm; ++i) out[(size_t)i * nn + r] = P[(size_t)r * mm + i];
  }
}

template <class T>
static T normT(const T* v, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)v[i] * (double)v[i];
  return (T)std::sqrt(s);
}

// ------------------------------------------------------------------------------------ region distances
// Box/Ball/Line (src/mesher.jl:39-46, :73-76, :112-122) evaluate in Float64 (their parameters are Float64
// vectors in every reference script); DistanceField (:767-769) in the promoted type of tree and query.
Num region_distance(const ibx_region& r, int nd, const double* x, bool xf32) {
  switch (r.kind) {
    case 0: {
      double s = 0;
      for (int d = 0; d < nd; ++d) s += (r.c[d] - x[d]) * (r.c[d] - x[d]);
      return {std::max(0.0, std::sqrt(s) - r.a[0]), false};
    }
    case 1: {
      double s = 0;
      for (int d = 0; d < nd; ++d) {
        double dd = x[d] - r.c[d];
        bool outside = (dd > r.a[d]) || (x[d] < r.c[d]);
        double v = std::min(std::fabs(dd), std::fabs(dd - r.a[d])) * (outside ? 1.0 : 0.0);
        s += v * v;
      }
      return {std::sqrt(s), false};
    }
    case 2: {
      double m[3], ss = 0, xi = 0;
      for (int d = 0; d < nd; ++d) { m[d] = r.a[d] - r.c[d]; ss += m[d] * m[d]; }
      for (int d = 0; d < nd; ++d) xi += (m[d] / ss) * (x[d] - r.c[d]);  // pinv(m) * v
      double s = 0;
      for (int d = 0; d < nd; ++d) {
        double q = xi < 0.0 ? r.c[d] : (xi > 1.0 ? r.a[d] : r.c[d] + m[d] * xi);
        s += (x[d] - q) * (x[d] - q);
      }
      return {std::sqrt(s), false};
    }
    case 3:
      return r.dfield->distance(x, xf32);
  }
  throw std::runtime_error("unknown region kind");
}

// proj2simplex (src/mesher.jl:544-596); arithmetic in T (the promoted type of simplex and point)
template <class T>
static void proj2simplex(const T* simp, int nv, int nd, const T* pt, T* out) {
  const T eps = (T)1e-14f;
  if (nv == 1) {
    for (int d = 0; d < nd; ++d) out[d] = simp[d];
    return;
  }
  if (nv == 2) {
    const T *p0 = simp, *p1 = simp + nd;
    T u[3], num = 0, den = 0;
    for (int d = 0; d < nd; ++d) {
      u[d] = p1[d] - p0[d];
      T a = (pt[d] - p0[d]) * u[d];
      T b = u[d] * u[d];
      num = d == 0 ? a : num + a;
      den = d == 0 ? b : den + b;
    }
    T xi = num / (den + eps);
    if (xi < -eps) {
      for (int d = 0; d < nd; ++d) out[d] = p0[d];
    } else if (xi > (T)1.0 + eps) {
      for (int d = 0; d < nd; ++d) out[d] = p1[d];
    } else {
      for (int d = 0; d < nd; ++d) out[d] = p0[d] + u[d] * xi;
    }
    return;
  }
  // triangle in 3-D: xi = pinv(M) * (pt - p0), M = [p1 - p0, p2 - p0]
  // Canonical rule: the reference takes pinv through LAPACK's SVD (LinearAlgebra.pinv, not vendored,
  // not bit-reproducible).  For a full-rank nd x 2 matrix pinv(M) = (M^T M)^-1 M^T; it is evaluated here in Float64 from
  // the T-rounded entries of M -- Gram matrix summed in dimension order, closed-form 2 x 2 inverse -- and rounded to T
  // once.  The oracle performs the same operations in the same order, so projections (and therefore image points
  // and donor sets) agree bit for bit.  (Nearly) rank-deficient triangles fall back to the Jacobi SVD with Julia's cut-off.
  const T* p0 = simp;
  double M[3 * 2], P[2 * 3];
  for (int d = 0; d < nd; ++d) {
    M[d * 2 + 0] = (double)(T)(simp[nd + d] - p0[d]);
    M[d * 2 + 1] = (double)(T)(simp[2 * nd + d] - p0[d]);
  }
  double rtol = (sizeof(T) == 4 ? 1.1920928955078125e-07 : 2.220446049250313e-16) * 2;
  {
    double ga = 0, gb = 0, gc = 0;
    for (int d = 0; d < nd; ++d) {
      const double m0 = M[d * 2 + 0], m1 = M[d * 2 + 1];
      ga = d == 0 ? m0 * m0 : ga + m0 * m0;
      gb = d == 0 ? m0 * m1 : gb + m0 * m1;
      gc = d == 0 ? m1 * m1 : gc + m1 * m1;
    }
    const double det = ga * gc - gb * gb;
    if (det > 1e-10 * (ga * gc)) {
      for (int d = 0; d < nd; ++d) {
        P[0 * nd + d] = (gc * M[d * 2 + 0] - gb * M[d * 2 + 1]) / det;
        P[1 * nd + d] = (ga * M[d * 2 + 1] - gb * M[d * 2 + 0]) / det;
      }
    } else {
      pinv_small(M, nd, 2, rtol, P);
    }
  }
  T xi[2];
  for (int j = 0; j < 2; ++j) {
    T acc = 0;
    for (int d = 0; d < nd; ++d) {
      T t = (T)P[j * nd + d] * (pt[d] - p0[d]);
      acc = d == 0 ? t : acc + t;
    }
    xi[j] = acc;
  }
  if (xi[0] < -eps || xi[1] < -eps || xi[0] + xi[1] > (T)1.0 + eps) {
    T best[3] = {0, 0, 0};
    T bd = INFINITY;
    for (int drop = 0; drop < 3; ++drop) {  // simplex_faces (:533-539)
      T face[6], pr[3], df[3];
      int c = 0;
      for (int j = 0; j < 3; ++j)
        if (j != drop) {
          for (int d = 0; d < nd; ++d) face[c * nd + d] = simp[j * nd + d];
          ++c;
        }
      proj2simplex<T>(face, 2, nd, pt, pr);
      for (int d = 0; d < nd; ++d) df[d] = pr[d] - pt[d];
      T dd = normT<T>(df, nd);
      if (dd < bd) {
        bd = dd;
        for (int d = 0; d < nd; ++d) best[d] = pr[d];
      }
    }
    for (int d = 0; d < nd; ++d) out[d] = best[d];
    return;
  }
  for (int d = 0; d < nd; ++d) {
    T t0 = (T)M[d * 2 + 0] * xi[0];
    T t1 = (T)M[d * 2 + 1] * xi[1];
    out[d] = p0[d] + (t0 + t1);
  }
}

}  // namespace ibx

using namespace ibx;

Num ibx_dfield::distance(const double* x, bool xf32) const {
  if (sphere) {
    double s = 0;
    int nd = 3;
    for (int d = 0; d < nd; ++d) s += (x[d] - sc[d]) * (x[d] - sc[d]);
    return {std::fabs(std::sqrt(s) - sr), false};
  }
  int64_t i;
  double dd;
  tree.knn(x, xf32, 1, &i, &dd);
  bool f = tree.f32 && xf32;
  return {f ? (double)std::sqrt((float)dd) : std::sqrt(dd), f};
}

bool ibx_dfield::distance_within(const double* x, bool xf32, double r, Num* out) const {
  if (sphere) {
    *out = distance(x, xf32);
    return out->v <= r;
  }
  // a point farther than r from the bounding box of the simplex centres is farther than r from every centre
  const double r2 = r * r * (1.0 + 1e-5) + 1e-300;
  if (tree.n == 0 || tree.box_d2(0, x) > r2) return false;
  int64_t i;
  double dd;
  if (tree.knn(x, xf32, 1, &i, &dd, r2) == 0) return false;
  bool f = tree.f32 && xf32;
  *out = {f ? (double)std::sqrt((float)dd) : std::sqrt(dd), f};
  return out->v <= r;
}

void ibx_dfield::projection(const double* x, bool xf32, double R, double* out) const {
  if (sphere) {
    double v[3], s = 0;
    for (int d = 0; d < 3; ++d) { v[d] = x[d] - sc[d]; s += v[d] * v[d]; }
    s = std::sqrt(s);
    for (int d = 0; d < 3; ++d) out[d] = s > 0 ? sc[d] + v[d] * (sr / s) : sc[d] + (d == 0 ? sr : 0.0);
    return;
  }
  int nd = stl->nd;
  bool f = tree.f32 && xf32;
  int64_t i0;
  double d2v;
  tree.knn(x, xf32, 1, &i0, &d2v);
  double d = f ? (double)std::sqrt((float)d2v) : std::sqrt(d2v);
  for (int k = 0; k < nd; ++k) out[k] = centers[i0 * nd + k];
  if (!(R > d)) return;
  // Candidates: every simplex whose centre is within R (src/mesher.jl:785), scanned in ascending index order with a strict
  // `<` -- i.e. the winner is the LOWEST-INDEX simplex among those at the minimal distance d_min, if d_min beats d (the
  // distance to the nearest centre), else the centre stays.  All points of a simplex lie within rmax of its centre, so a
  // simplex whose centre is farther than best + rmax cannot reach the current best: candidates are visited by increasing
  // centre distance and the scan stops there.  Every simplex at d_min has been evaluated by then (its centre is within
  // d_min + rmax), so picking the lowest index among the evaluated minima reproduces the full scan bit for bit -- at a
  // few evaluations per query instead of hundreds (C4 with an STL sphere: 640 s -> seconds at 3 M cells).
  const double margin = 1.0 + 1e-6;
  const double Rp = std::min(R, (d + rmax) * margin + 1e-30);
  std::vector<int64_t> cand;
  tree.inrange(x, xf32, Rp, cand);
  std::vector<std::pair<double, int64_t>> byd(cand.size());
  for (size_t q = 0; q < cand.size(); ++q) {
    double a = 0;
    for (int k = 0; k < nd; ++k) { double t = centers[cand[q] * nd + k] - x[k]; a += t * t; }
    byd[q] = {std::sqrt(a), cand[q]};
  }
  // few candidates (a finely triangulated surface: nearly all of them end up evaluated): ascending index order, no sort;
  // the `dd == dmin && s < imin` rule below makes the winner independent of the visiting order either way
  const bool sorted = cand.size() > 192;
  if (sorted) std::sort(byd.begin(), byd.end());
  double dmin = INFINITY, pmin[3] = {0, 0, 0};
  int64_t imin = -1;
  for (const auto& c : byd) {
    if (c.first > (std::min(dmin, d) + rmax) * margin + 1e-30) {
      if (sorted) break;
      continue;
    }
    const int64_t s = c.second;
    double pr[3], dd;
    if (f) {
      float simp[9], pt[3], o[3], df[3];
      for (int v = 0; v < nd; ++v)
        for (int k = 0; k < nd; ++k) simp[v * nd + k] = (float)stl->points[stl->simplices[s * nd + v] * nd + k];
      for (int k = 0; k < nd; ++k) pt[k] = (float)x[k];
      proj2simplex<float>(simp, nd, nd, pt, o);
      for (int k = 0; k < nd; ++k) { df[k] = o[k] - pt[k]; pr[k] = o[k]; }
      dd = (double)normT<float>(df, nd);
    } else {
      double simp[9], o[3], df[3];
      for (int v = 0; v < nd; ++v)
        for (int k = 0; k < nd; ++k) simp[v * nd + k] = stl->points[stl->simplices[s * nd + v] * nd + k];
      proj2simplex<double>(simp, nd, nd, x, o);
      for (int k = 0; k < nd; ++k) { df[k] = o[k] - x[k]; pr[k] = o[k]; }
      dd = normT<double>(df, nd);
    }
    if (dd < dmin || (dd == dmin && s < imin)) {
      dmin = dd;
      imin = s;
      for (int k = 0; k < nd; ++k) pmin[k] = pr[k];
    }
  }
  if (imin >= 0 && dmin < d)
    for (int k = 0; k < nd; ++k) out[k] = pmin[k];
}

// ------------------------------------------------------------------------------------ STL operations
static void simplex_centers_normals(const ibx_stl& s, std::vector<double>& centers, std::vector<double>& normals) {
  // centers_and_normals (src/mesher.jl:639-660), _simplex_normal(normalize = false) (:601-628)
  int nd = s.nd;
  int64_t ns = s.nsimp();
  centers.assign((size_t)ns * nd, 0.0);
  normals.assign((size_t)ns * nd, 0.0);
  auto run = [&](auto tag) {
    using T = decltype(tag);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < ns; ++i) {
      T p[9];
      for (int v = 0; v < nd; ++v)
        for (int d = 0; d < nd; ++d) p[v * nd + d] = (T)s.points[s.simplices[i * nd + v] * nd + d];
      for (int d = 0; d < nd; ++d) {
        T acc = p[d];
        for (int v = 1; v < nd; ++v) acc = acc + p[v * nd + d];
        centers[i * nd + d] = (double)(T)(acc / (T)nd);
      }
      if (nd == 2) {
        T vx = p[2] - p[0], vy = p[3] - p[1];
        normals[i * 2 + 0] = (double)vy;
        normals[i * 2 + 1] = (double)(-vx);
      } else {
        T a[3], b[3];
        for (int d = 0; d < 3; ++d) { a[d] = p[3 + d] - p[d]; b[d] = p[6 + d] - p[d]; }
        normals[i * 3 + 0] = (double)(T)(a[1] * b[2] - a[2] * b[1]);
        normals[i * 3 + 1] = (double)(T)(a[2] * b[0] - a[0] * b[2]);
        normals[i * 3 + 2] = (double)(T)(a[0] * b[1] - a[1] * b[0]);
      }
    }
  };
  if (s.f32) run(float{}); else run(double{});
}

namespace ibx {
void stl_centers_normals(const ibx_stl& s, std::vector<double>& c, std::vector<double>& n) {
  simplex_centers_normals(s, c, n);
}
}  // namespace ibx

static std::shared_ptr<ibx_stl> merge_points_impl(const std::vector<const ibx_stl*>& in, double tol, bool tol_f32,
                                                  bool clean) {
  // merge_points (src/mesher.jl:351-407): tag = Int64(round(pt / tolerance)), first point with a tag is kept
  auto out = std::make_shared<ibx_stl>();
  out->nd = in[0]->nd;
  out->f32 = in[0]->f32;
  int nd = out->nd;
  struct Key {
    int64_t v[3];
    bool operator==(const Key& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      uint64_t h = 1469598103934665603ull;
      for (int i = 0; i < 3; ++i) { h ^= (uint64_t)k.v[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
      return (size_t)h;
    }
  };
  // Points are numbered globally over the inputs; the representative of a point is the FIRST point carrying its tag,
  // and kept points are numbered in order of first appearance -- the result of the reference's sequential loop.  The
  // tags are computed in parallel, the points split into buckets by tag hash (lists in ascending point order), and
  // every bucket resolved by its own hash map in parallel.
  std::vector<int64_t> off(in.size() + 1, 0);
  for (size_t k = 0; k < in.size(); ++k) off[k + 1] = off[k] + in[k]->npoints();
  const int64_t NP = off.back();
  std::vector<Key> keys((size_t)NP);
  constexpr int NBUCKET = 1024;
  std::vector<uint16_t> bucket((size_t)NP);
  for (size_t k = 0; k < in.size(); ++k) {
    const ibx_stl* s = in[k];
    const int64_t np = s->npoints();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < np; ++i) {
      Key key{{0, 0, 0}};
      for (int d = 0; d < nd; ++d) {
        double q = (s->f32 && tol_f32) ? (double)((float)s->points[i * nd + d] / (float)tol) : s->points[i * nd + d] / tol;
        key.v[d] = (int64_t)std::nearbyint(q);
      }
      keys[(size_t)(off[k] + i)] = key;
      bucket[(size_t)(off[k] + i)] = (uint16_t)(KeyHash()(key) % NBUCKET);
    }
  }
  std::vector<int64_t> bptr(NBUCKET + 1, 0);
  for (int64_t i = 0; i < NP; ++i) ++bptr[bucket[(size_t)i] + 1];
  for (int q = 0; q < NBUCKET; ++q) bptr[q + 1] += bptr[q];
  std::vector<int64_t> blist((size_t)NP);
  {
    std::vector<int64_t> fill(bptr.begin(), bptr.end() - 1);
    for (int64_t i = 0; i < NP; ++i) blist[(size_t)fill[bucket[(size_t)i]]++] = i;
  }
  std::vector<int64_t> rep((size_t)NP);
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = 0; q < NBUCKET; ++q) {
    std::unordered_map<Key, int64_t, KeyHash> first;
    first.reserve((size_t)(bptr[q + 1] - bptr[q]));
    for (int64_t j = bptr[q]; j < bptr[q + 1]; ++j) {
      const int64_t i = blist[(size_t)j];
      rep[(size_t)i] = first.emplace(keys[(size_t)i], i).first->second;
    }
  }
  std::vector<int64_t> newid((size_t)NP);   // valid at representatives
  int64_t nkept = 0;
  for (int64_t i = 0; i < NP; ++i)
    if (rep[(size_t)i] == i) newid[(size_t)i] = nkept++;
  out->points.resize((size_t)nkept * nd);
  std::vector<int64_t> simp;
  for (size_t k = 0; k < in.size(); ++k) {
    const ibx_stl* s = in[k];
    const int64_t np = s->npoints();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < np; ++i)
      if (rep[(size_t)(off[k] + i)] == off[k] + i)
        for (int d = 0; d < nd; ++d) out->points[(size_t)(newid[(size_t)(off[k] + i)] * nd + d)] = s->points[i * nd + d];
    const size_t base = simp.size();
    simp.resize(base + s->simplices.size());
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)s->simplices.size(); ++j)
      simp[base + (size_t)j] = newid[(size_t)rep[(size_t)(off[k] + s->simplices[(size_t)j])]];
  }
  int64_t ns = (int64_t)simp.size() / nd;
  for (int64_t i = 0; i < ns; ++i) {
    bool ok = true;
    if (clean)
      for (int a = 0; a < nd && ok; ++a)
        for (int b = a + 1; b < nd; ++b)
          if (simp[i * nd + a] == simp[i * nd + b]) ok = false;
    if (ok)
      for (int d = 0; d < nd; ++d) out->simplices.push_back(simp[i * nd + d]);
  }
  return out;
}

template <class T>
static void refine_simplex(std::vector<T> simplex, int nd, Num h, double gm1, int nreg, const ibx_region* regs,
                           std::vector<T>& out) {
  // refine_to_length! (src/mesher.jl:438-495), depth-first [first; second]
  int nv = nd;
  std::vector<std::vector<T>> stack;
  stack.push_back(std::move(simplex));
  while (!stack.empty()) {
    std::vector<T> s = std::move(stack.back());
    stack.pop_back();
    double max_violation = 0.0;
    int index = -1;
    for (int i = 0; i < nv; ++i) {
      int inext = (i == nv - 1) ? 0 : i + 1;
      T ph[3], df[3];
      double phd[3];
      for (int d = 0; d < nd; ++d) {
        ph[d] = (s[i * nd + d] + s[inext * nd + d]) / (T)2;
        df[d] = s[inext * nd + d] - s[i * nd + d];
        phd[d] = (double)ph[d];
      }
      Num L{(double)normT<T>(df, nd), sizeof(T) == 4};
      Num hloc = h;
      for (int r = 0; r < nreg; ++r) {
        Num dist = region_distance(regs[r], nd, phd, sizeof(T) == 4);
        Num cand = nmax(nmul(nsub(dist, L), Num{gm1, false}), Num{regs[r].h, regs[r].h_is_f32 != 0});
        hloc = nmin(hloc, cand);
      }
      double violation = nsub(L, hloc).v;
      if (max_violation < violation) {
        max_violation = violation;
        index = i;
      }
    }
    if (index < 0) {
      out.insert(out.end(), s.begin(), s.end());
      continue;
    }
    int inext = (index == nv - 1) ? 0 : index + 1;
    std::vector<T> second = s;
    for (int d = 0; d < nd; ++d) {
      T pn = (s[index * nd + d] + s[inext * nd + d]) / (T)2;
      s[inext * nd + d] = pn;
      second[index * nd + d] = pn;
    }
    stack.push_back(std::move(second));
    stack.push_back(std::move(s));
  }
}

namespace ibx {
std::shared_ptr<ibx_stl> refine_to_length_impl(const ibx_stl& s, Num h, double tol, bool tol_f32, double growth_ratio,
                                               int nreg, const ibx_region* regs) {
  // refine_to_length (src/mesher.jl:503-528)
  int nd = s.nd;
  ibx_stl tmp;
  tmp.nd = nd;
  tmp.f32 = s.f32;
  double gm1 = growth_ratio - 1.0;
  auto run = [&](auto tag) {
    using T = decltype(tag);
    // every input simplex refines on its own: chunks in parallel, outputs concatenated in input order
    const int64_t ns = s.nsimp(), CH = 64, nch = (ns + CH - 1) / CH;
    std::vector<std::vector<T>> outs((size_t)nch);
    std::string err;   // an exception must not leave the parallel region
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < nch; ++c) {
      std::vector<T>& out = outs[(size_t)c];
      try {
        for (int64_t i = c * CH; i < std::min(ns, (c + 1) * CH); ++i) {
          std::vector<T> simp((size_t)nd * nd);
          for (int v = 0; v < nd; ++v)
            for (int d = 0; d < nd; ++d) simp[v * nd + d] = (T)s.points[s.simplices[i * nd + v] * nd + d];
          refine_simplex<T>(std::move(simp), nd, h, gm1, nreg, regs, out);
        }
      } catch (const std::exception& e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    size_t total = 0;
    for (auto& o : outs) total += o.size();
    tmp.points.reserve(total);
    for (auto& o : outs) tmp.points.insert(tmp.points.end(), o.begin(), o.end());
  };
  double t0 = omp_get_wtime();
  if (s.f32) run(float{}); else run(double{});
  if (getenv("IBX_BUILD_VERBOSE")) fprintf(stderr, "[ibx mesh ]   refine loop %.2f s, %lld points\n", omp_get_wtime() - t0, (long long)tmp.npoints());
  int64_t np = tmp.npoints();
  tmp.simplices.resize(np);
  std::iota(tmp.simplices.begin(), tmp.simplices.end(), (int64_t)0);
  return merge_points_impl({&tmp}, tol, tol_f32, true);
}

std::shared_ptr<ibx_dfield> make_dfield(std::shared_ptr<ibx_stl> stl) {
  auto d = std::make_shared<ibx_dfield>();
  d->stl = stl;
  std::vector<double> normals;
  simplex_centers_normals(*stl, d->centers, normals);
  d->tree.build(stl->nd, stl->nsimp(), d->centers.data(), stl->f32);
  {
    const int nd = stl->nd;
    double r2 = 0;
#pragma omp parallel for schedule(static) reduction(max : r2)
    for (int64_t i = 0; i < stl->nsimp(); ++i)
      for (int v = 0; v < nd; ++v) {
        double a = 0;
        for (int k = 0; k < nd; ++k) {
          double t = stl->points[stl->simplices[i * nd + v] * nd + k] - d->centers[i * nd + k];
          a += t * t;
        }
        r2 = std::max(r2, a);
      }
    d->rmax = std::sqrt(r2);
  }
  return d;
}
}  // namespace ibx

// ------------------------------------------------------------------------------------ file readers
static std::shared_ptr<ibx_stl> read_surface(const std::string& path) {
  // Stereolitography(fname) (src/mesher.jl:279-296), STLReader (:124-227)
  auto out = std::make_shared<ibx_stl>();
  out->f32 = true;
  auto ends_with = [&](const char* suf) {
    size_t n = strlen(suf);
    return path.size() >= n && path.compare(path.size() - n, n, suf) == 0;
  };
  if (ends_with(".dat") || ends_with(".DAT")) {
    std::ifstream fh(path);
    if (!fh) throw std::runtime_error("cannot open " + path);
    out->nd = 2;
    double x, y;
    while (fh >> x >> y) {
      out->points.push_back((double)(float)x);
      out->points.push_back((double)(float)y);
    }
    int64_t n = out->npoints();
    for (int64_t i = 0; i < n; ++i) {
      out->simplices.push_back(i);
      out->simplices.push_back((i + 1) % n);
    }
    return out;
  }
  std::ifstream fh(path, std::ios::binary);
  if (!fh) throw std::runtime_error("cannot open " + path);
  std::string raw((std::istreambuf_iterator<char>(fh)), std::istreambuf_iterator<char>());
  out->nd = 3;
  if (raw.size() >= 5 && raw.compare(0, 5, "solid") == 0) {
    std::istringstream ss(raw);
    std::string line;
    std::vector<int64_t> face;
    while (std::getline(ss, line)) {
      size_t b = line.find_first_not_of(" \t\r");
      if (b == std::string::npos) continue;
      line = line.substr(b);
      if (line.compare(0, 6, "vertex") == 0) {
        std::istringstream ls(line.substr(6));
        float v[3];
        ls >> v[0] >> v[1] >> v[2];
        face.push_back(out->npoints());
        for (int d = 0; d < 3; ++d) out->points.push_back((double)v[d]);
      } else if (line.compare(0, 12, "facet normal") == 0) {
        face.clear();
      } else if (line.compare(0, 7, "endloop") == 0) {
        if (face.size() == 3)
          for (int64_t v : face) out->simplices.push_back(v);
      }
    }
    return out;
  }
  if (raw.size() < 84) throw std::runtime_error("binary STL too short: " + path);
  uint32_t ntri;
  memcpy(&ntri, raw.data() + 80, 4);
  if (raw.size() < 84 + (size_t)ntri * 50) throw std::runtime_error("binary STL truncated: " + path);
  for (uint32_t k = 0; k < ntri; ++k) {
    const char* rec = raw.data() + 84 + (size_t)k * 50 + 12;
    for (int v = 0; v < 3; ++v) {
      float p[3];
      memcpy(p, rec + v * 12, 12);
      out->simplices.push_back(out->npoints());
      for (int d = 0; d < 3; ++d) out->points.push_back((double)p[d]);
    }
  }
  return out;
}

// ------------------------------------------------------------------------------------ C ABI
// Handles own a shared_ptr so that distance fields / meshes can share STLs safely.
struct StlBox { std::shared_ptr<ibx_stl> p; };
static std::map<const ibx_stl*, std::shared_ptr<ibx_stl>> g_stl;
static std::map<const ibx_dfield*, std::shared_ptr<ibx_dfield>> g_df;

namespace ibx {
std::shared_ptr<ibx_stl> lookup_stl(const ibx_stl* s) {
  auto it = g_stl.find(s);
  if (it == g_stl.end()) throw std::runtime_error("unknown ibx_stl handle");
  return it->second;
}
std::shared_ptr<ibx_dfield> lookup_dfield(const ibx_dfield* d) {
  auto it = g_df.find(d);
  if (it == g_df.end()) throw std::runtime_error("unknown ibx_dfield handle");
  return it->second;
}
ibx_dfield* register_dfield(std::shared_ptr<ibx_dfield> d) {
  g_df[d.get()] = d;
  return d.get();
}
}  // namespace ibx

static ibx_stl* reg_stl(std::shared_ptr<ibx_stl> s) {
  g_stl[s.get()] = s;
  return s.get();
}

extern "C" {

const char* ibx_last_error(void) { return ibx::g_err.c_str(); }
const char* ibx_version(void) { return "ibx-b200 0.1.0 (sm_100a)"; }

int ibx_stl_create(int nd, int64_t npoints, const double* points, int64_t nsimp, const int64_t* simplices, int is_f32,
                   ibx_stl** out) {
  IBX_TRY
  IBX_REQUIRE(nd == 2 || nd == 3, "nd must be 2 or 3");
  auto s = std::make_shared<ibx_stl>();
  s->nd = nd;
  s->f32 = is_f32 != 0;
  s->points.assign(points, points + npoints * nd);
  if (s->f32)
    for (auto& v : s->points) v = (double)(float)v;
  s->simplices.assign(simplices, simplices + nsimp * nd);
  for (int64_t v : s->simplices) IBX_REQUIRE(v >= 0 && v < npoints, "simplex index out of range");
  *out = reg_stl(s);
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_read(const char* path, ibx_stl** out) {
  IBX_TRY
  *out = reg_stl(read_surface(path));
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_free(ibx_stl* s) {
  g_stl.erase(s);
  return IBX_OK;
}

int ibx_stl_info(const ibx_stl* s, int* nd, int64_t* npoints, int64_t* nsimp, int* is_f32) {
  IBX_TRY
  auto p = lookup_stl(s);
  *nd = p->nd;
  *npoints = p->npoints();
  *nsimp = p->nsimp();
  *is_f32 = p->f32;
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_copy(const ibx_stl* s, double* points, int64_t* simplices) {
  IBX_TRY
  auto p = lookup_stl(s);
  std::copy(p->points.begin(), p->points.end(), points);
  std::copy(p->simplices.begin(), p->simplices.end(), simplices);
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_merge_points(int n, ibx_stl* const* in, double tolerance, int tol_is_f32, int clean_degenerate,
                         ibx_stl** out) {
  IBX_TRY
  IBX_REQUIRE(n >= 1, "need at least one input");
  std::vector<const ibx_stl*> v;
  for (int i = 0; i < n; ++i) v.push_back(lookup_stl(in[i]).get());
  *out = reg_stl(merge_points_impl(v, tolerance, tol_is_f32 != 0, clean_degenerate != 0));
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_feature_regions(const ibx_stl* sh, double angle_deg, double radius, int include_boundaries, ibx_stl** out) {
  IBX_TRY
  // feature_regions (src/mesher.jl:670-728)
  auto s = lookup_stl(sh);
  int nd = s->nd;
  int64_t ns = s->nsimp();
  double eps = 1.1920928955078125e-07;
  double angle = std::max(angle_deg, 1.0) * M_PI / 180.0;
  double max_cos = std::cos(0.05 * M_PI / 180.0);
  std::map<std::vector<int64_t>, int64_t> registry;
  std::vector<std::pair<int64_t, int64_t>> edges;
  for (int64_t i = 0; i < ns; ++i)
    for (int pv = 0; pv < nd; ++pv) {
      std::vector<int64_t> face;
      int64_t pivot = s->simplices[i * nd + pv];
      for (int v = 0; v < nd; ++v)
        if (s->simplices[i * nd + v] != pivot) face.push_back(s->simplices[i * nd + v]);
      std::sort(face.begin(), face.end());
      auto it = registry.find(face);
      if (it != registry.end()) {
        edges.emplace_back(it->second, i);
        registry.erase(it);
      } else {
        registry[face] = i;
      }
    }
  for (auto& kv : registry) edges.emplace_back(kv.second, kv.second);
  std::vector<double> centers, normals;
  simplex_centers_normals(*s, centers, normals);
  std::vector<char> inc(ns, 0);
  for (auto& e : edges) {
    int64_t i = e.first, j = e.second;
    double ni[3], nj[3], li = 0, lj = 0, dot = 0, dd = 0;
    for (int d = 0; d < nd; ++d) { li += normals[i * nd + d] * normals[i * nd + d]; lj += normals[j * nd + d] * normals[j * nd + d]; }
    li = std::sqrt(li) + eps;
    lj = std::sqrt(lj) + eps;
    for (int d = 0; d < nd; ++d) {
      ni[d] = normals[i * nd + d] / li;
      nj[d] = normals[j * nd + d] / lj;
      if (s->f32) { ni[d] = (double)(float)ni[d]; nj[d] = (double)(float)nj[d]; }
      dot += ni[d] * nj[d];
      double c = centers[i * nd + d] - centers[j * nd + d];
      dd += c * c;
    }
    if (s->f32) dot = (double)(float)dot;
    double theta = std::acos(std::min(dot, max_cos));
    double dist = std::sqrt(dd);
    if (s->f32) dist = (double)(float)dist;
    if ((i == j && include_boundaries) || (dist / theta < radius) || (theta > angle)) inc[i] = inc[j] = 1;
  }
  auto o = std::make_shared<ibx_stl>();
  o->nd = nd;
  o->f32 = s->f32;
  o->points = s->points;
  for (int64_t i = 0; i < ns; ++i)
    if (inc[i])
      for (int d = 0; d < nd; ++d) o->simplices.push_back(s->simplices[i * nd + d]);
  *out = reg_stl(o);
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_centers_normals(const ibx_stl* s, double* centers, double* normals) {
  IBX_TRY
  auto p = lookup_stl(s);
  std::vector<double> c, n;
  simplex_centers_normals(*p, c, n);
  std::copy(c.begin(), c.end(), centers);
  std::copy(n.begin(), n.end(), normals);
  return IBX_OK;
  IBX_CATCH
}

int ibx_dfield_create(const ibx_stl* s, ibx_dfield** out) {
  IBX_TRY
  *out = register_dfield(make_dfield(lookup_stl(s)));
  return IBX_OK;
  IBX_CATCH
}

int ibx_dfield_free(ibx_dfield* d) {
  g_df.erase(d);
  return IBX_OK;
}

int ibx_dfield_stl(const ibx_dfield* d, const ibx_stl** out) {
  IBX_TRY
  auto p = lookup_dfield(d);
  if (p->stl) { g_stl[p->stl.get()] = p->stl; *out = p->stl.get(); } else { *out = nullptr; }
  return IBX_OK;
  IBX_CATCH
}

int ibx_dfield_distance(const ibx_dfield* d, int64_t n, const float* x, double* out) {
  IBX_TRY
  auto p = lookup_dfield(d);
  int nd = p->sphere ? 3 : p->stl->nd;
  for (int64_t i = 0; i < n; ++i) {
    double q[3];
    for (int k = 0; k < nd; ++k) q[k] = x[i * nd + k];
    out[i] = p->distance(q, true).v;
  }
  return IBX_OK;
  IBX_CATCH
}

int ibx_stl_refine_to_length(const ibx_stl* s, double h, int h_is_f32, double tolerance, int tol_is_f32,
                             double growth_ratio, int nregions, const ibx_region* regions, ibx_stl** out) {
  IBX_TRY
  auto p = lookup_stl(s);
  *out = reg_stl(refine_to_length_impl(*p, Num{h, h_is_f32 != 0}, tolerance, tol_is_f32 != 0, growth_ratio, nregions, regions));
  return IBX_OK;
  IBX_CATCH
}

}  // extern "C"
