// Host mesher: block octree refinement and cell enumeration.
// Mirrors src/mesher.jl:811-1112 of the reference (refine_octree, refine_orderly, Mesh, get_cells).
#include "ibx_internal.h"
#include <cstdio>
#include <cstdlib>
#include <omp.h>

#include <numeric>

namespace ibx {
Num region_distance(const ibx_region& r, int nd, const double* x, bool xf32);
std::shared_ptr<ibx_stl> refine_to_length_impl(const ibx_stl& s, Num h, double tol, bool tol_f32, double growth_ratio,
                                               int nreg, const ibx_region* regs);
std::shared_ptr<ibx_dfield> make_dfield(std::shared_ptr<ibx_stl> stl);
std::shared_ptr<ibx_stl> lookup_stl(const ibx_stl* s);
std::shared_ptr<ibx_dfield> lookup_dfield(const ibx_dfield* d);
ibx_dfield* register_dfield(std::shared_ptr<ibx_dfield> d);

static std::map<const ibx_mesh*, std::shared_ptr<ibx_mesh>> g_mesh;
std::shared_ptr<ibx_mesh> lookup_mesh(const ibx_mesh* m) {
  auto it = g_mesh.find(m);
  if (it == g_mesh.end()) throw std::runtime_error("unknown ibx_mesh handle");
  return it->second;
}

struct Criterion {
  ibx_region region;
  Num h;
};

// refine_octree (src/mesher.jl:811-862): depth-first, children in Iterators.product order (first dim fastest).
// The subtrees below depth PAR_DEPTH are independent: the top of the tree is walked serially, every node met at that
// depth becomes a task, the tasks run in parallel and their leaves are spliced back in depth-first order -- the block
// list is the serial one.
namespace {
struct OctItem {
  float o[3], w[3];
  int depth;
  std::vector<int> active;  // indices into crit
};
struct OctOut {
  std::vector<float> o, w;
  std::vector<OctItem> tasks;                // nodes left unexpanded (serial top only)
  std::vector<std::pair<int64_t, int>> at;   // (number of leaves emitted before the task, task id)
};

void octree_dfs(const std::vector<Criterion>& crit, int nd, double gm1, OctItem root, int stop_depth, OctOut& out) {
  std::vector<OctItem> stack;
  stack.push_back(std::move(root));
  while (!stack.empty()) {
    OctItem it = std::move(stack.back());
    stack.pop_back();
    if (stop_depth >= 0 && it.depth == stop_depth) {
      out.at.emplace_back((int64_t)out.o.size() / nd, (int)out.tasks.size());
      out.tasks.push_back(std::move(it));
      continue;
    }
    float L = it.w[0], wmin = it.w[0];
    double ss = 0;
    for (int d = 0; d < nd; ++d) {
      L = std::max(L, it.w[d]);
      wmin = std::min(wmin, it.w[d]);
      ss += (double)it.w[d] * (double)it.w[d];
    }
    float R = (float)std::sqrt(ss) / 2.0f;  // norm(widths) / 2, circumradius
    double c[3];
    for (int d = 0; d < nd; ++d) c[d] = (double)(float)(it.o[d] + it.w[d] / 2.0f);
    std::vector<int> active;
    for (int ci : it.active) {
      Num dist = region_distance(crit[ci].region, nd, c, true);
      Num lmax = nmax(nmul(Num{gm1, false}, nsub(dist, Num{(double)R, true})), crit[ci].h);
      if (lmax.v < (double)L) active.push_back(ci);
    }
    if (active.empty()) {
      for (int d = 0; d < nd; ++d) { out.o.push_back(it.o[d]); out.w.push_back(it.w[d]); }
      continue;
    }
    int split[3] = {1, 1, 1};
    float nw[3];
    std::vector<float> axes[3];
    for (int d = 0; d < nd; ++d) {
      split[d] = (int)std::nearbyint((double)(it.w[d] / wmin)) + 1;
      nw[d] = it.w[d] / (float)split[d];
      double a = (double)it.o[d], b = (double)(float)(it.o[d] + it.w[d]);
      for (int s = 0; s < split[d]; ++s) {
        double t = (double)s / (double)split[d];
        axes[d].push_back((float)((1.0 - t) * a + t * b));  // LinRange lerp (Base.lerpi)
      }
    }
    int total = split[0] * split[1] * split[2];
    // push in reverse so that children pop in product order
    for (int idx = total - 1; idx >= 0; --idx) {
      int i0 = idx % split[0], i1 = (idx / split[0]) % split[1], i2 = idx / (split[0] * split[1]);
      int ii[3] = {i0, i1, i2};
      OctItem ch;
      for (int d = 0; d < nd; ++d) { ch.o[d] = axes[d][ii[d]]; ch.w[d] = nw[d]; }
      ch.depth = it.depth + 1;
      ch.active = active;
      stack.push_back(std::move(ch));
    }
  }
}
}  // namespace

static void refine_octree(const std::vector<Criterion>& crit, int nd, const float* origin, const float* widths,
                          double gm1, std::vector<float>& out_o, std::vector<float>& out_w) {
  constexpr int PAR_DEPTH = 5;
  OctItem root;
  for (int d = 0; d < nd; ++d) { root.o[d] = origin[d]; root.w[d] = widths[d]; }
  root.depth = 0;
  root.active.resize(crit.size());
  std::iota(root.active.begin(), root.active.end(), 0);
  OctOut top;
  octree_dfs(crit, nd, gm1, std::move(root), PAR_DEPTH, top);
  const int nt = (int)top.tasks.size();
  std::vector<OctOut> sub((size_t)nt);
  std::string err;
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < nt; ++t) {
    try {
      octree_dfs(crit, nd, gm1, std::move(top.tasks[t]), -1, sub[(size_t)t]);
    } catch (const std::exception& e) {
#pragma omp critical
      err = e.what();
    }
  }
  if (!err.empty()) throw std::runtime_error(err);
  int64_t pos = 0;   // leaves of the serial top copied so far
  auto copy_top = [&](int64_t upto) {
    out_o.insert(out_o.end(), top.o.begin() + pos * nd, top.o.begin() + upto * nd);
    out_w.insert(out_w.end(), top.w.begin() + pos * nd, top.w.begin() + upto * nd);
    pos = upto;
  };
  for (const auto& a : top.at) {
    copy_top(a.first);
    const OctOut& s = sub[(size_t)a.second];
    out_o.insert(out_o.end(), s.o.begin(), s.o.end());
    out_w.insert(out_w.end(), s.w.begin(), s.w.end());
  }
  copy_top((int64_t)top.o.size() / nd);
}

// get_cells, margin = 0 (src/mesher.jl:1064-1112) for ONE block: cpb x nd centres / widths, first dim fastest.  The
// float32 operations are CellGeom's (ibx_internal.h): the builder evaluates the same expressions cell by cell.
void mesh_block_cells(const ibx_mesh& m, int64_t b, float* centers, float* widths) {
  CellGeom G;
  G.init(m);
  const int nd = m.nd;
  const int64_t cpb = m.cells_per_block();
  const int bs = m.block_size;
  float cw[3] = {0.f, 0.f, 0.f};
  for (int d = 0; d < nd; ++d) cw[d] = G.w(b * cpb, d);
  for (int64_t l = 0; l < cpb; ++l) {
    int rem = (int)l;
    for (int d = 0; d < nd; ++d) {
      const int i = rem % bs;
      rem /= bs;
      if (centers) centers[l * nd + d] = G.cb(b, i, d);
      if (widths) widths[l * nd + d] = cw[d];
    }
  }
}

void mesh_cells(const ibx_mesh& m, float* centers, float* widths) {
  // block-major, first dim fastest in a block
  int nd = m.nd;
  int64_t cpb = m.cells_per_block(), nb = m.nblocks();
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < nb; ++b)
    mesh_block_cells(m, b, centers ? centers + b * cpb * nd : nullptr, widths ? widths + b * cpb * nd : nullptr);
}

}  // namespace ibx

using namespace ibx;

namespace {
const char kMeshMagic[8] = {'I', 'B', 'X', 'M', 'E', 'S', 'H', '1'};
template <class T> void put(FILE* f, const T& v) { if (fwrite(&v, sizeof(T), 1, f) != 1) throw std::runtime_error("write failed"); }
template <class T> void put_vec(FILE* f, const std::vector<T>& v) {
  int64_t n = (int64_t)v.size();
  put(f, n);
  if (n && fwrite(v.data(), sizeof(T), (size_t)n, f) != (size_t)n) throw std::runtime_error("write failed");
}
template <class T> void get(FILE* f, T& v) { if (fread(&v, sizeof(T), 1, f) != 1) throw std::runtime_error("unexpected end of mesh file"); }
template <class T> void get_vec(FILE* f, std::vector<T>& v) {
  int64_t n;
  get(f, n);
  if (n < 0 || n > ((int64_t)1 << 40)) throw std::runtime_error("corrupt mesh file");
  v.resize((size_t)n);
  if (n && fread(v.data(), sizeof(T), (size_t)n, f) != (size_t)n) throw std::runtime_error("unexpected end of mesh file");
}
}  // namespace

extern "C" {

int ibx_mesh_create(int nd, const float* origin, const float* widths, int nsurf, const ibx_surface* surfaces,
                    int nregions, const ibx_region* regions, double growth_ratio, double tolerance, int tol_is_f32,
                    int block_size, ibx_mesh** out) {
  IBX_TRY
  IBX_REQUIRE(nd == 2 || nd == 3, "nd must be 2 or 3");
  IBX_REQUIRE(block_size >= 1 && block_size <= 64, "block_size must be in 1 .. 64");
  auto m = std::make_shared<ibx_mesh>();
  m->nd = nd;
  m->block_size = block_size;
  for (int d = 0; d < nd; ++d) { m->origin[d] = origin[d]; m->widths[d] = widths[d]; }
  const bool verbose = getenv("IBX_BUILD_VERBOSE") != nullptr;   // phase times, like the Domain build
  double t_phase = omp_get_wtime();
  auto phase = [&](const char* what) {
    if (verbose) fprintf(stderr, "[ibx mesh ] %-28s %8.2f s\n", what, omp_get_wtime() - t_phase);
    t_phase = omp_get_wtime();
  };
  // refine_orderly (src/mesher.jl:878-918): surfaces by increasing h; every refined surface becomes a
  // refinement region (at h * ratio, ratio = 0.5f0) for the ones that follow
  const float ratio = 0.5f;
  auto scaled = [&](double h, bool f32, float fac) -> Num {
    if (f32) return {(double)((float)h * fac), true};
    return {h * (double)fac, false};
  };
  std::vector<ibx_region> regs;
  for (int r = 0; r < nregions; ++r) {
    ibx_region rr = regions[r];
    if (rr.kind == 3) rr.dfield = lookup_dfield(rr.dfield).get();
    Num hs = scaled(rr.h, rr.h_is_f32 != 0, ratio);
    rr.h = hs.v;
    rr.h_is_f32 = hs.f32;
    regs.push_back(rr);
  }
  std::vector<int> order(nsurf);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return surfaces[a].h < surfaces[b].h; });
  m->surf_names.resize(nsurf);
  m->surf_fields.resize(nsurf);
  for (int i : order) {
    const ibx_surface& sf = surfaces[i];
    Num h = scaled(sf.h, sf.h_is_f32 != 0, ratio);
    std::shared_ptr<ibx_dfield> df;
    if (sf.stl) {
      auto stl = refine_to_length_impl(*lookup_stl(sf.stl), h, tolerance, tol_is_f32 != 0, growth_ratio,
                                       (int)regs.size(), regs.data());
      phase("refine_to_length");
      df = make_dfield(stl);
      phase("distance field");
    } else {
      IBX_REQUIRE(nd == 3, "analytic sphere surfaces are 3-D only");
      df = std::make_shared<ibx_dfield>();
      df->sphere = true;
      for (int d = 0; d < 3; ++d) df->sc[d] = sf.sphere_c[d];
      df->sr = sf.sphere_r;
    }
    register_dfield(df);
    m->surf_names[i] = sf.name ? sf.name : "";
    m->surf_fields[i] = df;
    ibx_region rr{};
    rr.kind = 3;
    rr.dfield = df.get();
    rr.h = h.v;
    rr.h_is_f32 = h.f32;
    regs.push_back(rr);
  }
  // refinement criteria of the octree (src/mesher.jl:1013-1021): sizes times block_size
  std::vector<Criterion> crit;
  auto times_bs = [&](double h, bool f32) -> Num {
    if (f32) return {(double)((float)h * (float)block_size), true};
    return {h * (double)block_size, false};
  };
  for (int r = 0; r < nregions; ++r) {
    Criterion c;
    c.region = regions[r];
    if (c.region.kind == 3) c.region.dfield = lookup_dfield(regions[r].dfield).get();
    c.h = times_bs(regions[r].h, regions[r].h_is_f32 != 0);
    crit.push_back(c);
  }
  for (int i = 0; i < nsurf; ++i) {
    Criterion c;
    c.region = ibx_region{};
    c.region.kind = 3;
    c.region.dfield = m->surf_fields[i].get();
    c.h = times_bs(surfaces[i].h, surfaces[i].h_is_f32 != 0);
    crit.push_back(c);
  }
  refine_octree(crit, nd, m->origin, m->widths, growth_ratio - 1.0, m->block_origins, m->block_widths);
  phase("octree");
  g_mesh[m.get()] = m;
  *out = m.get();
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_from_blocks(const ibx_mesh* like, int block_size, ibx_mesh** out) {
  IBX_TRY
  auto src = lookup_mesh(like);
  IBX_REQUIRE(block_size >= 1 && block_size <= 64, "block_size must be in 1 .. 64");
  auto m = std::make_shared<ibx_mesh>(*src);
  m->block_size = block_size;
  g_mesh[m.get()] = m;
  *out = m.get();
  return IBX_OK;
  IBX_CATCH
}

// ---- mesh (de)serialisation.  In the reference a Mesh is plain data (src/mesher.jl:926-933) that Julia's `Serialization`
// stdlib writes as is; behind this ABI it is an opaque handle, so the library writes it itself: header, root box, block
// list, and per surface its name and refined STL (or the analytic-sphere parameters).  Distance fields (KD-trees) are
// rebuilt on load, exactly as Mesh(...) builds them.

int ibx_mesh_save(const ibx_mesh* mh, const char* path) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  FILE* f = fopen(path, "wb");
  IBX_REQUIRE(f != nullptr, std::string("cannot open ") + path + " for writing");
  try {
    fwrite(kMeshMagic, 1, 8, f);
    put(f, (int32_t)m->nd);
    put(f, (int32_t)m->block_size);
    for (int d = 0; d < 3; ++d) { put(f, m->origin[d]); put(f, m->widths[d]); }
    put_vec(f, m->block_origins);
    put_vec(f, m->block_widths);
    put(f, (int32_t)m->surf_names.size());
    for (size_t s = 0; s < m->surf_names.size(); ++s) {
      std::vector<char> name(m->surf_names[s].begin(), m->surf_names[s].end());
      put_vec(f, name);
      const ibx_dfield& df = *m->surf_fields[s];
      put(f, (int32_t)(df.sphere ? 1 : 0));
      if (df.sphere) {
        for (int d = 0; d < 3; ++d) put(f, df.sc[d]);
        put(f, df.sr);
      } else {
        put(f, (int32_t)df.stl->nd);
        put(f, (int32_t)(df.stl->f32 ? 1 : 0));
        put_vec(f, df.stl->points);
        put_vec(f, df.stl->simplices);
      }
    }
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_load(const char* path, ibx_mesh** out) {
  IBX_TRY
  FILE* f = fopen(path, "rb");
  IBX_REQUIRE(f != nullptr, std::string("cannot open ") + path);
  auto m = std::make_shared<ibx_mesh>();
  try {
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMeshMagic, 8) != 0) throw std::runtime_error(std::string(path) + " is not an ibx mesh file");
    int32_t nd, bs, ns;
    get(f, nd);
    get(f, bs);
    if (nd < 2 || nd > 3 || bs < 1 || bs > 64) throw std::runtime_error("corrupt mesh file");
    m->nd = nd;
    m->block_size = bs;
    for (int d = 0; d < 3; ++d) { get(f, m->origin[d]); get(f, m->widths[d]); }
    get_vec(f, m->block_origins);
    get_vec(f, m->block_widths);
    if (m->block_origins.size() != m->block_widths.size() || m->block_origins.size() % nd) throw std::runtime_error("corrupt mesh file");
    get(f, ns);
    for (int s = 0; s < ns; ++s) {
      std::vector<char> name;
      get_vec(f, name);
      m->surf_names.emplace_back(name.begin(), name.end());
      int32_t sphere;
      get(f, sphere);
      std::shared_ptr<ibx_dfield> df;
      if (sphere) {
        df = std::make_shared<ibx_dfield>();
        df->sphere = true;
        for (int d = 0; d < 3; ++d) get(f, df->sc[d]);
        get(f, df->sr);
      } else {
        auto stl = std::make_shared<ibx_stl>();
        int32_t snd, f32;
        get(f, snd);
        get(f, f32);
        stl->nd = snd;
        stl->f32 = f32 != 0;
        get_vec(f, stl->points);
        get_vec(f, stl->simplices);
        df = make_dfield(stl);
      }
      register_dfield(df);
      m->surf_fields.push_back(df);
    }
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
  g_mesh[m.get()] = m;
  *out = m.get();
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_free(ibx_mesh* m) {
  g_mesh.erase(m);
  return IBX_OK;
}

int ibx_mesh_info(const ibx_mesh* mh, int* nd, int* block_size, int64_t* nblocks, int64_t* ncells, int* nsurf) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  *nd = m->nd;
  *block_size = m->block_size;
  *nblocks = m->nblocks();
  *ncells = m->ncells();
  *nsurf = (int)m->surf_names.size();
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_blocks(const ibx_mesh* mh, float* block_origins, float* block_widths) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  std::copy(m->block_origins.begin(), m->block_origins.end(), block_origins);
  std::copy(m->block_widths.begin(), m->block_widths.end(), block_widths);
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_surface_name(const ibx_mesh* mh, int i, const char** name) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  IBX_REQUIRE(i >= 0 && i < (int)m->surf_names.size(), "surface index out of range");
  *name = m->surf_names[i].c_str();
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_surface_dfield(const ibx_mesh* mh, int i, const ibx_dfield** out) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  IBX_REQUIRE(i >= 0 && i < (int)m->surf_fields.size(), "surface index out of range");
  *out = m->surf_fields[i].get();
  return IBX_OK;
  IBX_CATCH
}

int ibx_mesh_cells(const ibx_mesh* mh, float* centers, float* widths) {
  IBX_TRY
  auto m = lookup_mesh(mh);
  mesh_cells(*m, centers, widths);
  return IBX_OK;
  IBX_CATCH
}

}  // extern "C"
