/* Oracle (TEST INFRASTRUCTURE, CPU baseline): plain-C restatement of the reference's partitioned Euler residual +
 * immersed-boundary ghost update, written the way the reference executes it -- one task per partition
 * (ThreadTools.tmap, src/ImmersedBoundary.jl:834 -> an OpenMP loop), gather -> array-at-a-time operators with
 * temporaries -> scatter of the image rows (:820-864).  It exists so that bench.py's cpu_baseline / --impl reference
 * legs time compiled code on all host cores instead of NumPy; it is validated bit for bit against the NumPy oracle
 * (tests/test_oracle_cpu_ref.py) and is never linked into, or called by, the product.
 *
 * Arrays are column-major (Julia layout): a[col * rows + row].  Float32 arithmetic in the reference's operation order
 * (compile with -ffp-contract=off); the HLL flux and the Green-Gauss sums taken of it are Float64 (src/cfd.jl:504-507).
 *   operators: at_faces :899-910, green_gauss :918-926, unsigned_green_gauss :934-942, JST_sensor :1077-1097,
 *   minmod :1099, MUSCL :1113-1157, impose_bc! :1197-1247; src/cfd.jl: state2primitive :137-151, primitive2state
 *   :106-123, speed_of_sound :62-64, FlowBC :243-300, inviscid_fluxes (HLL) :459-508; Accumulator src/accumulator.jl:78-111.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int64_t n_dom, n_img;
  const int64_t *domain, *image, *image_in_domain; /* global rows of the domain / image; image rows inside the domain */
  const float* spacing;                           /* n_dom x nd */
  int64_t nf[3];
  const int64_t *own[3], *nei[3];                 /* face -> owner / neighbour (domain-local) */
  const int64_t *lptr[3], *lidx[3], *rptr[3], *ridx[3]; /* cell -> faces on its low / high side (CSR), weights 1/len */
} ibxref_part;

typedef struct {
  int64_t n_ghost, n_img_dom, nnz;
  const int64_t *ghost, *image_domain, *ptr, *idx; /* ghost rows (global); donors (global); CSR into image_domain-local ids */
  const float *w, *normals, *eta;                  /* weights; n_ghost x nd; ghost_distance / image_distance */
  int normal_flow, n_pinf;
  float pinf[5];
} ibxref_bdry;

static inline float clampT(float T) { return T > 10.0f ? T : 10.0f; }
static inline float sgnf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

/* out[i] = sum_j v[idx[j]] * (1/len), products first, summed in list order; 0 for an empty list */
static void acc_f32(const int64_t* ptr, const int64_t* idx, int64_t n, const float* v, float* out) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t a = ptr[i], b = ptr[i + 1];
    float s = 0.0f;
    if (b > a) {
      float w = 1.0f / (float)(b - a);
      s = v[idx[a]] * w;
      for (int64_t j = a + 1; j < b; ++j) s = s + v[idx[j]] * w;
    }
    out[i] = s;
  }
}
static void acc_f64(const int64_t* ptr, const int64_t* idx, int64_t n, const double* v, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t a = ptr[i], b = ptr[i + 1];
    double s = 0.0;
    if (b > a) {
      float w = 1.0f / (float)(b - a);
      s = v[idx[a]] * (double)w;
      for (int64_t j = a + 1; j < b; ++j) s = s + v[idx[j]] * (double)w;
    }
    out[i] = s;
  }
}

static void s2p(float R, float gamma, int nd, int64_t n, const float* Q, float* P) {
  for (int64_t i = 0; i < n; ++i) {
    float rho = Q[i], E = Q[n + i], u[3], k = 0.0f;
    for (int d = 0; d < nd; ++d) {
      u[d] = Q[(2 + d) * n + i] / rho;
      k = d == 0 ? u[0] * u[0] : k + u[d] * u[d];
    }
    k = k / 2.0f;
    float p = (gamma - 1.0f) * (E - rho * k);
    P[i] = p;
    P[n + i] = clampT(p / (rho * R));
    for (int d = 0; d < nd; ++d) P[(2 + d) * n + i] = u[d];
  }
}
static inline void p2s_point(float R, float gamma, int nd, const float* P, float* Q) {
  float T = clampT(P[1]), k = P[2] * P[2];
  for (int d = 1; d < nd; ++d) k = k + P[2 + d] * P[2 + d];
  k = k / 2.0f;
  float rho = P[0] / (R * T);
  Q[0] = rho;
  Q[1] = rho * (R / (gamma - 1.0f) * T + k);
  for (int d = 0; d < nd; ++d) Q[2 + d] = rho * P[2 + d];
}

/* one partition: Qd (n_dom x nv, gathered) -> Rd, cd (n_dom) */
static void residual_partition(const ibxref_part* p, int nd, float Rg, float gamma, const float* Qd, float* Rd, float* cd, int use_sensor) {
  const int nv = nd + 2;
  const int64_t n = p->n_dom;
  int64_t nfmax = 0;
  for (int d = 0; d < nd; ++d) nfmax = p->nf[d] > nfmax ? p->nf[d] : nfmax;
  float* P = malloc(sizeof(float) * n * nv);
  float* D = malloc(sizeof(float) * n);
  float* a = malloc(sizeof(float) * n);
  float* t1 = malloc(sizeof(float) * n);
  float* t2 = malloc(sizeof(float) * n);
  float* gP = malloc(sizeof(float) * n * nv);
  float* f1 = malloc(sizeof(float) * nfmax);
  float* f2 = malloc(sizeof(float) * nfmax);
  float* Pf = malloc(sizeof(float) * nfmax * nv);
  float* PL = malloc(sizeof(float) * nfmax * nv);
  float* PR = malloc(sizeof(float) * nfmax * nv);
  double* Fx = malloc(sizeof(double) * nfmax * nv);
  double* d1 = malloc(sizeof(double) * n);
  double* d2 = malloc(sizeof(double) * n);
  s2p(Rg, gamma, nd, n, Qd, P);
  /* JST sensor, dim = 0 form: max over dims, floor 1e-7 */
  for (int64_t i = 0; i < n; ++i) D[i] = 1e-7f;
  for (int d = 0; d < nd; ++d) {
    const int64_t nf = p->nf[d];
    const float* h = p->spacing + (int64_t)d * n;
    for (int64_t f = 0; f < nf; ++f) { f1[f] = P[p->nei[d][f]] - P[p->own[d][f]]; f2[f] = fabsf(f1[f]); }
    acc_f32(p->rptr[d], p->ridx[d], n, f1, t1);
    acc_f32(p->lptr[d], p->lidx[d], n, f1, t2);
    for (int64_t i = 0; i < n; ++i) t1[i] = fabsf((t1[i] - t2[i]) / h[i]);
    acc_f32(p->rptr[d], p->ridx[d], n, f2, a);
    acc_f32(p->lptr[d], p->lidx[d], n, f2, t2);
    for (int64_t i = 0; i < n; ++i) {
      float nu = (1e-7f + t1[i]) / (1e-7f + (a[i] + t2[i]) / h[i]);
      D[i] = D[i] > nu ? D[i] : nu;
    }
  }
  const float gr = gamma * Rg;
  for (int64_t i = 0; i < n; ++i) a[i] = sqrtf(gr * clampT(P[n + i]));
  memset(Rd, 0, sizeof(float) * n * nv);
  memset(cd, 0, sizeof(float) * n);
  for (int d = 0; d < nd; ++d) {
    const int64_t nf = p->nf[d];
    const int64_t *own = p->own[d], *nei = p->nei[d];
    const float* h = p->spacing + (int64_t)d * n;
    /* cell_gradient = green_gauss(at_faces(P)) */
    for (int v = 0; v < nv; ++v) {
      const float* u = P + (int64_t)v * n;
      float* uf = Pf + (int64_t)v * nf;
      for (int64_t f = 0; f < nf; ++f) {
        float spo = h[own[f]], spn = h[nei[f]];
        uf[f] = (u[own[f]] * spn + u[nei[f]] * spo) / (spn + spo);
      }
      acc_f32(p->rptr[d], p->ridx[d], n, uf, t1);
      acc_f32(p->lptr[d], p->lidx[d], n, uf, t2);
      float* g = gP + (int64_t)v * n;
      for (int64_t i = 0; i < n; ++i) g[i] = (t1[i] - t2[i]) / h[i];
    }
    /* MUSCL with the sensor blend (high_order = false) */
    for (int v = 0; v < nv; ++v) {
      const float *u = P + (int64_t)v * n, *du = gP + (int64_t)v * n;
      for (int64_t f = 0; f < nf; ++f) {
        int64_t o = own[f], q = nei[f];
        float down = h[o] / 2.0f, dnei = h[q] / 2.0f;
        float uo = u[o], un = u[q];
        float gf = (un - uo) / (down + dnei);
        float gu = (2.0f * du[o] - gf) * down, Du = (2.0f * du[q] - gf) * dnei;
        float s = fminf(fabsf(Du), fabsf(gu)) * (sgnf(Du) + sgnf(gu)) / 2.0f;
        float l = uo + s, r = un - s;
        if (!use_sensor) { /* MUSCL(...; D = nothing): no blend (src/ImmersedBoundary.jl:1141) */
          PL[(int64_t)v * nf + f] = l;
          PR[(int64_t)v * nf + f] = r;
          continue;
        }
        float Df = fmaxf(fmaxf(D[o], D[q]), 1e-7f);
        float ufc = (uo * dnei + un * down) / (down + dnei);
        PL[(int64_t)v * nf + f] = l * Df + (1.0f - Df) * ufc;
        PR[(int64_t)v * nf + f] = r * Df + (1.0f - Df) * ufc;
      }
    }
    /* HLL, Float64 out */
    for (int64_t f = 0; f < nf; ++f) {
      float pl[5], pr[5], ql[5], qr[5];
      for (int v = 0; v < nv; ++v) { pl[v] = PL[(int64_t)v * nf + f]; pr[v] = PR[(int64_t)v * nf + f]; }
      p2s_point(Rg, gamma, nd, pl, ql);
      p2s_point(Rg, gamma, nd, pr, qr);
      float uL = pl[2 + d], uR = pr[2 + d];
      float aL = sqrtf(gr * clampT(pl[1])), aR = sqrtf(gr * clampT(pr[1]));
      double SR = fmin((double)(uR - aR), 0.0), SL = fmax((double)(uL + aL), 0.0);
      for (int v = 0; v < nv; ++v) {
        float l = ql[v], r = qr[v];
        if (v == 1) { l = l + pl[0]; r = r + pr[0]; }
        l = l * uL;
        r = r * uR;
        if (v == 2 + d) { l = l + pl[0]; r = r + pr[0]; }
        Fx[(int64_t)v * nf + f] = (SL * (double)l - SR * (double)r + SR * SL * (double)(qr[v] - ql[v])) / (SL - SR);
      }
    }
    /* R .-= green_gauss(F): Float64 sums, rounded into the Float32 residual */
    for (int v = 0; v < nv; ++v) {
      acc_f64(p->rptr[d], p->ridx[d], n, Fx + (int64_t)v * nf, d1);
      acc_f64(p->lptr[d], p->lidx[d], n, Fx + (int64_t)v * nf, d2);
      float* r = Rd + (int64_t)v * n;
      for (int64_t i = 0; i < n; ++i) r[i] = (float)((double)r[i] - (d1[i] - d2[i]) / (double)h[i]);
    }
    /* cfl += unsigned_green_gauss(|at_faces(u_d)| + at_faces(a)) */
    const float* ud = P + (int64_t)(2 + d) * n;
    for (int64_t f = 0; f < nf; ++f) {
      float spo = h[own[f]], spn = h[nei[f]];
      float uf = (ud[own[f]] * spn + ud[nei[f]] * spo) / (spn + spo);
      float af = (a[own[f]] * spn + a[nei[f]] * spo) / (spn + spo);
      f1[f] = fabsf(uf) + af;
    }
    acc_f32(p->rptr[d], p->ridx[d], n, f1, t1);
    acc_f32(p->lptr[d], p->lidx[d], n, f1, t2);
    for (int64_t i = 0; i < n; ++i) cd[i] = cd[i] + (t1[i] + t2[i]) / h[i];
  }
  free(P); free(D); free(a); free(t1); free(t2); free(gP); free(f1); free(f2); free(Pf); free(PL); free(PR); free(Fx); free(d1); free(d2);
}

/* dom(f, Q, R, cfl): one task per partition, gather -> residual -> scatter of the image rows */
int ibxref_residual_ex(int nparts, const ibxref_part* parts, int nd, float Rg, float gamma, int64_t N, const float* Q, float* R,
                       float* cfl, int nthreads, int use_sensor) {
  const int nv = nd + 2;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int ip = 0; ip < nparts; ++ip) {
    const ibxref_part* p = parts + ip;
    const int64_t n = p->n_dom;
    float* Qd = malloc(sizeof(float) * n * nv);
    float* Rd = malloc(sizeof(float) * n * nv);
    float* cd = malloc(sizeof(float) * n);
    for (int v = 0; v < nv; ++v)
      for (int64_t i = 0; i < n; ++i) Qd[(int64_t)v * n + i] = Q[(int64_t)v * N + p->domain[i]];
    residual_partition(p, nd, Rg, gamma, Qd, Rd, cd, use_sensor);
    for (int64_t k = 0; k < p->n_img; ++k) {
      int64_t g = p->image[k], l = p->image_in_domain[k];
      for (int v = 0; v < nv; ++v) R[(int64_t)v * N + g] = Rd[(int64_t)v * n + l];
      cfl[g] = cd[l];
    }
    free(Qd); free(Rd); free(cd);
  }
  return 0;
}

int ibxref_residual(int nparts, const ibxref_part* parts, int nd, float Rg, float gamma, int64_t N, const float* Q, float* R,
                    float* cfl, int nthreads) {
  return ibxref_residual_ex(nparts, parts, nd, Rg, gamma, N, Q, R, cfl, nthreads, 1);
}

/* FlowBC(P, normals) at one point (src/cfd.jl:243-300, no wall-shear scaling) */
static void flowbc_point(const ibxref_bdry* b, int nd, float Rg, float gamma, const float* P, const float* nrm, float* out) {
  float un;
  if (b->normal_flow) un = b->pinf[2];
  else {
    un = nrm[0] * b->pinf[2];
    for (int d = 1; d < nd; ++d) un = un + nrm[d] * b->pinf[2 + d];
  }
  float cur = P[2] * nrm[0];
  for (int d = 1; d < nd; ++d) cur = cur + P[2 + d] * nrm[d];
  float a = sqrtf(gamma * Rg * clampT(P[1]));
  float M = fabsf(un) / a;
  float sup = M > 1.0f ? 1.0f : 0.0f, sub = M <= 1.0f ? 1.0f : 0.0f;
  float ge = un >= 0.0f ? 1.0f : 0.0f, lt = un < 0.0f ? 1.0f : 0.0f;
  out[0] = ge * (sup * b->pinf[0] + sub * P[0]) + lt * (sup * P[0] + sub * b->pinf[0]);
  out[1] = (un > 0.0f ? 1.0f : 0.0f) * b->pinf[1] + (un <= 0.0f ? 1.0f : 0.0f) * P[1];
  if (b->normal_flow) {
    float corr = un - cur + 0.0f;
    for (int d = 0; d < nd; ++d) out[2 + d] = P[2 + d] + nrm[d] * corr;
  } else {
    for (int d = 0; d < nd; ++d) out[2 + d] = lt * P[2 + d] + ge * b->pinf[2 + d];
  }
}

/* one boundary family: P = state2primitive(Q); impose_bc!(dom, name, P) do b, Pi; bc(Pi, b.normals) end;
 * Q[ghosts] = primitive2state(P[ghosts]) -- all image reads before any ghost write (chunks = entries of `bs`) */
int ibxref_ghost_update(int nb, const ibxref_bdry* bs, int nd, float Rg, float gamma, int64_t N, float* Q, int nthreads) {
  const int nv = nd + 2;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  float* P = malloc(sizeof(float) * N * nv);
#pragma omp parallel
  {
    int nt = 1, it = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads(); it = omp_get_thread_num();
#endif
    int64_t lo = N * it / nt, hi = N * (it + 1) / nt;
    /* state2primitive on this thread's row range (column-major: strided by N) */
    for (int64_t i = lo; i < hi; ++i) {
      float rho = Q[i], E = Q[N + i], u[3], k = 0.0f;
      for (int d = 0; d < nd; ++d) {
        u[d] = Q[(2 + d) * N + i] / rho;
        k = d == 0 ? u[0] * u[0] : k + u[d] * u[d];
      }
      k = k / 2.0f;
      float p = (gamma - 1.0f) * (E - rho * k);
      P[i] = p;
      P[N + i] = clampT(p / (rho * Rg));
      for (int d = 0; d < nd; ++d) P[(2 + d) * N + i] = u[d];
    }
  }
  for (int ib = 0; ib < nb; ++ib) {
    const ibxref_bdry* b = bs + ib;
    const int64_t G = b->n_ghost;
    float* val = malloc(sizeof(float) * G * nv);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; ++g) {
      float ia[5], ba[5], nrm[3];
      for (int v = 0; v < nv; ++v) {
        float s = 0.0f;
        for (int64_t j = b->ptr[g]; j < b->ptr[g + 1]; ++j) {
          float t = P[(int64_t)v * N + b->image_domain[b->idx[j]]] * b->w[j];
          s = j == b->ptr[g] ? t : s + t;
        }
        ia[v] = s;
      }
      for (int d = 0; d < nd; ++d) nrm[d] = b->normals[(int64_t)d * G + g];
      flowbc_point(b, nd, Rg, gamma, ia, nrm, ba);
      float e = b->eta[g];
      for (int v = 0; v < nv; ++v) val[(int64_t)v * G + g] = e * ia[v] + (1.0f - e) * ba[v];
    }
    /* Jacobi within the family: the caller passes all chunks of one family in one call; values are written after
     * every chunk of the family has read (chunks of one family have disjoint ghosts) */
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < G; ++g) {
      float pg[5], qg[5];
      for (int v = 0; v < nv; ++v) pg[v] = val[(int64_t)v * G + g];
      p2s_point(Rg, gamma, nd, pg, qg);
      for (int v = 0; v < nv; ++v) Q[(int64_t)v * N + b->ghost[g]] = qg[v];
    }
    free(val);
  }
  free(P);
  return 0;
}
