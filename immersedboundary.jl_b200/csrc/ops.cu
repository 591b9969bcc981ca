// Table-driven per-operator kernels: the compatibility path behind arbitrary `dom(f, args...)` closures.
// One kernel per reference operator (src/ImmersedBoundary.jl:879-1157, src/accumulator.jl:78-130).
// This translation unit is compiled with -fmad=false: every product and sum rounds separately, in the
// reference's order, so that results are bit-identical to a Float32 Julia evaluation (IEEE division and
// square root are nvcc defaults).  All kernels are grid-stride, cell/face index fastest (coalesced SoA).
#include "device.cuh"

using namespace ibx;

namespace {

constexpr int TB = 256;

#define GRID(n) grid_for((n), TB, c->sm_count, 32), TB, 0, c->stream

__global__ void k_fill(float* __restrict__ a, int64_t n, float v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = v;
}

// out[i, col] = in[idx[i], col]
__global__ void k_gather_rows(const float* __restrict__ in, int64_t in_rows, const int32_t* __restrict__ idx,
                              float* __restrict__ out, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    out[t] = in[col * in_rows + idx[i]];
  }
}

// out[start + i, col] = in[idx[i], col]
__global__ void k_scatter_image(const float* __restrict__ in, int64_t in_rows, const int32_t* __restrict__ idx,
                                float* __restrict__ out, int64_t out_rows, int64_t start, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    out[col * out_rows + start + i] = in[col * in_rows + idx[i]];
  }
}

// at_faces: (uo * spn + un * spo) / (spn + spo)
__global__ void k_at_faces(const float* __restrict__ u, int64_t nrows, const float* __restrict__ sp,
                           const int32_t* __restrict__ own, const int32_t* __restrict__ nei, float* __restrict__ out,
                           int64_t nf, int cols) {
  int64_t tot = nf * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t f = t % nf, col = t / nf;
    int32_t o = own[f], n = nei[f];
    float spo = sp[o], spn = sp[n];
    float uo = u[col * nrows + o], un = u[col * nrows + n];
    out[t] = (uo * spn + un * spo) / (spn + spo);
  }
}

// mode 0: (spo + spn) / 2; 1: spo / 2; 2: spn / 2
__global__ void k_distance(const float* __restrict__ sp, const int32_t* __restrict__ own,
                           const int32_t* __restrict__ nei, float* __restrict__ out, int64_t nf, int mode) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    float spo = sp[own[f]], spn = sp[nei[f]];
    out[f] = mode == 0 ? (spo + spn) / 2.0f : (mode == 1 ? spo / 2.0f : spn / 2.0f);
  }
}

__global__ void k_face_gradient(const float* __restrict__ u, int64_t nrows, const float* __restrict__ sp,
                                const int32_t* __restrict__ own, const int32_t* __restrict__ nei,
                                float* __restrict__ out, int64_t nf, int cols) {
  int64_t tot = nf * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t f = t % nf, col = t / nf;
    int32_t o = own[f], n = nei[f];
    float fd = (sp[o] + sp[n]) / 2.0f;
    out[t] = (u[col * nrows + n] - u[col * nrows + o]) / fd;
  }
}

// weighted mean over a CSR face list: products first, then top-to-bottom sum (accumulator.jl:95-106)
__device__ __forceinline__ float list_mean(const float* __restrict__ uf, const int32_t* __restrict__ ptr,
                                           const int32_t* __restrict__ idx, int64_t cell) {
  int32_t b = ptr[cell], e = ptr[cell + 1];
  if (e == b) return 0.0f;
  float w = 1.0f / (float)(e - b);
  float acc = uf[idx[b]] * w;
  for (int32_t k = b + 1; k < e; ++k) acc = acc + uf[idx[k]] * w;
  return acc;
}

// green_gauss (sign = -1) / unsigned_green_gauss (sign = +1): (accr(uf) -+ accl(uf)) / spacing
__global__ void k_green_gauss(const float* __restrict__ uf, int64_t nf, const float* __restrict__ sp,
                              const int32_t* __restrict__ lptr, const int32_t* __restrict__ lidx,
                              const int32_t* __restrict__ rptr, const int32_t* __restrict__ ridx,
                              float* __restrict__ out, int64_t n, int cols, int unsigned_) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t % n, col = t / n;
    const float* ufc = uf + col * nf;
    float r = list_mean(ufc, rptr, ridx, cell), l = list_mean(ufc, lptr, lidx, cell);
    out[t] = (unsigned_ ? (r + l) : (r - l)) / sp[cell];
  }
}

// Float64 variant: what Julia does when `uf` is the Float64 HLL flux (weights and spacing promote)
__global__ void k_green_gauss64(const double* __restrict__ uf, int64_t nf, const float* __restrict__ sp,
                                const int32_t* __restrict__ lptr, const int32_t* __restrict__ lidx,
                                const int32_t* __restrict__ rptr, const int32_t* __restrict__ ridx,
                                double* __restrict__ out, int64_t n, int cols, int unsigned_) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t % n, col = t / n;
    const double* ufc = uf + col * nf;
    double side[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int32_t* ptr = s ? rptr : lptr;
      const int32_t* idx = s ? ridx : lidx;
      int32_t b = ptr[cell], e = ptr[cell + 1];
      double acc = 0.0;
      if (e > b) {
        double w = (double)(1.0f / (float)(e - b));
        acc = ufc[idx[b]] * w;
        for (int32_t k = b + 1; k < e; ++k) acc = acc + ufc[idx[k]] * w;
      }
      side[s] = acc;
    }
    out[t] = (unsigned_ ? (side[1] + side[0]) : (side[1] - side[0])) / (double)sp[cell];
  }
}

struct AnyPtr {
  void* p;
  int f64;
};
__device__ __forceinline__ double ldany(AnyPtr a, int64_t i) { return a.f64 ? ((const double*)a.p)[i] : (double)((const float*)a.p)[i]; }
__device__ __forceinline__ void stany(AnyPtr a, int64_t i, double v) {
  if (a.f64) ((double*)a.p)[i] = v; else ((float*)a.p)[i] = (float)v;
}
__device__ __forceinline__ double binop64(int op, double x, double y) {
  switch (op) {
    case 0: return x + y;
    case 1: return x - y;
    case 2: return x * y;
    case 3: return x / y;
    case 4: return fmax(x, y);
    default: return fmin(x, y);
  }
}
// elementwise with at least one Float64 operand: Julia promotes to Float64, then converts on assignment
__global__ void k_ew_binary_any(int op, AnyPtr a, int acols, AnyPtr b, int bcols, AnyPtr out, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    stany(out, t, binop64(op, ldany(a, (acols == 1 ? 0 : col) * n + i), ldany(b, (bcols == 1 ? 0 : col) * n + i)));
  }
}
__global__ void k_ew_scalar_any(int op, AnyPtr a, double s, int scalar_first, AnyPtr out, int64_t tot) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x)
    stany(out, t, scalar_first ? binop64(op, s, ldany(a, t)) : binop64(op, ldany(a, t), s));
}
__global__ void k_ew_unary_any(int op, AnyPtr a, AnyPtr out, int64_t tot) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    double x = ldany(a, t), r;
    switch (op) {
      case 0: r = fabs(x); break;
      case 1: r = -x; break;
      case 2: r = sqrt(x); break;
      case 3: r = (double)((x > 0.0) - (x < 0.0)); break;
      default: r = 1.0 / x; break;
    }
    stany(out, t, r);
  }
}

__device__ __forceinline__ float face_value(const float* __restrict__ u, const float* __restrict__ sp,
                                            const int32_t* __restrict__ own, const int32_t* __restrict__ nei, int32_t f) {
  int32_t o = own[f], n = nei[f];
  float spo = sp[o], spn = sp[n];
  return (u[o] * spn + u[n] * spo) / (spn + spo);
}

// cell_gradient = green_gauss(at_faces(u)) fused (no face temporary)
__global__ void k_cell_gradient(const float* __restrict__ u, const float* __restrict__ sp,
                                const int32_t* __restrict__ own, const int32_t* __restrict__ nei,
                                const int32_t* __restrict__ lptr, const int32_t* __restrict__ lidx,
                                const int32_t* __restrict__ rptr, const int32_t* __restrict__ ridx,
                                float* __restrict__ out, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t % n, col = t / n;
    const float* uc = u + col * n;
    float side[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int32_t* ptr = s ? rptr : lptr;
      const int32_t* idx = s ? ridx : lidx;
      int32_t b = ptr[cell], e = ptr[cell + 1];
      float acc = 0.0f;
      if (e > b) {
        float w = 1.0f / (float)(e - b);
        acc = face_value(uc, sp, own, nei, idx[b]) * w;
        for (int32_t k = b + 1; k < e; ++k) acc = acc + face_value(uc, sp, own, nei, idx[k]) * w;
      }
      side[s] = acc;
    }
    out[t] = (side[1] - side[0]) / sp[cell];
  }
}

struct DimTables {
  const float* sp;
  const int32_t *own, *nei, *lptr, *lidx, *rptr, *ridx;
};
struct AllDims {
  DimTables d[3];
  int nd;
};

__device__ __forceinline__ float jst_dim(const float* __restrict__ p, const DimTables& T, int64_t cell) {
  float g[2], a[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int32_t* ptr = s ? T.rptr : T.lptr;
    const int32_t* idx = s ? T.ridx : T.lidx;
    int32_t b = ptr[cell], e = ptr[cell + 1];
    float accg = 0.0f, acca = 0.0f;
    if (e > b) {
      float w = 1.0f / (float)(e - b);
      for (int32_t k = b; k < e; ++k) {
        int32_t f = idx[k];
        float fd = p[T.nei[f]] - p[T.own[f]];
        float tg = fd * w, ta = fabsf(fd) * w;
        accg = k == b ? tg : accg + tg;
        acca = k == b ? ta : acca + ta;
      }
    }
    g[s] = accg;
    a[s] = acca;
  }
  float sp = T.sp[cell];
  float gg = (g[1] - g[0]) / sp;
  float ugg = (a[1] + a[0]) / sp;
  return (1e-7f + fabsf(gg)) / (1e-7f + ugg);
}

// JST_sensor(part, p, dim) (src/ImmersedBoundary.jl:1077-1097); dim < 0: max over dims with floor 1e-7
__global__ void k_jst(const float* __restrict__ p, AllDims A, int dim, float* __restrict__ out, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t % n, col = t / n;
    const float* pc = p + col * n;
    float nu;
    if (dim >= 0) {
      nu = jst_dim(pc, A.d[dim], cell);
    } else {
      nu = 1e-7f;
      for (int d = 0; d < A.nd; ++d) nu = fmaxf(nu, jst_dim(pc, A.d[d], cell));
    }
    out[t] = nu;
  }
}

__device__ __forceinline__ float sgn(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

// MUSCL (src/ImmersedBoundary.jl:1113-1157)
__global__ void k_muscl(const float* __restrict__ u, const float* __restrict__ du, const float* __restrict__ D,
                        int64_t nrows, const float* __restrict__ sp, const int32_t* __restrict__ own,
                        const int32_t* __restrict__ nei, float* __restrict__ uL, float* __restrict__ uR, int64_t nf,
                        int cols, int high_order) {
  int64_t tot = nf * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t f = t % nf, col = t / nf;
    int32_t o = own[f], n = nei[f];
    float down = sp[o] / 2.0f, dnei = sp[n] / 2.0f;
    float uo = u[col * nrows + o], un = u[col * nrows + n];
    float duo = du[col * nrows + o], dun = du[col * nrows + n];
    float gf = (un - uo) / (down + dnei);
    float gu = (2.0f * duo - gf) * down;
    float Du = (2.0f * dun - gf) * dnei;
    float s = fminf(fabsf(Du), fabsf(gu)) * (sgn(Du) + sgn(gu)) / 2.0f;
    float l = uo + s, r = un - s;
    if (D) {
      float Df = fmaxf(fmaxf(D[o], D[n]), 1e-7f);
      float uf = (uo * dnei + un * down) / (down + dnei);
      if (high_order) uf = uf + (duo * down - dun * dnei) / 8.0f;
      l = l * Df + (1.0f - Df) * uf;
      r = r * Df + (1.0f - Df) * uf;
    }
    uL[t] = l;
    uR[t] = r;
  }
}

// ------------------------------------------------------------------ accumulators / IB
// out[row, col] = sum_k w[k] * (v[idx[k], col] (- v[row, col])) over CSR row; unweighted: plain sum
__global__ void k_accumulate(const float* __restrict__ v, int64_t vrows, const int32_t* __restrict__ ptr,
                             const int32_t* __restrict__ idx, const float* __restrict__ w, float* __restrict__ out,
                             int64_t n, int cols, int delta) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = t % n, col = t / n;
    const float* vc = v + col * vrows;
    int32_t b = ptr[row], e = ptr[row + 1];
    float acc = 0.0f;
    float self = (delta && w) ? vc[row] : 0.0f;
    for (int32_t k = b; k < e; ++k) {
      float val = vc[idx[k]];
      if (w) {
        if (delta) val = val - self;
        val = val * w[k];
      }
      acc = k == b ? val : acc + val;
    }
    out[t] = acc;
  }
}

// the same with the `f` and `op` keyword arguments of src/accumulator.jl:78-111: f applied to the gathered values (after the
// Delta subtraction, before the weights), rows reduced with op in list order.  F: 0 identity, 1 abs, 2 square, 3 sign.
// OP: 0 +, 1 max, 2 min, 3 *.  An empty row keeps the zero of `vnew .= 0`.
template <int F, int OP>
__global__ void k_accumulate_ex(const float* __restrict__ v, int64_t vrows, const int32_t* __restrict__ ptr,
                                const int32_t* __restrict__ idx, const float* __restrict__ w, float* __restrict__ out,
                                int64_t n, int cols, int delta) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = t % n, col = t / n;
    const float* vc = v + col * vrows;
    int32_t b = ptr[row], e = ptr[row + 1];
    float acc = 0.0f;
    float self = (delta && w) ? vc[row] : 0.0f;
    for (int32_t k = b; k < e; ++k) {
      float val = vc[idx[k]];
      if (w && delta) val = val - self;
      if (F == 1) val = fabsf(val);
      if (F == 2) val = val * val;
      if (F == 3) val = (float)((val > 0.0f) - (val < 0.0f));
      if (w) val = val * w[k];
      if (k == b) acc = val;
      else if (OP == 0) acc = acc + val;
      else if (OP == 1) acc = fmaxf(acc, val);
      else if (OP == 2) acc = fminf(acc, val);
      else acc = acc * val;
    }
    out[t] = acc;
  }
}

// a[ghost, col] = eta * ia + (1 - eta) * ba   (src/ImmersedBoundary.jl:1242-1245); ba array or scalar
__global__ void k_bc_blend(float* __restrict__ a, int64_t arows, const int32_t* __restrict__ ghost,
                           const float* __restrict__ eta, const float* __restrict__ ia, const float* __restrict__ ba,
                           float ba_scalar, int64_t G, int cols) {
  int64_t tot = G * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = t % G, col = t / G;
    float e = eta[g];
    float b = ba ? ba[t] : ba_scalar;
    a[col * arows + ghost[g]] = e * ia[t] + (1.0f - e) * b;
  }
}

// ------------------------------------------------------------------ elementwise glue
__device__ __forceinline__ float binop(int op, float x, float y) {
  switch (op) {
    case 0: return x + y;
    case 1: return x - y;
    case 2: return x * y;
    case 3: return x / y;
    case 4: return fmaxf(x, y);
    default: return fminf(x, y);
  }
}

__global__ void k_ew_binary(int op, const float* __restrict__ a, int acols, const float* __restrict__ b, int bcols,
                            float* __restrict__ out, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    float x = a[(acols == 1 ? 0 : col) * n + i], y = b[(bcols == 1 ? 0 : col) * n + i];
    out[t] = binop(op, x, y);
  }
}

__global__ void k_ew_scalar(int op, const float* __restrict__ a, float s, int scalar_first, float* __restrict__ out,
                            int64_t tot) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x)
    out[t] = scalar_first ? binop(op, s, a[t]) : binop(op, a[t], s);
}

__global__ void k_ew_unary(int op, const float* __restrict__ a, float* __restrict__ out, int64_t tot) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    float x = a[t];
    float r;
    switch (op) {
      case 0: r = fabsf(x); break;
      case 1: r = -x; break;
      case 2: r = sqrtf(x); break;
      case 3: r = sgn(x); break;
      default: r = 1.0f / x; break;
    }
    out[t] = r;
  }
}

__global__ void k_axpy(float alpha, const float* __restrict__ x, float* __restrict__ y, int64_t tot) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x)
    y[t] = y[t] + alpha * x[t];
}

__global__ void k_clamped_update(float* __restrict__ Q, const float* __restrict__ om, int omcols,
                                 const float* __restrict__ r, int64_t n, int cols) {
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n, col = t / n;
    float w = fminf(fmaxf(om[(omcols == 1 ? 0 : col) * n + i], 0.0f), 1.0f);
    Q[t] = Q[t] + w * r[t];
  }
}

// one stage of a local-time-step march: Q = Q0 + ((alpha / cfl_i) * R) * mask_i, in that order of roundings
__global__ void k_local_step(const float* Q0, const float* __restrict__ R, const float* __restrict__ cfl,
                             const float* __restrict__ mask, float alpha, float* Q, int64_t n, int cols) {   // Q may be Q0
  int64_t tot = n * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t % n;
    float w = alpha / cfl[i];
    float d = w * R[t];
    if (mask) d = d * mask[i];
    Q[t] = Q0[t] + d;
  }
}

// ------------------------------------------------------------------ reductions: warp shuffle -> block -> one slot per block
__device__ __forceinline__ double red_combine(int op, double a, double b) {
  switch (op) {
    // Julia's maximum / minimum propagate NaN (fmax / fmin would drop it and hide a diverged field from the solver loops)
    case 1: case 3: return (a != a || b != b) ? (double)NAN : fmax(a, b);
    case 2: return (a != a || b != b) ? (double)NAN : fmin(a, b);
    default: return a + b;
  }
}
__device__ __forceinline__ double red_identity(int op) {
  switch (op) {
    case 1: return -INFINITY;
    case 2: return INFINITY;
    case 3: return 0.0;
    default: return 0.0;
  }
}

// one column per blockIdx.y; partial results to part[col * gridDim.x + blockIdx.x]
__global__ void k_reduce(int op, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ wgt,
                         int64_t n, double* __restrict__ part) {
  int col = blockIdx.y;
  const float* ac = a + (int64_t)col * n;
  const float* bc = b ? b + (int64_t)col * n : nullptr;
  double acc = red_identity(op);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double x = ac[i];
    double v;
    switch (op) {
      case 3: v = fabs(x); break;
      case 4: v = x * x; break;
      case 5: v = x * (double)bc[i]; break;           // dot
      case 6: v = (double)(ac[i] * wgt[i]); break;    // weighted sum with float product (volume / surface integrals)
      default: v = x; break;
    }
    acc = red_combine(op, acc, v);
  }
  for (int off = 16; off > 0; off >>= 1) acc = red_combine(op, acc, __shfl_xor_sync(0xffffffffu, acc, off));
  __shared__ double sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (blockDim.x >> 5) ? sm[lane] : red_identity(op);
    for (int off = 16; off > 0; off >>= 1) acc = red_combine(op, acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (lane == 0) part[(int64_t)col * gridDim.x + blockIdx.x] = acc;
  }
}

__global__ void k_prod_spacing(const float* __restrict__ w, int64_t n, int nd, float* __restrict__ vol) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = w[i];
    for (int d = 1; d < nd; ++d) v = v * w[(int64_t)d * n + i];
    vol[i] = v;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------ helpers
static int reduce_impl(ibx_ctx* c, int op, const float* a, const float* b, const float* w, int64_t n, int cols,
                       double* out_per_col) {
  int gx = grid_for(n, TB, c->sm_count, 4);
  if ((int64_t)gx * cols > c->red_cap) gx = (int)std::max<int64_t>(1, c->red_cap / cols);
  dim3 grid(gx, cols);
  k_reduce<<<grid, TB, 0, c->stream>>>(op, a, b, w, n, c->d_red);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(c->h_red, c->d_red, (size_t)gx * cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int rop = (op >= 4) ? 0 : op;
  for (int col = 0; col < cols; ++col) {
    double acc = rop == 1 || rop == 3 ? (rop == 1 ? -INFINITY : 0.0) : (rop == 2 ? INFINITY : 0.0);
    for (int k = 0; k < gx; ++k) {
      double v = c->h_red[(size_t)col * gx + k];
      if (rop != 0 && (v != v || acc != acc)) acc = NAN;   // NaN propagates through max / min like Julia's maximum / minimum
      else acc = (rop == 1 || rop == 3) ? std::max(acc, v) : (rop == 2 ? std::min(acc, v) : acc + v);
    }
    out_per_col[col] = acc;
  }
  return IBX_OK;
}

#define PART(P, D, p)                                                                                      \
  if ((p) < 0 || (p) >= (int)(D).parts.size()) return fail(IBX_ERR_ARG, std::string(__func__) + ": partition index out of range"); \
  PartitionT& P = (D).parts[p]
#define DIMCHK(D, dim) \
  if ((dim) < 0 || (dim) >= (D).nd) return fail(IBX_ERR_ARG, std::string(__func__) + ": dim out of range")
#define SHAPE(cond, msg) \
  if (!(cond)) return fail(IBX_ERR_ARG, std::string(__func__) + ": shape mismatch: " + (msg))

static AllDims all_dims(const ibx_domain& D, const PartitionT& P) {
  AllDims A;
  A.nd = D.nd;
  int64_t n = (int64_t)P.domain.size();
  for (int d = 0; d < D.nd; ++d) {
    const FaceTable& T = P.dims[d];
    A.d[d] = {P.d_spacing + (int64_t)d * n, T.d_owners, T.d_neighbors, T.d_lptr, T.d_lidx, T.d_rptr, T.d_ridx};
  }
  return A;
}

extern "C" {

int ibx_array_fill(ibx_ctx* c, ibx_array a, float v) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  int64_t n = A.rows * A.cols;
  k_fill<<<GRID(n)>>>(A.p, n, v);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_array_column(ibx_ctx* c, ibx_array a, int col, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  GET_ARR(O, out);
  SHAPE(col >= 0 && col < A.cols && O.rows == A.rows && O.cols == 1, "column out of range or output not rows x 1");
  CU(cudaMemcpyAsync(O.p, A.p + (int64_t)col * A.rows, (size_t)A.rows * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  return IBX_OK;
}

int ibx_array_set_column(ibx_ctx* c, ibx_array a, int col, ibx_array src) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  GET_ARR(S, src);
  SHAPE(col >= 0 && col < A.cols && S.rows == A.rows && S.cols == 1, "column out of range or source not rows x 1");
  CU(cudaMemcpyAsync(A.p + (int64_t)col * A.rows, S.p, (size_t)A.rows * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  return IBX_OK;
}

int ibx_gather_domain(ibx_ctx* c, const ibx_domain* d, int p, ibx_array global_in, ibx_array local_out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  GET_ARR(G, global_in);
  GET_ARR(L, local_out);
  int64_t n = (int64_t)P.domain.size();
  SHAPE(G.rows == D.ncells && L.rows == n && L.cols == G.cols, "global must be ncells x nv, local n_domain x nv");
  k_gather_rows<<<GRID(n * G.cols)>>>(G.p, G.rows, P.d_domain, L.p, n, (int)G.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_scatter_image(ibx_ctx* c, const ibx_domain* d, int p, ibx_array local_in, ibx_array global_out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  GET_ARR(L, local_in);
  GET_ARR(G, global_out);
  SHAPE(G.rows == D.ncells && L.rows == (int64_t)P.domain.size() && L.cols == G.cols, "global must be ncells x nv, local n_domain x nv");
  k_scatter_image<<<GRID(P.n_image * G.cols)>>>(L.p, L.rows, P.d_image_in_domain, G.p, G.rows, P.image_start, P.n_image, (int)G.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

static int copy_table(ibx_ctx* c, const float* src, int64_t rows, int cols, ibx_array out, const char* fn) {
  GET_ARR(O, out);
  if (O.rows != rows || O.cols != cols) return fail(IBX_ERR_ARG, std::string(fn) + ": output must be n_domain x nd");
  CU(cudaMemcpyAsync(O.p, src, (size_t)rows * cols * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  return IBX_OK;
}

int ibx_partition_spacing(ibx_ctx* c, const ibx_domain* d, int p, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  return copy_table(c, P.d_spacing, (int64_t)P.domain.size(), D.nd, out, __func__);
}

int ibx_partition_centers(ibx_ctx* c, const ibx_domain* d, int p, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  return copy_table(c, P.d_centers, (int64_t)P.domain.size(), D.nd, out, __func__);
}

static int gather_faces(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out, bool neigh,
                        const char* fn) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR(U, u);
  GET_ARR(O, out);
  const FaceTable& T = P.dims[dim];
  int64_t nf = (int64_t)T.owners.size();
  if (U.rows != (int64_t)P.domain.size() || O.rows != nf || O.cols != U.cols)
    return fail(IBX_ERR_ARG, std::string(fn) + ": u must be n_domain x nv and out nfaces x nv");
  k_gather_rows<<<GRID(nf * U.cols)>>>(U.p, U.rows, neigh ? T.d_neighbors : T.d_owners, O.p, nf, (int)U.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_at_owners(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out) {
  return gather_faces(c, d, p, dim, u, out, false, __func__);
}
int ibx_at_neighbors(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out) {
  return gather_faces(c, d, p, dim, u, out, true, __func__);
}

int ibx_at_faces(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR(U, u);
  GET_ARR(O, out);
  const FaceTable& T = P.dims[dim];
  int64_t n = (int64_t)P.domain.size(), nf = (int64_t)T.owners.size();
  SHAPE(U.rows == n && O.rows == nf && O.cols == U.cols, "u n_domain x nv, out nfaces x nv");
  k_at_faces<<<GRID(nf * U.cols)>>>(U.p, n, P.d_spacing + (int64_t)dim * n, T.d_owners, T.d_neighbors, O.p, nf, (int)U.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

static int gg_impl(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array uf, ibx_array out, int uns, const char* fn) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR_ANY(U, uf);
  GET_ARR_ANY(O, out);
  const FaceTable& T = P.dims[dim];
  int64_t n = (int64_t)P.domain.size(), nf = (int64_t)T.owners.size();
  if (U.rows != nf || O.rows != n || O.cols != U.cols)
    return fail(IBX_ERR_ARG, std::string(fn) + ": uf must be nfaces x nv and out n_domain x nv");
  if (U.f64 != O.f64) return fail(IBX_ERR_ARG, std::string(fn) + ": uf and out must have the same element type");
  if (U.f64) {
    k_green_gauss64<<<GRID(n * U.cols)>>>((const double*)U.p, nf, P.d_spacing + (int64_t)dim * n, T.d_lptr, T.d_lidx, T.d_rptr, T.d_ridx,
                                          (double*)O.p, n, (int)U.cols, uns);
    LAUNCH_CHECK();
    return IBX_OK;
  }
  k_green_gauss<<<GRID(n * U.cols)>>>(U.p, nf, P.d_spacing + (int64_t)dim * n, T.d_lptr, T.d_lidx, T.d_rptr, T.d_ridx, O.p, n, (int)U.cols, uns);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_green_gauss(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array uf, ibx_array out) {
  return gg_impl(c, d, p, dim, uf, out, 0, __func__);
}
int ibx_unsigned_green_gauss(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array uf, ibx_array out) {
  return gg_impl(c, d, p, dim, uf, out, 1, __func__);
}

int ibx_cell_gradient(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR(U, u);
  GET_ARR(O, out);
  const FaceTable& T = P.dims[dim];
  int64_t n = (int64_t)P.domain.size();
  SHAPE(U.rows == n && O.rows == n && O.cols == U.cols, "u and out must be n_domain x nv");
  k_cell_gradient<<<GRID(n * U.cols)>>>(U.p, P.d_spacing + (int64_t)dim * n, T.d_owners, T.d_neighbors, T.d_lptr, T.d_lidx, T.d_rptr, T.d_ridx, O.p, n, (int)U.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

static int dist_impl(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out, int mode, const char* fn) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR(O, out);
  const FaceTable& T = P.dims[dim];
  int64_t n = (int64_t)P.domain.size(), nf = (int64_t)T.owners.size();
  if (O.rows != nf || O.cols != 1) return fail(IBX_ERR_ARG, std::string(fn) + ": out must be nfaces x 1");
  k_distance<<<GRID(nf)>>>(P.d_spacing + (int64_t)dim * n, T.d_owners, T.d_neighbors, O.p, nf, mode);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_face_distance(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out) { return dist_impl(c, d, p, dim, out, 0, __func__); }
int ibx_owner_distance(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out) { return dist_impl(c, d, p, dim, out, 1, __func__); }
int ibx_neighbor_distance(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out) { return dist_impl(c, d, p, dim, out, 2, __func__); }

int ibx_face_gradient(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR(U, u);
  GET_ARR(O, out);
  const FaceTable& T = P.dims[dim];
  int64_t n = (int64_t)P.domain.size(), nf = (int64_t)T.owners.size();
  SHAPE(U.rows == n && O.rows == nf && O.cols == U.cols, "u n_domain x nv, out nfaces x nv");
  k_face_gradient<<<GRID(nf * U.cols)>>>(U.p, n, P.d_spacing + (int64_t)dim * n, T.d_owners, T.d_neighbors, O.p, nf, (int)U.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_jst_sensor(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array pr, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  if (dim >= D.nd) return fail(IBX_ERR_ARG, "ibx_jst_sensor: dim out of range");
  GET_ARR(U, pr);
  GET_ARR(O, out);
  int64_t n = (int64_t)P.domain.size();
  SHAPE(U.rows == n && O.rows == n && O.cols == U.cols, "p and out must be n_domain x nv");
  k_jst<<<GRID(n * U.cols)>>>(U.p, all_dims(D, P), dim, O.p, n, (int)U.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_muscl(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array du, ibx_array Dh, int high_order,
              ibx_array uL, ibx_array uR) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  PART(P, D, p);
  DIMCHK(D, dim);
  GET_ARR(U, u);
  GET_ARR(DU, du);
  GET_ARR(L, uL);
  GET_ARR(R, uR);
  const float* Dp = nullptr;
  if (Dh) {
    GET_ARR(DD, Dh);
    SHAPE(DD.rows == U.rows && DD.cols == 1, "D must be a vector of n_domain entries (Union{AbstractVector, Nothing})");
    Dp = DD.p;
  }
  const FaceTable& T = P.dims[dim];
  int64_t n = (int64_t)P.domain.size(), nf = (int64_t)T.owners.size();
  SHAPE(U.rows == n && DU.rows == n && DU.cols == U.cols && L.rows == nf && R.rows == nf && L.cols == U.cols && R.cols == U.cols,
        "u, du n_domain x nv; uL, uR nfaces x nv");
  k_muscl<<<GRID(nf * U.cols)>>>(U.p, DU.p, Dp, n, P.d_spacing + (int64_t)dim * n, T.d_owners, T.d_neighbors, L.p, R.p, nf, (int)U.cols, high_order);
  LAUNCH_CHECK();
  return IBX_OK;
}

// ------------------------------------------------------------------ accumulators / IB ghost update
static int accumulate_raw(ibx_ctx* c, const int32_t* ptr, const int32_t* idx, const float* w, int64_t n_out,
                          const ibx_ctx::Arr& V, const ibx_ctx::Arr& O, int delta) {
  k_accumulate<<<GRID(n_out * V.cols)>>>(V.p, V.rows, ptr, idx, w, O.p, n_out, (int)V.cols, delta);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_accumulate(ibx_ctx* c, const ibx_accum* a, ibx_array v, int delta, ibx_array out) {
  CHECK_CTX(c);
  ibx_accum* A = find_accum(a);
  if (!A) return fail(IBX_ERR_ARG, "ibx_accumulate: unknown accumulator handle");
  if (!A->uploaded) return fail(IBX_ERR_STATE, "ibx_accumulate: accumulator not uploaded (ibx_accum_upload)");
  GET_ARR(V, v);
  GET_ARR(O, out);
  SHAPE(O.rows == A->n_out && O.cols == V.cols, "out must be n_output x nv");
  if (delta) SHAPE(V.rows >= A->n_out, "delta form needs v[row] for every output row");
  return accumulate_raw(c, A->d_ptr, A->d_idx, A->weighted ? A->d_w : nullptr, A->n_out, V, O, delta);
}

int ibx_accumulate_ex(ibx_ctx* c, const ibx_accum* a, ibx_array v, int delta, int f_kind, int op_kind, ibx_array out) {
  CHECK_CTX(c);
  ibx_accum* A = find_accum(a);
  if (!A) return fail(IBX_ERR_ARG, "ibx_accumulate_ex: unknown accumulator handle");
  if (!A->uploaded) return fail(IBX_ERR_STATE, "ibx_accumulate_ex: accumulator not uploaded (ibx_accum_upload)");
  if (f_kind < 0 || f_kind > 3 || op_kind < 0 || op_kind > 3)
    return fail(IBX_ERR_ARG, "ibx_accumulate_ex: f_kind in 0..3 (identity, abs, square, sign), op_kind in 0..3 (+, max, min, *)");
  if (f_kind == 0 && op_kind == 0) return ibx_accumulate(c, a, v, delta, out);
  GET_ARR(V, v);
  GET_ARR(O, out);
  SHAPE(O.rows == A->n_out && O.cols == V.cols, "out must be n_output x nv");
  if (delta) SHAPE(V.rows >= A->n_out, "delta form needs v[row] for every output row");
  const float* w = A->weighted ? A->d_w : nullptr;
#define GO(F, OP) k_accumulate_ex<F, OP><<<GRID(A->n_out * V.cols)>>>(V.p, V.rows, A->d_ptr, A->d_idx, w, O.p, A->n_out, (int)V.cols, delta)
#define ROW(F) switch (op_kind) { case 0: GO(F, 0); break; case 1: GO(F, 1); break; case 2: GO(F, 2); break; default: GO(F, 3); }
  switch (f_kind) { case 0: ROW(0) break; case 1: ROW(1) break; case 2: ROW(2) break; default: ROW(3) }
#undef ROW
#undef GO
  LAUNCH_CHECK();
  return IBX_OK;
}

#define BDRY(B, D, b, part)                                                                                    \
  if ((b) < 0 || (b) >= (int)(D).boundaries.size()) return fail(IBX_ERR_ARG, std::string(__func__) + ": boundary index out of range"); \
  if ((part) < 0 || (part) >= (int)(D).boundaries[b].parts.size()) return fail(IBX_ERR_ARG, std::string(__func__) + ": boundary partition out of range"); \
  BoundaryT& B = (D).boundaries[b].parts[part]

int ibx_bc_image_values(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  BDRY(B, D, b, part);
  GET_ARR(A, a);
  GET_ARR(I, ia);
  int64_t G = (int64_t)B.ghost.size();
  SHAPE(A.rows == D.ncells && I.rows == G && I.cols == A.cols, "a ncells x nv, ia nghost x nv");
  return accumulate_raw(c, B.d_ptr, B.d_idx_global, B.d_w, G, A, I, 0);
}

int ibx_bc_normals(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  BDRY(B, D, b, part);
  return copy_table(c, B.d_normals, (int64_t)B.ghost.size(), D.nd, out, __func__);
}

static int blend_impl(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia, const float* ba,
                      int64_t ba_rows, int ba_cols, float scalar, const char* fn) {
  GET_DOM(D, d);
  BDRY(B, D, b, part);
  GET_ARR(A, a);
  GET_ARR(I, ia);
  int64_t G = (int64_t)B.ghost.size();
  if (A.rows != D.ncells || I.rows != G || I.cols != A.cols || (ba && (ba_rows != G || ba_cols != A.cols)))
    return fail(IBX_ERR_ARG, std::string(fn) + ": a ncells x nv; ia, ba nghost x nv");
  k_bc_blend<<<GRID(G * A.cols)>>>(A.p, A.rows, B.d_ghost, B.d_eta, I.p, ba, scalar, G, (int)A.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_bc_blend(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia, ibx_array ba) {
  CHECK_CTX(c);
  GET_ARR(BA, ba);
  return blend_impl(c, d, b, part, a, ia, BA.p, BA.rows, (int)BA.cols, 0.0f, __func__);
}

int ibx_bc_blend_scalar(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia, float ba) {
  CHECK_CTX(c);
  return blend_impl(c, d, b, part, a, ia, nullptr, 0, 0, ba, __func__);
}

int ibx_surface_values(ibx_ctx* c, const ibx_domain* d, int s, int offset, ibx_array u, ibx_array out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  if (s < 0 || s >= (int)D.surfaces.size()) return fail(IBX_ERR_ARG, "ibx_surface_values: surface index out of range");
  ibx_accum& A = offset ? D.surfaces[s].offset_interp : D.surfaces[s].interp;
  GET_ARR(U, u);
  GET_ARR(O, out);
  SHAPE(U.rows == D.ncells && O.rows == A.n_out && O.cols == U.cols, "u ncells x nv, out npoints x nv");
  return accumulate_raw(c, A.d_ptr, A.d_idx, A.d_w, A.n_out, U, O, 0);
}

int ibx_surface_integral(ibx_ctx* c, const ibx_domain* d, int s, ibx_array u, float* out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  if (s < 0 || s >= (int)D.surfaces.size()) return fail(IBX_ERR_ARG, "ibx_surface_integral: surface index out of range");
  GET_ARR(U, u);
  int64_t n = (int64_t)D.surfaces[s].areas.size();
  SHAPE(U.rows == n, "u must have one row per surface point");
  std::vector<double> r(U.cols);
  std::vector<double> tmp(1);
  for (int col = 0; col < U.cols; ++col) {
    int rc = reduce_impl(c, 6, U.p + (int64_t)col * n, nullptr, D.surfaces[s].d_areas, n, 1, tmp.data());
    if (rc) return rc;
    out[col] = (float)tmp[0];
  }
  return IBX_OK;
}

// ------------------------------------------------------------------ elementwise + reductions
int ibx_ew_binary(ibx_ctx* c, int op, ibx_array a, ibx_array b, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, a);
  GET_ARR_ANY(B, b);
  GET_ARR_ANY(O, out);
  if (op < 0 || op > 5) return fail(IBX_ERR_ARG, "ibx_ew_binary: unknown op");
  int cols = (int)std::max(A.cols, B.cols);
  SHAPE(A.rows == B.rows && O.rows == A.rows && O.cols == cols && (A.cols == cols || A.cols == 1) && (B.cols == cols || B.cols == 1),
        "operands must have equal rows; a 1-column operand broadcasts over columns");
  if (A.f64 || B.f64 || O.f64) {
    k_ew_binary_any<<<GRID(A.rows * cols)>>>(op, AnyPtr{A.p, A.f64}, (int)A.cols, AnyPtr{B.p, B.f64}, (int)B.cols, AnyPtr{O.p, O.f64}, A.rows, cols);
    LAUNCH_CHECK();
    return IBX_OK;
  }
  k_ew_binary<<<GRID(A.rows * cols)>>>(op, A.p, (int)A.cols, B.p, (int)B.cols, O.p, A.rows, cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_ew_scalar(ibx_ctx* c, int op, ibx_array a, float s, int scalar_first, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, a);
  GET_ARR_ANY(O, out);
  if (op < 0 || op > 5) return fail(IBX_ERR_ARG, "ibx_ew_scalar: unknown op");
  SHAPE(O.rows == A.rows && O.cols == A.cols, "out must match a");
  if (A.f64 || O.f64) {
    k_ew_scalar_any<<<GRID(A.rows * A.cols)>>>(op, AnyPtr{A.p, A.f64}, (double)s, scalar_first, AnyPtr{O.p, O.f64}, A.rows * A.cols);
    LAUNCH_CHECK();
    return IBX_OK;
  }
  k_ew_scalar<<<GRID(A.rows * A.cols)>>>(op, A.p, s, scalar_first, O.p, A.rows * A.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_ew_unary(ibx_ctx* c, int op, ibx_array a, ibx_array out) {
  CHECK_CTX(c);
  GET_ARR_ANY(A, a);
  GET_ARR_ANY(O, out);
  if (op < 0 || op > 4) return fail(IBX_ERR_ARG, "ibx_ew_unary: unknown op");
  SHAPE(O.rows == A.rows && O.cols == A.cols, "out must match a");
  if (A.f64 || O.f64) {
    k_ew_unary_any<<<GRID(A.rows * A.cols)>>>(op, AnyPtr{A.p, A.f64}, AnyPtr{O.p, O.f64}, A.rows * A.cols);
    LAUNCH_CHECK();
    return IBX_OK;
  }
  k_ew_unary<<<GRID(A.rows * A.cols)>>>(op, A.p, O.p, A.rows * A.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_axpy(ibx_ctx* c, float alpha, ibx_array x, ibx_array y) {
  CHECK_CTX(c);
  GET_ARR(X, x);
  GET_ARR(Y, y);
  SHAPE(X.rows == Y.rows && X.cols == Y.cols, "x and y must match");
  k_axpy<<<GRID(X.rows * X.cols)>>>(alpha, X.p, Y.p, X.rows * X.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_clamped_update(ibx_ctx* c, ibx_array Q, ibx_array omega, ibx_array r) {
  CHECK_CTX(c);
  GET_ARR(A, Q);
  GET_ARR(W, omega);
  GET_ARR(R, r);
  SHAPE(A.rows == R.rows && A.cols == R.cols && W.rows == A.rows && (W.cols == A.cols || W.cols == 1), "Q, r equal; omega rows x (1|nv)");
  k_clamped_update<<<GRID(A.rows * A.cols)>>>(A.p, W.p, (int)W.cols, R.p, A.rows, (int)A.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_local_step_update(ibx_ctx* c, ibx_array Q0, ibx_array R, ibx_array cfl, ibx_array mask, float alpha, ibx_array Q) {
  CHECK_CTX(c);
  GET_ARR(A0, Q0);
  GET_ARR(Rr, R);
  GET_ARR(Cf, cfl);
  GET_ARR(A, Q);
  const float* m = nullptr;
  if (mask) {
    GET_ARR(Mk, mask);
    SHAPE(Mk.rows == A.rows && Mk.cols == 1, "mask is rows x 1");
    m = Mk.p;
  }
  SHAPE(A0.rows == A.rows && A0.cols == A.cols && Rr.rows == A.rows && Rr.cols == A.cols && Cf.rows == A.rows && Cf.cols == 1,
        "Q0, R, Q equal; cfl rows x 1");
  if (A.rows == 0) return IBX_OK;
  k_local_step<<<GRID(A.rows * A.cols)>>>(A0.p, Rr.p, Cf.p, m, alpha, A.p, A.rows, (int)A.cols);
  LAUNCH_CHECK();
  return IBX_OK;
}

int ibx_reduce(ibx_ctx* c, int op, ibx_array a, int per_column, double* out) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  if (op < 0 || op > 4) return fail(IBX_ERR_ARG, "ibx_reduce: unknown op");
  if (A.rows == 0) return fail(IBX_ERR_ARG, "ibx_reduce: empty array");
  if (per_column) return reduce_impl(c, op, A.p, nullptr, nullptr, A.rows, (int)A.cols, out);
  return reduce_impl(c, op, A.p, nullptr, nullptr, A.rows * A.cols, 1, out);
}

int ibx_dot(ibx_ctx* c, ibx_array a, ibx_array b, double* out) {
  CHECK_CTX(c);
  GET_ARR(A, a);
  GET_ARR(B, b);
  SHAPE(A.rows == B.rows && A.cols == B.cols, "a and b must match");
  return reduce_impl(c, 5, A.p, B.p, nullptr, A.rows * A.cols, 1, out);
}

int ibx_volume_integral(ibx_ctx* c, const ibx_domain* d, ibx_array Ah, float* out) {
  CHECK_CTX(c);
  GET_DOM(D, d);
  GET_ARR(A, Ah);
  SHAPE(A.rows == D.ncells, "A must have ncells rows");
  // cell volumes prod_d spacing_d are formed from the global width table (identical to the per-partition
  // products of src/ImmersedBoundary.jl:1417-1421)
  float* vol = ensure_scratch(c, D.ncells * (1 + D.nd));
  if (!vol) return fail(IBX_ERR_CUDA, "ibx_volume_integral: out of device memory");
  float* wcm = vol + D.ncells;
  std::vector<float> w((size_t)D.ncells * D.nd);
  for (int64_t i = 0; i < D.ncells; ++i)
    for (int k = 0; k < D.nd; ++k) w[(size_t)k * D.ncells + i] = D.widths[(size_t)i * D.nd + k];
  CU(cudaMemcpyAsync(wcm, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  k_prod_spacing<<<GRID(D.ncells)>>>(wcm, D.ncells, D.nd, vol);
  LAUNCH_CHECK();
  double tmp;
  for (int col = 0; col < A.cols; ++col) {
    int rc = reduce_impl(c, 6, A.p + (int64_t)col * A.rows, nullptr, vol, A.rows, 1, &tmp);
    if (rc) return rc;
    out[col] = (float)tmp;
  }
  return IBX_OK;
}

}  // extern "C"
