/*
 * ibx.h -- C ABI of libibx.so: the B200-native (sm_100a) implementation of
 * ImmersedBoundary.jl's partitioned residual-evaluation hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI; its
 * only backend seam is `conv_to_backend` / `to_backend` (src/arraybends.jl:14-77,
 * src/ImmersedBoundary.jl:820-864, :1197-1247) plus multiple dispatch on the array
 * type.  A Julia maintainer binds these entry points with `ccall` (see INTEGRATION.md
 * and immersedboundary.jl_b200/julia/ImmersedBoundaryB200.jl); the Python test harness
 * binds the same symbols with ctypes.
 *
 * Conventions
 *  - every function returns an int status: 0 = ok, non-zero = error; the message is
 *    available from ibx_last_error().  No exception crosses this boundary.
 *  - all indices are 0-based int32 ("no cell" = -1); the Julia wrapper converts
 *    from/to the reference's 1-based Int32 at the call.
 *  - field arrays are column-major float32 (rows = cells/faces, cols = variables):
 *    exactly Julia's `Matrix{Float32}(N, nv)` memory layout (README.md:172-173).
 *  - `dim` arguments are 0-based.
 *  - host pointers are borrowed for the duration of the call only.
 *  - device memory is owned by the library; callers hold opaque handles.
 *  - there is no CPU fallback: every compute entry point fails with IBX_ERR_CUDA when
 *    no sm_100-class device is usable.
 */
#ifndef IBX_H
#define IBX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBX_OK 0
#define IBX_ERR_ARG 1
#define IBX_ERR_CUDA 2
#define IBX_ERR_STATE 3
#define IBX_ERR_UNSUPPORTED 4
#define IBX_ERR_NCCL 5

typedef struct ibx_ctx ibx_ctx;           /* one per device / calling process       */
typedef struct ibx_stl ibx_stl;           /* host: Stereolitography                 */
typedef struct ibx_dfield ibx_dfield;     /* host: DistanceField (KD-tree on STL)   */
typedef struct ibx_mesh ibx_mesh;         /* host: block octree Mesh                */
typedef struct ibx_domain ibx_domain;     /* host tables (+ device copies)          */
typedef struct ibx_accum ibx_accum;       /* Accumulator (CSR; host + device)       */
typedef int64_t ibx_array;                /* device array handle (0 = invalid)      */

const char* ibx_last_error(void);
const char* ibx_version(void);

/* ------------------------------------------------------------------ geometry (host) */
/* Stereolitography (src/mesher.jl:238-296). points: npoints x nd row-major; f32 != 0 means the
 * source element type is Float32 (arithmetic is then carried out in float like Julia would). */
int ibx_stl_create(int nd, int64_t npoints, const double* points, int64_t nsimp, const int64_t* simplices,
                   int is_f32, ibx_stl** out);
int ibx_stl_read(const char* path, ibx_stl** out);            /* .dat (Selig), ASCII or binary STL */
int ibx_stl_free(ibx_stl* s);
int ibx_stl_info(const ibx_stl* s, int* nd, int64_t* npoints, int64_t* nsimp, int* is_f32);
int ibx_stl_copy(const ibx_stl* s, double* points, int64_t* simplices);
/* merge_points (src/mesher.jl:351-407) over `n` inputs */
int ibx_stl_merge_points(int n, ibx_stl* const* in, double tolerance, int tol_is_f32, int clean_degenerate,
                         ibx_stl** out);
/* feature_regions (src/mesher.jl:670-728) */
int ibx_stl_feature_regions(const ibx_stl* s, double angle_deg, double radius, int include_boundaries, ibx_stl** out);
/* centers_and_normals (src/mesher.jl:639-660): nsimp x nd each */
int ibx_stl_centers_normals(const ibx_stl* s, double* centers, double* normals);

/* Refinement criteria: distance function + local size (src/mesher.jl:27-122, :736-769).
 * kind: 0 Ball(center[nd], radius=a[0]); 1 Box(origin=c, widths=a[0..nd)); 2 Line(p1=c, p2=a[0..nd));
 *       3 DistanceField(dfield).  `h_is_f32` mirrors the element type of the size literal. */
typedef struct ibx_region {
  int kind;
  int h_is_f32;
  double c[3];
  double a[3];
  const ibx_dfield* dfield;
  double h;
} ibx_region;

/* DistanceField(stl; h = 0) (src/mesher.jl:748-760); refine_to_length (src/mesher.jl:503-528) */
int ibx_dfield_create(const ibx_stl* s, ibx_dfield** out);
int ibx_dfield_free(ibx_dfield* d);
int ibx_dfield_stl(const ibx_dfield* d, const ibx_stl** out);
int ibx_dfield_distance(const ibx_dfield* d, int64_t n, const float* x, double* out);   /* x: n x nd row-major */
int ibx_stl_refine_to_length(const ibx_stl* s, double h, int h_is_f32, double tolerance, int tol_is_f32,
                             double growth_ratio, int nregions, const ibx_region* regions, ibx_stl** out);

/* Mesh(origin, widths, surfaces...; refinement_regions, growth_ratio, block_size)
 * (src/mesher.jl:972-1046).  Surface i = (names[i], stls[i], hs[i]).  A surface may instead be
 * an analytic sphere (stls[i] == NULL, sphere_c/sphere_r give centre/radius): an extension used
 * for the synthetic throughput meshes, where the STL refinement the reference performs
 * (millions of triangles) is replaced by the exact distance/projection. */
typedef struct ibx_surface {
  const char* name;
  const ibx_stl* stl;
  double h;
  int h_is_f32;
  double sphere_c[3];
  double sphere_r;
} ibx_surface;
int ibx_mesh_create(int nd, const float* origin, const float* widths, int nsurf, const ibx_surface* surfaces,
                    int nregions, const ibx_region* regions, double growth_ratio, double tolerance, int tol_is_f32,
                    int block_size, ibx_mesh** out);
/* positional struct constructor used by multigrid (src/ImmersedBoundary.jl:1366-1368); shares the
 * distance fields of `like` */
int ibx_mesh_from_blocks(const ibx_mesh* like, int block_size, ibx_mesh** out);
int ibx_mesh_free(ibx_mesh* m);
/* Mesh (de)serialisation: a Mesh is plain data in the reference (src/mesher.jl:926-933; Julia's `Serialization` stdlib
 * writes it as is); behind this ABI it is an opaque handle, so the library writes / reads it itself (root box, block list,
 * per surface its name and refined STL or analytic-sphere parameters; distance fields are rebuilt on load). */
int ibx_mesh_save(const ibx_mesh* m, const char* path);
int ibx_mesh_load(const char* path, ibx_mesh** out);
int ibx_mesh_info(const ibx_mesh* m, int* nd, int* block_size, int64_t* nblocks, int64_t* ncells, int* nsurf);
int ibx_mesh_blocks(const ibx_mesh* m, float* block_origins, float* block_widths);     /* nblocks x nd row-major */
int ibx_mesh_surface_name(const ibx_mesh* m, int i, const char** name);
int ibx_mesh_surface_dfield(const ibx_mesh* m, int i, const ibx_dfield** out);        /* NULL for analytic surfaces */
/* get_cells (src/mesher.jl:1064-1112): centers/widths ncells x nd row-major */
int ibx_mesh_cells(const ibx_mesh* m, float* centers, float* widths);

/* ------------------------------------------------------------------ domain tables (host) */
/* Domain(msh; max_partition_size, partition_skirt_depth, ghost_layer_ratio, hypercube_families)
 * (src/ImmersedBoundary.jl:536-786).  Family f covers faces fam_ptr[f]..fam_ptr[f+1] of
 * (fam_dim[], fam_front[]).  build_partitions = 0 skips the per-partition tables (the fused,
 * block-structured kernels do not need them). */
int ibx_domain_build(const ibx_mesh* m, int64_t max_partition_size, int skirt_depth, float ghost_layer_ratio,
                     int nfam, const char* const* fam_names, const int* fam_ptr, const int* fam_dim,
                     const int* fam_front, int build_partitions, int build_surfaces, ibx_domain** out);
/* Same global cell numbering and block connectivity, but ghosts / image stencils are only built for the cells of
 * `rank`'s contiguous block range: the input of ibx_domain_shard when every rank builds its own tables (no
 * partition tables, no surfaces).  Cost: O(all blocks) + O(owned cells). */
int ibx_domain_build_for_rank(const ibx_mesh* m, float ghost_layer_ratio, int nfam, const char* const* fam_names,
                              const int* fam_ptr, const int* fam_dim, const int* fam_front, int rank, int nranks,
                              ibx_domain** out);
int ibx_domain_free(ibx_domain* d);
int ibx_domain_info(const ibx_domain* d, int* nd, int64_t* ncells, int64_t* nfaces, int* npartitions,
                    int* nboundaries, int* nsurfaces);
/* global face list (dim, owner, neigh), canonical order: interior by (owner, neigh), then box faces */
int ibx_domain_faces(const ibx_domain* d, int32_t* faces3);
int ibx_domain_cells(const ibx_domain* d, float* centers, float* widths);   /* ncells x nd row-major; either may be NULL */
/* block-structured connectivity used by the fused kernels: two_to_one = every block contact is same-level
 * or 2:1; block_faces: nblocks x 2nd x 7 = (kind, nb0..nb3, sub0, sub1), kind 0 box / 1 same / 2 coarser /
 * 3 finer / 4 irregular */
int ibx_domain_flags(const ibx_domain* d, int* two_to_one, int64_t* nblocks);
int ibx_domain_block_faces(const ibx_domain* d, int32_t* out);
/* partition p (0-based): sizes then tables (src/ImmersedBoundary.jl:383-392, :605-698) */
int ibx_partition_info(const ibx_domain* d, int p, int64_t* n_domain, int64_t* n_image, int64_t* image_start,
                       int64_t* nfaces_per_dim /* [nd] */);
int ibx_partition_tables(const ibx_domain* d, int p, int32_t* domain, int32_t* image_in_domain);
int ibx_partition_faces(const ibx_domain* d, int p, int dim, int32_t* owners, int32_t* neighbors);
/* left (side 0) / right (side 1) face lists as CSR over the partition's domain cells */
int ibx_partition_face_lists(const ibx_domain* d, int p, int dim, int side, int32_t* ptr /* n_domain+1 */,
                             int32_t* idx /* ptr[n_domain] */);
/* boundaries (src/ImmersedBoundary.jl:406-476): boundary b has nparts chunks of <= max_partition_size ghosts */
int ibx_boundary_name(const ibx_domain* d, int b, const char** name, int* nparts);
int ibx_boundary_info(const ibx_domain* d, int b, int part, int64_t* nghost, int64_t* n_image_domain, int64_t* nnz);
/* ghosts of the chunk whose k-th and (k+1)-th nearest donor candidates are exactly equidistant from the image point: the
 * only ghosts where NearestNeighbors.jl's traversal-order tie-break (nninterp.jl:121-123; not reproducible)
 * could pick a different donor than this library's documented (distance, lower index) rule */
int ibx_boundary_tie_count(const ibx_domain* d, int b, int part, int64_t* n_tied);
int ibx_boundary_tables(const ibx_domain* d, int b, int part, int32_t* ghost_indices, float* projections,
                        float* normals, float* image_distances, float* ghost_distances, int32_t* image_domain,
                        int32_t* interp_ptr, int32_t* interp_idx /* into image_domain */, float* interp_w);
/* surfaces (src/ImmersedBoundary.jl:335-376, :743-763) */
int ibx_surface_name(const ibx_domain* d, int s, const char** name);
int ibx_surface_info(const ibx_domain* d, int s, int64_t* npoints, int64_t* nnz, int64_t* nnz_offset);
int ibx_surface_tables(const ibx_domain* d, int s, float* points, float* offsets, float* normals, float* areas,
                       int32_t* ptr, int32_t* idx, float* w, int32_t* optr, int32_t* oidx, float* ow);

/* Interpolator(X, Xc; bias, linear, k) (src/nninterp.jl:85-138) on arbitrary point clouds; CSR out.
 * X: n x nd, Xc/bias: q x nd row-major (bias may be NULL).  k = 0 means 2^nd. */
int ibx_interpolator_build(int nd, int64_t n, const float* X, int64_t q, const float* Xc, const float* bias,
                           int linear, int k, ibx_accum** out);
/* multigrid transfer pair between two cell-centre clouds (src/ImmersedBoundary.jl:1391-1392) is two calls
 * of ibx_interpolator_build with linear = 0.
 * GeometricMultigrid.coarsener_and_prolongator (src/mgrid.jl:24-97) */
int ibx_mgrid_build(int nd, int64_t n, const float* X, int level, const float* volumes /* or NULL */,
                    ibx_accum** coarsener, ibx_accum** prolongator);
/* Accumulator(inds, weights) (src/accumulator.jl:39-65) from CSR; w may be NULL (unweighted) */
int ibx_accum_create(int64_t n_out, const int32_t* ptr, const int32_t* idx, const float* w, ibx_accum** out);
int ibx_accum_info(const ibx_accum* a, int64_t* n_out, int64_t* nnz, int* weighted);
int ibx_accum_tables(const ibx_accum* a, int32_t* ptr, int32_t* idx, float* w);
int ibx_accum_free(ibx_accum* a);

/* ------------------------------------------------------------------ device context + arrays */
int ibx_init(int device, ibx_ctx** out);
int ibx_finalize(ibx_ctx* c);
int ibx_sync(ibx_ctx* c);
int ibx_device_info(ibx_ctx* c, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);
/* number of kernels launched by this context since creation (for bench.py's gpu_launches) */
int ibx_launch_count(ibx_ctx* c, int64_t* out);
/* CUDA-event timing on the context's compute stream */
int ibx_timer_start(ibx_ctx* c);
int ibx_timer_stop(ibx_ctx* c, float* ms);

int ibx_array_alloc(ibx_ctx* c, int64_t rows, int64_t cols, ibx_array* out);
/* Float64 arrays exist for one reason: the reference's HLL flux is Float64 (src/cfd.jl:504-507) and so is every
 * Green-Gauss sum taken of it.  Accepted by: ibx_inviscid_fluxes_hll (F), ibx_green_gauss / ibx_unsigned_green_gauss
 * (uf and out), ibx_ew_binary / ibx_ew_scalar / ibx_ew_unary (any operand; Julia promotion, conversion on store),
 * ibx_array_copy / shape / free.  Every other entry point rejects them with IBX_ERR_UNSUPPORTED. */
int ibx_array_alloc_f64(ibx_ctx* c, int64_t rows, int64_t cols, ibx_array* out);
int ibx_array_is_f64(ibx_ctx* c, ibx_array a, int* out);
int ibx_array_download_f64(ibx_ctx* c, ibx_array a, double* host);
int ibx_array_free(ibx_ctx* c, ibx_array a);
int ibx_array_shape(ibx_ctx* c, ibx_array a, int64_t* rows, int64_t* cols);
int ibx_array_upload(ibx_ctx* c, ibx_array a, const float* host);        /* column-major rows x cols */
int ibx_array_download(ibx_ctx* c, ibx_array a, float* host);
/* Asynchronous forms for hosts that keep the state in (pinned) host memory and pipeline independent evaluations, also on a
 * rank-local shard (the sharded counterpart of ibx_euler_step_host_begin/_end):
 *   upload_async   copies on the upload stream; everything enqueued afterwards on the compute / halo streams waits for it.
 *                  The caller guarantees that no earlier work still reads `a` (e.g. by ibx_download_wait on its slot).
 *   download_async copies on the download stream once the compute stream's work enqueued so far has finished.
 *   download_fence marks the downloads enqueued so far as slot 0 / 1; download_wait blocks the host until they landed. */
int ibx_array_upload_async(ibx_ctx* c, ibx_array a, const float* host);
int ibx_array_download_async(ibx_ctx* c, ibx_array a, float* host);
int ibx_download_fence(ibx_ctx* c, int slot);
int ibx_download_wait(ibx_ctx* c, int slot);
int ibx_array_fill(ibx_ctx* c, ibx_array a, float v);
int ibx_array_copy(ibx_ctx* c, ibx_array dst, ibx_array src);
int ibx_array_devptr(ibx_ctx* c, ibx_array a, void** ptr);
int ibx_array_column(ibx_ctx* c, ibx_array a, int col, ibx_array out);        /* out (rows x 1) = a[:, col]  */
int ibx_array_set_column(ibx_ctx* c, ibx_array a, int col, ibx_array src);    /* a[:, col] = src (rows x 1)  */
/* pinned host staging buffers for the end-to-end path */
int ibx_host_alloc(int64_t bytes, void** out);
int ibx_host_free(void* p);

/* upload the tables of a domain / accumulator once (replaces the per-call to_backend(part, ...) of
 * src/ImmersedBoundary.jl:846-849) */
int ibx_domain_upload(ibx_ctx* c, ibx_domain* d);
int ibx_accum_upload(ibx_ctx* c, ibx_accum* a);

/* ------------------------------------------------------------------ partition runtime (K1) */
/* dargs = a[part.domain, :] and a[part.image, :] = da[part.image_in_domain, :]
 * (src/ImmersedBoundary.jl:836-859) */
int ibx_gather_domain(ibx_ctx* c, const ibx_domain* d, int p, ibx_array global_in, ibx_array local_out);
int ibx_scatter_image(ibx_ctx* c, const ibx_domain* d, int p, ibx_array local_in, ibx_array global_out);
int ibx_partition_spacing(ibx_ctx* c, const ibx_domain* d, int p, ibx_array out /* n_domain x nd */);
int ibx_partition_centers(ibx_ctx* c, const ibx_domain* d, int p, ibx_array out /* n_domain x nd */);

/* ------------------------------------------------------------------ grid operators (K2-K5)
 * u: n_domain x nv (or faces x nv); outputs must be pre-allocated with the right shape. */
int ibx_at_owners(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out);      /* :879 */
int ibx_at_neighbors(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out);   /* :889 */
int ibx_at_faces(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out);       /* :899 */
int ibx_green_gauss(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array uf, ibx_array out);   /* :918 */
int ibx_unsigned_green_gauss(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array uf, ibx_array out); /* :934 */
int ibx_cell_gradient(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out);  /* :965 */
int ibx_face_distance(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out);               /* :995 */
int ibx_owner_distance(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out);              /* :1010 */
int ibx_neighbor_distance(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array out);           /* :1024 */
int ibx_face_gradient(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array out);  /* :1039 */
/* JST_sensor(part, p, dim); dim = -1 is the reference's dim = 0 (max over dims) */
int ibx_jst_sensor(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array pr, ibx_array out);    /* :1077 */
/* MUSCL(part, u, du, dim; D, high_order); D = 0 handle means `nothing` */
int ibx_muscl(ibx_ctx* c, const ibx_domain* d, int p, int dim, ibx_array u, ibx_array du, ibx_array D,
              int high_order, ibx_array uL, ibx_array uR);                                           /* :1113 */

/* ------------------------------------------------------------------ cfd.jl (K6, K13) */
typedef struct ibx_fluid { float R, gamma; } ibx_fluid;                     /* src/cfd.jl:14-53 */
int ibx_state2primitive(ibx_ctx* c, ibx_fluid f, ibx_array Q, ibx_array P); /* :137 */
int ibx_primitive2state(ibx_ctx* c, ibx_fluid f, ibx_array P, ibx_array Q); /* :106 */
int ibx_speed_of_sound(ibx_ctx* c, ibx_fluid f, ibx_array T, ibx_array a);  /* :62 */
int ibx_inviscid_fluxes_hll(ibx_ctx* c, ibx_fluid f, ibx_array PL, ibx_array PR, int dim, ibx_array F);          /* :459 */
int ibx_inviscid_fluxes_sensor(ibx_ctx* c, ibx_fluid f, ibx_array PL, ibx_array PR, ibx_array nuL, ibx_array nuR,
                               int dim, ibx_array F);                                                            /* :516 */
/* FlowBC(fluid, Pinf; normal_flow)(P, normals) (src/cfd.jl:243-300); Pinf has nv (or 3 when normal_flow) entries */
int ibx_flowbc(ibx_ctx* c, ibx_fluid f, const float* Pinf, int n_pinf, int normal_flow, ibx_array P,
               ibx_array normals, ibx_array out);

/* the same call with the keyword arguments of src/cfd.jl:245-249: transpiration (per-point array, or 0 and the scalar)
 * and the wall-shear scaling of :290-296 when image_distances and du_dn are both given (both 0: none) */
int ibx_flowbc_ex(ibx_ctx* c, ibx_fluid f, const float* Pinf, int n_pinf, int normal_flow, ibx_array P, ibx_array normals,
                  ibx_array image_distances, ibx_array du_dn, ibx_array transpiration, float transpiration_scalar, ibx_array out);

/* ------------------------------------------------------------------ pointwise closures around the residual */
/* Sutherland / conductivity-polynomial constants of Fluid (src/cfd.jl:14-53): mu_ref, T_ref, S, k[0..nk-1] */
typedef struct ibx_transport { float mu_ref, T_ref, S; int nk; float k[4]; } ibx_transport;
int ibx_dynamic_viscosity(ibx_ctx* c, ibx_transport t, ibx_array T, ibx_array mu);   /* src/cfd.jl:71-77 */
int ibx_heat_conductivity(ibx_ctx* c, ibx_transport t, ibx_array T, ibx_array k);    /* :84-90 */
/* viscous_fluxes(fluid, P, Pgrad, dim; mu_t) (src/cfd.jl:664-736).  Pgrad: nd arrays shaped like P (gradient along
 * each axis).  normals != 0: the `dim::AbstractMatrix` method (N x nd direction matrix), `dim` ignored.
 * mu_t != 0: per-point eddy viscosity, else the scalar mu_t_scalar. */
int ibx_viscous_fluxes(ibx_ctx* c, ibx_transport t, ibx_array P, const ibx_array* Pgrad, int dim, ibx_array normals,
                       ibx_array mu_t, float mu_t_scalar, ibx_array F);
int ibx_jst_sensor3(ibx_ctx* c, ibx_array Pim1, ibx_array Pi, ibx_array Pip1, ibx_array out);   /* :563-573 */
/* g[i * nd + j] = d u_i / d x_j, nd * nd vectors (the reference's matrix of vectors) */
int ibx_shock_sensor(ibx_ctx* c, int nd, const ibx_array* g, ibx_array out);                    /* :589-617 */
int ibx_pressure_coefficient(ibx_ctx* c, float gamma, ibx_array p, float p_inf, float M_inf, ibx_array Cp);   /* :420-426 */
/* src/turbulence.jl */
typedef struct ibx_wall_params { float kappa, C, A, beta, beta_star, D, A_plus, omega; int n_iter; } ibx_wall_params;   /* :27-33 */
int ibx_wall_function_rey(ibx_ctx* c, ibx_wall_params w, ibx_array Rey, ibx_array y_plus, ibx_array u_plus, ibx_array mu_plus,
                          ibx_array k_plus, ibx_array dudy_plus);                                /* :27-72 */
int ibx_wall_function(ibx_ctx* c, ibx_wall_params w, ibx_array y, ibx_array u, ibx_array nu, ibx_array u_tau, ibx_array nu_t,
                      ibx_array k, ibx_array omega, ibx_array eps, ibx_array dudn);              /* :74-98 */
int ibx_shear_rate(ibx_ctx* c, int nd, const ibx_array* g, ibx_array out);                       /* :110-124 */
int ibx_smagorinsky(ibx_ctx* c, ibx_array Delta, ibx_array S, float Cs, ibx_array out);          /* :126-137 */
int ibx_standard_keps(ibx_ctx* c, ibx_array k, ibx_array eps, ibx_array S, float Cmu, float sigma_k, float sigma_eps, float C1,
                      float C2, ibx_array nu_k, ibx_array nu_eps, ibx_array Sk, ibx_array Seps, ibx_array nu_t);   /* :175-194 */
int ibx_wray_agarwal(ibx_ctx* c, ibx_array R, ibx_array S, ibx_array gradR, ibx_array gradS, float sigma_R, float C1, float kappa,
                     ibx_array nu_R, ibx_array S_out);                                           /* :222-241 */
int ibx_ducros_sensor(ibx_ctx* c, int nd, const ibx_array* g, ibx_array out);                    /* :253-283 */
int ibx_wale(ibx_ctx* c, ibx_array Delta, const ibx_array* g, float Cw, ibx_array out);          /* :292-337, 3-D only */

/* ------------------------------------------------------------------ accumulators, IB ghost update (K8, K9) */
/* out = acc(v)  (src/accumulator.jl:78-130); delta: subtract v[row] first (Delta = true) */
int ibx_accumulate(ibx_ctx* c, const ibx_accum* a, ibx_array v, int delta, ibx_array out);
/* the `f` and `op` keyword arguments of the same call (src/accumulator.jl:78-81): f is applied to the gathered values
 * (after the Delta subtraction, before the weights), the stencil rows are reduced with op in list order.
 * f_kind: 0 identity, 1 abs, 2 x -> x^2, 3 sign; op_kind: 0 +, 1 max, 2 min, 3 *.  (A closure cannot cross a C ABI: the
 * Julia wrapper maps `abs`, `abs2`, `sign`, `+`, `max`, `min`, `*` to these codes and, for any other ELEMENTWISE f with
 * Delta = false, applies f on the device array first -- f(v[stencil]) == f.(v)[stencil].) */
int ibx_accumulate_ex(ibx_ctx* c, const ibx_accum* a, ibx_array v, int delta, int f_kind, int op_kind, ibx_array out);
/* impose_bc! pieces (src/ImmersedBoundary.jl:1197-1247) for boundary b, chunk part */
int ibx_bc_image_values(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia);   /* :1228 */
int ibx_bc_normals(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array out /* nghost x nd */);
int ibx_bc_blend(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia, ibx_array ba); /* :1242 */
int ibx_bc_blend_scalar(ibx_ctx* c, const ibx_domain* d, int b, int part, ibx_array a, ibx_array ia, float ba);
/* Surface(u), at_offset (src/ImmersedBoundary.jl:368-376), surface_integral (:351-361) */
int ibx_surface_values(ibx_ctx* c, const ibx_domain* d, int s, int offset, ibx_array u, ibx_array out);
int ibx_surface_integral(ibx_ctx* c, const ibx_domain* d, int s, ibx_array u, float* out /* cols */);

/* ------------------------------------------------------------------ elementwise glue + reductions (K12) */
/* op: 0 add, 1 sub, 2 mul, 3 div, 4 max, 5 min.  b may have 1 column (broadcast over a's columns). */
int ibx_ew_binary(ibx_ctx* c, int op, ibx_array a, ibx_array b, ibx_array out);
int ibx_ew_scalar(ibx_ctx* c, int op, ibx_array a, float s, int scalar_first, ibx_array out);
/* op: 0 abs, 1 neg, 2 sqrt, 3 sign, 4 reciprocal */
int ibx_ew_unary(ibx_ctx* c, int op, ibx_array a, ibx_array out);
int ibx_axpy(ibx_ctx* c, float alpha, ibx_array x, ibx_array y);                 /* y += alpha * x */
int ibx_clamped_update(ibx_ctx* c, ibx_array Q, ibx_array omega, ibx_array r);   /* Q += clamp(omega,0,1) * r (solver.jl:82) */
/* One stage of a local-time-step march, Q = Q0 + ((alpha / cfl_i) * R) * mask_i  (mask = 0: no mask).  This is the update
 * the reference's drivers write as broadcasts around `FAS!`'s f(l, Q) = (R .* CFL ./ cfl, 1) (src/solver.jl:78-82,
 * test/advection.jl:28-46); Q may alias Q0. */
int ibx_local_step_update(ibx_ctx* c, ibx_array Q0, ibx_array R, ibx_array cfl, ibx_array mask, float alpha, ibx_array Q);
/* op: 0 sum, 1 max, 2 min, 3 max|.|, 4 sum of squares (double accumulation); one value per column, or over
 * everything when per_column = 0 */
int ibx_reduce(ibx_ctx* c, int op, ibx_array a, int per_column, double* out);
int ibx_dot(ibx_ctx* c, ibx_array a, ibx_array b, double* out);
/* volume_integral (src/ImmersedBoundary.jl:1415-1431): sum_cells A * prod_d spacing_d, per column */
int ibx_volume_integral(ibx_ctx* c, const ibx_domain* d, ibx_array A, float* out);

/* ------------------------------------------------------------------ point-implicit (K10) */
/* per-cell pinv of nv x nv blocks (src/point_implicit.jl:128-135); D: N x (nv*nv), block (p, j, i) at column j + nv*i */
int ibx_block_pinv(ibx_ctx* c, ibx_array D, int nv, ibx_array Dinv);
/* out[p, j] = sum_i Dinv[p, j, i] * v[p, i]  (src/point_implicit.jl:153-161) */
int ibx_block_apply(ibx_ctx* c, ibx_array Dinv, int nv, ibx_array v, ibx_array out);

/* ------------------------------------------------------------------ fused, block-structured residuals */
/* Run-time options of the fused Euler residual, per context (they replace the IBX_* environment switches of round 1).
 *   "arithmetic": 0 (default) reference-exact -- no FMA contraction, Float64 HLL combination and Green-Gauss sums like
 *                   src/cfd.jl:504-507 / src/ImmersedBoundary.jl:918-926: bit-identical to the oracle;
 *                 1 fast -- Float32 flux, FMA contraction, approximate reciprocals: within the north-star tolerance when the
 *                   error is scaled by the face fluxes, not bit-identical.
 *   "path":       0 (default) marching kernels where they apply (3-D, block size 8, power-of-two spacings), tile kernels
 *                   elsewhere; 1 tile kernels everywhere; 2 per-cell gather kernels.  1 and 2 are independent
 *                   implementations kept as fallbacks and cross-checks (same bits as 0).
 *   "sensor":     1 (default) MUSCL(...; D = JST_sensor(p)); 0 MUSCL(...; D = nothing)
 *                   (src/ImmersedBoundary.jl:1117,1141): plain limited reconstruction. */
int ibx_set_option(ibx_ctx* c, const char* name, int value);
int ibx_get_option(ibx_ctx* c, const char* name, int* value);
/* Linear advection residual of test/advection.jl:67-83 on the whole domain:
 * ud = -sum_dim GG(upwind MUSCL flux), spec = max_dim UGG(at_faces(C_dim)). u, ud, spec: N; C: N x nd. */
int ibx_residual_advection(ibx_ctx* c, const ibx_domain* d, ibx_array u, ibx_array C, ibx_array ud, ibx_array spec);
/* Canonical Euler residual: Q -> (R, cfl).  flux_kind 0 = HLL, 1 = sensor-Rusanov.
 * Domain of the default path ("path" = 0 on 3-D, block size 8): its divisions, square roots and the Float64 reciprocal
 * are the straight-line sequences ptxas emits when its own range check passes, without the check.  They are correctly
 * rounded for operands and results between 2^-100 and 2^100 in magnitude -- pressures, R T, gamma R T (the temperature
 * is clamped at 10 K first, src/cfd.jl:140) and wave-speed differences of any physical flow; a zero wave-speed
 * difference gives NaN like the reference's 0 / 0.  A state that has diverged beyond that range gets unspecified finite
 * values or NaN instead of the IEEE result: detect it with ibx_reduce (max / min / max|.| propagate NaN) or re-run the
 * step with "path" = 1 or 2, whose kernels keep the guarded library forms. */
int ibx_residual_euler(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, ibx_array Q, ibx_array R,
                       ibx_array cfl);
/* Configuration C5 -- canonical RANS residual (ours: the reference ships the pieces, not a composition; oracle/euler.py
 * rans_residual states it with the restated operators).  Q: ncells x 5 mean-flow state, qR = rho * R: the transported
 * Wray-Agarwal variable (src/turbulence.jl:197-241, nu_t = R).  R, cfl = ibx_residual_euler (HLL) plus, per dimension,
 * green_gauss(viscous_fluxes(at_faces(P), face_gradient(P, dim, cell_gradient(P)), dim; mu_t = at_faces(rho R)))
 * (src/cfd.jl:664-736, src/ImmersedBoundary.jl:1039-1069); RR = div[rho (nu + sigma_R R) grad R] - div(rho u R) (upwind on
 * MUSCL states) + rho S_WA(R, shear_rate(grad u), grad R, grad shear_rate).  3-D; runs on rank-local shards as well. */
int ibx_residual_rans(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, ibx_transport t, float sigma_R, float C1, float kappa,
                      ibx_array Q, ibx_array qR, ibx_array R, ibx_array RR, ibx_array cfl);
/* IB ghost update of the transported variable of C5 for boundary b: impose_bc!(dom, b, R) do ...; R_bc end on R = qR / rho
 * (src/ImmersedBoundary.jl:1197-1247; `R = 0` at walls, `R_inf = 3 nu` in the far field, src/turbulence.jl:203), written back
 * as rho_ghost * R_ghost.  Call after ibx_ghost_update_euler of the same boundary (it needs the ghost densities). */
int ibx_ghost_update_rans(ibx_ctx* c, const ibx_domain* d, int b, ibx_array Q, ibx_array qR, float R_bc);
/* IB ghost update on the conservative state for boundary b with FlowBC(fluid, Pinf; normal_flow):
 * image interpolation of P = state2primitive(Q), BC, eta-blend, primitive2state, written Jacobi-style. */
int ibx_ghost_update_euler(ibx_ctx* c, const ibx_domain* d, int b, ibx_fluid f, const float* Pinf, int n_pinf,
                           int normal_flow, ibx_array Q);
typedef struct ibx_bc_spec { int boundary; int normal_flow; int n_pinf; float Pinf[5]; } ibx_bc_spec;
/* One step of a solver loop on a whole (unsharded) domain: ghost updates of `bcs` in order, then the residual -- with the
 * ghost update running on a second, high-priority stream under the residual of the blocks that do not read a ghost cell
 * (3-D, block size 8; elsewhere the plain sequence).  Same bits as ibx_ghost_update_euler x nbc + ibx_residual_euler.
 * ibx_step_euler_sharded (below) is the same call on a rank-local shard, where the halo exchanges are hidden too. */
int ibx_step_euler(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs, ibx_array Q,
                   ibx_array R, ibx_array cfl);
/* End-to-end convenience for hosts that hold the state in host memory: H2D(Q) -> ghost updates of
 * every boundary in `bcs` order -> residual -> D2H(R, cfl).  Host buffers column-major. */
int ibx_euler_step_host(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                        const float* Q_host, float* R_host, float* cfl_host);
/* The same call split in two so that consecutive, independent evaluations (finite-difference JVP probes of
 * src/point_implicit.jl:98-114, Hutchinson samples :18-47) overlap: `begin` enqueues H2D(Q) -> ghost updates ->
 * residual -> D2H(R, cfl) on slot 0 or 1 and returns; `end` blocks until the slot's results are in the host buffers.
 * The host buffers must stay valid (and pinned, for the copies to be asynchronous) until `end`. */
int ibx_euler_step_host_begin(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                              const float* Q_host, float* R_host, float* cfl_host, int slot);
int ibx_euler_step_host_end(ibx_ctx* c, int slot);

/* ------------------------------------------------------------------ multi-GPU (one rank per GPU) */
/* Rank r owns the contiguous cell range of its blocks; halo = cells of other ranks within the 2-deep
 * skirt plus image donors.  The unique id is created on rank 0 and distributed by the host program. */
int ibx_comm_unique_id(char id[128]);
int ibx_comm_init(ibx_ctx* c, int rank, int nranks, const char id[128]);
int ibx_comm_finalize(ibx_ctx* c);
/* build the rank-local domain: owned block range [b0, b1) of `global`, halo tables for all peers */
int ibx_domain_shard(const ibx_domain* global, int rank, int nranks, ibx_domain** local);
int ibx_shard_info(const ibx_domain* local, int64_t* n_owned, int64_t* n_halo, int64_t* owned_start);
int ibx_shard_tables(const ibx_domain* local, int32_t* local_to_global /* n_owned + n_halo */);
int ibx_halo_sizes(const ibx_domain* local, int nranks, int64_t* send_counts, int64_t* recv_counts);
int ibx_halo_lists(const ibx_domain* local, int peer, int32_t* send_local, int32_t* recv_local);
/* tell this rank which of its cells `peer` needs (GLOBAL ids, in the peer's unpack order = recv_local order);
 * the host program moves these lists between ranks (torch.distributed / MPI / files) before ibx_domain_upload */
int ibx_shard_set_send(ibx_domain* local, int peer, int64_t n, const int32_t* global_ids);
/* exchange the halo rows of `a` (n_owned + n_halo rows): pack -> ncclSend/Recv -> unpack, on the comm stream.
 * ibx_residual_euler on the same array completes a posted exchange itself, after converting the owned rows (so the
 * exchange overlaps that kernel); every other consumer calls ibx_halo_end first. */
int ibx_halo_begin(ibx_ctx* c, const ibx_domain* local, ibx_array a);
int ibx_halo_end(ibx_ctx* c, const ibx_domain* local, ibx_array a);
int ibx_allreduce(ibx_ctx* c, int op /* 0 sum 1 max 2 min */, double* inout, int n);
/* One step of a sharded solver loop with the communication hidden behind compute: on the halo stream exchange(Q) -> ghost
 * updates of `bcs` (in order) -> exchange(Q); on the compute stream the residual of the blocks that read neither a ghost
 * nor a halo cell, then -- after the second exchange -- the rest.  Same results as ibx_halo_begin/_end +
 * ibx_ghost_update_euler + ibx_halo_begin + ibx_residual_euler, bit for bit.  Falls back to that sequence where the
 * phase split does not apply (other block sizes / option "path" != 0 / one rank).  exchange_between_families != 0: the halo
 * rows are also refreshed between consecutive families (needed when a ghost of one family interpolates from a ghost of an
 * earlier family owned by another rank; the host program detects this when it trades the halo lists). */
/* owned blocks in the early / late phase of ibx_step_euler_sharded (0, 0 if the phase split does not apply) */
int ibx_shard_phase_info(const ibx_domain* local, int64_t* n_flux_early, int64_t* n_flux_late);
int ibx_step_euler_sharded(ibx_ctx* c, const ibx_domain* local, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs,
                           int exchange_between_families, ibx_array Q, ibx_array R, ibx_array cfl);

/* Pseudo-time march of the Euler residual with local time steps, n_steps steps entirely on the device (the loop a user
 * writes around `FAS!`'s f(l, Q) = (R .* CFL ./ cfl, 1), src/solver.jl:78-82, or by hand as in test/advection.jl:28-46).
 * Per step: ghost updates of `bcs` on Q in place; Q0 = Q; for each of the nstages coefficients a: ibx_step_euler, then
 * Q = Q0 + ((a CFL / cfl) R) live  (live = 0: every cell advances; else an ncells x 1 array of 0 / 1 -- pass 0 on the ghost
 * cells, which only take boundary values).  use_graph != 0 replays one captured step as a CUDA graph (launch-bound small
 * meshes); *graph_used reports it.  Whole domains only. */
int ibx_march_euler(ibx_ctx* c, const ibx_domain* d, ibx_fluid f, int flux_kind, int nbc, const ibx_bc_spec* bcs, ibx_array Q,
                    ibx_array live, int64_t n_steps, float CFL, int nstages, const float* alphas, int use_graph, int* graph_used);

#ifdef __cplusplus
}
#endif
#endif /* IBX_H */
