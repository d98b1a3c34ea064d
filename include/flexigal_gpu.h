/*
 * flexigal_gpu.h -- C ABI of the B200 (sm_100a) Element-Free-Galerkin hot path.
 *
 * This is the drop-in boundary: FlexiGal's unchanged Julia API `ccall`s these symbols
 * (see INTEGRATION.md for the Julia shim); flexigal_b200/space.py binds the same symbols
 * through ctypes and is the tested caller.  Plain pointers and sizes only.
 *
 * The reference (pure Julia, /root/reference) has no FFI for this path; each entry point
 * names the Julia function whose body it replaces (file:line under /root/reference).
 *
 * Conventions are Julia's: arrays column-major (x[d*nnodes+i], gs[c*ngauss+g]), node ids
 * returned 1-based Int64, ascending inside each neighbour list (findall,
 * src/Shape_Functions.jl:26).  All pointers are HOST pointers owned by the caller for the
 * duration of the call only, unless a function says "device".  One fg_ctx = one GPU = one
 * host thread at a time.  Functions return FG_OK (0) or a negative FG_E_* code; the text
 * of the last failure is fg_last_error(ctx).  There is no CPU fallback: every entry point
 * launches CUDA kernels and fails with FG_E_CUDA when no sm_100-class device is usable.
 */
#ifndef FLEXIGAL_GPU_H
#define FLEXIGAL_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_API __attribute__((visibility("default")))

#define FG_OK 0
#define FG_E_ARG (-1)      /* bad argument (the reference's error(...) on bad shape/technique) */
#define FG_E_CUDA (-2)     /* CUDA runtime / launch failure, or no device */
#define FG_E_OOM (-3)      /* device allocation failed */
#define FG_E_STATE (-4)    /* call order violated (e.g. shape functions before neighbours) */
#define FG_E_CAPACITY (-5) /* a neighbour list / sparsity row exceeded a kernel limit */

#define FG_SHAPE_RECTANGULAR 0 /* :rectangular, src/Shape_Functions.jl:6  */
#define FG_SHAPE_CIRCULAR 1    /* :circular / :spherical, :15             */

#define FG_TECH_IMLS 0 /* :IMLS, src/Shape_Functions.jl:251 */
#define FG_TECH_MK 1   /* :MK,   src/Shape_Functions.jl:253 */

/* Bilinear integrands.  Every closure the weak-form DSL can build is bilinear in
 * B_a = [phi_a, d1 phi_a, ..] and B_b at a fixed Gauss point, so it is a coefficient tensor
 *   K_ab[i][j] = ( sum_{al,be} B_a[al] * C[(i*nb+al)*(D*nb) + (j*nb+be)] * B_b[be] ) * w*jac,
 * nb = 1+dim, a = test (row), b = trial (column).  The named kinds are closed forms of the
 * closures the reference's examples use (src/Fields_Operations.jl arithmetic). */
#define FG_OP_GENERIC 0    /* coef = (D*nb)^2 doubles, constant (stride 0) or per Gauss point */
#define FG_OP_GRADGRAD 1   /* c[0] * grad(dT).grad(T)       Fields_Operations.jl:321-323,1002; vector: :552-556 */
#define FG_OP_MASS 2       /* c[0] * dT*T  /  c[0]*(du.u)   :1004 ; :546-550 (penalty term, Assembling_Operators.jl:262) */
#define FG_OP_ADVECTION 3  /* dT * (v.grad T), v = c[0..dim)     examples/Poisson2D_No_Source.jl:30 */
#define FG_OP_ELASTICITY 4 /* grad(du) (.) sigma(u), G = c[0], lambda = c[1]   examples/Elasticity_Test.jl:25-26 */

/* Linear integrands: F[(node-1)*D+i] += ( sum_al c[i*nb+al] * B_a[al] ) * w*jac. */
#define FG_LIN_GENERIC 0 /* coef = D*nb doubles, constant (stride 0) or per Gauss point */
#define FG_LIN_SOURCE 1  /* c[i] * phi_a, i < D     examples/Poisson3D_Source.jl:13 */

typedef struct fg_ctx fg_ctx;

typedef struct fg_operator {
    int32_t kind;         /* FG_OP_* or FG_LIN_* */
    int32_t D;            /* field dimension: 1 scalar, 2/3 vector (VectorField{D}) */
    double c[4];          /* constants of the named kinds */
    const double *coef;   /* FG_*_GENERIC: host pointer to the coefficient tensor(s) */
    int64_t coef_stride;  /* 0 = one constant tensor; else doubles between Gauss points */
} fg_operator;

/* ---- context ------------------------------------------------------------------------- */
FG_API int fg_create(fg_ctx **ctx, int device);
FG_API int fg_destroy(fg_ctx *ctx);
FG_API const char *fg_last_error(const fg_ctx *ctx); /* ctx may be NULL: error of a failed fg_create */
/* Run all work of this context on the given cudaStream_t (NULL = the context's own stream). */
FG_API int fg_set_stream(fg_ctx *ctx, void *cuda_stream);
FG_API int fg_sync(fg_ctx *ctx);
/* Kernels of this library launched by the context so far (bench.py's gpu_launches). */
FG_API int fg_launch_count(fg_ctx *ctx, int64_t *count);
/* Kernel-selection switches (A/B forensics and the tests of the fallback paths; the defaults are the measured best):
 * "assembly_direct" 1 = global-atomics assembly only, "assembly_sym"/"assembly_reg"/"assembly_mma" 0 = without the symmetric /
 * register-block / FP64-MMA variants, "cluster_ucap" / "nb_ucap" = cluster block sizes, "nb_pointwise" 1 = point-wise search,
 * "nb_fused" 1 = single-pass search on recomputation, "assembly_pts" / "imls_warp" 0 = without the point-batched assembly / the warp-autonomous IMLS kernel, "coef_on_device" 1 = fg_operator.coef
 * is a DEVICE pointer on this context's GPU (no upload), "nb_rebin" 1 = re-bin the Gauss points on every search (as for new inputs), "fused" 0 = shape functions and assembly as separate kernels only. */
FG_API int fg_set_option(fg_ctx *ctx, const char *key, int value);

/* ---- inputs -------------------------------------------------------------------------- */
/* Node cloud and support sizes: the (x, dm) pair SHAPE_FUN receives
 * (src/Shape_Functions.jl:241; dm from Influence_Domains, src/FlexiGal.jl:164-199).
 * shape = RECTANGULAR: dm is nnodes x dim; CIRCULAR: dm is nnodes radii.  Bins the nodes. */
FG_API int fg_set_nodes(fg_ctx *ctx, int dim, int64_t nnodes, const double *x, const double *dm, int shape);
/* Gauss table of one integration set: gs is ngauss x (dim+2) = coords.., weight, jacobian
 * (src/Integration.jl:148-218).  set_id is a small non-negative handle chosen by the caller. */
FG_API int fg_set_gauss(fg_ctx *ctx, int set_id, int64_t ngauss, const double *gs);
/* Node subset of one set: its neighbour search sees only the listed nodes (1-based, any order), and DOM / the sparsity stay in
 * the numbering of the full cloud -- what build_space does for domain sets, src/FlexiGal.jl:250-252:
 *   SHAPE_FUN(gs, x[ids,:], dm[ids,:]);  DOM = [ids[idx] for idx in DOM_local].
 * Call between fg_set_gauss (which clears the subset) and fg_neighbours; ids = NULL removes the subset. */
FG_API int fg_set_node_subset(fg_ctx *ctx, int set_id, const int64_t *ids, int64_t n);
/* Frees everything the context holds for one set. */
FG_API int fg_drop_set(fg_ctx *ctx, int set_id);

/* ---- kernel 1: support-domain neighbour search = domain(), src/Shape_Functions.jl:3-27 -- */
FG_API int fg_neighbours(fg_ctx *ctx, int set_id, int64_t *total_L, int32_t *max_L);

/* ---- kernel 2: shape functions = Improved_Moving_Least_Squares / Moving_Kriging,
 *      src/Shape_Functions.jl:88-162 / :164-239, driven like SHAPE_FUN :241-263 ----------- */
FG_API int fg_shape_functions(fg_ctx *ctx, int set_id, int technique);
/* Number of Gauss points whose moment matrix was numerically singular in the last call
 * (their phi/dphi are NaN, as the reference's 1/0 would give). */
FG_API int fg_singular_count(fg_ctx *ctx, int set_id, int64_t *count);

/* (PHI, DPHI, DOM) of SHAPE_FUN as flat arrays: offsets[ngauss+1] (0-based), dom[total_L]
 * (1-based ascending), phi[total_L], dphi[total_L][dim].  Any pointer may be NULL. */
FG_API int fg_get_shapes(fg_ctx *ctx, int set_id, int64_t *offsets, int64_t *dom, double *phi, double *dphi);
/* The same for the Gauss points [g_lo, g_hi) only; offsets (g_hi-g_lo+1 entries, required) are relative to the first fetched entry. */
FG_API int fg_get_shapes_range(fg_ctx *ctx, int set_id, int64_t g_lo, int64_t g_hi, int64_t *offsets, int64_t *dom, double *phi, double *dphi);
/* Flexi_Measure's concatenation of several sets (src/FlexiGal.jl:207-224) into set dst_id. */
FG_API int fg_concat_sets(fg_ctx *ctx, const int *set_ids, int nsets, int dst_id);

/* ---- sparsity of sparse(row,col,val) for the COO of src/Assembling_Operators.jl:14-22 --- */
/* Node-level CSC pattern of one set: column j = sorted union of DOM_g over all g containing j. */
FG_API int fg_pattern(fg_ctx *ctx, int set_id, int64_t *nnz_nodes);
/* Same, but the witness Gauss points come from another set (a halo-extended superset on a
 * multi-GPU slab, so that columns shared by two ranks get the same layout on both). */
FG_API int fg_pattern_from(fg_ctx *ctx, int set_id, int points_set_id, int64_t *nnz_nodes);
/* DOF-level CSC (1-based Int64) for field dimension D: colptr[nnodes*D+1], rowval[nnz*D*D]. */
FG_API int fg_get_pattern(fg_ctx *ctx, int set_id, int D, int64_t *colptr, int64_t *rowval);

/* Column block [col_lo, col_hi) (0-based node columns) of the pattern / of the assembled values: the block a slab owns.  colptr
 * ((col_hi-col_lo)*D + 1 entries) is 1-based relative to the block, row ids are shifted by row_offset nodes (local -> global
 * numbering), so the blocks of all ranks concatenate into one global SparseMatrixCSC (SURVEY 8e "Return to Julia").
 * nnz_dof = DOF-level entries of the block; colptr / rowval may be NULL (size query). */
FG_API int fg_get_pattern_cols(fg_ctx *ctx, int set_id, int D, int64_t col_lo, int64_t col_hi, int64_t row_offset, int64_t *nnz_dof,
                               int64_t *colptr, int64_t *rowval);
FG_API int fg_get_values_cols(fg_ctx *ctx, int set_id, int64_t col_lo, int64_t col_hi, double *nzval);

/* ---- kernel 3: Bilinear_Assembler / Linear_Assembler, src/Assembling_Operators.jl:81-165 - */
/* nzval: nnz*D*D doubles in the order of fg_get_pattern, or NULL to keep the result on the
 * device (fg_get_values fetches it later). */
FG_API int fg_assemble_bilinear(fg_ctx *ctx, int set_id, const fg_operator *op, double *nzval);
FG_API int fg_get_values(fg_ctx *ctx, int set_id, double *nzval);
/* The same download on a second stream: returns at once, later calls (fg_assemble_linear, another set) overlap the copy; nzval is
 * complete after the next fg_sync.  No new assembly on this set before that fg_sync. */
FG_API int fg_get_values_async(fg_ctx *ctx, int set_id, double *nzval);
/* Assembly and download as one pipeline: the caller of Bilinear_Assembler (Assembling_Operators.jl:81-127) wants the values on the
 * host.  For scalar fields on the point-batched path the clusters are assembled in ranges, and each finished column block is mirrored
 * and copied out on the copy stream under the next range; otherwise the same as fg_assemble_bilinear + fg_get_values_async.
 * nzval (pinned memory for a truly asynchronous copy) is complete after the next fg_sync. */
FG_API int fg_assemble_bilinear_async(fg_ctx *ctx, int set_id, const fg_operator *op, double *nzval);
/* How the last bilinear assembly of the set was mapped: number of Gauss-point clusters reduced in
 * shared memory, Gauss points that took the direct (global-atomics) path, largest cluster union. */
FG_API int fg_assembly_stats(fg_ctx *ctx, int set_id, int64_t *nclusters, int64_t *direct_points, int32_t *max_union);
/* F: nnodes*D doubles, or NULL to keep it on the device. */
FG_API int fg_assemble_linear(fg_ctx *ctx, int set_id, const fg_operator *op, double *F);
FG_API int fg_get_vector(fg_ctx *ctx, int set_id, double *F);
/* maximum(A) of the matrix last assembled on the set as Julia's maximum(::SparseMatrixCSC) sees it (implicit zeros count):
 * the alpha = maximum(A)*100 of the penalty step, src/Assembling_Operators.jl:238, without downloading A. */
FG_API int fg_values_max(fg_ctx *ctx, int set_id, double *max_value);

/* ---- SURVEY 8(f) "next" rows --------------------------------------------------------------- */
/* f2: field evaluation at the Gauss points of a set = Get_Point_Values, src/Fields_Operations.jl:40-112:
 * values[g*D+i] = sum_a phi_a u[(dom_a-1)*D+i];  gradients[g*D*dim + d*D + i] = sum_a u[..] * d_d phi_a
 * (the memory order of Julia's SMatrix{D,dim}).  u_nodal has nnodes*D entries; either output may be NULL. */
FG_API int fg_evaluate_field(fg_ctx *ctx, int set_id, int D, const double *u_nodal, double *values, double *gradients);
/* f3: u = A \ F  for a symmetric positive definite A held as Julia's SparseMatrixCSC arrays (1-based Int64
 * colptr/rowval), Solve, src/Solvers.jl:24-31: Jacobi-preconditioned conjugate gradients on the device.
 * Stops at ||r|| <= rtol*||F|| or max_iter; reports iterations and the true relative residual. */
FG_API int fg_solve_spd(fg_ctx *ctx, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval, const double *rhs,
                        double *x, double rtol, int max_iter, int *iterations, double *rel_residual);

/* f3 with K resident: u = (sum_k scales[k] * A_k) \ rhs, A_k = the matrix last assembled on set set_ids[k] with field dimension D
 * (e.g. the domain operator and the boundary penalty operator of Linear_Problem, src/Assembling_Operators.jl:266 `A + Ap`).
 * Patterns and values stay on the device; only rhs goes up and x comes down.  scales may be NULL (all 1). */
FG_API int fg_solve_spd_sets(fg_ctx *ctx, const int *set_ids, const double *scales, int nsets, int D, const double *rhs, double *x,
                             double rtol, int max_iter, int *iterations, double *rel_residual);

/* f4: Gauss table of a set generated on the device = egauss / egauss_bound, src/Integration.jl:148-218.
 * conn is ncells x nverts (1-based Int64, column-major: Julia's `conn`); nverts selects the entity: 4 (2D) / 8 (3D)
 * cells, 2 (2D) / 4 (3D) boundary edges / faces.  gauss is pgauss(ngpts) (2 x ngpts, column-major).  Bit-identical
 * to the host table; fg_get_gauss copies it out (ngauss x (dim+2), column-major). */
FG_API int fg_generate_gauss(fg_ctx *ctx, int set_id, int64_t ncells, int nverts, const int64_t *conn, int ngpts, const double *gauss);
FG_API int fg_get_gauss(fg_ctx *ctx, int set_id, double *gs);
/* f4, FEM branch of build_space (src/FlexiGal.jl:245-246): BackgroundIntegration(iset, :FEM) = egauss_fem / egauss_bound_fem,
 * src/Integration.jl:264-385.  Same arguments as fg_generate_gauss; leaves on the device the Gauss table AND the element shape functions of
 * every Gauss point (PHI = the quad4 / hex8 / line2 functions, DPHI = inv(J)' * local gradient as the reference writes it, DOM = the cell's
 * connectivity in vertex order): the set is ready for fg_pattern / fg_assemble_* / fg_get_shapes without fg_neighbours / fg_shape_functions.
 * Its sparsity is the union of conn x conn.  FEM and EFG sets cannot be mixed in one fg_concat_sets. */
FG_API int fg_generate_fem(fg_ctx *ctx, int set_id, int64_t ncells, int nverts, const int64_t *conn, int ngpts, const double *gauss);

/* ---- device-side access for the multi-GPU plumbing (halo exchange runs in the host layer
 *      over torch.distributed/NCCL on these buffers) ------------------------------------ */
#define FG_BUF_COLPTR 0 /* int64[nnodes+1], 0-based  */
#define FG_BUF_ROWVAL 1 /* int32[nnz], 0-based       */
#define FG_BUF_NZVAL 2  /* double[nnz*D*D]           */
#define FG_BUF_FVEC 3   /* double[nnodes*D]          */
#define FG_BUF_OFFSETS 4 /* int64[ngauss+1]          */
#define FG_BUF_DOM 5    /* int32[total_L], 0-based   */
#define FG_BUF_PHI 6    /* double[total_L]           */
#define FG_BUF_DPHI 7   /* double[dim][total_L]: component-major on the device (fg_get_shapes returns [total_L][dim]) */
FG_API int fg_device_buffer(fg_ctx *ctx, int set_id, int which, void **dev_ptr, int64_t *count);

/* ---- multi-GPU inside the library: one process and one fg_ctx per GPU, NCCL bound at run time ----------- */
/* Rank 0 obtains a 128-byte id (ncclGetUniqueId) and hands it to the other ranks over its own transport; every rank then calls
 * fg_comm_init(ctx, id, rank, world) (collective). */
FG_API int fg_comm_unique_id(fg_ctx *ctx, void *id128);
FG_API int fg_comm_init(fg_ctx *ctx, const void *id128, int rank, int world);
FG_API int fg_comm_destroy(fg_ctx *ctx);
/* The exchange step of a slab partition: for every peer, the partial sums of the node columns [send_lo, send_hi) (halo columns
 * the peer owns) are sent, and what the peer sends is ADDED to the owned columns [recv_lo, recv_hi) -- ncclSend/ncclRecv in one
 * group on the context's stream plus one add kernel, no host synchronisation.  which = FG_BUF_NZVAL (values of the last bilinear
 * assembly; both ranks must have built these columns with the same layout, fg_pattern_from) or FG_BUF_FVEC.  Column ids are
 * 0-based local node ids; at most 4 peers per call. */
FG_API int fg_exchange_columns(fg_ctx *ctx, int set_id, int which, int npeers, const int *peers, const int64_t *send_lo,
                               const int64_t *send_hi, const int64_t *recv_lo, const int64_t *recv_hi);
/* value = max over all ranks (no-op without a communicator): the global maximum(A) of the penalty step. */
FG_API int fg_comm_allreduce_max(fg_ctx *ctx, double *value);

/* Library/ABI version and the compute capability it was built for (100 = sm_100a). */
FG_API int fg_version(int *abi, int *sm);

#ifdef __cplusplus
}
#endif
#endif /* FLEXIGAL_GPU_H */
