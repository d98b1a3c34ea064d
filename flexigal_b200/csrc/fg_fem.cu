// fg_fem.cu -- SURVEY 8(f) f4, FEM branch of build_space (src/FlexiGal.jl:245-246): mixed FEM / EFG spaces.
//
// BackgroundIntegration(iset, :FEM) = egauss_fem / egauss_bound_fem (src/Integration.jl:264-385): at every Gauss point of a
// background cell the shape functions are the cell's own Lagrange functions -- quad4 / hex8 in the volume, line2 / quad4 on
// boundary entities --, the "neighbour list" is the cell's connectivity in ITS vertex order (not ascending), and gs is the usual
// (coords, weight, |J|) table.  One thread per Gauss point builds all of it; the set then feeds the same assembly kernels as an
// EFG set.  The sparsity of such a set is the union of conn x conn over the cells: built from DOM by sorting the (column,row)
// pairs of one Gauss point per cell (the support-overlap witness search of fg_pattern.cu describes EFG supports, not elements).
#include "fg_internal.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

// kind: 0 quad4 in 2D, 1 hex8 in 3D, 2 line2 on a 2D boundary, 3 quad4 on a 3D boundary
template <int KIND>
__global__ void __launch_bounds__(128) k_fem_table(int64_t ncells, int l, int64_t nnodes, const double *__restrict__ x, const int64_t *__restrict__ conn,
                                                   const double *__restrict__ gauss /* [xi_0, w_0, xi_1, w_1, ..] */, double *__restrict__ gs,
                                                   int64_t *__restrict__ off, int *__restrict__ dom, double *__restrict__ phi, double *__restrict__ dphi) {
    constexpr int DIM = (KIND == 0 || KIND == 2) ? 2 : 3;
    constexpr int NV = KIND == 0 ? 4 : (KIND == 1 ? 8 : (KIND == 2 ? 2 : 4));
    constexpr int PD = KIND == 0 ? 2 : (KIND == 1 ? 3 : (KIND == 2 ? 1 : 2));       // parametric dimension
    int npts = 1;
    for (int d = 0; d < PD; ++d) npts *= l;
    const int64_t ng = ncells * npts, tl = ng * NV;
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g == 0) off[ng] = tl;
    if (g >= ng) return;
    const int64_t e = g / npts;
    const int q = (int)(g - e * npts);
    const int i = q % l, j = (q / l) % l, k = q / (l * l);        // i fastest (Integration.jl:282,303,340,355)
    const double xi = gauss[2 * i], eta = PD > 1 ? gauss[2 * j] : 0.0, zeta = PD > 2 ? gauss[2 * k] : 0.0;
    double w = gauss[2 * i + 1];
    if (PD > 1) w = w * gauss[2 * j + 1];
    if (PD > 2) w = w * gauss[2 * k + 1];
    double N[NV], dN[PD][NV];
    if constexpr (KIND == 2) {                 // fem_shape_line2, :220-224
        N[0] = 0.5 * (1 - xi); N[1] = 0.5 * (1 + xi);
        dN[0][0] = -0.5; dN[0][1] = 0.5;
    } else if constexpr (KIND == 0 || KIND == 3) {      // fem_shape_quad4, :227-239
        N[0] = 0.25 * (1 - xi) * (1 - eta); N[1] = 0.25 * (1 + xi) * (1 - eta); N[2] = 0.25 * (1 + xi) * (1 + eta); N[3] = 0.25 * (1 - xi) * (1 + eta);
        dN[0][0] = -0.25 * (1 - eta); dN[0][1] = 0.25 * (1 - eta); dN[0][2] = 0.25 * (1 + eta); dN[0][3] = -0.25 * (1 + eta);
        dN[1][0] = -0.25 * (1 - xi); dN[1][1] = -0.25 * (1 + xi); dN[1][2] = 0.25 * (1 + xi); dN[1][3] = 0.25 * (1 - xi);
    } else {                         // fem_shape_hex8, :242-258
        const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const double fx = 1 + sx[m] * xi, fy = 1 + sy[m] * eta, fz = 1 + sz[m] * zeta;
            N[m] = (1.0 / 8) * fx * fy * fz;
            dN[0][m] = sx[m] * (1.0 / 8) * fy * fz;
            dN[1][m] = sy[m] * (1.0 / 8) * fx * fz;
            dN[2][m] = sz[m] * (1.0 / 8) * fx * fy;
        }
    }
    int id[NV];
    double xe[NV][DIM];
#pragma unroll
    for (int m = 0; m < NV; ++m) {
        id[m] = (int)(conn[m * ncells + e] - 1);                   // conn is ncells x nverts, column-major, 1-based
#pragma unroll
        for (int d = 0; d < DIM; ++d) xe[m][d] = x[d * nnodes + id[m]];
    }
    double xg[DIM], T[PD][DIM];                                    // T[p][d] = sum_m dN[p][m] xe[m][d]  (rows of J / the tangents)
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < NV; ++m) s += N[m] * xe[m][d];
        xg[d] = s;
#pragma unroll
        for (int p = 0; p < PD; ++p) {
            double t = 0.0;
#pragma unroll
            for (int m = 0; m < NV; ++m) t += dN[p][m] * xe[m][d];
            T[p][d] = t;
        }
    }
    double meas, G[DIM][NV];                                       // G[d][m] = d phi_m / d x_d
    // The reference applies inv(J)' (the TRANSPOSE of the inverse) to the local gradient, with J[p][d] = d x_d / d xi_p (:287-295,
    // :308-317).  For the axis-aligned cells of create_model (also under separable maps) J is diagonal and this is the physical
    // gradient; for sheared cells it is not (that would be inv(J) itself).  Reproduced literally: parity with the reference.
    if constexpr (KIND == 0) {
        const double a = T[0][0], b = T[0][1], c = T[1][0], d2 = T[1][1];
        meas = a * d2 - b * c;
        const double inv = 1.0 / meas;
#pragma unroll
        for (int m = 0; m < NV; ++m) {
            G[0][m] = (d2 * dN[0][m] - c * dN[1][m]) * inv;          // inv(J)' = [d -c; -b a] / det
            G[1][m] = (-b * dN[0][m] + a * dN[1][m]) * inv;
        }
    } else if constexpr (KIND == 1) {          // :313-317
        const double c00 = T[1][1] * T[2][2] - T[1][2] * T[2][1];
        const double c01 = T[1][2] * T[2][0] - T[1][0] * T[2][2];
        const double c02 = T[1][0] * T[2][1] - T[1][1] * T[2][0];
        meas = T[0][0] * c00 + T[0][1] * c01 + T[0][2] * c02;
        const double inv = 1.0 / meas;
        double iJ[3][3];             // iJ = J^-1 with J[p][d] = T[p][d]; the reference's gradient is inv(J)' dN: grad_d = sum_p iJ[p][d] dN[p]
        iJ[0][0] = c00 * inv; iJ[1][0] = c01 * inv; iJ[2][0] = c02 * inv;
        iJ[0][1] = (T[0][2] * T[2][1] - T[0][1] * T[2][2]) * inv;
        iJ[1][1] = (T[0][0] * T[2][2] - T[0][2] * T[2][0]) * inv;
        iJ[2][1] = (T[0][1] * T[2][0] - T[0][0] * T[2][1]) * inv;
        iJ[0][2] = (T[0][1] * T[1][2] - T[0][2] * T[1][1]) * inv;
        iJ[1][2] = (T[0][2] * T[1][0] - T[0][0] * T[1][2]) * inv;
        iJ[2][2] = (T[0][0] * T[1][1] - T[0][1] * T[1][0]) * inv;
#pragma unroll
        for (int m = 0; m < NV; ++m)
#pragma unroll
            for (int d = 0; d < 3; ++d) G[d][m] = iJ[0][d] * dN[0][m] + iJ[1][d] * dN[1][m] + iJ[2][d] * dN[2][m];
    } else if constexpr (KIND == 2) {          // boundary edge in 2D: ds, tangential gradient  (:341-349)
        const double tx = T[0][0], ty = T[0][1];
        meas = sqrt(tx * tx + ty * ty);
#pragma unroll
        for (int m = 0; m < NV; ++m) { G[0][m] = dN[0][m] * tx / (meas * meas); G[1][m] = dN[0][m] * ty / (meas * meas); }
    } else {                         // boundary face in 3D: dA, surface gradient through the metric tensor  (:356-376)
        double g11 = 0, g12 = 0, g22 = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { g11 += T[0][d] * T[0][d]; g12 += T[0][d] * T[1][d]; g22 += T[1][d] * T[1][d]; }
        const double detG = g11 * g22 - g12 * g12;
        meas = sqrt(detG);
#pragma unroll
        for (int m = 0; m < NV; ++m) {
            const double c1 = (g22 / detG) * dN[0][m] + (-g12 / detG) * dN[1][m];
            const double c2 = (-g12 / detG) * dN[0][m] + (g11 / detG) * dN[1][m];
#pragma unroll
            for (int d = 0; d < DIM; ++d) G[d][m] = c1 * T[0][d] + c2 * T[1][d];
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) gs[d * ng + g] = xg[d];
    gs[DIM * ng + g] = w;
    gs[(DIM + 1) * ng + g] = meas;
    off[g] = g * NV;
#pragma unroll
    for (int m = 0; m < NV; ++m) {
        dom[g * NV + m] = id[m];
        phi[g * NV + m] = N[m];
#pragma unroll
        for (int d = 0; d < DIM; ++d) dphi[d * tl + g * NV + m] = G[d][m];
    }
}

extern "C" int fg_generate_fem(fg_ctx *ctx, int set_id, int64_t ncells, int nverts, const int64_t *conn, int ngpts, const double *gauss) {
    if (!ctx) return FG_E_ARG;
    NodeSet &N = ctx->nodes;
    if (!N.ready) return fg_fail(ctx, FG_E_STATE, "fg_generate_fem: call fg_set_nodes first");
    if (set_id < 0 || ncells < 0 || !conn || !gauss || ngpts < 1 || ngpts > 4) return fg_fail(ctx, FG_E_ARG, "fg_generate_fem: bad arguments");
    const int dim = N.dim;
    int kind;
    if (dim == 2 && nverts == 4) kind = 0; else if (dim == 3 && nverts == 8) kind = 1;
    else if (dim == 2 && nverts == 2) kind = 2; else if (dim == 3 && nverts == 4) kind = 3;
    else return fg_fail(ctx, FG_E_ARG, "fg_generate_fem: %d vertices per entity in %dD is not a cell or a boundary entity", nverts, dim);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    int npts = 1;
    for (int d = 0; d < (kind == 0 || kind == 3 ? 2 : (kind == 1 ? 3 : 1)); ++d) npts *= ngpts;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) { S = new GaussSet(); ctx->sets[set_id] = S; }
    const int64_t ng = ncells * npts, tl = ng * nverts;
    if (ng >= (1ll << 31)) return fg_fail(ctx, FG_E_ARG, "fg_generate_fem: too many Gauss points");
    S->n = ng; S->have_nb = S->have_sf = false; S->pat.ready = false; S->gbins_ready = false; S->cl.ready = S->cl.binned = false; S->nbcl.binned = false;
    S->nsub = 0; S->nb_clustered = false; S->dom_pending = false;
    S->total_L = tl; S->max_L = nverts;
    FG_CUDA(ctx, S->gs.reserve(ng * (dim + 2))); FG_CUDA(ctx, S->off.reserve(ng + 1)); FG_CUDA(ctx, S->cnt.reserve(ng));
    FG_CUDA(ctx, S->dom.reserve(tl)); FG_CUDA(ctx, S->phi.reserve(tl)); FG_CUDA(ctx, S->dphi.reserve(tl * dim));
    FG_CUDA(ctx, S->singular_dev.reserve(1));
    FG_CUDA(ctx, cudaMemsetAsync(S->singular_dev.p, 0, sizeof(int), ctx->stream));
    S->fem.clear();
    if (ng == 0) { FG_CUDA(ctx, cudaMemsetAsync(S->off.p, 0, sizeof(int64_t), ctx->stream)); S->have_nb = S->have_sf = true; return FG_OK; }
    FG_CUDA(ctx, ctx->ids64.reserve(ncells * nverts));
    FG_CUDA(ctx, ctx->stage.reserve(sizeof(double) * 2 * ngpts));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->ids64.p, conn, sizeof(int64_t) * ncells * nverts, cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->stage.p, gauss, sizeof(double) * 2 * ngpts, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)((ng + 127) / 128);
    const double *dg = (const double *)ctx->stage.p;
    switch (kind) {
    case 0: k_fem_table<0><<<grid, 128, 0, ctx->stream>>>(ncells, ngpts, N.n, N.x.p, ctx->ids64.p, dg, S->gs.p, S->off.p, S->dom.p, S->phi.p, S->dphi.p); break;
    case 1: k_fem_table<1><<<grid, 128, 0, ctx->stream>>>(ncells, ngpts, N.n, N.x.p, ctx->ids64.p, dg, S->gs.p, S->off.p, S->dom.p, S->phi.p, S->dphi.p); break;
    case 2: k_fem_table<2><<<grid, 128, 0, ctx->stream>>>(ncells, ngpts, N.n, N.x.p, ctx->ids64.p, dg, S->gs.p, S->off.p, S->dom.p, S->phi.p, S->dphi.p); break;
    default: k_fem_table<3><<<grid, 128, 0, ctx->stream>>>(ncells, ngpts, N.n, N.x.p, ctx->ids64.p, dg, S->gs.p, S->off.p, S->dom.p, S->phi.p, S->dphi.p); break;
    }
    FG_CHECK_LAUNCH(ctx);
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // conn / gauss are the caller's again
    GaussSet::FemSeg seg; seg.first = 0; seg.count = ng; seg.npts = npts;
    S->fem.push_back(seg);
    S->have_nb = S->have_sf = true;
    return FG_OK;
}

// ---- sparsity of a FEM set (or a concatenation of FEM sets): sorted unique (column, row) pairs of one Gauss point per cell ------------
__global__ void k_fem_pairs(int64_t ncells, int64_t first, int npts, const int64_t *__restrict__ off, const int *__restrict__ dom,
                            unsigned long long *__restrict__ keys, int64_t key0, int nv) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ncells * nv * nv) return;
    const int64_t c = t / (nv * nv);
    const int r = (int)(t - c * nv * nv), a = r / nv, b = r - a * nv;
    const int64_t o0 = off[first + c * npts];
    keys[key0 + t] = ((unsigned long long)(unsigned)dom[o0 + b] << 32) | (unsigned)dom[o0 + a];       // column (trial) high, row (test) low
}

__global__ void k_fem_csc(int64_t nuniq, int64_t nnodes, const unsigned long long *__restrict__ keys, int64_t *__restrict__ colptr, int *__restrict__ rowval) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > nuniq) return;
    const int64_t prev = k == 0 ? -1 : (int64_t)(keys[k - 1] >> 32);
    const int64_t cur = k == nuniq ? nnodes : (int64_t)(keys[k] >> 32);
    for (int64_t j = prev + 1; j <= cur; ++j) colptr[j] = k;          // colptr[j] = first entry of a column >= j; colptr[nnodes] = nuniq
    if (k < nuniq) rowval[k] = (int)(keys[k] & 0xffffffffull);
}

int fg_impl_pattern_fem(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    Pattern &P = S.pat;
    P.ready = false; P.D = 0; P.colptr_host.clear(); ++P.version;
    S.cl.tab_ready = false;
    const int nv = S.max_L;
    int64_t npairs = 0;
    for (auto &seg : S.fem) npairs += (seg.count / seg.npts) * (int64_t)nv * nv;
    FG_CUDA(ctx, P.colptr.reserve(N.n + 1));
    DevBuf<unsigned long long> &kin = ctx->fem_keys_in, &kout = ctx->fem_keys_out;
    FG_CUDA(ctx, kin.reserve(npairs + 1)); FG_CUDA(ctx, kout.reserve(npairs + 1));
    int64_t key0 = 0;
    for (auto &seg : S.fem) {
        const int64_t ncells = seg.count / seg.npts, n = ncells * nv * nv;
        if (n > 0) {
            k_fem_pairs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ncells, seg.first, seg.npts, S.off.p, S.dom.p, kin.p, key0, nv);
            FG_CHECK_LAUNCH(ctx);
        }
        key0 += n;
    }
    int64_t nuniq = 0;
    if (npairs > 0) {
        if (npairs >= (1ll << 31)) return fg_fail(ctx, FG_E_CAPACITY, "fg_pattern: too many element pairs");
        int bits = 33; while ((1ll << (bits - 32)) < N.n && bits < 64) ++bits;
        size_t b1 = 0, b2 = 0;
        FG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, b1, kin.p, kout.p, (int)npairs, 0, bits, ctx->stream));
        FG_CUDA(ctx, ctx->flags.reserve(64));
        int64_t *dcount = reinterpret_cast<int64_t *>(ctx->flags.p + 32);
        FG_CUDA(ctx, cub::DeviceSelect::Unique(nullptr, b2, kout.p, kin.p, dcount, (int)npairs, ctx->stream));
        FG_CUDA(ctx, ctx->temp.reserve(std::max(b1, b2)));
        FG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->temp.p, b1, kin.p, kout.p, (int)npairs, 0, bits, ctx->stream));
        FG_CUDA(ctx, cub::DeviceSelect::Unique(ctx->temp.p, b2, kout.p, kin.p, dcount, (int)npairs, ctx->stream));
        FG_CUDA(ctx, cudaMemcpyAsync(&nuniq, dcount, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    P.nnz = nuniq;
    FG_CUDA(ctx, P.rowval.reserve(nuniq));
    k_fem_csc<<<(unsigned)((nuniq + 1 + 255) / 256), 256, 0, ctx->stream>>>(nuniq, N.n, kin.p, P.colptr.p, P.rowval.p);
    FG_CHECK_LAUNCH(ctx);
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P.max_col = 0;
    P.ready = true; P.tmap_ready = false;
    return FG_OK;
}
