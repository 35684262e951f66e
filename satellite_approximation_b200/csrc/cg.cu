// Matrix-free preconditioned conjugate gradient on the block-sparse masked grid.  Replaces the Eigen sparse assembly
// and solve of the reference (laplace.cpp:58-114, poisson.cpp:179-270 -> Eigen ConjugateGradient.h:30-85 with
// DiagonalPreconditioner, BasicPreconditioners.h:39-86).
//
// The system, for p in the unknown set U:
//     d_p x_p - sum_{q in N4(p) & U} x_q = b_p,      d_p = number of in-image 4-neighbours of p (4 in the interior)
//     b_p = [Poisson: sum_{q in N4(p)} (g_p - g_q)] + sum_{q in N4(p) \ U} f_q
// x lives in the image plane itself (known cells keep f, unknown cells hold the iterate); r, p are planes that are
// zero outside U.  One CG iteration is two kernels over the active tiles of all bands (DESIGN.md "CG kernels"):
//     direction:  beta = rz_k / rz_{k-1};  p' = z + beta p  (ping-pong buffer; z = r/d for Jacobi);  pq = p'.Ap'
//     update:     alpha = rz_k / pq;  x += alpha p';  r -= alpha A p'  (A p' recomputed, never stored);
//                 rr = |r|^2,  rz = r.(r/d)
// A p is never stored (66 B per unknown and Jacobi iteration instead of the 89 B of the textbook five-pass formulation).
// All scalars stay on the device (BandScalars); the host only polls the `done` flags through pinned memory.
//
// This file holds the solver driver (solve_scene, prepare_solve, the scrub of the work vectors) and the FIRST-GENERATION
// kernels (k_init_guess, k_residual, k_direction, k_update: tile + halo staged through shared memory), kept as
// cg_variant = 1, the reference the strip kernels of cg_strip.cu are tested against.  The product path is cg_strip.cu.
#include "common.cuh"
#include "tile.cuh"

#include <cstdlib>

namespace satfill {

// ---------------------------------------------------------------------------------------------------------------
// Set-up: initial iterate, residual, right-hand-side norm.
// ---------------------------------------------------------------------------------------------------------------

// x0: Laplace solve() starts from zero (IterativeSolverBase.h:357-360); Poisson from the replacement image
// (poisson.cpp:239, 257).
#if SATFILL_LEGACY_VARIANTS
template <bool POISSON>
__global__ void __launch_bounds__(CG_THREADS) k_init_guess(Level lv, double* __restrict__ u, const double* __restrict__ g)
{
    int tile = lv.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
    int64_t boff = (int64_t)blockIdx.y * lv.plane;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int64_t idx = (r0 + threadIdx.y + j * CG_BLOCK_Y) * lv.pitch + c0 + threadIdx.x;
        if (lv.umask[idx])
            u[boff + idx] = POISSON ? g[boff + idx] : 0.0;
    }
}

// r = b - A x0 and |b|^2, |r|^2, r.(r/d) in one pass over u (and g): laplace.cpp:71-94 / poisson.cpp:241-251 for b,
// ConjugateGradient.h:38-61 for the rest.
template <bool POISSON>
__global__ void __launch_bounds__(CG_THREADS) k_residual(Level lv, const double* __restrict__ u, const double* __restrict__ g,
    double* __restrict__ rvec, float* __restrict__ rf, BandScalars* __restrict__ scal)
{
    __shared__ double s_red[CG_BLOCK_Y];
    int tile = lv.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
    int64_t boff = (int64_t)blockIdx.y * lv.plane;
    const double* ub = u + boff;
    const double* gb = POISSON ? g + boff : nullptr;
    double b2 = 0.0, r2 = 0.0, rz = 0.0;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int64_t r = r0 + threadIdx.y + j * CG_BLOCK_Y, c = c0 + threadIdx.x;
        int64_t idx = r * lv.pitch + c;
        if (lv.umask[idx]) {
            double d = diag_of(r, c, lv.rows, lv.cols);
            double un = ub[idx - lv.pitch], us = ub[idx + lv.pitch], uw = ub[idx - 1], ue = ub[idx + 1];
            double kn = lv.umask[idx - lv.pitch] ? 0.0 : un, ks = lv.umask[idx + lv.pitch] ? 0.0 : us;
            double kw = lv.umask[idx - 1] ? 0.0 : uw, ke = lv.umask[idx + 1] ? 0.0 : ue;
            double div = 0.0;
            if (POISSON)  // sum over in-image neighbours of (g_p - g_q); g is zero outside the image
                div = d * gb[idx] - (gb[idx - lv.pitch] + gb[idx + lv.pitch] + gb[idx - 1] + gb[idx + 1]);
            double b = div + (kn + ks + kw + ke);
            double res = div + (un + us + uw + ue) - d * ub[idx];
            b2 += b * b;
            r2 += res * res;
            rz += res * res * inv_diag(r, c, lv.rows, lv.cols);
            rvec[boff + idx] = res;
            if (rf)
                rf[boff + idx] = (float)res;
        }
    }
    double t;
    t = block_sum(b2, s_red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
        atomicAdd(&scal[blockIdx.y].bnorm2, t);
    t = block_sum(r2, s_red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
        atomicAdd(&scal[blockIdx.y].rr[0], t);
    t = block_sum(rz, s_red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
        atomicAdd(&scal[blockIdx.y].rz[0], t);
}

#endif  // SATFILL_LEGACY_VARIANTS

__global__ void k_finalize_setup(BandScalars* scal, int nbands, double tol, int mg)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbands)
        return;
    BandScalars& s = scal[b];
    if (mg)
        s.rz[0] = 0.0;  // r.z of iteration 0 comes from the first V-cycle, not from the Jacobi scaling
    if (s.bnorm2 == 0.0) {  // ConjugateGradient.h:43-49
        s.zero_rhs = 1;
        s.done = 1;
        s.iters = 0;
        s.rr_exit = 0.0;
        return;
    }
    double thr = tol * tol * s.bnorm2;  // ConjugateGradient.h:50-51
    s.thr = thr < DBL_MIN ? DBL_MIN : thr;
    if (s.rr[0] < s.thr) {  // ConjugateGradient.h:52-57
        s.done = 1;
        s.iters = 0;
        s.rr_exit = s.rr[0];
    }
}

// Zero right-hand side: Eigen returns x = 0 whatever the guess was (ConjugateGradient.h:43-49).
__global__ void __launch_bounds__(CG_THREADS) k_zero_unknowns(Level lv, double* __restrict__ u,
    const BandScalars* __restrict__ scal)
{
    if (!scal[blockIdx.y].zero_rhs)  // (almost always: the grid is a few CTAs per band that walk the tile list)
        return;
    const int64_t boff = (int64_t)blockIdx.y * lv.plane;
    for (int i = blockIdx.x; i < lv.n_tiles; i += gridDim.x) {
        const int tile = lv.tile_list[i];
        const int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
#pragma unroll
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            const int64_t idx = (r0 + threadIdx.y + j * CG_BLOCK_Y) * lv.pitch + c0 + threadIdx.x;
            if (lv.umask[idx])
                u[boff + idx] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The two kernels of one CG iteration.
// ---------------------------------------------------------------------------------------------------------------

#if SATFILL_LEGACY_VARIANTS
// k_direction: p' = z + beta p, pq = p'.Ap'.   JACOBI: z = r / d computed on the fly (zin = r).  Otherwise zin = z.
template <bool JACOBI, typename ZT>
__global__ void __launch_bounds__(CG_THREADS, 8) k_direction(Level lv, const ZT* __restrict__ zin,
    const double* __restrict__ p_old, double* __restrict__ p_new, BandScalars* __restrict__ scal, int k)
{
    __shared__ double sp[TILE_H + 2][SP];
    __shared__ double s_red[CG_BLOCK_Y];
    BandScalars& sc = scal[blockIdx.y];
    if (sc.done)
        return;
    int slot = k & 3;
    bool lead = blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0;
    if (k > 0 && sc.rr[slot] < sc.thr) {  // ConjugateGradient.h:72-73 (strict <), tested one launch later
        if (lead) {
            sc.rr_exit = sc.rr[slot];
            sc.iters = k - 1;
            __threadfence();
            sc.done = 1;
        }
        return;
    }
    double beta = k > 0 ? sc.rz[slot] / sc.rz[(k - 1) & 3] : 0.0;  // ConjugateGradient.h:77-79
    if (lead) {  // recycle the slot two iterations ahead
        int z2 = (k + 2) & 3;
        sc.rz[z2] = 0.0;
        sc.rr[z2] = 0.0;
        sc.pq[z2] = 0.0;
    }
    int tile = lv.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
    int64_t boff = (int64_t)blockIdx.y * lv.plane;
    const ZT* zb = zin + boff;
    const double* pb = p_old + boff;
    double* pn = p_new + boff;
    int64_t rows = lv.rows, cols = lv.cols;
    stage_tile<2>(
        sp, lv.umask, r0, c0, lv.pitch,
        [&](int64_t idx, double* v) {
            v[0] = (double)zb[idx];
            v[1] = pb[idx];
        },
        [&](const double* v, int64_t idx, int64_t r, int64_t c, bool interior) {
            double z = v[0];
            if (JACOBI)
                z *= inv_diag(r, c, rows, cols);
            double pv = z + beta * v[1];  // ConjugateGradient.h:80
            if (interior)
                pn[idx] = pv;
            return pv;
        });
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y + 1, lc = threadIdx.x + 1;
        double pc = sp[lr][lc];
        double d = diag_of(r0 + lr - 1, c0 + lc - 1, rows, cols);
        double q = d * pc - (sp[lr - 1][lc] + sp[lr + 1][lc] + sp[lr][lc - 1] + sp[lr][lc + 1]);
        acc += pc * q;  // p is zero outside U, so no mask is needed here
    }
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
        atomicAdd(&sc.pq[slot], t);
}

// k_update: alpha = rz / pq; x += alpha p; r -= alpha A p; new |r|^2 and (JACOBI) r.(r/d) into slot k+1.
// RF: also write the residual as float for the red-black cycle (mg_rb.cu), which then never reads a double.
template <bool JACOBI, bool RF>
__global__ void __launch_bounds__(CG_THREADS, 8) k_update(Level lv, double* __restrict__ u, const double* __restrict__ p,
    double* __restrict__ rvec, float* __restrict__ rf, BandScalars* __restrict__ scal, int k)
{
    __shared__ double sp[TILE_H + 2][SP];
    __shared__ double s_red[CG_BLOCK_Y];
    BandScalars& sc = scal[blockIdx.y];
    if (sc.done)
        return;
    int slot = k & 3, next = (k + 1) & 3;
    double alpha = sc.rz[slot] / sc.pq[slot];  // ConjugateGradient.h:68
    int tile = lv.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
    int64_t boff = (int64_t)blockIdx.y * lv.plane;
    const double* pb = p + boff;
    double* ub = u + boff;
    double* rb = rvec + boff;
    int64_t rows = lv.rows, cols = lv.cols;
    // x and r of the thread's own cells: issued before the staging of p so that all loads are in flight together
    const int64_t base = (r0 + threadIdx.y) * lv.pitch + c0 + threadIdx.x;
    const int64_t rstep = (int64_t)CG_BLOCK_Y * lv.pitch;
    uint8_t m[ROWS_PER_THREAD];
    double xv[ROWS_PER_THREAD], rv[ROWS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j)
        m[j] = lv.umask[base + j * rstep];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        xv[j] = m[j] ? ub[base + j * rstep] : 0.0;
        rv[j] = m[j] ? rb[base + j * rstep] : 0.0;
    }
    stage_tile<1>(
        sp, lv.umask, r0, c0, lv.pitch, [&](int64_t idx, double* v) { v[0] = pb[idx]; },
        [&](const double* v, int64_t, int64_t, int64_t, bool) { return v[0]; });
    __syncthreads();
    double r2 = 0.0, rz = 0.0;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y + 1, lc = threadIdx.x + 1;
        if (m[j]) {
            int64_t r = r0 + lr - 1, c = c0 + lc - 1;
            double pc = sp[lr][lc];
            double d = diag_of(r, c, rows, cols);
            double q = d * pc - (sp[lr - 1][lc] + sp[lr + 1][lc] + sp[lr][lc - 1] + sp[lr][lc + 1]);
            ub[base + j * rstep] = xv[j] + alpha * pc;  // ConjugateGradient.h:69
            double rn = rv[j] - alpha * q;               // ConjugateGradient.h:70
            rb[base + j * rstep] = rn;
            if (RF)
                rf[boff + base + j * rstep] = (float)rn;
            r2 += rn * rn;
            if (JACOBI)
                rz += rn * rn * inv_diag(r, c, rows, cols);
        }
    }
    double t = block_sum(r2, s_red);
    if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
        atomicAdd(&sc.rr[next], t);
    if (JACOBI) {
        t = block_sum(rz, s_red);
        if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
            atomicAdd(&sc.rz[next], t);
    }
}

#endif  // SATFILL_LEGACY_VARIANTS

// Multigrid loop: test the stop rule right after the update of iteration k - 1 (one thread per band), so that a band
// that has converged does not pay for another V-cycle before k_direction would notice.
__global__ void k_check_converged(BandScalars* scal, int nbands, int k)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbands)
        return;
    BandScalars& s = scal[b];
    if (!s.done && s.rr[k & 3] < s.thr) {  // ConjugateGradient.h:72-73
        s.rr_exit = s.rr[k & 3];
        s.iters = k - 1;
        s.done = 1;
    }
}

// The host polls the per-band scalars through pinned host memory that this kernel writes (cudaMallocHost memory is
// device-addressable under unified addressing): no copy-engine transfer inside the loop, so the poll never queues
// behind the band transfers that the host-pointer entry points keep in flight on the other streams (api.cu).
__global__ void k_publish_scalars(const BandScalars* __restrict__ scal, int nbands, BandScalars* __restrict__ host)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbands)
        host[b] = scal[b];
}

__global__ void k_final_check(BandScalars* scal, int nbands, int k_end)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbands)
        return;
    BandScalars& s = scal[b];
    if (s.done)
        return;
    double rr = s.rr[k_end & 3];
    s.rr_exit = rr;
    if (k_end > 0 && rr < s.thr) {
        s.iters = k_end - 1;
        s.done = 1;
    } else {
        s.iters = k_end;  // ran out of iterations (ConjugateGradient.h:65)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host driver
// ---------------------------------------------------------------------------------------------------------------
Level fine_level(const sa_scene* s)
{
    Level lv {};
    lv.rows = s->rows;
    lv.cols = s->cols;
    lv.pitch = s->pitch;
    lv.plane = s->plane;
    lv.tiles_x = s->tiles_x;
    lv.tiles_y = s->tiles_y;
    lv.n_tiles = s->n_active_tiles;
    lv.umask = s->mask0(s->umask);
    lv.tile_list = s->tile_list;
    lv.tile_yx = s->tile_list + 2 * (size_t)s->tiles_x * s->tiles_y;
    lv.fixed_diag = s->problem == SA_LAPLACE;
    lv.tbits = s->tbits;
    lv.tbitsT = s->tbits + s->tb_words;
    lv.tb_stride = s->tiles_x + 2;
    return lv;
}

// Zero the work vectors at the unknowns of the mask they were last used with (the tile lists and bit masks of that mask
// are still in place: sa_scene_set_mask only replaces the mask itself).
// `lean`: the solve that follows is the strip CG + red-black cycle again, which reads r, the float copy of r, the red
// half-planes and the coarse right-hand sides only through the unknown bits (k_update2, k_rb_down, k_rb_up mask every
// cell they use): only the vectors whose halos are read unmasked -- p, z, the coarse corrections -- have to be clean.
// r is then marked stale for the day a Jacobi solve (which reads r with its halo) follows.
static int scrub_work_vectors(sa_scene* s, bool lean)
{
    sa_ctx* ctx = s->ctx;
    const int nb = s->nbands;
    ScrubPlanes P {};
    if (!lean)
        P.d[P.nd++] = s->plane0(s->r, 0);
    else
        s->stale_r = s->stale_rb = true;
    if (s->work_dirty & WORK_PF) {
        P.f[P.nf++] = (float*)s->p[0] + s->pitch;
        P.f[P.nf++] = (float*)s->p[1] + s->pitch;
    } else {
        P.d[P.nd++] = s->plane0(s->p[0], 0);
        P.d[P.nd++] = s->plane0(s->p[1], 0);
    }
    if ((s->work_dirty & WORK_J64) && s->z && s->t) {
        P.d[P.nd++] = s->plane0(s->z, 0);
        P.d[P.nd++] = s->plane0(s->t, 0);
    }
    if ((s->work_dirty & WORK_RB) && s->z) {
        P.f[P.nf++] = (float*)s->z + s->pitch;                                   // z of the red-black cycle
        if (!lean) {
            P.f[P.nf++] = (float*)s->z + (int64_t)s->plane * nb + s->pitch;       // float copy of the residual
            if (s->t)  // (first-generation cycle only: the product library does not allocate it)
                P.h[P.nh++] = (float*)s->t + (s->pitch >> 1);                     // red half of the iterate
        }
    }
    SA_TRY(launch_scrub(ctx, fine_level(s), nb, P));
    if (s->hierarchy_built)
        for (sa_level_store& c : s->coarse) {
            if (c.lv.n_tiles == 0)
                break;
            ScrubPlanes Q {};
            if ((s->work_dirty & WORK_J64) && c.t) {
                Q.d[Q.nd++] = c.x + c.lv.pitch;
                Q.d[Q.nd++] = c.b + c.lv.pitch;
                Q.d[Q.nd++] = c.t + c.lv.pitch;
            }
            if (s->work_dirty & WORK_RB) {
                Q.f[Q.nf++] = (float*)c.x + c.lv.pitch;  // prolongation reads the coarse correction unmasked
                if (!lean) {
                    Q.f[Q.nf++] = (float*)c.b + c.lv.pitch;
                    if (c.t)
                        Q.h[Q.nh++] = (float*)c.t + (c.lv.pitch >> 1);
                }
            }
            SA_TRY(launch_scrub(ctx, c.lv, nb, Q));
        }
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

// The two buffers of the search direction (ping-pong).  The product path keeps p in float, so they are sized for float
// planes -- half of what round 1 allocated, 12.6 GB of a C3 scene -- and grow (contents dropped: they are only ever needed
// zero outside the unknown set) the first time a path wants double planes (Jacobi, the first-generation CG).  Never less
// than one double plane: precondition_scene hands its result back through p[0].
int ensure_p(sa_scene* s, size_t elem_bytes)
{
    sa_ctx* ctx = s->ctx;
    const size_t need = std::max((size_t)s->plane * s->nbands * elem_bytes, (size_t)s->plane * sizeof(double));
    if (s->p_bytes >= need)
        return SA_OK;
    if (s->p[0] || s->p[1])
        SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(s->p[0]);
    cudaFree(s->p[1]);
    s->p[0] = s->p[1] = nullptr;
    s->p_bytes = 0;
    SA_CUDA(ctx, cudaMalloc(&s->p[0], need));
    SA_CUDA(ctx, cudaMalloc(&s->p[1], need));
    SA_CUDA(ctx, cudaMemsetAsync(s->p[0], 0, need, ctx->stream));
    SA_CUDA(ctx, cudaMemsetAsync(s->p[1], 0, need, ctx->stream));
    s->p_bytes = need;
    return SA_OK;
}

static int clear_p(sa_scene* s)
{
    if (s->p_bytes) {
        SA_CUDA(s->ctx, cudaMemsetAsync(s->p[0], 0, s->p_bytes, s->ctx->stream));
        SA_CUDA(s->ctx, cudaMemsetAsync(s->p[1], 0, s->p_bytes, s->ctx->stream));
    }
    return SA_OK;
}

// Clear everything the multigrid cycles write (variant switch, or contents unknown).
static int clear_multigrid_vectors(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    size_t bytes = (size_t)s->plane * s->nbands * sizeof(double);
    if (s->z)
        SA_CUDA(ctx, cudaMemsetAsync(s->z, 0, bytes, ctx->stream));
    if (s->t)
        SA_CUDA(ctx, cudaMemsetAsync(s->t, 0, bytes, ctx->stream));
    for (sa_level_store& c : s->coarse) {
        size_t vec = (size_t)c.lv.plane * s->nbands * sizeof(double);
        SA_CUDA(ctx, cudaMemsetAsync(c.x, 0, vec, ctx->stream));
        SA_CUDA(ctx, cudaMemsetAsync(c.b, 0, vec, ctx->stream));
        if (c.t)
            SA_CUDA(ctx, cudaMemsetAsync(c.t, 0, vec, ctx->stream));
    }
    return SA_OK;
}

// Work vectors left unscrubbed by mask changes between solves of the warp-kernel red-black path (stale_all): any other
// path reads some of them unmasked, so it gets them cleared wholesale first.
static int clear_stale_vectors(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    size_t bytes = (size_t)s->plane * s->nbands * sizeof(double);
    SA_CUDA(ctx, cudaMemsetAsync(s->r, 0, bytes, ctx->stream));
    SA_TRY(clear_p(s));
    SA_TRY(clear_multigrid_vectors(s));
    s->stale_r = s->stale_rb = s->stale_all = false;
    s->work_dirty &= ~(WORK_JACOBI | WORK_RB | WORK_J64 | WORK_PF | WORK_RBW);
    return SA_OK;
}

int ensure_indexed(sa_scene* s, int next_kind)
{
    sa_ctx* ctx = s->ctx;
    if (s->indexed)
        return SA_OK;
    // work vectors must be zero outside the unknown set (and in tiles that are never visited)
    if ((s->work_dirty & WORK_FULL) || !s->ever_indexed) {
        size_t bytes = (size_t)s->plane * s->nbands * sizeof(double);
        SA_CUDA(ctx, cudaMemsetAsync(s->r, 0, bytes, ctx->stream));
        SA_TRY(clear_p(s));
        SA_TRY(clear_multigrid_vectors(s));
        s->stale_r = s->stale_rb = s->stale_all = false;
    } else if (s->work_dirty != WORK_CLEAN) {
        // before index_scene replaces the old tile lists
        const bool rb_again = (next_kind & ~WORK_RBW) == WORK_RB && (s->work_dirty & ~(WORK_PF | WORK_RBW)) == WORK_RB && (s->work_dirty & WORK_PF);
        if (rb_again && (next_kind & WORK_RBW) && (s->work_dirty & WORK_RBW)) {
            // strip CG + warp-per-tile cycle on both sides of the mask change: NOTHING has to be scrubbed.  Every kernel of
            // that path loads a work vector only where the current mask has an unknown in the loaded piece (pair, quad,
            // halo cell), every such piece has been rewritten by the current solve before it is read (whole sectors /
            // quads, zeros at the cells that are not unknowns), and the one place that multiplies an old value by zero
            // (p at iteration 0) only needs it to be finite.  Everything is then "stale" for any other path.
            s->stale_r = s->stale_rb = s->stale_all = true;
        } else {
            SA_TRY(scrub_work_vectors(s, rb_again));
        }
    }
    s->work_dirty = WORK_CLEAN;
    // a row-decomposed scene indexes only its own rows (dist.cu)
    SA_TRY(dist_prepare_window(s, next_kind != WORK_JACOBI));
    SA_TRY(index_scene(s));
    s->ever_indexed = true;
    s->hierarchy_built = false;
    s->dist_planned = false;
    return SA_OK;
}

// ---- diagnostic: one application of the preconditioner (sa_scene_precondition) ---------------------------------------
__global__ void __launch_bounds__(256) k_mask_plane(double* __restrict__ v, const uint8_t* __restrict__ umask, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!umask[i])
            v[i] = 0.0;
}

__global__ void __launch_bounds__(256) k_narrow_plane(const double* __restrict__ v, float* __restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)v[i];
}

template <typename ZT>
__global__ void __launch_bounds__(256) k_widen_plane(const ZT* __restrict__ z, const uint8_t* __restrict__ umask,
    double* __restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = umask[i] ? (double)z[i] : 0.0;
}

static int ensure_multigrid(sa_scene* s, const sa_options& o)
{
    sa_ctx* ctx = s->ctx;
    if (!s->z) {
        size_t bytes = (size_t)s->plane * s->nbands * sizeof(double);
        SA_CUDA(ctx, cudaMalloc(&s->z, bytes));  // the red-black cycle: z (float) in the first half, the float copy of r in the second
        SA_CUDA(ctx, cudaMemsetAsync(s->z, 0, bytes, ctx->stream));
#if SATFILL_LEGACY_VARIANTS  // the second plane of the first-generation cycles (ping-pong partner / red half of the iterate)
        SA_CUDA(ctx, cudaMalloc(&s->t, bytes));
        SA_CUDA(ctx, cudaMemsetAsync(s->t, 0, bytes, ctx->stream));
#endif
    }
    if (!s->hierarchy_built)
        SA_TRY(build_hierarchy(s, o));
    return SA_OK;
}

int precondition_scene(sa_scene* s, const sa_options& o)
{
    sa_ctx* ctx = s->ctx;
#if !SATFILL_LEGACY_VARIANTS
    if (o.mg_variant != SA_MG_RB32)
        return fail(ctx, SA_BAD_ARGUMENT, "this libsatfill holds the product kernels only (SATFILL_LEGACY_VARIANTS)");
#endif
    SA_TRY(ensure_multigrid(s, o));
    SA_TRY(ensure_p(s, sizeof(float)));  // (at least one double plane: the result goes out through p[0])
    if (s->stale_all && o.mg_variant != SA_MG_RB32)
        SA_TRY(clear_stale_vectors(s));
    s->work_dirty = WORK_FULL;  // the caller's vector goes through r and p: clear everything before the next solve
    const int64_t n = (int64_t)s->rows_p * s->pitch;  // the plane without its guard rows
    const uint8_t* um = s->mask0(s->umask);
    SA_LAUNCH(ctx, k_mask_plane, 1024, 256, 0, s->plane0(s->r, 0), um, n);
    SA_CUDA(ctx, cudaMemsetAsync(s->scal, 0, sizeof(BandScalars) * s->nbands, ctx->stream));
    if (s->n_unknowns == 0) {
        SA_CUDA(ctx, cudaMemsetAsync(s->p[0], 0, (size_t)s->plane * sizeof(double), ctx->stream));
        return SA_OK;
    }
    KernelTimer kt;
    kt.ctx = ctx;
    if (o.mg_variant != SA_MG_JACOBI64) {
        SA_LAUNCH(ctx, k_narrow_plane, 1024, 256, 0, s->plane0(s->r, 0), s->rb_rf(), n);
        if (o.mg_variant == SA_MG_RB32_CTA)
            SA_TRY(apply_vcycle_rb(s, o, kt, 0, s->nbands));
        else
            SA_TRY(apply_vcycle_rbw(s, o, kt, 0, s->nbands));
        SA_LAUNCH(ctx, k_widen_plane<float>, 1024, 256, 0, s->rb_z(), um, s->plane0(s->p[0], 0), n);
    } else {
        SA_TRY(apply_vcycle(s, o, kt, 0, s->nbands));
        SA_LAUNCH(ctx, k_widen_plane<double>, 1024, 256, 0, s->plane0(s->z, 0), um, s->plane0(s->p[0], 0), n);
    }
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

// Everything a solve needs before its first kernel: the index of the mask, the multigrid hierarchy, and work vectors that
// are clean for the preconditioner about to be used.  Idempotent, so the host-pointer entry points can call it once
// before they start set-up kernels for several band windows on another stream.
int prepare_solve(sa_scene* s, const sa_options& o)
{
    sa_ctx* ctx = s->ctx;
    const bool mg = o.precond == SA_PRECOND_MULTIGRID;
    const bool rb = mg && o.mg_variant != SA_MG_JACOBI64;
    const bool strip = o.cg_variant == 0;
    const bool rbw = rb && strip && o.mg_variant == SA_MG_RB32;
    SA_TRY(ensure_p(s, strip && rb ? sizeof(float) : sizeof(double)));  // (grows before anything below touches p)
    SA_TRY(ensure_indexed(s, !mg ? WORK_JACOBI : (rb && strip ? (WORK_RB | (rbw ? WORK_RBW : 0)) : WORK_J64)));
    if (s->stale_all && !rbw)
        SA_TRY(clear_stale_vectors(s));
    if (s->stale_r && !(rb && strip)) {  // Jacobi and the double cycle read r with its halo
        SA_CUDA(ctx, cudaMemsetAsync(s->r, 0, (size_t)s->plane * s->nbands * sizeof(double), ctx->stream));
        s->stale_r = false;
    }
    if (s->n_unknowns == 0)
        return SA_OK;
    if (mg)
        SA_TRY(ensure_multigrid(s, o));
    {
        // the two multigrid variants lay the same allocations out differently: clear them when the variant changes
        const int kind = !mg ? WORK_JACOBI : (rb ? WORK_RB : WORK_J64);
        const int other = kind == WORK_RB ? WORK_J64 : (kind == WORK_J64 ? WORK_RB : 0);
        if ((s->work_dirty & other) || (kind == WORK_J64 && s->stale_rb)) {
            SA_TRY(clear_multigrid_vectors(s));
            s->work_dirty &= ~other;
            s->stale_rb = false;
        }
        // ... and so do the two precisions of the search direction
        const bool pf_now = o.cg_variant == 0 && rb;
        if ((s->work_dirty & (WORK_JACOBI | WORK_RB | WORK_J64)) && ((s->work_dirty & WORK_PF) != 0) != pf_now) {
            SA_TRY(clear_p(s));
        }
        s->work_dirty = (s->work_dirty & ~(WORK_PF | WORK_RBW)) | kind | (pf_now ? WORK_PF : 0) | (rbw ? WORK_RBW : 0);
    }
    return SA_OK;
}

int solve_scene(sa_scene* s, const sa_options& o, sa_stats* stats)
{
    sa_ctx* ctx = s->ctx;
    // the band window [band0, band0 + nb): `stats` and every per-band base pointer below start at its first band
    const int nb = s->win_n(), b0 = s->band0;
    BandScalars* const scal = s->scal + b0;
    const bool poisson = s->problem == SA_POISSON;
    const bool mg = o.precond == SA_PRECOND_MULTIGRID;
    const bool rb = mg && o.mg_variant != SA_MG_JACOBI64;
    const bool strip = o.cg_variant == 0;
#if !SATFILL_LEGACY_VARIANTS
    if (!strip || (mg && o.mg_variant != SA_MG_RB32))
        return fail(ctx, SA_BAD_ARGUMENT, "this libsatfill holds the product kernels only: cg_variant = 1, SA_MG_JACOBI64 and SA_MG_RB32_CTA "
                                          "need the library built with SATFILL_LEGACY_VARIANTS (lib/libsatfill_legacy.so)");
#endif
    SA_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    SA_TRY(prepare_solve(s, o));
    const int64_t n = s->n_unknowns;
    // reference defaults: Laplace 2N (IterativeSolverBase.h:251), Poisson n/2 (poisson.cpp:207)
    int64_t max_it = o.max_iterations > 0 ? o.max_iterations : (poisson ? n / 2 : 2 * n);
    if (max_it > (int64_t)1 << 30)
        max_it = (int64_t)1 << 30;
    if (stats) {
        for (int b = 0; b < nb; ++b) {
            stats[b] = sa_stats {};
            stats[b].unknowns = n;
            stats[b].max_iterations = max_it;
            stats[b].tolerance = o.tolerance;
            stats[b].active_tiles = s->n_active_tiles;
        }
    }
    if (n == 0) {  // laplace.cpp:41-44: nothing to do
        if (stats)
            for (int b = 0; b < nb; ++b)
                stats[b].status = SA_EMPTY_MASK;
        return SA_EMPTY_MASK;
    }
    // one system split by rows across the ranks of the context's communicator (dist.cu)
    const bool dist = s->distributed && ctx->world > 1;
    if (dist && (b0 != 0 || nb != s->nbands))
        return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: band windows are not supported");
    if (mg && !rb && (b0 != 0 || nb != s->nbands))
        return fail(ctx, SA_BAD_ARGUMENT, "band windows need the Jacobi or the red-black multigrid preconditioner");
    if (dist) {
        if (!strip || (mg && !rb))
            return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: needs the strip CG kernels and the red-black cycle");
        if (!s->dist_planned || s->dist_mg != mg) {
            int pst = dist_plan_scene(s, mg);
            if (pst == SA_RETRY_UNWINDOWED) {
                // too small for the windowed set-up (or the preconditioner changed under the same mask): index the whole
                // mask on every rank, as a scene that is not split does, and slice the tile lists
                s->dist_no_window = true;
                s->indexed = false;
                s->work_dirty |= WORK_FULL;
                SA_TRY(prepare_solve(s, o));
                pst = dist_plan_scene(s, mg);
            }
            SA_TRY(pst);
        }
    }

    Level lv = dist ? dist_level(s, 0, fine_level(s)) : fine_level(s);
    const bool have_tiles = lv.n_tiles > 0;  // a rank's slice may hold no active tile
    dim3 grid((unsigned)(have_tiles ? lv.n_tiles : 1), (unsigned)nb), block(CG_BLOCK_X, CG_BLOCK_Y);
    double* u0 = s->plane0(s->u, b0);
    double* g0 = poisson ? s->plane0(s->g, b0) : nullptr;
    double* r0 = s->plane0(s->r, b0);
    // the search direction: double planes, or (strip kernels + red-black cycle) float planes in the same allocations
    const bool pf = strip && rb;
    double* pbuf[2] = { s->plane0(s->p[0], b0), s->plane0(s->p[1], b0) };
    float* pbuf_f[2] = { (float*)s->p[0] + s->pitch + (int64_t)b0 * s->plane, (float*)s->p[1] + s->pitch + (int64_t)b0 * s->plane };

    SA_CUDA(ctx, cudaMemsetAsync(scal, 0, sizeof(BandScalars) * nb, ctx->stream));
    if (strip) {
        // one pass: x0, r0 = b - A x0 from the KNOWN neighbours only (so no halo of x0 is needed), the three norms
        SA_TRY(launch_setup2(ctx, lv, nb, poisson, u0, g0, r0, rb ? s->rb_rf() : nullptr, scal));
    } else {
        if (have_tiles) {
            if (poisson)
                SA_LAUNCH_LEGACY(ctx, k_init_guess<true>, grid, block, 0, lv, u0, g0);
            else
                SA_LAUNCH_LEGACY(ctx, k_init_guess<false>, grid, block, 0, lv, u0, g0);
        }
        if (dist)  // the residual's stencil reads the iterate one row beyond the slice
            SA_TRY(dist_halo<double>(s, 0, u0, s->pitch, s->plane, 1, 1));
        if (have_tiles) {
            if (poisson)
                SA_LAUNCH_LEGACY(ctx, k_residual<true>, grid, block, 0, lv, u0, g0, r0, rb ? s->rb_rf() : nullptr, scal);
            else
                SA_LAUNCH_LEGACY(ctx, k_residual<false>, grid, block, 0, lv, u0, g0, r0, rb ? s->rb_rf() : nullptr, scal);
        }
    }
    if (dist) {  // |b|^2, |r0|^2, r0.z0 summed over the ranks, and the halo rows of the residual the first kernel reads
        if (rb)
            SA_TRY(dist_step(s, 0, DIST_VEC_RHS, s->rb_rf(), 4, s->pitch, s->plane, 3, 3, DIST_SETUP, 0, -1));
        else
            SA_TRY(dist_step(s, 0, DIST_VEC_RHS, r0, 8, s->pitch, s->plane, 1, 1, DIST_SETUP, 0, -1));
    }
    SA_LAUNCH(ctx, k_finalize_setup, (nb + 63) / 64, 64, 0, scal, nb, o.tolerance, mg ? 1 : 0);
    SA_CUDA(ctx, cudaGetLastError());
    SA_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    KernelTimer kt;
    kt.ctx = ctx;
    kt.on = o.profile != 0;
    // what the profile counts per launch: the unknowns THIS rank processes
    const int64_t n_units = dist ? (int64_t)((double)n * s->dist_unit_frac) : n;
    BandScalars* h_scal = (BandScalars*)ctx->pinned;
    BandScalars* h_scal_dev = nullptr;  // the same memory as the device addresses it (k_publish_scalars)
    SA_CUDA(ctx, cudaHostGetDevicePointer((void**)&h_scal_dev, h_scal, 0));
    // Jacobi iterations are short (two kernels): poll the flags every 32.  A multigrid iteration is a whole V-cycle:
    // poll every iteration, which also keeps the profile's per-class unit counts exact.
    const int check = o.check_every > 0 ? o.check_every : (mg ? 1 : 32);
    // The host reads the flags of batch i - 1 while batch i is already queued: the device never waits for the host to
    // notice, enqueue and launch (for a one-band window that wait was 8 % of the solve).  A batch that turns out to be
    // unnecessary costs a few dozen kernels that exit at their first instruction.  Ranks of a row-decomposed solve see the
    // same (all-reduced) flags for the same batch, so they still take every decision together.  The profile's
    // per-launch accounting keeps the in-step poll when bands can finish at different iterations (its unit counts follow
    // the live bands); with one band its events are simply read once the solve is over.
    const bool lookahead = (!kt.on || nb == 1) && nb <= 512;
    BandScalars* const h_slot[2] = { h_scal, h_scal + 512 };
    BandScalars* const h_slot_dev[2] = { h_scal_dev, h_scal_dev + 512 };
    int64_t k = 0, batch = 0;
    bool all_done = false;
    int live = nb;  // bands not done at the last poll
    // strip CG with the float search direction: x += alpha p every other pass (SATFILL_DEFER_X=0 switches it off)
    static const bool defer_env = [] { const char* e = std::getenv("SATFILL_DEFER_X"); return !e || std::atoi(e) != 0; }();
    const bool defer_x = defer_env && strip && mg && rb && pf;
    while (k < max_it && !all_done) {
        int64_t k_stop = k + check < max_it ? k + check : max_it;
        for (; k < k_stop; ++k) {
            int ki = (int)(k & 0x3fffffff);  // only k == 0 and k & 3 matter to the kernels; iters is re-based below
            const double* pin = pbuf[k & 1];
            double* pout = pbuf[(k + 1) & 1];
            const void* pin_v = pf ? (const void*)pbuf_f[k & 1] : (const void*)pin;
            void* pout_v = pf ? (void*)pbuf_f[(k + 1) & 1] : (void*)pout;
            if (mg) {
                // z = M^-1 r, rz[slot] accumulated by the cycle's last kernel
                const void* z;
                if (rb) {
                    if (o.mg_variant == SA_MG_RB32_CTA)
                        SA_TRY(apply_vcycle_rb(s, o, kt, ki & 3, live));
                    else
                        SA_TRY(apply_vcycle_rbw(s, o, kt, ki & 3, live));
                    z = s->rb_z();
                } else {
                    SA_TRY(apply_vcycle(s, o, kt, ki & 3, live));
                    z = s->plane0(s->z, b0);
                }
                float* rf = rb ? s->rb_rf() : nullptr;
                if (dist)  // r.z summed over the ranks and the halo row of z, one exchange
                    SA_TRY(dist_step(s, 0, DIST_VEC_SOL, rb ? (void*)s->rb_z() : nullptr, 4, s->pitch, s->plane, 1, 1, DIST_RZ, ki & 3, -1));
                kt.begin(KC_DIRECTION, n_units * live);
                if (strip)
                    SA_TRY(launch_direction2(ctx, lv, nb, false, z, rb, pin_v, pout_v, pf, scal, ki));
                else if (rb)
                    SA_LAUNCH_LEGACY(ctx, (k_direction<false, float>), grid, block, 0, lv, (const float*)z, pin, pout, scal, ki);
                else
                    SA_LAUNCH_LEGACY(ctx, (k_direction<false, double>), grid, block, 0, lv, (const double*)z, pin, pout, scal, ki);
                kt.end();
                if (dist)  // the halo row of p' and p'.Ap' in one exchange
                    SA_TRY(dist_step(s, 0, DIST_VEC_DIR, pf ? pout_v : (void*)pout, pf ? 4 : 8, s->pitch, s->plane, 1, 1, DIST_PQ, ki & 3,
                        (ki + 2) & 3));
                // x travels every other pass (k_update2, XM): even passes leave it alone, odd ones add both steps
                const int xm = defer_x ? ((k & 1) ? 2 : 1) : 0;
                kt.begin(xm == 1 ? KC_UPDATE_DEFERRED : KC_UPDATE, n_units * live);
                if (strip)
                    SA_TRY(launch_update2(ctx, lv, nb, false, u0, pout_v, pf, r0, rf, scal, ki, xm, pin_v));
                else if (rb)
                    SA_LAUNCH_LEGACY(ctx, (k_update<false, true>), grid, block, 0, lv, u0, pout, r0, rf, scal, ki);
                else
                    SA_LAUNCH_LEGACY(ctx, (k_update<false, false>), grid, block, 0, lv, u0, pout, r0, nullptr, scal, ki);
                kt.end();
                if (dist)  // the halo rows of the cycle's residual copy and |r|^2 in one exchange
                    SA_TRY(dist_step(s, 0, DIST_VEC_RHS, rf, 4, s->pitch, s->plane, 3, 3, DIST_RR, (ki + 1) & 3, -1));
                SA_LAUNCH(ctx, k_check_converged, (nb + 63) / 64, 64, 0, scal, nb, ki + 1);
            } else {
                kt.begin(KC_DIRECTION, n_units * live);
                if (strip)
                    SA_TRY(launch_direction2(ctx, lv, nb, true, r0, false, pin, pout, false, scal, ki));
                else
                    SA_LAUNCH_LEGACY(ctx, (k_direction<true, double>), grid, block, 0, lv, r0, pin, pout, scal, ki);
                kt.end();
                if (dist)
                    SA_TRY(dist_step(s, 0, DIST_VEC_DIR, pout, 8, s->pitch, s->plane, 1, 1, DIST_PQ, ki & 3, (ki + 2) & 3));
                kt.begin(KC_UPDATE, n_units * live);
                if (strip)
                    SA_TRY(launch_update2(ctx, lv, nb, true, u0, pout, false, r0, nullptr, scal, ki));
                else
                    SA_LAUNCH_LEGACY(ctx, (k_update<true, false>), grid, block, 0, lv, u0, pout, r0, nullptr, scal, ki);
                kt.end();
                if (dist) {  // ranks must agree on the stop: test it from the reduced norm after every iteration
                    SA_TRY(dist_step(s, 0, DIST_VEC_RHS, r0, 8, s->pitch, s->plane, 1, 1, DIST_RR_RZ, (ki + 1) & 3, -1));
                    SA_LAUNCH(ctx, k_check_converged, (nb + 63) / 64, 64, 0, scal, nb, ki + 1);
                }
            }
        }
        SA_CUDA(ctx, cudaGetLastError());
        const int slot = lookahead ? (int)(batch & 1) : 0;
        SA_LAUNCH(ctx, k_publish_scalars, (nb + 63) / 64, 64, 0, scal, nb, h_slot_dev[slot]);
        const BandScalars* seen = nullptr;
        if (lookahead) {
            SA_CUDA(ctx, cudaEventRecord(ctx->ev[4 + slot], ctx->stream));
            if (batch >= 1) {
                SA_CUDA(ctx, cudaEventSynchronize(ctx->ev[4 + (1 - slot)]));
                seen = h_slot[1 - slot];
            }
        } else {
            SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            kt.flush();
            seen = h_slot[0];
        }

        if (seen) {
            live = 0;
            for (int b = 0; b < nb; ++b)
                live += seen[b].done ? 0 : 1;
            all_done = live == 0;
        }
        ++batch;
    }
    if (defer_x)  // the step a band's last pass left behind
        SA_TRY(launch_flush_x(ctx, lv, nb, u0, pbuf_f[0], pbuf_f[1], scal));
    SA_LAUNCH(ctx, k_final_check, (nb + 63) / 64, 64, 0, scal, nb, (int)(k & 0x3fffffff));
    if (have_tiles)
        SA_LAUNCH(ctx, k_zero_unknowns, dim3(std::min(grid.x, 4u * (unsigned)ctx->sm_count), grid.y), block, 0, lv, u0, scal);
    SA_CUDA(ctx, cudaGetLastError());
    SA_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    SA_LAUNCH(ctx, k_publish_scalars, (nb + 63) / 64, 64, 0, scal, nb, h_scal_dev);
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    kt.flush();  // the per-launch events of a one-band profile that polled one batch behind (everything else: empty)
    float setup_ms = 0.f, solve_ms = 0.f;
    cudaEventElapsedTime(&setup_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&solve_ms, ctx->ev[1], ctx->ev[2]);
    int status = SA_OK;
    if (std::getenv("SATFILL_DEBUG"))
        for (int b = 0; b < nb; ++b)
            std::fprintf(stderr, "[satfill] band %d: bnorm2 %.6e thr %.3e rr_exit %.3e iters %d done %d zero_rhs %d rz %.3e %.3e %.3e %.3e\n",
                b, h_scal[b].bnorm2, h_scal[b].thr, h_scal[b].rr_exit, h_scal[b].iters, h_scal[b].done, h_scal[b].zero_rhs,
                h_scal[b].rz[0], h_scal[b].rz[1], h_scal[b].rz[2], h_scal[b].rz[3]);
    for (int b = 0; b < nb; ++b) {
        const BandScalars& sc = h_scal[b];
        int st = sc.done ? SA_OK : SA_NOT_CONVERGED;
        if (stats) {
            stats[b].iterations = sc.iters;
            stats[b].error = sc.bnorm2 > 0.0 ? sqrt(sc.rr_exit / sc.bnorm2) : 0.0;
            stats[b].solve_ms = solve_ms;
            stats[b].setup_ms = setup_ms;
            stats[b].status = st;
            for (int c = 0; c < KC_COUNT; ++c) {
                stats[b].kernel_ms[c] = kt.ms[c];
                stats[b].kernel_launches[c] = kt.n[c];
                stats[b].kernel_units[c] = kt.units[c];
            }
        }
        if (st != SA_OK)
            status = st;
    }
    if (status == SA_NOT_CONVERGED)
        fail(ctx, status, "conjugate gradient reached max_iterations before the tolerance");
    return status;
}

}  // namespace satfill
