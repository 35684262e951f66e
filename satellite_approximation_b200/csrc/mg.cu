// Geometric multigrid preconditioner for the masked 5-point operator (DESIGN.md "Multigrid").
//
//   z = M^-1 r   by one symmetric V(nu, nu)-cycle:  damped-Jacobi pre-smoothing from a zero iterate, residual,
//   full-weighting restriction, recursion, bilinear prolongation + correction, damped-Jacobi post-smoothing.
//
// Grid hierarchy: vertex-centred coarsening by 2 -- coarse cell (I, J) sits on fine cell (2I, 2J) and is an unknown of
// the coarse problem iff that fine cell is an unknown (mask injection).  Every level re-discretises the same
// unscaled 5-point operator (diagonal = neighbour count, off-diagonals -1 between unknowns, Dirichlet zero at every
// other cell), the restriction is the transpose of the bilinear prolongation (weights [1 2 1; 2 4 2; 1 2 1] / 4), which
// is the consistent scaling for unscaled operators, and pre- and post-smoother are the same symmetric iteration: the
// cycle is a symmetric positive definite operator, as CG requires.  All vectors of all levels follow the solver's
// convention -- zero outside the unknown set -- so that masks are only ever read at the cell being written.
//
// Components thinner than the coarse spacing drop out of the coarse grids; they sit close to Dirichlet data, where the
// smoother alone converges fast.  Measured on 30 % cloud-like masks: 8-10 CG iterations to 1e-6 instead of ~420
// with the Jacobi preconditioner.
#include "common.cuh"
#include "tile.cuh"

#include <algorithm>

namespace satfill {

constexpr double MG_OMEGA = 0.8;  // damped Jacobi: optimal smoothing factor 0.6 for the 5-point operator

__device__ __forceinline__ double lv_diag(const Level& lv, int64_t r, int64_t c)
{
    return lv.fixed_diag ? 4.0 : fmax(diag_of(r, c, lv.rows, lv.cols), 1.0);
}

// ---- hierarchy construction -------------------------------------------------------------------------------------

// coarse umask(I, J) = fine umask(2I, 2J); per-tile activity flags and unknown count.  One CTA per coarse tile.
//
// It also writes 1 / d of the coarse operator the red-black cycle uses (Level::winv).  Re-discretising the 5-point
// operator on the injected mask puts every Dirichlet boundary on a coarse grid point, i.e. up to one fine cell too far
// out; measured on cloud-like masks that mismatch costs almost half of the convergence rate (tools/mg_prototype.py:
// 11 -> 8 CG iterations).  The correction keeps the operator symmetric (only the diagonal changes): the diagonal is the
// sum of the conductances of the four arms of the coarse cell, in units of a regular coarse arm --
//     the fine cell half way along the arm is an unknown: 1 (whether the arm ends on a coarse unknown or on a boundary
//                                                           a full coarse spacing away);
//     it is known: the boundary sits at half the spacing -> 2;
//     it lies outside the image (Poisson only: no neighbour there, poisson.cpp:187-190), or the arm's end does: 0.
//
// shift > 0 (row-decomposed scenes, dist.cu): `fmask` is the RAW level-0 mask (non-zero = invalid) and the level is built
// from it directly -- coarse (I, J) <-> level-0 (I << shift, J << shift), a cell of the next finer level (i, j) <-> level-0
// (i << (shift - 1), j << (shift - 1)) -- which is the same mask injection, but needs no finer level to exist beyond the
// rows this rank owns.  laplace: the border rule of the unknown set (k_build_unknown_set).  The grid covers the tiles
// [first_tile, first_tile + gridDim.x); only tiles of rows [own_lo, own_hi) are flagged / counted.
__global__ void __launch_bounds__(CG_THREADS) k_coarsen_mask(const uint8_t* __restrict__ fmask, int64_t fpitch,
    int64_t frows, int64_t fcols, int fixed, uint8_t* __restrict__ cmask, int64_t crows, int64_t ccols, int64_t cpitch,
    int tiles_x, int32_t* __restrict__ tile_flags, unsigned long long* __restrict__ count64, uint32_t* __restrict__ tbits,
    uint32_t* __restrict__ tbitsT, float* __restrict__ winv, int shift, int64_t rows0, int64_t cols0, int laplace, int first_tile,
    int own_lo, int own_hi)
{
    // unknown of the next finer level at (i, j) (in range by construction of the callers' tests)
    auto fine_unknown = [&](int64_t i, int64_t j) -> bool {
        if (shift == 0)
            return fmask[i * fpitch + j] != 0;
        const int64_t R = i << (shift - 1), C = j << (shift - 1);
        if (R >= rows0 || C >= cols0 || !fmask[R * fpitch + C])
            return false;
        return !(laplace && (R == 0 || R == rows0 - 1 || C == 0 || C == cols0 - 1));
    };
    __shared__ int warp_cnt[CG_BLOCK_Y];
    __shared__ unsigned scol[TILE_W];
    if (threadIdx.y == 0)
        scol[threadIdx.x] = 0;
    __syncthreads();
    unsigned colbits = 0;
    const int tile = first_tile + (int)blockIdx.x;
    int tx = tile % tiles_x, ty = tile / tiles_x;
    int64_t c = (int64_t)tx * TILE_W + threadIdx.x;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int64_t r = (int64_t)ty * TILE_H + threadIdx.y + j * CG_BLOCK_Y;
        uint8_t m = 0;
        if (r < crows && c < ccols)
            m = fine_unknown(2 * r, 2 * c) ? 1 : 0;
        cmask[r * cpitch + c] = m;
        {
            float d = 0.f;
            if (m) {
                const int64_t fr = 2 * r, fc = 2 * c;
                const int dr[4] = { -1, 1, 0, 0 }, dc[4] = { 0, 0, -1, 1 };
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    int64_t mr = fr + dr[a], mc = fc + dc[a], er = fr + 2 * dr[a], ec = fc + 2 * dc[a];
                    bool mid_in = mr >= 0 && mr < frows && mc >= 0 && mc < fcols;
                    bool end_in = er >= 0 && er < frows && ec >= 0 && ec < fcols;
                    if (!mid_in)
                        d += fixed ? 2.f : 0.f;
                    else if (!fine_unknown(mr, mc))
                        d += 2.f;
                    else
                        d += (end_in || fixed) ? 1.f : 0.f;
                }
                d = d < 1.f ? 1.f : d;
            }
            winv[r * cpitch + c] = m ? 1.f / d : 0.f;
        }
        cnt += m;
        unsigned word = __ballot_sync(0xffffffffu, m);
        if (threadIdx.x == 0)
            tbits[((size_t)(ty + 1) * (tiles_x + 2) + tx + 1) * 32 + threadIdx.y + j * CG_BLOCK_Y] = word;
        colbits |= (unsigned)m << (threadIdx.y + j * CG_BLOCK_Y);
    }
    atomicOr(&scol[threadIdx.x], colbits);
    for (int o = 16; o; o >>= 1)
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (threadIdx.x == 0)
        warp_cnt[threadIdx.y] = cnt;
    __syncthreads();
    if (threadIdx.y == 0)
        tbitsT[((size_t)(ty + 1) * (tiles_x + 2) + tx + 1) * 32 + threadIdx.x] = scol[threadIdx.x];
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int total = 0;
        for (int w = 0; w < CG_BLOCK_Y; ++w)
            total += warp_cnt[w];
        const bool own = ty >= own_lo && ty < own_hi;
        tile_flags[tile] = own && total > 0;
        if (own && total > 0)
            atomicAdd(count64, (unsigned long long)total);
    }
}

void free_hierarchy(sa_scene* s)
{
    for (sa_level_store& L : s->coarse) {
        cudaFree(L.umask_alloc);
        cudaFree(L.tile_list);
        cudaFree(L.d_counters);
        cudaFree(L.tbits);
        cudaFree(L.x);
        cudaFree(L.b);
        cudaFree(L.t);
        cudaFree(L.winv);
    }
    s->coarse.clear();
    s->hierarchy_built = false;
}

// Allocation depends on the scene extents only; the masks / tile lists are rebuilt whenever the mask changes.
static int alloc_hierarchy(sa_scene* s, const sa_options& o)
{
    sa_ctx* ctx = s->ctx;
    int max_levels = o.mg_levels > 0 ? o.mg_levels : MAX_LEVELS;
    if (max_levels > MAX_LEVELS)
        max_levels = MAX_LEVELS;
    int64_t rows = s->rows, cols = s->cols;
    for (int l = 1; l < max_levels; ++l) {
        int64_t crows = (rows + 1) / 2, ccols = (cols + 1) / 2;
        if (crows < 3 || ccols < 3)
            break;
        sa_level_store L;
        L.lv.rows = crows;
        L.lv.cols = ccols;
        L.lv.pitch = round_up(ccols + 1, TILE_W);
        L.rows_p = round_up(crows, TILE_H);
        L.lv.plane = (L.rows_p + 2) * L.lv.pitch;
        L.lv.tiles_x = (int)(L.lv.pitch / TILE_W);
        L.lv.tiles_y = (int)(L.rows_p / TILE_H);
        L.lv.fixed_diag = s->problem == SA_LAPLACE;
        size_t vec = (size_t)L.lv.plane * s->nbands * sizeof(double);
        SA_CUDA(ctx, cudaMalloc(&L.umask_alloc, (size_t)L.lv.plane));
        SA_CUDA(ctx, cudaMemsetAsync(L.umask_alloc, 0, (size_t)L.lv.plane, ctx->stream));
        SA_CUDA(ctx, cudaMalloc(&L.tile_list, sizeof(int32_t) * 3 * (size_t)L.lv.tiles_x * L.lv.tiles_y));
        SA_CUDA(ctx, cudaMalloc(&L.d_counters, sizeof(int32_t) * 4 + sizeof(unsigned long long)));
        SA_CUDA(ctx, cudaMalloc(&L.x, vec));
        SA_CUDA(ctx, cudaMalloc(&L.b, vec));
#if SATFILL_LEGACY_VARIANTS  // the first-generation cycles' third vector; the product cycle keeps x and b only
        SA_CUDA(ctx, cudaMalloc(&L.t, vec));
        SA_CUDA(ctx, cudaMemsetAsync(L.t, 0, vec, ctx->stream));
#endif
        // cleared once; afterwards a mask change scrubs them through the old tile lists (cg.cu: scrub_work_vectors)
        SA_CUDA(ctx, cudaMemsetAsync(L.x, 0, vec, ctx->stream));
        SA_CUDA(ctx, cudaMemsetAsync(L.b, 0, vec, ctx->stream));
        SA_CUDA(ctx, cudaMalloc(&L.winv, (size_t)L.lv.plane * sizeof(float)));
        SA_CUDA(ctx, cudaMemsetAsync(L.winv, 0, (size_t)L.lv.plane * sizeof(float), ctx->stream));
        size_t words = (size_t)(L.lv.tiles_x + 2) * (L.lv.tiles_y + 2) * 32;
        L.tb_words = words;
        SA_CUDA(ctx, cudaMalloc(&L.tbits, 2 * words * sizeof(uint32_t)));
        SA_CUDA(ctx, cudaMemsetAsync(L.tbits, 0, 2 * words * sizeof(uint32_t), ctx->stream));
        L.lv.umask = L.umask_alloc + L.lv.pitch;
        L.lv.tile_list = L.tile_list;
        L.lv.tile_yx = L.tile_list + 2 * (size_t)L.lv.tiles_x * L.lv.tiles_y;
        L.lv.tbits = L.tbits;
        L.lv.tbitsT = L.tbits + words;
        L.lv.tb_stride = L.lv.tiles_x + 2;
        L.lv.winv = L.winv + L.lv.pitch;
        s->coarse.push_back(L);
        rows = crows;
        cols = ccols;
    }
    return SA_OK;
}

int build_hierarchy(sa_scene* s, const sa_options& o)
{
    sa_ctx* ctx = s->ctx;
    if (s->coarse.empty())
        SA_TRY(alloc_hierarchy(s, o));
    const uint8_t* fmask = s->mask0(s->umask);
    int64_t fpitch = s->pitch, frows = s->rows, fcols = s->cols;
    dim3 block(CG_BLOCK_X, CG_BLOCK_Y);
    // A row-decomposed scene that indexes only its own rows (dist.cu: dist_prepare_window) builds every level straight from
    // the raw mask: the split levels on the rank's rows (+ one tile row either side), the replicated ones everywhere.
    const bool windowed = s->dist_windowed;
    int level = 0;
    for (sa_level_store& L : s->coarse) {
        ++level;
        int n_tiles = L.lv.tiles_x * L.lv.tiles_y;
        int32_t* flags = L.tile_list + n_tiles;
        unsigned long long* count64 = reinterpret_cast<unsigned long long*>(L.d_counters + 4);
        SA_CUDA(ctx, cudaMemsetAsync(L.d_counters, 0, sizeof(int32_t) * 4 + sizeof(unsigned long long), ctx->stream));
        uint8_t* cmask = L.umask_alloc + L.lv.pitch;
        int own_lo = 0, own_hi = L.lv.tiles_y, win_lo = 0, win_hi = L.lv.tiles_y;
        if (windowed && level < (int)s->dl.size()) {
            own_lo = (int)(s->dl[(size_t)level].row_lo / TILE_H);
            own_hi = (int)(s->dl[(size_t)level].row_hi / TILE_H);
            win_lo = std::max(own_lo - 1, 0);
            win_hi = std::max(std::min(own_hi + 1, L.lv.tiles_y), win_lo);
        }
        const int first = win_lo * L.lv.tiles_x, count = (win_hi - win_lo) * L.lv.tiles_x;
        if (count > 0) {
            if (windowed)
                SA_LAUNCH(ctx, k_coarsen_mask, count, block, 0, s->mask0(s->mask), s->pitch, frows, fcols, L.lv.fixed_diag, cmask,
                    L.lv.rows, L.lv.cols, L.lv.pitch, L.lv.tiles_x, flags, count64, L.tbits, L.tbits + L.tb_words, L.winv + L.lv.pitch,
                    level, s->rows, s->cols, s->problem == SA_LAPLACE ? 1 : 0, first, own_lo, own_hi);
            else
                SA_LAUNCH(ctx, k_coarsen_mask, count, block, 0, fmask, fpitch, frows, fcols, L.lv.fixed_diag, cmask, L.lv.rows,
                    L.lv.cols, L.lv.pitch, L.lv.tiles_x, flags, count64, L.tbits, L.tbits + L.tb_words, L.winv + L.lv.pitch, 0,
                    s->rows, s->cols, 0, first, own_lo, own_hi);
        }
        SA_TRY(compact_tile_flags(ctx, flags, count, L.lv.tiles_x, L.tile_list, L.tile_list + 2 * n_tiles, L.d_counters, first));
        fmask = cmask;
        fpitch = L.lv.pitch;
        frows = L.lv.rows;
        fcols = L.lv.cols;
    }
    SA_CUDA(ctx, cudaGetLastError());
    // one read-back for all levels
    struct rb {
        int32_t c[4];
        unsigned long long n;
    };
    rb* h = (rb*)ctx->pinned;
    for (size_t l = 0; l < s->coarse.size(); ++l)
        SA_CUDA(ctx, cudaMemcpyAsync(&h[l], s->coarse[l].d_counters, sizeof(rb), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t l = 0; l < s->coarse.size(); ++l) {
        s->coarse[l].lv.n_tiles = h[l].c[0];
        s->coarse[l].n_unknowns = (int64_t)h[l].n;
    }
    s->hierarchy_built = true;
    return SA_OK;
}

#if SATFILL_LEGACY_VARIANTS
// ---- cycle kernels -------------------------------------------------------------------------------------------------

// FIRST: x_out = omega * b / d (one damped-Jacobi sweep from a zero iterate, pointwise).
// else : x_out = x_in + omega * (b - A x_in) / d.
// DOT  : additionally accumulate b . x_out into rz[slot] (level 0, last post-smoothing sweep: b is the CG residual).
template <bool FIRST, bool DOT>
__global__ void __launch_bounds__(CG_THREADS) k_mg_smooth(Level lv, const double* __restrict__ x_in,
    const double* __restrict__ b, double* __restrict__ x_out, BandScalars* __restrict__ scal, int slot)
{
    __shared__ double sp[TILE_H + 2][SP];
    __shared__ double s_red[CG_BLOCK_Y];
    if (scal[blockIdx.y].done)
        return;
    int tile = lv.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
    int64_t boff = (int64_t)blockIdx.y * lv.plane;
    const double* bb = b + boff;
    double* xo = x_out + boff;
    if (!FIRST) {
        const double* xi = x_in + boff;
        stage_tile<1>(
            sp, lv.umask, r0, c0, lv.pitch, [&](int64_t idx, double* v) { v[0] = xi[idx]; },
            [&](const double* v, int64_t, int64_t, int64_t, bool) { return v[0]; });
        __syncthreads();
    }
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y + 1, lc = threadIdx.x + 1;
        int64_t r = r0 + lr - 1, c = c0 + lc - 1;
        int64_t idx = r * lv.pitch + c;
        if (lv.umask[idx]) {
            double d = lv_diag(lv, r, c);
            double bv = bb[idx];
            double xn;
            if (FIRST) {
                xn = MG_OMEGA * bv / d;
            } else {
                double xc = sp[lr][lc];
                double ax = d * xc - (sp[lr - 1][lc] + sp[lr + 1][lc] + sp[lr][lc - 1] + sp[lr][lc + 1]);
                xn = xc + MG_OMEGA * (bv - ax) / d;
            }
            xo[idx] = xn;
            if (DOT)
                acc += bv * xn;
        }
    }
    if (DOT) {
        double t = block_sum(acc, s_red);
        if (threadIdx.x == 0 && threadIdx.y == 0 && t != 0.0)
            atomicAdd(&scal[blockIdx.y].rz[slot], t);
    }
}

// t = b - A x on the unknowns of the level
__global__ void __launch_bounds__(CG_THREADS) k_mg_residual(Level lv, const double* __restrict__ x,
    const double* __restrict__ b, double* __restrict__ t, const BandScalars* __restrict__ scal)
{
    __shared__ double sp[TILE_H + 2][SP];
    if (scal[blockIdx.y].done)
        return;
    int tile = lv.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
    int64_t boff = (int64_t)blockIdx.y * lv.plane;
    const double* xb = x + boff;
    stage_tile<1>(
        sp, lv.umask, r0, c0, lv.pitch, [&](int64_t idx, double* v) { v[0] = xb[idx]; },
        [&](const double* v, int64_t, int64_t, int64_t, bool) { return v[0]; });
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y + 1, lc = threadIdx.x + 1;
        int64_t r = r0 + lr - 1, c = c0 + lc - 1;
        int64_t idx = r * lv.pitch + c;
        if (lv.umask[idx]) {
            double d = lv_diag(lv, r, c);
            double xc = sp[lr][lc];
            double ax = d * xc - (sp[lr - 1][lc] + sp[lr + 1][lc] + sp[lr][lc - 1] + sp[lr][lc + 1]);
            t[boff + idx] = b[boff + idx] - ax;
        }
    }
}

// b_c(I, J) = sum_{di, dj in -1..1} w(di) w(dj) t_f(2I + di, 2J + dj),  w = (1/2, 1, 1/2)  (= P^T t_f).
// Runs over the active tiles of the COARSE level.
__global__ void __launch_bounds__(CG_THREADS) k_mg_restrict(Level lc, Level lf, const double* __restrict__ tf,
    double* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    if (scal[blockIdx.y].done)
        return;
    int tile = lc.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lc.tiles_x) * TILE_H, c0 = (int64_t)(tile % lc.tiles_x) * TILE_W;
    const double* f = tf + (int64_t)blockIdx.y * lf.plane;
    double* out = bc + (int64_t)blockIdx.y * lc.plane;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int64_t I = r0 + threadIdx.y + j * CG_BLOCK_Y, J = c0 + threadIdx.x;
        int64_t cidx = I * lc.pitch + J;
        if (lc.umask[cidx]) {
            const double* p = f + 2 * I * lf.pitch + 2 * J;
            double ul = (I | J) ? p[-lf.pitch - 1] : 0.0;  // (-1, -1) of band 0 lies before the allocation
            double up = 0.5 * ul + p[-lf.pitch] + 0.5 * p[-lf.pitch + 1];
            double mid = 0.5 * p[-1] + p[0] + 0.5 * p[1];
            double dn = 0.5 * p[lf.pitch - 1] + p[lf.pitch] + 0.5 * p[lf.pitch + 1];
            out[cidx] = 0.5 * up + mid + 0.5 * dn;
        }
    }
}

// x_f += P e_c (bilinear), on the unknowns of the fine level.  Runs over the active tiles of the FINE level.
__global__ void __launch_bounds__(CG_THREADS) k_mg_prolong(Level lf, Level lc, double* __restrict__ xf,
    const double* __restrict__ ec, const BandScalars* __restrict__ scal)
{
    if (scal[blockIdx.y].done)
        return;
    int tile = lf.tile_list[blockIdx.x];
    int64_t r0 = (int64_t)(tile / lf.tiles_x) * TILE_H, c0 = (int64_t)(tile % lf.tiles_x) * TILE_W;
    double* x = xf + (int64_t)blockIdx.y * lf.plane;
    const double* e = ec + (int64_t)blockIdx.y * lc.plane;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int64_t r = r0 + threadIdx.y + j * CG_BLOCK_Y, c = c0 + threadIdx.x;
        int64_t idx = r * lf.pitch + c;
        if (lf.umask[idx]) {
            int64_t I = r >> 1, J = c >> 1;
            const double* p = e + I * lc.pitch + J;
            double v;
            if ((r & 1) == 0)
                v = (c & 1) == 0 ? p[0] : 0.5 * (p[0] + p[1]);
            else
                v = (c & 1) == 0 ? 0.5 * (p[0] + p[lc.pitch]) : 0.25 * (p[0] + p[1] + p[lc.pitch] + p[lc.pitch + 1]);
            x[idx] += v;
        }
    }
}

// ---- the cycle ------------------------------------------------------------------------------------------------------

namespace {

struct LevelVecs {
    Level lv;
    int64_t units;  // unknowns x bands of the level
    double* x;  // iterate (level 0: z)
    double* b;  // right-hand side (level 0: the CG residual r)
    double* t;  // scratch
};

}  // namespace

int apply_vcycle(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands)
{
    sa_ctx* ctx = s->ctx;
    const int nb = s->nbands;
    const int nu = o.mg_smooth > 0 ? o.mg_smooth : 2;
    std::vector<LevelVecs> L;
    L.push_back({ fine_level(s), s->n_unknowns * live_bands, s->plane0(s->z, 0), s->plane0(s->r, 0), s->plane0(s->t, 0) });
    for (sa_level_store& c : s->coarse) {
        if (c.lv.n_tiles == 0)
            break;
        L.push_back({ c.lv, c.n_unknowns * live_bands, c.x + c.lv.pitch, c.b + c.lv.pitch, c.t + c.lv.pitch });
    }
    const int nl = (int)L.size();
    dim3 block(CG_BLOCK_X, CG_BLOCK_Y);
    BandScalars* scal = s->scal;

    // `sweeps` damped-Jacobi sweeps on level l; the first one may start from a zero iterate.  Ping-pongs between x and
    // t and returns with the result in x.  rz_slot >= 0: the last sweep also accumulates b.x into rz[rz_slot].
    auto smooth = [&](int l, int sweeps, bool from_zero, int rz_slot) -> int {
        LevelVecs& V = L[l];
        dim3 grid((unsigned)V.lv.n_tiles, (unsigned)nb);
        double* cur = V.x;
        double* oth = V.t;
        int writes = sweeps;
        if (from_zero) {
            // sweep 1 writes without reading an iterate: choose its target so that the last sweep writes V.x
            double* target = (writes % 2 == 1) ? V.x : V.t;
            kt.begin(KC_SMOOTH, V.units);
            if (sweeps == 1 && rz_slot >= 0)
                SA_LAUNCH(ctx, (k_mg_smooth<true, true>), grid, block, 0, V.lv, nullptr, V.b, target, scal, rz_slot);
            else
                SA_LAUNCH(ctx, (k_mg_smooth<true, false>), grid, block, 0, V.lv, nullptr, V.b, target, scal, 0);
            kt.end();
            cur = target;
            oth = (target == V.x) ? V.t : V.x;
            --writes;
        }
        for (int k = 0; k < writes; ++k) {
            bool last = k == writes - 1;
            kt.begin(KC_SMOOTH, V.units);
            if (last && rz_slot >= 0)
                SA_LAUNCH(ctx, (k_mg_smooth<false, true>), grid, block, 0, V.lv, cur, V.b, oth, scal, rz_slot);
            else
                SA_LAUNCH(ctx, (k_mg_smooth<false, false>), grid, block, 0, V.lv, cur, V.b, oth, scal, 0);
            kt.end();
            double* tmp = cur;
            cur = oth;
            oth = tmp;
        }
        if (cur != V.x) {  // odd number of read-modify sweeps: the result sits in t -> swap the level's buffers
            V.t = V.x;
            V.x = cur;
        }
        return SA_OK;
    };

    const bool fused = nu == 2 && !o.mg_unfused;
    // descend
    for (int l = 0; l < nl - 1; ++l) {
        if (fused) {
            kt.begin(l == 0 ? KC_MG_DOWN : KC_MG_DOWN_COARSE, L[l].units);
            SA_TRY(launch_mg_down(ctx, L[l].lv, L[l + 1].lv, nb, L[l].b, L[l].x, L[l + 1].b, scal));
            kt.end();
            continue;
        }
        SA_TRY(smooth(l, nu, true, -1));
        LevelVecs& F = L[l];
        LevelVecs& C = L[l + 1];
        dim3 gf((unsigned)F.lv.n_tiles, (unsigned)nb), gc((unsigned)C.lv.n_tiles, (unsigned)nb);
        kt.begin(KC_TRANSFER, F.units);
        SA_LAUNCH(ctx, k_mg_residual, gf, block, 0, F.lv, F.x, F.b, F.t, scal);
        kt.end();
        kt.begin(KC_TRANSFER, C.units);
        SA_LAUNCH(ctx, k_mg_restrict, gc, block, 0, C.lv, F.lv, F.t, C.b, scal);
        kt.end();
    }
    // coarsest level: smooth hard (the grid is tiny)
    {
        int l = nl - 1;
        int sweeps = nl == 1 ? 2 * nu : 2 * nu + 28;
        SA_TRY(smooth(l, sweeps, true, nl == 1 ? rz_slot : -1));
    }
    // ascend
    for (int l = nl - 2; l >= 0; --l) {
        LevelVecs& F = L[l];
        LevelVecs& C = L[l + 1];
        if (fused) {
            kt.begin(l == 0 ? KC_MG_UP : KC_MG_UP_COARSE, F.units);
            SA_TRY(launch_mg_up(ctx, F.lv, C.lv, nb, F.x, F.b, C.x, F.t, scal, l == 0 ? rz_slot : -1));
            kt.end();
            double* tmp = F.x;  // the result sits in t: swap the level's buffers
            F.x = F.t;
            F.t = tmp;
            continue;
        }
        dim3 gf((unsigned)F.lv.n_tiles, (unsigned)nb);
        kt.begin(KC_TRANSFER, F.units);
        SA_LAUNCH(ctx, k_mg_prolong, gf, block, 0, F.lv, C.lv, F.x, C.x, scal);
        kt.end();
        SA_TRY(smooth(l, nu, false, l == 0 ? rz_slot : -1));
    }
    SA_CUDA(ctx, cudaGetLastError());
    // level 0 must end in s->z, which k_direction reads: if the ping-pong left it in s->t, swap the scene's buffers
    if (L[0].x != s->plane0(s->z, 0)) {
        double* tmp = s->z;
        s->z = s->t;
        s->t = tmp;
    }
    // coarse levels: persist swapped roles
    for (int l = 1; l < nl; ++l) {
        sa_level_store& c = s->coarse[l - 1];
        if (L[l].x != c.x + c.lv.pitch) {
            double* tmp = c.x;
            c.x = c.t;
            c.t = tmp;
        }
    }
    return SA_OK;
}

#else
int apply_vcycle(sa_scene* s, const sa_options&, KernelTimer&, int, int)
{
    return fail(s->ctx, SA_BAD_ARGUMENT, "SA_MG_JACOBI64 needs a library built with SATFILL_LEGACY_VARIANTS");
}
#endif  // SATFILL_LEGACY_VARIANTS

}  // namespace satfill
