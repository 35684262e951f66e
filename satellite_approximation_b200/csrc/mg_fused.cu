// Fused multigrid kernels: one V(2,2)-cycle touches every level with exactly two kernels.
//
//   k_mg_down:  x  = two damped-Jacobi sweeps from zero on  A x = b          (pre-smoothing)
//               t  = b - A x                                                   (residual, never stored)
//               bc = P^T t                                                     (full-weighting restriction)
//   k_mg_up:    x' = x + P e_c                                                 (bilinear prolongation + correction)
//               x' = two damped-Jacobi sweeps on A x' = b                      (post-smoothing)
//               [level 0] rz += b . x'                                         (the r.z of the CG iteration)
//
// Temporal blocking in shared memory: a CTA stages its 32 x 32 tile together with a halo as deep as the chain of
// stencil applications it fuses (3 cells going down, 2 going up), runs the sweeps on shrinking regions and writes
// only its own tile.  Results are identical to running the sweeps one kernel at a time (the neighbouring CTAs
// recompute the overlap), but a level costs  R b + W x + W bc = 18 B  and  R x + R b + R e + W x' = 26 B  per unknown
// instead of the 140 B of the seven single-sweep kernels (mg.cu keeps those for nu != 2 and for the coarsest level).
// The unknown set comes from the per-tile bit masks (Level::tbits, L2 resident), so loads of known cells are
// predicated off and no bounds checks are needed.
//
// Thread mapping: 40 x 4 threads; thread (x, y) owns column x of the staged region and rows y, y + 4, ...  All shared
// memory addresses are then a per-thread base plus compile-time offsets -- the first version of these kernels walked
// the regions with a linear index (div / mod per cell) and was instruction-issue bound at ~1 TB/s (profiles/).
// Rows of the region without any unknown are skipped by the whole CTA.
#include "common.cuh"
#include "tile.cuh"

namespace satfill {

constexpr double FW = 0.8;  // damped-Jacobi weight, same as mg.cu
constexpr int FX = 40, FY = 4, FTHREADS = FX * FY;

// Row masks of the (32 + 2H)^2 neighbourhood of tile (ty, tx): bit (col + H) of mrow[row + H] <=> cell
// (r0 + row, c0 + col) is an unknown, for row, col in [-H, 32 + H).
template <int H>
__device__ __forceinline__ void load_region_mask(const Level& lv, int ty, int tx, unsigned long long* mrow, int t)
{
    for (int row = t; row < TILE_H + 2 * H; row += FTHREADS) {
        int gr = row - H;
        int tyy = ty + (gr < 0 ? -1 : (gr >= TILE_H ? 1 : 0));
        const uint32_t* w = lv.tbits + ((size_t)(tyy + 1) * lv.tb_stride + (tx + 1)) * 32 + (gr & 31);
        unsigned long long C = w[0], L = w[-32], R = w[32];
        mrow[row] = (L >> (32 - H)) | (C << H) | ((R & ((1ull << H) - 1)) << (32 + H));
    }
}

// diagonal and its inverse; FIXED: Laplace (every unknown has four in-image neighbours)
template <bool FIXED>
__device__ __forceinline__ void diag_pair(const Level& lv, int64_t r, int64_t c, double& d, double& inv)
{
    if (FIXED) {
        d = 4.0;
        inv = 0.25;
    } else {
        int n = (r > 0) + (r < lv.rows - 1) + (c > 0) + (c < lv.cols - 1);
        d = n < 1 ? 1.0 : (double)n;
        inv = n == 4 ? 0.25 : (n == 3 ? (1.0 / 3.0) : (n == 2 ? 0.5 : 1.0));
    }
}

template <bool FIXED>
__global__ void __launch_bounds__(FTHREADS) k_mg_down(Level lf, Level lc, const double* __restrict__ b,
    double* __restrict__ x_out, double* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    constexpr int H = 3, W = TILE_W + 2 * H;  // 38
    constexpr int S = W + 1;                  // shared row stride
    __shared__ unsigned long long mrow[W];
    __shared__ double B[W * S];
    __shared__ double X1[W * S];  // sweep 1; later reused for the residual
    __shared__ double X2[W * S];  // sweep 2, stored at the 38-grid position of the cell
    if (scal[blockIdx.y].done)
        return;
    const int x = threadIdx.x, y = threadIdx.y, t = y * FX + x;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    load_region_mask<H>(lf, ty, tx, mrow, t);
    __syncthreads();
    const int64_t gc = c0 - H + x;  // global column of this thread
    {
        // all loads of the thread are issued back to back (predicated, no branches) before any of them is used
        constexpr int NK = (W + FY - 1) / FY;
        const double* bp = b + (int64_t)blockIdx.y * lf.plane + (r0 - H + y) * lf.pitch + gc;
        const int64_t step = (int64_t)FY * lf.pitch;
        double v[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            int row = y + k * FY;
            bool on = x < W && row < W && ((mrow[row < W ? row : 0] >> x) & 1);
            v[k] = on ? bp[k * step] : 0.0;
        }
        if (x < W) {
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                int row = y + k * FY;
                if (row < W) {
                    double d, inv;
                    diag_pair<FIXED>(lf, r0 - H + row, gc, d, inv);
                    B[row * S + x] = v[k];
                    X1[row * S + x] = FW * inv * v[k];
                }
            }
        }
    }
    __syncthreads();
    if (x >= 1 && x < W - 1) {  // sweep 2 on the 36 x 36 region
#pragma unroll
        for (int k = 0; k < (W + FY - 1) / FY; ++k) {
            int row = y + k * FY;
            if (row >= 1 && row < W - 1) {
                unsigned long long m = mrow[row];
                double x2 = 0.0;
                if ((m >> x) & 1) {
                    double d, inv;
                    diag_pair<FIXED>(lf, r0 - H + row, gc, d, inv);
                    const double* p = X1 + row * S + x;
                    double xc = p[0];
                    double ax = d * xc - ((p[-S] + p[S]) + (p[-1] + p[1]));
                    x2 = xc + FW * inv * (B[row * S + x] - ax);
                }
                X2[row * S + x] = x2;
            }
        }
    }
    __syncthreads();
    double* R = X1;  // residual on the 34 x 34 region (X1 is dead)
    double* xo = x_out + (int64_t)blockIdx.y * lf.plane + (r0 - H) * lf.pitch + gc;
    if (x >= 2 && x < W - 2) {
#pragma unroll
        for (int k = 0; k < (W + FY - 1) / FY; ++k) {
            int row = y + k * FY;
            if (row >= 2 && row < W - 2) {
                double res = 0.0;
                if ((mrow[row] >> x) & 1) {
                    double d, inv;
                    diag_pair<FIXED>(lf, r0 - H + row, gc, d, inv);
                    const double* p = X2 + row * S + x;
                    double xc = p[0];
                    double ax = d * xc - ((p[-S] + p[S]) + (p[-1] + p[1]));
                    res = B[row * S + x] - ax;
                    if (row >= H && row < H + TILE_H && x >= H && x < H + TILE_W)
                        xo[row * lf.pitch] = xc;  // the CTA's own tile
                }
                R[row * S + x] = res;
            }
        }
    }
    __syncthreads();
    // restriction: coarse cell (ci, cj) of this tile sits on fine tile cell (2 ci, 2 cj) = 38-grid (2 ci + 3, 2 cj + 3)
    double* bco = bc + (int64_t)blockIdx.y * lc.plane + (r0 >> 1) * lc.pitch + (c0 >> 1);
    for (int i = t; i < (TILE_H / 2) * (TILE_W / 2); i += FTHREADS) {
        int ci = i >> 4, cj = i & 15;
        int a = 2 * ci + H, c = 2 * cj + H;
        if ((mrow[a] >> c) & 1) {  // mask injection: coarse unknown <=> fine (2I, 2J) unknown
            const double* p = R + a * S + c;
            double up = 0.5 * p[-S - 1] + p[-S] + 0.5 * p[-S + 1];
            double mid = 0.5 * p[-1] + p[0] + 0.5 * p[1];
            double dn = 0.5 * p[S - 1] + p[S] + 0.5 * p[S + 1];
            bco[ci * lc.pitch + cj] = 0.5 * up + mid + 0.5 * dn;
        }
    }
}

template <bool FIXED, bool DOT>
__global__ void __launch_bounds__(FTHREADS) k_mg_up(Level lf, Level lc, const double* __restrict__ x_in,
    const double* __restrict__ b, const double* __restrict__ ec, double* __restrict__ x_out,
    BandScalars* __restrict__ scal, int slot)
{
    constexpr int H = 2, W = TILE_W + 2 * H;  // 36
    constexpr int S = W + 1;
    constexpr int EW = W / 2 + 1;  // 19 coarse cells cover the region
    constexpr int ES = EW + 2;
    __shared__ unsigned long long mrow[W];
    __shared__ double X[W * S];
    __shared__ double Bv[W * S];
    __shared__ double X3[W * S];
    __shared__ double E[EW * ES];
    __shared__ double s_red[FTHREADS / 32];
    if (scal[blockIdx.y].done)
        return;
    const int x = threadIdx.x, y = threadIdx.y, t = y * FX + x;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    load_region_mask<H>(lf, ty, tx, mrow, t);
    // the coarse correction under the region: coarse rows r0/2 - 1 .. r0/2 + 17 (zero outside the coarse grid)
    {
        const double* e = ec + (int64_t)blockIdx.y * lc.plane;
        const int64_t I0 = (r0 >> 1) - 1, J0 = (c0 >> 1) - 1;
        for (int i = t; i < EW * EW; i += FTHREADS) {
            int ei = i / EW, ej = i - ei * EW;
            int64_t I = I0 + ei, J = J0 + ej;
            E[ei * ES + ej] = (I >= 0 && I < lc.rows && J >= 0 && J < lc.cols) ? e[I * lc.pitch + J] : 0.0;
        }
    }
    __syncthreads();
    const int64_t gc = c0 - H + x;
    const int64_t boff = (int64_t)blockIdx.y * lf.plane + (r0 - H) * lf.pitch + gc;
    {
        // x + P e on the 36 x 36 region, b on the inner 34 x 34; loads first (predicated, back to back), then use
        constexpr int NK = W / FY;
        const double* xp = x_in + boff + (int64_t)y * lf.pitch;
        const double* bp = b + boff + (int64_t)y * lf.pitch;
        const int64_t step = (int64_t)FY * lf.pitch;
        double xv[NK], bv[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            int row = y + k * FY;
            bool on = x < W && ((mrow[row] >> x) & 1);
            bool inner = on && row >= 1 && row < W - 1 && x >= 1 && x < W - 1;
            xv[k] = on ? xp[k * step] : 0.0;
            bv[k] = inner ? bp[k * step] : 0.0;
        }
        if (x < W) {
            const int ej = x >> 1, oj = x & 1;  // c0 - 2 is even: parity of the local index = global parity
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                int row = y + k * FY;
                double v = 0.0;
                if ((mrow[row] >> x) & 1) {
                    const double* p = E + (row >> 1) * ES + ej;
                    int oi = (row & 1) * ES;
                    v = xv[k] + 0.25 * ((p[0] + p[oj]) + (p[oi] + p[oi + oj]));  // bilinear, branch free
                }
                X[row * S + x] = v;
                Bv[row * S + x] = bv[k];
            }
        }
    }
    __syncthreads();
    if (x >= 1 && x < W - 1) {  // post-smoothing sweep 1 on the 34 x 34 region
#pragma unroll
        for (int k = 0; k < W / FY; ++k) {
            int row = y + k * FY;
            if (row >= 1 && row < W - 1) {
                double x3 = 0.0;
                if ((mrow[row] >> x) & 1) {
                    double d, inv;
                    diag_pair<FIXED>(lf, r0 - H + row, gc, d, inv);
                    const double* p = X + row * S + x;
                    double xc = p[0];
                    double ax = d * xc - ((p[-S] + p[S]) + (p[-1] + p[1]));
                    x3 = xc + FW * inv * (Bv[row * S + x] - ax);
                }
                X3[row * S + x] = x3;
            }
        }
    }
    __syncthreads();
    double* xo = x_out + boff;
    double acc = 0.0;
    if (x >= H && x < H + TILE_W) {  // sweep 2 on the tile itself
#pragma unroll
        for (int k = 0; k < TILE_H / FY; ++k) {
            int row = H + y + k * FY;
            if ((mrow[row] >> x) & 1) {
                double d, inv;
                diag_pair<FIXED>(lf, r0 - H + row, gc, d, inv);
                const double* p = X3 + row * S + x;
                double xc = p[0];
                double ax = d * xc - ((p[-S] + p[S]) + (p[-1] + p[1]));
                double bv = Bv[row * S + x];
                double x4 = xc + FW * inv * (bv - ax);
                xo[row * lf.pitch] = x4;
                if (DOT)
                    acc += bv * x4;
            }
        }
    }
    if (DOT) {
        for (int o = 16; o; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((t & 31) == 0)
            s_red[t >> 5] = acc;
        __syncthreads();
        if (t == 0) {
            double s = 0.0;
            for (int w = 0; w < FTHREADS / 32; ++w)
                s += s_red[w];
            if (s != 0.0)
                atomicAdd(&scal[blockIdx.y].rz[slot], s);
        }
    }
}

int launch_mg_down(sa_ctx* ctx, const Level& lf, const Level& lc, int nbands, const double* b, double* x_out, double* bc,
    const BandScalars* scal)
{
    dim3 grid((unsigned)lf.n_tiles, (unsigned)nbands), block(FX, FY);
    if (lf.fixed_diag)
        SA_LAUNCH(ctx, k_mg_down<true>, grid, block, 0, lf, lc, b, x_out, bc, scal);
    else
        SA_LAUNCH(ctx, k_mg_down<false>, grid, block, 0, lf, lc, b, x_out, bc, scal);
    return SA_OK;
}

int launch_mg_up(sa_ctx* ctx, const Level& lf, const Level& lc, int nbands, const double* x_in, const double* b,
    const double* ec, double* x_out, BandScalars* scal, int rz_slot)
{
    dim3 grid((unsigned)lf.n_tiles, (unsigned)nbands), block(FX, FY);
    const int slot = rz_slot >= 0 ? rz_slot : 0;
    if (lf.fixed_diag) {
        if (rz_slot >= 0)
            SA_LAUNCH(ctx, (k_mg_up<true, true>), grid, block, 0, lf, lc, x_in, b, ec, x_out, scal, slot);
        else
            SA_LAUNCH(ctx, (k_mg_up<true, false>), grid, block, 0, lf, lc, x_in, b, ec, x_out, scal, slot);
    } else {
        if (rz_slot >= 0)
            SA_LAUNCH(ctx, (k_mg_up<false, true>), grid, block, 0, lf, lc, x_in, b, ec, x_out, scal, slot);
        else
            SA_LAUNCH(ctx, (k_mg_up<false, false>), grid, block, 0, lf, lc, x_in, b, ec, x_out, scal, slot);
    }
    return SA_OK;
}

}  // namespace satfill
