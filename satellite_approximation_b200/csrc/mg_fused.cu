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
#include "common.cuh"
#include "tile.cuh"

namespace satfill {

constexpr double FW = 0.8;  // damped-Jacobi weight, same as mg.cu

// Row masks of the (32 + 2H)^2 neighbourhood of tile (ty, tx): bit (col + H) of mrow[row + H] <=> cell
// (r0 + row, c0 + col) is an unknown, for row, col in [-H, 32 + H).
template <int H>
__device__ __forceinline__ void load_region_mask(const Level& lv, int ty, int tx, unsigned long long* mrow)
{
    int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    for (int row = t; row < TILE_H + 2 * H; row += CG_THREADS) {
        int gr = row - H;
        int tyy = ty + (gr < 0 ? -1 : (gr >= TILE_H ? 1 : 0));
        const uint32_t* w = lv.tbits + ((size_t)(tyy + 1) * lv.tb_stride + (tx + 1)) * 32 + (gr & 31);
        unsigned long long C = w[0], L = w[-32], R = w[32];
        mrow[row] = (L >> (32 - H)) | (C << H) | ((R & ((1ull << H) - 1)) << (32 + H));
    }
}

__device__ __forceinline__ double f_diag(const Level& lv, int64_t r, int64_t c)
{
    return lv.fixed_diag ? 4.0 : fmax(diag_of(r, c, lv.rows, lv.cols), 1.0);
}

__global__ void __launch_bounds__(CG_THREADS) k_mg_down(Level lf, Level lc, const double* __restrict__ b,
    double* __restrict__ x_out, double* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    constexpr int H = 3, W = TILE_W + 2 * H;  // 38
    __shared__ unsigned long long mrow[W];
    __shared__ double B[W][W + 1];
    __shared__ double X1[W][W + 1];  // sweep 1; later reused for the residual
    __shared__ double X2[W - 2][W - 1];
    if (scal[blockIdx.y].done)
        return;
    const int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    load_region_mask<H>(lf, ty, tx, mrow);
    __syncthreads();
    const double* bb = b + (int64_t)blockIdx.y * lf.plane + (r0 - H) * lf.pitch + (c0 - H);
    for (int i = t; i < W * W; i += CG_THREADS) {
        int row = i / W, col = i - row * W;
        double v = 0.0, x1 = 0.0;
        if ((mrow[row] >> col) & 1) {
            v = bb[row * lf.pitch + col];
            x1 = FW * v / f_diag(lf, r0 - H + row, c0 - H + col);
        }
        B[row][col] = v;
        X1[row][col] = x1;
    }
    __syncthreads();
    for (int i = t; i < (W - 2) * (W - 2); i += CG_THREADS) {  // sweep 2 on the 36 x 36 region
        int rr = i / (W - 2), cc = i - rr * (W - 2);
        int row = rr + 1, col = cc + 1;
        double x2 = 0.0;
        if ((mrow[row] >> col) & 1) {
            double d = f_diag(lf, r0 - H + row, c0 - H + col);
            double xc = X1[row][col];
            double ax = d * xc - (X1[row - 1][col] + X1[row + 1][col] + X1[row][col - 1] + X1[row][col + 1]);
            x2 = xc + FW * (B[row][col] - ax) / d;
        }
        X2[rr][cc] = x2;
    }
    __syncthreads();
    double(*R)[W + 1] = X1;  // residual on the 34 x 34 region (X1 is dead)
    double* xo = x_out + (int64_t)blockIdx.y * lf.plane + (r0 - 1) * lf.pitch + (c0 - 1);
    for (int i = t; i < (W - 4) * (W - 4); i += CG_THREADS) {
        int rr = i / (W - 4), cc = i - rr * (W - 4);
        int row = rr + 2, col = cc + 2;  // position in the 38-grid; X2 index = (row - 1, col - 1)
        double res = 0.0;
        if ((mrow[row] >> col) & 1) {
            double d = f_diag(lf, r0 - H + row, c0 - H + col);
            double xc = X2[row - 1][col - 1];
            double ax = d * xc - (X2[row - 2][col - 1] + X2[row][col - 1] + X2[row - 1][col - 2] + X2[row - 1][col]);
            res = B[row][col] - ax;
            if (rr >= 1 && rr <= TILE_H && cc >= 1 && cc <= TILE_W)
                xo[rr * lf.pitch + cc] = xc;  // the CTA's own tile
        }
        R[rr][cc] = res;
    }
    __syncthreads();
    // restriction: coarse cell (ci, cj) of this tile sits on fine tile cell (2 ci, 2 cj) = R[2 ci + 1][2 cj + 1]
    double* bco = bc + (int64_t)blockIdx.y * lc.plane + (r0 >> 1) * lc.pitch + (c0 >> 1);
    for (int i = t; i < (TILE_H / 2) * (TILE_W / 2); i += CG_THREADS) {
        int ci = i / (TILE_W / 2), cj = i - ci * (TILE_W / 2);
        if ((mrow[2 * ci + H] >> (2 * cj + H)) & 1) {  // mask injection: coarse unknown <=> fine (2I, 2J) unknown
            int a = 2 * ci + 1, c = 2 * cj + 1;
            double up = 0.5 * R[a - 1][c - 1] + R[a - 1][c] + 0.5 * R[a - 1][c + 1];
            double mid = 0.5 * R[a][c - 1] + R[a][c] + 0.5 * R[a][c + 1];
            double dn = 0.5 * R[a + 1][c - 1] + R[a + 1][c] + 0.5 * R[a + 1][c + 1];
            bco[ci * lc.pitch + cj] = 0.5 * up + mid + 0.5 * dn;
        }
    }
}

template <bool DOT>
__global__ void __launch_bounds__(CG_THREADS) k_mg_up(Level lf, Level lc, const double* __restrict__ x_in,
    const double* __restrict__ b, const double* __restrict__ ec, double* __restrict__ x_out,
    BandScalars* __restrict__ scal, int slot)
{
    constexpr int H = 2, W = TILE_W + 2 * H;  // 36
    constexpr int EW = W / 2 + 1;             // 19 coarse cells cover the region
    __shared__ unsigned long long mrow[W];
    __shared__ double X[W][W + 1];
    __shared__ double Bv[W - 2][W - 1];
    __shared__ double X3[W - 2][W - 1];
    __shared__ double E[EW][EW + 2];
    __shared__ double s_red[CG_BLOCK_Y];
    if (scal[blockIdx.y].done)
        return;
    const int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    load_region_mask<H>(lf, ty, tx, mrow);
    // the coarse correction under the region: coarse rows r0/2 - 1 .. r0/2 + 17 (zero outside the coarse grid)
    {
        const double* e = ec + (int64_t)blockIdx.y * lc.plane;
        const int64_t I0 = (r0 >> 1) - 1, J0 = (c0 >> 1) - 1;
        for (int i = t; i < EW * EW; i += CG_THREADS) {
            int ei = i / EW, ej = i - ei * EW;
            int64_t I = I0 + ei, J = J0 + ej;
            E[ei][ej] = (I >= 0 && I < lc.rows && J >= 0 && J < lc.cols) ? e[I * lc.pitch + J] : 0.0;
        }
    }
    __syncthreads();
    const int64_t boff = (int64_t)blockIdx.y * lf.plane;
    const double* xi = x_in + boff + (r0 - H) * lf.pitch + (c0 - H);
    for (int i = t; i < W * W; i += CG_THREADS) {  // x + P e on the 36 x 36 region
        int row = i / W, col = i - row * W;
        double v = 0.0;
        if ((mrow[row] >> col) & 1) {
            int ei = row >> 1, ej = col >> 1;  // r0 - 2 and c0 - 2 are even: parity of the local index = global parity
            double pe;
            if ((row & 1) == 0)
                pe = (col & 1) == 0 ? E[ei][ej] : 0.5 * (E[ei][ej] + E[ei][ej + 1]);
            else
                pe = (col & 1) == 0 ? 0.5 * (E[ei][ej] + E[ei + 1][ej])
                                    : 0.25 * (E[ei][ej] + E[ei][ej + 1] + E[ei + 1][ej] + E[ei + 1][ej + 1]);
            v = xi[row * lf.pitch + col] + pe;
        }
        X[row][col] = v;
    }
    const double* bb = b + boff + (r0 - 1) * lf.pitch + (c0 - 1);
    for (int i = t; i < (W - 2) * (W - 2); i += CG_THREADS) {
        int rr = i / (W - 2), cc = i - rr * (W - 2);
        Bv[rr][cc] = ((mrow[rr + 1] >> (cc + 1)) & 1) ? bb[rr * lf.pitch + cc] : 0.0;
    }
    __syncthreads();
    for (int i = t; i < (W - 2) * (W - 2); i += CG_THREADS) {  // post-smoothing sweep 1 on the 34 x 34 region
        int rr = i / (W - 2), cc = i - rr * (W - 2);
        int row = rr + 1, col = cc + 1;
        double x3 = 0.0;
        if ((mrow[row] >> col) & 1) {
            double d = f_diag(lf, r0 - H + row, c0 - H + col);
            double xc = X[row][col];
            double ax = d * xc - (X[row - 1][col] + X[row + 1][col] + X[row][col - 1] + X[row][col + 1]);
            x3 = xc + FW * (Bv[rr][cc] - ax) / d;
        }
        X3[rr][cc] = x3;
    }
    __syncthreads();
    double* xo = x_out + boff + r0 * lf.pitch + c0;
    double acc = 0.0;
    for (int i = t; i < TILE_H * TILE_W; i += CG_THREADS) {  // sweep 2 on the tile itself
        int ri = i / TILE_W, ci = i - ri * TILE_W;
        if ((mrow[ri + H] >> (ci + H)) & 1) {
            int a = ri + 1, c = ci + 1;  // index in X3 / Bv
            double d = f_diag(lf, r0 + ri, c0 + ci);
            double xc = X3[a][c];
            double ax = d * xc - (X3[a - 1][c] + X3[a + 1][c] + X3[a][c - 1] + X3[a][c + 1]);
            double bv = Bv[a][c];
            double x4 = xc + FW * (bv - ax) / d;
            xo[ri * lf.pitch + ci] = x4;
            if (DOT)
                acc += bv * x4;
        }
    }
    if (DOT) {
        double s = block_sum(acc, s_red);
        if (t == 0 && s != 0.0)
            atomicAdd(&scal[blockIdx.y].rz[slot], s);
    }
}

int launch_mg_down(sa_ctx* ctx, const Level& lf, const Level& lc, int nbands, const double* b, double* x_out, double* bc,
    const BandScalars* scal)
{
    dim3 grid((unsigned)lf.n_tiles, (unsigned)nbands), block(CG_BLOCK_X, CG_BLOCK_Y);
    SA_LAUNCH(ctx, k_mg_down, grid, block, 0, lf, lc, b, x_out, bc, scal);
    return SA_OK;
}

int launch_mg_up(sa_ctx* ctx, const Level& lf, const Level& lc, int nbands, const double* x_in, const double* b,
    const double* ec, double* x_out, BandScalars* scal, int rz_slot)
{
    dim3 grid((unsigned)lf.n_tiles, (unsigned)nbands), block(CG_BLOCK_X, CG_BLOCK_Y);
    if (rz_slot >= 0)
        SA_LAUNCH(ctx, k_mg_up<true>, grid, block, 0, lf, lc, x_in, b, ec, x_out, scal, rz_slot);
    else
        SA_LAUNCH(ctx, k_mg_up<false>, grid, block, 0, lf, lc, x_in, b, ec, x_out, scal, 0);
    return SA_OK;
}

}  // namespace satfill
