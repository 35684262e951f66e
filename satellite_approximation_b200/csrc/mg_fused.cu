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
// Thread mapping: 40 x 4 threads; thread (x, y) owns column x of the staged region and a CONTIGUOUS run of rows
// [RG y, RG y + RG).  Consequences, each of which removed instructions from what was an issue-bound kernel
// (profiles/r1_fused_v1_*: 62 % issue utilisation at 18 % DRAM utilisation):
//   * shared-memory addresses are a per-thread base plus compile-time offsets (no div / mod per cell);
//   * the unknown set of the thread's column is ONE 64-bit register (Level::tbitsT, three L2-resident words), tested
//     with a compile-time bit index -- no mask array in shared memory, no barrier before the global loads;
//   * the north / centre / south values of the 5-point stencil slide through registers down the column: 4 shared
//     loads per cell instead of 6;
//   * all global loads of a thread are issued back to back before the first use, and cells are processed branch
//     free (results of known cells are discarded by a select), so the unrolled body schedules as one block.
#include "common.cuh"
#include "tile.cuh"

#if SATFILL_LEGACY_VARIANTS
namespace satfill {

constexpr double FW = 0.8;  // damped-Jacobi weight, same as mg.cu
constexpr int FX = 40, FY = 4, FTHREADS = FX * FY;

// diagonal and omega / diagonal; FIXED: Laplace (every unknown has four in-image neighbours)
template <bool FIXED>
__device__ __forceinline__ void diag_pair(const Level& lv, int64_t r, int64_t c, double& d, double& winv)
{
    if (FIXED) {
        d = 4.0;
        winv = FW * 0.25;
    } else {
        int n = (r > 0) + (r < lv.rows - 1) + (c > 0) + (c < lv.cols - 1);
        d = n < 1 ? 1.0 : (double)n;
        winv = FW * (n == 4 ? 0.25 : (n == 3 ? (1.0 / 3.0) : (n == 2 ? 0.5 : 1.0)));
    }
}

template <bool FIXED>
__global__ void __launch_bounds__(FTHREADS, 8) k_mg_down(Level lf, Level lc, const double* __restrict__ b,
    double* __restrict__ x_out, double* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    constexpr int H = 3, W = TILE_W + 2 * H;  // 38
    constexpr int S = W + 1;                  // shared row stride
    constexpr int RG = (W + FY - 1) / FY;     // 10 rows per thread
    constexpr int RA = RG * FY + 1;           // allocated rows: the sliding window reads one row past the last
    __shared__ double X1[RA * S];  // sweep 1; later reused for the residual
    __shared__ double X2[RA * S];  // sweep 2
    if (scal[blockIdx.y].done)
        return;
    const int x = threadIdx.x, y = threadIdx.y, t = y * FX + x;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    const int row0 = RG * y;
    const unsigned long long cm = x < W ? region_col_mask<H>(lf, ty, tx, x) : 0ull;
    const unsigned my = (unsigned)(cm >> row0);  // bit k <=> (row0 + k, x) is an unknown; rows >= W have no bits
    const int64_t gr = r0 - H + row0, gc = c0 - H + x;
    const int sbase = row0 * S + x;
    double v[RG];  // the right-hand side of the thread's own cells stays in registers (it is only used pointwise)
    {
        const double* bp = b + (int64_t)blockIdx.y * lf.plane + gr * lf.pitch + gc;
#pragma unroll
        for (int k = 0; k < RG; ++k)
            v[k] = ((my >> k) & 1) ? bp[k * lf.pitch] : 0.0;
        if (x < W) {
#pragma unroll
            for (int k = 0; k < RG; ++k) {
                double d, winv;
                diag_pair<FIXED>(lf, gr + k, gc, d, winv);
                X1[sbase + k * S] = winv * v[k];  // one damped-Jacobi sweep from zero
            }
        }
    }
    __syncthreads();
    if (x >= 1 && x < W - 1) {  // sweep 2 on rows / columns 1 .. 36
        const double* p = X1 + sbase;
        double n = y > 0 ? p[-S] : 0.0, c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            double sv = p[(k + 1) * S];
            double d, winv;
            diag_pair<FIXED>(lf, gr + k, gc, d, winv);
            double ax = d * c - ((n + sv) + (p[k * S - 1] + p[k * S + 1]));
            double x2 = c + winv * (v[k] - ax);
            bool on = ((my >> k) & 1) && row0 + k >= 1 && row0 + k < W - 1;
            X2[sbase + k * S] = on ? x2 : 0.0;
            n = c;
            c = sv;
        }
    }
    __syncthreads();
    double* R = X1;  // residual on rows / columns 2 .. 35 (X1 is dead)
    if (x >= 2 && x < W - 2) {
        const double* p = X2 + sbase;
        double* xo = x_out + (int64_t)blockIdx.y * lf.plane + gr * lf.pitch + gc;
        const bool own_col = x >= H && x < H + TILE_W;
        double n = y > 0 ? p[-S] : 0.0, c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            double sv = p[(k + 1) * S];
            double d, winv;
            diag_pair<FIXED>(lf, gr + k, gc, d, winv);
            double ax = d * c - ((n + sv) + (p[k * S - 1] + p[k * S + 1]));
            double res = v[k] - ax;
            int row = row0 + k;
            bool on = ((my >> k) & 1) && row >= 2 && row < W - 2;
            R[sbase + k * S] = on ? res : 0.0;
            if (on && own_col && row >= H && row < H + TILE_H)
                xo[k * lf.pitch] = c;  // the CTA's own tile
            n = c;
            c = sv;
        }
    }
    __syncthreads();
    // restriction: coarse cell (ci, cj) of this tile sits on fine tile cell (2 ci, 2 cj) = region (2 ci + 3, 2 cj + 3)
    double* bco = bc + (int64_t)blockIdx.y * lc.plane + (r0 >> 1) * lc.pitch + (c0 >> 1);
    const uint32_t* rowbits = lf.tbits + ((size_t)(ty + 1) * lf.tb_stride + (tx + 1)) * 32;
    for (int i = t; i < (TILE_H / 2) * (TILE_W / 2); i += FTHREADS) {
        int ci = i >> 4, cj = i & 15;
        if ((rowbits[2 * ci] >> (2 * cj)) & 1) {  // mask injection: coarse unknown <=> fine (2I, 2J) unknown
            const double* p = R + (2 * ci + H) * S + 2 * cj + H;
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
    constexpr int RG = W / FY;        // 9 rows per thread
    constexpr int RA = RG * FY + 1;
    constexpr int EW = W / 2 + 1;     // 19 coarse cells cover the region
    constexpr int ES = EW + 2;
    __shared__ double X[RA * S];
    __shared__ double X3[RA * S];
    __shared__ double E[EW * ES];
    __shared__ double s_red[FTHREADS / 32];
    if (scal[blockIdx.y].done)
        return;
    const int x = threadIdx.x, y = threadIdx.y, t = y * FX + x;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    const int row0 = RG * y;
    const unsigned long long cm = x < W ? region_col_mask<H>(lf, ty, tx, x) : 0ull;
    const unsigned my = (unsigned)(cm >> row0);
    const int64_t gr = r0 - H + row0, gc = c0 - H + x;
    const int sbase = row0 * S + x;
    const int64_t goff = (int64_t)blockIdx.y * lf.plane + gr * lf.pitch + gc;
    // global loads first: x on the 36 x 36 region, b on the inner 34 x 34, the coarse correction under the region
    double xv[RG], bv[RG];
    {
        const double* xp = x_in + goff;
        const double* bp = b + goff;
        const bool inner_col = x >= 1 && x < W - 1;
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            bool on = (my >> k) & 1;
            xv[k] = on ? xp[k * lf.pitch] : 0.0;
            bv[k] = (on && inner_col && row0 + k >= 1 && row0 + k < W - 1) ? bp[k * lf.pitch] : 0.0;
        }
        const double* e = ec + (int64_t)blockIdx.y * lc.plane;
        const int64_t I0 = (r0 >> 1) - 1, J0 = (c0 >> 1) - 1;
        for (int i = t; i < EW * EW; i += FTHREADS) {
            int ei = i / EW, ej = i - ei * EW;
            int64_t I = I0 + ei, J = J0 + ej;
            E[ei * ES + ej] = (I >= 0 && I < lc.rows && J >= 0 && J < lc.cols) ? e[I * lc.pitch + J] : 0.0;
        }
    }
    __syncthreads();
    if (x < W) {  // X = x + P e (bilinear; r0 - 2 and c0 - 2 are even, so local parity = global parity)
        const int ej = x >> 1, oj = x & 1;
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            int row = row0 + k;
            const double* p = E + (row >> 1) * ES + ej;
            int oi = (row & 1) * ES;
            double pe = 0.25 * ((p[0] + p[oj]) + (p[oi] + p[oi + oj]));
            X[sbase + k * S] = ((my >> k) & 1) ? xv[k] + pe : 0.0;
        }
    }
    __syncthreads();
    if (x >= 1 && x < W - 1) {  // post-smoothing sweep 1 on rows / columns 1 .. 34
        const double* p = X + sbase;
        double n = y > 0 ? p[-S] : 0.0, c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            double sv = p[(k + 1) * S];
            double d, winv;
            diag_pair<FIXED>(lf, gr + k, gc, d, winv);
            double ax = d * c - ((n + sv) + (p[k * S - 1] + p[k * S + 1]));
            double x3 = c + winv * (bv[k] - ax);
            bool on = ((my >> k) & 1) && row0 + k >= 1 && row0 + k < W - 1;
            X3[sbase + k * S] = on ? x3 : 0.0;
            n = c;
            c = sv;
        }
    }
    __syncthreads();
    double acc = 0.0;
    if (x >= H && x < H + TILE_W) {  // sweep 2 on the tile itself
        const double* p = X3 + sbase;
        double* xo = x_out + goff;
        double n = y > 0 ? p[-S] : 0.0, c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            double sv = p[(k + 1) * S];
            double d, winv;
            diag_pair<FIXED>(lf, gr + k, gc, d, winv);
            double ax = d * c - ((n + sv) + (p[k * S - 1] + p[k * S + 1]));
            double x4 = c + winv * (bv[k] - ax);
            int row = row0 + k;
            if (((my >> k) & 1) && row >= H && row < H + TILE_H) {
                xo[k * lf.pitch] = x4;
                if (DOT)
                    acc += bv[k] * x4;
            }
            n = c;
            c = sv;
        }
    }
    if (DOT) {
        for (int o = 16; o; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((t & 31) == 0)
            s_red[t >> 5] = acc;
        __syncthreads();
        if (t == 0) {
            double sum = 0.0;
            for (int w = 0; w < FTHREADS / 32; ++w)
                sum += s_red[w];
            if (sum != 0.0)
                atomicAdd(&scal[blockIdx.y].rz[slot], sum);
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
#endif  // SATFILL_LEGACY_VARIANTS
