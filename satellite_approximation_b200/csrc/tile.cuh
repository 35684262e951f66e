// Device helpers shared by the CG (cg.cu) and multigrid (mg.cu) kernels: diagonal of the 5-point operator, CTA-wide
// sums by warp shuffle, and staging of a 32 x 32 tile with its one-cell halo in shared memory.
#pragma once
#include "common.cuh"

namespace satfill {

__device__ __forceinline__ double inv_diag(int64_t r, int64_t c, int64_t rows, int64_t cols)
{
    // in-image neighbour count: poisson.cpp:187-190 (valid_neighbours, utils.h:35-50); 4 for every Laplace unknown.
    int d = (r > 0) + (r < rows - 1) + (c > 0) + (c < cols - 1);
    // Eigen's DiagonalPreconditioner uses 1 for a zero diagonal (BasicPreconditioners.h:66-70)
    return d == 4 ? 0.25 : (d == 3 ? (1.0 / 3.0) : (d == 2 ? 0.5 : 1.0));
}
__device__ __forceinline__ double diag_of(int64_t r, int64_t c, int64_t rows, int64_t cols)
{
    return (double)((r > 0) + (r < rows - 1) + (c > 0) + (c < cols - 1));
}

__device__ __forceinline__ double block_sum(double v, double* s_red /* 8 doubles */)
{
    for (int o = 16; o; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // protect s_red reuse
    if (threadIdx.x == 0)
        s_red[threadIdx.y] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.y == 0) {
        t = threadIdx.x < CG_BLOCK_Y ? s_red[threadIdx.x] : 0.0;
        for (int o = 4; o; o >>= 1)
            t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread (0, 0)
}

// Stage the 34 x 34 neighbourhood of a tile in shared memory.  `f(idx, r, c)` yields the value of the staged vector
// at plane offset idx; it is evaluated for the 32 x 32 interior (coalesced 256 B rows) and the 4 x 32 halo cells.
constexpr int SP = TILE_W + 3;  // padded row length of the staged tile (odd multiple keeps 8-byte banks spread)

template <typename F>
__device__ __forceinline__ void stage_tile(double (*sp)[SP], int64_t r0, int64_t c0, int64_t pitch, F f)
{
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y;
        int64_t r = r0 + lr, c = c0 + threadIdx.x;
        sp[lr + 1][threadIdx.x + 1] = f(r * pitch + c, r, c, true);
    }
    int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    if (t < 128) {
        int e = t >> 5, i = t & 31;
        int lr, lc;
        if (e == 0) { lr = -1; lc = i; }
        else if (e == 1) { lr = TILE_H; lc = i; }
        else if (e == 2) { lr = i; lc = -1; }
        else { lr = i; lc = TILE_W; }
        int64_t r = r0 + lr, c = c0 + lc;
        sp[lr + 1][lc + 1] = f(r * pitch + c, r, c, false);
    }
}


}  // namespace satfill
