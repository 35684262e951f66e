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

// Stage the 34 x 34 neighbourhood of a tile in shared memory.  `f(idx, r, c, interior)` yields the value of the staged
// vector at plane offset idx; it is evaluated for the 32 x 32 interior (coalesced 256 B rows) and the 4 x 32 halo
// cells, but ONLY at cells of the unknown set: every staged vector is zero elsewhere by construction, so the loads
// of known cells are predicated off and their 32-byte sectors never leave HBM (tiles are ~50 % known cells on
// cloud-like masks).  The mask bytes are read first so that the predicated loads can all be in flight together.
constexpr int SP = TILE_W + 3;  // padded row length of the staged tile (odd multiple keeps 8-byte banks spread)

template <typename F>
__device__ __forceinline__ void stage_tile(double (*sp)[SP], const uint8_t* __restrict__ umask, int64_t r0, int64_t c0,
    int64_t pitch, F f)
{
    uint8_t m[ROWS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j)
        m[j] = umask[(r0 + threadIdx.y + j * CG_BLOCK_Y) * pitch + c0 + threadIdx.x];
    int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    int hr = 0, hc = 0;
    uint8_t hm = 0;
    if (t < 128) {
        int e = t >> 5, i = t & 31;
        if (e == 0) { hr = -1; hc = i; }
        else if (e == 1) { hr = TILE_H; hc = i; }
        else if (e == 2) { hr = i; hc = -1; }
        else { hr = i; hc = TILE_W; }
        hm = umask[(r0 + hr) * pitch + c0 + hc];
    }
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y;
        int64_t r = r0 + lr, c = c0 + threadIdx.x;
        sp[lr + 1][threadIdx.x + 1] = m[j] ? f(r * pitch + c, r, c, true) : 0.0;
    }
    if (t < 128) {
        int64_t r = r0 + hr, c = c0 + hc;
        sp[hr + 1][hc + 1] = hm ? f(r * pitch + c, r, c, false) : 0.0;
    }
}

}  // namespace satfill
