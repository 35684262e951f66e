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
        for (int o = CG_BLOCK_Y / 2; o; o >>= 1)
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
    // 4 x 32 halo cells: edge e = 0 top, 1 bottom, 2 left, 3 right; HALO_PER_THREAD of them per thread
    constexpr int HALO_PER_THREAD = (128 + CG_THREADS - 1) / CG_THREADS;
    int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    int hr[HALO_PER_THREAD], hc[HALO_PER_THREAD];
    uint8_t hm[HALO_PER_THREAD];
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k) {
        int h = t + k * CG_THREADS;
        int e = h >> 5, i = h & 31;
        hr[k] = e == 0 ? -1 : (e == 1 ? TILE_H : i);
        hc[k] = e == 2 ? -1 : (e == 3 ? TILE_W : i);
        hm[k] = h < 128 ? umask[(r0 + hr[k]) * pitch + c0 + hc[k]] : 0;
    }
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y;
        int64_t r = r0 + lr, c = c0 + threadIdx.x;
        sp[lr + 1][threadIdx.x + 1] = m[j] ? f(r * pitch + c, r, c, true) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k) {
        if (t + k * CG_THREADS < 128) {
            int64_t r = r0 + hr[k], c = c0 + hc[k];
            sp[hr[k] + 1][hc[k] + 1] = hm[k] ? f(r * pitch + c, r, c, false) : 0.0;
        }
    }
}

}  // namespace satfill
