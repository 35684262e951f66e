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

// Stage the 34 x 34 neighbourhood of a tile in shared memory.  For each staged cell, `load(idx, v)` reads NV values
// (plane offset idx) and `combine(v, idx, r, c, interior)` turns them into the staged value.  The 32 x 32 interior is
// read as coalesced 256 B rows, plus 4 x 32 halo cells -- but ONLY at cells of the unknown set: every staged vector is
// zero elsewhere by construction, so the loads of known cells are predicated off and their 32-byte sectors never
// leave HBM (tiles are ~50 % known cells on cloud-like masks).  The schedule is: all mask bytes, then ALL loads of the
// thread back to back (predicated, branch free), then the combines -- a load that sits in the same branch as its use
// serialises the thread on one HBM round trip per cell.
constexpr int SP = TILE_W + 3;  // padded row length of the staged tile (odd multiple keeps 8-byte banks spread)

template <int NV, typename FL, typename FC>
__device__ __forceinline__ void stage_tile(double (*sp)[SP], const uint8_t* __restrict__ umask, int64_t r0, int64_t c0,
    int64_t pitch, FL load, FC combine)
{
    // 4 x 32 halo cells: edge e = 0 top, 1 bottom, 2 left, 3 right; HALO_PER_THREAD of them per thread
    constexpr int HALO_PER_THREAD = (128 + CG_THREADS - 1) / CG_THREADS;
    const int t = threadIdx.y * CG_BLOCK_X + threadIdx.x;
    uint8_t m[ROWS_PER_THREAD], hm[HALO_PER_THREAD];
    int hr[HALO_PER_THREAD], hc[HALO_PER_THREAD];
    const int64_t base = (r0 + threadIdx.y) * pitch + c0 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j)
        m[j] = umask[base + (int64_t)j * CG_BLOCK_Y * pitch];
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k) {
        int h = t + k * CG_THREADS;
        int e = h >> 5, i = h & 31;
        hr[k] = e == 0 ? -1 : (e == 1 ? TILE_H : i);
        hc[k] = e == 2 ? -1 : (e == 3 ? TILE_W : i);
        hm[k] = h < 128 ? umask[(r0 + hr[k]) * pitch + c0 + hc[k]] : 0;
    }
    double v[ROWS_PER_THREAD][NV], hv[HALO_PER_THREAD][NV];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
#pragma unroll
        for (int q = 0; q < NV; ++q)
            v[j][q] = 0.0;
        if (m[j])
            load(base + (int64_t)j * CG_BLOCK_Y * pitch, v[j]);
    }
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k) {
#pragma unroll
        for (int q = 0; q < NV; ++q)
            hv[k][q] = 0.0;
        if (hm[k])
            load((r0 + hr[k]) * pitch + c0 + hc[k], hv[k]);
    }
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int lr = threadIdx.y + j * CG_BLOCK_Y;
        sp[lr + 1][threadIdx.x + 1] =
            m[j] ? combine(v[j], base + (int64_t)j * CG_BLOCK_Y * pitch, r0 + lr, c0 + threadIdx.x, true) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < HALO_PER_THREAD; ++k) {
        if (t + k * CG_THREADS < 128)
            sp[hr[k] + 1][hc[k] + 1] =
                hm[k] ? combine(hv[k], (r0 + hr[k]) * pitch + c0 + hc[k], r0 + hr[k], c0 + hc[k], false) : 0.0;
    }
}

// Unknown set of column x of the (32 + 2H)^2 neighbourhood of tile (ty, tx): bit (row + H) <=> cell
// (r0 + row, c0 - H + x) is an unknown, row in [-H, 32 + H).
template <int H>
__device__ __forceinline__ unsigned long long region_col_mask(const Level& lv, int ty, int tx, int x)
{
    int gc = x - H;
    int txx = tx + (gc < 0 ? -1 : (gc >= TILE_W ? 1 : 0));
    const uint32_t* w = lv.tbitsT + ((size_t)(ty + 1) * lv.tb_stride + (txx + 1)) * 32 + (gc & 31);
    const size_t vs = (size_t)lv.tb_stride * 32;
    unsigned long long C = w[0], N = *(w - vs), Sx = w[vs];
    return (N >> (32 - H)) | (C << H) | ((Sx & ((1ull << H) - 1)) << (32 + H));
}

}  // namespace satfill
