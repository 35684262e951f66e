// The two kernels of one CG iteration (see cg.cu for the algebra), second generation: no shared-memory staging.
//
// One CTA of 128 threads per active 32 x 32 tile and band.  Thread (cx, strip) owns the 16-byte aligned column pair
// (2cx, 2cx + 1) of the four rows [4 strip, 4 strip + 4): a half warp spans one tile row (16 x 16 B = 256 B, fully
// coalesced), a warp two strips.  The 5-point operator needs
//   * north / south neighbours: the thread's own rows, plus one row above and below its strip (re-read by the
//     neighbouring strip's thread: an L1 / L2 hit, never a second HBM transfer);
//   * west / east neighbours: the other cell of the pair, or the adjacent lane's pair through a warp shuffle; the
//     two edge lanes of a row read the tile's halo column instead.
// So a warp never waits for another warp (no __syncthreads before the reduction), every global access is a
// predicated 16-byte (8-byte for float planes) load or store, and all loads of a thread -- 14 to 18 of them -- are
// issued back to back before the first use.  The unknown set of a column is one register (Level::tbitsT): loads of
// known cells are predicated off, so their sectors never leave HBM.
//
// ncu on the first generation (cg.cu: tile + halo staged through shared memory, one 8-byte access per cell, byte
// masks): 2.1 / 3.1 TB/s of algorithmic traffic with three dependent global round trips (tile list -> mask -> data)
// and a barrier per CTA.
#include "common.cuh"
#include "tile.cuh"

namespace satfill {

namespace {

constexpr int ST_THREADS = 128;
constexpr int ST_RG = 4;  // rows per thread

__device__ __forceinline__ double2 ldnc2_if(const double* p, unsigned pred)
{
    double2 v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@q ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(v.x), "=d"(v.y)
        : "l"(p), "r"(pred));
    return v;
}
__device__ __forceinline__ double2 ld2_if(const double* p, unsigned pred)  // data this kernel also writes: no .nc
{
    double2 v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@q ld.global.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(v.x), "=d"(v.y)
        : "l"(p), "r"(pred));
    return v;
}
__device__ __forceinline__ double ldnc_if(const double* p, unsigned pred)
{
    double v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}"
        : "=d"(v)
        : "l"(p), "r"(pred));
    return v;
}
__device__ __forceinline__ double2 ldnc2_if(const float* p, unsigned pred)  // float plane, widened
{
    float x, y;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
        "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
        : "=f"(x), "=f"(y)
        : "l"(p), "r"(pred));
    return make_double2((double)x, (double)y);
}
__device__ __forceinline__ double ldnc_if(const float* p, unsigned pred)
{
    float v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"(pred));
    return (double)v;
}

// Everything a thread needs to know about where it is.
struct Strip {
    int cx, row0;          // column pair index (0..15), first own row (tile-local)
    int64_t gr, gc;        // global row of row0, global column of the pair's left cell
    unsigned mL, mR;       // unknown bits of the two columns: bit j <=> tile row row0 - 1 + j, j = 0..5
    int toff;              // element offset of (row0 - 1, left cell) from the tile's origin, in a plane of pitch `pitch`
    int pitch;
};

__device__ __forceinline__ Strip make_strip(const Level& lv, int tile_index, int& ty, int& tx)
{
    Strip s;
    const int t = threadIdx.x;
    const int yx = lv.tile_yx[tile_index];
    ty = yx >> 16;
    tx = yx & 0xffff;
    s.cx = t & 15;
    s.row0 = (t >> 4) * ST_RG;
    s.pitch = (int)lv.pitch;
    s.gr = (int64_t)ty * TILE_H + s.row0;
    s.gc = (int64_t)tx * TILE_W + 2 * s.cx;
    // frame column of tile column c is c + 1 (halo 1); bit (row + 1) of the mask <=> tile row `row`
    unsigned long long cl = region_col_mask<1>(lv, ty, tx, 2 * s.cx + 1), cr = region_col_mask<1>(lv, ty, tx, 2 * s.cx + 2);
    s.mL = (unsigned)(cl >> s.row0) & 63u;
    s.mR = (unsigned)(cr >> s.row0) & 63u;
    s.toff = (s.row0 - 1) * s.pitch + 2 * s.cx;
    return s;
}

__device__ __forceinline__ double block_sum4(double v, double* s_red /* 4 */)
{
    for (int o = 16; o; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
        s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    return (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);  // valid everywhere
}

template <bool FIXED>
__device__ __forceinline__ void diag_cols(const Level& lv, const Strip& s, int& dL, int& dR)
{
    if (FIXED) {
        dL = dR = 2;
    } else {
        dL = (s.gc > 0) + (s.gc < lv.cols - 1);
        dR = (s.gc + 1 > 0) + (s.gc + 1 < lv.cols - 1);
    }
}
template <bool FIXED>
__device__ __forceinline__ int diag_row(const Level& lv, int64_t r)
{
    return FIXED ? 2 : (r > 0) + (r < lv.rows - 1);
}
__device__ __forceinline__ double inv_of(int d) { return d == 4 ? 0.25 : (d == 3 ? (1.0 / 3.0) : (d == 2 ? 0.5 : 1.0)); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// k_direction2:  beta = rz_k / rz_{k-1};  p' = z + beta p  (halo cells recomputed);  pq = p'.Ap'
//   JACOBI: z = r / d on the fly (zin = r).
// ---------------------------------------------------------------------------------------------------------------
template <bool JACOBI, bool FIXED, typename ZT>
__global__ void __launch_bounds__(ST_THREADS) k_direction2(Level lv, const ZT* __restrict__ zin,
    const double* __restrict__ p_old, double* __restrict__ p_new, BandScalars* __restrict__ scal, int k)
{
    __shared__ double s_red[4];
    BandScalars& sc = scal[blockIdx.y];
    if (sc.done)
        return;
    const int slot = k & 3;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (k > 0 && sc.rr[slot] < sc.thr) {  // ConjugateGradient.h:72-73 (strict <), tested one launch later
        if (lead) {
            sc.rr_exit = sc.rr[slot];
            sc.iters = k - 1;
            __threadfence();
            sc.done = 1;
        }
        return;
    }
    const double beta = k > 0 ? sc.rz[slot] / sc.rz[(k - 1) & 3] : 0.0;  // ConjugateGradient.h:77-79
    if (lead) {  // recycle the slot two iterations ahead
        int z2 = (k + 2) & 3;
        sc.rz[z2] = 0.0;
        sc.rr[z2] = 0.0;
        sc.pq[z2] = 0.0;
    }
    int ty, tx;
    const Strip s = make_strip(lv, blockIdx.x, ty, tx);
    const int64_t origin = (int64_t)blockIdx.y * lv.plane + (int64_t)ty * TILE_H * lv.pitch + (int64_t)tx * TILE_W;
    const ZT* zb = zin + origin;
    const double* pb = p_old + origin;
    const unsigned any = s.mL | s.mR;
    const bool west = s.cx == 0, east = s.cx == 15;
    // ---- all loads: pairs of rows row0-1 .. row0+4, and (edge lanes) the halo column of the own rows
    double2 zv[6], pv[6];
    double ze[ST_RG], pe[ST_RG];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        zv[j] = ldnc2_if(zb + (s.toff + j * s.pitch), (any >> j) & 1);
        pv[j] = ldnc2_if(pb + (s.toff + j * s.pitch), (any >> j) & 1);
    }
    {
        // the halo cell can only matter if the own edge cell is an unknown
        const int eoff = s.toff + (west ? -1 : 2);
        const unsigned em = (west ? s.mL : (east ? s.mR : 0u)) >> 1;
#pragma unroll
        for (int j = 0; j < ST_RG; ++j) {
            ze[j] = ldnc_if(zb + (eoff + (j + 1) * s.pitch), (em >> j) & 1);
            pe[j] = ldnc_if(pb + (eoff + (j + 1) * s.pitch), (em >> j) & 1);
        }
    }
    int dcL, dcR;
    diag_cols<FIXED>(lv, s, dcL, dcR);
    // ---- p' = z + beta p on the 6 x 2 cells and the edge column
    double2 pn[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double zl = zv[j].x, zr = zv[j].y;
        if (JACOBI) {
            int dr = diag_row<FIXED>(lv, s.gr - 1 + j);
            zl *= inv_of(dr + dcL);
            zr *= inv_of(dr + dcR);
        }
        pn[j].x = zl + beta * pv[j].x;  // ConjugateGradient.h:80
        pn[j].y = zr + beta * pv[j].y;
    }
    double pedge[ST_RG];
#pragma unroll
    for (int j = 0; j < ST_RG; ++j) {
        double z = ze[j];
        if (JACOBI) {
            int64_t c = west ? s.gc - 1 : s.gc + 2;
            z *= inv_of(diag_row<FIXED>(lv, s.gr + j) + (FIXED ? 2 : (c > 0) + (c < lv.cols - 1)));
        }
        pedge[j] = z + beta * pe[j];
    }
    // ---- store the own rows, accumulate p'.Ap'
    double* po = p_new + origin;
    double acc = 0.0;
#pragma unroll
    for (int j = 1; j <= ST_RG; ++j) {
        if ((any >> j) & 1)
            *reinterpret_cast<double2*>(po + (s.toff + j * s.pitch)) = pn[j];
        double wl = __shfl_up_sync(0xffffffffu, pn[j].y, 1);    // lane - 1's right cell
        double er = __shfl_down_sync(0xffffffffu, pn[j].x, 1);  // lane + 1's left cell
        if (west)
            wl = pedge[j - 1];
        if (east)
            er = pedge[j - 1];
        int dr = diag_row<FIXED>(lv, s.gr - 1 + j);
        double ql = (double)(dr + dcL) * pn[j].x - ((pn[j - 1].x + pn[j + 1].x) + (wl + pn[j].y));
        double qr = (double)(dr + dcR) * pn[j].y - ((pn[j - 1].y + pn[j + 1].y) + (pn[j].x + er));
        acc += pn[j].x * ql + pn[j].y * qr;  // p' is zero outside the unknown set: no mask needed
    }
    double tot = block_sum4(acc, s_red);
    if (threadIdx.x == 0 && tot != 0.0)
        atomicAdd(&sc.pq[slot], tot);
}

// ---------------------------------------------------------------------------------------------------------------
// k_update2:  alpha = rz / pq;  x += alpha p;  r -= alpha A p  (A p recomputed);  |r|^2 and (JACOBI) r.(r/d)
//   RF: also write the residual as float for the red-black cycle.
// ---------------------------------------------------------------------------------------------------------------
template <bool JACOBI, bool FIXED, bool RF>
__global__ void __launch_bounds__(ST_THREADS) k_update2(Level lv, double* __restrict__ u, const double* __restrict__ p,
    double* __restrict__ rvec, float* __restrict__ rf, BandScalars* __restrict__ scal, int k)
{
    __shared__ double s_red[4];
    BandScalars& sc = scal[blockIdx.y];
    if (sc.done)
        return;
    const int slot = k & 3, next = (k + 1) & 3;
    const double alpha = sc.rz[slot] / sc.pq[slot];  // ConjugateGradient.h:68
    int ty, tx;
    const Strip s = make_strip(lv, blockIdx.x, ty, tx);
    const int64_t origin = (int64_t)blockIdx.y * lv.plane + (int64_t)ty * TILE_H * lv.pitch + (int64_t)tx * TILE_W;
    const double* pb = p + origin;
    double* ub = u + origin;
    double* rb = rvec + origin;
    const unsigned any = s.mL | s.mR;
    const bool west = s.cx == 0, east = s.cx == 15;
    double2 pv[6], xv[ST_RG], rv[ST_RG];
    double pe[ST_RG];
#pragma unroll
    for (int j = 0; j < 6; ++j)
        pv[j] = ldnc2_if(pb + (s.toff + j * s.pitch), (any >> j) & 1);
#pragma unroll
    for (int j = 0; j < ST_RG; ++j) {
        xv[j] = ld2_if(ub + (s.toff + (j + 1) * s.pitch), (any >> (j + 1)) & 1);
        rv[j] = ld2_if(rb + (s.toff + (j + 1) * s.pitch), (any >> (j + 1)) & 1);
    }
    {
        const int eoff = s.toff + (west ? -1 : 2);
        const unsigned em = (west ? s.mL : (east ? s.mR : 0u)) >> 1;
#pragma unroll
        for (int j = 0; j < ST_RG; ++j)
            pe[j] = ldnc_if(pb + (eoff + (j + 1) * s.pitch), (em >> j) & 1);
    }
    int dcL, dcR;
    diag_cols<FIXED>(lv, s, dcL, dcR);
    double r2 = 0.0, rz = 0.0;
#pragma unroll
    for (int j = 1; j <= ST_RG; ++j) {
        double wl = __shfl_up_sync(0xffffffffu, pv[j].y, 1);
        double er = __shfl_down_sync(0xffffffffu, pv[j].x, 1);
        if (west)
            wl = pe[j - 1];
        if (east)
            er = pe[j - 1];
        int dr = diag_row<FIXED>(lv, s.gr - 1 + j);
        double ql = (double)(dr + dcL) * pv[j].x - ((pv[j - 1].x + pv[j + 1].x) + (wl + pv[j].y));
        double qr = (double)(dr + dcR) * pv[j].y - ((pv[j - 1].y + pv[j + 1].y) + (pv[j].x + er));
        // a cell of the pair that is not an unknown keeps its value: p is zero there, and its r stays zero
        double2 xn, rn;
        xn.x = ((s.mL >> j) & 1) ? xv[j - 1].x + alpha * pv[j].x : xv[j - 1].x;                   // ConjugateGradient.h:69
        xn.y = ((s.mR >> j) & 1) ? xv[j - 1].y + alpha * pv[j].y : xv[j - 1].y;
        rn.x = ((s.mL >> j) & 1) ? rv[j - 1].x - alpha * ql : 0.0;                                // ConjugateGradient.h:70
        rn.y = ((s.mR >> j) & 1) ? rv[j - 1].y - alpha * qr : 0.0;
        if ((any >> j) & 1) {
            const int off = s.toff + j * s.pitch;
            *reinterpret_cast<double2*>(ub + off) = xn;
            *reinterpret_cast<double2*>(rb + off) = rn;
            if (RF)
                *reinterpret_cast<float2*>(rf + origin + off) = make_float2((float)rn.x, (float)rn.y);
        }
        r2 += rn.x * rn.x + rn.y * rn.y;
        if (JACOBI)
            rz += rn.x * rn.x * inv_of(dr + dcL) + rn.y * rn.y * inv_of(dr + dcR);
    }
    double tot = block_sum4(r2, s_red);
    if (threadIdx.x == 0 && tot != 0.0)
        atomicAdd(&sc.rr[next], tot);
    if (JACOBI) {
        __syncthreads();
        tot = block_sum4(rz, s_red);
        if (threadIdx.x == 0 && tot != 0.0)
            atomicAdd(&sc.rz[next], tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
int launch_direction2(sa_ctx* ctx, const Level& lv, int nbands, bool jacobi, const void* zin, bool z_is_float,
    const double* p_old, double* p_new, BandScalars* scal, int k)
{
    if (lv.n_tiles == 0)
        return SA_OK;
    dim3 grid((unsigned)lv.n_tiles, (unsigned)nbands);
    if (jacobi) {
        if (lv.fixed_diag)
            SA_LAUNCH(ctx, (k_direction2<true, true, double>), grid, ST_THREADS, 0, lv, (const double*)zin, p_old, p_new, scal, k);
        else
            SA_LAUNCH(ctx, (k_direction2<true, false, double>), grid, ST_THREADS, 0, lv, (const double*)zin, p_old, p_new, scal, k);
    } else if (z_is_float) {
        if (lv.fixed_diag)
            SA_LAUNCH(ctx, (k_direction2<false, true, float>), grid, ST_THREADS, 0, lv, (const float*)zin, p_old, p_new, scal, k);
        else
            SA_LAUNCH(ctx, (k_direction2<false, false, float>), grid, ST_THREADS, 0, lv, (const float*)zin, p_old, p_new, scal, k);
    } else {
        if (lv.fixed_diag)
            SA_LAUNCH(ctx, (k_direction2<false, true, double>), grid, ST_THREADS, 0, lv, (const double*)zin, p_old, p_new, scal, k);
        else
            SA_LAUNCH(ctx, (k_direction2<false, false, double>), grid, ST_THREADS, 0, lv, (const double*)zin, p_old, p_new, scal, k);
    }
    return SA_OK;
}

int launch_update2(sa_ctx* ctx, const Level& lv, int nbands, bool jacobi, double* u, const double* p, double* r, float* rf,
    BandScalars* scal, int k)
{
    if (lv.n_tiles == 0)
        return SA_OK;
    dim3 grid((unsigned)lv.n_tiles, (unsigned)nbands);
    if (jacobi) {
        if (lv.fixed_diag)
            SA_LAUNCH(ctx, (k_update2<true, true, false>), grid, ST_THREADS, 0, lv, u, p, r, rf, scal, k);
        else
            SA_LAUNCH(ctx, (k_update2<true, false, false>), grid, ST_THREADS, 0, lv, u, p, r, rf, scal, k);
    } else if (lv.fixed_diag) {
        if (rf)
            SA_LAUNCH(ctx, (k_update2<false, true, true>), grid, ST_THREADS, 0, lv, u, p, r, rf, scal, k);
        else
            SA_LAUNCH(ctx, (k_update2<false, true, false>), grid, ST_THREADS, 0, lv, u, p, r, rf, scal, k);
    } else {
        if (rf)
            SA_LAUNCH(ctx, (k_update2<false, false, true>), grid, ST_THREADS, 0, lv, u, p, r, rf, scal, k);
        else
            SA_LAUNCH(ctx, (k_update2<false, false, false>), grid, ST_THREADS, 0, lv, u, p, r, rf, scal, k);
    }
    return SA_OK;
}

}  // namespace satfill
