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

#include <cstdlib>

namespace satfill {

namespace {

#ifndef SATFILL_ST_RG
#define SATFILL_ST_RG 4
#endif
constexpr int ST_RG = SATFILL_ST_RG;  // rows per thread (4: 128-thread CTAs, 2: 256-thread CTAs with half the registers)
constexpr int ST_THREADS = 16 * (TILE_H / ST_RG);
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_NR = ST_RG + 2;                    // rows a thread looks at: its own and one above / below
constexpr unsigned ST_NRM = (1u << ST_NR) - 1;      // ... as a bit mask (bit j <=> tile row row0 - 1 + j)
constexpr unsigned ST_OWN = ((1u << ST_RG) - 1) << 1;  // the own rows among them
static_assert(TILE_H % ST_RG == 0 && ST_THREADS % 32 == 0 && ST_NR <= 8, "unsupported rows per thread");
// resident CTAs per SM the two kernels are compiled for (register budget) and launched with (persistent grids)
#ifndef SATFILL_DIR_CTAS
#define SATFILL_DIR_CTAS 6
#endif
#ifndef SATFILL_UPD_CTAS
#define SATFILL_UPD_CTAS 5
#endif
#ifndef SATFILL_L2_PREFETCH
#define SATFILL_L2_PREFETCH 0
#endif
#ifndef SATFILL_SECTOR_STORES
#define SATFILL_SECTOR_STORES 1
#endif
constexpr int ST_DIR_CTAS = SATFILL_DIR_CTAS, ST_UPD_CTAS = SATFILL_UPD_CTAS;
// a double z costs 12 more registers than a float one
template <typename ZT>
constexpr int dir_ctas() { return sizeof(ZT) == 8 && ST_DIR_CTAS > 6 ? 6 : ST_DIR_CTAS; }

__device__ __forceinline__ double2 ldnc2_if(const double* p, unsigned mask, unsigned bit)
{
    double2 v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@q ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(v.x), "=d"(v.y)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ double2 ld2_if(const double* p, unsigned mask, unsigned bit)  // data this kernel also writes: no .nc
{
    double2 v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "@q ld.global.v2.f64 {%0, %1}, [%2];\n\t}"
        : "=d"(v.x), "=d"(v.y)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ double ldnc_if(const double* p, unsigned mask, unsigned bit)
{
    double v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}"
        : "=d"(v)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ double ld_if(const double* p, unsigned mask, unsigned bit)  // data this kernel also writes: no .nc
{
    double v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@q ld.global.f64 %0, [%1];\n\t}"
        : "=d"(v)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ double2 ldnc2_if(const float* p, unsigned mask, unsigned bit)  // float plane, widened
{
    float x, y;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
        "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
        : "=f"(x), "=f"(y)
        : "l"(p), "r"(mask), "r"(bit));
    return make_double2((double)x, (double)y);
}
__device__ __forceinline__ float2 ldnc2f_if(const float* p, unsigned mask, unsigned bit)  // float plane, as stored
{
    float2 v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
        "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
        : "=f"(v.x), "=f"(v.y)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ float ldncf_if(const float* p, unsigned mask, unsigned bit)
{
    float v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ double ldnc_if(const float* p, unsigned mask, unsigned bit)
{
    float v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"(mask), "r"(bit));
    return (double)v;
}

__device__ __forceinline__ void store_pair(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ void store_pair(float* p, double2 v) { *reinterpret_cast<float2*>(p) = make_float2((float)v.x, (float)v.y); }

// What a thread needs to know about a tile: its coordinates and the unknown bits of the thread's two columns.
struct TileBits {
    int yx;      // ty << 16 | tx
    unsigned m;  // mL | mR << 8: unknown bits of the two columns, bit j <=> tile row row0 - 1 + j, j = 0..5;
                 // bits 16 .. 16 + ST_RG - 1 (edge lanes): the halo column's cell of own row j is an unknown
    __device__ __forceinline__ int ty() const { return yx >> 16; }
    __device__ __forceinline__ int tx() const { return yx & 0xffff; }
    __device__ __forceinline__ unsigned mL() const { return m & ST_NRM; }
    __device__ __forceinline__ unsigned mR() const { return (m >> 8) & ST_NRM; }
    __device__ __forceinline__ unsigned any() const { return (m | (m >> 8)) & ST_NRM; }
    // own rows (bit j <=> row0 + j) whose halo-column neighbour -- column -1 for the west lane, column 32 for the east lane --
    // is an unknown: a halo value is only ever loaded where it is one (the vectors are zero elsewhere BY DEFINITION, but
    // after a mask change only the sectors that hold an unknown of the new mask have been rewritten: cg.cu, stale_all)
    __device__ __forceinline__ unsigned halo() const { return (m >> 16) & ((1u << ST_RG) - 1); }
    // element offset of the tile's origin inside a band plane (fits 32 bits: a plane has < 2^31 elements)
    __device__ __forceinline__ int origin(int pitch) const { return ty() * (TILE_H * pitch) + tx() * TILE_W; }
};

// Column masks of the thread's aligned column pair (2cx, 2cx + 1), rows row0 - 1 .. row0 + 4, from the transposed
// per-tile bit words of the tile and of the tiles above and below it: three 8-byte loads that hit L2 (or L1).
__device__ __forceinline__ TileBits load_tile_bits(const Level& lv, int yx, int cx, int row0)
{
    TileBits b;
    b.yx = yx;
    const uint32_t* w = lv.tbitsT + ((size_t)(b.ty() + 1) * lv.tb_stride + (b.tx() + 1)) * 32 + 2 * cx;
    const size_t vs = (size_t)lv.tb_stride * 32;
    const uint2 C = __ldg(reinterpret_cast<const uint2*>(w));
    const uint2 N = __ldg(reinterpret_cast<const uint2*>(w - vs));
    const uint2 S = __ldg(reinterpret_cast<const uint2*>(w + vs));
    // bit (row + 1) of the 34-bit column <=> tile row `row`
    unsigned long long cl = ((unsigned long long)N.x >> 31) | ((unsigned long long)C.x << 1) | ((unsigned long long)(S.x & 1u) << 33);
    unsigned long long cr = ((unsigned long long)N.y >> 31) | ((unsigned long long)C.y << 1) | ((unsigned long long)(S.y & 1u) << 33);
    b.m = ((unsigned)(cl >> row0) & ST_NRM) | (((unsigned)(cr >> row0) & ST_NRM) << 8);
    if (cx == 0 || cx == 15) {  // the halo column: column 31 of the tile to the west, column 0 of the tile to the east
        const uint32_t h = __ldg(lv.tbitsT + ((size_t)(b.ty() + 1) * lv.tb_stride + (b.tx() + (cx == 0 ? 0 : 2))) * 32 + (cx == 0 ? 31 : 0));
        b.m |= ((h >> row0) & ((1u << ST_RG) - 1)) << 16;
    }
    return b;
}

__device__ __forceinline__ double block_sum4(double v, double* s_red /* ST_WARPS */)
{
    for (int o = 16; o; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // s_red may still be read from the previous band
    if ((threadIdx.x & 31) == 0)
        s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < ST_WARPS; w += 2)
        t += s_red[w] + s_red[w + 1];
    return t;  // valid everywhere
}

template <bool FIXED>
__device__ __forceinline__ int diag_col(const Level& lv, int64_t c)
{
    return FIXED ? 2 : (c > 0) + (c < lv.cols - 1);
}
template <bool FIXED>
__device__ __forceinline__ int diag_row(const Level& lv, int64_t r)
{
    return FIXED ? 2 : (r > 0) + (r < lv.rows - 1);
}
__device__ __forceinline__ double inv_of(int d) { return d == 4 ? 0.25 : (d == 3 ? (1.0 / 3.0) : (d == 2 ? 0.5 : 1.0)); }

// The persistent tile loop shared by both kernels: CTA c visits tiles c, c + G, c + 2G, ... of the level's raster-ordered
// list (neighbouring CTAs work on neighbouring tiles at the same time, so halo rows meet in L2).  Bytes in flight are
// what bounds these kernels on cloud-like masks (half of a tile's loads are predicated off, and the register file caps
// what a resident thread can have outstanding), so the loop runs a three-deep pipeline that costs no registers for
// data: tile coordinates are fetched three visits ahead, the column masks two visits ahead, and `prefetch(tb)` asks L2
// for the data of the NEXT visit (prefetch.global.L2, predicated like the loads) right after the current tile's loads
// have been issued.  The dependent chain per tile is then  L2 -> arithmetic -> store  instead of
// list -> masks -> HBM -> arithmetic -> store.
template <typename Body, typename Prefetch>
__device__ __forceinline__ void for_each_tile(const Level& lv, int cx, int row0, Body body, Prefetch prefetch)
{
    const int n = lv.n_tiles, G = (int)gridDim.x;
    const int last = n - 1;
    int i1 = (int)blockIdx.x + G, i2 = i1 + G;
    TileBits tb = load_tile_bits(lv, lv.tile_yx[blockIdx.x], cx, row0);
    TileBits tb1 = load_tile_bits(lv, lv.tile_yx[i1 < n ? i1 : last], cx, row0);
    int yx2 = lv.tile_yx[i2 < n ? i2 : last];
    while (true) {
        const int i3 = i2 + G;
        const int yx3 = lv.tile_yx[i3 < n ? i3 : last];
        const TileBits tb2 = load_tile_bits(lv, yx2, cx, row0);
        body(tb, tb1, i1 < n, prefetch);
        if (i1 >= n)
            break;
        tb = tb1;
        tb1 = tb2;
        yx2 = yx3;
        i1 = i2;
        i2 = i3;
    }
}

// OR of a row-bit mask over the lanes that share a 32-byte sector: two lanes for double pairs, four for float pairs
__device__ __forceinline__ unsigned sector_or2(unsigned m)
{
    return SATFILL_SECTOR_STORES ? (m | __shfl_xor_sync(0xffffffffu, m, 1)) : m;
}
__device__ __forceinline__ unsigned sector_or4(unsigned m2)
{
    return SATFILL_SECTOR_STORES ? (m2 | __shfl_xor_sync(0xffffffffu, m2, 2)) : m2;
}

__device__ __forceinline__ void prefetch_l2_if(const void* p, unsigned pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q prefetch.global.L2 [%0];\n\t}" ::"l"(p), "r"(pred));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// k_direction2:  beta = rz_k / rz_{k-1};  p' = z + beta p  (halo cells recomputed);  pq = p'.Ap'
//   JACOBI: z = r / d on the fly (zin = r).
// Persistent: gridDim.x <= n_tiles CTAs, each walks all bands (a band that has converged costs one flag read).
// ---------------------------------------------------------------------------------------------------------------
//   PT = float: the search direction is STORED in single precision (the multigrid path, whose z is single precision
//   anyway).  x, r, A p and every dot product stay double and use the rounded p, so r = b - A x holds as exactly as
//   before; the rounding only perturbs the direction by 6e-8 relative, which CG does not notice
//   (tools/mg_prototype.py: identical iteration counts down to a 1e-14 residual).
template <bool JACOBI, bool FIXED, typename ZT, typename PT>
__global__ void __launch_bounds__(ST_THREADS, dir_ctas<ZT>()) k_direction2(Level lv, int nbands, const ZT* __restrict__ zin,
    const PT* __restrict__ p_old, PT* __restrict__ p_new, BandScalars* __restrict__ scal, int k)
{
    __shared__ double s_red[ST_WARPS];
    const int slot = k & 3;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const int cx = threadIdx.x & 15, row0 = (threadIdx.x >> 4) * ST_RG;
    const int pitch = (int)lv.pitch;
    const int toff = (row0 - 1) * pitch + 2 * cx;  // (row0 - 1, left cell) from the tile's origin
    const bool west = cx == 0, east = cx == 15;
    for (int band = 0; band < nbands; ++band) {
        BandScalars& sc = scal[band];
        if (sc.done)
            continue;
        if (k > 0 && sc.rr[slot] < sc.thr) {  // ConjugateGradient.h:72-73 (strict <), tested one launch later
            if (lead) {
                sc.rr_exit = sc.rr[slot];
                sc.iters = k - 1;
                __threadfence();
                sc.done = 1;
            }
            continue;
        }
        const double beta = k > 0 ? sc.rz[slot] / sc.rz[(k - 1) & 3] : 0.0;  // ConjugateGradient.h:77-79
        if (lead) {  // recycle the slot two iterations ahead
            int z2 = (k + 2) & 3;
            sc.rz[z2] = 0.0;
            sc.rr[z2] = 0.0;
            sc.pq[z2] = 0.0;
        }
        const int64_t band_off = (int64_t)band * lv.plane;
        const ZT* zband = zin + band_off;
        const PT* pband = p_old + band_off;
        PT* poband = p_new + band_off;
        double acc = 0.0;
        // L2 prefetch of the next visit's tile: the own rows of z and p (the halo rows are other threads' own rows)
        auto prefetch = [&](const TileBits& nx) {
            const int o = nx.origin(pitch) + toff;
            const unsigned any = nx.any();
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j) {
                prefetch_l2_if(zband + (o + j * pitch), (any >> j) & 1);
                prefetch_l2_if(pband + (o + j * pitch), (any >> j) & 1);
            }
        };
        for_each_tile(lv, cx, row0, [&](const TileBits& tb, const TileBits& nx, bool has_next, auto& pf_next) {
            const int origin = tb.origin(pitch);
            const ZT* zb = zband + origin;
            const PT* pb = pband + origin;
            const unsigned any = tb.any();
            const int64_t gr = (int64_t)tb.ty() * TILE_H + row0, gc = (int64_t)tb.tx() * TILE_W + 2 * cx;
            const int dcL = diag_col<FIXED>(lv, gc), dcR = diag_col<FIXED>(lv, gc + 1);
            const int eoff = toff + (west ? -1 : 2);
            // the halo cell can only matter if the own edge cell is an unknown -- and holds a value only if it is one itself
            const unsigned em = ((west ? tb.mL() : (east ? tb.mR() : 0u)) >> 1) & tb.halo();
            double2 pn[ST_NR];
            double pedge[ST_RG];
            if constexpr (sizeof(PT) == 4 && sizeof(ZT) == 4 && !JACOBI) {
                // float z, float p: p' = z + beta p in single precision, exactly as it is stored (conversions between
                // float and double are a quarter-rate pipe: two per cell instead of eight); widened once for A p'
                float2 zf[ST_NR], pf[ST_NR];
                float zef[ST_RG], pef[ST_RG];
#pragma unroll
                for (int j = 0; j < ST_NR; ++j) {
                    zf[j] = ldnc2f_if(zb + (toff + j * pitch), any, 1u << j);
                    pf[j] = ldnc2f_if(pb + (toff + j * pitch), any, 1u << j);
                }
#pragma unroll
                for (int j = 0; j < ST_RG; ++j) {
                    zef[j] = ldncf_if(zb + (eoff + (j + 1) * pitch), em, 1u << j);
                    pef[j] = ldncf_if(pb + (eoff + (j + 1) * pitch), em, 1u << j);
                }
                if (SATFILL_L2_PREFETCH && has_next)
                    pf_next(nx);
                const float bf = (float)beta;
                const unsigned st = sector_or4(sector_or2(any));
                PT* po = poband + origin;
#pragma unroll
                for (int j = 0; j < ST_NR; ++j) {
                    float2 v = make_float2(fmaf(bf, pf[j].x, zf[j].x), fmaf(bf, pf[j].y, zf[j].y));  // ConjugateGradient.h:80
                    if (j >= 1 && j <= ST_RG && ((st >> j) & 1))
                        *reinterpret_cast<float2*>(po + (toff + j * pitch)) = v;
                    pn[j] = make_double2((double)v.x, (double)v.y);
                }
#pragma unroll
                for (int j = 0; j < ST_RG; ++j)
                    pedge[j] = (double)fmaf(bf, pef[j], zef[j]);
            } else {
                // ---- all loads: pairs of rows row0-1 .. row0+4, and (edge lanes) the halo column of the own rows
                double2 zv[ST_NR], pv[ST_NR];
                double ze[ST_RG], pe[ST_RG];
#pragma unroll
                for (int j = 0; j < ST_NR; ++j) {
                    zv[j] = ldnc2_if(zb + (toff + j * pitch), any, 1u << j);
                    pv[j] = ldnc2_if(pb + (toff + j * pitch), any, 1u << j);
                }
#pragma unroll
                for (int j = 0; j < ST_RG; ++j) {
                    ze[j] = ldnc_if(zb + (eoff + (j + 1) * pitch), em, 1u << j);
                    pe[j] = ldnc_if(pb + (eoff + (j + 1) * pitch), em, 1u << j);
                }
                if (SATFILL_L2_PREFETCH && has_next)
                    pf_next(nx);
                // ---- p' = z + beta p on the 6 x 2 cells and the edge column
#pragma unroll
                for (int j = 0; j < ST_NR; ++j) {
                    double zl = zv[j].x, zr = zv[j].y;
                    if (JACOBI) {
                        int dr = diag_row<FIXED>(lv, gr - 1 + j);
                        zl *= inv_of(dr + dcL);
                        zr *= inv_of(dr + dcR);
                    }
                    pn[j].x = zl + beta * pv[j].x;  // ConjugateGradient.h:80
                    pn[j].y = zr + beta * pv[j].y;
                    if (sizeof(PT) == 4) {  // p'.Ap' of the direction as it is stored
                        pn[j].x = (double)(float)pn[j].x;
                        pn[j].y = (double)(float)pn[j].y;
                    }
                }
#pragma unroll
                for (int j = 0; j < ST_RG; ++j) {
                    double z = ze[j];
                    if (JACOBI)
                        z *= inv_of(diag_row<FIXED>(lv, gr + j) + diag_col<FIXED>(lv, west ? gc - 1 : gc + 2));
                    pedge[j] = z + beta * pe[j];
                    if (sizeof(PT) == 4)
                        pedge[j] = (double)(float)pedge[j];
                }
                PT* po = poband + origin;
                const unsigned st = sizeof(PT) == 4 ? sector_or4(sector_or2(any)) : sector_or2(any);
#pragma unroll
                for (int j = 1; j <= ST_RG; ++j)
                    if ((st >> j) & 1)
                        store_pair(po + (toff + j * pitch), pn[j]);
            }
            // ---- accumulate p'.Ap' over the own rows (whole 32-byte sectors were stored above: a partially written
            //      sector costs HBM a read-modify-write; the extra cells are not unknowns and get the zero the invariant
            //      demands)
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j) {
                double wl = __shfl_up_sync(0xffffffffu, pn[j].y, 1);    // lane - 1's right cell
                double er = __shfl_down_sync(0xffffffffu, pn[j].x, 1);  // lane + 1's left cell
                if (west)
                    wl = pedge[j - 1];
                if (east)
                    er = pedge[j - 1];
                int dr = diag_row<FIXED>(lv, gr - 1 + j);
                double ql = (double)(dr + dcL) * pn[j].x - ((pn[j - 1].x + pn[j + 1].x) + (wl + pn[j].y));
                double qr = (double)(dr + dcR) * pn[j].y - ((pn[j - 1].y + pn[j + 1].y) + (pn[j].x + er));
                acc += pn[j].x * ql + pn[j].y * qr;  // p' is zero outside the unknown set: no mask needed
            }
        }, prefetch);
        double tot = block_sum4(acc, s_red);
        if (threadIdx.x == 0 && tot != 0.0)
            atomicAdd(&sc.pq[slot], tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_update2:  alpha = rz / pq;  x += alpha p;  r -= alpha A p  (A p recomputed);  |r|^2 and (JACOBI) r.(r/d)
//   RF: also write the residual as float for the red-black cycle.
// ---------------------------------------------------------------------------------------------------------------
//   XM (the float-direction path): x is only ever accumulated, so its 16 bytes per unknown travel every OTHER pass.
//     XM = 1 (even passes): x is neither loaded nor stored; alpha stays in alpha_hist[k & 1], the direction in its buffer
//     (the next k_direction2 writes the OTHER buffer), pend is raised.
//     XM = 2 (odd passes): x += alpha_{k-1} p_{k-1} + alpha_k p_k (4 more bytes: p_{k-1}), pend is lowered.
//     A band whose last pass was an XM = 1 pass gets its step from k_flush_x after the loop.
template <bool JACOBI, bool FIXED, bool RF, typename PT, int XM = 0>
__global__ void __launch_bounds__(ST_THREADS, ST_UPD_CTAS) k_update2(Level lv, int nbands, double* __restrict__ u,
    const PT* __restrict__ p, double* __restrict__ rvec, float* __restrict__ rf, BandScalars* __restrict__ scal, int k,
    const PT* __restrict__ p_prev = nullptr)
{
    __shared__ double s_red[ST_WARPS];
    const int slot = k & 3, next = (k + 1) & 3;
    const int cx = threadIdx.x & 15, row0 = (threadIdx.x >> 4) * ST_RG;
    const int pitch = (int)lv.pitch;
    const int toff = (row0 - 1) * pitch + 2 * cx;
    const bool west = cx == 0, east = cx == 15;
    for (int band = 0; band < nbands; ++band) {
        BandScalars& sc = scal[band];
        if (sc.done)
            continue;
        const double alpha = sc.rz[slot] / sc.pq[slot];  // ConjugateGradient.h:68
        const double alpha_prev = XM == 2 ? sc.alpha_hist[(k + 1) & 1] : 0.0;  // written by the previous launch
        if (XM != 0 && blockIdx.x == 0 && threadIdx.x == 0) {  // every CTA visits every live band: one writer
            sc.alpha_hist[k & 1] = alpha;
            sc.pend = XM == 1 ? 1 : 0;
            sc.pend_buf = (k + 1) & 1;  // cg.cu: pass k's direction lives in p buffer (k + 1) & 1
        }
        const int64_t band_off = (int64_t)band * lv.plane;
        double r2 = 0.0, rz = 0.0;
        const PT* pband = p + band_off;
        const PT* qband = XM == 2 ? p_prev + band_off : nullptr;
        double* uband = u + band_off;
        double* rband = rvec + band_off;
        float* rfband = RF ? rf + band_off : nullptr;
        auto prefetch = [&](const TileBits& nx) {
            const int o = nx.origin(pitch) + toff;
            const unsigned any = nx.any();
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j) {
                prefetch_l2_if(pband + (o + j * pitch), (any >> j) & 1);
                prefetch_l2_if(uband + (o + j * pitch), (any >> j) & 1);
                prefetch_l2_if(rband + (o + j * pitch), (any >> j) & 1);
            }
        };
        for_each_tile(lv, cx, row0, [&](const TileBits& tb, const TileBits& nx, bool has_next, auto& pf) {
            const int origin = tb.origin(pitch);
            const PT* pb = pband + origin;
            double* ub = uband + origin;
            double* rb = rband + origin;
            const unsigned any = tb.any();
            const unsigned mL = tb.mL(), mR = tb.mR();
            // whole 32-byte sectors are written (see k_direction2): x of a cell that is not an unknown is written back as
            // read, so x is loaded wherever its sector is stored (the sector travels anyway)
            const unsigned st2 = sector_or2(any), st4 = RF ? sector_or4(st2) : 0u;
            double2 pv[ST_NR], xv[ST_RG], rv[ST_RG], qv[ST_RG];
            double pe[ST_RG];
#pragma unroll
            for (int j = 0; j < ST_NR; ++j)
                pv[j] = ldnc2_if(pb + (toff + j * pitch), any, 1u << j);
#pragma unroll
            for (int j = 0; j < ST_RG; ++j) {
                if (XM != 1)
                    xv[j] = ld2_if(ub + (toff + (j + 1) * pitch), st2, 1u << (j + 1));
                if (XM == 2)
                    qv[j] = ldnc2_if(qband + origin + (toff + (j + 1) * pitch), any, 1u << (j + 1));
                rv[j] = ld2_if(rb + (toff + (j + 1) * pitch), any, 1u << (j + 1));
            }
            {
                const int eoff = toff + (west ? -1 : 2);
                const unsigned em = ((west ? mL : (east ? mR : 0u)) >> 1) & tb.halo();
#pragma unroll
                for (int j = 0; j < ST_RG; ++j)
                    pe[j] = ldnc_if(pb + (eoff + (j + 1) * pitch), em, 1u << j);
            }
            if (SATFILL_L2_PREFETCH && has_next)
                pf(nx);
            const int64_t gr = (int64_t)tb.ty() * TILE_H + row0, gc = (int64_t)tb.tx() * TILE_W + 2 * cx;
            const int dcL = diag_col<FIXED>(lv, gc), dcR = diag_col<FIXED>(lv, gc + 1);
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j) {
                double wl = __shfl_up_sync(0xffffffffu, pv[j].y, 1);
                double er = __shfl_down_sync(0xffffffffu, pv[j].x, 1);
                if (west)
                    wl = pe[j - 1];
                if (east)
                    er = pe[j - 1];
                int dr = diag_row<FIXED>(lv, gr - 1 + j);
                double ql = (double)(dr + dcL) * pv[j].x - ((pv[j - 1].x + pv[j + 1].x) + (wl + pv[j].y));
                double qr = (double)(dr + dcR) * pv[j].y - ((pv[j - 1].y + pv[j + 1].y) + (pv[j].x + er));
                // a cell of the pair that is not an unknown keeps its value: p is zero there, and its r stays zero
                double2 xn, rn;
                if (XM != 1) {
                    // XM = 2: the step of the previous pass first, as if it had been added then
                    const double2 x0 = XM == 2 ? make_double2(xv[j - 1].x + alpha_prev * qv[j - 1].x, xv[j - 1].y + alpha_prev * qv[j - 1].y)
                                               : xv[j - 1];
                    xn.x = ((mL >> j) & 1) ? x0.x + alpha * pv[j].x : xv[j - 1].x;                      // ConjugateGradient.h:69
                    xn.y = ((mR >> j) & 1) ? x0.y + alpha * pv[j].y : xv[j - 1].y;
                }
                rn.x = ((mL >> j) & 1) ? rv[j - 1].x - alpha * ql : 0.0;                                // ConjugateGradient.h:70
                rn.y = ((mR >> j) & 1) ? rv[j - 1].y - alpha * qr : 0.0;
                const int off = toff + j * pitch;
                if ((st2 >> j) & 1) {
                    if (XM != 1)
                        *reinterpret_cast<double2*>(ub + off) = xn;
                    *reinterpret_cast<double2*>(rb + off) = rn;
                }
                if (RF && ((st4 >> j) & 1))
                    *reinterpret_cast<float2*>(rfband + origin + off) = make_float2((float)rn.x, (float)rn.y);
                r2 += rn.x * rn.x + rn.y * rn.y;
                if (JACOBI)
                    rz += rn.x * rn.x * inv_of(dr + dcL) + rn.y * rn.y * inv_of(dr + dcR);
            }
        }, prefetch);
        double tot = block_sum4(r2, s_red);
        if (threadIdx.x == 0 && tot != 0.0)
            atomicAdd(&sc.rr[next], tot);
        if (JACOBI) {
            tot = block_sum4(rz, s_red);
            if (threadIdx.x == 0 && tot != 0.0)
                atomicAdd(&sc.rz[next], tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_setup2: the set-up of a solve in one pass over the active tiles (replaces k_init_guess + k_residual of cg.cu):
//   x0 (Laplace: 0, IterativeSolverBase.h:357-360; Poisson: the replacement image, poisson.cpp:239,257) written into
//   the unknown cells of u;  r0 = b - A x0;  |b|^2, |r0|^2, r0.(r0/d)  (laplace.cpp:71-94 / poisson.cpp:241-251 for b,
//   ConjugateGradient.h:38-61 for the rest).  With x0 = g the residual collapses to a sum over the KNOWN neighbours:
//       r0_p = sum_{q in N(p) known} (f_q - g_q)        (Laplace: g = 0, so r0 = b)
//       b_p  = d_p g_p - sum_{q in N(p)} g_q + sum_{q in N(p) known} f_q
//   A neighbour outside the image is a "known" cell holding zeros (guard rows / columns), so it drops out by itself.
//   Unknown cells of u hold whatever the previous fill left there: they are only ever used through the known mask.
// ---------------------------------------------------------------------------------------------------------------
template <bool POISSON, bool RF>
__global__ void __launch_bounds__(ST_THREADS, 4) k_setup2(Level lv, int nbands, double* __restrict__ u,
    const double* __restrict__ g, double* __restrict__ rvec, float* __restrict__ rf, BandScalars* __restrict__ scal)
{
    constexpr bool FIXED = !POISSON;
    __shared__ double s_red[ST_WARPS];
    const int cx = threadIdx.x & 15, row0 = (threadIdx.x >> 4) * ST_RG;
    const int pitch = (int)lv.pitch;
    const int toff = (row0 - 1) * pitch + 2 * cx;
    const bool west = cx == 0, east = cx == 15;
    for (int band = 0; band < nbands; ++band) {
        BandScalars& sc = scal[band];
        const int64_t band_off = (int64_t)band * lv.plane;
        double* uband = u + band_off;
        const double* gband = POISSON ? g + band_off : nullptr;
        double* rband = rvec + band_off;
        float* rfband = RF ? rf + band_off : nullptr;
        double b2 = 0.0, r2 = 0.0, rz = 0.0;
        auto no_prefetch = [](const TileBits&) {};
        for_each_tile(lv, cx, row0, [&](const TileBits& tb, const TileBits&, bool, auto&) {
            const int origin = tb.origin(pitch);
            const unsigned mL = tb.mL(), mR = tb.mR(), any = tb.any();
            // unknown bits of the columns west of the pair's left cell and east of its right cell, own rows
            unsigned mW = __shfl_up_sync(0xffffffffu, mR, 1), mE = __shfl_down_sync(0xffffffffu, mL, 1);
            if (west || east) {
                const uint32_t w = __ldg(lv.tbitsT + ((size_t)(tb.ty() + 1) * lv.tb_stride + (tb.tx() + (west ? 0 : 2))) * 32 + (west ? 31 : 0));
                const unsigned e = ((w >> row0) & ((1u << ST_RG) - 1)) << 1;
                if (west)
                    mW = e;
                else
                    mE = e;
            }
            const unsigned own = any & ST_OWN;  // own rows that hold an unknown of the pair
            // rows to load: those, the rows above / below them, and the rows an adjacent lane's unknowns look at (u and g
            // hold real pixel values at known cells, unlike the solver's work vectors)
            const unsigned ldm = (own | (own << 1) | (own >> 1) | __shfl_up_sync(0xffffffffu, own, 1) | __shfl_down_sync(0xffffffffu, own, 1)) & ST_NRM;
            const int64_t gr = (int64_t)tb.ty() * TILE_H + row0, gc = (int64_t)tb.tx() * TILE_W + 2 * cx;
            const unsigned em = (west ? mL : (east ? mR : 0u)) >> 1;  // own rows whose edge cell is an unknown
            // of the image only KNOWN values are looked at (x0 replaces the unknown cells): a pair of two unknowns, or a
            // halo cell that is an unknown itself, is not loaded -- inside a hole that is every load of u
            const unsigned ldu = ldm & ~(mL & mR), emu = em & ~((west ? mW : mE) >> 1);
            double* ub = uband + origin;
            double2 uv[ST_NR], gv[ST_NR];
            double ue[ST_RG], ge[ST_RG];
#pragma unroll
            for (int j = 0; j < ST_NR; ++j) {
                uv[j] = ld2_if(ub + (toff + j * pitch), ldu, 1u << j);
                gv[j] = POISSON ? ldnc2_if(gband + origin + (toff + j * pitch), ldm, 1u << j) : make_double2(0.0, 0.0);
            }
            {
                const int eoff = toff + (west ? -1 : 2);
#pragma unroll
                for (int j = 0; j < ST_RG; ++j) {
                    ue[j] = ld_if(ub + (eoff + (j + 1) * pitch), emu, 1u << j);
                    ge[j] = POISSON ? ldnc_if(gband + origin + (eoff + (j + 1) * pitch), em, 1u << j) : 0.0;
                }
            }
            const int dcL = diag_col<FIXED>(lv, gc), dcR = diag_col<FIXED>(lv, gc + 1);
            const unsigned st2 = sector_or2(own), st4 = RF ? sector_or4(st2) : 0u;
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j) {
                double uw = __shfl_up_sync(0xffffffffu, uv[j].y, 1), ueast = __shfl_down_sync(0xffffffffu, uv[j].x, 1);
                double gw = 0.0, geast = 0.0;
                if (POISSON) {
                    gw = __shfl_up_sync(0xffffffffu, gv[j].y, 1);
                    geast = __shfl_down_sync(0xffffffffu, gv[j].x, 1);
                }
                if (west) {
                    uw = ue[j - 1];
                    gw = ge[j - 1];
                }
                if (east) {
                    ueast = ue[j - 1];
                    geast = ge[j - 1];
                }
                const int dr = diag_row<FIXED>(lv, gr - 1 + j);
                // known-neighbour sums of f and of (f - g): north, south, west, east
                auto knownsum = [&](double n, double s_, double w, double e, unsigned kn, unsigned ks, unsigned kw, unsigned ke) {
                    return ((kn ? 0.0 : n) + (ks ? 0.0 : s_)) + ((kw ? 0.0 : w) + (ke ? 0.0 : e));
                };
                const unsigned nL = (mL >> (j - 1)) & 1, sL = (mL >> (j + 1)) & 1, wL = (mW >> j) & 1, eL = (mR >> j) & 1;
                const unsigned nR = (mR >> (j - 1)) & 1, sR = (mR >> (j + 1)) & 1, wR = (mL >> j) & 1, eR = (mE >> j) & 1;
                const double fL = knownsum(uv[j - 1].x, uv[j + 1].x, uw, uv[j].y, nL, sL, wL, eL);
                const double fR = knownsum(uv[j - 1].y, uv[j + 1].y, uv[j].x, ueast, nR, sR, wR, eR);
                double bL = fL, bR = fR, resL = fL, resR = fR;
                if (POISSON) {
                    const double kgL = knownsum(gv[j - 1].x, gv[j + 1].x, gw, gv[j].y, nL, sL, wL, eL);
                    const double kgR = knownsum(gv[j - 1].y, gv[j + 1].y, gv[j].x, geast, nR, sR, wR, eR);
                    const double divL = (double)(dr + dcL) * gv[j].x - ((gv[j - 1].x + gv[j + 1].x) + (gw + gv[j].y));
                    const double divR = (double)(dr + dcR) * gv[j].y - ((gv[j - 1].y + gv[j + 1].y) + (gv[j].x + geast));
                    bL = divL + fL;
                    bR = divR + fR;
                    resL = fL - kgL;
                    resR = fR - kgR;
                }
                const bool unkL = (mL >> j) & 1, unkR = (mR >> j) & 1;
                if (!unkL)
                    bL = resL = 0.0;
                if (!unkR)
                    bR = resR = 0.0;
                b2 += bL * bL + bR * bR;
                r2 += resL * resL + resR * resR;
                rz += resL * resL * inv_of(dr + dcL) + resR * resR * inv_of(dr + dcR);
                const int off = toff + j * pitch;
                if ((st2 >> j) & 1) {
                    // x0 into the unknown cells; a known cell of the sector is written back as read
                    double2 xn = make_double2(unkL ? (POISSON ? gv[j].x : 0.0) : uv[j].x, unkR ? (POISSON ? gv[j].y : 0.0) : uv[j].y);
                    *reinterpret_cast<double2*>(ub + off) = xn;
                    *reinterpret_cast<double2*>(rband + origin + off) = make_double2(resL, resR);
                }
                if (RF && ((st4 >> j) & 1))
                    *reinterpret_cast<float2*>(rfband + origin + off) = make_float2((float)resL, (float)resR);
            }
        }, no_prefetch);
        double t = block_sum4(b2, s_red);
        if (threadIdx.x == 0 && t != 0.0)
            atomicAdd(&sc.bnorm2, t);
        t = block_sum4(r2, s_red);
        if (threadIdx.x == 0 && t != 0.0)
            atomicAdd(&sc.rr[0], t);
        t = block_sum4(rz, s_red);
        if (threadIdx.x == 0 && t != 0.0)
            atomicAdd(&sc.rz[0], t);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_scrub: zero work vectors at the unknowns of a level (whole sectors).  Run through the tile list of the PREVIOUS
// mask when the mask changes: a solve leaves its work vectors non-zero only at its own unknowns, so this restores the
// "zero outside the unknown set" invariant for any new mask at a third of the bytes of clearing the planes.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS) k_scrub(Level lv, ScrubPlanes P)
{
    const int cx = threadIdx.x & 15, row0 = (threadIdx.x >> 4) * ST_RG;
    const int pitch = (int)lv.pitch, pitch2 = pitch >> 1;
    const TileBits tb = load_tile_bits(lv, lv.tile_yx[blockIdx.x], cx, row0);
    const unsigned any = tb.any(), st2 = any | __shfl_xor_sync(0xffffffffu, any, 1), st4 = st2 | __shfl_xor_sync(0xffffffffu, st2, 2);
    const int64_t o = (int64_t)blockIdx.y * lv.plane + tb.origin(pitch) + (row0 - 1) * pitch + 2 * cx;
    const int64_t o2 = (int64_t)blockIdx.y * (lv.plane >> 1) + tb.ty() * (TILE_H * pitch2) + tb.tx() * (TILE_W / 2) + (row0 - 1) * pitch2 + cx;
#pragma unroll
    for (int j = 1; j <= ST_RG; ++j) {
        if ((st2 >> j) & 1)
            for (int q = 0; q < P.nd; ++q)
                *reinterpret_cast<double2*>(P.d[q] + o + j * pitch) = make_double2(0.0, 0.0);
        if ((st4 >> j) & 1)
            for (int q = 0; q < P.nf; ++q)
                *reinterpret_cast<float2*>(P.f[q] + o + j * pitch) = make_float2(0.f, 0.f);
        if ((any >> j) & 1)
            for (int q = 0; q < P.nh; ++q)
                P.h[q][o2 + j * pitch2] = 0.f;
    }
}

int launch_scrub(sa_ctx* ctx, const Level& lv, int nbands, const ScrubPlanes& planes)
{
    if (lv.n_tiles == 0 || planes.nd + planes.nf + planes.nh == 0)
        return SA_OK;
    SA_LAUNCH(ctx, k_scrub, dim3((unsigned)lv.n_tiles, (unsigned)nbands), ST_THREADS, 0, lv, planes);
    return SA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Grid: as many CTAs as are resident at once (the occupancy the kernel was compiled for x the SM count), never more than
// there are tiles, so that every CTA owns a tile of every band.
static unsigned strip_grid(const sa_ctx* ctx, const Level& lv, int ctas_per_sm)
{
    int g = ctx->grid_sms * ctas_per_sm;
    return (unsigned)(g < lv.n_tiles ? g : lv.n_tiles);
}

int launch_setup2(sa_ctx* ctx, const Level& lv, int nbands, bool poisson, double* u, const double* g, double* r, float* rf,
    BandScalars* scal)
{
    if (lv.n_tiles == 0)
        return SA_OK;
    const unsigned grid = strip_grid(ctx, lv, 4);
    if (poisson) {
        if (rf)
            SA_LAUNCH(ctx, (k_setup2<true, true>), grid, ST_THREADS, 0, lv, nbands, u, g, r, rf, scal);
        else
            SA_LAUNCH(ctx, (k_setup2<true, false>), grid, ST_THREADS, 0, lv, nbands, u, g, r, rf, scal);
    } else {
        if (rf)
            SA_LAUNCH(ctx, (k_setup2<false, true>), grid, ST_THREADS, 0, lv, nbands, u, g, r, rf, scal);
        else
            SA_LAUNCH(ctx, (k_setup2<false, false>), grid, ST_THREADS, 0, lv, nbands, u, g, r, rf, scal);
    }
    return SA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The io kernels of the direct mode run beside the solves (api.cu).  Both are bound by PCIe latency, not by anything an SM
// does, so each is a handful of 1024-thread CTAs: an io CTA has an SM to itself and the solve kernels, whose grids are
// sized for the SMs that are left (sa_ctx::grid_sms), never wait for a slot it holds.  (io CTAs sprinkled over many SMs
// cost every one of those SMs a solve CTA, and a solve CTA that does not fit runs as a second wave: measured +50 % on the
// solve with 16 x 256 threads of scatter beside it.)
//   fetch: reads of host memory, ~3 ms per band of a 10980^2 tile with 6 x 1024 threads, slower with fewer AND with many more;
//   scatter: posted writes, 64 warps keep PCIe busy (38 GB/s of unknown pixels alone, the same as tools/probe measures).
// ---------------------------------------------------------------------------------------------------------------
int io_ctas(bool scatter)
{
    static const int f = [] { const char* e = std::getenv("SATFILL_FETCH_CTAS"); return e && std::atoi(e) > 0 ? std::atoi(e) : 6; }();
    static const int sc = [] { const char* e = std::getenv("SATFILL_SCATTER_CTAS"); return e && std::atoi(e) > 0 ? std::atoi(e) : 2; }();
    return scatter ? sc : f;
}

// ---------------------------------------------------------------------------------------------------------------
// k_fetch_direct: the way in of the direct mode (api.cu).  The caller's page-locked arrays are device-addressable, so
// nothing is copied: this kernel reads, straight from host memory, exactly the pixels k_setup2 will look at -- a pair of
// cells that are both unknowns is never fetched, so what crosses PCIe is the ring of known pixels around the unknown set
// (and, for Poisson, g on the unknown set and around it) -- and drops them into the image (and guidance) planes.  It runs
// on a few CTAs beside the solve of the previous band window: the volume is tiny (2 % of the image on cloud-like masks)
// but every read is a PCIe round trip.  Known pixels elsewhere in the planes are stale and never read.
// ---------------------------------------------------------------------------------------------------------------
constexpr int IO_THREADS = 1024;  // an io CTA takes a whole SM (see io_ctas below)
template <bool POISSON>
__global__ void __launch_bounds__(IO_THREADS, 1) k_fetch_direct(Level lv, int nbands, double* __restrict__ u, double* __restrict__ g,
    HostBands src)
{
    // groups of ST_THREADS threads walk the tiles independently (no CTA-wide barrier; the shuffles stay inside a warp)
    constexpr int GROUPS = IO_THREADS / ST_THREADS;
    const int tid = threadIdx.x % ST_THREADS;
    const int cx = tid & 15, row0 = (tid >> 4) * ST_RG;
    const int pitch = (int)lv.pitch, fpitch = (int)src.pitch;
    const int toff = (row0 - 1) * pitch + 2 * cx, ftoff = (row0 - 1) * fpitch + 2 * cx;
    const bool west = cx == 0, east = cx == 15;
    // trip count is uniform over a group (and a warp lies inside one group): shuffles are safe
    for (int i = blockIdx.x * GROUPS + threadIdx.x / ST_THREADS; i < lv.n_tiles; i += gridDim.x * GROUPS) {
        const TileBits tb = load_tile_bits(lv, lv.tile_yx[i], cx, row0);
        const unsigned mL = tb.mL(), mR = tb.mR(), any = tb.any();
        unsigned mW = __shfl_up_sync(0xffffffffu, mR, 1), mE = __shfl_down_sync(0xffffffffu, mL, 1);
        if (west || east) {
            const uint32_t w = __ldg(lv.tbitsT + ((size_t)(tb.ty() + 1) * lv.tb_stride + (tb.tx() + (west ? 0 : 2))) * 32 + (west ? 31 : 0));
            const unsigned e = ((w >> row0) & ((1u << ST_RG) - 1)) << 1;
            if (west)
                mW = e;
            else
                mE = e;
        }
        const unsigned own = any & ST_OWN;
        unsigned ldm = (own | (own << 1) | (own >> 1) | __shfl_up_sync(0xffffffffu, own, 1) | __shfl_down_sync(0xffffffffu, own, 1)) & ST_NRM;
        const int64_t gr = (int64_t)tb.ty() * TILE_H + row0, gc = (int64_t)tb.tx() * TILE_W + 2 * cx;
        // the caller's array has no guard rows / padding columns: stay inside it (the planes hold zeros out there)
        unsigned inside = 0;
#pragma unroll
        for (int j = 0; j < ST_NR; ++j)
            inside |= (gr - 1 + j >= 0 && gr - 1 + j < src.rows && gc + 1 < src.cols) ? (1u << j) : 0u;
        ldm &= inside;
        unsigned em = (west ? mL : (east ? mR : 0u)) >> 1;  // own rows whose edge cell is an unknown
        const int64_t ec = west ? gc - 1 : gc + 2;
        if (ec < 0 || ec >= src.cols)
            em = 0;
        em &= inside >> 1;
        const unsigned ldf = ldm & ~(mL & mR);               // a pair of two unknowns holds nothing the equations read
        const unsigned emf = em & ~((west ? mW : mE) >> 1);  // nor does a halo cell that is an unknown itself
        if ((ldm | em) == 0)
            continue;
        const int origin = tb.origin(pitch), forigin = tb.origin(fpitch);
        const int eoff = toff + (west ? -1 : 2), feoff = ftoff + (west ? -1 : 2);
        for (int band = 0; band < nbands; ++band) {
            double* ub = u + (int64_t)band * lv.plane + origin;
            const double* fb = src.f[band] + forigin;
            double2 v[ST_NR];
            double e[ST_RG];
#pragma unroll
            for (int j = 0; j < ST_NR; ++j)
                v[j] = ldnc2_if(fb + (ftoff + j * fpitch), ldf, 1u << j);
#pragma unroll
            for (int j = 0; j < ST_RG; ++j)
                e[j] = ldnc_if(fb + (feoff + (j + 1) * fpitch), emf, 1u << j);
#pragma unroll
            for (int j = 0; j < ST_NR; ++j)
                if ((ldf >> j) & 1)
                    *reinterpret_cast<double2*>(ub + (toff + j * pitch)) = v[j];
#pragma unroll
            for (int j = 0; j < ST_RG; ++j)
                if ((emf >> j) & 1)
                    ub[eoff + (j + 1) * pitch] = e[j];
            if (POISSON) {
                double* gb = g + (int64_t)band * lv.plane + origin;
                const double* hg = src.g[band] + forigin;
#pragma unroll
                for (int j = 0; j < ST_NR; ++j)
                    v[j] = ldnc2_if(hg + (ftoff + j * fpitch), ldm, 1u << j);
#pragma unroll
                for (int j = 0; j < ST_RG; ++j)
                    e[j] = ldnc_if(hg + (feoff + (j + 1) * fpitch), em, 1u << j);
#pragma unroll
                for (int j = 0; j < ST_NR; ++j)
                    if ((ldm >> j) & 1)
                        *reinterpret_cast<double2*>(gb + (toff + j * pitch)) = v[j];
#pragma unroll
                for (int j = 0; j < ST_RG; ++j)
                    if ((em >> j) & 1)
                        gb[eoff + (j + 1) * pitch] = e[j];
            }
        }
    }
}

int launch_fetch_direct(sa_ctx* ctx, cudaStream_t stream, const Level& lv, int nbands, bool poisson, double* u, double* g,
    const HostBands& src)
{
    if (lv.n_tiles == 0 || nbands == 0)
        return SA_OK;
    if (nbands > HOST_BANDS_MAX)
        return fail(ctx, SA_BAD_ARGUMENT, "direct fetch: too many bands in one window");
    const unsigned grid = (unsigned)io_ctas(false);
    if (poisson)
        k_fetch_direct<true><<<grid, IO_THREADS, 0, stream>>>(lv, nbands, u, g, src);
    else
        k_fetch_direct<false><<<grid, IO_THREADS, 0, stream>>>(lv, nbands, u, g, src);
    ctx->launches += 1;
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// k_scatter_direct: the way out of the direct mode.  The unknown pixels -- and nothing else -- are stored from the image
// plane straight into the caller's page-locked arrays: 16 bytes where both cells of a pair are unknowns, 8 bytes where
// one is.  Known pixels never cross PCIe in either direction (laplace.cpp:117-119 / poisson.cpp:273-283 only write the
// invalid pixels too).
// The stores go out in raster order of the caller's array: the warps of the grid walk the image row by row, each taking K
// consecutive 64-column segments (K x 512 bytes) of one row at a time, a lane per aligned column pair; the K loads are in
// flight together.  At any moment the whole grid writes into a few neighbouring rows, i.e. a few pages of host memory.
// ---------------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(IO_THREADS, 1) k_scatter_direct(Level lv, int nbands, const double* __restrict__ u, HostBands dst)
{
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int pairs_x = (lv.tiles_x + 1) >> 1, groups_x = (pairs_x + K - 1) / K;
    const int pitch = (int)lv.pitch, hpitch = (int)dst.pitch;
    const int n_groups = (int)dst.rows * groups_x;
    const int sh = (2 * lane) & 31;
    for (int g = warp; g < n_groups; g += nwarps) {
        const int r = g / groups_x, tp0 = (g - r * groups_x) * K;
        // the row's 32-column words of the tiles of the K segments: lanes 0..15 hold the left tile's, 16..31 the right one's
        const uint32_t* w = lv.tbits + ((size_t)((r >> 5) + 1) * lv.tb_stride + (2 * tp0 + (lane >> 4) + 1)) * 32 + (r & 31);
        unsigned bits[K];
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            bits[k] = tp0 + k < pairs_x ? (__ldg(w + (size_t)k * 64) >> sh) & 3u : 0u;
            any |= bits[k] != 0;
        }
        if (!__any_sync(0xffffffffu, any))
            continue;
        const int o = r * pitch + tp0 * (2 * TILE_W) + 2 * lane, ho = r * hpitch + tp0 * (2 * TILE_W) + 2 * lane;
        for (int band = 0; band < nbands; ++band) {
            const double* ub = u + (int64_t)band * lv.plane + o;
            double* hb = dst.f[band] + ho;
            double2 v[K];
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (bits[k])
                    v[k] = *reinterpret_cast<const double2*>(ub + k * (2 * TILE_W));
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double* h = hb + k * (2 * TILE_W);
                if (bits[k] == 3u)
                    *reinterpret_cast<double2*>(h) = v[k];
                else if (bits[k] == 1u)
                    h[0] = v[k].x;
                else if (bits[k] == 2u)
                    h[1] = v[k].y;
            }
        }
    }
}

int launch_scatter_direct(sa_ctx* ctx, cudaStream_t stream, const Level& lv, int nbands, const double* u, const HostBands& dst)
{
    if (lv.n_tiles == 0 || nbands == 0)
        return SA_OK;
    if (nbands > HOST_BANDS_MAX)
        return fail(ctx, SA_BAD_ARGUMENT, "direct scatter: too many bands in one window");
    k_scatter_direct<8><<<(unsigned)io_ctas(true), IO_THREADS, 0, stream>>>(lv, nbands, u, dst);
    ctx->launches += 1;
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

int launch_direction2(sa_ctx* ctx, const Level& lv, int nbands, bool jacobi, const void* zin, bool z_is_float,
    const void* p_old, void* p_new, bool p_is_float, BandScalars* scal, int k)
{
    if (lv.n_tiles == 0)
        return SA_OK;
    const unsigned grid = strip_grid(ctx, lv, z_is_float ? dir_ctas<float>() : dir_ctas<double>());
#define SA_DIR(J, F, ZT, PT)                                                                                      \
    SA_LAUNCH(ctx, (k_direction2<J, F, ZT, PT>), grid, ST_THREADS, 0, lv, nbands, (const ZT*)zin, (const PT*)p_old, \
        (PT*)p_new, scal, k)
    const bool fixed = lv.fixed_diag != 0;
    if (jacobi) {
        if (p_is_float)
            return fail(ctx, SA_BAD_ARGUMENT, "direction: the Jacobi path keeps its search direction in double");
        if (fixed)
            SA_DIR(true, true, double, double);
        else
            SA_DIR(true, false, double, double);
    } else if (z_is_float) {
        if (p_is_float) {
            if (fixed)
                SA_DIR(false, true, float, float);
            else
                SA_DIR(false, false, float, float);
        } else {
            if (fixed)
                SA_DIR(false, true, float, double);
            else
                SA_DIR(false, false, float, double);
        }
    } else {
        if (p_is_float)
            return fail(ctx, SA_BAD_ARGUMENT, "direction: a double z goes with a double search direction");
        if (fixed)
            SA_DIR(false, true, double, double);
        else
            SA_DIR(false, false, double, double);
    }
#undef SA_DIR
    return SA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// k_flush_x: the step a band's last pass left behind (k_update2, XM = 1): x += alpha p at the unknowns of the bands whose
// pend flag is up.  Runs once after the loop; costs nothing for bands whose last pass was an XM = 2 pass.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS) k_flush_x(Level lv, int nbands, double* __restrict__ u, const float* __restrict__ p0,
    const float* __restrict__ p1, const BandScalars* __restrict__ scal)
{
    const int cx = threadIdx.x & 15, row0 = (threadIdx.x >> 4) * ST_RG;
    const int pitch = (int)lv.pitch;
    for (int band = 0; band < nbands; ++band) {
        const BandScalars& sc = scal[band];
        if (!sc.pend)
            continue;
        const int buf = sc.pend_buf;
        const double alpha = sc.alpha_hist[(buf + 1) & 1];  // pass k: slot k & 1, buffer (k + 1) & 1
        const float* pband = (buf ? p1 : p0) + (int64_t)band * lv.plane;
        double* uband = u + (int64_t)band * lv.plane;
        for (int i = blockIdx.x; i < lv.n_tiles; i += gridDim.x) {
            const TileBits tb = load_tile_bits(lv, lv.tile_yx[i], cx, row0);
            const unsigned mL = tb.mL(), mR = tb.mR(), any = tb.any() & ST_OWN;
            if (!any)
                continue;
            const int o = tb.origin(pitch) + (row0 - 1) * pitch + 2 * cx;
            double2 xv[ST_RG], pv[ST_RG];
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j) {
                xv[j - 1] = ld2_if(uband + (o + j * pitch), any, 1u << j);
                pv[j - 1] = ldnc2_if(pband + (o + j * pitch), any, 1u << j);
            }
#pragma unroll
            for (int j = 1; j <= ST_RG; ++j)
                if ((any >> j) & 1) {
                    double2 xn = xv[j - 1];
                    if ((mL >> j) & 1)
                        xn.x += alpha * pv[j - 1].x;
                    if ((mR >> j) & 1)
                        xn.y += alpha * pv[j - 1].y;
                    *reinterpret_cast<double2*>(uband + (o + j * pitch)) = xn;
                }
        }
    }
}

int launch_flush_x(sa_ctx* ctx, const Level& lv, int nbands, double* u, const float* p0, const float* p1, const BandScalars* scal)
{
    if (lv.n_tiles == 0 || nbands == 0)
        return SA_OK;
    SA_LAUNCH(ctx, k_flush_x, strip_grid(ctx, lv, 8), ST_THREADS, 0, lv, nbands, u, p0, p1, scal);
    return SA_OK;
}

int launch_update2(sa_ctx* ctx, const Level& lv, int nbands, bool jacobi, double* u, const void* p, bool p_is_float, double* r,
    float* rf, BandScalars* scal, int k, int xm, const void* p_prev)
{
    if (lv.n_tiles == 0)
        return SA_OK;
    const unsigned grid = strip_grid(ctx, lv, ST_UPD_CTAS);
    if (xm != 0) {  // the product path only: float direction, float copy of the residual
        if (jacobi || !p_is_float || !rf || (xm == 2 && !p_prev))
            return fail(ctx, SA_BAD_ARGUMENT, "update: deferred x is for the float-direction path");
#define SA_UPDX(F, X)                                                                                                   \
    SA_LAUNCH(ctx, (k_update2<false, F, true, float, X>), grid, ST_THREADS, 0, lv, nbands, u, (const float*)p, r, rf, scal, k, \
        (const float*)p_prev)
        const bool fx = lv.fixed_diag != 0;
        if (xm == 1) {
            if (fx)
                SA_UPDX(true, 1);
            else
                SA_UPDX(false, 1);
        } else {
            if (fx)
                SA_UPDX(true, 2);
            else
                SA_UPDX(false, 2);
        }
#undef SA_UPDX
        return SA_OK;
    }
#define SA_UPD(J, F, R, PT) \
    SA_LAUNCH(ctx, (k_update2<J, F, R, PT>), grid, ST_THREADS, 0, lv, nbands, u, (const PT*)p, r, rf, scal, k, (const PT*)nullptr)
    const bool fixed = lv.fixed_diag != 0;
    if (jacobi) {
        if (p_is_float)
            return fail(ctx, SA_BAD_ARGUMENT, "update: the Jacobi path keeps its search direction in double");
        if (fixed)
            SA_UPD(true, true, false, double);
        else
            SA_UPD(true, false, false, double);
    } else if (p_is_float) {
        if (fixed) {
            if (rf)
                SA_UPD(false, true, true, float);
            else
                SA_UPD(false, true, false, float);
        } else {
            if (rf)
                SA_UPD(false, false, true, float);
            else
                SA_UPD(false, false, false, float);
        }
    } else {
        if (fixed) {
            if (rf)
                SA_UPD(false, true, true, double);
            else
                SA_UPD(false, true, false, double);
        } else {
            if (rf)
                SA_UPD(false, false, true, double);
            else
                SA_UPD(false, false, false, double);
        }
    }
#undef SA_UPD
    return SA_OK;
}

}  // namespace satfill
