// Integer kernels: layout transposition, mask normalisation, unknown-set construction, block-sparse tile list and
// raster-order numbering.  Everything here is exact integer work and must be bit-identical to the oracle
// (oracle/satfill_oracle.c: so_mask_scan, so_unknown_numbering), which restates laplace.cpp:33-52 and
// poisson.cpp:162-177.  HBM-bound byte work: one coalesced read of the mask, one coalesced write of the table.
#include "common.cuh"

#include <algorithm>

namespace satfill {

// ---------------------------------------------------------------------------------------------------------------
// 32 x 32 shared-memory tile transpose: dst(c, r) = src(r, c).  Column-major sources (the reference's MatX,
// utils/types.h:31) are uploaded as their row-major transpose and flipped once on the device; integer tables go the
// other way on download.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_transpose(const T* __restrict__ src, int64_t src_rows, int64_t src_cols,
    int64_t src_pitch, T* __restrict__ dst, int64_t dst_pitch)
{
    __shared__ T tile[32][33];
    int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < src_rows && c < src_cols)
            tile[j][threadIdx.x] = src[r * src_pitch + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        int64_t c = c0 + j, r = r0 + threadIdx.x;  // dst row = src col
        if (r < src_rows && c < src_cols)
            dst[c * dst_pitch + r] = tile[threadIdx.x][j];
    }
}

template <typename T>
static int transpose_any(sa_ctx* ctx, const T* src, int64_t src_rows, int64_t src_cols, int64_t src_pitch, T* dst,
    int64_t dst_pitch)
{
    if (src_rows == 0 || src_cols == 0)
        return SA_OK;
    dim3 grid((unsigned)((src_cols + 31) / 32), (unsigned)((src_rows + 31) / 32)), block(32, 8);
    SA_LAUNCH(ctx, k_transpose<T>, grid, block, 0, src, src_rows, src_cols, src_pitch, dst, dst_pitch);
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

int transpose_u8(sa_ctx* ctx, const uint8_t* s, int64_t r, int64_t c, int64_t sp, uint8_t* d, int64_t dp)
{
    return transpose_any<uint8_t>(ctx, s, r, c, sp, d, dp);
}
int transpose_i32(sa_ctx* ctx, const int32_t* s, int64_t r, int64_t c, int64_t sp, int32_t* d, int64_t dp)
{
    return transpose_any<int32_t>(ctx, s, r, c, sp, d, dp);
}

// ---------------------------------------------------------------------------------------------------------------
// Unknown set + tile list.  One CTA per 32 x 32 tile.  unknown(p) = invalid(p) and, for Laplace, p not on the image
// border (laplace.cpp:23-29, 98-100: border cells are identity rows even when invalid).  The raw mask is normalised
// to 0/1 in place on the way.  Each CTA records whether its tile holds an unknown; a single-CTA ballot scan then
// lists the active tiles in raster order (reproducible work order, neighbouring CTAs touch neighbouring memory).
// ---------------------------------------------------------------------------------------------------------------
// A row-decomposed scene (dist.cu) launches it on its own tile rows plus one either side (first_tile = the first tile of
// that window); only tiles of rows [own_lo, own_hi) are flagged, i.e. listed.
__global__ void __launch_bounds__(CG_THREADS) k_build_unknown_set(uint8_t* __restrict__ mask, uint8_t* __restrict__ umask,
    int64_t rows, int64_t cols, int64_t pitch, int tiles_x, int laplace, int32_t* __restrict__ tile_flags,
    unsigned long long* __restrict__ count64, uint32_t* __restrict__ tbits, uint32_t* __restrict__ tbitsT, int first_tile, int own_lo,
    int own_hi)
{
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
        uint8_t m = 0, um = 0;
        if (r < rows && c < cols) {
            m = mask[r * pitch + c] != 0;
            bool border = r == 0 || r == rows - 1 || c == 0 || c == cols - 1;
            um = m && !(laplace && border);
        }
        // padding stays 0 in both tables
        mask[r * pitch + c] = m;
        umask[r * pitch + c] = um;
        cnt += um;
        unsigned word = __ballot_sync(0xffffffffu, um);
        if (threadIdx.x == 0)
            tbits[((size_t)(ty + 1) * (tiles_x + 2) + tx + 1) * 32 + threadIdx.y + j * CG_BLOCK_Y] = word;
        colbits |= (unsigned)um << (threadIdx.y + j * CG_BLOCK_Y);
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

// unknowns of the whole mask (a row-decomposed scene indexes only its own rows, but every rank needs the size of the system)
__global__ void __launch_bounds__(256) k_count_unknowns(const uint8_t* __restrict__ mask, int64_t rows, int64_t cols, int64_t pitch,
    int laplace, unsigned long long* __restrict__ count64)
{
    unsigned long long n = 0;
    const int64_t r_lo = laplace ? 1 : 0, r_hi = laplace ? rows - 1 : rows, c_lo = laplace ? 1 : 0, c_hi = laplace ? cols - 1 : cols;
    for (int64_t r = r_lo + blockIdx.x; r < r_hi; r += gridDim.x)
        for (int64_t c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x)
            n += mask[r * pitch + c] != 0;
    for (int o = 16; o; o >>= 1)
        n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n)
        atomicAdd(count64, n);
}

__global__ void __launch_bounds__(1024) k_compact_flags(const int32_t* __restrict__ flags, int n_tiles, int tiles_x,
    int32_t* __restrict__ tile_list, int32_t* __restrict__ tile_yx, int32_t* __restrict__ n_active, int first_tile)
{
    // single CTA, raster order: running offset + block-wide ballot scan
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0)
        base = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // tiles [first_tile, first_tile + n_tiles) of the grid
    for (int start = first_tile; start < first_tile + n_tiles; start += 1024) {
        int i = start + threadIdx.x;
        int f = (i < first_tile + n_tiles) ? flags[i] : 0;
        unsigned b = __ballot_sync(0xffffffffu, f);
        int pre = __popc(b & ((1u << lane) - 1));
        if (lane == 0)
            warp_tot[warp] = __popc(b);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w)
            woff += warp_tot[w];
        if (f) {
            tile_list[base + woff + pre] = i;
            tile_yx[base + woff + pre] = ((i / tiles_x) << 16) | (i % tiles_x);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w)
                t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {  // n_active[0..2] = count, first and last active tile (raster order)
        n_active[0] = base;
        n_active[1] = base ? tile_list[0] : 0;
        n_active[2] = base ? tile_list[base - 1] : 0;
    }
}

// Shared by the fine level (index_scene) and the multigrid coarse levels: flags -> raster-ordered list + count.
int compact_tile_flags(sa_ctx* ctx, const int32_t* flags, int n_tiles, int tiles_x, int32_t* tile_list, int32_t* tile_yx,
    int32_t* d_n_active, int first_tile)
{
    SA_LAUNCH(ctx, k_compact_flags, 1, 1024, 0, flags, n_tiles, tiles_x, tile_list, tile_yx, d_n_active, first_tile);
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

int index_scene(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    int n_tiles = s->tiles_x * s->tiles_y;
    int32_t* flags = s->tile_list + n_tiles;  // tile_list is allocated with 3 * n_tiles entries
    SA_CUDA(ctx, cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(int32_t), ctx->stream));
    SA_CUDA(ctx, cudaMemsetAsync(s->d_count64, 0, sizeof(unsigned long long), ctx->stream));
    dim3 block(CG_BLOCK_X, CG_BLOCK_Y);
    // a row-decomposed scene: its own tile rows (+ one either side for the neighbourhoods of its edge tiles)
    int own_lo = 0, own_hi = s->tiles_y, win_lo = 0, win_hi = s->tiles_y;
    if (s->dist_windowed && !s->dl.empty()) {
        own_lo = (int)(s->dl[0].row_lo / TILE_H);
        own_hi = (int)(s->dl[0].row_hi / TILE_H);
        win_lo = std::max(own_lo - 1, 0);
        win_hi = std::min(own_hi + 1, s->tiles_y);
        if (win_hi < win_lo)
            win_hi = win_lo;
    }
    const int first = win_lo * s->tiles_x, count = (win_hi - win_lo) * s->tiles_x;
    if (count > 0)
        SA_LAUNCH(ctx, k_build_unknown_set, count, block, 0, s->mask0(s->mask), s->mask0(s->umask), s->rows, s->cols, s->pitch,
            s->tiles_x, s->problem == SA_LAPLACE ? 1 : 0, flags, s->d_count64, s->tbits, s->tbits + s->tb_words, first, own_lo, own_hi);
    SA_CUDA(ctx, cudaGetLastError());
    SA_TRY(compact_tile_flags(ctx, flags, count, s->tiles_x, s->tile_list, s->tile_list + 2 * n_tiles, s->d_counters, first));
    struct readback {
        int32_t counters[4];
        unsigned long long n;
    }* h = (readback*)ctx->pinned;
    SA_CUDA(ctx, cudaMemcpyAsync(h->counters, s->d_counters, sizeof(h->counters), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaMemcpyAsync(&h->n, s->d_count64, sizeof(h->n), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s->n_active_tiles = h->counters[0];
    s->n_unknowns = (int64_t)h->n;
    if (s->dist_windowed)  // the size of the whole system, not of this rank's rows
        SA_TRY(count_unknowns(s, &s->n_unknowns));
    s->indexed = true;
    return SA_OK;
}

int count_unknowns(sa_scene* s, int64_t* out)
{
    sa_ctx* ctx = s->ctx;
    SA_CUDA(ctx, cudaMemsetAsync(s->d_count64, 0, sizeof(unsigned long long), ctx->stream));
    SA_LAUNCH(ctx, k_count_unknowns, 4 * ctx->sm_count, 256, 0, s->mask0(s->mask), s->rows, s->cols, s->pitch,
        s->problem == SA_LAPLACE ? 1 : 0, s->d_count64);
    unsigned long long* h = (unsigned long long*)ctx->pinned;
    SA_CUDA(ctx, cudaMemcpyAsync(h, s->d_count64, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = (int64_t)*h;
    return SA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Raster-order numbering (poisson.cpp:162-177) and invalid-pixel list + bounding box (laplace.cpp:33-52).
// numbering(r, c) = exclusive prefix sum of the row-major mask.  Three phases: per-row counts, scan of the row
// counts (one CTA), per-row ballot scan that writes the table / the list.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_counts(const uint8_t* __restrict__ mask, int64_t cols, int64_t pitch,
    unsigned long long* __restrict__ row_count, int* __restrict__ bbox /* minr maxr minc maxc */)
{
    __shared__ int s_cnt[8], s_minc[8], s_maxc[8];
    int64_t r = blockIdx.x;
    int cnt = 0, minc = INT_MAX, maxc = -1;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
        if (mask[r * pitch + c]) {
            ++cnt;
            minc = min(minc, (int)c);
            maxc = max(maxc, (int)c);
        }
    }
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        minc = min(minc, __shfl_xor_sync(0xffffffffu, minc, o));
        maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_cnt[warp] = cnt;
        s_minc[warp] = minc;
        s_maxc[warp] = maxc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            cnt += s_cnt[w];
            minc = min(minc, s_minc[w]);
            maxc = max(maxc, s_maxc[w]);
        }
        row_count[r] = (unsigned long long)cnt;
        if (cnt) {
            atomicMin(&bbox[0], (int)r);
            atomicMax(&bbox[1], (int)r);
            atomicMin(&bbox[2], minc);
            atomicMax(&bbox[3], maxc);
        }
    }
}

__global__ void __launch_bounds__(1024) k_scan_rows(unsigned long long* __restrict__ row_count, int64_t rows,
    unsigned long long* __restrict__ total)
{
    // exclusive scan in place, single CTA
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long base;
    if (threadIdx.x == 0)
        base = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t start = 0; start < rows; start += 1024) {
        int64_t i = start + threadIdx.x;
        unsigned long long v = (i < rows) ? row_count[i] : 0ull, incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            warp_tot[warp] = incl;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < warp; ++w)
            woff += warp_tot[w];
        if (i < rows)
            row_count[i] = base + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < 32; ++w)
                t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *total = base;
}

__global__ void __launch_bounds__(256) k_row_number(const uint8_t* __restrict__ mask, int64_t cols, int64_t pitch,
    const unsigned long long* __restrict__ row_offset, int32_t* __restrict__ numbering /* dense rows x cols or null */,
    int64_t* __restrict__ pixels /* (row, col) pairs or null */, int64_t capacity)
{
    __shared__ int warp_tot[8];
    __shared__ unsigned long long base;
    int64_t r = blockIdx.x;
    if (threadIdx.x == 0)
        base = row_offset[r];
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t start = 0; start < cols; start += blockDim.x) {
        int64_t c = start + threadIdx.x;
        int f = (c < cols) ? (mask[r * pitch + c] != 0) : 0;
        unsigned b = __ballot_sync(0xffffffffu, f);
        int pre = __popc(b & ((1u << lane) - 1));
        if (lane == 0)
            warp_tot[warp] = __popc(b);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w)
            woff += warp_tot[w];
        unsigned long long k = base + (unsigned long long)(woff + pre);
        if (c < cols && numbering)
            numbering[r * cols + c] = f ? (int32_t)k : -1;
        if (f && pixels && (int64_t)k < capacity) {
            pixels[2 * k] = r;
            pixels[2 * k + 1] = c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w)
                t += warp_tot[w];
            base += (unsigned long long)t;
        }
        __syncthreads();
    }
}

int device_scan_u64(sa_ctx* ctx, unsigned long long* data, int64_t n, unsigned long long* total)
{
    SA_LAUNCH(ctx, k_scan_rows, 1, 1024, 0, data, n, total);
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

__global__ void k_init_bbox(int* bbox, int rows, int cols)
{
    bbox[0] = rows;
    bbox[1] = -1;
    bbox[2] = cols;
    bbox[3] = -1;
}

int device_numbering(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t pitch, int32_t* numbering,
    int64_t* out_pixels, int64_t capacity, int64_t* out_count, int64_t bbox[4])
{
    unsigned long long* row_count = nullptr;
    int* d_bbox = nullptr;
    if (rows == 0 || cols == 0) {
        if (out_count) *out_count = 0;
        if (bbox) { bbox[0] = rows; bbox[1] = -1; bbox[2] = cols; bbox[3] = -1; }
        return SA_OK;
    }
    SA_CUDA(ctx, cudaMallocAsync(&row_count, (size_t)(rows + 1) * sizeof(unsigned long long), ctx->stream));
    SA_CUDA(ctx, cudaMallocAsync(&d_bbox, 4 * sizeof(int), ctx->stream));
    SA_LAUNCH(ctx, k_init_bbox, 1, 1, 0, d_bbox, (int)rows, (int)cols);
    SA_LAUNCH(ctx, k_row_counts, (unsigned)rows, 256, 0, mask, cols, pitch, row_count, d_bbox);
    SA_TRY(device_scan_u64(ctx, row_count, rows, row_count + rows));
    if (numbering || out_pixels)
        SA_LAUNCH(ctx, k_row_number, (unsigned)rows, 256, 0, mask, cols, pitch, row_count, numbering, out_pixels,
            capacity);
    SA_CUDA(ctx, cudaGetLastError());
    struct readback {
        unsigned long long total;
        int bbox[4];
    }* h = (readback*)ctx->pinned;
    SA_CUDA(ctx, cudaMemcpyAsync(&h->total, row_count + rows, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaMemcpyAsync(h->bbox, d_bbox, sizeof(h->bbox), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaFreeAsync(row_count, ctx->stream));
    SA_CUDA(ctx, cudaFreeAsync(d_bbox, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_count) *out_count = (int64_t)h->total;
    if (bbox)
        for (int i = 0; i < 4; ++i)
            bbox[i] = h->bbox[i];
    return SA_OK;
}

}  // namespace satfill
