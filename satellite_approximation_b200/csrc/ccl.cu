// Connected-component labelling of the invalid mask.  Replaces approx::find_connected_components (laplace.h:11-20),
// which the reference declares and tests (tests/approximation.h:55-75) but never defines; contract (SURVEY.md 8a A3,
// restated by oracle/satfill_oracle.c so_label_components): 4-connectivity (the stencil's coupling, utils.h:38-44),
// background 0, labels 1..K numbered by the first pixel of each component in row-major raster order.
//
// Label-equivalence union-find, all integer, bit-exact by construction:
//   1. k_ccl_init     every invalid pixel points at the first pixel of its horizontal run (warp ballot per 32 columns,
//                     stitched across 32-column segments by the merge step)
//   2. k_ccl_merge    union(pixel, upper neighbour) at run starts and wherever the upper run changes, union across
//                     segment seams; the union always links the larger root to the smaller one (atomicMin), so
//                     the root of a component is its first pixel in raster order
//   3. k_ccl_flatten  every pixel -> its root
//   4. roots are ranked in raster order with the same row-count / scan / row-number passes as the unknown numbering
//      (mask_index.cu) and every pixel takes rank(root) + 1.
// HBM-bound byte/integer work: 1 B read + 4 B written per pixel per pass.
#include "common.cuh"

namespace satfill {

__device__ __forceinline__ int ccl_find(const int32_t* L, int i)
{
    // L2-coherent loads: the table is being rewritten by atomicMin in other CTAs while it is walked
    int p = __ldcg(L + i);
    while (p != i) {
        i = p;
        p = __ldcg(L + i);
    }
    return i;
}

__device__ __forceinline__ void ccl_union(int32_t* L, int a, int b)
{
    bool done;
    do {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// grid: (ceil(cols / 32), ceil(rows / 8)), block (32, 8): one warp per 32-column segment of a row.
__global__ void __launch_bounds__(256) k_ccl_init(const uint8_t* __restrict__ mask, int64_t rows, int64_t cols,
    int64_t pitch, int32_t* __restrict__ L)
{
    int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y, c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (r >= rows)
        return;
    bool m = c < cols && mask[r * pitch + c] != 0;
    unsigned bits = __ballot_sync(0xffffffffu, m);
    if (c >= cols)
        return;
    int lane = threadIdx.x;
    int32_t v = -1;
    if (m) {
        // start of the run inside this segment: position after the highest zero bit below `lane`
        unsigned below = ~bits & ((1u << lane) - 1);
        int start = below ? 32 - __clz(below) : 0;
        v = (int32_t)(r * cols + (c - lane + start));
    }
    L[r * cols + c] = v;
}

__global__ void __launch_bounds__(256) k_ccl_merge(const uint8_t* __restrict__ mask, int64_t rows, int64_t cols,
    int64_t pitch, int32_t* __restrict__ L)
{
    int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y, c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (r >= rows || c >= cols)
        return;
    if (!mask[r * pitch + c])
        return;
    int i = (int)(r * cols + c);
    bool left = c > 0 && mask[r * pitch + c - 1] != 0;
    // seam between two 32-column segments of the same run
    if (threadIdx.x == 0 && left)
        ccl_union(L, i, i - 1);
    if (r > 0 && mask[(r - 1) * pitch + c] != 0) {
        // one union per contact between runs is enough: do it where either run starts
        bool up_left = c > 0 && mask[(r - 1) * pitch + c - 1] != 0;
        if (!left || !up_left)
            ccl_union(L, i, i - (int)cols);
    }
}

__global__ void __launch_bounds__(256) k_ccl_flatten(int64_t total, int32_t* __restrict__ L)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total)
        return;
    int v = L[i];
    if (v >= 0)
        L[i] = ccl_find(L, (int)i);
}

// roots per row (a root is a pixel that points at itself)
__global__ void __launch_bounds__(256) k_ccl_root_counts(const int32_t* __restrict__ L, int64_t cols,
    unsigned long long* __restrict__ row_count)
{
    __shared__ int s_cnt[8];
    int64_t r = blockIdx.x;
    int cnt = 0;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x)
        cnt += L[r * cols + c] == (int32_t)(r * cols + c);
    for (int o = 16; o; o >>= 1)
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0)
        s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            cnt += s_cnt[w];
        row_count[r] = (unsigned long long)cnt;
    }
}

// labels[root] = 1 + number of roots before it in raster order
__global__ void __launch_bounds__(256) k_ccl_root_rank(const int32_t* __restrict__ L, int64_t cols,
    const unsigned long long* __restrict__ row_offset, int32_t* __restrict__ labels)
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
        int f = (c < cols) ? (L[r * cols + c] == (int32_t)(r * cols + c)) : 0;
        unsigned b = __ballot_sync(0xffffffffu, f);
        int pre = __popc(b & ((1u << lane) - 1));
        if (lane == 0)
            warp_tot[warp] = __popc(b);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w)
            woff += warp_tot[w];
        if (f)
            labels[r * cols + c] = (int32_t)(base + (unsigned long long)(woff + pre)) + 1;
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

__global__ void __launch_bounds__(256) k_ccl_relabel(int64_t total, const int32_t* __restrict__ L,
    int32_t* __restrict__ labels)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total)
        return;
    int root = L[i];
    if (root < 0)
        labels[i] = 0;
    else if (root != (int)i)
        labels[i] = labels[root];  // root entries were written by k_ccl_root_rank and are never rewritten here
}

int device_label_components(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t pitch,
    int32_t* labels, int32_t* out_num_labels)
{
    if (out_num_labels)
        *out_num_labels = 0;
    int64_t total = rows * cols;
    if (total == 0)
        return SA_OK;
    if (total > (int64_t)INT32_MAX)
        return fail(ctx, SA_BAD_ARGUMENT, "label_components: rows * cols must fit a 32-bit label table");
    int32_t* L = nullptr;
    unsigned long long* row_count = nullptr;
    SA_CUDA(ctx, cudaMallocAsync(&L, (size_t)total * sizeof(int32_t), ctx->stream));
    SA_CUDA(ctx, cudaMallocAsync(&row_count, (size_t)(rows + 1) * sizeof(unsigned long long), ctx->stream));
    dim3 block(32, 8), grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 7) / 8));
    unsigned lin = (unsigned)((total + 255) / 256);
    SA_LAUNCH(ctx, k_ccl_init, grid, block, 0, mask, rows, cols, pitch, L);
    SA_LAUNCH(ctx, k_ccl_merge, grid, block, 0, mask, rows, cols, pitch, L);
    SA_LAUNCH(ctx, k_ccl_flatten, lin, 256, 0, total, L);
    SA_LAUNCH(ctx, k_ccl_root_counts, (unsigned)rows, 256, 0, L, cols, row_count);
    SA_TRY(device_scan_u64(ctx, row_count, rows, row_count + rows));
    SA_LAUNCH(ctx, k_ccl_root_rank, (unsigned)rows, 256, 0, L, cols, row_count, labels);
    SA_LAUNCH(ctx, k_ccl_relabel, lin, 256, 0, total, L, labels);
    SA_CUDA(ctx, cudaGetLastError());
    unsigned long long* h = (unsigned long long*)ctx->pinned;
    SA_CUDA(ctx, cudaMemcpyAsync(h, row_count + rows, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaFreeAsync(L, ctx->stream));
    SA_CUDA(ctx, cudaFreeAsync(row_count, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_num_labels)
        *out_num_labels = (int32_t)*h;
    return SA_OK;
}

}  // namespace satfill
