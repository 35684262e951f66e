// What bounds the way home of the unknown pixels?  (tools/probe: measurement only, not part of the library)
//   1. copy-engine bandwidth device -> pinned host, host -> device
//   2. a kernel storing straight into mapped pinned host memory: contiguous, and in runs with gaps (the shape of a
//      scatter of unknown pixels into an image), 8 and 16 bytes per lane, few and many CTAs
//   3. host threads scattering a contiguous buffer into runs of an image (memcpy, 1 .. 16 threads)
//   4. 1 and 3 at the same time
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe/pcie_probe tools/probe/pcie_probe.cu -lpthread
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <typename T>
__global__ void k_store(T* __restrict__ dst, const T* __restrict__ src, size_t n, int run, int gap)
{
    // element i of src -> run/gap pattern in dst: block of `run` elements, then `gap` elements skipped (gap 0: contiguous)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t o = gap ? (i / run) * (size_t)(run + gap) + i % run : i;
        dst[o] = src[i];
    }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void host_scatter(const double* src, double* dst, size_t n, int run, int gap, int threads)
{
    std::vector<std::thread> th;
    const size_t nruns = n / run;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=] {
            const size_t r0 = nruns * t / threads, r1 = nruns * (t + 1) / threads;
            for (size_t r = r0; r < r1; ++r)
                std::memcpy(dst + r * (size_t)(run + gap), src + r * (size_t)run, (size_t)run * sizeof(double));
        });
    for (auto& x : th)
        x.join();
}

int main()
{
    const size_t n = (size_t)1 << 27;  // 1 GiB of doubles
    const int run = 50, gap = 50;
    double *d = nullptr, *h = nullptr, *himg = nullptr;
    CK(cudaMalloc(&d, n * 8));
    CK(cudaMemset(d, 1, n * 8));
    CK(cudaHostAlloc(&h, n * 8, cudaHostAllocMapped));
    CK(cudaHostAlloc(&himg, 2 * n * 8 + 4096, cudaHostAllocMapped));
    std::memset(h, 0, n * 8);
    std::memset(himg, 0, 2 * n * 8);
    cudaStream_t s, s2;
    CK(cudaStreamCreate(&s));
    CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms;
    std::printf("host threads available: %u\n", std::thread::hardware_concurrency());
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, s));
        CK(cudaMemcpyAsync(h, d, n * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf("copy engine D2H 1 GiB: %.1f GB/s\n", n * 8 / ms * 1e-6);
        CK(cudaEventRecord(e0, s));
        CK(cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf("copy engine H2D 1 GiB: %.1f GB/s\n", n * 8 / ms * 1e-6);
    }
    double* dh = nullptr;
    double* dimg = nullptr;
    CK(cudaHostGetDevicePointer(&dh, h, 0));
    CK(cudaHostGetDevicePointer(&dimg, himg, 0));
    for (int g : { 0, gap })
        for (int width : { 8, 16 })
            for (int ctas : { 8, 32, 148, 592 }) {
                double* dst = g ? dimg : dh;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0, s));
                    if (width == 8)
                        k_store<double><<<ctas, 256, 0, s>>>(dst, d, n / 4, run, g);
                    else
                        k_store<double2><<<ctas, 256, 0, s>>>((double2*)dst, (const double2*)d, n / 8, run / 2, g / 2);
                    CK(cudaEventRecord(e1, s));
                    CK(cudaStreamSynchronize(s));
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                }
                std::printf("kernel store to host, %s, %2d B/lane, %3d CTAs: %.1f GB/s\n", g ? "runs of 400 B + gaps" : "contiguous         ", width, ctas,
                    n / 4 * 8 / ms * 1e-6);
            }
    for (int threads : { 1, 2, 4, 8, 12, 16 }) {
        double best = 1e9;
        for (int rep = 0; rep < 2; ++rep) {
            const double t0 = now();
            host_scatter(h, himg, n, run, gap, threads);
            best = std::min(best, now() - t0);
        }
        std::printf("host scatter (runs of 400 B), %2d threads: %.1f GB/s of pixels\n", threads, n * 8 / best * 1e-9);
    }
    for (int threads : { 4, 8, 12 }) {
        CK(cudaEventRecord(e0, s));
        for (int k = 0; k < 3; ++k) {
            CK(cudaMemcpyAsync(d + n / 2, h + n / 2, n * 4, cudaMemcpyHostToDevice, s2));
            CK(cudaMemcpyAsync(h, d, n * 4, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaEventRecord(e1, s));
        const double t0 = now();
        host_scatter(h, himg, n, run, gap, threads);
        const double t = now() - t0;
        CK(cudaStreamSynchronize(s));
        CK(cudaStreamSynchronize(s2));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::printf("both at once, %2d threads: host scatter %.1f GB/s, copy engine D2H %.1f GB/s (1.5 GiB, H2D running beside it)\n", threads,
            n * 8 / t * 1e-9, 3.0 * n * 4 / ms * 1e-6);
    }
    return 0;
}
