// h2d_probe.cu -- host-link diagnostic for the GPU box (not part of the product, not a pytest).
// Question (missing #2): the float32 (3,600,1000) image upload (7.2 MB) ran at ~1.5 GB/s through
// torch's pinned copy.  Is that the link, the copy engine, the allocation kind, or the calling code?
// Measures, with CUDA events on the copying stream (and wall clock around a sync for the call overhead):
//   * cudaMemcpyAsync H2D / D2H from cudaHostAlloc(Default | WriteCombined | Mapped) and cudaHostRegister memory,
//     sizes 64 KB .. 256 MB
//   * the 7.2 MB copy split in 1 MB chunks, and split over 2 / 4 streams
//   * an SM copy kernel reading MAPPED pinned host memory directly (zero-copy, no copy engine)
//   * pageable memcpy for reference
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o h2d_probe h2d_probe.cu      Run: ./h2d_probe
#include <cuda_runtime.h>

#include <algorithm>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

__global__ void zero_copy_read(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) dst[i] = src[i];
}

// median of `reps` event-timed runs of fn() (ms)
template <class F>
static float timed(F fn, cudaStream_t st, int reps = 9) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<float> v;
    fn();
    CK(cudaStreamSynchronize(st));
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, st));
        fn();
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        v.push_back(ms);
    }
    std::sort(v.begin(), v.end());
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    return v[v.size() / 2];
}

#include <algorithm>

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, asyncEngineCount %d, canMapHostMemory %d, unifiedAddressing %d\n", prop.name, prop.asyncEngineCount,
           prop.canMapHostMemory, prop.unifiedAddressing);
    const size_t MAXB = 256u << 20;
    void* dev;
    CK(cudaMalloc(&dev, MAXB));
    cudaStream_t st, st2[4];
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& s : st2) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

    struct Kind { const char* name; unsigned flags; int reg; };
    const Kind kinds[] = {{"cudaHostAlloc(Default)", cudaHostAllocDefault, 0},
                          {"cudaHostAlloc(WriteCombined)", cudaHostAllocWriteCombined, 0},
                          {"cudaHostAlloc(Mapped|Portable)", cudaHostAllocMapped | cudaHostAllocPortable, 0},
                          {"malloc + cudaHostRegister", 0, 1}};
    const size_t sizes[] = {64u << 10, 703125, 1u << 20, 7200000, 32u << 20, 256u << 20};
    for (const Kind& k : kinds) {
        void* h = nullptr;
        if (k.reg) {
            CK(cudaSuccess);
            if (posix_memalign(&h, 4096, MAXB) != 0) return 1;
            memset(h, 1, MAXB);
            CK(cudaHostRegister(h, MAXB, cudaHostRegisterDefault));
        } else {
            CK(cudaHostAlloc(&h, MAXB, k.flags));
            memset(h, 1, MAXB);
        }
        printf("== %s\n", k.name);
        for (size_t n : sizes) {
            float h2d = timed([&] { CK(cudaMemcpyAsync(dev, h, n, cudaMemcpyHostToDevice, st)); }, st);
            float d2h = timed([&] { CK(cudaMemcpyAsync(h, dev, n, cudaMemcpyDeviceToHost, st)); }, st);
            // wall clock of call + sync (what a serial caller sees)
            double t0 = now_ms();
            for (int r = 0; r < 5; ++r) {
                CK(cudaMemcpyAsync(dev, h, n, cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));
            }
            double wall = (now_ms() - t0) / 5;
            printf("  %10zu B: H2D %8.3f ms %7.2f GB/s | D2H %8.3f ms %7.2f GB/s | H2D call+sync wall %8.3f ms\n", n, h2d,
                   n / h2d / 1e6, d2h, n / d2h / 1e6, wall);
        }
        // the 7.2 MB image: chunked, multi-stream, zero-copy
        const size_t n = 7200000;
        float chunked = timed([&] {
            for (size_t o = 0; o < n; o += (1u << 20))
                CK(cudaMemcpyAsync((char*)dev + o, (char*)h + o, std::min<size_t>(1u << 20, n - o), cudaMemcpyHostToDevice, st));
        }, st);
        printf("  7.2 MB in 1 MB chunks, one stream: %.3f ms %.2f GB/s\n", chunked, n / chunked / 1e6);
        for (int ns : {2, 4}) {
            cudaEvent_t e0, e1, ej[4];
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            for (auto& e : ej) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            std::vector<float> v;
            for (int r = 0; r < 7; ++r) {
                CK(cudaEventRecord(e0, st));
                const size_t part = (n / ns + 15) / 16 * 16;
                for (int s = 0; s < ns; ++s) {
                    CK(cudaStreamWaitEvent(st2[s], e0, 0));
                    const size_t o = s * part, len = std::min(part, n - o);
                    CK(cudaMemcpyAsync((char*)dev + o, (char*)h + o, len, cudaMemcpyHostToDevice, st2[s]));
                    CK(cudaEventRecord(ej[s], st2[s]));
                    CK(cudaStreamWaitEvent(st, ej[s], 0));
                }
                CK(cudaEventRecord(e1, st));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                v.push_back(ms);
            }
            std::sort(v.begin(), v.end());
            printf("  7.2 MB over %d streams: %.3f ms %.2f GB/s\n", ns, v[v.size() / 2], n / v[v.size() / 2] / 1e6);
        }
        void* hd = nullptr;
        if (cudaHostGetDevicePointer(&hd, h, 0) == cudaSuccess && hd != nullptr) {
            for (int blocks : {148, 592, 2368}) {
                float zc = timed([&] { zero_copy_read<<<blocks, 256, 0, st>>>((const uint4*)hd, (uint4*)dev, n / 16); }, st);
                printf("  zero-copy SM read of mapped host memory, 7.2 MB, %4d CTAs: %.3f ms %.2f GB/s\n", blocks, zc, n / zc / 1e6);
            }
            float zc = timed([&] { zero_copy_read<<<592, 256, 0, st>>>((const uint4*)hd, (uint4*)dev, (32u << 20) / 16); }, st);
            printf("  zero-copy SM read, 32 MB: %.3f ms %.2f GB/s\n", zc, (32u << 20) / zc / 1e6);
        } else {
            cudaGetLastError();
            printf("  (no device pointer for this allocation)\n");
        }
        if (k.reg) {
            CK(cudaHostUnregister(h));
            free(h);
        } else {
            CK(cudaFreeHost(h));
        }
    }
    {
        void* h = malloc(MAXB);
        memset(h, 1, MAXB);
        printf("== pageable malloc\n");
        for (size_t n : {(size_t)7200000, (size_t)(32u << 20)}) {
            double t0 = now_ms();
            for (int r = 0; r < 5; ++r) CK(cudaMemcpy(dev, h, n, cudaMemcpyHostToDevice));
            double w = (now_ms() - t0) / 5;
            printf("  %10zu B: cudaMemcpy H2D wall %.3f ms %.2f GB/s\n", n, w, n / w / 1e6);
        }
        free(h);
    }
    // periodic stalls: 400 back-to-back 7.2 MB pinned copies, wall clock per call+sync; report percentiles
    {
        void* h;
        CK(cudaHostAlloc(&h, 8u << 20, cudaHostAllocDefault));
        memset(h, 1, 8u << 20);
        std::vector<double> v;
        for (int r = 0; r < 400; ++r) {
            double t0 = now_ms();
            CK(cudaMemcpyAsync(dev, h, 7200000, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            v.push_back(now_ms() - t0);
        }
        std::vector<double> s = v;
        std::sort(s.begin(), s.end());
        printf("== 400 serial 7.2 MB pinned copies (call+sync wall): min %.3f p50 %.3f p90 %.3f p99 %.3f max %.3f ms\n", s[0], s[200],
               s[360], s[396], s[399]);
        CK(cudaFreeHost(h));
    }
    return 0;
}
