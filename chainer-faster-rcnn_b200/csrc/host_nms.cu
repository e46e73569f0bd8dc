// host_nms.cu -- the HOST-array NMS entry points: `_nms` (the reference's only FFI, models/gpu_nms.hpp:9-10, called from
// models/gpu_nms.pyx:16-29) and frcnn_cpu_nms_host (behind models.cpu_nms.cpu_nms: /root/reference models/cpu_nms.pyx:18-69,
// called per class by forward.py:48-57 and by models/proposal_layer.py:176-178).
//
// The caller hands over HOST arrays and wants a HOST keep list back, 20 times per image with n <= 300 rows
// (forward.py's per-class loop), so what matters here is the round trip, not the arithmetic.  Per calling thread the library
// keeps one context: a stream, a MAPPED pinned staging block (dets in, keep list + completion flag out) and a device
// workspace -- nothing is allocated or freed per call (the reference mallocs and frees its device scratch on every call,
// nms_kernel.cu:100-108,142-143).  n <= 2048: ONE single-CTA kernel reads the boxes straight from the mapped host block
// (zero-copy over PCIe), ranks them by counting (score descending, ties: lower index first), runs the greedy suppression
// with cpu_nms.pyx's exact float32 IoU / double compare (or `_nms`'s float `>`), writes the keep list into the mapped
// block and raises a flag the host thread polls: no cudaMemcpy, no cudaStreamSynchronize on the fast path.
// n > 2048: the chip-wide pipeline of frcnn_nms (sort, 64x64 IoU bitmask, device-side scan) on the context's buffers.
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace frcnn {

constexpr int kSmallMax = 2048, kSmallThreads = 256;

__device__ __forceinline__ float hn_iou(const float4 a, const float4 b) {     // models/cpu_nms.pyx:37-41,58-65
    const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.0f), __fadd_rn(__fsub_rn(a.w, a.y), 1.0f));
    const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// dets: [n, dim] floats in mapped host memory (x1,y1,x2,y2[,score]); out[0] = completion flag (ticket), out[1] = count,
// out[2..] = keep list (original indices, descending score).
__global__ void __launch_bounds__(kSmallThreads) nms_small_kernel(const float* __restrict__ dets, int n, int dim, int presorted,
                                                                  double thr_d, float thr_f, int mode, volatile int* out,
                                                                  int ticket, int use_mask) {
    extern __shared__ __align__(16) unsigned char sm[];
    float4* sbox = reinterpret_cast<float4*>(sm);               // [n] boxes by rank
    float4* rbox = sbox + n;                                    // [n] boxes by row
    float* sc = reinterpret_cast<float*>(rbox + n);             // [n] scores by row
    int* order = reinterpret_cast<int*>(sc + n);                // [n] row by rank
    unsigned char* dead = reinterpret_cast<unsigned char*>(order + n);
    const int tid = threadIdx.x;
    // the rows live in MAPPED HOST memory: read the dense [n*dim] floats once, coalesced (consecutive lanes = consecutive
    // words -> few, large PCIe reads), into shared memory, then regroup
    float* flat = reinterpret_cast<float*>(sbox);            // sbox [n] float4 is not live yet: n*dim <= n*5 floats fit in 2n float4
    for (int k = tid; k < n * dim; k += kSmallThreads) flat[k] = dets[k];
    __syncthreads();
    float4 my_box[(kSmallMax + kSmallThreads - 1) / kSmallThreads];
    float my_sc[(kSmallMax + kSmallThreads - 1) / kSmallThreads];
    {
        int q = 0;
        for (int r = tid; r < n; r += kSmallThreads, ++q) {
            const float* d = flat + (size_t)r * dim;
            my_box[q] = make_float4(d[0], d[1], d[2], d[3]);
            my_sc[q] = presorted ? 0.0f : d[4];
        }
    }
    __syncthreads();
    {
        int q = 0;
        for (int r = tid; r < n; r += kSmallThreads, ++q) {
            rbox[r] = my_box[q];
            sc[r] = my_sc[q];
        }
    }
    __syncthreads();
    for (int r = tid; r < n; r += kSmallThreads) {
        int rank = r;
        if (!presorted) {
            const float s = sc[r];
            rank = 0;
            for (int q = 0; q < n; ++q) {
                const float t = sc[q];
                rank += (t > s) || (t == s && q < r);
            }
        }
        order[rank] = r;
        sbox[rank] = rbox[r];
        dead[rank] = 0;
    }
    __syncthreads();
    if (use_mask) {
        // suppression bitmask in shared memory (row i = the later ranks box i suppresses, upper triangle), then ONE warp walks
        // the ranks with the `removed` words in registers: no block-wide barrier on the greedy chain
        const int nw = (n + 63) >> 6;
        unsigned long long* mask = reinterpret_cast<unsigned long long*>(dead + ((n + 15) & ~15));
        // item = (word w, row i), i fastest: a warp's lanes test consecutive rows against the same 64 columns (broadcast reads of
        // sbox[j], conflict-free reads of sbox[i]); product pre-test wherever it provably equals the exact decision
        const float thr_lo = thr_f * (1.0f - 1e-5f), thr_hi = thr_f * (1.0f + 1e-5f);
        const bool fast_ok = thr_f > 1e-3f;
        for (int it = tid; it < n * nw; it += kSmallThreads) {
            const int w = it / n, i = it - w * n;
            unsigned long long bits = 0ull;
            if (w >= (i >> 6)) {
                const float4 a = sbox[i];
                const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.0f), __fadd_rn(__fsub_rn(a.w, a.y), 1.0f));
                const int j0 = w << 6, j1 = min(n, j0 + 64);
                for (int j = max(j0, i + 1); j < j1; ++j) {
                    const float4 b = sbox[j];
                    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
                    const float ww = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
                    const float hh = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
                    const float inter = __fmul_rn(ww, hh);
                    if (fast_ok && inter == 0.0f) continue;
                    const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
                    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
                    bool sup;
                    if (fast_ok && uni > 0.0f && inter > thr_hi * uni) sup = true;
                    else if (fast_ok && uni > 0.0f && inter < thr_lo * uni) sup = false;
                    else {
                        const float ovr = __fdiv_rn(inter, uni);
                        sup = mode == FRCNN_NMS_GE_DOUBLE ? ((double)ovr >= thr_d) : (ovr > thr_f);
                    }
                    if (sup) bits |= 1ull << (j - j0);
                }
            }
            mask[(size_t)i * nw + w] = bits;
        }
        __syncthreads();
        if (tid < 32) {
            unsigned long long removed = 0ull;            // lane w: ranks [64w, 64w + 64)
            int* s_keep = reinterpret_cast<int*>(rbox);   // rbox is dead: the keep list is collected in shared memory ...
            int nk = 0;
            for (int i = 0; i < n; ++i) {
                const unsigned long long word = __shfl_sync(0xffffffffu, removed, i >> 6);
                if ((word >> (i & 63)) & 1ull) continue;
                if (tid == 0) s_keep[nk] = order[i];
                ++nk;
                if (tid < nw) removed |= mask[(size_t)i * nw + tid];
            }
            __syncwarp();
            // ... and leaves for the host in a few wide PCIe writes (one 4-byte system-memory store per kept box from a single
            // lane was the most expensive part of this kernel)
            for (int k = tid; k < nk; k += 32) out[2 + k] = s_keep[k];
            if (tid == 0) out[1] = nk;
            __threadfence_system();                        // keep list + count visible to the host before the flag
            __syncwarp();
            if (tid == 0) out[0] = ticket;
        }
        return;
    }
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;                      // uniform: dead[] is only written ahead of a barrier
        const float4 a = sbox[i];
        if (tid == 0) out[2 + nk] = order[i];
        ++nk;
        for (int j = i + 1 + tid; j < n; j += kSmallThreads) {
            if (dead[j]) continue;
            const float ovr = hn_iou(a, sbox[j]);
            if (mode == FRCNN_NMS_GE_DOUBLE ? ((double)ovr >= thr_d) : (ovr > thr_f)) dead[j] = 1;
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[1] = nk;
        __threadfence_system();                    // keep list + count visible to the host before the flag
        out[0] = ticket;
    }
}

struct HostNmsCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    float* h_in = nullptr;          // mapped pinned: dets
    int* h_out = nullptr;           // mapped pinned: flag, count, keep[]
    float* d_in = nullptr;          // device aliases of the two blocks
    int* d_out = nullptr;
    int cap = 0;                    // rows the blocks hold
    char* d_ws = nullptr;           // device scratch of the large-n path
    size_t ws_bytes = 0;
    int ticket = 0;
    size_t smem_attr = 48 * 1024;   // largest dynamic shared-memory size set on nms_small_kernel by this thread so far
    // No destructor on purpose: cudaFreeHost / cudaFree at THREAD EXIT synchronise the whole device and hold the context lock
    // for milliseconds, stalling every other caller thread that is still running (measured: two of six threads took 3x as
    // long, profiles/r02_api_threads_diag_before_plan_pool.txt).  A thread's ~50 KB of pinned staging and its stream stay
    // allocated until the process exits.
};

static int ctx_prepare(HostNmsCtx& c, int device_id, int n, bool large) {
    if (c.device != device_id) {
        if (c.device >= 0) {
            cudaSetDevice(c.device);
            if (c.h_in) cudaFreeHost(c.h_in);
            if (c.h_out) cudaFreeHost(c.h_out);
            if (c.d_ws) cudaFree(c.d_ws);
            if (c.stream) cudaStreamDestroy(c.stream);
            c.h_in = nullptr; c.h_out = nullptr; c.d_ws = nullptr; c.stream = nullptr; c.cap = 0; c.ws_bytes = 0;
        }
        FRCNN_CUDA_OK(cudaSetDevice(device_id));
        FRCNN_CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        c.device = device_id;
    } else {
        FRCNN_CUDA_OK(cudaSetDevice(device_id));
    }
    if (n > c.cap) {
        const int cap = n < 1024 ? 1024 : (n + 1023) / 1024 * 1024;
        if (c.h_in) cudaFreeHost(c.h_in);
        if (c.h_out) cudaFreeHost(c.h_out);
        c.h_in = nullptr; c.h_out = nullptr; c.cap = 0;
        FRCNN_CUDA_OK(cudaHostAlloc((void**)&c.h_in, sizeof(float) * 5 * (size_t)cap, cudaHostAllocMapped));
        FRCNN_CUDA_OK(cudaHostAlloc((void**)&c.h_out, sizeof(int) * ((size_t)cap + 2), cudaHostAllocMapped));
        FRCNN_CUDA_OK(cudaHostGetDevicePointer((void**)&c.d_in, c.h_in, 0));
        FRCNN_CUDA_OK(cudaHostGetDevicePointer((void**)&c.d_out, c.h_out, 0));
        c.h_out[0] = 0;
        c.cap = cap;
    }
    if (large) {
        const size_t need = align_up(sizeof(float) * 5 * (size_t)n, 256) + align_up(sizeof(int) * (size_t)n, 256) + 256 +
                            frcnn_nms_workspace_bytes(n);
        if (need > c.ws_bytes) {
            if (c.d_ws) cudaFree(c.d_ws);
            c.d_ws = nullptr; c.ws_bytes = 0;
            FRCNN_CUDA_OK(cudaMalloc((void**)&c.d_ws, need));
            c.ws_bytes = need;
        }
    }
    return FRCNN_OK;
}

static int nms_host_impl(const float* dets_host, int n, int dim, double thresh, int mode, int presorted, int* keep_out_host,
                         int device_id) {
    if (n < 0 || dim < 4 || (!presorted && dim < 5)) { set_error("nms host: bad n=%d dim=%d", n, dim); return FRCNN_ERR_ARG; }
    if (n == 0) return 0;
    if (n > 16384) { set_error("nms host: n=%d > 16384 (the device pipeline's limit; see INTEGRATION.md)", n); return FRCNN_ERR_ARG; }
    if (!dets_host || !keep_out_host) { set_error("nms host: NULL argument"); return FRCNN_ERR_ARG; }
    static thread_local HostNmsCtx ctx;
    const bool large = n > kSmallMax;
    // device_id < 0: the calling thread's CURRENT device (cpu_nms has no device argument: a process that owns GPU 3 must
    // not be moved to GPU 0 by its per-class NMS).  An explicit id (`_nms`, gpu_nms.pyx:16) is honoured and the caller's
    // current device restored afterwards.
    int prev_dev = 0;
    FRCNN_CUDA_OK(cudaGetDevice(&prev_dev));
    if (device_id < 0) device_id = prev_dev;
    struct Restore {
        int dev, prev;
        ~Restore() { if (dev != prev) cudaSetDevice(prev); }
    } restore{device_id, prev_dev};
    int rc = ctx_prepare(ctx, device_id, n, large);
    if (rc != FRCNN_OK) return rc;
    if (!large) {
        // stage rows densely as [n, dim'] (dim' = 4 or 5) in the mapped block
        const int sd = presorted ? 4 : 5;
        if (dim == sd) {
            memcpy(ctx.h_in, dets_host, sizeof(float) * (size_t)sd * n);
        } else {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < sd; ++j) ctx.h_in[(size_t)sd * i + j] = dets_host[(size_t)dim * i + j];
        }
        const int ticket = ++ctx.ticket == 0 ? ++ctx.ticket : ctx.ticket;
        size_t smem = (size_t)n * (2 * sizeof(float4) + sizeof(float) + sizeof(int)) + ((n + 15) & ~15) + 16;
        const size_t mask_bytes = (size_t)n * ((n + 63) / 64) * 8;
        const int use_mask = smem + mask_bytes <= 160 * 1024;
        if (use_mask) smem += mask_bytes;
        if (smem > ctx.smem_attr) {                      // raise the kernel's dynamic shared-memory limit only when needed
            FRCNN_CUDA_OK(cudaFuncSetAttribute(nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx.smem_attr = smem;
        }
        FRCNN_CUDA_OK(launch_plain(nms_small_kernel, dim3(1), dim3(kSmallThreads), smem, ctx.stream, (const float*)ctx.d_in, n, sd,
                                   presorted, thresh, (float)thresh, mode, (volatile int*)ctx.d_out, ticket, use_mask));
        // poll the completion flag in mapped memory; if it does not show up soon, fall back to a stream sync (which also
        // surfaces an execution error)
        volatile int* flag = ctx.h_out;
        const auto t0 = std::chrono::steady_clock::now();
        long spins = 0;
        while (flag[0] != ticket) {
            if ((++spins & 0x3ff) == 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.25) {
                FRCNN_CUDA_OK(cudaStreamSynchronize(ctx.stream));
                if (flag[0] != ticket) { set_error("nms host: kernel finished without raising its flag"); return FRCNN_ERR_CUDA; }
                break;
            }
        }
        const int num = flag[1];
        memcpy(keep_out_host, (const void*)(ctx.h_out + 2), sizeof(int) * (size_t)num);
        return num;
    }
    // ---- large n: the chip-wide pipeline on the context's buffers.  [n,5] staging: a pre-sorted input without scores
    // gets strictly descending synthetic scores so the internal (stable) sort is the identity.
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < 4; ++j) ctx.h_in[5 * (size_t)i + j] = dets_host[(size_t)dim * i + j];
        ctx.h_in[5 * (size_t)i + 4] = presorted ? (float)(n - i) : dets_host[(size_t)dim * i + 4];
    }
    const size_t dets_b = align_up(sizeof(float) * 5 * (size_t)n, 256), keep_b = align_up(sizeof(int) * (size_t)n, 256);
    float* d_dets = reinterpret_cast<float*>(ctx.d_ws);
    int* d_keep = reinterpret_cast<int*>(ctx.d_ws + dets_b);
    int* d_num = reinterpret_cast<int*>(ctx.d_ws + dets_b + keep_b);
    void* d_ws = ctx.d_ws + dets_b + keep_b + 256;
    FRCNN_CUDA_OK(cudaMemcpyAsync(d_dets, ctx.h_in, sizeof(float) * 5 * (size_t)n, cudaMemcpyHostToDevice, ctx.stream));
    rc = frcnn_nms(d_dets, n, thresh, mode, 0, d_keep, d_num, d_ws, ctx.ws_bytes - (dets_b + keep_b + 256), ctx.stream);
    if (rc != FRCNN_OK) return rc;
    FRCNN_CUDA_OK(cudaMemcpyAsync(ctx.h_out + 1, d_num, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    FRCNN_CUDA_OK(cudaMemcpyAsync(ctx.h_out + 2, d_keep, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx.stream));
    FRCNN_CUDA_OK(cudaStreamSynchronize(ctx.stream));
    const int num = ctx.h_out[1];
    memcpy(keep_out_host, ctx.h_out + 2, sizeof(int) * (size_t)num);
    return num;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                     float nms_overlap_thresh, int device_id) {
    FRCNN_ENTRY();
    int r = nms_host_impl(boxes_host, boxes_num, boxes_dim, (double)nms_overlap_thresh, FRCNN_NMS_GT_FLOAT, 1, keep_out,
                          device_id);
    *num_out = r < 0 ? -1 : r;
}

extern "C" int frcnn_cpu_nms_host(const float* dets_host, int n, double thresh, int* keep_out_host, int device_id) {
    FRCNN_ENTRY();
    return nms_host_impl(dets_host, n, 5, thresh, FRCNN_NMS_GE_DOUBLE, 0, keep_out_host, device_id);
}

// ------------------------------------------------------------------------------------------ host staging helpers
// Pinned host blocks and explicit async copies for the host-array front ends (engine.ForwardPlan.forward_host,
// engine.StreamRunner): the image upload and the result download go through cudaMemcpyAsync on the plan's stream from /
// to memory this library pinned itself, so that their cost is the link's (measured: 7.2 MB in 0.136 ms = 53 GB/s,
// profiles/r02_h2d_probe.txt) and does not depend on a framework's pinned-memory heuristics.
extern "C" void* frcnn_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) return nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        set_error("frcnn_host_alloc: cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

extern "C" int frcnn_host_free(void* p) {
    FRCNN_ENTRY();
    if (p != nullptr) FRCNN_CUDA_OK(cudaFreeHost(p));
    return FRCNN_OK;
}

extern "C" int frcnn_memcpy_h2d_async(void* dst_device, const void* src_host, size_t bytes, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(dst_device && src_host, "frcnn_memcpy_h2d_async: NULL pointer");
    FRCNN_CUDA_OK(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return FRCNN_OK;
}

extern "C" int frcnn_memcpy_d2h_async(void* dst_host, const void* src_device, size_t bytes, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(dst_host && src_device, "frcnn_memcpy_d2h_async: NULL pointer");
    FRCNN_CUDA_OK(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return FRCNN_OK;
}

extern "C" int frcnn_stream_synchronize(void* stream) {
    FRCNN_ENTRY();
    FRCNN_CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return FRCNN_OK;
}

// Pageable -> pinned staging copy on a few host threads that SLEEP between jobs (condition variable, no spinning).  The
// first version used the framework's parallel copy: 64 OpenMP workers that spin for milliseconds after every parallel
// region; with one such region per image the process ran into ~70 ms stalls (CPU-quota throttling on the GPU boxes,
// profiles/r02_host_api_profile.txt).  4 workers move the 7.2 MB image in ~0.15-0.2 ms and cost nothing when idle.
namespace frcnn {
class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();          // leaked on purpose: worker threads must not be joined at exit
        return *p;
    }
    void copy(char* dst, const char* src, size_t bytes) {
        const int parts = kWorkers + 1;
        const size_t chunk = ((bytes / parts) + 4095) & ~size_t(4095);
        if (bytes < (1u << 20) || chunk == 0) { memcpy(dst, src, bytes); return; }
        std::unique_lock<std::mutex> run(run_mu_);      // one job at a time (callers from several threads queue here)
        {
            std::lock_guard<std::mutex> g(mu_);
            dst_ = dst; src_ = src; bytes_ = bytes; chunk_ = chunk;
            pending_ = kWorkers;
            ++job_;
        }
        cv_.notify_all();
        const size_t o = (size_t)kWorkers * chunk;       // the caller copies the last part itself
        if (o < bytes) memcpy(dst + o, src + o, bytes - o);
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [&] { return pending_ == 0; });
    }

  private:
    static constexpr int kWorkers = 3;
    CopyPool() {
        for (int i = 0; i < kWorkers; ++i) std::thread([this, i] { work(i); }).detach();
    }
    void work(int i) {
        unsigned long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> g(mu_);
            cv_.wait(g, [&] { return job_ != seen; });
            seen = job_;
            char* d = dst_; const char* s = src_;
            const size_t bytes = bytes_, chunk = chunk_;
            g.unlock();
            const size_t o = (size_t)i * chunk;
            if (o < bytes) memcpy(d + o, s + o, o + chunk <= bytes ? chunk : bytes - o);
            g.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_;
    char* dst_ = nullptr; const char* src_ = nullptr;
    size_t bytes_ = 0, chunk_ = 0;
    int pending_ = 0;
    unsigned long job_ = 0;
};
}  // namespace frcnn

extern "C" int frcnn_host_copy(void* dst, const void* src, size_t bytes) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(dst && src, "frcnn_host_copy: NULL pointer");
    frcnn::CopyPool::get().copy(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    return FRCNN_OK;
}

// Upload of a PAGEABLE host buffer through a pinned staging block, pipelined: chunk k is copied into the staging block by
// the calling thread while chunk k-1 is already on the wire (cudaMemcpyAsync on `stream`), so the image costs
// max(host memcpy, DMA) + one chunk instead of their sum -- and no helper thread is involved (the GPU boxes give the process
// a 16-CPU quota; a thread pool that wakes for every image costs more than it saves there, profiles/r02_host_api_profile.txt).
// `staging` must hold `bytes` (the whole image: the chunks of an image never share staging memory, so the host never waits).
extern "C" int frcnn_upload_pageable(void* dst_device, const void* src_host, void* staging_pinned, size_t bytes, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(dst_device && src_host && staging_pinned, "frcnn_upload_pageable: NULL pointer");
    const size_t chunk = 1u << 20;
    for (size_t o = 0; o < bytes; o += chunk) {
        const size_t n = bytes - o < chunk ? bytes - o : chunk;
        memcpy(static_cast<char*>(staging_pinned) + o, static_cast<const char*>(src_host) + o, n);
        FRCNN_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(dst_device) + o, static_cast<char*>(staging_pinned) + o, n,
                                      cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    }
    return FRCNN_OK;
}
