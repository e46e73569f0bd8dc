// resnet_ops.cu -- the memory-bound kernels a ResNet trunk adds around the tensor-core convolutions (SURVEY.md 8f rank 2,
// BASELINE config #4).  The reference's trunk is chainer.links.model.vision.resnet.ResNetLayers (un-vendored;
// /root/reference models/resnet.py:11-45 only subclasses it and returns 'res5'); its published structure is restated:
//   conv1 7x7/2 pad 3 + BN + ReLU          -> frcnn_pack_image_im2col (K = 147 -> 160) + a 1x1 tensor-core GEMM (BN folded)
//   pool1 max 3x3/2, pad 0, cover_all      -> frcnn_maxpool3x3s2_ceil
//   bottleneck 1x1(/s) - 3x3 - 1x1 + shortcut, BN folded into every conv, ReLU after the add
//                                          -> frcnn_conv2d / frcnn_conv2d_res (residual add fused into the epilogue);
//                                             a stride-2 1x1 convolution = frcnn_subsample2x + a stride-1 1x1 convolution
// All three kernels are HBM-bound: 16-byte accesses along the channel axis of the NHWC bf16 hi/lo planes.
#include "common.cuh"

namespace frcnn {

__device__ __forceinline__ void load8f(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long off, float v[8]) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi + off);
    const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(hb[j]);
    if (lo) {
        const uint4 l = *reinterpret_cast<const uint4*>(lo + off);
        const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += __bfloat162float(lb[j]);
    }
}

__device__ __forceinline__ void store8f(__nv_bfloat16* hi, __nv_bfloat16* lo, long off, const float v[8]) {
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16(v[j], h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l);
}

// General first-layer im2col: (C,H,W) fp32 -> [Ho][Wo][Kp] bf16 hi/lo, K index (r*ks + s)*C + c, zero padded borders,
// zeros for k >= ks*ks*C.  One thread per 8 consecutive k of one output pixel.
__global__ void im2col_kernel(const float* __restrict__ x, int C, int H, int W, int ks, int stride, int pad, int Ho, int Wo,
                              int Kp, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int k8n = Kp / 8;
    const long total = (long)Ho * Wo * k8n;
    const int kmax = ks * ks * C;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int k8 = (int)(t % k8n);
        const long p = t / k8n;
        const int wo = (int)(p % Wo), ho = (int)(p / Wo);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k8 * 8 + j;
            float val = 0.f;
            if (k < kmax) {
                const int tap = k / C, c = k - tap * C;
                const int hh = ho * stride - pad + tap / ks, ww = wo * stride - pad + tap % ks;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = x[((long)c * H + hh) * W + ww];
            }
            v[j] = val;
        }
        store8f(hi, lo, p * Kp + k8 * 8, v);
    }
}

// OIHW (Cout, C, ks, ks) fp32 -> [1][Cout][Kp] bf16 hi/lo in the K order of im2col_kernel; scale[o] (optional) multiplies
// output channel o (a folded test-mode BatchNorm).
__global__ void pack_weights_im2col_general_kernel(const float* __restrict__ w, const float* __restrict__ scale, int Cout, int C,
                                                   int ks, int Kp, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * Kp) return;
    const int o = i / Kp, k = i % Kp;
    float v = 0.f;
    if (k < ks * ks * C) {
        const int tap = k / C, c = k % C;
        v = w[((long)o * C + c) * ks * ks + tap];
        if (scale) v = __fmul_rn(v, scale[o]);
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
}

// F.max_pooling_2d(x, 3, stride=2) with Chainer's defaults pad=0, cover_all=True: Ho = ceil((H-3)/2) + 1, windows clipped
// at the bottom / right border.
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, int H, int W, int C,
                                    int Ho, int Wo, __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl) {
    const int C8 = C / 8;
    const long total = (long)Ho * Wo * C8;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const long p = i / C8;
        const int wo = (int)(p % Wo), ho = (int)(p / Wo);
        float m[8];
        bool first = true;
        for (int dy = 0; dy < 3; ++dy) {
            const int h = ho * 2 + dy;
            if (h >= H) break;
            for (int dx = 0; dx < 3; ++dx) {
                const int w = wo * 2 + dx;
                if (w >= W) break;
                float v[8];
                load8f(xh, xl, ((long)h * W + w) * C + c8 * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = first ? v[j] : fmaxf(m[j], v[j]);
                first = false;
            }
        }
        store8f(yh, yl, p * C + c8 * 8, m);
    }
}

// pixels (2h, 2w): the input of a stride-2, pad-0 1x1 convolution
__global__ void subsample2x_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, int H, int W, int C,
                                   int Ho, int Wo, __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl) {
    const int C8 = C / 8;
    const long total = (long)Ho * Wo * C8;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const long p = i / C8;
        const int wo = (int)(p % Wo), ho = (int)(p / Wo);
        const long src = ((long)(2 * ho) * W + 2 * wo) * C + c8 * 8, dst = p * C + c8 * 8;
        *reinterpret_cast<uint4*>(yh + dst) = *reinterpret_cast<const uint4*>(xh + src);
        if (xl) *reinterpret_cast<uint4*>(yl + dst) = *reinterpret_cast<const uint4*>(xl + src);
    }
}

static int grid_cover(long total, int block) {
    long g = (total + block - 1) / block;
    const long cap = 148l * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_pack_image_im2col(const float* x_chw, int C, int H, int W, int ksize, int stride, int pad, int K_pad, void* y_hi,
                            void* y_lo, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(x_chw && y_hi && C > 0 && H > 0 && W > 0 && ksize > 0 && stride > 0 && pad >= 0, "pack_image_im2col: bad arguments");
    FRCNN_REQUIRE(K_pad % 8 == 0 && K_pad >= ksize * ksize * C, "pack_image_im2col: K_pad must be a multiple of 8 and >= ksize^2*C");
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    FRCNN_REQUIRE(Ho > 0 && Wo > 0, "pack_image_im2col: image smaller than the filter");
    const long total = (long)Ho * Wo * (K_pad / 8);
    im2col_kernel<<<grid_cover(total, 256), 256, 0, (cudaStream_t)stream>>>(x_chw, C, H, W, ksize, stride, pad, Ho, Wo, K_pad,
                                                                         (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_pack_conv_weights_im2col(const float* w_oihw, const float* scale, int Cout, int Cin, int ksize, int K_pad, void* w_hi,
                                   void* w_lo, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(w_oihw && w_hi && Cout > 0 && Cin > 0 && ksize > 0 && K_pad % 8 == 0 && K_pad >= ksize * ksize * Cin,
                  "pack_conv_weights_im2col: bad arguments");
    pack_weights_im2col_general_kernel<<<cdiv(Cout * K_pad, 256), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, scale, Cout, Cin, ksize, K_pad, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_maxpool3x3s2_ceil(const void* x_hi, const void* x_lo, int H, int W, int C, void* y_hi, void* y_lo, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(x_hi && y_hi && H >= 3 && W >= 3 && C > 0 && C % 8 == 0, "maxpool3x3s2_ceil: bad arguments (H=%d W=%d C=%d)", H, W, C);
    FRCNN_REQUIRE((x_lo == nullptr) == (y_lo == nullptr), "maxpool3x3s2_ceil: lo planes must both be given or both NULL");
    const int Ho = (H - 3 + 1) / 2 + 1, Wo = (W - 3 + 1) / 2 + 1;            // ceil((H-3)/2) + 1
    const long total = (long)Ho * Wo * (C / 8);
    maxpool3x3s2_kernel<<<grid_cover(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, H, W, C, Ho, Wo, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_subsample2x(const void* x_hi, const void* x_lo, int H, int W, int C, void* y_hi, void* y_lo, void* stream) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(x_hi && y_hi && H > 0 && W > 0 && C > 0 && C % 8 == 0, "subsample2x: bad arguments");
    FRCNN_REQUIRE((x_lo == nullptr) == (y_lo == nullptr), "subsample2x: lo planes must both be given or both NULL");
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const long total = (long)Ho * Wo * (C / 8);
    subsample2x_kernel<<<grid_cover(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, H, W, C, Ho, Wo, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

}  // extern "C"
