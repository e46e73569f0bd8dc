// train_backward.cu -- the memory-bound kernels of the conv backward pass (train_rpn.py step).
//
// The dense parts of backward run on the SAME tcgen05 kernel as forward (conv_gemm_sm100.cu):
//   data gradient    dX = conv(dY, W rotated 180 deg, in/out swapped)      -> frcnn_conv2d with frcnn_pack_conv_weights_dgrad
//   weight gradient  dW[r][s] = sum_pixels dY[p] (x) X[p + (r-1, s-1)]      -> frcnn_gemm_nt_splitk over a zero-padded,
//                                                                            pixel-contiguous ("transposed") copy of dY and X
// This file holds what surrounds them (all HBM-bound, 16-byte accesses along channels on the NHWC side):
//   frcnn_grad_prepare      ReLU mask (and 2x2 ceil-mode max-pool routing) of an incoming gradient, hi/lo re-split, NHWC
//                           output for the next dgrad and the transposed padded copy for the wgrad GEMM
//                           (Chainer F.relu / F.max_pooling_2d backward, /root/reference models/vgg16.py:38-82 graph)
//   frcnn_act_transpose3    forward activation -> three column-shifted transposed planes (the GEMM's B operand)
//   frcnn_wgrad_reduce      split-K partials -> dW in the reference's OIHW float32 layout (fixed order: deterministic)
//   frcnn_bias_grad         db[c] = sum over pixels of dY
//   frcnn_sgd_momentum      MomentumSGD(lr, momentum) + WeightDecay(rate) hook (train_rpn.py:165-167), float32 masters
#include "common.cuh"

namespace frcnn {

__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }

struct PrepArgs {
    // source gradient: bf16 hi/lo NHWC [Hs][Ws][C]  (Hs,Ws = H,W or the pooled size), or fp32 [H*W][ld] (src_f32)
    const __nv_bfloat16 *g_hi, *g_lo;
    const float* g_f32; int ld_f32;
    // forward activation the gradient flows into (post-ReLU output Y [H][W][C]); NULL = no mask
    const __nv_bfloat16 *y_hi, *y_lo;
    // pooled forward activation P [ceil(H/2)][ceil(W/2)][C] when the gradient arrives at pooled resolution; NULL otherwise
    const __nv_bfloat16 *p_hi, *p_lo;
    int H, W, C, Wp;
    long Kp;
    __nv_bfloat16 *o_hi, *o_lo;      // NHWC [H][W][C] out (optional)
    __nv_bfloat16 *t_hi, *t_lo;      // transposed [planes][C][Kp] out (optional)
    int planes;                      // 1: plane at q; 3: planes shifted by -1/0/+1 pixel
};

constexpr int TP = 64;               // pixels per tile (along w)
constexpr int TC = 64;               // channels per tile
constexpr int PITCH = TP + 2;

// value of the (masked / routed) gradient at (h, w, channel c8*8 + j), j < 8
__device__ __forceinline__ void prep_values(const PrepArgs& a, int h, int w, int c8, float v[8]) {
    const long off = ((long)h * a.W + w) * a.C + c8 * 8;
    float yv[8];
    bool has_y = a.y_hi != nullptr;
    if (has_y) {
        const uint4 yh = *reinterpret_cast<const uint4*>(a.y_hi + off);
        const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&yh);
        uint4 yl = make_uint4(0, 0, 0, 0);
        if (a.y_lo) yl = *reinterpret_cast<const uint4*>(a.y_lo + off);
        const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&yl);
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = bf(hb[j]) + bf(lb[j]);
    }
    if (a.g_f32 != nullptr) {
        const float* g = a.g_f32 + ((long)h * a.W + w) * a.ld_f32 + c8 * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c8 * 8 + j < a.ld_f32) ? g[j] : 0.f;
    } else if (a.p_hi == nullptr) {
        const uint4 gh = *reinterpret_cast<const uint4*>(a.g_hi + off);
        const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&gh);
        uint4 gl = make_uint4(0, 0, 0, 0);
        if (a.g_lo) gl = *reinterpret_cast<const uint4*>(a.g_lo + off);
        const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&gl);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = bf(hb[j]) + bf(lb[j]);
    } else {
        // gradient at pooled resolution: route to the FIRST maximum of the 2x2 (ceil-mode) window in scan order
        const int Wq = (a.W + 1) / 2;
        const int ph = h >> 1, pw = w >> 1;
        const long poff = ((long)ph * Wq + pw) * a.C + c8 * 8;
        const uint4 gh = *reinterpret_cast<const uint4*>(a.g_hi + poff);
        const __nv_bfloat16* ghb = reinterpret_cast<const __nv_bfloat16*>(&gh);
        uint4 gl = make_uint4(0, 0, 0, 0);
        if (a.g_lo) gl = *reinterpret_cast<const uint4*>(a.g_lo + poff);
        const __nv_bfloat16* glb = reinterpret_cast<const __nv_bfloat16*>(&gl);
        const uint4 pmh = *reinterpret_cast<const uint4*>(a.p_hi + poff);
        const __nv_bfloat16* pmhb = reinterpret_cast<const __nv_bfloat16*>(&pmh);
        uint4 pml = make_uint4(0, 0, 0, 0);
        if (a.p_lo) pml = *reinterpret_cast<const uint4*>(a.p_lo + poff);
        const __nv_bfloat16* pmlb = reinterpret_cast<const __nv_bfloat16*>(&pml);
        float pm[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pm[j] = bf(pmhb[j]) + bf(pmlb[j]);
        // earlier window elements (scan order (0,0),(0,1),(1,0),(1,1)) that already equal the maximum take the gradient
        bool earlier[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) earlier[j] = false;
        const int h0 = ph * 2, w0 = pw * 2;
        for (int e = 0; e < 4; ++e) {
            const int eh = h0 + (e >> 1), ew = w0 + (e & 1);
            if (eh == h && ew == w) break;
            if (eh >= a.H || ew >= a.W) continue;
            const long eo = ((long)eh * a.W + ew) * a.C + c8 * 8;
            const uint4 eh4 = *reinterpret_cast<const uint4*>(a.y_hi + eo);
            const __nv_bfloat16* ehb = reinterpret_cast<const __nv_bfloat16*>(&eh4);
            uint4 el4 = make_uint4(0, 0, 0, 0);
            if (a.y_lo) el4 = *reinterpret_cast<const uint4*>(a.y_lo + eo);
            const __nv_bfloat16* elb = reinterpret_cast<const __nv_bfloat16*>(&el4);
#pragma unroll
            for (int j = 0; j < 8; ++j) earlier[j] = earlier[j] || (bf(ehb[j]) + bf(elb[j]) == pm[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (!earlier[j] && yv[j] == pm[j]) ? bf(ghb[j]) + bf(glb[j]) : 0.f;
    }
    if (has_y) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = yv[j] > 0.f ? v[j] : 0.f;        // F.relu backward: gy * (y > 0)
    }
}

__global__ void __launch_bounds__(256) grad_prepare_kernel(PrepArgs a) {
    __shared__ __nv_bfloat16 s_hi[TC][PITCH];
    __shared__ __nv_bfloat16 s_lo[TC][PITCH];
    const int wt = blockIdx.x, h = blockIdx.y, ct = blockIdx.z;
    const int w0 = wt * TP, c0 = ct * TC;
    const int t = threadIdx.x;
    for (int it = 0; it < 2; ++it) {
        const int px = (t >> 3) + it * 32, c8l = t & 7;
        const int w = w0 + px, c8 = (c0 >> 3) + c8l;
        float v[8];
        const bool live = w < a.W && c8 * 8 < a.C;
        if (live) {
            prep_values(a, h, w, c8, v);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        __align__(16) __nv_bfloat16 hi[8];
        __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16(v[j], hi[j], lo[j]);
        if (live && a.o_hi != nullptr) {
            const long off = ((long)h * a.W + w) * a.C + c8 * 8;
            *reinterpret_cast<uint4*>(a.o_hi + off) = *reinterpret_cast<const uint4*>(hi);
            if (a.o_lo) *reinterpret_cast<uint4*>(a.o_lo + off) = *reinterpret_cast<const uint4*>(lo);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s_hi[c8l * 8 + j][px] = hi[j];
            s_lo[c8l * 8 + j][px] = lo[j];
        }
    }
    if (a.t_hi == nullptr) return;
    __syncthreads();
    // transposed stores: a channel row of this tile is <= 64 consecutive k positions starting at q0 - shift; they are
    // written as 16-byte vectors on the 8-element-aligned groups that lie fully inside the row, element-wise at the ends.
    const int warp = t >> 5, lane = t & 31;
    const long q0 = (long)(h + 1) * a.Wp + (w0 + 1);
    const int n = min(TP, a.W - w0);                 // valid pixels of this tile
    const int rsub = lane >> 3, gl = lane & 7;       // 4 channel rows per warp pass, 8 lanes per row
    for (int pl = 0; pl < a.planes; ++pl) {
        const int shift = a.planes == 3 ? pl - 1 : 0;                // plane pl at k holds the operand at k + pl - 1
        const long d0 = q0 - shift;                                  // destination of pixel 0
        const long g0 = d0 & ~7l;                                    // first aligned group
        for (int ch = warp * 4 + rsub; ch < TC; ch += 32) {
            if (c0 + ch >= a.C) continue;
            const long row = ((long)pl * a.C + c0 + ch) * a.Kp;
            for (int grp = gl; grp < 9; grp += 8) {
                const long k0 = g0 + 8 * grp;
                const int px0 = (int)(k0 - d0);                      // pixel of the group's first element (may be < 0)
                if (px0 >= n || px0 + 8 <= 0) continue;
                if (px0 >= 0 && px0 + 8 <= n) {
                    __align__(16) __nv_bfloat16 vh[8];
                    __align__(16) __nv_bfloat16 vl[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { vh[e] = s_hi[ch][px0 + e]; vl[e] = s_lo[ch][px0 + e]; }
                    *reinterpret_cast<uint4*>(a.t_hi + row + k0) = *reinterpret_cast<const uint4*>(vh);
                    if (a.t_lo) *reinterpret_cast<uint4*>(a.t_lo + row + k0) = *reinterpret_cast<const uint4*>(vl);
                } else {
                    for (int e = 0; e < 8; ++e) {
                        const int px = px0 + e;
                        if (px >= 0 && px < n) {
                            a.t_hi[row + k0 + e] = s_hi[ch][px];
                            if (a.t_lo) a.t_lo[row + k0 + e] = s_lo[ch][px];
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ split-K partials -> dW (OIHW fp32)
__global__ void wgrad_reduce_kernel(const float* __restrict__ parts, int groups, int splits, int Mp, int M, int ld, int N,
                                    float* __restrict__ dw, float scale) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)M * N * groups;
    if (t >= total) return;
    const int g = (int)(t % groups);
    const long mn = t / groups;
    const int n = (int)(mn % N), m = (int)(mn / N);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += parts[(((long)g * splits + s) * Mp + m) * ld + n];
    dw[t] = acc * scale;                         // OIHW: ((m*N + n)*groups + g), g = r*3 + s
}

// ------------------------------------------------------------------------------------------ bias gradient from the transposed dY
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ t_hi, const __nv_bfloat16* __restrict__ t_lo,
                                                        long Kp, float* __restrict__ db, float scale) {
    __shared__ float s_w[8];
    const int c = blockIdx.x;
    const __nv_bfloat16* rh = t_hi + (long)c * Kp;
    const __nv_bfloat16* rl = t_lo ? t_lo + (long)c * Kp : nullptr;
    float acc = 0.f;
    for (long k = (long)threadIdx.x * 8; k < Kp; k += 256 * 8) {        // Kp % 64 == 0
        const uint4 h4 = *reinterpret_cast<const uint4*>(rh + k);
        const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h4);
        uint4 l4 = make_uint4(0, 0, 0, 0);
        if (rl) l4 = *reinterpret_cast<const uint4*>(rl + k);
        const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l4);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += bf(hb[j]) + bf(lb[j]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += s_w[w];
        db[c] = a * scale;
    }
}

// ------------------------------------------------------------------------------------------ optimizer
__global__ void sgd_momentum_kernel(float* __restrict__ w, float* __restrict__ v, const float* __restrict__ g, long n,
                                    float lr, float momentum, float weight_decay) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // chainer.optimizer.WeightDecay hook: g += rate * w;  MomentumSGD.update_one: v = momentum*v - lr*g; w += v
    const float gi = __fmaf_rn(weight_decay, w[i], g[i]);
    const float vi = __fsub_rn(__fmul_rn(momentum, v[i]), __fmul_rn(lr, gi));
    v[i] = vi;
    w[i] = __fadd_rn(w[i], vi);
}

// ------------------------------------------------------------------------------------------ dgrad weight packing
// out[t][ci][co] (bf16 hi/lo, co padded to Cout_pad) = W[co][ci][kh-1-r][kw-1-s], t = r*kw + s: the data gradient of a
// stride-1 "same" convolution is the convolution of dY with the 180-degree rotated, in/out-swapped filter.
__global__ void pack_weights_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int Cout_pad,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)taps * Cin * Cout_pad;
    if (t >= total) return;
    const int co = (int)(t % Cout_pad);
    const int ci = (int)((t / Cout_pad) % Cin);
    const int tap = (int)(t / ((long)Cout_pad * Cin));
    float v = 0.f;
    if (co < Cout) v = w[((long)co * Cin + ci) * taps + (taps - 1 - tap)];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[t] = h;
    if (lo) lo[t] = l;
}

}  // namespace frcnn

using namespace frcnn;

static long padded_pixels(int H, int W, int* wp_out) {
    const int Wp = (W + 2 + 7) / 8 * 8;
    if (wp_out) *wp_out = Wp;
    const long raw = (long)(H + 2) * Wp;
    return (raw + 63) / 64 * 64;
}

extern "C" {

long frcnn_padded_pixels(int H, int W, int* row_pitch) { return padded_pixels(H, W, row_pitch); }

int frcnn_grad_prepare(const void* g_hi, const void* g_lo, const float* g_f32, int ld_f32, const void* y_hi, const void* y_lo,
                       const void* p_hi, const void* p_lo, int H, int W, int C, void* o_hi, void* o_lo, void* t_hi, void* t_lo,
                       int planes, void* stream) {
    FRCNN_REQUIRE(H > 0 && W > 0 && C > 0 && C % 8 == 0, "grad_prepare: bad shape H=%d W=%d C=%d (C %% 8 == 0)", H, W, C);
    FRCNN_REQUIRE((g_hi != nullptr) != (g_f32 != nullptr), "grad_prepare: give the source as bf16 planes OR fp32");
    FRCNN_REQUIRE(!g_f32 || (ld_f32 >= 1 && !p_hi), "grad_prepare: fp32 source needs ld_f32 and no pooling");
    FRCNN_REQUIRE(!p_hi || y_hi, "grad_prepare: pooled routing needs the un-pooled forward activation");
    FRCNN_REQUIRE(planes == 1 || planes == 3, "grad_prepare: planes must be 1 or 3");
    FRCNN_REQUIRE(o_hi || t_hi, "grad_prepare: no output requested");
    PrepArgs a;
    a.g_hi = (const __nv_bfloat16*)g_hi; a.g_lo = (const __nv_bfloat16*)g_lo; a.g_f32 = g_f32; a.ld_f32 = ld_f32;
    a.y_hi = (const __nv_bfloat16*)y_hi; a.y_lo = (const __nv_bfloat16*)y_lo;
    a.p_hi = (const __nv_bfloat16*)p_hi; a.p_lo = (const __nv_bfloat16*)p_lo;
    a.H = H; a.W = W; a.C = C;
    a.Kp = padded_pixels(H, W, &a.Wp);
    a.o_hi = (__nv_bfloat16*)o_hi; a.o_lo = (__nv_bfloat16*)o_lo; a.t_hi = (__nv_bfloat16*)t_hi; a.t_lo = (__nv_bfloat16*)t_lo;
    a.planes = planes;
    dim3 grid(cdiv(W, TP), H, cdiv(C, TC));
    FRCNN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "grad_prepare: image too tall / too many channels for one launch");
    grad_prepare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_wgrad_reduce(const float* parts, int groups, int splits, int M_parts, int M, int ld, int N, float scale, float* dw,
                       void* stream) {
    FRCNN_REQUIRE(parts && dw && groups >= 1 && splits >= 1 && M > 0 && M_parts >= M && N > 0 && ld >= N, "wgrad_reduce: bad arguments");
    const long total = (long)M * N * groups;
    wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parts, groups, splits, M_parts, M, ld, N, dw, scale);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_bias_grad(const void* t_hi, const void* t_lo, int C, long Kp, float scale, float* db, void* stream) {
    FRCNN_REQUIRE(t_hi && db && C > 0 && Kp > 0 && Kp % 64 == 0, "bias_grad: bad arguments");
    bias_grad_kernel<<<C, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)t_hi, (const __nv_bfloat16*)t_lo, Kp, db, scale);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_sgd_momentum(float* w, float* v, const float* g, long n, float lr, float momentum, float weight_decay, void* stream) {
    FRCNN_REQUIRE(w && v && g && n >= 0, "sgd_momentum: bad arguments");
    if (n == 0) return FRCNN_OK;
    sgd_momentum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, v, g, n, lr, momentum, weight_decay);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

int frcnn_pack_conv_weights_dgrad(const float* w_oihw, int Cout, int Cin, int kh, int kw, int Cout_pad, void* w_hi, void* w_lo,
                                  void* stream) {
    FRCNN_REQUIRE(w_oihw && w_hi && Cout > 0 && Cin > 0 && kh > 0 && kw > 0 && Cout_pad >= Cout && Cout_pad % 8 == 0,
                  "pack_conv_weights_dgrad: bad arguments");
    const long total = (long)kh * kw * Cin * Cout_pad;
    pack_weights_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, Cout, Cin, kh * kw, Cout_pad, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo);
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

}  // extern "C"
