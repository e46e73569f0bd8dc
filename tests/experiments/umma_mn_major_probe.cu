// Hardware probe (not product code): MN-major ("transposed") tcgen05.mma operands read straight from a pixel-major tile.
//
// Question (DESIGN.md "what comes next", training item): can the weight-gradient contraction
//     dW[co][ci] = sum_pixels dY[p][co] * X[p][ci]
// be issued on NHWC tiles as they are -- rows = pixels (the K axis), 128 B of channels per row, SWIZZLE_128B exactly as the
// forward kernel's TMA loads them -- by flagging both operands MN-major in the instruction descriptor (bits 15 / 16), so that
// no transposed copy of dY / X is ever written?
//
// Test: A tile = X[k][m] (64 pixel rows x 128 channels: two 64-channel blocks of [64][128 B], LBO apart), B tile = E[k][n]
// (64 pixel rows x 64 channels) with E = identity on (k, n).  D[m][n] = sum_k X[k][m] * E[k][n] = X[n][m]: the accumulator
// must hold the TRANSPOSE of X.  Descriptor fields swept: SBO (stride between 8-row K groups) and LBO (stride between
// 64-element MN blocks); the K advance per MMA (16 rows) is +2048 B on the start address.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I chainer-faster-rcnn_b200/csrc -o umma_mn_probe umma_mn_major_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include "sm100_ptx.cuh"

struct Cfg { int sbo_bytes; int lbo_bytes; int k_step_bytes; int b_mn_major; };

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, int lbo_bytes, int sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((uint32_t)lbo_bytes >> 4) << 16;
    d |= (uint64_t)((uint32_t)sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tme,
                                                 const __grid_constant__ CUtensorMap tmek, Cfg cfg, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                  // 2 x [64 rows][128 B] (channel blocks 0 and 1), 8192 B each
    uint8_t* sb = smem + 16384;          // [64 rows][128 B]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
    uint64_t* bar2 = bar + 1;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(tptr, 64); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(bar, 16384 + 8192);
        ptx::tma_load_2d(sa, &tmx, bar, 0, 0);               // channels 0..63,  pixel rows 0..63
        ptx::tma_load_2d(sa + 8192, &tmx, bar, 64, 0);       // channels 64..127
        ptx::tma_load_2d(sb, cfg.b_mn_major ? &tme : &tmek, bar, 0, 0);
        ptx::mbar_wait(bar, 0);
        ptx::tc_fence_after();
        // idesc: f32 accumulate, bf16 A/B, a_major = MN (bit 15), b_major = MN (bit 16) or K-major identity
        uint32_t idesc = ptx::make_idesc_f16(128, 64, 1) | (1u << 15) | (cfg.b_mn_major ? (1u << 16) : 0u);
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_mn(ptx::smem_u32(sa) + k * cfg.k_step_bytes, cfg.lbo_bytes, cfg.sbo_bytes);
            const uint64_t db = cfg.b_mn_major ? desc_mn(ptx::smem_u32(sb) + k * cfg.k_step_bytes, cfg.lbo_bytes, cfg.sbo_bytes)
                                               : (ptx::make_smem_desc(ptx::smem_u32(sb), 128) + 2 * k);
            ptx::mma_f16_ss(tmem, da, db, idesc, k != 0);
        }
        ptx::mma_commit(bar2);
    }
    ptx::mbar_wait(bar2, 0);
    ptx::tc_fence_after();
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

// Second question: the forward kernel's HALO trick for MN-major operands.  A = [128 pixel rows][128 B] x 2 channel blocks;
// the MMA's K axis walks rows start + (k/8)*pitch_groups + k%8 (8-row groups `sbo` bytes apart, first row `start` not a
// multiple of 8) -- a tap of a 3x3 filter over a (16+2) x (8+2) halo patch is start = r*10+s, sbo = 1280 B.
__global__ void __launch_bounds__(128, 1) probe_halo(const __grid_constant__ CUtensorMap tmx2, const __grid_constant__ CUtensorMap tmek,
                                                      int start_row, int sbo_bytes, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                  // 2 x [128 rows][128 B], 16384 B each
    uint8_t* sb = smem + 32768;          // identity [64][128 B], K-major
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
    uint64_t* bar2 = bar + 1;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(tptr, 64); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(bar, 32768 + 8192);
        ptx::tma_load_2d(sa, &tmx2, bar, 0, 0);
        ptx::tma_load_2d(sa + 16384, &tmx2, bar, 64, 0);
        ptx::tma_load_2d(sb, &tmek, bar, 0, 0);
        ptx::mbar_wait(bar, 0);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::make_idesc_f16(128, 64, 1) | (1u << 15);
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_mn(ptx::smem_u32(sa) + start_row * 128 + k * 2 * sbo_bytes, 16384, sbo_bytes);
            const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(sb), 128) + 2 * k;
            ptx::mma_f16_ss(tmem, da, db, idesc, k != 0);
        }
        ptx::mma_commit(bar2);
    }
    ptx::mbar_wait(bar2, 0);
    ptx::tc_fence_after();
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    // X[k][m]: 64 pixel rows x 128 channels, small integers (exact in bf16): X[k][m] = ((k * 7 + m * 3) % 17) - 8
    std::vector<__nv_bfloat16> hx(64 * 128), he(64 * 64);
    for (int k = 0; k < 64; ++k) for (int m = 0; m < 128; ++m) hx[k * 128 + m] = __float2bfloat16((float)(((k * 7 + m * 3) % 17) - 8));
    for (int k = 0; k < 64; ++k) for (int n = 0; n < 64; ++n) he[k * 64 + n] = __float2bfloat16(k == n ? 1.0f : 0.0f);
    __nv_bfloat16 *dx, *de; float* dout;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&de, he.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(de, he.data(), he.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmx, tme;
    cuuint32_t es[2] = {1, 1};
    { cuuint64_t dims[2] = {128, 64}; cuuint64_t str[1] = {256}; cuuint32_t box[2] = {64, 64};
      enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    { cuuint64_t dims[2] = {64, 64}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 64};
      enc(&tme, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, de, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    std::vector<float> ho(128 * 64);
    const Cfg cfgs[] = {
        {1024, 8192, 2048, 1},     // the canonical reading of cute's Major-MN SW128 layout: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-B units
        {1024, 8192, 2048, 0},     // A MN-major, B = K-major identity (the identity is symmetric)
        {8192, 1024, 2048, 1},     // LBO / SBO swapped
        {1024, 8192, 1024, 1},     // K advance of one 8-row group
        {2048, 8192, 2048, 1},
    };
    for (const Cfg& cfg : cfgs) {
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, 48 * 1024>>>(tmx, tme, tme, cfg, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d lbo %d kstep %d bmn %d: CUDA error %s\n", cfg.sbo_bytes, cfg.lbo_bytes, cfg.k_step_bytes, cfg.b_mn_major, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) ok += ho[m * 64 + n] == __bfloat162float(hx[n * 128 + m]);
        printf("SBO %5d LBO %5d K-step %5d B %s : D == X^T on %5d / 8192 elements | D[0][0..3] = %g %g %g %g (want %g %g %g %g) D[64][0] = %g (want %g) D[1][0] = %g (want %g)\n",
               cfg.sbo_bytes, cfg.lbo_bytes, cfg.k_step_bytes, cfg.b_mn_major ? "MN-major" : "K-major ", ok, ho[0], ho[1], ho[2], ho[3],
               __bfloat162float(hx[0]), __bfloat162float(hx[128]), __bfloat162float(hx[256]), __bfloat162float(hx[384]), ho[64 * 64],
               __bfloat162float(hx[64]), ho[64], __bfloat162float(hx[1]));
    }
    // ---- halo-style start row / group stride on an MN-major operand
    std::vector<__nv_bfloat16> hx2(128 * 128);
    for (int k = 0; k < 128; ++k) for (int m = 0; m < 128; ++m) hx2[k * 128 + m] = __float2bfloat16((float)(((k * 5 + m * 11) % 23) - 11));
    __nv_bfloat16* dx2;
    cudaMalloc(&dx2, hx2.size() * 2);
    cudaMemcpy(dx2, hx2.data(), hx2.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmx2;
    { cuuint64_t dims[2] = {128, 128}; cuuint64_t str[1] = {256}; cuuint32_t box[2] = {64, 128};
      enc(&tmx2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx2, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    cudaFuncSetAttribute(probe_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int starts[] = {0, 1, 10, 11, 12, 21, 22};
    const int sbos[] = {1024, 1280};
    for (int sbo : sbos) for (int st : starts) {
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe_halo<<<1, 128, 64 * 1024>>>(tmx2, tme, st, sbo, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("halo start %d sbo %d: CUDA error %s\n", st, sbo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            const int row = st + (n / 8) * (sbo / 128) + (n % 8);
            ok += ho[m * 64 + n] == __bfloat162float(hx2[row * 128 + m]);
        }
        printf("MN-major halo: start row %2d SBO %4d : D[m][n] == X[start + (n/8)*%d + n%%8][m] on %5d / 8192 elements\n", st, sbo, sbo / 128, ok);
    }
    return 0;
}
