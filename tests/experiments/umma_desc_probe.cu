// Hardware probe (not product code): which rows does tcgen05.mma read when the K-major SWIZZLE_128B
// smem descriptor starts at a row that is not a multiple of 8 (start address not 1024-B aligned) and/or
// uses a stride-byte-offset (SBO) other than 1024?  D = A * I  (B = 64x64 identity)  =>  D row m == the
// smem row the MMA actually read for logical row m.  X[i][c] = i + c/128 identifies row i and column c.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I chainer-faster-rcnn_b200/csrc -o umma_probe umma_desc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include "sm100_ptx.cuh"

struct Cfg { int start_row; int sbo_bytes; int base_off_mode; };   // base_off_mode: 0 = field 0, 1 = (addr>>7)&7

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmi,
                                                 Cfg cfg, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sx = smem;                 // 256 rows x 128 B
    uint8_t* si = smem + 32768;         // identity 64 x 128 B
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
        ptx::tma_load_2d(sx, &tmx, bar, 0, 0);
        ptx::tma_load_2d(si, &tmi, bar, 0, 0);
        ptx::mbar_wait(bar, 0);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sx) + cfg.start_row * 128;
        uint64_t da = 0;
        da |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)(cfg.sbo_bytes >> 4) << 32;
        da |= (uint64_t)1 << 46;
        if (cfg.base_off_mode == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
        da |= (uint64_t)2 << 61;
        const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(si), 128);
        const uint32_t idesc = ptx::make_idesc_f16(128, 64, 1);
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
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
    const int R = 256;
    std::vector<__nv_bfloat16> hx(R * 64), hi(64 * 64);
    for (int i = 0; i < R; ++i) for (int c = 0; c < 64; ++c) hx[i * 64 + c] = __float2bfloat16((float)i + c / 128.0f);   // exact in bf16? i<256 (8 bits) + c/128 (<0.5, 7 bits) -> not exact; decode by rounding
    for (int i = 0; i < 64; ++i) for (int c = 0; c < 64; ++c) hi[i * 64 + c] = __float2bfloat16(i == c ? 1.0f : 0.0f);
    // use two planes instead: rows encoded exactly: X[i][c] = (c == 0) ? i : (c == 1 ? 1000 + i : c)  -> col 0 gives the row id, other cols give the column id
    for (int i = 0; i < R; ++i) for (int c = 0; c < 64; ++c) hx[i * 64 + c] = __float2bfloat16(c == 0 ? (float)i : (float)c);
    __nv_bfloat16 *dx, *di; float* dout;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&di, hi.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(di, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmx, tmi;
    cuuint64_t dims[2] = {64, (cuuint64_t)R}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dims2[2] = {64, 64}; cuuint32_t box2[2] = {64, 64};
    enc(&tmi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, di, dims2, str, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    std::vector<float> ho(128 * 64);
    const int starts[] = {0, 1, 2, 3, 5, 8, 10, 11, 21};
    const int sbos[] = {1024, 1280, 2048};
    for (int bo = 0; bo < 2; ++bo) for (int sbo : sbos) for (int st : starts) {
        Cfg cfg{st, sbo, bo};
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, 48 * 1024>>>(tmx, tmi, cfg, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("start %d sbo %d bo %d: CUDA error %s\n", st, sbo, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int ok_rows = 0, ok_cols = 0;
        for (int m = 0; m < 128; ++m) {
            const int want = st + (m / 8) * (sbo / 128) + (m % 8);
            if ((int)ho[m * 64] == want) ++ok_rows;
            bool colsok = true;
            for (int c = 1; c < 64; ++c) if ((int)ho[m * 64 + c] != c) colsok = false;
            if (colsok) ++ok_cols;
        }
        printf("start_row %2d sbo %4d base_off_mode %d : rows as linear model %3d/128, columns intact %3d/128 | first rows read:", st, sbo, bo, ok_rows, ok_cols);
        for (int m = 0; m < 10; ++m) printf(" %d", (int)ho[m * 64]);
        printf(" | row8..: %d %d | cols of row0: %d %d %d %d\n", (int)ho[8 * 64], (int)ho[9 * 64], (int)ho[1], (int)ho[8], (int)ho[9], (int)ho[16]);
    }
    return 0;
}
