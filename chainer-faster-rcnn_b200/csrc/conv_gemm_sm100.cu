// conv_gemm_sm100.cu -- implicit-GEMM 3x3 / 1x1 convolution (and plain GEMM) on 5th-gen tensor
// cores: TMA -> 128B-swizzled shared memory -> tcgen05.mma (kind::f16, bf16 operands, fp32
// accumulators in TMEM) -> tcgen05.ld epilogue with fused bias + ReLU (+ bf16 hi/lo split,
// + optional fused 2x2 ceil-mode max-pool) -> swizzled smem staging -> TMA bulk tensor store.
//
// Replaces, for the forward path: L.Convolution2D at /root/reference models/vgg16.py:39-67 and
// models/region_proposal_network.py:53-57, L.Linear at models/faster_rcnn.py:33-36, and (fused)
// F.MaxPooling2D(2,2) at models/vgg16.py:43,48,55,62.
//
// Mapping (NHWC activations, tap-major K-major weights):
//   M = output pixels, tiled as TH x TW patches of 128 pixels;  N = output channels (BN per tile);
//   K = taps x Cin, walked as (tap, 64-channel block) "k-blocks".
//   The A operand of k-block (tap=(r,s), cb) is ONE 3-D TMA box {BK ch, TW, TH} of the input at
//   spatial offset (r-1, s-1): out-of-bounds rows/columns are zero-filled by the TMA unit, which
//   is exactly the conv's zero padding -- no im2col buffer, no halo handling in the kernel.
//   The box lands in smem as 128 rows x 128 B, i.e. the canonical K-major SWIZZLE_128B UMMA tile.
//
// Kernel structure: persistent, one CTA per SM, 10 warps:
//   warp 0    : TMA producer (one elected lane)         smem ring, full/empty mbarriers
//   warp 1    : TMEM allocator + MMA issuer (one lane)  ring of accumulators in TMEM
//   warps 2-9 : epilogue, two groups of four (TMEM lane group = warp % 4; group g takes the 32-column chunks g, g+2, ...);
//               overlaps the next tile's main loop
//
// "bf16x3" mode (lo planes present): per k-block the stage holds A_hi, A_lo, B_hi, B_lo and the
// issuer runs A_hi*B_hi into the main accumulator and A_lo*B_hi + A_hi*B_lo into a separate
// correction accumulator (the tensor-core accumulator truncates on every add).
//
// Epilogue stores: a thread owns one pixel row of the accumulator.  Per 32-channel chunk it writes
// its 64 B (hi) + 64 B (lo) into a SWIZZLE_64B staging tile (bank-conflict free), and one thread
// issues a 3-D TMA store {32 ch, TW, TH} which clips the ragged image edge -- 2 bulk stores per
// chunk instead of 1024 16-byte STG (the first version was store-issue bound on the 64-channel layers).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace frcnn {

// Division by a launch-constant divisor without the ~25-instruction software division the compiler emits for `x / p.field`
// (every role recomputes the tile coordinates once per tile: five divisions were a fifth of the epilogue's instructions on
// the 64-channel layers).  Granlund-Montgomery round-up multiplier: exact for every 32-bit unsigned x and d >= 1.
struct FastDiv {
    uint32_t m, s1, s2, d;
};
__device__ __forceinline__ int fd_div(int x, const FastDiv& f) {
    const uint32_t q = __umulhi(f.m, (uint32_t)x);
    return (int)(((((uint32_t)x - q) >> f.s1) + q) >> f.s2);
}
static inline FastDiv make_fastdiv(int d_) {
    FastDiv f;
    const uint32_t d = d_ > 0 ? (uint32_t)d_ : 1u;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                       // l = ceil(log2 d)
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l - f.s1;
    f.d = d;
    return f;
}

struct ConvParams {
    int H, W, Cout;
    int taps, ksize, cin_blocks;
    int kw, pad_h, pad_w;                // tap -> (r, s) = (tap / kw, tap % kw); A box origin (w0 + s - pad_w, h0 + r - pad_h)
    int TH, TW, tiles_h, tiles_w, n_tiles, num_tiles;
    FastDiv fd_tiles_w, fd_n_tiles, fd_tiles_per_part;   // set by finish_params() once the three divisors are final
    int num_stages, a_stages, x3, relu, pool;
    int acc_bufs, acc_cols, tmem_cols;   // TMEM ring: acc_bufs buffers of acc_cols columns (x3: main | correction)
    // long-K GEMMs (fc6: K = 25,088 .. 100,352): the tensor core's fp32 accumulator truncates on every add, an error that
    // grows with the number of adds x the accumulator's magnitude.  The k-blocks rotate over `nacc` main accumulators
    // (acc_chunk blocks each turn; accumulator a >= 1 sits behind the correction accumulator), summed once in the epilogue with
    // round-to-nearest adds: nacc-fold smaller bias.  nacc == 1: the ordinary single accumulator.
    int nacc, acc_chunk;
    int ld_f32, n_cover;
    int store_bf16, store_lo;            // bf16 outputs go through the staged TMA store
    int stage64;                         // 1: both epilogue groups fill ONE staging tile of 128-byte rows (64 channels) per 64-column block
    int b_res;                           // 1 (per-tap path, one N tile, few k-blocks): ALL weight tiles are loaded once per CTA and stay
                                         // resident in shared memory; the ring then carries A tiles only (conv1_1: 4750 tiles x 24 KB saved)
    float* y_f32;
    const float* bias;
    const int* m_valid;
    const __nv_bfloat16 *res_hi, *res_lo;   // optional residual [H][W][Cout] (hi + lo) added before the ReLU (ResNet shortcut)
    // split-K GEMM mode (frcnn_gemm_nt_splitk, per-tap path only): the tile space is n_parts x tiles_per_part,
    // part = group * splits + split; a split covers k-blocks [split*kb_per_split, ...) of kb_total; group g shifts
    // the B operand's K coordinate by (g/3-1)*g_row_stride and reads plane g%3 of B (the 3x3 taps of a padded pixel
    // axis: a TMA box must start 16-B aligned, so the +-1 column shifts are three pre-shifted planes); every part
    // writes its own fp32 slab y_f32 + part*part_stride.  n_parts == 1: an ordinary convolution.
    int n_parts, tiles_per_part, splits, kb_per_split, kb_total, g_row_stride;
    long part_stride;
};

constexpr int kNumThreads = 320;          // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two groups of four)
constexpr int kTileM = 128;
constexpr int kStagePlane = kTileM * 64;          // one staging plane: 128 rows x 32 bf16
constexpr int kStagingBytes = 2 * 2 * kStagePlane;   // 2 buffers x (hi, lo)
constexpr int kBarrierBytes = 512;
// HALO mode (3x3, Cin >= 64): the pixel tile is 16 rows x 8 columns and ONE (16+2) x (8+2) halo patch of the
// input (180 pixel rows of 128 B, SWIZZLE_128B) is loaded per 64-channel block and shared by all 9 taps: the A
// descriptor of tap (r,s) starts at halo row r*10+s and strides 10 rows (SBO = 1280 B) between its 8-row groups.
// (tests/experiments/umma_desc_probe.cu: with the descriptor's base-offset field 0 the MMA reads rows linearly
// from any 128-B-aligned start with any SBO, swizzle intact.)  Cuts the L2->smem traffic of A by 6.4x.
constexpr int kHaloTH = 16, kHaloTW = 8, kHaloW = kHaloTW + 2, kHaloH = kHaloTH + 2;
constexpr int kHaloTxBytes = kHaloW * kHaloH * 128;          // bytes one halo TMA box delivers (23,040)
constexpr int kHaloBytes = (kHaloTxBytes + 1023) / 1024 * 1024;   // slot size, keeps 1024-B alignment (23,552)

template <int BN, int BK, int CG = 1>
struct Cfg {
    static constexpr int ROW_BYTES = BK * 2;
    static constexpr int A_BYTES = kTileM * ROW_BYTES;
    static constexpr int B_BYTES = (BN / CG) * ROW_BYTES;     // a CTA of a pair holds half of the B tile
};

// CG = 1: one CTA per tile (M = 128).  CG = 2: a CTA pair (cluster of 2, cta_group::2) per 256-pixel tile --
// each CTA feeds its own 128 pixel rows of A and HALF of the weight tile, which halves the shared-memory
// operand traffic per MMA (the 1-CTA kernel is smem-bandwidth bound at N = 128).
// split-K GEMM mode helpers (ConvParams::n_parts > 1); an ordinary convolution has one part covering everything
__device__ __forceinline__ int part_k0(const ConvParams& p, int part) {
    return p.n_parts > 1 ? (part % p.splits) * p.kb_per_split : 0;
}
__device__ __forceinline__ int part_kblocks(const ConvParams& p, int part, int num_kb) {
    if (p.n_parts <= 1) return num_kb;
    const int k0 = (part % p.splits) * p.kb_per_split;
    return min(p.kb_per_split, p.kb_total - k0);
}
__device__ __forceinline__ int part_b_shift(const ConvParams& p, int part) {
    if (p.n_parts <= 1 || p.g_row_stride == 0) return 0;
    return ((part / p.splits) / 3 - 1) * p.g_row_stride;       // multiple of 8 elements: TMA box starts stay 16-B aligned
}
__device__ __forceinline__ int part_b_plane(const ConvParams& p, int part) {
    return (p.n_parts > 1 && p.g_row_stride != 0) ? (part / p.splits) % 3 : -1;
}

template <int CG>
__device__ __forceinline__ void tma_ld(void* s, const void* d, uint64_t* bar, int c0, int c1, int c2) {
    if constexpr (CG == 2) ptx::tma_load_3d_2sm(s, d, bar, c0, c1, c2); else ptx::tma_load_3d(s, d, bar, c0, c1, c2);
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 2) ptx::mma_f16_ss_2sm(d, a, b, idesc, acc); else ptx::mma_f16_ss(d, a, b, idesc, acc);
}
template <int CG>
__device__ __forceinline__ void mma_cm(uint64_t* bar) {
    if constexpr (CG == 2) ptx::mma_commit_2sm(bar); else ptx::mma_commit(bar);
}

__device__ __forceinline__ uint64_t make_halo_desc(uint32_t smem_addr) {   // SWIZZLE_128B, SBO = 10 rows
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((kHaloW * 128) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b, float& ra, float& rb) {
    // returns packed (bf16(a), bf16(b)); ra/rb = residuals a - hi, b - hi (exact in fp32)
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
    ra = a - __uint_as_float(u << 16);
    rb = b - __uint_as_float(u & 0xFFFF0000u);
    return u;
}

template <int BN, int BK, bool HALO, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const __grid_constant__ CUtensorMap tm_y_hi, const __grid_constant__ CUtensorMap tm_y_lo,
                 const ConvParams p) {
    using C = Cfg<BN, BK, CG>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-B alignment required by SWIZZLE_128B tiles.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    // non-HALO: one ring, a stage = A tile(s) + B tile(s).  HALO: an A ring of halo patches (one per
    // 64-channel block) followed by a B ring (one weight tile per tap).
    const int planes = p.x3 ? 2 : 1;
    const int a_stage_bytes = HALO ? planes * kHaloBytes : 0;
    const int AS = HALO ? p.a_stages : 0;
    const bool bres = !HALO && p.b_res != 0;
    const int bres_bytes = bres ? p.taps * p.cin_blocks * planes * C::B_BYTES : 0;       // resident weight tiles, k-block major
    const int stage_bytes = HALO ? planes * C::B_BYTES : (bres ? planes * C::A_BYTES : planes * (C::A_BYTES + C::B_BYTES));
    const int a_lo_off = bres ? C::A_BYTES : C::A_BYTES + C::B_BYTES;                     // A_lo inside a per-tap stage
    const int S = p.num_stages;
    uint8_t* bres_base = smem + (size_t)AS * a_stage_bytes;
    uint8_t* ring = bres_base + bres_bytes;
    uint8_t* staging = ring + (size_t)S * stage_bytes;                 // 1024-aligned (all slot sizes are multiples of 1 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + (p.store_bf16 ? kStagingBytes : 0));
    uint64_t* full_bar = bars;            // [S]
    uint64_t* empty_bar = bars + S;       // [S]
    uint64_t* tfull_bar = bars + 2 * S;   // [2]
    uint64_t* tempty_bar = bars + 2 * S + 2;  // [2]
    uint64_t* afull_bar = bars + 2 * S + 4;   // [AS]
    uint64_t* aempty_bar = afull_bar + AS;    // [AS]
    uint64_t* bres_bar = aempty_bar + AS;     // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;   // position in the CTA pair
    const bool leader = rank == 0;
    const int tile0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a_hi);
        ptx::prefetch_tensormap(&tm_b_hi);
        if (p.x3) {
            ptx::prefetch_tensormap(&tm_a_lo);
            ptx::prefetch_tensormap(&tm_b_lo);
        }
        if (p.store_bf16) ptx::prefetch_tensormap(&tm_y_hi);
        if (p.store_lo) ptx::prefetch_tensormap(&tm_y_lo);
        for (int i = 0; i < S; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 8 * CG);      // every epilogue warp of the pair arrives on the leader
        }
        for (int i = 0; i < AS; ++i) {
            ptx::mbar_init(&afull_bar[i], 1);
            ptx::mbar_init(&aempty_bar[i], 1);
        }
        ptx::mbar_init(bres_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            ptx::tmem_alloc_2sm(tmem_ptr, (uint32_t)p.tmem_cols);
            ptx::tmem_relinquish_2sm();
        } else {
            ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) ptx::cluster_sync();      // peer barriers initialised before any remote arrive / TMA signal
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // programmatic dependent launch: everything above overlapped the previous kernel's drain; from here on this kernel
    // reads the previous kernel's output (TMA loads, residual, m_valid) and overwrites buffers it may still be reading
    grid_dep_wait();

    const int num_kb = p.taps * p.cin_blocks;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            (void)as; (void)aphase;
            if (bres) {                  // one N tile: every weight tile of the layer, once, for all of this CTA's pixel tiles
                if (leader) ptx::mbar_arrive_expect_tx(bres_bar, (uint32_t)(bres_bytes * CG));
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
                    uint8_t* sb = bres_base + (size_t)kb * planes * C::B_BYTES;
                    tma_ld<CG>(sb, &tm_b_hi, bres_bar, cb * BK, (int)rank * (BN / CG), tap);
                    if (p.x3) tma_ld<CG>(sb + C::B_BYTES, &tm_b_lo, bres_bar, cb * BK, (int)rank * (BN / CG), tap);
                }
            }
            for (int tile = tile0; tile < p.num_tiles; tile += tile_step) {
                const int part = fd_div(tile, p.fd_tiles_per_part), t2 = tile - part * p.tiles_per_part;
                const int mq = fd_div(t2, p.fd_n_tiles), nt = t2 - mq * p.n_tiles;
                const int mt = mq * CG + (int)rank;      // an out-of-range tile of an odd pair loads zeros, stores nothing
                const int th_i = fd_div(mt, p.fd_tiles_w);
                const int h0 = th_i * p.TH;
                const int w0 = (mt - th_i * p.tiles_w) * p.TW;
                const int n0 = nt * BN + (int)rank * (BN / CG);        // this CTA's half of the weight tile
                if constexpr (HALO) {
                    for (int cb = 0; cb < p.cin_blocks; ++cb) {
                        ptx::mbar_wait(&aempty_bar[as], aphase ^ 1);
                        uint8_t* sa = smem + (size_t)as * a_stage_bytes;
                        if (leader) ptx::mbar_arrive_expect_tx(&afull_bar[as], (uint32_t)(planes * kHaloTxBytes * CG));
                        tma_ld<CG>(sa, &tm_a_hi, &afull_bar[as], cb * BK, w0 - 1, h0 - 1);
                        if (p.x3) tma_ld<CG>(sa + kHaloBytes, &tm_a_lo, &afull_bar[as], cb * BK, w0 - 1, h0 - 1);
                        if (++as == AS) { as = 0; aphase ^= 1; }
                        for (int tap = 0; tap < 9; ++tap) {
                            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* st = ring + (size_t)stage * stage_bytes;
                            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(stage_bytes * CG));
                            tma_ld<CG>(st, &tm_b_hi, &full_bar[stage], cb * BK, n0, tap);
                            if (p.x3) tma_ld<CG>(st + C::B_BYTES, &tm_b_lo, &full_bar[stage], cb * BK, n0, tap);
                            if (++stage == S) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
                const int nkb = part_kblocks(p, part, num_kb);
                const int ka = part_k0(p, part) * BK, kboff = ka + part_b_shift(p, part);
                const int bplane = part_b_plane(p, part);
                for (int kb = 0; kb < nkb; ++kb) {
                    const int tap = kb / p.cin_blocks;
                    const int btap = bplane >= 0 ? bplane : tap;
                    const int cb = kb - tap * p.cin_blocks;
                    const int r = tap / p.kw, s = tap - r * p.kw;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = ring + (size_t)stage * stage_bytes;
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(stage_bytes * CG));
                    tma_ld<CG>(st, &tm_a_hi, &full_bar[stage], ka + cb * BK, w0 + s - p.pad_w, h0 + r - p.pad_h);
                    if (!bres) tma_ld<CG>(st + C::A_BYTES, &tm_b_hi, &full_bar[stage], kboff + cb * BK, n0, btap);
                    if (p.x3) {
                        uint8_t* st2 = st + a_lo_off;
                        tma_ld<CG>(st2, &tm_a_lo, &full_bar[stage], ka + cb * BK, w0 + s - p.pad_w, h0 + r - p.pad_h);
                        if (!bres) tma_ld<CG>(st2 + C::A_BYTES, &tm_b_lo, &full_bar[stage], kboff + cb * BK, n0, btap);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        constexpr uint32_t idesc = ptx::make_idesc_f16(kTileM * CG, BN, /*bf16*/ 1);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        (void)as; (void)aphase;
        if (bres && leader) {
            ptx::mbar_wait(bres_bar, 0);                 // the resident weight tiles have landed
            ptx::tc_fence_after();
        }
        for (int tile = tile0; leader && tile < p.num_tiles; tile += tile_step) {
            ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
            const uint32_t d_corr = d_tmem + BN;   // bf16x3: lo*hi + hi*lo accumulate separately (see epilogue)
            if constexpr (HALO) {
                for (int cb = 0; cb < p.cin_blocks; ++cb) {
                    ptx::mbar_wait(&afull_bar[as], aphase);
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)as * a_stage_bytes);
                    for (int tap = 0; tap < 9; ++tap) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {
                            const uint32_t arow = sa + (uint32_t)(((tap / 3) * kHaloW + (tap % 3)) * 128);
                            const uint32_t st = ptx::smem_u32(ring + (size_t)stage * stage_bytes);
                            const uint64_t a_hi = make_halo_desc(arow);
                            const uint64_t b_hi = ptx::make_smem_desc(st, C::ROW_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                mma_ss<CG>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (cb | tap | k) != 0);
                            if (p.x3) {
                                const uint64_t a_lo = make_halo_desc(arow + kHaloBytes);
                                const uint64_t b_lo = ptx::make_smem_desc(st + C::B_BYTES, C::ROW_BYTES);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    mma_ss<CG>(d_corr, a_lo + 2 * k, b_hi + 2 * k, idesc, (cb | tap | k) != 0);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) mma_ss<CG>(d_corr, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
                            }
                            mma_cm<CG>(&empty_bar[stage]);                    // weight slot free
                            if (tap == 8) mma_cm<CG>(&aempty_bar[as]);         // halo patch free
                            if (tap == 8 && cb == p.cin_blocks - 1) mma_cm<CG>(&tfull_bar[acc]);
                        }
                        __syncwarp();
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                    if (++as == AS) { as = 0; aphase ^= 1; }
                }
                if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            const int nkb = part_kblocks(p, fd_div(tile, p.fd_tiles_per_part), num_kb);
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t st = ptx::smem_u32(ring + (size_t)stage * stage_bytes);
                    const uint32_t sbw = bres ? ptx::smem_u32(bres_base + (size_t)kb * planes * C::B_BYTES) : st + C::A_BYTES;
                    const uint64_t a_hi = ptx::make_smem_desc(st, C::ROW_BYTES);
                    const uint64_t b_hi = ptx::make_smem_desc(sbw, C::ROW_BYTES);
                    const int turn = kb / p.acc_chunk, ai = turn % p.nacc;
                    const uint32_t d_main = d_tmem + (uint32_t)(ai == 0 ? 0 : (ai + (p.x3 ? 1 : 0)) * BN);
                    const bool fresh = turn < p.nacc && kb == turn * p.acc_chunk;     // first k-block into this accumulator
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle span: +2 in (addr>>4)
                        mma_ss<CG>(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc, !(fresh && k == 0));
                    }
                    if (p.x3) {
                        const uint64_t a_lo = ptx::make_smem_desc(st + a_lo_off, C::ROW_BYTES);
                        const uint64_t b_lo = ptx::make_smem_desc(bres ? sbw + C::B_BYTES : st + 2 * C::A_BYTES + C::B_BYTES, C::ROW_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_ss<CG>(d_corr, a_lo + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_ss<CG>(d_corr, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
                    }
                    mma_cm<CG>(&empty_bar[stage]);            // frees the smem slot when the MMAs retire
                    if (kb == nkb - 1) mma_cm<CG>(&tfull_bar[acc]);   // accumulator complete
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================================ epilogue ================================
        // Eight epilogue warps in two groups of four: a group covers all 128 accumulator rows (TMEM lane group = warp % 4) and
        // takes every other 32-column chunk (group g: columns g*32, g*32 + 64, ...), with its own staging buffer, named barrier
        // and bulk-store issuer.  With one warp per scheduler the epilogue issued an instruction every 5.5 cycles (ncu on
        // conv1_1: no eligible warp 72 % of the time, profiles/r02_ncu_full_conv1_1_c8_details.txt): the N = 64 layers and the last
        // tile of every launch are bound by it.
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const int grp = (warp - 2) >> 2;         // epilogue group 0 / 1
        const int row = lg * 32 + lane;          // accumulator row == pixel within the tile
        const int row_h = row / p.TW, row_w = row - row_h * p.TW;     // its position inside the TH x TW patch (once per kernel)
        const int m_valid = p.m_valid ? *p.m_valid : 0x7fffffff;
        const bool issuer = (warp == 2 + 4 * grp && lane == 0);     // the group's thread that owns its bulk-store groups
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < p.num_tiles; tile += tile_step) {
            const int part = fd_div(tile, p.fd_tiles_per_part), t2 = tile - part * p.tiles_per_part;
            const int mq = fd_div(t2, p.fd_n_tiles), nt = t2 - mq * p.n_tiles;
            const int mt = mq * CG + (int)rank;
            const int th_i = fd_div(mt, p.fd_tiles_w);
            const int h0 = th_i * p.TH, w0 = (mt - th_i * p.tiles_w) * p.TW;
            const int h = h0 + row_h;
            const int w = w0 + row_w;
            const int n0 = nt * BN;
            const bool in_img = (h < p.H) && (w < p.W);
            const long pix = (long)h * p.W + w;
            const bool live = in_img && pix < m_valid;

            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            // TMEM -> registers, one 32-column chunk ahead: the loads of chunk c+1 (main and correction accumulator) are issued
            // as soon as chunk c has been copied out of the landing registers, so their latency hides behind chunk c's bias /
            // ReLU / split / staging work instead of standing at the head of every chunk (the N = 64 layers are epilogue-bound)
            uint32_t rm[32], rc[32];
            const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * p.acc_cols);
            if (n0 + grp * 32 < p.n_cover && grp * 32 < BN) {
                ptx::tmem_ld_32x32b_x32(tbase + (uint32_t)(grp * 32), rm);
                if (p.x3) ptx::tmem_ld_32x32b_x32(tbase + (uint32_t)(grp * 32) + BN, rc);
            }
#pragma unroll 1
            for (int c0 = grp * 32; c0 < BN; c0 += 64) {
                if (n0 + c0 >= p.n_cover) break;   // warp-uniform: nothing is stored past the covered columns
                uint32_t r[32];
                const uint32_t taddr = tbase + (uint32_t)c0;
                ptx::tmem_ld_wait();
                const int n = n0 + c0;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]);
                if (p.x3) {
                    // The tensor-core accumulator truncates on every add; keeping the ~2^-8-sized
                    // correction products in their own accumulator and adding them here with one
                    // round-to-nearest fp32 add keeps that bias at the single-pass level.
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rc[j]);
                }
                for (int a = 1; a < p.nacc; ++a) {            // rotated main accumulators of a long-K GEMM
                    ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)((a + (p.x3 ? 1 : 0)) * BN), r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
                }
                if (c0 + 64 < BN && n0 + c0 + 64 < p.n_cover) {      // this group's next chunk: its loads go out now
                    ptx::tmem_ld_32x32b_x32(taddr + 64, rm);
                    if (p.x3) ptx::tmem_ld_32x32b_x32(taddr + 64 + BN, rc);
                }
                if (p.bias != nullptr) {                    // split-K partial sums carry no bias (the reduction adds it once)
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = __ldg(b4 + j);
                        v[4 * j + 0] += b.x;
                        v[4 * j + 1] += b.y;
                        v[4 * j + 2] += b.z;
                        v[4 * j + 3] += b.w;
                    }
                }
                if (p.res_hi != nullptr && in_img && n < p.Cout) {
                    // residual add (h + shortcut) of a bottleneck block, fused ahead of the ReLU
                    const long ro = pix * p.Cout + n;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 rh = __ldg(reinterpret_cast<const uint4*>(p.res_hi + ro) + q);
                        const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&rh);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[8 * q + j] += __bfloat162float(hb[j]);
                        if (p.res_lo != nullptr) {
                            const uint4 rl = __ldg(reinterpret_cast<const uint4*>(p.res_lo + ro) + q);
                            const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&rl);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[8 * q + j] += __bfloat162float(lb[j]);
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                }
                if (!live) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.0f;
                }
                if (in_img && p.y_f32 != nullptr && n < p.ld_f32) {
                    float4* dst = reinterpret_cast<float4*>(p.y_f32 + (long)part * p.part_stride + pix * p.ld_f32 + n);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (p.store_bf16 && n < p.Cout) {          // CTA-uniform condition
                    int srow = row;                         // row of the staging tile this thread fills
                    bool writer = true;
                    if (p.pool) {
                        // F.MaxPooling2D(2,2) ceil mode fused: the tile is TH x TW pixels with TW = 16 or 8, a
                        // warp holds 32/TW consecutive tile rows, so the 2x2 window is lanes {l, l^1, l^TW, l^(TW+1)}.
                        // Out-of-image pixels were zeroed above and every valid value is >= 0 (ReLU), so the
                        // max over the valid part of a partial window is unchanged (Chainer cover_all=True).
                        const int tw_mask = p.TW;                // 16 or 8 (power of two)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                            v[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, tw_mask));
                        }
                        writer = (lane & (1 | tw_mask)) == 0;
                        // pooled tile is (TH/2) x (TW/2): row-major index of this thread's window
                        srow = (row_h >> 1) * (p.TW >> 1) + (row_w >> 1);
                    }
                    // stage64 (Cout % 64 == 0): the two groups' 32-channel chunks are the two halves of ONE staging tile with
                    // 128-byte rows (SWIZZLE_128B) and leave in one bulk store per plane: half as many (twice as long) rows for
                    // the TMA unit -- the N = 64 layers are bound by the rate at which it retires store rows.  Otherwise each
                    // group owns a staging tile of 64-byte rows (SWIZZLE_64B) and issues its own stores.
                    const bool s64 = p.stage64 != 0;
                    uint8_t* sb = s64 ? staging : staging + grp * (2 * kStagePlane);
                    const int bar_id = s64 ? 3 : 1 + grp, bar_n = s64 ? 256 : 128;
                    const bool my_store = s64 ? (warp == 2 && lane == 0) : issuer;
                    // the previous bulk store out of this buffer must be done READING it (it had this chunk's TMEM loads, bias,
                    // ReLU and split to finish in)
                    if (my_store) ptx::bulk_wait_group_read<0>();
                    ptx::named_bar_sync(bar_id, bar_n);
                    if (writer) {
                        if (s64) {
                            const int sw = srow & 7;               // SWIZZLE_128B: 16-B chunk index ^= address bits [7:9]
                            const uint32_t rowp = ptx::smem_u32(sb) + (uint32_t)(srow * 128);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float ra[8];
                                uint4 hv;
                                hv.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1], ra[0], ra[1]);
                                hv.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3], ra[2], ra[3]);
                                hv.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5], ra[4], ra[5]);
                                hv.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7], ra[6], ra[7]);
                                const int c16 = (grp * 4 + q) ^ sw;
                                ptx::st_shared_v4(rowp + (uint32_t)(c16 << 4), hv);
                                if (p.store_lo) {
                                    float d0, d1;
                                    uint4 lv;
                                    lv.x = pack_bf16x2(ra[0], ra[1], d0, d1);
                                    lv.y = pack_bf16x2(ra[2], ra[3], d0, d1);
                                    lv.z = pack_bf16x2(ra[4], ra[5], d0, d1);
                                    lv.w = pack_bf16x2(ra[6], ra[7], d0, d1);
                                    ptx::st_shared_v4(rowp + (uint32_t)(2 * kStagePlane + (c16 << 4)), lv);
                                }
                            }
                        } else {
                            const int sw = (srow >> 1) & 3;        // SWIZZLE_64B: 16-B chunk index ^= address bits [7:8]
                            const uint32_t rowp = ptx::smem_u32(sb) + (uint32_t)(srow * 64);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float ra[8];
                                uint4 hv;
                                hv.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1], ra[0], ra[1]);
                                hv.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3], ra[2], ra[3]);
                                hv.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5], ra[4], ra[5]);
                                hv.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7], ra[6], ra[7]);
                                ptx::st_shared_v4(rowp + (uint32_t)((q ^ sw) << 4), hv);
                                if (p.store_lo) {
                                    float d0, d1;
                                    uint4 lv;
                                    lv.x = pack_bf16x2(ra[0], ra[1], d0, d1);
                                    lv.y = pack_bf16x2(ra[2], ra[3], d0, d1);
                                    lv.z = pack_bf16x2(ra[4], ra[5], d0, d1);
                                    lv.w = pack_bf16x2(ra[6], ra[7], d0, d1);
                                    ptx::st_shared_v4(rowp + (uint32_t)(kStagePlane + ((q ^ sw) << 4)), lv);
                                }
                            }
                        }
                    }
                    ptx::fence_proxy_async_smem();             // generic-proxy smem writes -> visible to the TMA unit
                    ptx::named_bar_sync(bar_id, bar_n);
                    if (my_store) {
                        const int ow = p.pool ? (w0 >> 1) : w0, oh = p.pool ? (h0 >> 1) : h0;
                        const int nn = s64 ? n - grp * 32 : n;      // first channel of the stored block
                        ptx::tma_store_3d(&tm_y_hi, sb, nn, ow, oh);
                        if (p.store_lo) ptx::tma_store_3d(&tm_y_lo, sb + (s64 ? 2 * kStagePlane : kStagePlane), nn, ow, oh);
                        ptx::bulk_commit_group();
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) ptx::mbar_arrive_leader(&tempty_bar[acc]); else ptx::mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1; }
        }
        if (issuer) ptx::bulk_wait_group<0>();     // staging smem must outlive the last bulk store
    }

    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) ptx::cluster_sync();      // the peer may still signal our barriers / read our smem until here
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, (uint32_t)p.tmem_cols);
        else ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    // Resolved through the runtime so the library has no link-time dependency on libcuda.so
    // (it must dlopen on a CPU-only box for the symbol/ABI tests).
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 3-D bf16 tensor map: dims (innermost first) {d0,d1,d2}, row pitch d0 elements, box {b0,b1,b2};
// swizzle span = the box's inner extent in bytes (32 / 64 / 128).
// Encoded maps are cached per thread, keyed by (base, dims, box): an eager caller of frcnn_forward_vgg16 re-encodes
// nothing after its first image (six driver calls per conv launch otherwise); graph replay never comes here.
struct TmapKey {
    const void* base;
    uint64_t d0, d1, d2;
    uint32_t b0, b1, b2;
    bool operator==(const TmapKey& o) const {
        return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && b0 == o.b0 && b1 == o.b1 && b2 == o.b2;
    }
};
constexpr int kTmapCacheSlots = 512;
struct TmapCache {
    TmapKey key[kTmapCacheSlots];
    CUtensorMap map[kTmapCacheSlots];
    bool used[kTmapCacheSlots];
};
static TmapCache* tmap_cache() {
    static thread_local TmapCache* c = nullptr;
    if (c == nullptr) c = new TmapCache();          // value-initialised: used[] all false
    return c;
}

static int make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                        uint32_t b1, uint32_t b2, uint64_t stride1_bytes = 0, uint64_t stride2_bytes = 0) {
    // stride1_bytes / stride2_bytes != 0: explicit byte strides of dimensions 1 and 2 (a SLIDING-WINDOW map when stride1 is
    // smaller than the dimension-0 extent: neighbouring "rows" overlap in memory -- verified on B200 by
    // tests/experiments/tma_overlap_probe.cu); the cache key folds them into d1 / d2's upper bits
    const TmapKey key{base, d0, d1 | (stride1_bytes << 32), d2 | (stride2_bytes << 32), b0, b1, b2};
    uint64_t hsh = reinterpret_cast<uintptr_t>(base) * 0x9E3779B97F4A7C15ull;
    hsh ^= (d0 * 31 + d1) * 0xC2B2AE3D27D4EB4Full + d2 * 0x165667B19E3779F9ull + b0 * 131 + b1 * 17 + b2;
    const int slot = (int)((hsh >> 20) % kTmapCacheSlots);
    TmapCache* cache = tmap_cache();
    if (cache->used[slot] && cache->key[slot] == key) {
        *tm = cache->map[slot];
        return FRCNN_OK;
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return FRCNN_ERR_CUDA;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes ? stride1_bytes : d0 * 2, stride2_bytes ? stride2_bytes : d0 * d1 * 2};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = (b0 * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (b0 * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims {%llu,%llu,%llu} box {%u,%u,%u} base %p", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2, base);
        return FRCNN_ERR_CUDA;
    }
    cache->key[slot] = key;
    cache->map[slot] = *tm;
    cache->used[slot] = true;
    return FRCNN_OK;
}

// Tuning overrides (frcnn_conv2d_set_*): per calling thread, read when a launch is enqueued (and therefore fixed inside
// a captured graph) -- two engines driven from different threads cannot disturb each other.
static int env_int(const char* name) {
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}
// FRCNN_CONV_MAX_CTAS: a process-wide default for frcnn_conv2d_set_max_ctas (each thread starts from it)
static thread_local int g_force_bn = 0, g_force_th = 0, g_force_tw = 0, g_force_cg = 0, g_max_ctas = env_int("FRCNN_CONV_MAX_CTAS"),
                        g_smem_reserve = 0;

static int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int BN, int BK, bool HALO, int CG>
static int launch_conv(const CUtensorMap* tm, ConvParams p, cudaStream_t stream) {
    using C = Cfg<BN, BK, CG>;
    p.fd_tiles_w = make_fastdiv(p.tiles_w);
    p.fd_n_tiles = make_fastdiv(p.n_tiles);
    p.fd_tiles_per_part = make_fastdiv(p.tiles_per_part);
    const int planes = p.x3 ? 2 : 1;
    const bool bres = !HALO && p.b_res != 0;
    const int bres_bytes = bres ? p.taps * p.cin_blocks * planes * C::B_BYTES : 0;
    const int stage_bytes = HALO ? planes * C::B_BYTES : (bres ? planes * C::A_BYTES : planes * (C::A_BYTES + C::B_BYTES));
    const int fixed = 1024 /*align slack*/ + kBarrierBytes + (p.store_bf16 ? kStagingBytes : 0) + bres_bytes;
    // shared-memory budget of the persistent CTA: everything by default; frcnn_conv2d_set_smem_reserve leaves a slice of
    // every SM to other kernels, so that the small kernels of ANOTHER image in flight (decode, NMS, RoI pooling, a host
    // caller's per-class NMS) can become resident next to a convolution instead of waiting for one of its CTAs to retire
    const int smem_budget = 227 * 1024 - g_smem_reserve;
    int a_stages = 0, a_bytes = 0;
    if (HALO) {
        // two halo slots when at least 3 weight slots still fit, else one
        a_stages = ((smem_budget - fixed - 2 * planes * kHaloBytes) / stage_bytes >= 3) ? 2 : 1;
        a_bytes = a_stages * planes * kHaloBytes;
    }
    int stages = (smem_budget - fixed - a_bytes) / stage_bytes;
    if (stages > 20) stages = 20;
    if (stages < 2) {
        set_error("conv tile BN=%d BK=%d halo=%d does not fit 2 pipeline stages", BN, BK, (int)HALO);
        return FRCNN_ERR_ARG;
    }
    p.num_stages = stages;
    p.a_stages = a_stages;
    const size_t smem = (size_t)stages * stage_bytes + a_bytes + fixed;
    p.acc_cols = p.x3 ? 2 * BN : BN;
    // long K (>= 256 k-blocks of 64 on the per-tap path, or a split-K GEMM that asks for it): rotate over up to 3 main accumulators
    p.acc_chunk = 1 << 30;
    if (!HALO && ((p.n_parts == 1 && p.taps * p.cin_blocks >= 256) || (p.n_parts > 1 && p.nacc > 1))) {
        int nacc = p.n_parts > 1 ? p.nacc : 3;
        while (nacc > 1 && (p.acc_cols + (nacc - 1) * BN > 512)) --nacc;
        if (nacc > 1) {
            p.nacc = nacc;
            p.acc_chunk = 8;
            p.acc_cols += (nacc - 1) * BN;
        } else {
            p.nacc = 1;
        }
    } else {
        p.nacc = 1;
    }
    p.acc_bufs = (2 * p.acc_cols <= 512) ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.acc_bufs * p.acc_cols) p.tmem_cols *= 2;
    auto kern = conv_gemm_kernel<BN, BK, HALO, CG>;
    FRCNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas = device_sm_count();
    if (g_max_ctas > 0 && g_max_ctas < ctas) ctas = g_max_ctas;      // several images in flight: each launch takes a share of the SMs
    const int units = ctas / CG > 0 ? ctas / CG : 1;     // CTAs (CG = 1) or CTA pairs (CG = 2) resident at once
    const int grid = CG * (p.num_tiles < units ? p.num_tiles : units);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CG;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    FRCNN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p));
    FRCNN_LAUNCH_OK();
    return FRCNN_OK;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" void frcnn_conv2d_set_cta_group(int cta_group) { g_force_cg = cta_group; }

extern "C" void frcnn_conv2d_set_max_ctas(int max_ctas) { g_max_ctas = max_ctas; }

extern "C" void frcnn_conv2d_set_smem_reserve(int bytes) { g_smem_reserve = bytes < 0 ? 0 : (bytes > 96 * 1024 ? 96 * 1024 : bytes); }

extern "C" void frcnn_conv2d_set_tile(int block_n, int tile_h, int tile_w) {
    g_force_bn = block_n;
    g_force_th = tile_h;
    g_force_tw = tile_w;
}

namespace {
struct GemmExtra {          // split-K GEMM mode of the same kernel (frcnn_gemm_nt_splitk)
    int groups, row_stride, splits;
    long part_stride;
    int bn = 0;             // 0: the weight-gradient rule (128 / 64); else the N tile to use (64, 128, 160, 256)
    int nacc = 1;           // > 1: every split rotates its k-blocks over up to this many main accumulators (long K)
};
struct ResExtra {           // residual input of frcnn_conv2d_res
    const void *hi, *lo;
};
struct WinExtra {           // frcnn_conv3x3_c8: 3x3 convolution over a compact [H][W+2][8] image through a sliding-window map
    int row_pixels;         // pixels per stored row (W + 2: one zero column on each side)
};
}  // namespace

static int conv2d_impl(const void* x_hi, const void* x_lo, int H, int W, int Cin, const void* w_hi,
                       const void* w_lo, const float* bias, int Cout, int ksize, int relu, int fuse_pool2x2,
                       void* y_hi, void* y_lo, float* y_f32, int ld_f32, const int* m_valid, void* stream_,
                       const GemmExtra* ge, const ResExtra* re = nullptr, const WinExtra* we = nullptr) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    FRCNN_REQUIRE(x_hi && w_hi && (bias || ge), "frcnn_conv2d: x_hi, w_hi and bias are required");
    FRCNN_REQUIRE((x_lo == nullptr) == (w_lo == nullptr), "frcnn_conv2d: x_lo and w_lo must both be given (bf16x3) or both NULL");
    FRCNN_REQUIRE(ksize == 1 || ksize == 3, "frcnn_conv2d: ksize must be 1 or 3 (got %d)", ksize);
    FRCNN_REQUIRE(H > 0 && W > 0 && Cin > 0 && Cout > 0, "frcnn_conv2d: bad shape H=%d W=%d Cin=%d Cout=%d", H, W, Cin, Cout);
    FRCNN_REQUIRE(Cin % 8 == 0, "frcnn_conv2d: Cin must be a multiple of 8 (16-byte TMA rows), got %d", Cin);
    FRCNN_REQUIRE(y_hi || y_f32, "frcnn_conv2d: no output requested");
    FRCNN_REQUIRE(!y_hi || Cout % 32 == 0, "frcnn_conv2d: bf16 output needs Cout %% 32 == 0 (got %d)", Cout);
    FRCNN_REQUIRE(!y_f32 || (ld_f32 % 32 == 0 && ld_f32 >= Cout), "frcnn_conv2d: ld_f32 must be a multiple of 32 and >= Cout");
    FRCNN_REQUIRE(!y_lo || y_hi, "frcnn_conv2d: y_lo without y_hi");
    FRCNN_REQUIRE(!fuse_pool2x2 || (y_hi && relu && !y_f32 && !m_valid),
                  "frcnn_conv2d: fuse_pool2x2 needs a bf16 output, relu=1, no fp32 output and no m_valid");

    const int BK = (Cin >= 64) ? 64 : (Cin >= 32 ? 32 : 16);
    FRCNN_REQUIRE(BK != 32 || Cin == 32, "frcnn_conv2d: Cin in (32,64) is not supported (use 16, 32 or >= 64)");

    // ---- pixel tile: minimise padded pixels (the fused pool needs the 8x16 tile: 2x2 windows inside a warp)
    static const int shapes[6][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}, {32, 4}};
    int TH = 8, TW = 16;
    const bool forced_tile = g_force_th > 0 && g_force_tw > 0 && g_force_th * g_force_tw == kTileM;
    // HALO (shared halo patch for the 9 taps) needs the 16x8 tile; a forced other tile selects the per-tap path
    const bool halo = ksize == 3 && BK == 64 && !we && (!forced_tile || (g_force_th == kHaloTH && g_force_tw == kHaloTW));
    if (halo) {
        TH = kHaloTH;
        TW = kHaloTW;
    } else if (fuse_pool2x2) {
        TH = 8;
        TW = 16;
    } else if (forced_tile) {
        TH = g_force_th;
        TW = g_force_tw;
    } else {
        long best = -1;
        for (auto& s : shapes) {
            long t = (long)cdiv(H, s[0]) * cdiv(W, s[1]);
            if (best < 0 || t < best) { best = t; TH = s[0]; TW = s[1]; }
        }
    }
    const int tiles_h = cdiv(H, TH), tiles_w = cdiv(W, TW);
    const long m_tiles = (long)tiles_h * tiles_w;

    // ---- N tile: minimise waves x measured per-k-block cost
    const int cout_cover = y_f32 ? (ld_f32 > Cout ? ld_f32 : Cout) : Cout;
    int BN = 0;
    if (g_force_bn == 64 || g_force_bn == 128 || g_force_bn == 160 || g_force_bn == 256) {
        BN = g_force_bn;
    } else if (ge != nullptr && ge->bn != 0) {
        BN = ge->bn;
    } else if (ge != nullptr) {
        // split-K GEMM mode has tiles to spare (groups x splits): N = 128 halves the re-reads of A and is the measured
        // optimum on every VGG weight-gradient shape (tests/gpu_wgrad_tune.py); N = 256 loses the accumulator double buffer
        BN = cout_cover >= 128 ? 128 : 64;
    } else {
        const int sms = device_sm_count();
        double best = 0;
        const int cand[3] = {256, 128, 64};
        // measured cycles per k-block per CTA on B200 (tests/gpu_tune_conv.py): the main loop is
        // L2->smem bandwidth bound, not MMA bound, so the cost is NOT proportional to BN.
        const double cost_bf16[3] = {880, 770, 700}, cost_x3[3] = {2230, 1190, 1030};
        const double* cost = (x_lo != nullptr) ? cost_x3 : cost_bf16;
        for (int i = 0; i < 3; ++i) {
            if (cand[i] > 64 && cand[i] > cout_cover && cand[i] / 2 >= cout_cover) continue;  // too wide
            long tiles = m_tiles * cdiv(cout_cover, cand[i]);
            double t = (double)cdiv((int)tiles, sms) * cost[i];
            if (BN == 0 || t < best) { best = t; BN = cand[i]; }
        }
    }

    ConvParams p;
    p.H = H; p.W = W; p.Cout = Cout;
    p.ksize = ksize; p.taps = ksize * ksize; p.cin_blocks = cdiv(Cin, BK);
    p.kw = ksize; p.pad_h = p.pad_w = (ksize - 1) / 2;
    if (we != nullptr) {
        // K = 3 image rows x (4 pixels x 8 channels): one k-block of 32 per row r, its A box starting at stored pixel w0 of row
        // h0 + r - 1 (the left zero column is stored, rows -1 and H are the TMA unit's out-of-bounds zero fill)
        FRCNN_REQUIRE(Cin == 32 && ksize == 3 && !fuse_pool2x2 && !m_valid && !ge && !re, "conv3x3_c8: bad configuration");
        p.taps = 3; p.kw = 1; p.pad_h = 1; p.pad_w = 0;
    }
    p.TH = TH; p.TW = TW; p.tiles_h = tiles_h; p.tiles_w = tiles_w;
    p.n_tiles = cdiv(cout_cover, BN);
    // CTA pairs (cta_group::2) for the wide-N tiles of layers with at least two pixel tiles
    int CG = (BK == 64 && m_tiles >= 2) ? 2 : 1;      // measured: pairs win at every N (conv1_2, N = 64: -27 %)
    if (g_force_cg == 1) CG = 1;
    if (g_force_cg == 2 && BK == 64) CG = 2;
    FRCNN_REQUIRE(m_tiles * p.n_tiles < (1l << 30), "frcnn_conv2d: too many tiles");
    p.num_tiles = (int)(cdiv((int)m_tiles, CG) * p.n_tiles);     // tiles (CG = 1) or pair-tiles (CG = 2)
    p.tiles_per_part = p.num_tiles;
    p.n_parts = 1; p.splits = 1; p.kb_per_split = p.kb_total = p.taps * p.cin_blocks; p.g_row_stride = 0; p.part_stride = 0;
    if (ge != nullptr) {
        FRCNN_REQUIRE(ksize == 1 && BK == 64 && Cin % 64 == 0 && y_f32 && !y_hi && !m_valid, "gemm_nt_splitk: bad configuration");
        p.kb_total = p.cin_blocks;
        p.kb_per_split = cdiv(p.kb_total, ge->splits);
        p.splits = cdiv(p.kb_total, p.kb_per_split);             // every split non-empty
        p.n_parts = ge->groups * p.splits;
        p.g_row_stride = ge->groups == 9 ? ge->row_stride : 0;
        FRCNN_REQUIRE(ge->groups == 1 || (ge->row_stride > 0 && ge->row_stride % 8 == 0),
                      "gemm_nt_splitk: row_stride must be a positive multiple of 8 (16-byte aligned TMA box starts)");
        p.part_stride = ge->part_stride;
        FRCNN_REQUIRE((long)p.num_tiles * p.n_parts < (1l << 30), "gemm_nt_splitk: too many tiles");
        p.num_tiles *= p.n_parts;
    }
    p.num_stages = 0;
    p.nacc = ge != nullptr ? ge->nacc : 1;
    p.x3 = x_lo != nullptr;
    p.relu = relu;
    p.pool = fuse_pool2x2 ? 1 : 0;
    p.ld_f32 = ld_f32;
    p.n_cover = cdiv(cout_cover, 32) * 32;
    p.store_bf16 = y_hi != nullptr;
    p.store_lo = y_lo != nullptr;
    p.stage64 = (y_hi != nullptr && y_f32 == nullptr && Cout % 64 == 0) ? 1 : 0;
    // resident weights: the compact first layer (3 k-blocks, one N tile): 24 KB fetched once per CTA instead of once per tile
    p.b_res = (we != nullptr && p.n_tiles == 1 && ge == nullptr && p.taps * p.cin_blocks <= 16) ? 1 : 0;
    p.y_f32 = y_f32;
    p.bias = bias;
    p.m_valid = m_valid;
    p.res_hi = re ? (const __nv_bfloat16*)re->hi : nullptr;
    p.res_lo = re ? (const __nv_bfloat16*)re->lo : nullptr;
    FRCNN_REQUIRE(!re || (re->hi && !fuse_pool2x2 && Cout % 32 == 0), "frcnn_conv2d_res: residual needs res_hi, Cout %% 32 == 0 and no fused pool");
    p.acc_bufs = p.acc_cols = p.tmem_cols = 0;

    CUtensorMap tm[6];
    int rc;
    const int abw = halo ? kHaloW : TW, abh = halo ? kHaloH : TH;     // A box: the tile, or the tile + 1-pixel halo
    const uint64_t as1 = we ? 16 : 0, as2 = we ? (uint64_t)we->row_pixels * 16 : 0;     // sliding window: one pixel (8 ch) per step
    if ((rc = make_tmap_3d(&tm[0], x_hi, Cin, W, H, BK, abw, abh, as1, as2)) != FRCNN_OK) return rc;
    const int b_planes = (ge != nullptr && ge->groups == 9) ? 3 : p.taps;
    if ((rc = make_tmap_3d(&tm[2], w_hi, Cin, Cout, b_planes, BK, BN / CG, 1)) != FRCNN_OK) return rc;
    if (p.x3) {
        if ((rc = make_tmap_3d(&tm[1], x_lo, Cin, W, H, BK, abw, abh, as1, as2)) != FRCNN_OK) return rc;
        if ((rc = make_tmap_3d(&tm[3], w_lo, Cin, Cout, b_planes, BK, BN / CG, 1)) != FRCNN_OK) return rc;
    } else {
        tm[1] = tm[0];
        tm[3] = tm[2];
    }
    if (y_hi) {
        const int Ho = p.pool ? (H + 1) / 2 : H, Wo = p.pool ? (W + 1) / 2 : W;
        const int bw = p.pool ? TW / 2 : TW, bh = p.pool ? TH / 2 : TH;
        const int sc = p.stage64 ? 64 : 32;             // channels per store row: 128-byte (SWIZZLE_128B) or 64-byte rows
        if ((rc = make_tmap_3d(&tm[4], y_hi, Cout, Wo, Ho, sc, bw, bh)) != FRCNN_OK) return rc;
        if (y_lo) {
            if ((rc = make_tmap_3d(&tm[5], y_lo, Cout, Wo, Ho, sc, bw, bh)) != FRCNN_OK) return rc;
        } else {
            tm[5] = tm[4];
        }
    } else {
        tm[4] = tm[0];
        tm[5] = tm[0];
    }

#define FRCNN_DISPATCH(BN_, BK_, HALO_, CG_) \
    if (BN == BN_ && BK == BK_ && halo == HALO_ && CG == CG_) return launch_conv<BN_, BK_, HALO_, CG_>(tm, p, stream);
    FRCNN_DISPATCH(256, 64, true, 2)
    FRCNN_DISPATCH(128, 64, true, 2)
    FRCNN_DISPATCH(256, 64, false, 2)
    FRCNN_DISPATCH(128, 64, false, 2)
    FRCNN_DISPATCH(64, 64, true, 2)
    FRCNN_DISPATCH(64, 64, false, 2)
    FRCNN_DISPATCH(160, 64, false, 2)
    FRCNN_DISPATCH(160, 64, false, 1)
    FRCNN_DISPATCH(256, 64, true, 1)
    FRCNN_DISPATCH(128, 64, true, 1)
    FRCNN_DISPATCH(64, 64, true, 1)
    FRCNN_DISPATCH(256, 64, false, 1)
    FRCNN_DISPATCH(128, 64, false, 1)
    FRCNN_DISPATCH(64, 64, false, 1)
    FRCNN_DISPATCH(128, 32, false, 1)
    FRCNN_DISPATCH(64, 32, false, 1)
    FRCNN_DISPATCH(256, 16, false, 1)
    FRCNN_DISPATCH(128, 16, false, 1)
    FRCNN_DISPATCH(64, 16, false, 1)
#undef FRCNN_DISPATCH
    set_error("frcnn_conv2d: no kernel for BN=%d BK=%d", BN, BK);
    return FRCNN_ERR_ARG;
}

extern "C" int frcnn_conv2d(const void* x_hi, const void* x_lo, int H, int W, int Cin, const void* w_hi,
                            const void* w_lo, const float* bias, int Cout, int ksize, int relu, int fuse_pool2x2,
                            void* y_hi, void* y_lo, float* y_f32, int ld_f32, const int* m_valid, void* stream_) {
    FRCNN_ENTRY();
    return conv2d_impl(x_hi, x_lo, H, W, Cin, w_hi, w_lo, bias, Cout, ksize, relu, fuse_pool2x2, y_hi, y_lo, y_f32, ld_f32,
                       m_valid, stream_, nullptr);
}

extern "C" int frcnn_conv2d_res(const void* x_hi, const void* x_lo, int H, int W, int Cin, const void* w_hi, const void* w_lo,
                                const float* bias, int Cout, int ksize, int relu, const void* res_hi, const void* res_lo,
                                void* y_hi, void* y_lo, void* stream_) {
    FRCNN_ENTRY();
    ResExtra re{res_hi, res_lo};
    return conv2d_impl(x_hi, x_lo, H, W, Cin, w_hi, w_lo, bias, Cout, ksize, relu, 0, y_hi, y_lo, nullptr, 0, nullptr, stream_,
                       nullptr, &re);
}

extern "C" int frcnn_conv3x3_c8(const void* x_hi, const void* x_lo, int H, int W, const void* w_hi, const void* w_lo,
                                const float* bias, int Cout, int relu, void* y_hi, void* y_lo, void* stream_) {
    FRCNN_ENTRY();
    WinExtra we{W + 2};
    return conv2d_impl(x_hi, x_lo, H, W, 32, w_hi, w_lo, bias, Cout, 3, relu, 0, y_hi, y_lo, nullptr, 0, nullptr, stream_, nullptr,
                       nullptr, &we);
}

extern "C" int frcnn_gemm_nt_splitk_splits(int K, int splits) {
    FRCNN_ENTRY();
    const int kb = cdiv(K, 64), per = cdiv(kb, splits < 1 ? 1 : splits);
    return cdiv(kb, per);
}

extern "C" int frcnn_gemm_nt_splitk(const void* a_hi, const void* a_lo, int M, int K, const void* b_hi, const void* b_lo,
                                    int N, int groups, int row_stride, int splits, const float* zero_bias, float* parts,
                                    int ld, void* stream_) {
    FRCNN_ENTRY();
    FRCNN_REQUIRE(groups == 1 || groups == 9, "gemm_nt_splitk: groups must be 1 or 9 (got %d)", groups);
    FRCNN_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0, "gemm_nt_splitk: bad shape M=%d N=%d K=%d (K %% 64 == 0)", M, N, K);
    FRCNN_REQUIRE(splits >= 1 && parts && zero_bias, "gemm_nt_splitk: splits >= 1, parts and a zero bias vector of ld floats are required");
    GemmExtra ge{groups, row_stride, splits, (long)M * ld};
    return conv2d_impl(a_hi, a_lo, 1, M, K, b_hi, b_lo, zero_bias, N, 1, 0, 0, nullptr, nullptr, parts, ld, nullptr, stream_, &ge);
}

// Internal entry of linear_swapab.cu: parts[split][M][ld] = A[M,K] . B[N,K]^T over that split's K range, no bias, N tile
// `bn`, each split rotating over `nacc` main accumulators.  Returns the effective number of splits through *splits_out.
namespace frcnn {
int gemm_nt_splitk_parts(const void* a_hi, const void* a_lo, int M, int K, const void* b_hi, const void* b_lo, int N, int splits,
                         int bn, int nacc, float* parts, int ld, int* splits_out, void* stream_) {
    FRCNN_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && splits >= 1 && parts, "gemm_nt_splitk_parts: bad arguments");
    GemmExtra ge{1, 0, splits, (long)M * ld};
    ge.bn = bn;
    ge.nacc = nacc;
    if (splits_out) *splits_out = frcnn_gemm_nt_splitk_splits(K, splits);
    return conv2d_impl(a_hi, a_lo, 1, M, K, b_hi, b_lo, nullptr, N, 1, 0, 0, nullptr, nullptr, parts, ld, nullptr, stream_, &ge);
}
}  // namespace frcnn
