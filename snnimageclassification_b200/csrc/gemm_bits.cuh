// gemm_bits.cuh -- the two contractions of the path fed from BIT-PACKED spike rasters.
//
//   k_proj_bits    I_in[r][n]   = sum_k x[r][k] W_in[k][n]       (K1)   x, Z given as (rows, ceil(K/32)) uint32 words,
//   k_wgrad_bits   dW_in[m][n]  = sum_r x[r][m] gI[r][n]         (K4)   bit l of word w = element 32 w + l
//                  dW_rec[j][n] = sum_r Z_{t-1}[r][j] gI[r][n]
//
// The reference builds the fp32 raster on the host (src/datasets/datasets.py:93-97) and multiplies it at every step
// (src/modules/spiking_layers.py:163/233); the fp32 kernels of gemm_tc.cuh read it once per contraction.  A raster is
// one bit of information per element, so here the spike operand never exists as fp32 in HBM or L2: every CTA expands
// its words into the shared-memory tile the tensor pipe reads (generic-proxy stores + fence.proxy.async), which removes
// 32/33 of the operand bytes of a tile row and lets one weight stage serve TWO row tiles (the accumulators of both
// fit TMEM), halving the L2 -> SM weight traffic that bounded the fp32 tiles.
//
// Numerics.  k_proj_bits: kind::f16 -- spikes {0,1} are exact in fp16; every column n of W_in is scaled by a power of
// two 2^s(n) so that its largest element lies in [2^14, 2^15) and split into two fp16 planes hi = rn16(w 2^s),
// lo = rn16(w 2^s - hi): |w 2^s - hi - lo| <= 2^-22 |w 2^s| (or 2^-25, fp16's subnormal spacing, for elements 2^17
// times smaller than the column's largest) -- the same 22 bits as the two tf32 planes of k_proj_tc.  Products are exact,
// accumulation is fp32 in TMEM, the epilogue adds the two column groups and multiplies by 2^-s(n) (exact).
// k_wgrad_bits: kind::tf32 with the exact two-plane tf32 split of gI, as k_wgrad_tc.
#pragma once
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace snnk {
namespace tc {

constexpr int kBitsBlockK = 64;   // fp16 elements per k-block = one 128-byte swizzle row = two raster words
constexpr int kUmmaKf16 = 16;     // K per tcgen05.mma for kind::f16

// Instruction descriptor for kind::f16: D fp32, A and B fp16, both K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n)
{
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Threads of the bit-fed kernels: warp 0 producer, warp 1 MMA issue, then four expander / epilogue warps PER row tile
// (one warp per scheduler could not expand a tile in the time the tensor pipe needs for it; ncu, profiles/r02_*).
template <int MT> constexpr int bits_threads() { return 64 + MT * 128; }

// A spike enters the fp16 tile as 2.0 (0x4000: ONE set bit, so two raster bits become two halves with a multiply and a
// mask); the factor 2 is divided out, exactly, together with the column scale in the epilogue.
constexpr float kSpikeHalfValue = 2.0f;
__device__ __forceinline__ uint32_t bits2_to_half2(uint32_t t)      // t = b0 + 2 b1  ->  b0 << 14 | b1 << 30
{
    return (t * 0x20004000u) & 0x40004000u;
}

// ---- weight planes ---------------------------------------------------------------------------------------------------
// W_in (K,H) fp32 -> per-column power-of-two scale + two fp16 planes of the scaled transpose, stored as the K-major
// SWIZZLE_128B tiles the tensor pipe reads: planes[p][kb][h][64 halves], the eight 16-byte chunks of a row permuted by
// (chunk ^ (h & 7)), so the B operand of one k-block and n-tile is ONE contiguous range (one bulk copy per plane).
// One CTA per hidden unit: max |w| over the column, then the split.
__global__ void __launch_bounds__(256) k_split_w_h(const float* __restrict__ W, int K, int H, int kblocks,
                                                  __half* __restrict__ planes, float* __restrict__ inv_scale)
{
    const int h = blockIdx.x;
    __shared__ float s_max[8];
    float m = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmaxf(m, fabsf(W[(size_t)k * H + h]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_max[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) m = fmaxf(m, s_max[q]);
    int s = 0;
    if (m > 0.f && m < 3.0e38f) {
        int e;
        frexpf(m, &e);              // m = f 2^e, f in [0.5, 1)
        s = 15 - e;                 // m 2^s in [2^14, 2^15)
        s = s > 120 ? 120 : (s < -120 ? -120 : s);
    }
    if (threadIdx.x == 0) inv_scale[h] = ldexpf(1.0f, -s) / kSpikeHalfValue;   // exact: both are powers of two
    const size_t plane = (size_t)kblocks * H * kBitsBlockK;
    for (int k = threadIdx.x; k < kblocks * kBitsBlockK; k += blockDim.x) {
        const float w = k < K ? ldexpf(W[(size_t)k * H + h], s) : 0.f;      // exact scaling
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));            // the subtraction is exact
        const int kb = k / kBitsBlockK, kk = k - kb * kBitsBlockK;
        const size_t o = ((size_t)kb * H + h) * kBitsBlockK + (size_t)((((kk >> 3) ^ (h & 7)) << 3) | (kk & 7));
        planes[o] = hi;
        planes[plane + o] = lo;
    }
}

// ---- K1 from bits -----------------------------------------------------------------------------------------------------
template <int H, int MT>
struct ProjBitsCfg {
    static constexpr int P = 2;
    static constexpr uint32_t kABytes = kBlockM * 128;            // one row tile of one k-block: 128 rows x 64 fp16
    static constexpr uint32_t kBBytes = P * H * 128;              // both planes of the n-tile
    static constexpr uint32_t kStageBytes = MT * kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
    static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 256;
    static constexpr uint32_t kColsPerTile = tmem_cols_for(P * H);
    static constexpr uint32_t kTmemCols = MT * kColsPerTile;
    static_assert(P * H <= 256, "UMMA N");
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation");
};

// bits : (M, wd) uint32 raster words.  Warp 0: bulk copies of the weight planes; warp 1: TMEM + MMA issue; warps 2-5:
// expansion of the raster words into the A tiles (thread = row), then the epilogue.
template <int H, int MT>
__global__ void __launch_bounds__(bits_threads<MT>(), 1)
k_proj_bits(const uint32_t* __restrict__ bits, int wd, const __half* __restrict__ planes, const float* __restrict__ inv_scale,
            float* __restrict__ C, int M, int kblocks, int ldc)
{
    using Cfg = ProjBitsCfg<H, MT>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * Cfg::kStageBytes);
    uint64_t* full_b = bars;                   // [kStages]  bulk copies -> MMA
    uint64_t* full_a = bars + kStages;         // [kStages]  4 expander warps -> MMA
    uint64_t* empty = bars + 2 * kStages;      // [kStages]  MMA commit -> producer and expanders
    uint64_t* tmem_full = bars + 3 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * (MT * kBlockM);
    const int n0 = blockIdx.y * H;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // see k_proj_tc

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_b + s, 1); mbar_init(full_a + s, 4 * MT); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(empty + s, ph ^ 1);
                unsigned char* st = smem + (size_t)s * Cfg::kStageBytes + MT * Cfg::kABytes;
                mbar_expect_tx(full_b + s, Cfg::kBBytes);
#pragma unroll
                for (int p = 0; p < Cfg::P; ++p)
                    bulk_g2s(st + p * (H * 128), planes + ((size_t)p * kblocks + kb) * ((size_t)ldc * kBitsBlockK) + (size_t)n0 * kBitsBlockK,
                             H * 128, full_b + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(Cfg::P * H);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full_b + s, ph);
                mbar_wait(full_a + s, ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                const uint32_t b0 = a0 + MT * Cfg::kABytes;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int kk = 0; kk < kBitsBlockK / kUmmaKf16; ++kk) {
                        const uint64_t adesc = make_smem_desc(a0 + mt * Cfg::kABytes + kk * kUmmaKf16 * 2, 16, 1024);
                        const uint64_t bdesc = make_smem_desc(b0 + kk * kUmmaKf16 * 2, 16, 1024);
                        umma_f16(tmem_base + mt * Cfg::kColsPerTile, adesc, bdesc, idesc, (kb | kk) != 0);
                    }
                umma_commit(empty + s);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;                    // TMEM lane quarter of this warp = its rows of its tile
        const int mt = (warp - 2) >> 2;            // warps 2-5: row tile 0, warps 6-9: row tile 1
        const int trow = 32 * q + lane;
        const int row = m0 + mt * kBlockM + trow;
        const bool rvalid = row < M;
        const uint32_t* rowp = bits + (size_t)(rvalid ? row : 0) * wd;
        uint32_t w0 = (rvalid && 0 < wd) ? __ldg(rowp) : 0u;
        uint32_t w1 = (rvalid && 1 < wd) ? __ldg(rowp + 1) : 0u;
        const uint32_t dst0 = smem_u32(smem) + mt * Cfg::kABytes + trow * 128;
        const uint32_t sw = (uint32_t)(trow & 7);
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (kb / kStages) & 1;
            const uint32_t c0 = w0, c1 = w1;
            if (kb + 1 < kblocks) {   // the next k-block's words, in flight while this one is expanded
                w0 = (rvalid && 2 * kb + 2 < wd) ? __ldg(rowp + 2 * kb + 2) : 0u;
                w1 = (rvalid && 2 * kb + 3 < wd) ? __ldg(rowp + 2 * kb + 3) : 0u;
            }
            mbar_wait(empty + s, ph ^ 1);
            const uint32_t rowdst = dst0 + (uint32_t)s * Cfg::kStageBytes;
#pragma unroll
            for (int c = 0; c < 8; ++c) {      // chunk c = elements 8c .. 8c+7 = byte (c & 3) of word (c >> 2)
                const uint32_t v8 = (c < 4 ? c0 : c1) >> (8 * (c & 3));
                const uint32_t o0 = bits2_to_half2(v8 & 3u), o1 = bits2_to_half2((v8 >> 2) & 3u);
                const uint32_t o2 = bits2_to_half2((v8 >> 4) & 3u), o3 = bits2_to_half2((v8 >> 6) & 3u);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowdst + (((uint32_t)c ^ sw) << 4)), "r"(o0), "r"(o1),
                             "r"(o2), "r"(o3) : "memory");
            }
            fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor pipe's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a + s);
        }

        mbar_wait(tmem_full, 0);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < H; cc += 32) {
            float v[32], u[32];
            const uint32_t ta = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + mt * Cfg::kColsPerTile + cc;
            tmem_ld32(ta, v);
            tmem_ld32(ta + H, u);
            const float4* sc4 = reinterpret_cast<const float4*>(inv_scale + n0 + cc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 sc = __ldg(sc4 + j);
                v[4 * j + 0] = __fmul_rn(__fadd_rn(v[4 * j + 0], u[4 * j + 0]), sc.x);
                v[4 * j + 1] = __fmul_rn(__fadd_rn(v[4 * j + 1], u[4 * j + 1]), sc.y);
                v[4 * j + 2] = __fmul_rn(__fadd_rn(v[4 * j + 2], u[4 * j + 2]), sc.z);
                v[4 * j + 3] = __fmul_rn(__fadd_rn(v[4 * j + 3], u[4 * j + 3]), sc.w);
            }
            if (rvalid) {
                float4* dst = reinterpret_cast<float4*>(C + (size_t)row * ldc + n0 + cc);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---- K4 from bits -----------------------------------------------------------------------------------------------------
// Reduction index r = (b, t); a k-block is 32 consecutive time steps of one sample (as in k_wgrad_tc).  The B operand
// (the two tf32 planes of gI, MN-major) arrives by 4-D TMA boxes exactly as there.  The A operand -- 128 input
// features (raster words of x) or 128 hidden units (words of zbits, shifted by one step: Z_{t-1}) x 32 steps -- is
// expanded from raster words into the layout TMA's SWIZZLE_128B_ATOM_32B would have produced: four boxes of
// 32 features x 32 steps, a step's 128-byte row at t * 128, its four 32-byte atoms permuted by (atom ^ (t & 3))
// (UMMA layout SWIZZLE_128B_BASE32B; gemm_tc.cuh).  MT feature tiles share every gI stage.
template <int H, int MT>
struct WgradBitsCfg {
    static constexpr int P = 2;
    static constexpr uint32_t kBoxBytes = kBlockK * 32 * 4;                  // 4 KB: 32 steps x 32 elements
    static constexpr uint32_t kBBytes = P * (H / 32) * kBoxBytes;
    static constexpr uint32_t kStageBytes = MT * kATileBytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
    static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 256;
    static constexpr uint32_t kColsPerTile = tmem_cols_for(P * H);
    static constexpr uint32_t kTmemCols = MT * kColsPerTile;
    static_assert(P * H <= 256, "UMMA N");
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation");
};

struct WgradBitsParams {
    int N, T, B;
    int mtiles_x, mtiles_z;    // feature tiles taking A from the input raster / from the spike raster of the layer
    int m_total;               // N + (recurrent ? H_full : 0)
    int H_full;
    int samples_per_split;
    const uint32_t* xbits; int wd_x;     // (B*T, wd_x) words of the input raster
    const uint32_t* zbits; int wd_z;     // (B*T, wd_z) words of the layer's own spikes
    float* part;               // [S][m_total][H_full]
};

template <int H, int MT>
__global__ void __launch_bounds__(bits_threads<MT>(), 1)
k_wgrad_bits(const __grid_constant__ CUtensorMap map_g, const WgradBitsParams p)
{
    using Cfg = WgradBitsCfg<H, MT>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * Cfg::kStageBytes);
    uint64_t* full_b = bars;
    uint64_t* full_a = bars + kStages;
    uint64_t* empty = bars + 2 * kStages;
    uint64_t* tmem_full = bars + 3 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntile = p.mtiles_x + p.mtiles_z;
    const int tile0 = blockIdx.x * MT;
    const int n0 = blockIdx.z * H;
    const int b_lo = blockIdx.y * p.samples_per_split;
    const int b_hi = min(b_lo + p.samples_per_split, p.B);
    const int tblocks = (p.T + kBlockK - 1) / kBlockK;
    const int kblocks = max(b_hi - b_lo, 0) * tblocks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_g);
        for (int s = 0; s < kStages; ++s) { mbar_init(full_b + s, 1); mbar_init(full_a + s, 4 * MT); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                const int b = b_lo + kb / tblocks, t0 = (kb % tblocks) * kBlockK;
                mbar_wait(empty + s, ph ^ 1);
                unsigned char* st = smem + (size_t)s * Cfg::kStageBytes + MT * kATileBytes;
                mbar_expect_tx(full_b + s, Cfg::kBBytes);
#pragma unroll
                for (int q = 0; q < Cfg::P * (H / 32); ++q)
                    tma_load_4d(st + q * Cfg::kBoxBytes, &map_g, full_b + s, n0 + 32 * (q % (H / 32)), t0, b, q / (H / 32));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(Cfg::P * H, 1, 1);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full_b + s, ph);
                mbar_wait(full_a + s, ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                const uint32_t b0 = a0 + MT * kATileBytes;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    if (tile0 + mt >= ntile) break;
#pragma unroll
                    for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
                        const uint64_t adesc = make_smem_desc(a0 + mt * kATileBytes + kk * 1024, Cfg::kBoxBytes, 512, kLayoutSw128Base32);
                        const uint64_t bdesc = make_smem_desc(b0 + kk * 1024, Cfg::kBoxBytes, 512, kLayoutSw128Base32);
                        umma_tf32(tmem_base + mt * Cfg::kColsPerTile, adesc, bdesc, idesc, (kb | kk) != 0);
                    }
                }
                umma_commit(empty + s);
            }
            umma_commit(tmem_full);
        }
    } else {
        // expansion: this warp fills box j = q of ITS tile (warps 2-5: tile 0, warps 6-9: tile 1), lane = time step
        const int q = warp & 3;
        const int mt = (warp - 2) >> 2;
        const int tile = tile0 + mt;
        const bool fx = tile < p.mtiles_x;
        const bool live = tile < ntile;             // a CTA of the last tile pair may own one tile only
        const int wd = fx ? p.wd_x : p.wd_z;
        const int wcol = (fx ? tile : tile - p.mtiles_x) * (kBlockM / 32) + q;
        const int shift = fx ? 0 : 1;               // the recurrent operand is the PREVIOUS step's raster
        const uint32_t* src = (live && wcol < wd) ? (fx ? p.xbits : p.zbits) : nullptr;
        auto fetch = [&](int kb) -> uint32_t {
            const int b = b_lo + kb / tblocks, tt = (kb % tblocks) * kBlockK + lane - shift;
            return (src && tt >= 0 && tt < p.T) ? __ldg(src + ((size_t)b * p.T + tt) * wd + wcol) : 0u;
        };
        uint32_t nxt = kblocks > 0 ? fetch(0) : 0u;
        const uint32_t hsel = (lane >> 2) & 1;      // lanes t and t+4 share (t & 3): they start on different 16-byte halves
        const uint32_t dst0 = smem_u32(smem) + mt * kATileBytes + q * Cfg::kBoxBytes + lane * 128;
        if (live) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                const uint32_t cur = nxt;
                if (kb + 1 < kblocks) nxt = fetch(kb + 1);
                mbar_wait(empty + s, ph ^ 1);
                const uint32_t rowdst = dst0 + (uint32_t)s * Cfg::kStageBytes;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const uint32_t h = (uint32_t)hh ^ hsel;
                        const uint32_t v4 = cur >> (8 * a + 4 * h);      // elements 8a+4h .. 8a+4h+3 in its low four bits
                        const uint32_t o0 = (0u - (v4 & 1u)) & 0x3F800000u, o1 = (0u - ((v4 >> 1) & 1u)) & 0x3F800000u;
                        const uint32_t o2 = (0u - ((v4 >> 2) & 1u)) & 0x3F800000u, o3 = (0u - ((v4 >> 3) & 1u)) & 0x3F800000u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowdst + ((((uint32_t)a ^ (lane & 3)) << 5) | (h << 4))),
                                     "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                    }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a + s);
            }
        } else {
            // no tile: the MMA warp still waits for 4*MT arrivals per stage
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(empty + s, ph ^ 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a + s);
            }
        }

        if (kblocks > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        if (live) {
            const int m = (fx ? tile : tile - p.mtiles_x) * kBlockM + 32 * q + lane;
            const int mlim = fx ? p.N : p.H_full;
            const int mbase = fx ? 0 : p.N;
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 32) {
                float v[32];
                if (kblocks > 0) {
                    float u[32];
                    const uint32_t ta = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + mt * Cfg::kColsPerTile + c0;
                    tmem_ld32(ta, v);
                    tmem_ld32(ta + H, u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                if (m < mlim) {
                    float4* dst = reinterpret_cast<float4*>(p.part + ((size_t)blockIdx.y * p.m_total + mbase + m) * p.H_full + n0 + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace tc
}  // namespace snnk
