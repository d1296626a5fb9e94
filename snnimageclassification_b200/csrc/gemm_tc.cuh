// gemm_tc.cuh -- tcgen05 / TMA / TMEM versions of the two contractions of the path (sm_100a only).
//
//   k_proj_tc    I_in[r][n]   = sum_k x[r][k] W_in[k][n]         (K1; spiking_layers.py:163/233 for all T at once)
//   k_wgrad_tc   dW_in[m][n]  = sum_r x[r][m] gI[r][n]           (K4; MmBackward of the projection)
//                dW_rec[j][n] = sum_r Z_{t-1}[r][j] gI[r][n]     (K4; MmBackward of the recurrent matmul)
//
// Numerics: kind::tf32 with fp32 accumulation in TMEM.  The spike operand (x, Z) is exactly {0,1}; the fp32
// operand is split into tf32 planes whose sum is exact to 2^-22 (W_in: the first two of three round-to-nearest planes;
// gI: truncated high plane + remainder), so every product is exact and only the accumulation order differs from
// the fp32 SIMT kernels.  The planes of one operand sit back to back in shared memory and enter ONE MMA of N = 2H;
// the epilogue adds the two TMEM column groups.  x is read straight from the user's fp32 tensor by TMA -- no
// conversion pass.  While the MMA pipeline runs, the
// otherwise idle epilogue warps check every x tile in shared memory for values that are not exactly
// representable in tf32; if any is found a device flag is raised and the caller's fp32 SIMT kernel (which
// is always launched behind this one and exits immediately when the flag is clear) recomputes the result.
//
// Structure (both kernels): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma
// issuer, warps 2-5 = exactness check + epilogue (tcgen05.ld -> registers -> global).  smem ring of
// kStages {A tile, B planes} with full/empty mbarriers; accumulator 128 x 2H fp32 in TMEM.  Both kernels can be
// gated on a frame-run table (runs.cuh): the launch returns at once unless the table's ok word asks for it.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace snnk {
namespace tc {

constexpr int kThreads = 192;
constexpr int kBlockM = 128;      // UMMA M (cta_group::1)
constexpr int kBlockK = 32;       // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;         // K per tcgen05.mma for kind::tf32
constexpr uint32_t kATileBytes = kBlockM * kBlockK * 4;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Bounded wait: a pipeline bug must trap (the launch then fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// 1-D bulk async copy global -> shared (TMA without a tensor map); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version field = 1).
//   K-major : rows of 128 B (one k-block), 8-row groups SBO = 1024 B apart, LBO unused (encoded as 1)
//   MN-major: 32-element (128 B) MN blocks LBO bytes apart, 8-k groups SBO = 1024 B apart
constexpr uint32_t kLayoutSw128 = 2;         // 16-byte chunks swizzled within 128 B  (TMA SWIZZLE_128B)
constexpr uint32_t kLayoutSw128Base32 = 1;   // 32-byte chunks swizzled within 128 B  (TMA SWIZZLE_128B_ATOM_32B):
                                             // the only layout tcgen05 accepts for MN-major tf32 operands
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = kLayoutSw128)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type & 7u) << 61;
    return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, M = 128.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n, int a_mn_major, int b_mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of fp32 accumulator -> 32 registers per thread (thread = TMEM lane = output row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// round to nearest even at tf32 precision (11 significant bits); the result has its low 13 mantissa bits clear
__device__ __forceinline__ float rn_tf32(float x)
{
    uint32_t u = __float_as_uint(x);
    u += 0xFFFu + ((u >> 13) & 1u);
    return __uint_as_float(u & 0xFFFFE000u);
}

__host__ __device__ constexpr uint32_t tmem_cols_for(int n) { return n <= 32 ? 32u : (n <= 64 ? 64u : (n <= 128 ? 128u : 256u)); }

// A-tile exactness check run by the four epilogue warps while the tensor pipe works: any fp32 value with
// mantissa bits below tf32 precision raises the flag (the swizzle only permutes 16-byte chunks, so a flat
// scan of the tile sees every element).
__device__ __forceinline__ uint32_t scan_inexact(const void* tile, uint32_t bytes, int tid128)
{
    const uint4* p = static_cast<const uint4*>(tile);
    uint32_t bad = 0;
    for (uint32_t i = tid128; i < bytes / 16; i += 128) {
        const uint4 v = p[i];
        bad |= (v.x | v.y | v.z | v.w) & 0x1FFFu;
    }
    return bad;
}

// ---- K1 ------------------------------------------------------------------------------------------------------------
// map_x : 2-D (K, M) fp32, box (32, 128), SWIZZLE_128B        -> A tile, K-major
// planes: tf32-planes of W_in^T pre-tiled and pre-swizzled by k_split_w -> B planes, K-major, one bulk copy each
template <int H, int P>
struct ProjCfg {
    static constexpr uint32_t kBPlaneBytes = H * kBlockK * 4;
    static constexpr uint32_t kStageBytes = kATileBytes + P * kBPlaneBytes;
    // the tile time is bytes in flight / latency: as many stages as shared memory holds (8 x 24 KB for 32-column tiles)
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 256;
};

template <int H, int P>
__global__ void __launch_bounds__(kThreads, 1)
k_proj_tc(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ planes, float* __restrict__ C,
          int M, int kblocks, int ldc, unsigned int* __restrict__ inexact_flag, const int* __restrict__ run_table,
          int run_variant, const float* __restrict__ a_tiled)
{
    // Programmatic dependent launch: the recurrence kernel queued behind this one may start NOW (on SMs this grid does
    // not occupy) and run its prologue -- weight staging, label statistics -- until its griddepcontrol.wait, which
    // returns when this grid has completed.  A no-op when the dependent was launched the ordinary way.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // frame-dedup gating (runs.cuh): run_variant 0 = dense kernel, skipped when the table says the compact kernels
    // run; 1 = compact kernel over M = n_rows rows, skipped otherwise.  Decided before any barrier or TMEM allocation.
    if (run_table) {
        const int ok = run_table[1];
        if (ok != run_variant) return;
        if (run_variant == 1) M = min(M, run_table[0]);
        if ((int)blockIdx.x * kBlockM >= M) return;
    }
    // H is the N extent of this CTA's tile; blockIdx.y selects the tile, ldc is the full hidden width
    using Cfg = ProjCfg<H, P>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * Cfg::kStageBytes);
    uint64_t* full = bars;                     // [kStages]  TMA -> MMA / checker
    uint64_t* empty = bars + kStages;          // [kStages]  MMA commit + 4 checker warps -> TMA
    uint64_t* tmem_full = bars + 2 * kStages;  // accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM;
    const int n0 = blockIdx.y * H;
    // The P weight planes sit back to back in the stage, i.e. they ARE one K-major B operand of N = P*H rows: one MMA
    // per k-step computes a . p0 and a . p1 side by side in TMEM (a tf32 MMA of K = 8 costs the same ~150 cycles for
    // N = 32 ... 256, so halving the instruction count halves the tile time); the epilogue adds the column groups.
    static_assert(P * H <= 256, "UMMA N");
    constexpr uint32_t kTmemCols = tmem_cols_for(P * H);
    const bool check = inexact_flag != nullptr;   // null: the caller vouches for {0,1} inputs (SNNK_F_INPUT_BINARY)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, check ? 1 + 4 : 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
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
                unsigned char* st = smem + (size_t)s * Cfg::kStageBytes;
                mbar_expect_tx(full + s, Cfg::kStageBytes);
                if (a_tiled)   // pre-tiled, pre-swizzled A (k_gather_rows_tiled): one contiguous 16 KB copy
                    bulk_g2s(st, a_tiled + ((size_t)blockIdx.x * kblocks + kb) * (kBlockM * kBlockK), kATileBytes, full + s);
                else
                    tma_load_2d(st, &map_x, full + s, kb * kBlockK, m0);
#pragma unroll
                for (int p = 0; p < P; ++p)
                    bulk_g2s(st + kATileBytes + p * Cfg::kBPlaneBytes,
                             planes + ((size_t)p * kblocks + kb) * ((size_t)ldc * kBlockK) + (size_t)n0 * kBlockK,
                             Cfg::kBPlaneBytes, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(P * H, 0, 0);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                const uint32_t b0 = a0 + kATileBytes;
#pragma unroll
                for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
                    const uint64_t adesc = make_smem_desc(a0 + kk * kUmmaK * 4, 16, 1024);
                    const uint64_t bdesc = make_smem_desc(b0 + kk * kUmmaK * 4, 16, 1024);
                    umma_tf32(tmem_base, adesc, bdesc, idesc, (kb | kk) != 0);
                }
                umma_commit(empty + s);      // frees the stage once the MMAs above have read it
            }
            umma_commit(tmem_full);
        }
    } else {
        // warps 2..5: exactness check of every A tile, then the epilogue
        if (check) {
            const int tid128 = threadIdx.x - 64;
            uint32_t bad = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full + s, ph);
                bad |= scan_inexact(smem + (size_t)s * Cfg::kStageBytes, kATileBytes, tid128);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            if (__any_sync(0xffffffffu, bad != 0) && lane == 0) atomicOr(inexact_flag, 1u);
        }

        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row = m0 + 32 * q + lane;
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + c0, v);
#pragma unroll
            for (int p = 1; p < P; ++p) {
                float u[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + p * H + c0, u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
            if (row < M) {
                float4* dst = reinterpret_cast<float4*>(C + (size_t)row * ldc + n0 + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// W_in (K,H) fp32 -> tf32 planes of its transpose, stored the way the tensor pipe wants them in shared memory:
// planes[p][kb][h][32] -- per k-block of 32 one K-major tile row of 128 B per hidden unit, its eight 16-byte chunks
// permuted by (chunk ^ (h & 7)), i.e. exactly what a SWIZZLE_128B tensor-map load would have produced.  A CTA's B
// operand of one k-block (hidden units n0 .. n0+H) is then ONE contiguous range, fetched by a single 1-D bulk copy
// instead of a tensor-map box of H strided 128-byte rows (the TMA unit's request rate, not bytes, bounded the tile).
// w = p0 + p1 + p2 exactly; the projection uses the first two planes (round-to-nearest split: the dropped
// remainder is at most 2^-22 |w|, i.e. at the level of fp32's own rounding and ten times below the
// tensor pipe's accumulation noise; see DESIGN.md "Numerics").
__global__ void __launch_bounds__(256) k_split_w(const float* __restrict__ W, int K, int H, int Kpad,
                                                float* __restrict__ planes)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * Kpad) return;
    const int h = idx / Kpad, k = idx - h * Kpad;
    const float w = k < K ? W[(size_t)k * H + h] : 0.f;
    const float p0 = rn_tf32(w);
    const float r = w - p0;             // exact
    const float p1 = rn_tf32(r);        // |w - p0 - p1| <= 2^-22 |w|
    const float p2 = (r - p1);          // exact remainder, itself tf32-representable: p0 + p1 + p2 == w
    const int kb = k / kBlockK, kk = k - kb * kBlockK;
    const size_t o = ((size_t)kb * H + h) * kBlockK + (size_t)((((kk >> 2) ^ (h & 7)) << 2) | (kk & 3));
    const size_t plane = (size_t)H * Kpad;
    planes[o] = p0;
    planes[plane + o] = p1;
    planes[2 * plane + o] = p2;
}

// ---- K4 ------------------------------------------------------------------------------------------------------------
// Reduction index r = (b, t); every k-block is 32 consecutive time steps of ONE sample, loaded through 3-D/4-D
// tensor maps so that t < 0 (the Z_{t-1} shift) and t >= T are zero-filled by the TMA unit.
// All three maps use SWIZZLE_128B_ATOM_32B and the MMA descriptors layout type SWIZZLE_128B_BASE32B: the only
// shared-memory layout tcgen05 accepts for MN-major tf32 operands (established with tools/umma_probe.cu):
// 32-element MN blocks LBO = 4 KB apart, groups of 4 k-rows SBO = 512 B apart, 1 KB per UMMA_K = 8.
// map_x : 3-D (N, T, B) fp32, box (32, 32, 1)        -> A tile (4 boxes), MN-major
// map_z : 3-D (H, T, B) fp32 spike trace, box (32, 32, 1), read at t0 - 1
// map_g : 4-D (H, T, B, P) fp32 tf32-planes of gI, box (32, 32, 1, 1) -> B planes, MN-major
template <int H, int P>
struct WgradCfg {
    static constexpr uint32_t kBoxBytes = kBlockK * 32 * 4;                  // 4 KB: 32 t x 32 elements
    static constexpr uint32_t kBBytes = P * (H / 32) * kBoxBytes;
    static constexpr uint32_t kStageBytes = kATileBytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
    static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 256;
};

struct WgradTcParams {
    int N, T, B;
    int mtiles_x;              // tiles taking A from x; the remaining blockIdx.x take it from the spike trace
    int m_total;               // N + H_full
    int H_full;                // full hidden width (the template H is the N extent of one CTA tile)
    int samples_per_split;
    float* part;               // [S][m_total][H]
    unsigned int* inexact_flag;   // raised when an x tile holds values that are not tf32-exact
    const int* run_table;      // frame-dedup gating (runs.cuh) or null
    int run_gate;              // the launch runs only when the table's ok word equals this (0: dense, 1: dedup variant)
    int run_clip;              // 1: one "sample" of T compact rows, 32-row blocks dealt round-robin to the splits, only
                               //    the first n_rows rows are contracted
};

template <int H, int P>
__global__ void __launch_bounds__(kThreads, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z,
           const __grid_constant__ CUtensorMap map_g, const WgradTcParams p)
{
    using Cfg = WgradCfg<H, P>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * Cfg::kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool from_x = (int)blockIdx.x < p.mtiles_x;
    const int m0 = from_x ? blockIdx.x * kBlockM : (blockIdx.x - p.mtiles_x) * kBlockM;
    const int n0 = blockIdx.z * H;
    const int b_lo = blockIdx.y * p.samples_per_split;
    const int b_hi = min(b_lo + p.samples_per_split, p.B);
    const int tblocks = (p.T + kBlockK - 1) / kBlockK;
    int kblocks = (b_hi - b_lo) * tblocks;
    if (p.run_table) {
        if (p.run_table[1] != p.run_gate) return;
        if (p.run_clip) {   // compact rows: split y contracts the 32-row blocks y, y + S, y + 2S, ... below n_rows
            const int nblk = (p.run_table[0] + kBlockK - 1) / kBlockK;
            kblocks = (int)blockIdx.y < nblk ? (nblk - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
        }
    }
    // the P planes of gI are adjacent MN blocks of the stage: one B operand of N = P*H (see k_proj_tc)
    static_assert(P * H <= 256, "UMMA N");
    constexpr uint32_t kTmemCols = tmem_cols_for(P * H);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_z);
        prefetch_tmap(&map_g);
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, p.inexact_flag ? 1 + 4 : 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                const int b = p.run_clip ? 0 : b_lo + kb / tblocks;
                const int t0 = p.run_clip ? ((int)blockIdx.y + kb * (int)gridDim.y) * kBlockK : (kb % tblocks) * kBlockK;
                mbar_wait(empty + s, ph ^ 1);
                unsigned char* st = smem + (size_t)s * Cfg::kStageBytes;
                mbar_expect_tx(full + s, Cfg::kStageBytes);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (from_x) tma_load_3d(st + j * Cfg::kBoxBytes, &map_x, full + s, m0 + 32 * j, t0, b);
                    else tma_load_3d(st + j * Cfg::kBoxBytes, &map_z, full + s, m0 + 32 * j, t0 - 1, b);
                }
#pragma unroll
                for (int q = 0; q < P * (H / 32); ++q)
                    tma_load_4d(st + kATileBytes + q * Cfg::kBoxBytes, &map_g, full + s, n0 + 32 * (q % (H / 32)), t0, b,
                                q / (H / 32));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(P * H, 1, 1);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                const uint32_t b0 = a0 + kATileBytes;
#pragma unroll
                for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
                    const uint64_t adesc = make_smem_desc(a0 + kk * 1024, Cfg::kBoxBytes, 512, kLayoutSw128Base32);
                    const uint64_t bdesc = make_smem_desc(b0 + kk * 1024, Cfg::kBoxBytes, 512, kLayoutSw128Base32);
                    umma_tf32(tmem_base, adesc, bdesc, idesc, (kb | kk) != 0);
                }
                umma_commit(empty + s);
            }
            umma_commit(tmem_full);
        }
    } else {
        // warps 2..5: exactness check of the x tiles (the spike trace is {0,1} by construction), then the epilogue
        if (p.inexact_flag) {
            const int tid128 = threadIdx.x - 64;
            uint32_t bad = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full + s, ph);
                if (from_x) bad |= scan_inexact(smem + (size_t)s * Cfg::kStageBytes, kATileBytes, tid128);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            if (__any_sync(0xffffffffu, bad != 0) && lane == 0) atomicOr(p.inexact_flag, 1u);
        }
        if (kblocks > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        const int q = warp & 3;
        const int m = m0 + 32 * q + lane;
        const int mlim = from_x ? p.N : p.H_full;
        const int mbase = from_x ? 0 : p.N;
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 32) {
            float v[32];
            if (kblocks > 0) {
                tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + c0, v);
#pragma unroll
                for (int pl = 1; pl < P; ++pl) {
                    float u[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + pl * H + c0, u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                }
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace snnk
