// recur_mma.cuh -- K2 on the tensor cores: the forward recurrence for H = 128 with the recurrent matvec as warp MMAs.
//
// Why: in the SIMT kernel (recur_fwd.cuh) every thread re-reads the whole previous spike vector from shared memory
// at every step -- 64 KB of LDS per row and step, which is what bounds that kernel (ncu: the LSU return path, not
// the FMA pipe).  Here one CTA owns a 16-row batch tile and computes  S(16 x 128) = Z_{t-1}(16 x 128) . W(128 x 128)
// per step with mma.sync.m16n8k16 (bf16 x bf16 -> fp32): the spike tile is exact in bf16, W_rec (.) mask is held as
// three bf16 planes whose sum is the fp32 weight to 2^-24, register-resident as B fragments for the whole sequence
// (96 registers per thread), and the spike tile is fetched with 4 ldmatrix per warp and step: 2 KB of LDS per row and
// step instead of 64 KB.  Warps 0-7 each own 16 neurons (two n8 tiles) and keep their membrane / adaptation state in
// the accumulator fragment layout; warp 8 runs the leaky readout  y_t = kappa y_{t-1} + Z_t W_out + b  the same way,
// one step behind, and tracks the max over time.  (tcgen05 is the wrong tool for this step: its M = 128 tile would
// leave one CTA per 128 rows and 128 accumulator columns per epilogue thread; the per-step tile here is 16 x 128.)
//
// Numerics: products are exact; accumulation order is the tensor pipe's, so results match the fp32 SIMT kernel to
// ~1e-6 relative (not bit for bit) -- the same class as the tcgen05 projection.  Selected with SNNK_F_TENSOR_CORE.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace snnk {

constexpr int kMmaRows = 16;        // batch rows per CTA: one m16 tile
constexpr int kMmaZStride = 136;    // bf16 per shared-memory row of the spike tile (272 B: conflict-free ldmatrix)
constexpr int kMmaChunk = 4;        // time steps per input-current bulk copy
constexpr int kMmaRing = 3;
constexpr int kMmaH = 128;
constexpr int kMmaThreads = 288;    // 8 neuron warps + 1 readout warp

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t saddr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}

// w = hi + mid + lo with three bf16 (8 + 8 + 8 significant bits): exact to 2^-24 |w|
__device__ __forceinline__ void split_bf16x3(float w, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo)
{
    hi = __float2bfloat16_rn(w);
    float r = w - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r);
    r -= __bfloat162float(mid);
    lo = __float2bfloat16_rn(r);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi)
{
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// B fragments (k16 x n8, "col") of the three planes of a 128 x ncols fp32 matrix Wm[k][n] for one n8 tile:
// thread (g = lane >> 2, tig = lane & 3) holds k = 16 kt + 2 tig + {0,1} and + 8, n = n0 + g.
__device__ __forceinline__ void load_b_frags(const float* __restrict__ Wm, int ld, int n, bool n_ok, int tig,
                                             uint32_t (&bf)[8][3][2])
{
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
        __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = 16 * kt + 2 * tig + (q & 1) + 8 * (q >> 1);
            const float w = n_ok ? __ldg(Wm + (size_t)k * ld + n) : 0.f;
            split_bf16x3(w, h[q], m[q], l[q]);
        }
        bf[kt][0][0] = pack_bf16(h[0], h[1]); bf[kt][0][1] = pack_bf16(h[2], h[3]);
        bf[kt][1][0] = pack_bf16(m[0], m[1]); bf[kt][1][1] = pack_bf16(m[2], m[3]);
        bf[kt][2][0] = pack_bf16(l[0], l[1]); bf[kt][2][1] = pack_bf16(l[2], l[3]);
    }
}

constexpr size_t fwd_mma_smem_bytes()
{
    return sizeof(__nv_bfloat16) * 2 * kMmaRows * kMmaZStride +
           sizeof(float) * (size_t)kMmaRing * kMmaRows * kMmaChunk * kMmaH + sizeof(uint64_t) * (kMmaRing + 1);
}

// grid = ceil(B / 16), block = 288
template <bool REC>
__global__ void __launch_bounds__(kMmaThreads, 1) k_recur_fwd_mma(const FwdParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int H = kMmaH;
    const int T = p.T, O = p.O, B = p.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int b0 = blockIdx.x * kMmaRows;
    const int nvalid = min(kMmaRows, B - b0);

    __nv_bfloat16* s_zb = reinterpret_cast<__nv_bfloat16*>(smem_raw);                         // [2][16][136]
    float* s_in = reinterpret_cast<float*>(s_zb + 2 * kMmaRows * kMmaZStride);                 // [ring][16][chunk][128]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_in + kMmaRing * kMmaRows * kMmaChunk * H); // [ring]

    const int nchunks = (T + kMmaChunk - 1) / kMmaChunk;
    auto issue_chunk = [&](int c) {     // one thread: kMmaChunk consecutive steps of every valid row
        const int slot = c % kMmaRing, t0 = c * kMmaChunk;
        const uint32_t bytes = (uint32_t)(min(kMmaChunk, T - t0) * H * sizeof(float));
        tc::mbar_expect_tx(s_bar + slot, bytes * nvalid);
        for (int r = 0; r < nvalid; ++r)
            tc::bulk_g2s(s_in + ((slot * kMmaRows + r) * kMmaChunk) * H, p.I_in + ((size_t)(b0 + r) * T + t0) * H, bytes,
                         s_bar + slot);
    };
    if (tid == 256) {
        for (int s = 0; s < kMmaRing; ++s) tc::mbar_init(s_bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int c = 0; c < kMmaRing && c < nchunks; ++c) issue_chunk(c);
    }
    // spike tile buffers: buffer 1 holds Z_{-1} (initial state), read by step 0
    for (int idx = tid; idx < 2 * kMmaRows * kMmaZStride; idx += kMmaThreads) s_zb[idx] = __float2bfloat16_rn(0.f);
    __syncthreads();
    if (p.Z0) {
        for (int idx = tid; idx < kMmaRows * H; idx += kMmaThreads) {
            const int r = idx / H, c = idx - r * H;
            if (b0 + r < B) s_zb[(kMmaRows + r) * kMmaZStride + c] = __float2bfloat16_rn(p.Z0[(size_t)(b0 + r) * H + c]);
        }
    }

    // ldmatrix source address of this lane inside a spike tile (A operand, m16 x k16 per k-tile)
    const int a_row = (lane & 7) + 8 * ((lane >> 3) & 1), a_kofs = 8 * (lane >> 4);
    const uint32_t a_base = tc::smem_u32(s_zb) + (uint32_t)(a_row * kMmaZStride + a_kofs) * 2;
    constexpr uint32_t kBufBytes = kMmaRows * kMmaZStride * 2;

    if (warp < 8) {
        // ---------------- neuron warps: neurons 16 warp .. 16 warp + 15 ----------------
        uint32_t bf[2][8][3][2];
        if constexpr (REC) {
#pragma unroll
            for (int j = 0; j < 2; ++j) load_b_frags(p.W_eff, H, 16 * warp + 8 * j + g, true, tig, bf[j]);
        }
        const float beta = (p.alif && p.beta) ? __ldg(p.beta) : 0.f;
        float v[2][4], a[2][4], zp[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = g + 8 * (e >> 1), col = 16 * warp + 8 * j + 2 * tig + (e & 1);
                const bool ok = b0 + row < B;
                const size_t s = (size_t)(ok ? b0 + row : 0) * H + col;
                v[j][e] = (ok && p.V0) ? p.V0[s] : 0.f;
                a[j][e] = (ok && p.a0) ? p.a0[s] : 0.f;
                zp[j][e] = (ok && p.Z0) ? p.Z0[s] : 0.f;
            }
        __syncthreads();
        uint16_t* zbits16 = reinterpret_cast<uint16_t*>(p.zbits);

        for (int t = 0; t < T; ++t) {
            const int c = t / kMmaChunk, tt = t - c * kMmaChunk, slot = c % kMmaRing;
            if (tt == 0) tc::mbar_wait(s_bar + slot, (c / kMmaRing) & 1);
            float rec[2][4];
            if constexpr (REC) {
                float acc[2][3][4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[j][pl][e] = 0.f;
                const uint32_t abuf = a_base + ((t + 1) & 1) * kBufBytes;
#pragma unroll
                for (int kt = 0; kt < 8; ++kt) {
                    uint32_t af[4];
                    ldsm4(af, abuf + kt * 32);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl) mma16816(acc[j][pl], af, bf[j][kt][pl]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        rec[j][e] = __fadd_rn(__fadd_rn(acc[j][0][e], acc[j][1][e]), acc[j][2][e]);
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) rec[j][e] = 0.f;
            }
            float zn[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = g + 8 * (e >> 1), col = 16 * warp + 8 * j + 2 * tig + (e & 1);
                    const float cur = (b0 + row < B) ? s_in[((slot * kMmaRows + row) * kMmaChunk + tt) * H + col] : 0.f;
                    // same update as k_recur_fwd (spiking_layers.py:169/239-242)
                    const float t1 = __fmul_rn(p.alpha, v[j][e]);
                    const float t2 = __fadd_rn(t1, cur);
                    const float t3 = __fadd_rn(t2, rec[j][e]);
                    const float vn = __fmul_rn(t3, __fsub_rn(1.0f, zp[j][e]));
                    float thr = p.theta;
                    if (p.alif) {
                        a[j][e] = __fadd_rn(__fmul_rn(p.rho, a[j][e]), zp[j][e]);
                        thr = __fadd_rn(p.theta, __fmul_rn(beta, a[j][e]));
                    }
                    zn[j][e] = vn >= thr ? 1.0f : 0.0f;
                    v[j][e] = vn;
                    zp[j][e] = zn[j][e];
                }
                // traces: (e0,e1) and (e2,e3) are adjacent columns of rows g and g + 8
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int row = g + 8 * hrow, col = 16 * warp + 8 * j + 2 * tig;
                    if (b0 + row < B) {
                        if (p.traces) {
                            const size_t o = ((size_t)(b0 + row) * T + t) * H + col;
                            *reinterpret_cast<float2*>(p.V + o) = make_float2(v[j][2 * hrow], v[j][2 * hrow + 1]);
                            *reinterpret_cast<float2*>(p.Z + o) = make_float2(zn[j][2 * hrow], zn[j][2 * hrow + 1]);
                            if (p.alif) *reinterpret_cast<float2*>(p.a + o) = make_float2(a[j][2 * hrow], a[j][2 * hrow + 1]);
                        }
                    }
                    // spike tile of this step for the next one (bf16 pair = one 32-bit store)
                    *reinterpret_cast<uint32_t*>(s_zb + ((t & 1) * kMmaRows + row) * kMmaZStride + col) =
                        pack_bf16(__float2bfloat16_rn(zn[j][2 * hrow]), __float2bfloat16_rn(zn[j][2 * hrow + 1]));
                }
            }
            // bit-packed raster: 16 bits (this warp's neurons) per row; lane r < 16 assembles row r from the ballots
            unsigned bal[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) bal[j][e] = __ballot_sync(0xffffffffu, zn[j][e] != 0.f);
            if (lane < 16) {
                const int gg = lane & 7, hi_rows = lane >> 3;
                unsigned out = 0;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const unsigned word = hi_rows ? bal[j][2 + e2] : bal[j][e2];
                        const unsigned nib = (word >> (4 * gg)) & 0xFu;          // bit tig -> neuron 8 j + 2 tig + e2
                        const unsigned sp = (nib & 1u) | ((nib & 2u) << 1) | ((nib & 4u) << 2) | ((nib & 8u) << 3);
                        out |= sp << (8 * j + e2);
                    }
                if (b0 + lane < B) zbits16[((size_t)(b0 + lane) * T + t) * (H / 16) + warp] = (uint16_t)out;
            }
            __syncthreads();
        }
    } else {
        // ---------------- readout warp: y_t = kappa y_{t-1} + Z_t W_out + b, one step behind the neuron warps ----------------
        uint32_t bf[2][8][3][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) load_b_frags(p.W_out, O, 8 * j + g, 8 * j + g < O, tig, bf[j]);
        float yv[2][4], mx[2][4], bias[2][4];
        int mt[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = 8 * j + 2 * tig + (e & 1);
                yv[j][e] = 0.f; mx[j][e] = 0.f; mt[j][e] = 0;
                bias[j][e] = col < O ? __ldg(p.b_out + col) : 0.f;
            }
        __syncthreads();
        for (int t = 0; t <= T; ++t) {
            if (t < T) {
                const int c = t / kMmaChunk, tt = t - c * kMmaChunk;
                // every thread is past its reads of chunk c - 1 (the barrier that ended step t - 1): refill that slot
                if (tt == 0 && lane == 0 && c >= 1 && c - 1 + kMmaRing < nchunks) issue_chunk(c - 1 + kMmaRing);
            }
            if (t >= 1) {
                const int ty = t - 1;                       // spikes of step ty are in buffer ty & 1 == (t + 1) & 1
                float acc[2][3][4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[j][pl][e] = 0.f;
                const uint32_t abuf = a_base + ((t + 1) & 1) * kBufBytes;
#pragma unroll
                for (int kt = 0; kt < 8; ++kt) {
                    uint32_t af[4];
                    ldsm4(af, abuf + kt * 32);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl) mma16816(acc[j][pl], af, bf[j][kt][pl]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = g + 8 * (e >> 1), col = 8 * j + 2 * tig + (e & 1);
                        const float s = __fadd_rn(__fadd_rn(acc[j][0][e], acc[j][1][e]), acc[j][2][e]);
                        const float y = __fadd_rn(__fadd_rn(__fmul_rn(p.kappa, yv[j][e]), s), bias[j][e]);
                        yv[j][e] = y;
                        if (ty == 0 || y > mx[j][e]) { mx[j][e] = y; mt[j][e] = ty; }      // first max wins (snn.py:228)
                        if (col < O && b0 + row < B) p.y[((size_t)(b0 + row) * T + ty) * O + col] = y;
                    }
            }
            if (t < T) __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = g + 8 * (e >> 1), col = 8 * j + 2 * tig + (e & 1);
                if (col < O && b0 + row < B) {
                    p.logits[(size_t)(b0 + row) * O + col] = mx[j][e];
                    p.tstar[(size_t)(b0 + row) * O + col] = mt[j][e];
                }
            }
    }
}

}  // namespace snnk
