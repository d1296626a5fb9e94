// recur_tc.cuh -- K2 / K3 on the tensor cores for H = 128: the recurrent matvec of every time step as warp MMAs with the
// NEURONS on M and the batch rows on N.
//
// Why this shape.  Measured on B200 (tools/ubench.cu, profiles/r02_ubench.txt): mma.sync.m16n8k16 issues every 8.1
// cycles per SM sub-partition (1012 MAC/cycle/SM, 8x the fp32 FMA rate) with 21 cycles of latency, a 288-thread
// bar.sync costs 31 cycles.  The fp32 SIMT kernels (recur_fwd.cuh / recur_bwd.cuh) spend >= 128 FMA-pipe cycles per
// row and step and ran at ~1000 / ~1300 cycles per step with two rows per SM: latency-bound, 4x off even that floor.
// Here one CTA owns EIGHT batch rows (the n8 of the MMA) for all T steps; warp w owns neurons 16w..16w+15 (the m16),
// the masked recurrent matrix lives in its A fragments for the whole sequence, and a step is
//     S^T (128 x 8) = W^T (128 x 128) . Z_{t-1}^T (128 x 8)
// i.e. 16 (forward) / 24 (backward) MMAs per warp against B fragments fetched with 4 / 8 ldmatrix from a 2 KB tile in
// shared memory, 4 state elements per thread, one stmatrix that publishes the new spikes / gradients, one barrier.
// A batch of 256 rows then occupies 32 SMs for ~450-600 cycles per step instead of 148 SMs for 1000-1300.
// (tcgen05 does not fit this step: its M = 128 tile would be the neurons as well, but the round trip
// mma -> commit -> mbarrier -> tcgen05.ld -> registers -> st.shared -> fence.proxy.async per step costs more than the
// 130-190 cycles of MMA time it saves; the tile per step is 128 x 8 x 128.)
//
// Numerics (tensor-core mode only; the fp32 SIMT kernels stay the bit-exact mode).  Spikes are exact in fp16.  A
// weight is split as  w s = hi + lo / 2048  with hi, lo in fp16 and s a power of two that puts max|W| at 2^13..2^14:
// 22 significant bits, the same class as the tf32-plane tcgen05 GEMMs (gemm_tc.cuh).  Products are exact, sums are
// fp32 in the tensor pipe, so results differ from the fp32 kernels by summation order only (~1e-6 relative).  The
// backward operand gI is real-valued: it is split the same way with a scale chosen per CTA and per step from the
// largest exponent in the tile (an OR of exponent-class bits across the CTA before the tile is published), and the three products
// hi.hi, hi.lo, lo.hi are accumulated (the dropped lo.lo term is 2^-22 relative).
//
// Replaces, like recur_fwd.cuh / recur_bwd.cuh: the time loop of SNN.forward (src/modules/snn.py:209-214) around
// LIFLayer/ALIFLayer.forward (src/modules/spiking_layers.py:156-171, :229-243) and ReadoutLayer.forward (:402-408),
// and autograd's reverse sweep for batch_loss.backward() (snn.py:413) with the surrogates of spike_funcs.py:59-62/75-79.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace snnk {

constexpr int kTcH = 128;          // hidden width of these kernels
constexpr int kTcRows = 8;         // batch rows per CTA = N of the MMA
constexpr int kTcThreads = 288;    // 8 neuron warps + 1 service warp
constexpr int kTcTileStride = 136; // halves per row of a spike / gradient tile (272 B: conflict-free ldmatrix)

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}

__device__ __forceinline__ void stsm_x2_trans(uint32_t saddr, uint32_t r0, uint32_t r1)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(saddr), "r"(r0), "r"(r1) : "memory");
}

__device__ __forceinline__ void stsm_x4_trans(uint32_t saddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(r0), "r"(r1),
                 "r"(r2), "r"(r3)
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(__half lo, __half hi)
{
    return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}

// x s = hi + lo / 2048 (hi, lo fp16): 22 significant bits when x s is in fp16's normal range
__device__ __forceinline__ void split_h2(float xs, __half& hi, __half& lo)
{
    hi = __float2half_rn(xs);
    lo = __float2half_rn(__fmul_rn(__fsub_rn(xs, __half2float(hi)), 2048.0f));
}

// power of two s with  max * s  in [2^13, 2^14)  (1 for max == 0); exact to multiply and divide by
__device__ __forceinline__ float pow2_scale_for(float mx)
{
    if (!(mx > 0.f)) return 1.0f;
    int e = (int)((__float_as_uint(mx) >> 23) & 0xFFu) - 127;     // floor(log2(mx)) for normal numbers
    int k = 13 - e;
    k = k > 120 ? 120 : (k < -120 ? -120 : k);
    return __uint_as_float((uint32_t)(k + 127) << 23);
}

// Block-wide maximum of a non-negative value (all kTcThreads threads call it; s_red: 9 floats of shared memory).
__device__ __forceinline__ float block_max_tc(float v, float* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = s_red[0];
#pragma unroll
    for (int q = 1; q < kTcThreads / 32; ++q) m = fmaxf(m, s_red[q]);
    __syncthreads();
    return m;
}

// A fragments (m16 x k16, "row") of the two fp16 planes of Wm[k][i] (row-major, leading dimension ld) for the 16
// output neurons i0..i0+15 and K = 16 * KT: thread (g = lane >> 2, tig = lane & 3) holds rows m = g, g + 8 and
// k = 16 kt + 2 tig + {0, 1} (+ 8).  raw[] receives the fp32 values (so the caller can find the scale first).
template <int KT>
__device__ __forceinline__ void load_a_raw(const float* __restrict__ Wm, int ld, int i0, int kmax, int g, int tig,
                                           float (&raw)[KT][8])
{
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            // register order a0 a1 a2 a3, two halves each: a0 = (g, k..k+1), a1 = (g+8, k..), a2 = (g, k+8..), a3 = (g+8, k+8..)
            const int k = 16 * kt + 2 * tig + (q & 1) + 8 * (q >> 2);
            const int m = g + 8 * ((q >> 1) & 1);
            raw[kt][q] = k < kmax ? __ldg(Wm + (size_t)k * ld + i0 + m) : 0.f;
        }
}

template <int KT>
__device__ __forceinline__ void split_a(const float (&raw)[KT][8], float s, uint32_t (&ah)[KT][4], uint32_t (&al)[KT][4])
{
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            __half h0, l0, h1, l1;
            split_h2(__fmul_rn(raw[kt][2 * r], s), h0, l0);
            split_h2(__fmul_rn(raw[kt][2 * r + 1], s), h1, l1);
            ah[kt][r] = pack_h2(h0, h1);
            al[kt][r] = pack_h2(l0, l1);
        }
}

template <int KT>
__device__ __forceinline__ float absmax_raw(const float (&raw)[KT][8])
{
    float m = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int q = 0; q < 8; ++q) m = fmaxf(m, fabsf(raw[kt][q]));
    return m;
}

constexpr size_t fwd_tc_smem_bytes(int T)
{
    return sizeof(__half) * 2 * kTcRows * kTcTileStride                                   // spike tiles
           + sizeof(float) * 2 * 8 * 32 * 4                                               // readout partials [2][warp][lane][4]
           + sizeof(int) * (size_t)((kTcRows * T + 3) & ~3)                               // compact row of every (row, step)
           + sizeof(float) * 16;
}

// ---- forward --------------------------------------------------------------------------------------------------------
// grid = ceil(B / 8), block = 288.  Recurrent layers only (without the matvec the SIMT kernel has nothing to lose).
template <bool ALIF>
__global__ void __launch_bounds__(kTcThreads, 1) k_recur_fwd_tc(const FwdParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int H = kTcH;
    const int T = p.T, O = p.O, B = p.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int b0 = blockIdx.x * kTcRows;
    const int nvalid = min(kTcRows, B - b0);

    __half* s_z = reinterpret_cast<__half*>(smem_raw);                                          // [2][8][136]
    float* s_yp = reinterpret_cast<float*>(s_z + 2 * kTcRows * kTcTileStride);                   // [2][8][32][4]
    int* s_r2c = reinterpret_cast<int*>(s_yp + 2 * 8 * 32 * 4);                                  // [8][T]
    float* s_red = reinterpret_cast<float*>(s_r2c + ((kTcRows * T + 3) & ~3));                   // [16]

    // The input current is read straight from global memory / L2 into registers, two steps ahead of its use (four
    // coalesced 32-byte segments per warp load).  A shared-memory ring fed by per-(row, step) bulk copies was measured
    // first: the SM's TMA unit takes ~270 cycles per 512-byte cp.async.bulk, 8 of them per step -> 2000 cycles per step.
    // frame-dedup variant: row table[b*T + t] of the compact projection I_u instead of row b*T + t of I_in.
    const bool compact = p.run_table != nullptr && p.run_table[1] == 1;
    if (compact)
        for (int idx = tid; idx < nvalid * T; idx += kTcThreads)
            s_r2c[idx] = __ldg(p.run_table + kRunHdrInts + (size_t)b0 * T + idx);      // rows b0.. are consecutive
    // spike tiles: buffer 1 holds Z_{-1} (the initial state), read by step 0
    for (int idx = tid; idx < 2 * kTcRows * kTcTileStride; idx += kTcThreads) s_z[idx] = __float2half_rn(0.f);
    __syncthreads();
    if (p.Z0)
        for (int idx = tid; idx < kTcRows * H; idx += kTcThreads) {
            const int r = idx / H, c = idx - r * H;
            if (b0 + r < B) s_z[(kTcRows + r) * kTcTileStride + c] = __float2half_rn(p.Z0[(size_t)(b0 + r) * H + c]);
        }

    // ldmatrix source of this lane inside a tile: matrix j = lane >> 3 covers k = 8 j .. 8 j + 7 of a 32-wide k group
    const uint32_t tile_base = tc::smem_u32(s_z) + (uint32_t)((lane & 7) * kTcTileStride + 8 * (lane >> 3)) * 2;
    constexpr uint32_t kTileBytes = kTcRows * kTcTileStride * 2;

    if (warp < 8) {
        // ---------------- neuron warps: neurons 16 warp .. 16 warp + 15 ----------------
        const int i0 = 16 * warp;
        uint32_t ah[8][4], al[8][4];          // W_eff^T fragments, two fp16 planes: 64 registers for the whole sequence
        uint32_t oh[4], ol[4];                // W_out^T fragment of this warp's 16 neurons (readout partial)
        float inv_s, inv_so;
        {
            float raw[8][8];
            load_a_raw<8>(p.W_eff, H, i0, H, g, tig, raw);
            const float mx = block_max_tc(absmax_raw<8>(raw), s_red);
            const float s = pow2_scale_for(mx);
            inv_s = __fdiv_rn(1.0f, s);
            split_a<8>(raw, s, ah, al);
            // readout: A[m = class][k = neuron i0 + k] = W_out[i0 + k][m]
            float ro[1][8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = 2 * tig + (q & 1) + 8 * (q >> 2), m = g + 8 * ((q >> 1) & 1);
                ro[0][q] = m < O ? __ldg(p.W_out + (size_t)(i0 + k) * O + m) : 0.f;
            }
            const float mo = block_max_tc(absmax_raw<1>(ro), s_red);
            const float so = pow2_scale_for(mo);
            inv_so = __fdiv_rn(1.0f, so);
            uint32_t th[1][4], tl[1][4];
            split_a<1>(ro, so, th, tl);
#pragma unroll
            for (int r = 0; r < 4; ++r) { oh[r] = th[0][r]; ol[r] = tl[0][r]; }
        }
        const float beta = (ALIF && p.beta) ? __ldg(p.beta) : 0.f;
        // element e = 2 nh + rh: neuron i0 + g + 8 nh, row 2 tig + rh   (the accumulator fragment layout)
        float v[4], a[4], zp[4];
        bool ok[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int row = 2 * tig + (e & 1), col = i0 + g + 8 * (e >> 1);
            ok[e] = b0 + row < B;
            const size_t s = (size_t)(ok[e] ? b0 + row : 0) * H + col;
            v[e] = (ok[e] && p.V0) ? p.V0[s] : 0.f;
            a[e] = (ok[e] && p.a0) ? p.a0[s] : 0.f;
            zp[e] = (ok[e] && p.Z0) ? p.Z0[s] : 0.f;
        }
        const size_t o_base = ((size_t)(b0 + 2 * tig) * T) * H + i0 + g;      // (row 2 tig, t = 0, neuron i0 + g)
        const size_t o_row = (size_t)T * H;
        __syncthreads();      // tiles and the compact-row table initialised
        auto load_cur = [&](int tl, float (&dst)[4]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dst[e] = 0.f;
                if (tl < T && ok[e]) {
                    const int rh = e & 1;
                    const float* src = compact ? p.I_u + (size_t)s_r2c[(2 * tig + rh) * T + tl] * H + i0 + g
                                               : p.I_in + o_base + (size_t)tl * H + rh * o_row;
                    dst[e] = __ldg(src + 8 * (e >> 1));
                }
            }
        };
        float cur[4], nx1[4], nx2[4];
        load_cur(0, cur);
        load_cur(1, nx1);

        for (int t = 0; t <= T; ++t) {
            load_cur(t + 2, nx2);
            // B fragments of Z_{t-1}: bfr[kt] = (k = 16 kt + 2 tig + {0,1}, n = g), (k + 8 ..)
            uint32_t bfr[8][2];
            const uint32_t tb = tile_base + ((t + 1) & 1) * kTileBytes;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t r[4];
                ldsm_x4(r, tb + q * 64);
                bfr[2 * q][0] = r[0]; bfr[2 * q][1] = r[1]; bfr[2 * q + 1][0] = r[2]; bfr[2 * q + 1][1] = r[3];
            }
            if (t >= 1) {
                // readout partial of step t-1 over this warp's 16 neurons: (classes x rows) += W_out^T[:, i0..] Z_{t-1}^T
                // (its B fragment = k-tile `warp` of the tile, fetched by address: indexing bfr[] with the warp number
                // would put the array into local memory)
                float yh[4] = {0.f, 0.f, 0.f, 0.f}, yl[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t* own = reinterpret_cast<const uint32_t*>(s_z + (((t + 1) & 1) * kTcRows + g) * kTcTileStride + i0 + 2 * tig);
                const uint32_t bo[2] = {own[0], own[4]};
                mma_f16(yh, oh, bo);
                mma_f16(yl, ol, bo);
                float4 part;
                part.x = __fmul_rn(fmaf(yl[0], 1.0f / 2048.0f, yh[0]), inv_so);
                part.y = __fmul_rn(fmaf(yl[1], 1.0f / 2048.0f, yh[1]), inv_so);
                part.z = __fmul_rn(fmaf(yl[2], 1.0f / 2048.0f, yh[2]), inv_so);
                part.w = __fmul_rn(fmaf(yl[3], 1.0f / 2048.0f, yh[3]), inv_so);
                reinterpret_cast<float4*>(s_yp)[((t & 1) * 8 + warp) * 32 + lane] = part;
            }
            if (t < T) {
                float ch[4] = {0.f, 0.f, 0.f, 0.f}, cl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kt = 0; kt < 8; ++kt) {
                    mma_f16(ch, ah[kt], bfr[kt]);
                    mma_f16(cl, al[kt], bfr[kt]);
                }
                uint32_t zpk[2];
#pragma unroll
                for (int nh = 0; nh < 2; ++nh) {
                    __half zh[2];
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const int e = 2 * nh + rh;
                        const float rec = __fmul_rn(fmaf(cl[e], 1.0f / 2048.0f, ch[e]), inv_s);
                        // V' = (alpha V + I_in + I_rec)(1 - Z.detach())          spiking_layers.py:169/239
                        const float t1 = __fmul_rn(p.alpha, v[e]);
                        const float t2 = __fadd_rn(t1, cur[e]);
                        const float t3 = __fadd_rn(t2, rec);
                        const float vn = __fmul_rn(t3, __fsub_rn(1.0f, zp[e]));
                        float thr = p.theta;
                        if constexpr (ALIF) {
                            a[e] = __fadd_rn(__fmul_rn(p.rho, a[e]), zp[e]);          // :240
                            thr = __fadd_rn(p.theta, __fmul_rn(beta, a[e]));          // :241
                        }
                        const float zn = vn >= thr ? 1.0f : 0.0f;                     // spike_funcs.py:27-28
                        if (p.traces && ok[e]) {
                            const size_t o = o_base + (size_t)t * H + rh * o_row + 8 * nh;
                            p.V[o] = vn;
                            p.Z[o] = zn;
                            if constexpr (ALIF) p.a[o] = a[e];
                        }
                        v[e] = vn;
                        zp[e] = zn;
                        zh[rh] = __float2half_rn(zn);
                    }
                    zpk[nh] = pack_h2(zh[0], zh[1]);
                }
                // publish Z_t: tile[t & 1][row n][i0 + 8 j + m]  (matrix j: lanes 8 j .. 8 j + 7 give the row addresses)
                const uint32_t sa = tc::smem_u32(s_z) + (uint32_t)(((t & 1) * kTcRows + (lane & 7)) * kTcTileStride + i0 +
                                                                   8 * ((lane >> 3) & 1)) * 2;
                stsm_x2_trans(sa, zpk[0], zpk[1]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) { cur[e] = nx1[e]; nx1[e] = nx2[e]; }
            __syncthreads();
        }
    } else {
        // ---------------- service warp: bit-packed raster, readout scan + max over time ----------------
        // (two block_max_tc calls above contain __syncthreads: take part in them)
        block_max_tc(0.f, s_red);
        block_max_tc(0.f, s_red);
        float yv[4], mx[4], bias[4];
        int mt[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int cls = g + 8 * (e >> 1);
            yv[e] = 0.f; mx[e] = 0.f; mt[e] = 0;
            bias[e] = cls < O ? __ldg(p.b_out + cls) : 0.f;
        }
        auto finish_y = [&](int ty, int buf) {      // y_ty = kappa y_{ty-1} + sum of the 8 warps' partials + b
            float4 acc = reinterpret_cast<const float4*>(s_yp)[(buf * 8 + 0) * 32 + lane];
#pragma unroll
            for (int w = 1; w < 8; ++w) {
                const float4 q = reinterpret_cast<const float4*>(s_yp)[(buf * 8 + w) * 32 + lane];
                acc.x = __fadd_rn(acc.x, q.x); acc.y = __fadd_rn(acc.y, q.y);
                acc.z = __fadd_rn(acc.z, q.z); acc.w = __fadd_rn(acc.w, q.w);
            }
            const float s4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cls = g + 8 * (e >> 1), row = 2 * tig + (e & 1);
                const float y = __fadd_rn(__fadd_rn(__fmul_rn(p.kappa, yv[e]), s4[e]), bias[e]);     // spiking_layers.py:407
                yv[e] = y;
                if (ty == 0 || y > mx[e]) { mx[e] = y; mt[e] = ty; }                                 // first max wins (snn.py:228)
                if (cls < O && b0 + row < B) p.y[((size_t)(b0 + row) * T + ty) * O + cls] = y;
            }
        };
        __syncthreads();
        for (int t = 0; t <= T; ++t) {
            if (t >= 1) {
                // bit-packed raster of step t - 1 from its fp16 tile (1.0 = 0x3C00: bit 13 of each half)
                const int row = lane >> 2, qd = lane & 3;
                const uint4* src = reinterpret_cast<const uint4*>(s_z + (((t + 1) & 1) * kTcRows + row) * kTcTileStride + 32 * qd);
                uint32_t word = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 h = src[q];
                    const uint32_t xs[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        word |= (((xs[j] >> 13) & 1u) | ((xs[j] >> 28) & 2u)) << (8 * q + 2 * j);
                }
                if (b0 + row < B) p.zbits[((size_t)(b0 + row) * T + (t - 1)) * (H / 32) + qd] = word;
            }
            if (t >= 2) finish_y(t - 2, (t - 1) & 1);
            __syncthreads();
        }
        finish_y(T - 1, T & 1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int cls = g + 8 * (e >> 1), row = 2 * tig + (e & 1);
            if (cls < O && b0 + row < B) {
                p.logits[(size_t)(b0 + row) * O + cls] = mx[e];
                p.tstar[(size_t)(b0 + row) * O + cls] = mt[e];
            }
        }
    }
}

// ---- readout adjoint scan (pre-pass of the tensor-core sweep) ----------------------------------------------------------
// gy_t = seed_t + kappa gy_{t+1}  (spiking_layers.py:407 backwards) for every (row, class) -> gy_scan (B, T, kOMax), zero
// in the padded classes.  It does not depend on the recurrence, so it runs before the sweep; dW_out and db are
// contractions of it with the spike raster (k_wout_grad, recur_gen.cuh) and run BESIDE the sweep on the idle SMs.
__global__ void __launch_bounds__(256) k_gy_scan(int B, int T, int O, float kappa, const float* __restrict__ g_y,
                                                const float* __restrict__ g_logits, const int32_t* __restrict__ tstar,
                                                const float* __restrict__ g_scale, float* __restrict__ gy_scan)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = idx / kOMax, c = idx - b * kOMax;
    if (b >= B) return;
    float* out = gy_scan + (size_t)b * T * kOMax + c;
    if (c >= O) {
        for (int t = 0; t < T; ++t) out[(size_t)t * kOMax] = 0.f;
        return;
    }
    float g = 0.f;
    if (g_y) {
        const float* src = g_y + (size_t)b * T * O + c;
        for (int t = T - 1; t >= 0; --t) {
            g = __fadd_rn(__ldg(src + (size_t)t * O), __fmul_rn(kappa, g));
            out[(size_t)t * kOMax] = g;
        }
    } else {
        const float scale = g_scale ? __ldg(g_scale) : 1.0f;
        const int ts = __ldg(tstar + (size_t)b * O + c);
        const float seed = __fmul_rn(__ldg(g_logits + (size_t)b * O + c), scale);
        for (int t = T - 1; t >= 0; --t) {
            g = __fadd_rn(t == ts ? seed : 0.f, __fmul_rn(kappa, g));
            out[(size_t)t * kOMax] = g;
        }
    }
}

// ---- backward -------------------------------------------------------------------------------------------------------
constexpr size_t bwd_tc_smem_bytes(int T)
{
    return sizeof(__half) * 2 * 2 * kTcRows * kTcTileStride                                 // gI tiles [buf][plane][8][136]
           + sizeof(__half) * 2 * (size_t)T * kTcRows * 16                                  // gy planes [plane][T][8][16]
           + sizeof(uint32_t) * (size_t)kTcRows * (((T + 1) * (kTcH / 32) + 3) & ~3)         // spike words [8][T+1][4] (slot 0: Z_{-1})
           + sizeof(uint32_t) * (size_t)kTcRows * ((T + 31) / 32 + 1)                        // run-start bits
           + sizeof(float) * 16;
}

// grid = ceil(B / 8), block = 288.  gy_scan: output of k_gy_scan.  Writes gI (one or two tf32 planes) and, with a
// frame-run table, the run sums; dW_out / db are NOT produced here (k_wout_grad).
template <bool ALIF, int SURR>
__global__ void __launch_bounds__(kTcThreads, 1) k_recur_bwd_tc(const BwdParams p, const float* __restrict__ gy_scan)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int H = kTcH, W32 = kTcH / 32;
    const int T = p.T, B = p.B, O = p.O;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int b0 = blockIdx.x * kTcRows;
    const int nvalid = min(kTcRows, B - b0);
    const int zstride = ((T + 1) * W32 + 3) & ~3;      // words per row of the spike-word table
    const int TW = (T + 31) / 32 + 1;

    __half* s_g = reinterpret_cast<__half*>(smem_raw);                                       // [2][2][8][136]
    __half* s_gy = s_g + 2 * 2 * kTcRows * kTcTileStride;                                    // [2][T][8][16]
    uint32_t* s_zw = reinterpret_cast<uint32_t*>(s_gy + 2 * (size_t)T * kTcRows * 16);       // [8][zstride]
    uint32_t* s_start = s_zw + (size_t)kTcRows * zstride;                                    // [8][TW]
    float* s_red = reinterpret_cast<float*>(s_start + kTcRows * TW);                         // [12] + exponent-class words [2]
    uint32_t* s_cls = reinterpret_cast<uint32_t*>(s_red + 12);

    const bool run_sums = p.run_table != nullptr && p.run_table[1] == 1;
    if (tid == 256) s_cls[0] = s_cls[1] = 0u;
    // gradient tiles start at zero (gI_T = 0)
    for (int idx = tid; idx < 2 * 2 * kTcRows * kTcTileStride; idx += kTcThreads) s_g[idx] = __float2half_rn(0.f);
    // spike words: slot 0 of a row is Z_{-1} (initial state), slot t + 1 is Z_t
    for (int idx = tid; idx < kTcRows * (T + 1) * W32; idx += kTcThreads) {
        const int r = idx / ((T + 1) * W32), rem = idx - r * ((T + 1) * W32);
        const int ts = rem / W32, wd = rem - ts * W32;
        uint32_t w = 0u;
        if (b0 + r < B) {
            if (ts > 0) w = __ldg(p.zbits + ((size_t)(b0 + r) * T + ts - 1) * W32 + wd);
            else if (p.Z0) {
                for (int l = 0; l < 32; ++l)
                    if (__ldg(p.Z0 + (size_t)(b0 + r) * H + wd * 32 + l) != 0.f) w |= 1u << l;
            }
        }
        s_zw[r * zstride + rem] = w;
    }
    if (run_sums) {
        for (int idx = tid; idx < kTcRows * TW; idx += kTcThreads) s_start[idx] = 0u;
    }
    __syncthreads();
    if (run_sums) {
        // bit t of a row's word: step t is the first of its run of equal input frames
        for (int idx = tid; idx < nvalid * T; idx += kTcThreads) {
            const int r = idx / T, t = idx - r * T;
            const int* rc = p.run_table + kRunHdrInts + (size_t)(b0 + r) * T;
            if (t == 0 || __ldg(rc + t) != __ldg(rc + t - 1)) atomicOr(s_start + r * TW + (t >> 5), 1u << (t & 31));
        }
        if (blockIdx.x == 0) {   // the weight-gradient GEMM contracts whole 32-row blocks: zero the tail of the last one
            const int n_rows = p.run_table[0], n_pad = (n_rows + 31) & ~31;
            for (int idx = tid; idx < (n_pad - n_rows) * H; idx += kTcThreads) {
                p.Gu_hi[(size_t)n_rows * H + idx] = 0.f;
                p.Gu_lo[(size_t)n_rows * H + idx] = 0.f;
            }
        }
    }
    // readout adjoint of the tile as two fp16 planes: B operand (k = class, n = row) of every step
    float gymax = 0.f;
    for (int idx = tid; idx < nvalid * T * kOMax; idx += kTcThreads) gymax = fmaxf(gymax, fabsf(__ldg(gy_scan + (size_t)b0 * T * kOMax + idx)));
    gymax = block_max_tc(gymax, s_red);
    const float s_gyscale = pow2_scale_for(gymax);
    for (int idx = tid; idx < kTcRows * T * kOMax; idx += kTcThreads) {
        const int r = idx / (T * kOMax), rem = idx - r * (T * kOMax);
        const int t = rem / kOMax, c = rem - t * kOMax;
        __half hi = __float2half_rn(0.f), lo = hi;
        if (r < nvalid) split_h2(__fmul_rn(__ldg(gy_scan + (size_t)(b0 + r) * T * kOMax + rem), s_gyscale), hi, lo);
        s_gy[(t * kTcRows + r) * 16 + c] = hi;
        s_gy[((size_t)T * kTcRows + t * kTcRows + r) * 16 + c] = lo;
    }

    constexpr uint32_t kTileBytes = kTcRows * kTcTileStride * 2;       // one plane of one buffer
    const uint32_t tile_base = tc::smem_u32(s_g) + (uint32_t)((lane & 7) * kTcTileStride + 8 * (lane >> 3)) * 2;
    // ldmatrix.x4 source inside the gy planes: matrices (hi k 0-7, hi k 8-15, lo k 0-7, lo k 8-15)
    const uint32_t gy_base = tc::smem_u32(s_gy) + (uint32_t)(((lane >> 4) * T * kTcRows + (lane & 7)) * 16 + 8 * ((lane >> 3) & 1)) * 2;

    if (warp < 8) {
        const int i0 = 16 * warp;
        uint32_t ah[8][4], al[8][4];          // rows i0.. of W_eff as A fragments: A[m][k] = W_eff[i0 + m][k] = W_effT[k][i0 + m]
        uint32_t oh[4], ol[4];                // A[m][k = class] = W_out[i0 + m][k]
        float inv_sw, inv_so;
        {
            float raw[8][8];
            load_a_raw<8>(p.W_effT, H, i0, H, g, tig, raw);
            const float s = pow2_scale_for(block_max_tc(absmax_raw<8>(raw), s_red));
            inv_sw = __fdiv_rn(1.0f, s);
            split_a<8>(raw, s, ah, al);
            float ro[1][8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = 2 * tig + (q & 1) + 8 * (q >> 2), m = g + 8 * ((q >> 1) & 1);
                ro[0][q] = k < O ? __ldg(p.W_out + (size_t)(i0 + m) * O + k) : 0.f;
            }
            const float so = pow2_scale_for(block_max_tc(absmax_raw<1>(ro), s_red));
            inv_so = __fdiv_rn(__fdiv_rn(1.0f, so), s_gyscale);       // both scales of the readout-adjoint product
            uint32_t th[1][4], tl[1][4];
            split_a<1>(ro, so, th, tl);
#pragma unroll
            for (int r = 0; r < 4; ++r) { oh[r] = th[0][r]; ol[r] = tl[0][r]; }
        }
        const float beta = (ALIF && p.beta) ? __ldg(p.beta) : 0.f;
        float gv[4], racc[4];
        int crow[2];
        uint32_t sbits[2];
        bool ok[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { gv[e] = 0.f; racc[e] = 0.f; ok[e] = b0 + 2 * tig + (e & 1) < B; }
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            crow[rh] = (run_sums && ok[rh]) ? __ldg(p.run_table + kRunHdrInts + (size_t)(b0 + 2 * tig + rh) * T + T - 1) : 0;
            sbits[rh] = 0u;
        }
        const size_t o_base = ((size_t)(b0 + 2 * tig) * T) * H + i0 + g;
        const size_t o_row = (size_t)T * H;
        const int zword = warp >> 1, zsh = 16 * (warp & 1) + g;      // this thread's neurons in a spike word: bits zsh, zsh + 8
        float inv_sg = 1.0f;      // 1 / scale of the gradient tile being READ (gI_{t+1}); the first tile is all zero
        __syncthreads();          // planes and tables visible
        // saved traces V_t (and a_t): straight from global memory into registers, two steps ahead (see k_recur_fwd_tc)
        auto load_va = [&](int tl, float (&dv)[4], float (&da)[4]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dv[e] = 0.f; da[e] = 0.f;
                if (tl >= 0 && ok[e]) {
                    const size_t o = o_base + (size_t)tl * H + (e & 1) * o_row + 8 * (e >> 1);
                    dv[e] = __ldg(p.V + o);
                    if constexpr (ALIF) da[e] = __ldg(p.a + o);
                }
            }
        };
        float vcur[4], acur[4], vn1[4], an1[4], vn2[4], an2[4];
        load_va(T - 1, vcur, acur);
        load_va(T - 2, vn1, an1);

        for (int t = T - 1; t >= 0; --t) {
            load_va(t - 2, vn2, an2);
            // ---- gZ^T = W_eff gI_{t+1}^T  +  W_out gy_t^T ----
            float chh[4] = {0.f, 0.f, 0.f, 0.f}, chl[4] = {0.f, 0.f, 0.f, 0.f}, clh[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t tb = tile_base + ((t + 1) & 1) * 2 * kTileBytes;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t rh4[4], rl4[4];
                ldsm_x4(rh4, tb + q * 64);
                ldsm_x4(rl4, tb + kTileBytes + q * 64);
                const uint32_t bh0[2] = {rh4[0], rh4[1]}, bh1[2] = {rh4[2], rh4[3]};
                const uint32_t bl0[2] = {rl4[0], rl4[1]}, bl1[2] = {rl4[2], rl4[3]};
                mma_f16(chh, ah[2 * q], bh0);
                mma_f16(chl, ah[2 * q], bl0);
                mma_f16(clh, al[2 * q], bh0);
                mma_f16(chh, ah[2 * q + 1], bh1);
                mma_f16(chl, ah[2 * q + 1], bl1);
                mma_f16(clh, al[2 * q + 1], bh1);
            }
            float ohh[4] = {0.f, 0.f, 0.f, 0.f}, ox[4] = {0.f, 0.f, 0.f, 0.f};
            {
                uint32_t r4[4];
                ldsm_x4(r4, gy_base + (uint32_t)(t * kTcRows * 16) * 2);
                const uint32_t byh[2] = {r4[0], r4[1]}, byl[2] = {r4[2], r4[3]};
                mma_f16(ohh, oh, byh);
                mma_f16(ox, oh, byl);
                mma_f16(ox, ol, byh);
            }
            const float sc_rec = __fmul_rn(inv_sw, inv_sg);
            float gi[4];
            uint32_t cls = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int nh = e >> 1, rh = e & 1, row = 2 * tig + rh;
                const float rec = __fmul_rn(fmaf(__fadd_rn(chl[e], clh[e]), 1.0f / 2048.0f, chh[e]), sc_rec);
                const float rdo = __fmul_rn(fmaf(ox[e], 1.0f / 2048.0f, ohh[e]), inv_so);
                float s = __fadd_rn(rdo, rec);
                const size_t o = o_base + (size_t)t * H + rh * o_row + 8 * nh;
                if (p.g_Z && ok[e]) s = __fadd_rn(s, __ldg(p.g_Z + o));
                const float vt = vcur[e];
                float thr = p.theta;
                if constexpr (ALIF) thr = __fadd_rn(p.theta, __fmul_rn(beta, acur[e]));
                const uint32_t wt = s_zw[row * zstride + (t + 1) * W32 + zword], wp = s_zw[row * zstride + t * W32 + zword];
                const float zt = (float)((wt >> (zsh + 8 * nh)) & 1u), zprev = (float)((wp >> (zsh + 8 * nh)) & 1u);
                const float sg = surrogate_grad(SURR, p.gamma, vt, thr);
                const float carry = __fmul_rn(__fmul_rn(p.alpha, gv[e]), __fsub_rn(1.0f, zt));
                float gq = __fadd_rn(__fmul_rn(s, sg), carry);
                if (p.g_V && ok[e]) gq = __fadd_rn(gq, __ldg(p.g_V + o));
                gv[e] = gq;
                gi[e] = ok[e] ? __fmul_rn(gq, __fsub_rn(1.0f, zprev)) : 0.f;
                if (ok[e]) {
                    if (p.gI_lo) {      // exact two-plane tf32 split for the weight-gradient GEMM
                        const float hi = __uint_as_float(__float_as_uint(gi[e]) & 0xFFFFE000u);
                        p.gI[o] = hi;
                        p.gI_lo[o] = __fsub_rn(gi[e], hi);
                    } else {
                        p.gI[o] = gi[e];
                    }
                }
                cls |= gi[e] != 0.f ? 1u << (((__float_as_uint(gi[e]) >> 23) & 0xFFu) >> 3) : 0u;
                if (run_sums) {      // sum of gI over the run of equal input frames this step belongs to
                    racc[e] = __fadd_rn(racc[e], gi[e]);
                    if (nh == 0 && (t == T - 1 || (t & 31) == 31)) sbits[rh] = s_start[row * TW + (t >> 5)];
                    if ((sbits[rh] >> (t & 31)) & 1u) {
                        if (ok[e]) {
                            const float hi = __uint_as_float(__float_as_uint(racc[e]) & 0xFFFFE000u);
                            const size_t ro = (size_t)crow[rh] * H + i0 + g + 8 * nh;
                            p.Gu_hi[ro] = hi;
                            p.Gu_lo[ro] = __fsub_rn(racc[e], hi);
                        }
                        racc[e] = 0.f;
                        if (nh == 1) --crow[rh];
                    }
                }
            }
            // tile-wide scale for gI_t from the largest exponent class present (classes of 8 binades): warp OR
            // (redux.sync), one shared-memory atomicOr per warp, the barrier, one read
            const uint32_t wor = __reduce_or_sync(0xffffffffu, cls);
            if (lane == 0 && wor) atomicOr(s_cls + (t & 1), wor);
            __syncthreads();
            const uint32_t mask = s_cls[t & 1];
            if (tid == 0) s_cls[(t + 1) & 1] = 0u;      // the word of step t - 1 (last read before the barrier that ended step t + 1)
            float sg_new = 1.0f;
            if (mask) {
                const int top = 31 - __clz(mask);                  // exponents 8 top .. 8 top + 7: |gI| < 2^(8 top + 8 - 127)
                int kexp = 133 - 8 * top;                          // scaled maximum < 2^14
                kexp = kexp > 126 ? 126 : kexp;
                sg_new = __uint_as_float((uint32_t)(kexp + 127) << 23);
            }
            uint32_t pk[4];
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
                __half h0, l0, h1, l1;
                split_h2(__fmul_rn(gi[2 * nh], sg_new), h0, l0);
                split_h2(__fmul_rn(gi[2 * nh + 1], sg_new), h1, l1);
                pk[nh] = pack_h2(h0, h1);
                pk[2 + nh] = pack_h2(l0, l1);
            }
            // matrices: (hi, neurons i0..+7), (hi, i0+8..), (lo, i0..), (lo, i0+8..); lane 8 j + n gives row n of matrix j
            const uint32_t sa = tc::smem_u32(s_g) + (uint32_t)((((t & 1) * 2 + (lane >> 4)) * kTcRows + (lane & 7)) * kTcTileStride +
                                                              i0 + 8 * ((lane >> 3) & 1)) * 2;
            stsm_x4_trans(sa, pk[0], pk[1], pk[2], pk[3]);
            inv_sg = __fdiv_rn(1.0f, sg_new);
#pragma unroll
            for (int e = 0; e < 4; ++e) { vcur[e] = vn1[e]; acur[e] = an1[e]; vn1[e] = vn2[e]; an1[e] = an2[e]; }
            __syncthreads();
        }
    } else {
        // ninth warp: only takes part in the barriers (the block-wide reductions of the prologue count 288 threads)
        block_max_tc(0.f, s_red);
        block_max_tc(0.f, s_red);
        __syncthreads();
        for (int t = T - 1; t >= 0; --t) {
            __syncthreads();
            __syncthreads();
        }
    }
}

}  // namespace snnk
