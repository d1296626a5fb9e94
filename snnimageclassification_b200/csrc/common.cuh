// common.cuh -- shared device helpers and kernel parameter blocks for libsnnk.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace snnk {

constexpr int kOMax = 16;   // readout units are padded to 16 lanes in shared memory / registers

struct FwdParams {
    int B, T, H, O;
    int alif, traces;
    float alpha, rho, theta, kappa;
    const float* I_in;      // (B,T,H) input current from the projection GEMM
    const float* W_rec;     // (H,H) raw
    const float* rec_mask;  // (H,H) or null
    const float* beta;      // device scalar or null
    const float* W_out;     // (H,O)
    const float* b_out;     // (O)
    const float* V0; const float* a0; const float* Z0;   // (B,H) or null
    float* V; float* a; float* Z;                        // (B,T,H) traces (traces != 0)
    uint32_t* zbits;        // (B,T,H/32)
    float* y;               // (B,T,O)
    float* logits;          // (B,O)
    int32_t* tstar;         // (B,O)
};

struct BwdParams {
    int B, T, H, O;
    int alif, surrogate;
    float alpha, theta, gamma, kappa;
    const float* W_rec; const float* rec_mask; const float* beta; const float* W_out;
    const float* Z0;
    const float* V; const float* a; const uint32_t* zbits;
    const float* g_y;                               // (B,T,O) dense seeds, or null
    const float* g_logits; const int32_t* tstar;    // (B,O) sparse seeds, or null
    const float* g_V; const float* g_Z;             // optional (B,T,H) seeds
    float* gI;          // (B,T,H)
    float* gI_lo;       // (B,T,H) or null: when set, gI receives trunc_tf32(gI) and gI_lo the exact remainder
    float* part_wout;   // [grid][H][O]
    float* part_db;     // [grid*R][O]
};

// Surrogate derivatives, spike_funcs.py:59-62 (FastSigmoid) and :75-79 (Phi, epsilon = 1e-5).
__device__ __forceinline__ float surrogate_grad(int kind, float gamma, float v, float thr)
{
    if (kind == 0) {
        float d = __fadd_rn(__fmul_rn(gamma, fabsf(__fsub_rn(v, thr))), 1.0f);
        return __fdiv_rn(1.0f, __fmul_rn(d, d));
    }
    float te = __fadd_rn(thr, 1e-5f);
    float r = __fsub_rn(1.0f, fabsf(__fdiv_rn(__fsub_rn(v, thr), te)));
    r = r < 0.f ? 0.f : r;
    return __fmul_rn(__fdiv_rn(gamma, te), r);
}

// sum_k w[k] * z[k] with four accumulators over k mod 4 (the order oracle/snn_oracle.c fixes).
// The broadcast vector is fetched NB float4 at a time before the FFMAs that use them, so NB shared-memory
// loads are in flight together instead of one 29-cycle LDS round trip per 4 FFMAs.
template <int H, int NB = 4>
__device__ __forceinline__ float dot_rec4(const float (&w)[H], const float4* __restrict__ zv)
{
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    constexpr int NV = H / 4;
    constexpr int BATCH = NB < NV ? NB : NV;
#pragma unroll
    for (int k0 = 0; k0 < NV; k0 += BATCH) {
        float4 z[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) z[j] = zv[k0 + j];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int k = 4 * (k0 + j);
            s0 = fmaf(w[k + 0], z[j].x, s0);
            s1 = fmaf(w[k + 1], z[j].y, s1);
            s2 = fmaf(w[k + 2], z[j].z, s2);
            s3 = fmaf(w[k + 3], z[j].w, s3);
        }
    }
    return __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
}

}  // namespace snnk
