// gemm_simt.cuh -- fp32 CUDA-core GEMMs for the two contractions of the path.
//
// These are the always-exact kernels: every output is one fp32 FMA chain in ascending reduction index,
// which is the order oracle/snn_oracle.c fixes, so the projection is bit-identical to the oracle.  They
// serve (a) any input that is not exactly representable on the tensor-core path and (b) as the on-GPU
// check of the tcgen05 kernels.
//
//   k_proj_simt   I_in[r][n]  = sum_k x[r][k] W_in[k][n]          r = (b,t)    (spiking_layers.py:163/233, all T at once)
//   k_wgrad_simt  dW_in[m][n] = sum_r x[r][m] gI[r][n]                         (MmBackward of the same matmul)
//                 dW_rec[j][n]= sum_r Z_{t-1}[r][j] gI[r][n]                   (MmBackward of :165/235)
#pragma once
#include "common.cuh"

namespace snnk {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 16;
constexpr int kGemmThreads = 256;

// ---- projection: C[M x Hn] = A[M x K] (row-major) * W[K x Hn] (row-major) --------------------------------------
template <int BN>
__global__ void __launch_bounds__(kGemmThreads) k_proj_simt(const float* __restrict__ A,
                                                           const float* __restrict__ W,
                                                           float* __restrict__ C, int M, int K, int Hn,
                                                           const unsigned int* __restrict__ run_if_flag)
{
    // behind the tcgen05 kernel this launch is a fallback: it only runs when that kernel raised the flag
    if (run_if_flag && *run_if_flag == 0) return;
    constexpr int TN = BN / 4;                  // thread columns
    constexpr int TMG = kGemmThreads / TN;      // thread row groups
    constexpr int TM = kGemmBM / TMG;           // rows per thread
    constexpr int LDA = kGemmBM + 4;
    __shared__ __align__(16) float As[kGemmBK][LDA];
    __shared__ __align__(16) float Bs[kGemmBK][BN];

    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * BN;
    const int tn = tid % TN, tm = tid / TN;

    float acc[TM][4];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    const int arow = tid >> 1, akh = (tid & 1) * 8;     // A tile loader: 128 rows x 16 k, 8 k per thread
    for (int k0 = 0; k0 < K; k0 += kGemmBK) {
        {
            const int r = r0 + arow;
            const float* src = A + (size_t)r * K + k0 + akh;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + akh + j;
                As[akh + j][arow] = (r < M && k < K) ? __ldg(src + j) : 0.f;
            }
        }
        for (int idx = tid; idx < kGemmBK * BN; idx += kGemmThreads) {
            const int kk = idx / BN, nn = idx - kk * BN;
            const int k = k0 + kk;
            Bs[kk][nn] = (k < K) ? __ldg(W + (size_t)k * Hn + n0 + nn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGemmBK; ++kk) {
            float av[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm * TM + 4 * q]);
                av[4 * q + 0] = a4.x; av[4 * q + 1] = a4.y; av[4 * q + 2] = a4.z; av[4 * q + 3] = a4.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
#pragma unroll
            for (int a = 0; a < TM; ++a) {
                acc[a][0] = fmaf(av[a], b4.x, acc[a][0]);
                acc[a][1] = fmaf(av[a], b4.y, acc[a][1]);
                acc[a][2] = fmaf(av[a], b4.z, acc[a][2]);
                acc[a][3] = fmaf(av[a], b4.w, acc[a][3]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        const int r = r0 + tm * TM + a;
        if (r < M)
            *reinterpret_cast<float4*>(C + (size_t)r * Hn + n0 + tn * 4) =
                make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }
}

// ---- input gradient (stacked hidden layers): gX[r][n] = sum_i gI[r][i] * W_in[n][i] ----------------------------
// MmBackward of spiking_layers.py:163/233 w.r.t. its input: what layer l+1 hands down to the spikes of layer l.
// A = gI (M x K, K = H of this layer), W_in is (Nn x K) row-major and used transposed; any Nn (guarded).
__global__ void __launch_bounds__(kGemmThreads) k_input_grad_simt(const float* __restrict__ A, const float* __restrict__ W_in,
                                                                 float* __restrict__ C, int M, int K, int Nn)
{
    constexpr int BN = 64, TN = BN / 4, TMG = kGemmThreads / TN, TM = kGemmBM / TMG, LDA = kGemmBM + 4;
    __shared__ __align__(16) float As[kGemmBK][LDA];
    __shared__ __align__(16) float Bs[kGemmBK][BN];
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * BN;
    const int tn = tid % TN, tm = tid / TN;
    float acc[TM][4];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int arow = tid >> 1, akh = (tid & 1) * 8;
    for (int k0 = 0; k0 < K; k0 += kGemmBK) {
        {
            const int r = r0 + arow;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + akh + j;
                As[akh + j][arow] = (r < M && k < K) ? __ldg(A + (size_t)r * K + k) : 0.f;
            }
        }
        for (int idx = tid; idx < kGemmBK * BN; idx += kGemmThreads) {
            const int nn = idx / kGemmBK, kk = idx - nn * kGemmBK;      // k fastest: coalesced along a row of W_in
            const int k = k0 + kk, n = n0 + nn;
            Bs[kk][nn] = (k < K && n < Nn) ? __ldg(W_in + (size_t)n * K + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGemmBK; ++kk) {
            float av[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm * TM + 4 * q]);
                av[4 * q + 0] = a4.x; av[4 * q + 1] = a4.y; av[4 * q + 2] = a4.z; av[4 * q + 3] = a4.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
#pragma unroll
            for (int a = 0; a < TM; ++a) {
                acc[a][0] = fmaf(av[a], b4.x, acc[a][0]);
                acc[a][1] = fmaf(av[a], b4.y, acc[a][1]);
                acc[a][2] = fmaf(av[a], b4.z, acc[a][2]);
                acc[a][3] = fmaf(av[a], b4.w, acc[a][3]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        const int r = r0 + tm * TM + a;
        if (r >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + tn * 4 + b;
            if (n < Nn) C[(size_t)r * Nn + n] = acc[a][b];
        }
    }
}

// ---- weight gradients: split-K  C_part[s][m][n] = sum_{r in split s} A[r][m] * G[r][n] -------------------------
struct WgradParams {
    int BT, T, N, H;            // rows r = b*T + t
    int mtiles_x;               // ceil(N / 128) tiles take A from x, the rest from the spike raster
    int rows_per_split;
    const float* x;             // (BT, N)
    const uint32_t* zbits;      // (BT, H/32)
    const float* Z0;            // (B, H) or null
    const float* gI;            // (BT, H)
    const float* gI_lo;         // (BT, H) low tf32 plane of gI (tensor-core mode: gI holds the high plane) or null
    const unsigned int* run_if_flag;   // fallback gating, see k_proj_simt
    float* part;                // [S][N + H][H]   (rows N.. only when recurrent)
    int m_total;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads) k_wgrad_simt(const WgradParams p)
{
    constexpr int TN = BN / 4;
    constexpr int TMG = kGemmThreads / TN;
    constexpr int TM = kGemmBM / TMG;
    __shared__ __align__(16) float As[kGemmBK][kGemmBM];
    __shared__ __align__(16) float Bs[kGemmBK][BN];

    if (p.run_if_flag && *p.run_if_flag == 0) return;
    const int tid = threadIdx.x;
    const bool from_x = (int)blockIdx.x < p.mtiles_x;
    const int m0 = from_x ? blockIdx.x * kGemmBM : (blockIdx.x - p.mtiles_x) * kGemmBM;
    const int mlim = from_x ? p.N : p.H;
    const int n0 = blockIdx.y * BN;
    const int rbeg = blockIdx.z * p.rows_per_split;
    const int rend = min(rbeg + p.rows_per_split, p.BT);
    const int tn = tid % TN, tm = tid / TN;
    const int W32 = p.H / 32;

    float acc[TM][4];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int r0 = rbeg; r0 < rend; r0 += kGemmBK) {
        for (int idx = tid; idx < kGemmBK * kGemmBM; idx += kGemmThreads) {
            const int kk = idx / kGemmBM, mm = idx - kk * kGemmBM;
            const int r = r0 + kk, m = m0 + mm;
            float v = 0.f;
            if (r < rend && m < mlim) {
                if (from_x) {
                    v = __ldg(p.x + (size_t)r * p.N + m);
                } else {
                    const int t = r % p.T;
                    if (t > 0) v = (float)((__ldg(p.zbits + (size_t)(r - 1) * W32 + (m >> 5)) >> (m & 31)) & 1u);
                    else v = p.Z0 ? __ldg(p.Z0 + (size_t)(r / p.T) * p.H + m) : 0.f;
                }
            }
            As[kk][mm] = v;
        }
        for (int idx = tid; idx < kGemmBK * BN; idx += kGemmThreads) {
            const int kk = idx / BN, nn = idx - kk * BN;
            const int r = r0 + kk;
            float g = 0.f;
            if (r < rend) {
                g = __ldg(p.gI + (size_t)r * p.H + n0 + nn);
                if (p.gI_lo) g += __ldg(p.gI_lo + (size_t)r * p.H + n0 + nn);   // hi + lo is exact
            }
            Bs[kk][nn] = g;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGemmBK; ++kk) {
            float av[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm * TM + 4 * q]);
                av[4 * q + 0] = a4.x; av[4 * q + 1] = a4.y; av[4 * q + 2] = a4.z; av[4 * q + 3] = a4.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
#pragma unroll
            for (int a = 0; a < TM; ++a) {
                acc[a][0] = fmaf(av[a], b4.x, acc[a][0]);
                acc[a][1] = fmaf(av[a], b4.y, acc[a][1]);
                acc[a][2] = fmaf(av[a], b4.z, acc[a][2]);
                acc[a][3] = fmaf(av[a], b4.w, acc[a][3]);
            }
        }
        __syncthreads();
    }
    const int mbase = from_x ? 0 : p.N;
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        const int m = m0 + tm * TM + a;
        if (m < mlim)
            *reinterpret_cast<float4*>(p.part + ((size_t)blockIdx.z * p.m_total + mbase + m) * p.H + n0 + tn * 4) =
                make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }
}

}  // namespace snnk
