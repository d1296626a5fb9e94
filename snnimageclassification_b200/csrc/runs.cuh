// runs.cuh -- frame-dedup fast path of the input projection and its weight gradient.
//
// The production encoder (ToSpikes with tau = 0.02, datasets.py:21, :72-86) emits rasters whose frames repeat over
// long runs of time steps (periods are 1 or n_steps-1: at most three distinct frames per sample).  Consecutive equal
// frames (b, t-1) == (b, t) have equal input currents, so
//   forward : I_in is computed for the first row of every run only (compact rows) and copied to the rest of the run;
//   backward: dW_in = sum_r x_r^T gI_r = sum_runs x_run^T (sum_{r in run} gI_r)   -- the contraction shrinks from
//             B*T rows to the number of runs; dW_rec still contracts over all rows (the spike trace does not repeat).
// Which rows repeat is data: the encoder records it (changed[b][t], k_encode), k_frame_runs turns it into the run
// table below, and every kernel of both variants (dedup and dense) is launched and gated on the table's `ok` word
// on the DEVICE, so the choice needs no host synchronisation and the step stays capturable in one CUDA graph.
//
// Run table (int32): [0] n_rows (number of runs in the batch)  [1] ok (1: n_rows <= cap, use the dedup kernels)
//                    [2] cap  [3] 0   [4 .. 4+B*T) compact row of every dense row b*T+t
//                    then rep[cap] = first dense row of each run, then len[cap] = its length.
#pragma once
#include "common.cuh"

namespace snnk {

constexpr int kRunHdr = 4;
__host__ __device__ inline int run_cap(long long BT)
{
    long long c = (BT / 4 + 127) / 128 * 128;
    return (int)(c < 128 ? 128 : c);
}
__host__ __device__ inline size_t run_table_ints(long long BT) { return (size_t)kRunHdr + (size_t)BT + 2 * (size_t)run_cap(BT); }

// One CTA: each thread owns a contiguous chunk of samples, counts their runs, a block scan gives its first compact
// row, a second pass writes the table.  B*T bytes are read twice; the kernel runs once per encoded batch.
__global__ void __launch_bounds__(1024) k_frame_runs(int B, int T, const unsigned char* __restrict__ changed,
                                                    int* __restrict__ table)
{
    __shared__ int s_cnt[1024];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int per = (B + nthr - 1) / nthr;
    const int b_lo = min(tid * per, B), b_hi = min(b_lo + per, B);
    const int cap = run_cap((long long)B * T);
    int* row2c = table + kRunHdr;
    int* rep = row2c + (size_t)B * T;
    int* len = rep + cap;
    int mine = 0;
    for (int b = b_lo; b < b_hi; ++b) {
        const unsigned char* c = changed + (size_t)b * T;
        int n = 1;
        for (int t = 1; t < T; ++t) n += c[t] != 0;
        mine += n;
    }
    s_cnt[tid] = mine;
    __syncthreads();
    // inclusive Hillis-Steele scan over the threads
    for (int o = 1; o < nthr; o <<= 1) {
        const int v = tid >= o ? s_cnt[tid - o] : 0;
        __syncthreads();
        s_cnt[tid] += v;
        __syncthreads();
    }
    int r = s_cnt[tid] - mine;
    for (int b = b_lo; b < b_hi; ++b) {
        const unsigned char* c = changed + (size_t)b * T;
        int start = 0;
        for (int t = 0; t < T; ++t) {
            if (t > 0 && c[t] != 0) {
                if (r < cap) { rep[r] = b * T + start; len[r] = t - start; }
                ++r;
                start = t;
            }
            row2c[(size_t)b * T + t] = r;
        }
        if (r < cap) { rep[r] = b * T + start; len[r] = T - start; }
        ++r;
    }
    if (tid == nthr - 1) {
        const int total = s_cnt[tid];
        table[0] = total;
        table[1] = total <= cap ? 1 : 0;
        table[2] = cap;
        table[3] = 0;
    }
}

// X_u[r] = X[rep[r]] for r < n_rows; rows up to the next multiple of 32 are zero-filled (the weight-gradient GEMM
// contracts over whole 32-row blocks).  One CTA per compact row, N/4 float4 per row.
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ X, const int* __restrict__ table, int BT,
                                                    int N, float* __restrict__ Xu)
{
    if (table[1] != 1) return;
    const int n_rows = table[0], r = blockIdx.x;
    if (r >= ((n_rows + 31) & ~31)) return;
    float4* dst = reinterpret_cast<float4*>(Xu + (size_t)r * N);
    if (r < n_rows) {
        const int* rep = table + kRunHdr + BT;
        const float4* src = reinterpret_cast<const float4*>(X + (size_t)rep[r] * N);
        for (int i = threadIdx.x; i < N / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < N / 4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// I_in[row] = I_u[compact(row)]: H/4 threads per dense row
__global__ void __launch_bounds__(256) k_expand_rows(const float* __restrict__ Iu, const int* __restrict__ table, int BT,
                                                    int H, float* __restrict__ I_in)
{
    if (table[1] != 1) return;
    const int per_row = H / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = idx / per_row;
    if (row >= BT) return;
    const int q = (int)(idx - row * per_row);
    const int r = table[kRunHdr + row];
    reinterpret_cast<float4*>(I_in + (size_t)row * H)[q] = __ldg(reinterpret_cast<const float4*>(Iu + (size_t)r * H) + q);
}

// G_u[r] = sum over the rows of run r of gI (= hi + lo plane), re-split into two tf32 planes; zero rows up to the
// next multiple of 32.  One thread per (compact row, neuron), ascending t.
__global__ void __launch_bounds__(128) k_run_sum(const float* __restrict__ g_hi, const float* __restrict__ g_lo,
                                                const int* __restrict__ table, int BT, int H, float* __restrict__ Gu_hi,
                                                float* __restrict__ Gu_lo)
{
    if (table[1] != 1) return;
    const int n_rows = table[0], r = blockIdx.x;
    if (r >= ((n_rows + 31) & ~31)) return;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float s = 0.f;
        if (r < n_rows) {
            const int cap = table[2];
            const int* rep = table + kRunHdr + BT;
            const int row = rep[r], n = rep[cap + r];
            const float* ph = g_hi + (size_t)row * H + h;
            const float* pl = g_lo + (size_t)row * H + h;
            int j = 0;
            for (; j + 4 <= n; j += 4) {
                float a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { a[u] = __ldg(ph + (size_t)(j + u) * H); b[u] = __ldg(pl + (size_t)(j + u) * H); }
#pragma unroll
                for (int u = 0; u < 4; ++u) s += a[u] + b[u];
            }
            for (; j < n; ++j) s += __ldg(ph + (size_t)j * H) + __ldg(pl + (size_t)j * H);
        }
        const float hi = __uint_as_float(__float_as_uint(s) & 0xFFFFE000u);
        Gu_hi[(size_t)r * H + h] = hi;
        Gu_lo[(size_t)r * H + h] = s - hi;
    }
}

}  // namespace snnk
