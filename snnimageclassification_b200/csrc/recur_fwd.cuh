// recur_fwd.cuh -- K2: persistent fused forward recurrence.
//
// Replaces the time loop of SNN.forward (src/modules/snn.py:209-214) around LIFLayer/ALIFLayer.forward
// (src/modules/spiking_layers.py:156-171, :229-243) and ReadoutLayer.forward (:402-408).
//
// IzhikevichLayer.forward (:330-353) is the MODE = 2 variant of the same kernel.
//
// One CTA owns R batch rows for all T steps (rows are independent, so there is no grid-wide sync).
// Thread i owns hidden neuron i: its membrane/adaptation state lives in registers for the whole
// sequence.  The masked recurrent matrix is register-resident too, column-blocked over lane quads
// (common.cuh, dot_rec16_cb): a step is H/2 packed FFMA2 against a quarter of the previous spike vector
// read from shared memory, three quad shuffles, the elementwise update and one __syncthreads.  The input
// current of the row (a contiguous T x H block written by the projection GEMM, or the compact rows of the
// frame-dedup variant) is streamed through a shared-memory ring by 1-D bulk async copies (cp.async.bulk +
// mbarrier), kChunk steps per copy and kRing copies in flight, so HBM/L2 latency never sits on the
// step-to-step critical path.  Spikes are bit-packed with warp ballots; all T spike words stay in
// shared memory, so the leaky readout (linear in the spikes) is evaluated after the loop instead of
// inside the latency-critical chain.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace snnk {

constexpr int kChunk = 8;   // time steps per bulk copy
constexpr int kRing = 4;    // ring slots (copies in flight)

template <int H, int R>
constexpr size_t fwd_smem_bytes(int T, int O)
{
    return sizeof(float) * (size_t)(2 * R * H) + sizeof(uint32_t) * (size_t)((R * T * (H / 32) + 3) & ~3) +
           sizeof(float) * (size_t)((H * O + R * T * O + 3) & ~3) + sizeof(float) * (size_t)(kRing * R * kChunk * H) +
           sizeof(uint64_t) * (kRing + 2) + sizeof(float) * (size_t)H * H +   // + staging of the recurrent matrix
           sizeof(int) * (size_t)((R * T + 3) & ~3);                          // + compact row of every step (runs.cuh)
}

// Epilogue shared by the forward kernels: bit-packed raster out, leaky readout (spiking_layers.py:407) and max over
// time (snn.py:228) from the T spike words kept in shared memory.  tid/nthr: thread index and block size.
template <int H, int R>
__device__ __forceinline__ void fwd_tail(const FwdParams& p, const uint32_t* s_mask, const float* s_wout, float* s_s, int b0,
                                         int tid, int nthr, float* s_logit = nullptr)
{
    constexpr int W32 = H / 32;
    const int T = p.T, O = p.O, B = p.B;
    // bit-packed raster for the backward pass
    for (int idx = tid; idx < R * T * W32; idx += nthr) {
        const int r = idx / (T * W32), rem = idx - r * (T * W32);
        if (b0 + r < B) p.zbits[(size_t)(b0 + r) * T * W32 + rem] = s_mask[idx];
    }

    // Leaky readout, spiking_layers.py:407:  y_t = kappa y_{t-1} + Z_t @ W_out + b.
    // (A) s[t][c] = sum_j Z_t[j] W_out[j][c], ascending j.  nthr / O threads per class, each holding its class's
    // column of W_out in registers (the recurrent weights are dead by now) and walking its share of the T spike
    // words with one predicated add per bit: adding w only where the bit is set equals adding (bit ? w : 0) for every
    // bit, because the running sum is never -0.
    {
        const int tpc = nthr / O, c = tid / tpc, u = tid - c * tpc;
        if (c < O) {
            float wc[H];
#pragma unroll
            for (int j = 0; j < H; ++j) wc[j] = s_wout[j * O + c];
            for (int rt = u; rt < R * T; rt += tpc) {
                const uint32_t* mw = s_mask + rt * W32;
                float sum = 0.f;
#pragma unroll
                for (int wd = 0; wd < W32; ++wd) {
                    const uint32_t m = mw[wd];
#pragma unroll
                    for (int l = 0; l < 32; ++l)
                        if (m & (1u << l)) sum = __fadd_rn(sum, wc[wd * 32 + l]);
                }
                s_s[rt * O + c] = sum;
            }
        }
    }
    __syncthreads();
    // (B) the scan over t (sequential per (row, class)) + max over time, snn.py:228 (first max wins)
    for (int idx = tid; idx < R * O; idx += nthr) {
        const int r = idx / O, c = idx - r * O;
        const float bc = __ldg(p.b_out + c);
        float yv = 0.f, mx = 0.f;
        int mt = 0;
        for (int t = 0; t < T; ++t) {
            float* sp = s_s + (r * T + t) * O + c;
            yv = __fadd_rn(__fadd_rn(__fmul_rn(p.kappa, yv), *sp), bc);
            *sp = yv;
            if (t == 0 || yv > mx) { mx = yv; mt = t; }
        }
        if (b0 + r < B) {
            p.logits[(size_t)(b0 + r) * O + c] = mx;
            p.tstar[(size_t)(b0 + r) * O + c] = mt;
        }
        if (s_logit) s_logit[r * kOMax + c] = mx;      // for the fused head
    }
    __syncthreads();
    // (C) coalesced write of the output trace
    for (int idx = tid; idx < R * T * O; idx += nthr) {
        const int r = idx / (T * O), rem = idx - r * (T * O);
        if (b0 + r < B) p.y[(size_t)(b0 + r) * T * O + rem] = s_s[idx];
    }
}

// ---- fused head (snn.py:228 -> :258 -> :297) ---------------------------------------------------------------------------
// The arithmetic and the summation order of k_head_nll (encode_head.cuh), moved into the tail of the forward kernel:
// a separate one-CTA launch between the forward pass and the BPTT sweep cost ~10 us of an 180 us step.
// head_count: batch-wide label statistics (every CTA counts them itself, in its prologue).
// head_tail : thread r < R evaluates row b0 + r exactly like k_head_nll's per-row body; the last CTA to take a ticket
//             sums the per-row NLL terms as k_head_nll's 256 threads would (virtual thread v: rows v, v+256, ...; xor
//             shuffle tree per virtual warp; the eight warp sums in order) -- the loss is bit-identical to the
//             stand-alone kernel's and run-to-run deterministic.
__device__ __forceinline__ void head_count(const FwdParams& p, int* s_hd, int tid, int nthr)
{
    constexpr long long kIgnore = -100;
    int nv = 0, nbad = 0;
    for (int b = tid; b < p.B; b += nthr) {
        const long long l = p.labels[b];
        nv += (l >= 0 && l < p.O);
        nbad += (l != kIgnore && (l < 0 || l >= p.O));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        nbad += __shfl_xor_sync(0xffffffffu, nbad, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_hd[0], nv); atomicAdd(&s_hd[1], nbad); }
}

template <int R>
__device__ __forceinline__ void head_tail(const FwdParams& p, const float* s_logit, int* s_hd, double* s_hp, int b0, int tid, int nthr)
{
    constexpr long long kIgnore = -100;
    const int O = p.O, B = p.B;
    const int n_valid = s_hd[0];
    const bool bad = s_hd[1] != 0;
    if (tid < R && b0 + tid < B) {
        const int b = b0 + tid;
        float lg[kOMax];
        float mx = -INFINITY;
        for (int c = 0; c < O; ++c) { lg[c] = s_logit[tid * kOMax + c]; mx = fmaxf(mx, lg[c]); }
        float se = 0.f;
        for (int c = 0; c < O; ++c) se += expf(lg[c] - mx);
        const float lse = logf(se);
        const long long lab = p.labels[b];
        const bool row_ok = lab >= 0 && lab < O, row_bad = !row_ok && lab != kIgnore;
        float nll = 0.f;
        for (int c = 0; c < O; ++c) {
            const float lp = (lg[c] - mx) - lse;
            if (p.logp) p.logp[(size_t)b * O + c] = lp;
            if (c == lab) nll = -lp;
            if (p.g_logits) {
                float g = row_ok ? __fdiv_rn(expf(lp) - (c == lab ? 1.0f : 0.0f), (float)n_valid) : 0.0f;
                if (row_bad) g = __int_as_float(0x7fc00000);
                p.g_logits[(size_t)b * O + c] = g;
            }
        }
        p.part_nll[b] = nll;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) s_hd[2] = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_hd[2]) return;
    __threadfence();
    const int lane = tid & 31;
    for (int vw = tid >> 5; vw < 8; vw += nthr >> 5) {
        double acc = 0.0;
        for (int b = vw * 32 + lane; b < B; b += 256) acc += (double)__ldcg(p.part_nll + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_hp[vw] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int q = 0; q < 8; ++q) sum += s_hp[q];
        const float lossf = bad ? __int_as_float(0x7fc00000) : (float)(sum / (double)n_valid);
        *p.loss = lossf;
        *p.ticket = 0u;
        if (p.mailbox) {      // {launch number, loss} as one 8-byte store into pinned host memory (see k_head_nll)
            const unsigned int seq = *p.mail_counter + 1u;
            *p.mail_counter = seq;
            *reinterpret_cast<volatile unsigned long long*>(p.mailbox) =
                (static_cast<unsigned long long>(seq) << 32) | __float_as_uint(lossf);
            __threadfence_system();
        }
    }
}

// MODE: 0 LIF, 1 ALIF, 2 Izhikevich -- compile-time, because a run-time branch on the layer type inside the step
// loop measurably slows the issue-bound kernel (6 % for one extra uniform branch)
template <int H, int R, bool REC, int MODE = 1>
__global__ void __launch_bounds__(H, 256 / H) k_recur_fwd(const FwdParams p)
{
    constexpr bool IZH = MODE == 2, ALIF = MODE == 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W32 = H / 32;
    const int T = p.T, O = p.O, B = p.B;
    const int i = threadIdx.x, lane = i & 31, warp = i >> 5;
    const int b0 = blockIdx.x * R;
    const int nvalid = min(R, B - b0);

    float* s_z = reinterpret_cast<float*>(smem_raw);                                // [2][R][H]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_z + 2 * R * H);                // [R][T][W32]
    float* s_wout = reinterpret_cast<float*>(s_mask + ((R * T * W32 + 3) & ~3));    // [H][O]
    float* s_s = s_wout + H * O;                                                    // [R][T][O]
    float* s_in = s_wout + ((H * O + R * T * O + 3) & ~3);                          // [kRing][R][kChunk][H], 16-B aligned
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_in + kRing * R * kChunk * H);   // [kRing] + 1 for the weights
    float* s_w = reinterpret_cast<float*>(s_bar + kRing + 2);                       // [H][H] staging (16-B aligned), prologue only
    int* s_r2c = reinterpret_cast<int*>(s_w + H * H);                               // [R][T] compact row of (row, t)

    // frame-dedup variant: the projection was evaluated once per run of equal frames; every step fetches its run's row
    const bool compact = p.run_table != nullptr && p.run_table[1] == 1;
    const int nchunks = (T + kChunk - 1) / kChunk;
    // bulk copy of chunk c (kChunk consecutive steps of every valid row) into ring slot c % kRing; thread 0 only
    auto issue_chunk = [&](int c) {
        const int slot = c % kRing, t0 = c * kChunk;
        const int nt = min(kChunk, T - t0);
        const uint32_t bytes = (uint32_t)(nt * H * sizeof(float));
        tc::mbar_expect_tx(s_bar + slot, bytes * nvalid);
        if (!compact) {
            for (int r = 0; r < nvalid; ++r)
                tc::bulk_g2s(s_in + ((slot * R + r) * kChunk) * H, p.I_in + ((size_t)(b0 + r) * T + t0) * H, bytes,
                             s_bar + slot);
        } else {
            for (int r = 0; r < nvalid; ++r)
                for (int tt = 0; tt < nt; ++tt)
                    tc::bulk_g2s(s_in + ((slot * R + r) * kChunk + tt) * H, p.I_u + (size_t)s_r2c[r * T + t0 + tt] * H,
                                 (uint32_t)(H * sizeof(float)), s_bar + slot);
        }
    };
    if (compact)
        for (int idx = i; idx < nvalid * T; idx += H)
            s_r2c[idx] = __ldg(p.run_table + kRunHdrInts + (size_t)b0 * T + idx);   // rows b0.. are consecutive
    __shared__ int s_hd[4];                 // fused head: valid labels, bad labels, "this is the last CTA"
    __shared__ double s_hp[8];
    __shared__ float s_logit[R * kOMax];
    if (i == 0) {
        s_hd[0] = s_hd[1] = s_hd[2] = 0;
        for (int s = 0; s <= kRing; ++s) tc::mbar_init(s_bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (REC) {   // the whole masked recurrent matrix in one bulk copy (64 KB at H = 128)
            tc::mbar_expect_tx(s_bar + kRing, (uint32_t)(H * H * sizeof(float)));
            tc::bulk_g2s(s_w, p.W_eff, (uint32_t)(H * H * sizeof(float)), s_bar + kRing);
        }
    }
    __syncthreads();   // barrier inits (and the compact-row table) visible before anyone issues or waits
    if (p.labels) head_count(p, s_hd, i, H);      // its global loads overlap the weight staging below
    // Frame-dedup variant: the compact rows of one sample are CONSECUTIVE rows of I_u (three for the production encoder),
    // so they are fetched ONCE, with one bulk copy per sample, into the memory of the ring and indexed per step.  A copy
    // per (row, step) -- what the ring does for this variant -- costs the SM's TMA unit ~270 cycles per 512-byte
    // cp.async.bulk (measured, profiles/r02_*): two CTAs per SM made that 540 of the ~1000 cycles of a step.
    constexpr int kCacheRows = kRing * kChunk;      // compact rows per sample that fit the ring's memory
    int first[R];
    bool cached = compact;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        first[r] = (compact && r < nvalid) ? s_r2c[r * T] : 0;
        if (compact && r < nvalid && s_r2c[r * T + T - 1] - first[r] + 1 > kCacheRows) cached = false;
    }
    if (i == 0) {
        if (cached) {
            uint32_t total = 0;
            for (int r = 0; r < nvalid; ++r) total += (uint32_t)(s_r2c[r * T + T - 1] - first[r] + 1) * H * sizeof(float);
            tc::mbar_expect_tx(s_bar, total);
            for (int r = 0; r < nvalid; ++r)
                tc::bulk_g2s(s_in + r * kCacheRows * H, p.I_u + (size_t)first[r] * H,
                             (uint32_t)(s_r2c[r * T + T - 1] - first[r] + 1) * H * sizeof(float), s_bar);
        } else {
            for (int c = 0; c < kRing && c < nchunks; ++c) issue_chunk(c);
        }
    }

    // column i of W_rec (.) rec_mask -> registers for the whole sequence
    float w[REC ? H : 16];
    if constexpr (REC) {
        tc::mbar_wait(s_bar + kRing, 0);
        load_w_cb<REC ? H : 16>(w, s_w, i);
    }
    const float beta = (ALIF && p.beta) ? __ldg(p.beta) : 0.f;

    float v[R], a[R], zp[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int b = b0 + r;
        valid[r] = b < B;
        const size_t s = (size_t)(valid[r] ? b : 0) * H + i;
        v[r] = (valid[r] && p.V0) ? p.V0[s] : (IZH ? p.iz.vr : 0.f);   // Izhikevich starts at v_rest (:309)
        a[r] = (valid[r] && p.a0) ? p.a0[s] : 0.f;
        zp[r] = (valid[r] && p.Z0) ? p.Z0[s] : 0.f;
        if (REC) s_z[1 * R * H + r * H + i] = zp[r];   // step 0 reads buffer (0+1)&1
    }
    for (int idx = i; idx < H * O; idx += H) s_wout[idx] = __ldg(p.W_out + idx);
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const int c = t / kChunk, tt = t - c * kChunk, slot = c % kRing;
        float cur[R];
        if (cached) {
            if (t == 0) tc::mbar_wait(s_bar, 0);
#pragma unroll
            for (int r = 0; r < R; ++r) cur[r] = valid[r] ? s_in[(r * kCacheRows + s_r2c[r * T + t] - first[r]) * H + i] : 0.f;
        } else {
            if (tt == 0) {
                // every thread is past its last read of chunk c-1 (REC: the step barrier; otherwise sync here), so
                // its slot can be refilled with chunk c-1+kRing; then wait for chunk c to have landed
                if (!REC) __syncthreads();
                if (i == 0 && c >= 1 && c - 1 + kRing < nchunks) issue_chunk(c - 1 + kRing);
                tc::mbar_wait(s_bar + slot, (c / kRing) & 1);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) cur[r] = valid[r] ? s_in[((slot * R + r) * kChunk + tt) * H + i] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float rec = 0.0f;
            if constexpr (REC) {
                const float4* zv = reinterpret_cast<const float4*>(s_z + ((t + 1) & 1) * R * H + r * H);
                rec = dot_rec16_cb<REC ? H : 16>(w, zv, i & 3);
            }
            // V' = (alpha V + I_in + I_rec) (1 - Z.detach())     spiking_layers.py:169/239
            float vn, thr = p.theta;
            if constexpr (IZH) {
                // IzhikevichLayer.forward, spiking_layers.py:344-349 (same operation order as oracle/snn_oracle.c)
                const float I = __fadd_rn(cur[r], rec);
                const float d1 = __fsub_rn(v[r], p.iz.vr), d2 = __fsub_rn(v[r], p.iz.vth);
                const float q = __fsub_rn(__fmul_rn(__fmul_rn(p.iz.k, d1), d2), a[r]);
                const float inc = __fdiv_rn(__fmul_rn(p.iz.dt, __fadd_rn(q, I)), p.iz.C);
                vn = __fadd_rn(__fmul_rn(__fadd_rn(v[r], inc), __fsub_rn(1.0f, zp[r])), __fmul_rn(p.iz.c, zp[r]));
                const float du = __fmul_rn(p.iz.a, __fsub_rn(__fmul_rn(p.iz.b, d1), a[r]));
                a[r] = __fadd_rn(__fadd_rn(a[r], __fmul_rn(p.iz.dt, du)), __fmul_rn(p.iz.d, zp[r]));
                thr = p.iz.vpeak;
            } else {
                const float t1 = __fmul_rn(p.alpha, v[r]);
                const float t2 = __fadd_rn(t1, cur[r]);
                const float t3 = __fadd_rn(t2, rec);
                vn = __fmul_rn(t3, __fsub_rn(1.0f, zp[r]));
                if constexpr (ALIF) {
                    a[r] = __fadd_rn(__fmul_rn(p.rho, a[r]), zp[r]);          // :240
                    thr = __fadd_rn(p.theta, __fmul_rn(beta, a[r]));          // :241
                }
            }
            const float zn = vn >= thr ? 1.0f : 0.0f;                     // spike_funcs.py:27-28
            if (p.traces && valid[r]) {
                const size_t o = ((size_t)(b0 + r) * T + t) * H + i;
                p.V[o] = vn;
                p.Z[o] = zn;
                if (IZH || ALIF) p.a[o] = a[r];
            }
            const unsigned m = __ballot_sync(0xffffffffu, zn != 0.f);
            if (lane == 0) s_mask[(r * T + t) * W32 + warp] = m;
            if (REC) s_z[(t & 1) * R * H + r * H + i] = zn;
            v[r] = vn;
            zp[r] = zn;
        }
        if (REC) __syncthreads();
    }
    __syncthreads();

    fwd_tail<H, R>(p, s_mask, s_wout, s_s, b0, i, H, s_logit);
    if (p.labels) head_tail<R>(p, s_logit, s_hd, s_hp, b0, i, H);
}

// A k-split variant (two, then eight lanes per neuron; 8 warps per row) was built and measured twice: 76 us and 74 us
// against 65 us / 59 us for the kernel above at B = 256.  The step is bound by instruction issue (each SM sub-partition
// runs its two warps at the 1-per-2-cycles rate of the FMA pipe), so spreading a row over more threads only adds the
// duplicated tail and wider barriers; see DESIGN.md.

}  // namespace snnk
