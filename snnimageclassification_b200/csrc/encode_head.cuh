// encode_head.cuh -- K5 (image -> spike train encoder) and K6 (fused classification head).
#pragma once
#include "common.cuh"

namespace snnk {

// ---- K5: ToSpikes.__call__ (src/datasets/datasets.py:93-97) for a whole batch ---------------------------------
// pixels_to_firing_periods (datasets.py:42-54).  The arithmetic type follows the input dtype, as numpy does.
__device__ __forceinline__ long long period_of(double x, double t_max, double tau, double thr, double eps)
{
    const bool below = x < thr;                                   // :49
    const double lo = thr + eps;
    const double xc = x < lo ? lo : (x > 1.0e9 ? 1.0e9 : x);      // :50
    double Tv = tau * log(xc / (xc - thr));                       // :51
    if (below) Tv = t_max;                                        // :52
    return (long long)Tv;                                         // :54 (truncation)
}

// float32 inputs: numpy keeps float32 (python-float parameters are weak scalars), so every constant is
// rounded to float32 first.  The logarithm is the correctly rounded fp32 log (fp64 log rounded once),
// exactly as oracle/snn_oracle.c defines it.
__device__ __forceinline__ long long period_of(float x, double t_max, double tau, double thr, double eps)
{
    const float thr_f = (float)thr, lo = (float)(thr + eps), hi = (float)1.0e9;
    const bool below = x < thr_f;
    const float xc = x < lo ? lo : (x > hi ? hi : x);
    const float d = __fsub_rn(xc, thr_f);
    const float q = __fdiv_rn(xc, d);
    const float l = (float)log((double)q);
    float Tv = __fmul_rn((float)tau, l);
    if (below) Tv = (float)t_max;
    return (long long)Tv;
}

// One thread per (item, pixel): the latency/period is computed once, then the thread writes its T raster
// entries; consecutive threads own consecutive pixels, so every store instruction is fully coalesced.
//   non-periodic (datasets.py:81-86): one spike at t = period if period < n_steps
//   periodic     (datasets.py:72-79): p = clamp(period, 1, n_steps-1); spike iff t >= p and (t - p) % p == 0
// int64 inputs already hold latencies/periods (firing_times_to_spikes / firing_periods_to_spikes called directly)
__device__ __forceinline__ long long period_of(long long x, double, double, double, double) { return x; }

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_encode(const TIn* __restrict__ x, long long n_items, long long n_pix,
                                               int n_steps, double t_max, double tau, double thr, double eps,
                                               int periodic, TOut* __restrict__ out,
                                               long long* __restrict__ periods, unsigned char* __restrict__ changed)
{
    const long long pix = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    const long long item = blockIdx.x;
    if (pix >= n_pix || item >= n_items) return;
    const long long per = period_of(x[item * n_pix + pix], t_max, tau, thr, eps);
    if (periods) periods[item * n_pix + pix] = per;
    TOut* col = out + item * (long long)n_steps * n_pix + pix;
    // changed[item][t] = 1 when frame t of the item differs from frame t-1 in at least one pixel (zeroed by the
    // caller; every writer stores the same byte) -- the run table of the frame-dedup path is built from it
    unsigned char* chg = changed ? changed + item * (long long)n_steps : nullptr;
    if (!periodic) {
        for (int t = 0; t < n_steps; ++t) {
            const bool s = (long long)t == per;
            col[(long long)t * n_pix] = (TOut)(s ? 1 : 0);
            if (chg && t > 0 && (s || (long long)(t - 1) == per)) chg[t] = 1;
        }
    } else {
        long long p = per > n_steps - 1 ? n_steps - 1 : per;
        p = p < 1 ? 1 : p;
        long long next = p;
        bool prev = false;
        for (int t = 0; t < n_steps; ++t) {
            const bool s = (long long)t == next;
            if (s) next += p;
            col[(long long)t * n_pix] = (TOut)(s ? 1 : 0);
            if (chg && t > 0 && s != prev) chg[t] = 1;
            prev = s;
        }
    }
}

// ---- bit-packed raster: 32 pixels per uint32 word, bit l of word w = pixel 32 w + l -------------------
// The same encoder with the time loop ending in a warp ballot instead of 32 stores: (n_items, n_steps, ceil(n_pix/32))
// words -- 1/32 of the fp32 raster -- for rasters that are stored or moved (host <-> device) rather than consumed at once.
template <typename TIn>
__global__ void __launch_bounds__(256) k_encode_bits(const TIn* __restrict__ x, long long n_items, long long n_pix,
                                                    int n_steps, double t_max, double tau, double thr, double eps,
                                                    int periodic, uint32_t* __restrict__ out)
{
    const long long pix = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    const long long item = blockIdx.x;
    const bool live = pix < n_pix;              // dead lanes keep taking part in the ballots
    const int words = (int)((n_pix + 31) / 32);
    const long long per = live ? period_of(x[item * n_pix + pix], t_max, tau, thr, eps) : -1;
    long long p = per > n_steps - 1 ? n_steps - 1 : per;
    p = p < 1 ? 1 : p;
    long long next = periodic ? p : per;
    uint32_t* row = out + item * (long long)n_steps * words + (pix >> 5);
    const bool writer = (threadIdx.x & 31) == 0 && (pix >> 5) < words;
    for (int t = 0; t < n_steps; ++t) {
        const bool s = live && (long long)t == next;
        if (s && periodic) next += p;
        const unsigned m = __ballot_sync(0xffffffffu, s);
        if (writer) row[(long long)t * words] = m;
    }
}

// bits (n_rows, ceil(n_pix/32)) -> fp32 raster (n_rows, n_pix); one thread per output element, coalesced stores
__global__ void __launch_bounds__(256) k_unpack_raster(const uint32_t* __restrict__ bits, long long n_rows, int n_pix,
                                                      float* __restrict__ out)
{
    const int words = (n_pix + 31) / 32;
    const long long total = n_rows * (long long)n_pix;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / n_pix;
        const int c = (int)(e - r * n_pix);
        out[e] = (float)((__ldg(bits + r * words + (c >> 5)) >> (c & 31)) & 1u);
    }
}

// ---- lazy raster (snnk_encode_runs with lazy != 0) ------------------------------------------------------------------
// Pass A: the change flags alone, straight from the latency/period of every pixel -- no loop over time: a pixel of
// period p >= 2 toggles at every multiple of p and one step later, a pixel of period 1 only at t = 1, a latency-coded
// pixel at its firing time and one step later.  Same flags as k_encode records.
template <typename TIn>
__global__ void __launch_bounds__(256) k_encode_flags(const TIn* __restrict__ x, long long n_items, long long n_pix,
                                                     int n_steps, double t_max, double tau, double thr, double eps,
                                                     int periodic, unsigned char* __restrict__ changed)
{
    const long long pix = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    const long long item = blockIdx.x;
    if (pix >= n_pix || item >= n_items) return;
    const long long per = period_of(x[item * n_pix + pix], t_max, tau, thr, eps);
    unsigned char* chg = changed + item * (long long)n_steps;
    if (!periodic) {
        if (per >= 0 && per < n_steps) {
            if (per > 0) chg[per] = 1;
            if (per + 1 < n_steps) chg[per + 1] = 1;
        }
    } else {
        long long p = per > n_steps - 1 ? n_steps - 1 : per;
        p = p < 1 ? 1 : p;
        if (p == 1) {
            if (n_steps > 1) chg[1] = 1;
        } else {
            for (long long t = p; t < n_steps; t += p) {
                chg[t] = 1;
                if (t + 1 < n_steps) chg[t + 1] = 1;
            }
        }
    }
}

// Pass B, after k_frame_runs: the whole raster if the table is not ok (the dense kernels will read every row),
// otherwise only the first row of every run -- the only rows the frame-dedup kernels ever read.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_encode_rows(const TIn* __restrict__ x, long long n_items, long long n_pix,
                                                    int n_steps, double t_max, double tau, double thr, double eps,
                                                    int periodic, TOut* __restrict__ out, const int* __restrict__ table)
{
    const long long pix = (long long)blockIdx.y * blockDim.x + threadIdx.x;
    const long long item = blockIdx.x;
    if (pix >= n_pix || item >= n_items) return;
    const long long per = period_of(x[item * n_pix + pix], t_max, tau, thr, eps);
    long long p = per > n_steps - 1 ? n_steps - 1 : per;
    p = p < 1 ? 1 : p;
    TOut* col = out + item * (long long)n_steps * n_pix + pix;
    if (table[1] != 1) {
        long long next = periodic ? p : per;
        for (int t = 0; t < n_steps; ++t) {
            const bool s = (long long)t == next;
            if (s && periodic) next += p;
            col[(long long)t * n_pix] = (TOut)(s ? 1 : 0);
        }
        return;
    }
    const int* row2c = table + 4;
    const int* rep = row2c + n_items * n_steps;
    const int r0 = row2c[item * n_steps], r1 = row2c[item * n_steps + n_steps - 1];
    for (int r = r0; r <= r1; ++r) {
        const long long t = rep[r] - item * n_steps;
        const bool s = periodic ? (t >= p && t % p == 0) : t == per;
        col[t * n_pix] = (TOut)(s ? 1 : 0);
    }
}

// ---- SpikeFunction.apply stand-alone (src/modules/spike_funcs.py:12-29, :46-62, :65-79) --------------------------
// thr is a tensor of the same shape as v, or a single element (thr_n == 1) broadcast over it.
__global__ void __launch_bounds__(256) k_spike_fwd(const float* __restrict__ v, const float* __restrict__ thr,
                                                  long long n, long long thr_n, float* __restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = v[e] >= thr[thr_n == 1 ? 0 : e] ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(256) k_spike_bwd(int kind, const float* __restrict__ v,
                                                  const float* __restrict__ thr, const float* __restrict__ gamma,
                                                  const float* __restrict__ g, long long n, long long thr_n,
                                                  float* __restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = __fmul_rn(g[e], surrogate_grad(kind, gamma[0], v[e], thr[thr_n == 1 ? 0 : e]));
}

// ---- K6: log_softmax (snn.py:258) + NLLLoss mean (snn.py:297) + d loss / d logits -----------------------------
// One CTA; B is a few thousand at most.  The loss is reduced in a fixed order (deterministic).
__global__ void __launch_bounds__(256) k_head_nll(int B, int O, const float* __restrict__ logits,
                                                 const long long* __restrict__ labels, float* __restrict__ logp,
                                                 float* __restrict__ loss, float* __restrict__ g_logits,
                                                 unsigned long long* mailbox, unsigned int* counter)
{
    __shared__ double s_part[256];
    __shared__ int s_cnt[9];
    // torch.nn.NLLLoss semantics for the labels (snn.py:297 uses the defaults): rows labelled ignore_index (-100) are
    // left out of the mean and get a zero gradient; any other label outside [0, O) is an error -- torch raises, here
    // the loss and that row's gradient become NaN so the mistake cannot pass silently
    constexpr long long kIgnore = -100;
    int nv = 0, nbad = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const long long l = labels[b];
        nv += (l >= 0 && l < O);
        nbad += (l != kIgnore && (l < 0 || l >= O));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = nv;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += s_cnt[q];
        s_cnt[8] = t;
    }
    const bool bad = __syncthreads_or(nbad != 0) != 0;
    const int n_valid = s_cnt[8];      // == B for well-formed labels
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float lg[kOMax];
        float mx = -INFINITY;
        for (int c = 0; c < O; ++c) { lg[c] = logits[(size_t)b * O + c]; mx = fmaxf(mx, lg[c]); }
        float se = 0.f;
        for (int c = 0; c < O; ++c) se += expf(lg[c] - mx);
        const float lse = logf(se);
        const long long lab = labels[b];
        const bool row_ok = lab >= 0 && lab < O, row_bad = !row_ok && lab != kIgnore;
        for (int c = 0; c < O; ++c) {
            const float lp = (lg[c] - mx) - lse;
            if (logp) logp[(size_t)b * O + c] = lp;
            if (c == lab) acc += -(double)lp;
            if (g_logits) {
                float g = row_ok ? __fdiv_rn(expf(lp) - (c == lab ? 1.0f : 0.0f), (float)n_valid) : 0.0f;
                if (row_bad) g = __int_as_float(0x7fc00000);
                g_logits[(size_t)b * O + c] = g;
            }
        }
    }
    // fixed-shape reduction tree (shuffles inside a warp, then the warp sums in order): deterministic, and not a serial
    // walk over 256 shared-memory words
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += s_part[q];
        const float lossf = bad ? __int_as_float(0x7fc00000) : (float)(s / (double)n_valid);
        *loss = lossf;
        if (mailbox) {
            // {launch number, loss} as ONE 8-byte store into pinned host memory: the host reads the step's loss by
            // polling this word instead of synchronising with the stream, i.e. while the backward pass still runs
            const unsigned int seq = *counter + 1u;
            *counter = seq;
            *reinterpret_cast<volatile unsigned long long*>(mailbox) =
                (static_cast<unsigned long long>(seq) << 32) | __float_as_uint(lossf);
            __threadfence_system();
        }
    }
}

// ---- optimizer step of SNN._exec_batch (snn.py:414; the reference's default is Adam(lr, weight_decay=1e-5), :299) -----
// torch.optim.Adam (no amsgrad, L2 weight decay) for up to kAdamMaxTensors parameter tensors in ONE launch; the
// per-tensor step counters live on the device (float32, as torch keeps them when capturable) so the launch can sit
// inside a captured CUDA graph.  Tensors whose gradient is absent (the never-trained beta) are simply not listed.
constexpr int kAdamMaxTensors = 16;
struct AdamTensors {
    float* p[kAdamMaxTensors]; const float* g[kAdamMaxTensors]; float* m[kAdamMaxTensors]; float* v[kAdamMaxTensors];
    float* step[kAdamMaxTensors]; long long n[kAdamMaxTensors]; long long start[kAdamMaxTensors + 1];
    int count;
};

// CTAs that have finished the current k_adam_step launch; the last one bumps the step counters and resets it.  One
// instance per device (module global); optimizer steps on one device are issued one after the other.
__device__ unsigned int g_adam_done = 0;

__global__ void __launch_bounds__(256) k_adam_step(const AdamTensors t, float lr, float beta1, float beta2, float eps,
                                                  float weight_decay)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < t.start[t.count]) {
        int k = 0;
        while (e >= t.start[k + 1]) ++k;
        const long long i = e - t.start[k];
        const float step = *t.step[k] + 1.0f;       // every thread reads the old counter; the last CTA out bumps it
        float g = t.g[k][i];
        const float p = t.p[k][i];
        g = fmaf(weight_decay, p, g);
        const float m = fmaf(1.0f - beta1, g - t.m[k][i], t.m[k][i]);
        const float v = fmaf(1.0f - beta2, g * g, beta2 * t.v[k][i]);
        t.m[k][i] = m;
        t.v[k][i] = v;
        const float bc1 = 1.0f - powf(beta1, step), bc2 = 1.0f - powf(beta2, step);
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        t.p[k][i] = p - (lr / bc1) * (m / denom);
    }
    // all reads of the step counters by this CTA are done once its threads pass the barrier
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&g_adam_done, 1u) == gridDim.x - 1) {
            g_adam_done = 0;
            for (int k = 0; k < t.count; ++k) *t.step[k] += 1.0f;
        }
    }
}

// ---- data-parallel optimizer step: gradient exchange + mean + Adam in ONE kernel -------------------
// The only exchange step of the path is the mean of the weight gradients over the ranks.  Instead of an NCCL
// all-reduce followed by the optimizer launch, the thread that owns gradient element e PUSHES it over NVLink straight
// into slot `rank` of every peer's exchange buffer (peer-mapped symmetric memory) as one 8-byte word {value, epoch},
// then polls the same element of the other ranks' slots in its OWN buffer until their epoch tags match, sums the
// `world` values in rank order (bit-identical on every rank, so the replicated weights never diverge), stores the
// mean back as the gradient and applies Adam.  A 64-bit store is single-copy atomic, so the tag that arrives with the
// value IS the synchronisation: no flags, no memory fences (a system-scope fence behind remote stores was measured at
// ~5 us, and the flag protocol needs three in a row), no grid-wide dependency -- the latency is one NVLink crossing.
//   exchange buffer per rank: [2 parities][world slots][total] x uint64
//   parity = epoch & 1: a rank can run at most one launch ahead of a peer (it needs the peer's values of launch e+1,
//   which the peer sends after it finished reading launch e), so two buffers suffice; the epoch counter lives in
//   `state` and survives graph replays.  Buffers start zeroed and epochs start at 1.
constexpr int kDpMaxWorld = 16;
struct AdamDp {
    unsigned long long* slots[kDpMaxWorld];   // peer r's exchange buffer as mapped into this process
    unsigned* state;                // local: [0] epoch, [1] unused, [2] leave counter, [3] timeout marker,
                                    // [4..11] globaltimer ns of the last launch as seen by CTA 0: start, pushed, peers seen, done
    int rank, world;
    int rsag;                       // 1: two-phase exchange (reduce-scatter to the element's owner, all-gather of the mean)
};

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) k_adam_step_dp(const AdamTensors t, const AdamDp dp, float lr, float beta1,
                                                     float beta2, float eps, float weight_decay)
{
    const long long total = t.start[t.count];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned epoch = *(volatile unsigned*)(dp.state) + 1u;   // bumped by the last CTA to leave
    const size_t par = (size_t)(epoch & 1u) * dp.world;
    unsigned long long* stamp = reinterpret_cast<unsigned long long*>(dp.state + 4);
    const bool scribe = blockIdx.x == 0 && threadIdx.x == 0;
    if (scribe) stamp[0] = global_ns();
    const unsigned long long tag = (unsigned long long)epoch << 32;
    const size_t mine = (par + dp.rank) * (size_t)total;
    const unsigned long long* local = dp.slots[dp.rank];
    const float inv = 1.0f / (float)dp.world;

    int k = 0;
    for (long long e = e0; e < total; e += stride) {
        while (e >= t.start[k + 1]) ++k;
        const long long i = e - t.start[k];
        const float g_own = t.g[k][i];
        const unsigned long long word = tag | (unsigned long long)__float_as_uint(g_own);
        unsigned long long t0 = 0;
        // waits until the word at src carries this epoch's tag; 20 s: a peer died -- fail loudly instead of hanging
        auto await = [&](const unsigned long long* src) -> float {
            unsigned long long w = ld_relaxed_sys(src);
            while ((unsigned)(w >> 32) != epoch) {
                if (t0 == 0) t0 = global_ns();
                else if (global_ns() - t0 > 20000000000ull) {
                    dp.state[3] = epoch;
                    __threadfence_system();
                    __trap();
                }
                w = ld_relaxed_sys(src);
            }
            return __uint_as_float((unsigned)w);
        };
        float g = 0.0f;
        if (!dp.rsag) {
            // 1. push {value, epoch} to every peer
            for (int r = 0; r < dp.world; ++r)
                if (r != dp.rank) st_relaxed_sys(dp.slots[r] + mine + e, word);
            if (scribe && e == e0) stamp[1] = global_ns();
            // 2. gather the peers' values of this element from the local buffer, in rank order
            for (int r = 0; r < dp.world; ++r)
                g += r == dp.rank ? g_own : await(local + (par + r) * (size_t)total + e);
            if (scribe && e == e0) stamp[2] = global_ns();
            // 3. mean
            g *= inv;
        } else {
            // Two phases over the same tagged words: every element has an OWNER rank (contiguous slices).  A rank pushes
            // its value to the owner only; the owner sums the world values in rank order, and pushes the MEAN into the
            // slot its own raw value would have had in every peer's buffer (free there: nobody but the owner writes to
            // slot (source = owner, e) of a non-owner).  1/world of the all-to-all's bytes leave every GPU twice
            // (6.6 MB -> 1.7 MB at 8 GPUs) for one more one-way trip; all ranks use the owner's bits.
            const long long slice = (total + dp.world - 1) / dp.world;
            const int owner = (int)(e / slice);
            if (owner != dp.rank) {
                st_relaxed_sys(dp.slots[owner] + mine + e, word);
                if (scribe && e == e0) stamp[1] = global_ns();
                g = await(local + (par + owner) * (size_t)total + e);
            } else {
                if (scribe && e == e0) stamp[1] = global_ns();
                for (int r = 0; r < dp.world; ++r)
                    g += r == dp.rank ? g_own : await(local + (par + r) * (size_t)total + e);
                g *= inv;
                const unsigned long long mword = tag | (unsigned long long)__float_as_uint(g);
                for (int r = 0; r < dp.world; ++r)
                    if (r != dp.rank) st_relaxed_sys(dp.slots[r] + mine + e, mword);
            }
            if (scribe && e == e0) stamp[2] = global_ns();
        }
        // Adam on the mean
        const_cast<float*>(t.g[k])[i] = g;      // the caller sees the averaged gradient, as after an all-reduce
        const float step = *t.step[k] + 1.0f;
        const float p = t.p[k][i];
        g = fmaf(weight_decay, p, g);
        const float m = fmaf(1.0f - beta1, g - t.m[k][i], t.m[k][i]);
        const float v = fmaf(1.0f - beta2, g * g, beta2 * t.v[k][i]);
        t.m[k][i] = m;
        t.v[k][i] = v;
        const float bc1 = 1.0f - powf(beta1, step), bc2 = 1.0f - powf(beta2, step);
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        t.p[k][i] = p - (lr / bc1) * (m / denom);
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(dp.state + 2, 1u) == gridDim.x - 1) {
        dp.state[2] = 0;
        dp.state[0] = epoch;
        for (int kk = 0; kk < t.count; ++kk) *t.step[kk] += 1.0f;   // every thread of the grid has read the old counters
    }
    if (scribe) stamp[3] = global_ns();
}

}  // namespace snnk
