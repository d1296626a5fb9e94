// ubench.cu -- B200 micro-measurements that decide the shape of the recurrence kernels:
//   mma.sync m16n8k16 bf16 latency / throughput per SM sub-partition, FFMA vs FFMA2 rate, bar.sync and shuffle
//   latency, ldmatrix / stmatrix latency, and the latency of a flag barrier between co-resident CTAs through L2.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// NCH independent accumulator chains per warp, ITERS rounds: cycles per mma per warp and per SMSP
template <int NCH>
__global__ void k_hmma(int iters, long long* out, float* sink)
{
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, threadIdx.x};
    float acc[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) mma16816(acc[c], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += acc[c][0] + acc[c][3];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int PACKED>
__global__ void k_ffma(int iters, long long* out, float* sink)
{
    float2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = make_float2(threadIdx.x * 1e-3f, c);
    const float2 w = make_float2(1.0001f, 0.9999f), z = make_float2(1e-6f, 2e-6f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (PACKED) acc[c] = __ffma2_rn(acc[c], w, z);
            else { acc[c].x = fmaf(acc[c].x, w.x, z.x); acc[c].y = fmaf(acc[c].y, w.y, z.y); }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c].x + acc[c].y;
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

__global__ void k_bar(int iters, long long* out)
{
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(int iters, long long* out, float* sink)
{
    float v = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0f;
    const long long t1 = clock64();
    if (v == 123.456f) sink[0] = v;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// ldmatrix -> mma -> stmatrix -> bar.sync round trip (the dependent chain of one recurrence step, minimal form)
__global__ void k_step_chain(int iters, long long* out, float* sink)
{
    __shared__ __align__(16) uint16_t tile[2][16][136];
    for (int i = threadIdx.x; i < 2 * 16 * 136; i += blockDim.x) (&tile[0][0][0])[i] = 0x3f80;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[4];
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&tile[it & 1][lane & 7][8 * (lane >> 3)]);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
        uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
        mma16816(acc, a, b0);
        mma16816(acc, a, b1);
        const uint32_t pk = acc[0] > 1e30f ? 0x3f803f80u : 0x00003f80u;
        const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&tile[(it + 1) & 1][lane & 7][0]);
        asm volatile("stmatrix.sync.aligned.m8n8.x1.trans.shared.b16 [%0], {%1};" ::"r"(saddr), "r"(pk) : "memory");
        __syncthreads();
    }
    const long long t1 = clock64();
    if (acc[0] == 123.456f) sink[0] = acc[0];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// flag barrier between co-resident CTAs: groups of `gsize` consecutive CTAs, one counter per group
__global__ void k_grid_barrier(int iters, int gsize, unsigned int* counters, long long* out)
{
    const int grp = blockIdx.x / gsize;
    unsigned int* ctr = counters + grp * 32;   // 128 B apart
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            const unsigned int want = (unsigned int)gsize * (unsigned int)(it + 1);
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < want);
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

static double avg(const std::vector<long long>& v) { double s = 0; for (auto x : v) s += x; return s / v.size(); }

int main()
{
    long long* d_out; float* d_sink; unsigned int* d_ctr;
    CK(cudaMalloc(&d_out, 4096 * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 16));
    CK(cudaMalloc(&d_ctr, 32 * 32 * sizeof(unsigned int)));
    std::vector<long long> h(4096);
    const int iters = 2000;
    auto fetch = [&](int n) { cudaMemcpy(h.data(), d_out, n * sizeof(long long), cudaMemcpyDeviceToHost); std::vector<long long> r(h.begin(), h.begin() + n); return avg(r); };
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);

    // mma.sync: 1 warp, 1 chain = latency; more chains / warps = throughput
    k_hmma<1><<<1, 32>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    printf("hmma m16n8k16 bf16: dependent latency %.1f cycles\n", fetch(1) / iters);
    for (int warps : {1, 2, 4, 8, 12, 16}) {
        k_hmma<6><<<148, warps * 32>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
        const double cyc = fetch(148) / iters;
        printf("hmma 6 chains, %2d warps/SM: %.1f cycles per 6 mma per warp -> %.2f cycles per mma per SMSP\n", warps, cyc,
               cyc / 6.0 / ((warps + 3) / 4));
    }
    for (int warps : {4, 8, 16}) {
        k_ffma<0><<<148, warps * 32>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
        const double c0 = fetch(148) / iters;
        k_ffma<1><<<148, warps * 32>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
        const double c1 = fetch(148) / iters;
        printf("%2d warps/SM: 16 FFMA %.1f cycles, 8 FFMA2 %.1f cycles per warp-iteration\n", warps, c0, c1);
    }
    for (int thr : {128, 256, 288}) {
        k_bar<<<148, thr>>>(iters, d_out); CK(cudaDeviceSynchronize());
        printf("bar.sync, %d threads: %.1f cycles\n", thr, fetch(148) / iters);
    }
    k_shfl<<<148, 128>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    printf("shfl + fadd dependent: %.1f cycles\n", fetch(148) / iters);
    for (int thr : {32, 128, 256}) {
        k_step_chain<<<148, thr>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
        printf("ldmatrix -> 2 mma -> stmatrix -> bar.sync, %d threads: %.1f cycles per round\n", thr, fetch(148) / iters);
    }
    for (int gsize : {8, 32, 128}) {
        CK(cudaMemset(d_ctr, 0, 32 * 32 * sizeof(unsigned int)));
        int it2 = 500;
        unsigned int* ctr = d_ctr; long long* out = d_out;
        void* args[] = {&it2, &gsize, &ctr, &out};
        CK(cudaLaunchCooperativeKernel((void*)k_grid_barrier, dim3(128), dim3(128), args, 0, nullptr));
        CK(cudaDeviceSynchronize());
        printf("flag barrier over groups of %3d CTAs (128 CTAs resident): %.0f cycles per barrier\n", gsize, fetch(128) / it2);
    }
    return 0;
}
