// umma_probe.cu -- bring-up probe for tcgen05.mma operand layouts (not part of the product).
// One CTA, one k-block: TMA-loads A[K=32][M=128] and B[K=32][N=128] (MN-major, 4 boxes of 32x32 each) or
// K-major tiles, dumps the smem images, runs 4 tf32 MMAs with descriptors built from RUNTIME parameters and
// compares the accumulator with the host result for a sweep of candidate encodings.
//
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma_probe tools/umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../snnimageclassification_b200/csrc/gemm_tc.cuh"

using namespace snnk::tc;

struct ProbeParams {
    int mn_major;               // 1: tiles are [k][mn] (MN contiguous), 4 boxes; 0: tiles are [mn][k], one box
    uint32_t lbo, sbo, kstep;   // descriptor fields in bytes; kstep = start-address advance per UMMA_K
    uint32_t layout;            // UMMA layout type
    uint32_t idesc;
    float* C;                   // [128][128]
    float* dumpA;               // 4096 floats
};

__global__ void __launch_bounds__(128, 1) k_probe(const __grid_constant__ CUtensorMap mapA,
                                                  const __grid_constant__ CUtensorMap mapB, const ProbeParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = smem;
    unsigned char* sB = smem + 16384;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint64_t* bar2 = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 32768);
        if (p.mn_major) {
            for (int j = 0; j < 4; ++j) {
                tma_load_2d(sA + j * 4096, &mapA, bar, 32 * j, 0);
                tma_load_2d(sB + j * 4096, &mapB, bar, 32 * j, 0);
            }
        } else {
            tma_load_2d(sA, &mapA, bar, 0, 0);
            tma_load_2d(sB, &mapB, bar, 0, 0);
        }
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 4096; i += 128) p.dumpA[i] = reinterpret_cast<float*>(sA)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_smem_desc(smem_u32(sA) + kk * p.kstep, p.lbo, p.sbo, p.layout);
            const uint64_t bd = make_smem_desc(smem_u32(sB) + kk * p.kstep, p.lbo, p.sbo, p.layout);
            umma_tf32(tmem_base, ad, bd, p.idesc, kk != 0);
        }
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0);
    tc_fence_after();
    const int row = 32 * warp + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + c0, v);
        for (int j = 0; j < 32; ++j) p.C[row * 128 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn enc;

static CUtensorMap map2d(const float* base, uint64_t inner, uint64_t outer, uint32_t bi, uint32_t bo,
                         CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t str[1] = {inner * 4};
    cuuint32_t box[2] = {bi, bo};
    cuuint32_t ones[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, str, box, ones,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

int main()
{
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    enc = (EncodeTiledFn)fp;
    const int K = 32, M = 128, N = 128;
    std::vector<float> A(K * M), Bm(K * N), At(M * K), Bt(N * K), Cref(M * N, 0.f);
    srand(1);
    for (int k = 0; k < K; ++k)
        for (int m = 0; m < M; ++m) { A[k * M + m] = (float)(rand() % 3 == 0); At[m * K + k] = A[k * M + m]; }
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            float v = (float)((rand() % 2001) - 1000) / 1024.0f;      // exactly representable in tf32
            Bm[k * N + n] = v; Bt[n * K + k] = v;
        }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += A[k * M + m] * Bm[k * N + n];
            Cref[m * N + n] = s;
        }
    float *dA, *dB, *dAt, *dBt, *dC, *dDump;
    cudaMalloc(&dA, K * M * 4); cudaMalloc(&dB, K * N * 4); cudaMalloc(&dAt, K * M * 4); cudaMalloc(&dBt, K * N * 4);
    cudaMalloc(&dC, M * N * 4); cudaMalloc(&dDump, 16384);
    cudaMemcpy(dA, A.data(), K * M * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bm.data(), K * N * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dAt, At.data(), K * M * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBt, Bt.data(), K * N * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);

    struct Cand { const char* name; int mn; uint32_t lbo, sbo, kstep; uint32_t layout = 2; int sw = 3; };
    const Cand cands[] = {
        {"K-major   lbo=16   sbo=1024 kstep=32   (known good from k_proj_tc)", 0, 16, 1024, 32},
        {"MN-major  lbo=4096 sbo=1024 kstep=1024", 1, 4096, 1024, 1024},
        {"MN-major  lbo=1024 sbo=4096 kstep=1024", 1, 1024, 4096, 1024},
        {"MN-major  lbo=4096 sbo=4096 kstep=1024", 1, 4096, 4096, 1024},
        {"MN-major  lbo=1024 sbo=1024 kstep=1024", 1, 1024, 1024, 1024},
        {"MN-major  lbo=16   sbo=4096 kstep=1024", 1, 16, 4096, 1024},
        {"MN-major  lbo=4096 sbo=16   kstep=1024", 1, 4096, 16, 1024},
        {"MN-major  lbo=4096 sbo=128  kstep=1024", 1, 4096, 128, 1024},
        {"MN-major  lbo=128  sbo=4096 kstep=1024", 1, 128, 4096, 1024},
        {"MN-major BASE32B atom32 lbo=4096 sbo=512  kstep=1024", 1, 4096, 512, 1024, 1, 4},
        {"MN-major BASE32B atom32 lbo=4096 sbo=1024 kstep=1024", 1, 4096, 1024, 1024, 1, 4},
        {"MN-major BASE32B atom32 lbo=512  sbo=4096 kstep=1024", 1, 512, 4096, 1024, 1, 4},
        {"MN-major BASE32B atom32 lbo=1024 sbo=4096 kstep=1024", 1, 1024, 4096, 1024, 1, 4},
        {"MN-major BASE32B sw128  lbo=4096 sbo=512  kstep=1024", 1, 4096, 512, 1024, 1, 3},
        {"MN-major SW128   atom32 lbo=4096 sbo=1024 kstep=1024", 1, 4096, 1024, 1024, 2, 4},
    };
    for (const Cand& c : cands) {
        const CUtensorMapSwizzle sw = (CUtensorMapSwizzle)c.sw;
        CUtensorMap ma = c.mn ? map2d(dA, M, K, 32, 32, sw) : map2d(dAt, K, M, 32, 128, sw);
        CUtensorMap mb = c.mn ? map2d(dB, N, K, 32, 32, sw) : map2d(dBt, K, N, 32, 128, sw);
        ProbeParams p{};
        p.mn_major = c.mn; p.lbo = c.lbo; p.sbo = c.sbo; p.kstep = c.kstep; p.layout = c.layout;
        p.idesc = make_idesc_tf32(N, c.mn, c.mn);
        p.C = dC; p.dumpA = dDump;
        cudaMemset(dC, 0xff, M * N * 4);
        k_probe<<<1, 128, 40000>>>(ma, mb, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<float> C(M * N), dump(4096);
        cudaMemcpy(C.data(), dC, M * N * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(dump.data(), dDump, 16384, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0; int nz = 0, nan = 0;
        for (int i = 0; i < M * N; ++i) {
            if (std::isnan(C[i])) { ++nan; continue; }
            maxerr = fmax(maxerr, fabs((double)C[i] - Cref[i])); maxref = fmax(maxref, fabs((double)Cref[i]));
            nz += C[i] != 0.f;
        }
        double dsum = 0; for (float v : dump) dsum += fabs(v);
        printf("%-70s rel err %.3e  nonzero %d nan %d  |smemA| sum %.1f (expect %.1f)\n", c.name, maxerr / maxref, nz, nan,
               dsum, [&] { double s = 0; for (float v : A) s += v; return s; }());
        if (c.mn && c.sw == 4 && c.lbo == 4096 && c.sbo == 512) {
            // 32-byte-atom swizzle: element (k, m) expected at block j, row k, 32-B chunk ((m%32)/8)^(k%4)
            int bad = 0;
            for (int k = 0; k < K; ++k)
                for (int m = 0; m < M; ++m) {
                    int j = m / 32, mm = m % 32, chunk = (mm / 8) ^ (k % 4);
                    float v = dump[j * 1024 + k * 32 + chunk * 8 + (mm % 8)];
                    bad += v != A[k * M + m];
                }
            printf("   MN-major smem image mismatches vs expected 32B-atom swizzle: %d\n", bad);
        }
        if (c.mn && c.sw == 3 && c.lbo == 4096 && c.sbo == 1024) {
            // layout check of the MN-major smem image: element (k, m) expected at block j=m/32, row k, chunk ((m%32)/4)^(k%8)
            int bad = 0;
            for (int k = 0; k < K; ++k)
                for (int m = 0; m < M; ++m) {
                    int j = m / 32, mm = m % 32, chunk = (mm / 4) ^ (k % 8);
                    float v = dump[j * 1024 + k * 32 + chunk * 4 + (mm % 4)];
                    bad += v != A[k * M + m];
                }
            printf("   MN-major smem image mismatches vs expected swizzle: %d\n", bad);
        }
    }
    return 0;
}
