// TMA load-rate microbenchmark (tuning tool): 148 persistent CTAs stream a [M x 320] bf16 matrix (row pitch 328)
// through a ring of S stages of 128 x 64 boxes (SWIZZLE_128B), exactly the A-operand traffic of gemm_nt, with a
// consumer thread that only waits and frees.  `reread` CTAs read the same tile (the weight-slice CTAs of one group).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I news-recommendation_b200/csrc -o build/tmabench tools/tmabench.cu
#define NR_OWNS_WATCHDOG 1
#include <stdio.h>
#include <stdlib.h>

#include "nr_common.cuh"

using namespace nr;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int num_tiles, int kchunks,
                                                    int stages, int reread) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16], empty[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int tile0 = blockIdx.x / reread, tstep = gridDim.x / reread;
    if (warp == 0 && lane == 0) {
        int st = 0; uint32_t ph = 0;
        for (int tile = tile0; tile < num_tiles; tile += tstep)
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(&empty[st], ph ^ 1, 1);
                mbar_arrive_expect_tx(&full[st], 16384);
                tma_load_2d(smem + st * 16384, &tm, &full[st], kc * 64, tile * 128);
                if (++st == stages) { st = 0; ph ^= 1; }
            }
    } else if (warp == 1 && lane == 0) {
        int st = 0; uint32_t ph = 0;
        for (int tile = tile0; tile < num_tiles; tile += tstep)
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(&full[st], ph, 2);
                mbar_arrive(&empty[st]);
                if (++st == stages) { st = 0; ph ^= 1; }
            }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeFn fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&fn), cudaEnableDefault, &q));
    void* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    const int M = 563200;
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int K = cfg == 0 ? 320 : 960, ld = cfg == 0 ? 328 : 968;
        void* A;
        CK(cudaMalloc(&A, static_cast<size_t>(M) * ld * 2));
        CK(cudaMemset(A, 0, static_cast<size_t>(M) * ld * 2));
        for (int promo = 0; promo < 2; ++promo) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(M)};
            cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
            cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
            if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
            CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * 16384 + 1024));
            for (int reread : {1, 4})
                for (int stages : {2, 4, 8, 12}) {
                    cudaEvent_t a, b;
                    CK(cudaEventCreate(&a));
                    CK(cudaEventCreate(&b));
                    float best = 1e9f;
                    for (int it = 0; it < 3; ++it) {
                        CK(cudaMemsetAsync(flush, it, 256u << 20));
                        CK(cudaEventRecord(a));
                        tma_kernel<<<148, 64, stages * 16384 + 1024>>>(tm, M / 128, K / 64, stages, reread);
                        CK(cudaEventRecord(b));
                        CK(cudaDeviceSynchronize());
                        float ms;
                        CK(cudaEventElapsedTime(&ms, a, b));
                        if (ms < best) best = ms;
                    }
                    const double gb = static_cast<double>(M) * K * 2 / 1e9;
                    printf("K=%3d promo=%s reread=%d stages=%2d  %.3f ms  HBM %.0f GB/s  L2->SM %.0f GB/s\n", K, promo ? "256B" : "128B",
                           reread, stages, best, gb / best * 1e3, gb * reread / best * 1e3);
                }
        }
        CK(cudaFree(A));
    }
    return 0;
}
