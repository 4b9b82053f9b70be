// tcgen05.mma issue-rate microbenchmark (tuning tool): one CTA per SM, one thread issues `iters` x `kblocks` x 4
// MMAs of shape 128 x N x 16 (bf16, SS mode, K-major SWIZZLE_128B operands resident in shared memory) and reports
// cycles per MMA.  Answers: what does one k-step cost at the slice widths the planner picks (80 / 160 / 240 / 256)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I news-recommendation_b200/csrc -o build/mmabench tools/mmabench.cu
#define NR_OWNS_WATCHDOG 1
#include <stdio.h>
#include <stdlib.h>

#include "nr_common.cuh"

using namespace nr;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

// mode 0: A and B both from smem, new A stage every k-block (ring of `stages`), B slice resident
// mode bit0: tcgen05.commit to an mbarrier after every k-block (as the GEMM's stage release does)
// mode bit2: warps 4-11 poll an mbarrier with try_wait (the GEMM's epilogue warps waiting for an accumulator)
// mode bit1: warp 2 streams 16 KB bulk copies global->smem (the GEMM's TMA producer traffic) while the MMAs run
__global__ void __launch_bounds__(384, 1) mma_kernel(int N, int kblocks, int stages, int iters, long long* out, int mode,
                                                     const uint8_t* gsrc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, kbar[8], cbar[2];
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sB = smem;                         // kblocks boxes of N x 128 B
    uint8_t* sA = smem + kblocks * N * 128;     // stages x 16 KB
    for (int i = threadIdx.x; i < (kblocks * N * 128 + stages * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (threadIdx.x == 32) {
        mbar_init(&bar, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&kbar[i], 1);
        mbar_init(&cbar[0], 1);
        mbar_init(&cbar[1], 1);
        done = 0;
        fence_barrier_init();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const long long t0 = clock64();
        int st = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (it & 1) * 256;
            uint32_t acc = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const uint32_t a = smem_u32(sA + st * 16384), b = smem_u32(sB + kb * N * 128);
                for (int k = 0; k < 4; ++k) {
                    umma_bf16(d, make_sw128_desc(a + k * 32, 0, 1024), make_sw128_desc(b + k * 32, 0, 1024), idesc, acc);
                    acc = 1;
                }
                if (mode & 1) umma_commit(&kbar[st]);
                if (++st == stages) st = 0;
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0, 1);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        done = 1;
    } else if (warp == 2 && lane == 0 && (mode & 2)) {
        // two 16 KB landing buffers behind the operands; keep 2 copies in flight until the MMA thread is done
        uint8_t* land = sA + stages * 16384;
        uint32_t ph[2] = {0, 0};
        long long n = 0;
        const uint8_t* src = gsrc + static_cast<size_t>(blockIdx.x) * (8u << 20);
        for (int b = 0; b < 2; ++b) {
            mbar_arrive_expect_tx(&cbar[b], 16384);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(land + b * 16384)), "l"(src + (n++ % 512) * 16384), "r"(16384), "r"(smem_u32(&cbar[b]))
                         : "memory");
        }
        int b = 0;
        while (!done) {
            mbar_wait(&cbar[b], ph[b], 2);
            ph[b] ^= 1;
            mbar_arrive_expect_tx(&cbar[b], 16384);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(land + b * 16384)), "l"(src + (n++ % 512) * 16384), "r"(16384), "r"(smem_u32(&cbar[b]))
                         : "memory");
            b ^= 1;
        }
        mbar_wait(&cbar[0], ph[0], 3);
        mbar_wait(&cbar[1], ph[1], 3);
        out[148 + blockIdx.x] = n;
    }
    else if (warp >= 4 && (mode & 4)) {
        while (!done) {
            if (mbar_try_wait(&kbar[7], 1 ^ 1)) break;  // phase 0 never completes (nobody arrives on kbar[7] when stages <= 7)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    CK(cudaMalloc(&out, 2 * 148 * 8));
    uint8_t* gsrc;
    CK(cudaMalloc(&gsrc, 148ull * (8u << 20)));
    CK(cudaMemset(gsrc, 0, 148ull * (8u << 20)));
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    const int shapes[][2] = {{32, 5}, {48, 5}, {64, 5}, {128, 5}, {160, 4}, {240, 5}};
    for (auto& s : shapes) {
        const int N = s[0], kb = s[1], stages = 4, iters = 400;
        const size_t smem = static_cast<size_t>(kb) * N * 128 + (stages + 2) * 16384 + 1024;
        for (int mode : {0, 1, 5}) {
            const int grid = 148;
            mma_kernel<<<grid, 384, smem>>>(N, kb, stages, iters, out, mode, gsrc);
            CK(cudaDeviceSynchronize());
            long long h[296];
            CK(cudaMemcpy(h, out, 2 * 148 * 8, cudaMemcpyDeviceToHost));
            double avg = 0, copies = 0;
            for (int i = 0; i < grid; ++i) { avg += h[i]; copies += h[148 + i]; }
            avg /= grid;
            copies /= grid;
            const double per = avg / (iters * kb * 4.0);
            printf("N=%3d kblocks=%2d mode=%d (commit/kblock=%d, bulk copies=%d)  %.1f cyc/MMA  (floor %d)  %.0f%% of peak", N, kb,
                   mode, mode & 1, (mode >> 1) & 1, per, N / 2, 100.0 * (N / 2.0) / per);
            if (mode & 2) printf("  copy traffic %.1f B/cyc/SM", copies * 16384.0 / avg);
            printf("\n");
        }
    }
    return 0;
}
