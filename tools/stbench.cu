// Store-path microbenchmark for the gemm_nt epilogue (tuning tool, not part of the library).
// 148 persistent CTAs x 256 "epilogue" threads write a [M x N] bf16 matrix in the same (tile, slice, row-per-thread)
// order as the tcgen05 GEMM epilogue, with different store strategies:
//   0  row-per-thread 2 x STG.128 per 16 columns          1  row-per-thread STG.256
//   2  per-warp smem transpose, 8 rows x 64 B per instr   3  per-warp TMA store, 32 cols x 32 rows (SWIZZLE_64B)
//   4  per-warp TMA store, 64 cols x 32 rows (SWIZZLE_128B)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stbench tools/stbench.cu && ./stbench
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float val(int row, int col) { return static_cast<float>((row * 7 + col * 3) & 1023) * 0.125f; }

struct P {
    __nv_bfloat16* out;
    int M, N, ld, n_stride, n_slices, num_tiles;
};

template <int V>
__global__ void __launch_bounds__(256, 1) st_kernel(const __grid_constant__ CUtensorMap tm, const P p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3, half = warp >> 2;
    const int slice = blockIdx.x % p.n_slices, tile0 = blockIdx.x / p.n_slices, tstep = gridDim.x / p.n_slices;
    const int col0 = slice * p.n_stride;
    const int nch = p.n_stride / 32, mid = (nch + 1) / 2;
    const int ch0 = half ? mid : 0, ch1 = half ? nch : mid;
    uint8_t* wbuf = smem + warp * 8192;  // 2 x 4 KB per warp
    int bufi = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tstep) {
        const int r = quarter * 32 + lane;
        const int grow = tile * 128 + r;
        for (int ch = ch0; ch < ch1; ++ch) {
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = val(grow, col0 + ch * 32 + j);
            const int col = col0 + ch * 32;
            if (V == 0 || V == 1) {
                if (grow < p.M) {
                    __nv_bfloat16* o = p.out + static_cast<size_t>(grow) * p.ld + col;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = pack2(x[g * 16 + 2 * j], x[g * 16 + 2 * j + 1]);
                        if (V == 0) {
                            *reinterpret_cast<uint4*>(o + g * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                            *reinterpret_cast<uint4*>(o + g * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                        } else {
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + g * 16), "r"(w[0]),
                                         "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                         : "memory");
                        }
                    }
                }
            } else if (V == 2 || V == 3) {
                // 32 rows x 64 B, 16-byte chunk c of row r lives at r*64 + ((c ^ (r>>1)) & 3)*16  (== SWIZZLE_64B)
                uint8_t* buf = wbuf + bufi * 4096;
                if (V == 3) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the buffer used 2 chunks ago is free
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = make_uint4(pack2(x[c * 8], x[c * 8 + 1]), pack2(x[c * 8 + 2], x[c * 8 + 3]),
                                               pack2(x[c * 8 + 4], x[c * 8 + 5]), pack2(x[c * 8 + 6], x[c * 8 + 7]));
                    *reinterpret_cast<uint4*>(buf + lane * 64 + ((c ^ (lane >> 1)) & 3) * 16) = v;
                }
                if (V == 2) {
                    __syncwarp();
                    const int c = lane & 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + (lane >> 2);
                        const uint4 v = *reinterpret_cast<const uint4*>(buf + rr * 64 + ((c ^ (rr >> 1)) & 3) * 16);
                        const int gr = tile * 128 + quarter * 32 + rr;
                        if (gr < p.M) *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(gr) * p.ld + col + c * 8) = v;
                    }
                    __syncwarp();
                } else {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&tm)),
                                     "r"(smem_u32(buf)), "r"(col), "r"(tile * 128 + quarter * 32)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                bufi ^= 1;
            } else if (V == 4) {
                // 32 rows x 128 B (two 32-column chunks), SWIZZLE_128B: chunk c (0..7) of row r at r*128 + ((c ^ (r&7))*16)
                const int sub = (ch - ch0) & 1;
                uint8_t* buf = wbuf + bufi * 4096;
                if (sub == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = make_uint4(pack2(x[c * 8], x[c * 8 + 1]), pack2(x[c * 8 + 2], x[c * 8 + 3]),
                                               pack2(x[c * 8 + 4], x[c * 8 + 5]), pack2(x[c * 8 + 6], x[c * 8 + 7]));
                    *reinterpret_cast<uint4*>(buf + lane * 128 + (((sub * 4 + c) ^ (lane & 7)) * 16)) = v;
                }
                if (sub == 1 || ch == ch1 - 1) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&tm)),
                                     "r"(smem_u32(buf)), "r"(col - sub * 32), "r"(tile * 128 + quarter * 32)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    bufi ^= 1;
                }
            }
        }
    }
    if (V >= 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void cmp_kernel(const uint4* a, const uint4* b, size_t n, unsigned long long* bad) {
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    unsigned long long c = 0;
    for (; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 x = a[i], y = b[i];
        c += (x.x != y.x) + (x.y != y.y) + (x.z != y.z) + (x.w != y.w);
    }
    if (c) atomicAdd(bad, c);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(void* base, int rows, int cols, int ld, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
    static EncodeFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&fn), cudaEnableDefault, &q));
    }
    CUtensorMap m;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

template <int V>
static float run(const CUtensorMap& tm, const P& p, int grid, size_t smem, void* flush) {
    CK(cudaFuncSetAttribute(st_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaMemsetAsync(flush, it, 256u << 20));
        CK(cudaEventRecord(a));
        st_kernel<V><<<grid, 256, smem>>>(tm, p);
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    const int M = 563200, n_slices = 4, n_stride = 256, N = n_slices * n_stride;
    void* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    unsigned long long* bad;
    CK(cudaMalloc(&bad, 8));
    for (int ld : {1024, 1040}) {
        const size_t bytes = static_cast<size_t>(M) * ld * 2;
        __nv_bfloat16 *ref, *out;
        CK(cudaMalloc(&ref, bytes));
        CK(cudaMalloc(&out, bytes));
        CK(cudaMemset(ref, 0, bytes));
        P p{ref, M, N, ld, n_stride, n_slices, (M + 127) / 128};
        const int grid = 148;
        const size_t smem = 200 * 1024;
        CUtensorMap dummy = make_map(ref, M, N, ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        const double gb = static_cast<double>(M) * N * 2 / 1e9;
        float t0 = run<0>(dummy, p, grid, smem, flush);
        printf("ld=%d V0 STG.128 x2      %.3f ms  %.0f GB/s\n", ld, t0, gb / t0 * 1e3);
        p.out = out;
        auto check = [&](const char* name, float t) {
            CK(cudaMemset(bad, 0, 8));
            cmp_kernel<<<1024, 256>>>(reinterpret_cast<const uint4*>(ref), reinterpret_cast<const uint4*>(out), bytes / 16, bad);
            unsigned long long h;
            CK(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
            printf("ld=%d %-20s %.3f ms  %.0f GB/s  mismatched words %llu\n", ld, name, t, gb / t * 1e3, h);
            CK(cudaMemset(out, 0, bytes));
        };
        CK(cudaMemset(out, 0, bytes));
        check("V1 STG.256", run<1>(dummy, p, grid, smem, flush));
        check("V2 smem transpose", run<2>(dummy, p, grid, smem, flush));
        CUtensorMap m3 = make_map(out, M, N, ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        check("V3 TMA 32x32 sw64", run<3>(m3, p, grid, smem, flush));
        CUtensorMap m4 = make_map(out, M, N, ld, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
        check("V4 TMA 64x32 sw128", run<4>(m4, p, grid, smem, flush));
        CK(cudaFree(ref));
        CK(cudaFree(out));
    }
    return 0;
}
