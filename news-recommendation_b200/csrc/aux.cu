// Memory-bound companions of the tcgen05 GEMMs: operand preparation, the embedding-row gather,
// the per-(sequence, head) self-attention core (fwd/bwd), pooling-backward row dots, the dot-product
// click scorer.  All HBM-bound integer/byte or small-reduction work: coalesced 16-byte accesses,
// warp-shuffle reductions, no tensor cores.
#include <algorithm>

#include "nr_common.cuh"
#include "nr_ops.h"

namespace nr {

extern int g_launches;

// ------------------------------------------------------------------------------------------------
// fp32 parameter -> zero-padded bf16 operand (optionally transposed)
// ------------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, int R, int C, int lds, __nv_bfloat16* __restrict__ dst,
                                int ld, int transpose) {
    const long long n_rows = transpose ? C : R;
    const long long total = n_rows * ld;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / ld;
        const int c = static_cast<int>(i - r * ld);
        float v = 0.f;
        if (!transpose) {
            if (c < C) v = src[r * lds + c];
        } else {
            if (c < R) v = src[static_cast<long long>(c) * lds + r];
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}
// non-transposed, 16-byte aligned rows: one thread per 8-column chunk (two float4 loads -> one 16-byte store), no division
// per element.  The embedding table (85 MB of fp32 -> 43 MB of bf16) is rebuilt after every optimizer step: 0.07 -> 0.03 ms.
__global__ void __launch_bounds__(256) cast_pad_rows_kernel(const float* __restrict__ src, long long R, int C, int lds,
                                                            __nv_bfloat16* __restrict__ dst, int ld) {
    const int chunks = ld >> 3;
    const long long total = R * chunks;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const long long r = i / chunks;
        const int col = static_cast<int>(i - r * chunks) * 8;
        const float* sp = src + r * lds + col;
        float v[8];
        if (col + 8 <= C) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(sp)), b = __ldg(reinterpret_cast<const float4*>(sp) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = col + j < C ? sp[j] : 0.f;
        }
        *reinterpret_cast<uint4*>(dst + r * ld + col) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}
int cast_pad_bf16(const float* src, int R, int C, int lds, void* dst, int ld, int transpose, cudaStream_t stream) {
    const long long total = static_cast<long long>(transpose ? C : R) * ld;
    if (total == 0) return 0;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    ProfScope ps("cast_pad", R, C, ld, stream);
    if (!transpose && (ld & 7) == 0 && (lds & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const long long items = static_cast<long long>(R) * (ld >> 3);
        cast_pad_rows_kernel<<<static_cast<int>(std::min<long long>((items + 255) / 256, 148 * 16)), 256, 0, stream>>>(
            src, R, C, lds, static_cast<__nv_bfloat16*>(dst), ld);
        ++g_launches;
        NR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    cast_pad_kernel<<<blocks, 256, 0, stream>>>(src, R, C, lds, static_cast<__nv_bfloat16*>(dst), ld, transpose);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Several small operand casts in ONE launch (blockIdx.y = matrix): the packed / transposed projection and pooling weights of
// an encoder are four matrices of a few hundred rows each -- as four launches they cost more in launch gaps than in work,
// and they are rebuilt after every optimizer step.
constexpr int kCastMany = 8;
struct CastManyArgs {
    const float* src[kCastMany];
    __nv_bfloat16* dst[kCastMany];
    int R[kCastMany], C[kCastMany], lds[kCastMany], ld[kCastMany], transpose[kCastMany];
};
__global__ void __launch_bounds__(256) cast_pad_many_kernel(CastManyArgs a) {
    const int m = blockIdx.y;
    const float* __restrict__ src = a.src[m];
    __nv_bfloat16* __restrict__ dst = a.dst[m];
    const int R = a.R[m], C = a.C[m], lds = a.lds[m], ld = a.ld[m], transpose = a.transpose[m];
    const long long total = static_cast<long long>(transpose ? C : R) * ld;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const long long r = i / ld;
        const int c = static_cast<int>(i - r * ld);
        float v = 0.f;
        if (!transpose) {
            if (c < C) v = src[r * lds + c];
        } else {
            if (c < R) v = src[static_cast<long long>(c) * lds + r];
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}
int cast_pad_bf16_many(int n, const float* const* src, const int* R, const int* C, const int* lds, void* const* dst, const int* ld,
                       const int* transpose, cudaStream_t stream) {
    NR_REQUIRE(n >= 0 && n <= kCastMany, "cast_pad_bf16_many: %d matrices (at most %d per call)", n, kCastMany);
    if (n == 0) return 0;
    CastManyArgs a;
    long long biggest = 0;
    for (int i = 0; i < n; ++i) {
        NR_REQUIRE(src[i] && dst[i] && R[i] >= 1 && C[i] >= 1 && ld[i] >= (transpose[i] ? R[i] : C[i]), "cast_pad_bf16_many: bad matrix %d", i);
        a.src[i] = src[i];
        a.dst[i] = static_cast<__nv_bfloat16*>(dst[i]);
        a.R[i] = R[i], a.C[i] = C[i], a.lds[i] = lds[i], a.ld[i] = ld[i], a.transpose[i] = transpose[i];
        biggest = std::max(biggest, static_cast<long long>(transpose[i] ? C[i] : R[i]) * ld[i]);
    }
    ProfScope ps("cast_pad_many", n, static_cast<int>(std::min<long long>(biggest, 1 << 30)), 0, stream);
    const dim3 grid(static_cast<unsigned>(std::min<long long>((biggest + 255) / 256, 148 * 4)), static_cast<unsigned>(n));
    cast_pad_many_kernel<<<grid, 256, 0, stream>>>(a);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// fp32 rows [n_seq][T][D] (arbitrary element strides) -> bf16 rows [n_seq*T x ld] with a ones column at D.
// One thread per 8-column chunk column, a block walks kRowsThreads / chunks rows per iteration (no division in the loop).
constexpr int kRowsThreads = 320;
__global__ void __launch_bounds__(kRowsThreads) rows_to_bf16_kernel(const float* __restrict__ src, long long n_rows, int T, int D,
                                                                  long long s_seq, long long s_tok, long long s_col,
                                                                  __nv_bfloat16* __restrict__ dst, int ld) {
    const int chunks = ld >> 3;
    const int rows_per_it = kRowsThreads / chunks;
    const int rl = threadIdx.x / chunks, c = threadIdx.x - rl * chunks;
    if (rl >= rows_per_it) return;
    const int col = c * 8;
    for (long long r = static_cast<long long>(blockIdx.x) * rows_per_it + rl; r < n_rows;
         r += static_cast<long long>(gridDim.x) * rows_per_it) {
        const long long seq = r / T;
        const long long tok = r - seq * T;
        const float* sp = src + seq * s_seq + tok * s_tok;
        float v[8];
        if (s_col == 1 && col + 8 <= D && ((reinterpret_cast<uintptr_t>(sp + col) & 15) == 0)) {
            const float4 a = *reinterpret_cast<const float4*>(sp + col), b = *reinterpret_cast<const float4*>(sp + col + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = col + j < D ? sp[(col + j) * s_col] : (col + j == D ? 1.0f : 0.f);
        }
        *reinterpret_cast<uint4*>(dst + r * ld + col) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}
int rows_to_bf16(const float* src, long long n_seq, int T, int D, long long s_seq, long long s_tok, long long s_col,
                 void* dst, int ld, cudaStream_t stream) {
    const long long n = n_seq * T;
    if (n == 0) return 0;
    NR_REQUIRE(ld >= D + 1 && ld % 8 == 0 && ld / 8 <= kRowsThreads, "rows_to_bf16: pitch %d for D=%d plus the ones column", ld, D);
    const int rows_per_it = kRowsThreads / (ld / 8);
    const int blocks = static_cast<int>(std::min<long long>((n + rows_per_it - 1) / rows_per_it, 148 * 6));
    ProfScope ps("rows_to_bf16", static_cast<int>(n), D, ld, stream);
    rows_to_bf16_kernel<<<blocks, kRowsThreads, 0, stream>>>(src, n, T, D, s_seq, s_tok, s_col, static_cast<__nv_bfloat16*>(dst),
                                                             ld);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Embedding-row gather: one warp per token, 16-byte lanes.  The copy is bit exact (bf16 table rows).
// Column D of every gathered row is set to 1.0 (bias-gradient trick), columns after it stay 0.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float drop_mult(uint64_t seed, uint32_t thresh, float scale, long long row, int ld, int col) {
    const uint64_t bits = dropout_bits4(seed, (static_cast<uint64_t>(row) * ld + col) >> 2);
    return (((bits >> (16 * (col & 3))) & 0xffffu) >= thresh) ? scale : 0.f;
}

// one thread = one 16-byte chunk COLUMN: a block covers kGatherThreads / chunks consecutive token rows per iteration
// (consecutive threads copy consecutive chunks: coalesced), so the chunk index and the row slot are computed once and
// the loop carries no division (the flat-index version spent most of its instructions on 64-bit div/mod);
// dropout draws ONE counter hash per 4 aligned columns (two per chunk)
constexpr int kGatherThreads = 320;
__global__ void __launch_bounds__(kGatherThreads) gather_rows_kernel(const long long* __restrict__ ids, long long n_tok, int T,
                                                                   const uint4* __restrict__ table, int V, int D, int ld,
                                                                   uint4* __restrict__ X, int padded, float p, uint64_t seed,
                                                                   int* bad_flag) {
    const int chunks = ld >> 3;  // 16-byte chunks per row
    const int rows_per_it = kGatherThreads / chunks;
    const int rl = threadIdx.x / chunks, c = threadIdx.x - rl * chunks;
    if (rl >= rows_per_it) return;
    const uint32_t thresh = static_cast<uint32_t>(p * 65536.0f + 0.5f);
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    const long long stride = static_cast<long long>(gridDim.x) * rows_per_it;
    long long tok = static_cast<long long>(blockIdx.x) * rows_per_it + rl;
    // software pipeline: the (id -> table row) loads of the NEXT row are in flight while this one is stored
    long long n_id = tok < n_tok ? __ldg(ids + tok) : 0;
    if (tok < n_tok && (n_id < 0 || n_id >= V)) { atomicExch(bad_flag, 1); n_id = 0; }
    uint4 n_u = tok < n_tok ? __ldg(table + n_id * chunks + c) : make_uint4(0, 0, 0, 0);
    const int col = c * 8;
    for (; tok < n_tok; tok += stride) {
        uint4 u = n_u;
        const long long tok2 = tok + stride;
        if (tok2 < n_tok) {
            n_id = __ldg(ids + tok2);
            if (n_id < 0 || n_id >= V) { atomicExch(bad_flag, 1); n_id = 0; }
            n_u = __ldg(table + n_id * chunks + c);
        }
        long long xr = tok;
        int t = 0;
        if (padded) {
            const long long seg = tok / T;
            t = static_cast<int>(tok - seg * T);
            xr = seg * (T + 2) + 1 + t;
        }
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
        if (p > 0.f) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t bits = dropout_bits4(seed, static_cast<uint64_t>(xr) * (ld >> 2) + (col >> 2) + h);  // ld % 8 == 0
                const uint32_t lo = static_cast<uint32_t>(bits), hi = static_cast<uint32_t>(bits >> 32);
                float2 f0 = unpack_bf16x2(w[2 * h]), f1 = unpack_bf16x2(w[2 * h + 1]);
                f0.x *= ((lo & 0xffffu) >= thresh) ? scale : 0.f;
                f0.y *= ((lo >> 16) >= thresh) ? scale : 0.f;
                f1.x *= ((hi & 0xffffu) >= thresh) ? scale : 0.f;
                f1.y *= ((hi >> 16) >= thresh) ? scale : 0.f;
                w[2 * h] = pack_bf16x2(f0.x, f0.y);
                w[2 * h + 1] = pack_bf16x2(f1.x, f1.y);
            }
        }
        if (D >= col && D < col + 8) {  // ones column, zeros behind it
            __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(w);
            for (int j = D - col; j < 8; ++j) e[j] = __float2bfloat16_rn(j == D - col ? 1.0f : 0.f);
        }
        X[xr * chunks + c] = make_uint4(w[0], w[1], w[2], w[3]);
        if (padded) {
            if (t == 0) X[(xr - 1) * chunks + c] = make_uint4(0, 0, 0, 0);
            if (t == T - 1) X[(xr + 1) * chunks + c] = make_uint4(0, 0, 0, 0);
        }
    }
}
int gather_rows(const long long* ids, long long n_tok, int T, const void* table, int V, int D, int ld_table, void* X,
                int ld_x, int padded, DropoutCfg drop, int* bad_id_flag, cudaStream_t stream) {
    if (n_tok == 0) return 0;
    NR_REQUIRE(ld_table == ld_x && ld_x % 8 == 0 && ld_x >= D + 1 && ld_x / 8 <= kGatherThreads,
               "gather_rows: pitch %d/%d for D=%d", ld_table, ld_x, D);
    const int rows_per_it = kGatherThreads / (ld_x / 8);
    const int blocks = static_cast<int>(std::min<long long>((n_tok + rows_per_it - 1) / rows_per_it, 148 * 6));
    ProfScope ps("gather_rows", static_cast<int>(n_tok), D, ld_x, stream);
    gather_rows_kernel<<<blocks, kGatherThreads, 0, stream>>>(ids, n_tok, T, static_cast<const uint4*>(table), V, D, ld_x,
                                                   static_cast<uint4*>(X), padded, drop.p, drop.seed, bad_id_flag);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pooling backward, scalar part: dw_r = dOut[seg] . X_r ; dscore_r = w_r (dw_r - sum_seg w dw)
// ------------------------------------------------------------------------------------------------
// Two phases inside one block (= a few segments): (1) every warp takes rows round-robin and keeps the 16-byte loads of
// up to 4 rows in flight before reducing (the first version walked rows one by one and was latency bound at ~1.3 TB/s);
// (2) one thread per row turns the row dots into dscore.
constexpr int kDsSegs = 6;    // segments per block iteration
__global__ void __launch_bounds__(256) pool_dscore_kernel(const __nv_bfloat16* __restrict__ X, int lda, int D,
                                                          long long n_seg, int seg_len, const float* __restrict__ w,
                                                          const float* __restrict__ dout, int ldo,
                                                          float* __restrict__ dscore) {
    extern __shared__ float s_dw[];  // [kDsSegs * seg_len]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = D >> 3;
    const long long n_groups = (n_seg + kDsSegs - 1) / kDsSegs;
    for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const long long seg0 = grp * kDsSegs;
        const int nsg = static_cast<int>(min(static_cast<long long>(kDsSegs), n_seg - seg0));
        const int nrows = nsg * seg_len;
        for (int r0 = warp * 4; r0 < nrows; r0 += 8 * 4) {
            uint4 u[4][2];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = r0 + k;
                const uint4* xr = reinterpret_cast<const uint4*>(X + (seg0 * seg_len + r) * static_cast<long long>(lda));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = lane + 32 * h;
                    u[k][h] = (r < nrows && c < chunks) ? __ldg(xr + c) : make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = r0 + k;
                if (r >= nrows) break;
                const float* dob = dout + (seg0 + r / seg_len) * ldo;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = lane + 32 * h;
                    if (c < chunks) {
                        const float4 d0 = *reinterpret_cast<const float4*>(dob + c * 8);
                        const float4 d1 = *reinterpret_cast<const float4*>(dob + c * 8 + 4);
                        const uint4 v = u[k][h];
                        const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
                        float a = acc[k];
                        a = fmaf(f0.x, d0.x, a); a = fmaf(f0.y, d0.y, a); a = fmaf(f1.x, d0.z, a); a = fmaf(f1.y, d0.w, a);
                        a = fmaf(f2.x, d1.x, a); a = fmaf(f2.y, d1.y, a); a = fmaf(f3.x, d1.z, a); a = fmaf(f3.y, d1.w, a);
                        acc[k] = a;
                    }
                }
                // columns beyond 64 chunks (D > 512) and the D % 8 tail
                const __nv_bfloat16* xe = X + (seg0 * seg_len + r) * static_cast<long long>(lda);
                for (int c = 64 + lane; c < chunks; c += 32) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xe) + c);
                    const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16x2(vw[j]);
                        acc[k] = fmaf(f.x, dob[c * 8 + 2 * j], acc[k]);
                        acc[k] = fmaf(f.y, dob[c * 8 + 2 * j + 1], acc[k]);
                    }
                }
                for (int c = chunks * 8 + lane; c < D; c += 32) acc[k] = fmaf(__bfloat162float(xe[c]), dob[c], acc[k]);
                const float tot = warp_sum(acc[k]);
                if (lane == 0) s_dw[r] = tot;
            }
        }
        __syncthreads();
        for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
            const int sgi = r / seg_len;
            const float* wr = w + (seg0 + sgi) * seg_len;
            const float* dwr = s_dw + sgi * seg_len;
            float dot = 0.f;
            for (int t = 0; t < seg_len; ++t) dot = fmaf(wr[t], dwr[t], dot);
            dscore[seg0 * seg_len + r] = wr[r - sgi * seg_len] * (s_dw[r] - dot);
        }
        __syncthreads();
    }
}
// Title-level shape (seg_len <= 32, D <= 512): one WARP per segment, no shared memory and no block barrier.  The
// segment's dOut chunk stays in registers for all its rows, 5 rows x 2 16-byte loads are in flight per lane, and the
// per-row totals come out of one 31-shuffle transpose reduction (lane t ends with dw_t) instead of 5 shuffles per row.
__global__ void __launch_bounds__(256) pool_dscore_warp_kernel(const __nv_bfloat16* __restrict__ X, int lda, int D,
                                                               long long n_seg, int seg_len, const float* __restrict__ w,
                                                               const float* __restrict__ dout, int ldo,
                                                               float* __restrict__ dscore) {
    const int lane = threadIdx.x & 31;
    const int chunks = (D + 7) >> 3;  // the last one may be half valid (D % 8 == 4): its dOut half is zeroed below
    const long long wstride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    for (long long seg = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < n_seg; seg += wstride) {
        const float* dob = dout + seg * ldo;
        float4 dd[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            dd[h][0] = (c < chunks && c * 8 + 4 <= D) ? *reinterpret_cast<const float4*>(dob + c * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
            dd[h][1] = (c < chunks && c * 8 + 8 <= D) ? *reinterpret_cast<const float4*>(dob + c * 8 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const uint4* xs = reinterpret_cast<const uint4*>(X + seg * seg_len * static_cast<long long>(lda));
        const int pitch16 = lda >> 3;
        float pr[32];
#pragma unroll
        for (int tb = 0; tb < 32; tb += 4) {
            uint4 u[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = tb + k;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = lane + 32 * h;
                    u[k][h] = (t < seg_len && c < chunks) ? __ldg(xs + static_cast<long long>(t) * pitch16 + c) : make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 v = u[k][h];
                    const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
                    a = fmaf(f0.x, dd[h][0].x, a); a = fmaf(f0.y, dd[h][0].y, a); a = fmaf(f1.x, dd[h][0].z, a); a = fmaf(f1.y, dd[h][0].w, a);
                    a = fmaf(f2.x, dd[h][1].x, a); a = fmaf(f2.y, dd[h][1].y, a); a = fmaf(f3.x, dd[h][1].z, a); a = fmaf(f3.y, dd[h][1].w, a);
                }
                pr[tb + k] = a;
            }
            if (tb + 4 >= seg_len) {  // warp-uniform: the remaining row slots stay zero
#pragma unroll
                for (int t = tb + 4; t < 32; ++t) pr[t] = 0.f;
                break;
            }
        }
        const float dw = warp_transpose_sum32(pr);  // lane t: dOut . X_t
        const float wt = lane < seg_len ? __ldg(w + seg * seg_len + lane) : 0.f;
        const float dot = warp_sum(wt * dw);
        if (lane < seg_len) dscore[seg * seg_len + lane] = wt * (dw - dot);
    }
}

int pool_dscore(const void* X, int lda, int D, long long n_seg, int seg_len, const float* w, const float* dout, int ldo,
                float* dscore, cudaStream_t stream) {
    if (n_seg == 0) return 0;
    NR_REQUIRE(seg_len <= 128 && D % 4 == 0 && ldo % 4 == 0 && lda % 8 == 0, "pool_dscore: seg_len=%d D=%d ldo=%d lda=%d", seg_len, D, ldo, lda);
    const long long groups = (n_seg + kDsSegs - 1) / kDsSegs;
    const int blocks = static_cast<int>(std::min<long long>(groups, 148 * 8));
    ProfScope ps("pool_dscore", static_cast<int>(n_seg), seg_len, D, stream);
    if (seg_len <= 32 && D <= 512) {
        const int wblocks = static_cast<int>(std::min<long long>((n_seg + 7) / 8, 148 * 8));
        pool_dscore_warp_kernel<<<wblocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(X), lda, D, n_seg, seg_len, w, dout, ldo,
                                                             dscore);
        ++g_launches;
        NR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    pool_dscore_kernel<<<blocks, 256, sizeof(float) * kDsSegs * seg_len, stream>>>(
        static_cast<const __nv_bfloat16*>(X), lda, D, n_seg, seg_len, w, dout, ldo, dscore);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Extended weight gradient [rows][ld] (columns [0,D) = dW, column D = db from the ones-column trick) -> accumulated
// into the parameters' own .grad storage, and cleared for the next step (the buffer is a persistent workspace).
// Replaces, per weight, ~4 framework kernels (2 slice copies + 2 AccumulateGrad adds + the zero fill).
// ------------------------------------------------------------------------------------------------
__global__ void accumulate_ext_grad_kernel(float* __restrict__ ext, int rows, int ld, int D, float* __restrict__ dW,
                                           float* __restrict__ db) {
    const long long n = static_cast<long long>(rows) * (D + 1);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / (D + 1)), c = static_cast<int>(i - static_cast<long long>(r) * (D + 1));
        float* src = ext + static_cast<size_t>(r) * ld + c;
        const float v = *src;
        *src = 0.f;
        if (c < D) dW[static_cast<size_t>(r) * D + c] += v;
        else if (db != nullptr) db[r] += v;
    }
}
int accumulate_ext_grad(float* ext, int rows, int ld, int D, float* dW, float* db, cudaStream_t stream) {
    if (rows == 0) return 0;
    NR_REQUIRE(ld >= D + 1, "accumulate_ext_grad: pitch %d < D+1 = %d", ld, D + 1);
    const long long n = static_cast<long long>(rows) * (D + 1);
    const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
    accumulate_ext_grad_kernel<<<blocks, 256, 0, stream>>>(ext, rows, ld, D, dW, db);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ReLU backward + cast:  dst = dy * (relu_out > 0)  -> zero padded bf16 rows
// ------------------------------------------------------------------------------------------------
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ ro, long long n, int N, int ld_dy,
                                __nv_bfloat16* __restrict__ dst, int ldn) {
    const long long total = n * ldn;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / ldn;
        const int c = static_cast<int>(i - r * ldn);
        float v = 0.f;
        if (c < N) {
            v = dy[r * ld_dy + c];
            if (ro != nullptr && !(ro[r * ld_dy + c] > 0.f)) v = 0.f;
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}
int relu_bwd_to_bf16(const float* dy, const float* relu_out, long long n, int N, int ld_dy, void* dst, int ldn, cudaStream_t stream) {
    if (n == 0) return 0;
    ProfScope ps("relu_bwd_to_bf16", static_cast<int>(n), N, ldn, stream);
    const int blocks = static_cast<int>(std::min<long long>((n * ldn + 255) / 256, 148 * 16));
    relu_bwd_kernel<<<blocks, 256, 0, stream>>>(dy, relu_out, n, N, ld_dy, static_cast<__nv_bfloat16*>(dst), ldn);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 embedding rows (category / user tables): gather forward, red.add scatter backward
// ------------------------------------------------------------------------------------------------
__global__ void emb_f32_fwd_kernel(const long long* __restrict__ ids, long long n, const float* __restrict__ table, int V, int D,
                                   float* __restrict__ out, int* bad_flag) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long i = w0; i < n; i += nw) {
        long long id = ids[i];
        if (id < 0 || id >= V) {
            if (lane == 0) atomicExch(bad_flag, 1);
            id = 0;
        }
        for (int c = lane; c < D; c += 32) out[i * D + c] = table[id * D + c];
    }
}
__global__ void emb_f32_bwd_kernel(const long long* __restrict__ ids, long long n, const float* __restrict__ dout, int V, int D,
                                   float* __restrict__ dtable) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long i = w0; i < n; i += nw) {
        const long long id = ids[i];
        if (id <= 0 || id >= V) continue;  // padding_idx; out-of-range ids were flagged by the forward lookup
        for (int c = lane; c < D; c += 32) red_add_f32(dtable + id * D + c, dout[i * D + c]);
    }
}
int embedding_f32_fwd(const long long* ids, long long n, const float* table, int V, int D, float* out, int* bad_id_flag,
                      cudaStream_t stream) {
    if (n == 0) return 0;
    ProfScope ps("embedding_f32_fwd", static_cast<int>(n), D, V, stream);
    const int blocks = static_cast<int>(std::min<long long>((n + 7) / 8, 148 * 8));
    emb_f32_fwd_kernel<<<blocks, 256, 0, stream>>>(ids, n, table, V, D, out, bad_id_flag);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int embedding_f32_bwd(const long long* ids, long long n, const float* dout, int V, int D, float* dtable, cudaStream_t stream) {
    if (n == 0) return 0;
    ProfScope ps("embedding_f32_bwd", static_cast<int>(n), D, 0, stream);
    const int blocks = static_cast<int>(std::min<long long>((n + 7) / 8, 148 * 8));
    emb_f32_bwd_kernel<<<blocks, 256, 0, stream>>>(ids, n, dout, V, D, dtable);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// dot-product click predictor (reference dot_product.py:8-19): one warp per (impression, candidate)
// ------------------------------------------------------------------------------------------------
__global__ void dot_fwd_kernel(const float* __restrict__ cand, const float* __restrict__ user, int B, int C, int D,
                               float* __restrict__ logits) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= B * C) return;
    const int b = wid / C;
    const float* cv = cand + static_cast<size_t>(wid) * D;
    const float* uv = user + static_cast<size_t>(b) * D;
    float a = 0.f;
    for (int c = lane; c < D; c += 32) a = fmaf(cv[c], uv[c], a);
    a = warp_sum(a);
    if (lane == 0) logits[wid] = a;
}
__global__ void dot_bwd_kernel(const float* __restrict__ cand, const float* __restrict__ user,
                               const float* __restrict__ dlogits, int B, int C, int D, float* __restrict__ dcand,
                               float* __restrict__ duser) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const float u = user[static_cast<size_t>(b) * D + c];
        float du = 0.f;
        for (int j = 0; j < C; ++j) {
            const float g = dlogits[b * C + j];
            dcand[(static_cast<size_t>(b) * C + j) * D + c] = g * u;
            du = fmaf(g, cand[(static_cast<size_t>(b) * C + j) * D + c], du);
        }
        duser[static_cast<size_t>(b) * D + c] = du;
    }
}
int dot_score_fwd(const float* cand, const float* user, int B, int C, int D, float* logits, cudaStream_t stream) {
    if (B * C == 0) return 0;
    ProfScope ps("dot_fwd", B, C, D, stream);
    dot_fwd_kernel<<<ceil_div(B * C * 32, 256), 256, 0, stream>>>(cand, user, B, C, D, logits);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int dot_score_bwd(const float* cand, const float* user, const float* dlogits, int B, int C, int D, float* dcand,
                  float* duser, cudaStream_t stream) {
    if (B == 0) return 0;
    ProfScope ps("dot_bwd", B, C, D, stream);
    dot_bwd_kernel<<<B, 128, 0, stream>>>(cand, user, dlogits, B, C, D, dcand, duser);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Precise user encoder (NRMS precise mode): operand and attention kernels that keep the history-level path at fp32 accuracy.
// The level is 4.6 % of the model's FLOPs (512 users x 50 news vectors), so plain CUDA-core arithmetic is affordable.
// ------------------------------------------------------------------------------------------------
// fp32 rows [n_seq][T][D] (element strides) -> bf16 [rows][2*ld]: columns [0, D) = hi = bf16(x), column D = 1.0, zeros up to
// ld; columns [ld, ld + D) = lo = bf16(x - hi), zeros up to 2*ld.  Against the K-concatenated weight operand [W | W] the GEMM
// computes (hi + lo) . W^T: the input enters with ~16 mantissa bits instead of 8.
__global__ void __launch_bounds__(256) rows_to_bf16_hilo_kernel(const float* __restrict__ src, long long n_rows, int T, int D,
                                                                long long s_seq, long long s_tok, long long s_col,
                                                                __nv_bfloat16* __restrict__ dst, int ld) {
    const int chunks = (2 * ld) >> 3;
    const long long total = n_rows * chunks;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / chunks;
        const int c = static_cast<int>(i - r * chunks) * 8;
        const bool lo = c >= ld;
        const int col = lo ? c - ld : c;
        const long long seq = r / T;
        const float* sp = src + seq * s_seq + (r - seq * T) * s_tok;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cc = col + j;
            float x = cc < D ? sp[cc * s_col] : 0.f;
            const float hi = bf16_round(x);
            v[j] = lo ? (x - hi) : (cc == D ? 1.0f : hi);
        }
        *reinterpret_cast<uint4*>(dst + r * (2 * ld) + c) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}
int rows_to_bf16_hilo(const float* src, long long n_seq, int T, int D, long long s_seq, long long s_tok, long long s_col, void* dst,
                      int ld, cudaStream_t stream) {
    const long long n = n_seq * T;
    if (n == 0) return 0;
    NR_REQUIRE(ld >= D + 1 && ld % 8 == 0, "rows_to_bf16_hilo: pitch %d for D=%d plus the ones column", ld, D);
    ProfScope ps("rows_to_bf16_hilo", static_cast<int>(n), D, ld, stream);
    const int blocks = static_cast<int>(std::min<long long>((n * (2 * ld / 8) + 255) / 256, 148 * 8));
    rows_to_bf16_hilo_kernel<<<blocks, 256, 0, stream>>>(src, n, T, D, s_seq, s_tok, s_col, static_cast<__nv_bfloat16*>(dst), ld);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// the low plane alone: dst bf16 [rows][ld], columns [0, D) = bf16(x - bf16(x)), zeros up to ld (no ones column: the bias
// belongs to the hi pass).  Operand of the second pass of a two-pass hi/lo product (gru.cu).
__global__ void __launch_bounds__(256) rows_to_bf16_lo_kernel(const float* __restrict__ src, long long n_rows, int T, int D,
                                                              long long s_seq, long long s_tok, long long s_col,
                                                              __nv_bfloat16* __restrict__ dst, int ld) {
    const int chunks = ld >> 3;
    const long long total = n_rows * chunks;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const long long r = i / chunks;
        const int col = static_cast<int>(i - r * chunks) * 8;
        const long long seq = r / T;
        const float* sp = src + seq * s_seq + (r - seq * T) * s_tok;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x = col + j < D ? sp[(col + j) * s_col] : 0.f;
            v[j] = x - bf16_round(x);
        }
        *reinterpret_cast<uint4*>(dst + r * ld + col) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}
int rows_to_bf16_lo(const float* src, long long n_seq, int T, int D, long long s_seq, long long s_tok, long long s_col, void* dst, int ld,
                    cudaStream_t stream) {
    const long long n = n_seq * T;
    if (n == 0) return 0;
    NR_REQUIRE(ld >= D && ld % 8 == 0, "rows_to_bf16_lo: pitch %d for D=%d", ld, D);
    ProfScope ps("rows_to_bf16_lo", static_cast<int>(n), D, ld, stream);
    const int blocks = static_cast<int>(std::min<long long>((n * (ld / 8) + 255) / 256, 148 * 8));
    rows_to_bf16_lo_kernel<<<blocks, 256, 0, stream>>>(src, n, T, D, s_seq, s_tok, s_col, static_cast<__nv_bfloat16*>(dst), ld);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// fp32 multi-head self-attention core (multihead_self.py:15-23) on fp32 Q|K|V rows [n_seq*T][ld] (Q at column 0, K at d, V at
// 2d): one CTA per (sequence, head), thread i owns query row i (T <= 64): scores, exp-softmax with the +1e-8, P.V -- all in
// fp32 registers / shared memory.  The context leaves as bf16 hi + lo planes (pitch ldc, ones column at d in the hi plane).
constexpr int kF32MaxT = 64, kF32MaxDk = 32;
__global__ void __launch_bounds__(64) mhsa_f32_fwd_kernel(const float* __restrict__ qkv, int ld, int sec, int T, int heads, int dk,
                                                          __nv_bfloat16* __restrict__ c_hi, __nv_bfloat16* __restrict__ c_lo, int ldc) {
    __shared__ float sk[kF32MaxT][kF32MaxDk + 1], sv[kF32MaxT][kF32MaxDk + 1];
    const int seq = blockIdx.x / heads, h = blockIdx.x - seq * heads;
    const int d = heads * dk;
    const float* base = qkv + static_cast<size_t>(seq) * T * ld + h * dk;
    for (int i = threadIdx.x; i < T * dk; i += blockDim.x) {
        const int r = i / dk, c = i - r * dk;
        sk[r][c] = base[static_cast<size_t>(r) * ld + sec + c];
        sv[r][c] = base[static_cast<size_t>(r) * ld + 2 * sec + c];
    }
    __syncthreads();
    const int i = threadIdx.x;
    if (i >= T) return;
    float q[kF32MaxDk];
#pragma unroll
    for (int c = 0; c < kF32MaxDk; ++c) q[c] = c < dk ? base[static_cast<size_t>(i) * ld + c] : 0.f;
    const float rs = rsqrtf(static_cast<float>(dk));
    float s[kF32MaxT];
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < kF32MaxDk; ++c) acc = fmaf(q[c], c < dk ? sk[j][c] : 0.f, acc);
        s[j] = acc * rs;
        m = fmaxf(m, s[j]);
    }
    float l = 0.f;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        s[j] = __expf(s[j] - m);
        l += s[j];
    }
    const float inv = 1.f / (l + 1e-8f * __expf(-m));  // == exp(S) / (sum exp(S) + 1e-8)
    float o[kF32MaxDk];
#pragma unroll
    for (int c = 0; c < kF32MaxDk; ++c) o[c] = 0.f;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        const float p = s[j] * inv;
#pragma unroll
        for (int c = 0; c < kF32MaxDk; ++c) o[c] = fmaf(p, c < dk ? sv[j][c] : 0.f, o[c]);
    }
    const size_t row = (static_cast<size_t>(seq) * T + i) * ldc;
#pragma unroll
    for (int c = 0; c < kF32MaxDk; ++c) {
        if (c < dk) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(o[c]);
            c_hi[row + h * dk + c] = hi;
            c_lo[row + h * dk + c] = __float2bfloat16_rn(o[c] - __bfloat162float(hi));
        }
    }
    if (h == 0) {
        for (int c = d; c < ldc; ++c) {
            c_hi[row + c] = __float2bfloat16_rn(c == d ? 1.0f : 0.f);
            c_lo[row + c] = __float2bfloat16_rn(0.f);
        }
    }
}
// The same for d_k % 4 == 0 (the reference's 20): keys / values are read from shared memory as float4 (the scalar form above
// issues one ld.shared per FMA and is bound by it: 0.45 ms for the 512 x 15 history-level heads), the softmax runs online
// (one pass, no per-thread score array in local memory), exp(S)/(sum exp(S) + 1e-8) in its max-subtracted form.
template <int DK>
__global__ void __launch_bounds__(64) mhsa_f32_fwd_v4_kernel(const float* __restrict__ qkv, int ld, int sec, int T, int heads,
                                                             __nv_bfloat16* __restrict__ c_hi, __nv_bfloat16* __restrict__ c_lo, int ldc) {
    constexpr int V4 = DK / 4;
    __shared__ float4 sk[kF32MaxT][V4], sv[kF32MaxT][V4];
    const int seq = blockIdx.x / heads, h = blockIdx.x - seq * heads;
    const int d = heads * DK;
    const float* base = qkv + static_cast<size_t>(seq) * T * ld + h * DK;
    for (int i = threadIdx.x; i < T * V4; i += blockDim.x) {
        const int r = i / V4, c = i - r * V4;
        const float* kr = base + static_cast<size_t>(r) * ld + sec + 4 * c;
        const float* vr = base + static_cast<size_t>(r) * ld + 2 * sec + 4 * c;
        sk[r][c] = *reinterpret_cast<const float4*>(kr);  // 16-byte aligned: ld, sec multiples of 4, DK % 4 == 0
        sv[r][c] = *reinterpret_cast<const float4*>(vr);
    }
    __syncthreads();
    const int i = threadIdx.x;
    if (i >= T) return;
    float4 q[V4], o[V4];
    const float sc = rsqrtf(static_cast<float>(DK)) * 1.4426950408889634f;  // scores in the log2 domain
#pragma unroll
    for (int c = 0; c < V4; ++c) {
        const float4 qr = *reinterpret_cast<const float4*>(base + static_cast<size_t>(i) * ld + 4 * c);
        q[c] = make_float4(qr.x * sc, qr.y * sc, qr.z * sc, qr.w * sc);
        o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float m = -INFINITY, l = 0.f;
#pragma unroll 2
    for (int j = 0; j < T; ++j) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < V4; ++c) {
            const float4 k = sk[j][c];
            a0 = fmaf(q[c].x, k.x, a0); a1 = fmaf(q[c].y, k.y, a1);
            a0 = fmaf(q[c].z, k.z, a0); a1 = fmaf(q[c].w, k.w, a1);
        }
        const float sj = a0 + a1;
        const float mn = fmaxf(m, sj);
        const float r = exp2f(m - mn), pj = exp2f(sj - mn);  // first key: m = -inf -> r = 0
        l = fmaf(l, r, pj);
        m = mn;
#pragma unroll
        for (int c = 0; c < V4; ++c) {
            const float4 v = sv[j][c];
            o[c].x = fmaf(o[c].x, r, pj * v.x); o[c].y = fmaf(o[c].y, r, pj * v.y);
            o[c].z = fmaf(o[c].z, r, pj * v.z); o[c].w = fmaf(o[c].w, r, pj * v.w);
        }
    }
    const float inv = 1.f / (l + 1e-8f * exp2f(-m));  // == exp(S) / (sum exp(S) + 1e-8)
    const size_t row = (static_cast<size_t>(seq) * T + i) * ldc;
#pragma unroll
    for (int c = 0; c < V4; ++c) {
        const float ov[4] = {o[c].x * inv, o[c].y * inv, o[c].z * inv, o[c].w * inv};
        uint32_t hw[2], lw[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            hw[e] = pack_bf16x2(ov[2 * e], ov[2 * e + 1]);
            const float2 hf = unpack_bf16x2(hw[e]);
            lw[e] = pack_bf16x2(ov[2 * e] - hf.x, ov[2 * e + 1] - hf.y);
        }
        *reinterpret_cast<uint2*>(c_hi + row + h * DK + 4 * c) = make_uint2(hw[0], hw[1]);  // 8-byte aligned: ldc % 8 == 0, DK % 4 == 0
        *reinterpret_cast<uint2*>(c_lo + row + h * DK + 4 * c) = make_uint2(lw[0], lw[1]);
    }
    if (h == 0) {
        for (int c = d; c < ldc; ++c) {
            c_hi[row + c] = __float2bfloat16_rn(c == d ? 1.0f : 0.f);
            c_lo[row + c] = __float2bfloat16_rn(0.f);
        }
    }
}
int mhsa_f32_fwd(const float* qkv, int ld, int sec, long long n_seq, int T, int heads, int dk, void* c_hi, void* c_lo, int ldc,
                 cudaStream_t stream) {
    if (n_seq == 0) return 0;
    NR_REQUIRE(T >= 1 && T <= kF32MaxT && dk >= 1 && dk <= kF32MaxDk && ldc >= heads * dk + 1 && n_seq * heads < (1ll << 31),
               "mhsa_f32_fwd: T=%d dk=%d ldc=%d", T, dk, ldc);
    ProfScope ps("mhsa_f32_fwd", static_cast<int>(n_seq), T, heads * dk, stream);
    if (dk == 20 && ldc % 8 == 0 && ld % 4 == 0 && sec % 4 == 0) {
        mhsa_f32_fwd_v4_kernel<20><<<static_cast<int>(n_seq * heads), 64, 0, stream>>>(qkv, ld, sec, T, heads, static_cast<__nv_bfloat16*>(c_hi),
                                                                                       static_cast<__nv_bfloat16*>(c_lo), ldc);
        ++g_launches;
        NR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    mhsa_f32_fwd_kernel<<<static_cast<int>(n_seq * heads), 64, 0, stream>>>(qkv, ld, sec, T, heads, dk, static_cast<__nv_bfloat16*>(c_hi),
                                                                           static_cast<__nv_bfloat16*>(c_lo), ldc);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// evaluation stage 3, batched (reference src/evaluate.py:245-265 scores ONE impression per get_prediction call and
// synchronises on .tolist() after each): scores[i] = news[cand[i]] . user[seg(i)] for the candidates of MANY impressions
// in one launch.  The news vectors stay in ONE device matrix (row = news index) instead of a Python dict of rows;
// seg_offsets[s] .. seg_offsets[s+1] delimit the candidates of impression s.  One warp per candidate, 16-byte loads.
// ------------------------------------------------------------------------------------------------
__global__ void segment_dot_kernel(const float* __restrict__ news, long long n_news, int D, const long long* __restrict__ cand,
                                   long long n_cand, const long long* __restrict__ seg_offsets, long long n_seg,
                                   const float* __restrict__ user, float* __restrict__ scores, int* __restrict__ bad_flag) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long i = w0; i < n_cand; i += nw) {
        long long lo = 0, hi = n_seg;  // the impression of candidate i: last s with seg_offsets[s] <= i
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(seg_offsets + mid) <= i) lo = mid; else hi = mid;
        }
        long long nid = __ldg(cand + i);
        if (nid < 0 || nid >= n_news) {
            if (lane == 0) atomicExch(bad_flag, 1);
            nid = 0;
        }
        const float* nv = news + nid * D;
        const float* uv = user + lo * D;
        float acc = 0.f;
        if ((D & 3) == 0) {
            for (int c = lane * 4; c < D; c += 128) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(nv + c)), b = __ldg(reinterpret_cast<const float4*>(uv + c));
                acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
            }
        } else {
            for (int c = lane; c < D; c += 32) acc = fmaf(nv[c], uv[c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) scores[i] = acc;
    }
}
int segment_dot(const float* news, long long n_news, int D, const long long* cand, long long n_cand, const long long* seg_offsets,
                long long n_seg, const float* user, float* scores, int* bad_flag, cudaStream_t stream) {
    if (n_cand == 0) return 0;
    ProfScope ps("segment_dot", static_cast<int>(n_cand), static_cast<int>(n_seg), D, stream);
    const int blocks = static_cast<int>(std::min<long long>((n_cand + 7) / 8, 148 * 16));
    segment_dot_kernel<<<blocks, 256, 0, stream>>>(news, n_news, D, cand, n_cand, seg_offsets, n_seg, user, scores, bad_flag);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Batch feed: the reference hands the model slot-major lists of per-slot (B, L) int64 tensors (default_collate over
// src/dataset.py:64-85; pinned by the DataLoader, src/train.py:165-171).  ONE launch reads the payload of every slot
// straight from page-locked host memory (unified addressing: the kernel's loads cross PCIe, no host staging copy, no
// per-slot cudaMemcpyAsync) and writes the impression-major block the encoders consume:
//   out[(b*H + h)*L + t] = clicked[h][b][t]            rows [0, B*H)
//   out[B*H*L + (b*C + c)*L + t] = candidates[c][b][t]  rows [B*H, B*(H+C))
// ------------------------------------------------------------------------------------------------
constexpr int kSlotTable = 64;
struct SlotTable {
    const long long* p[kSlotTable];
};
__global__ void __launch_bounds__(256) pack_slots_kernel(SlotTable tab, int n, int slot0, int H, int C, int B, int L, long long* __restrict__ out) {
    const long long per = static_cast<long long>(B) * L, total = per * n;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int sl = static_cast<int>(i / per);
        const long long rem = i - sl * per;
        const int b = static_cast<int>(rem / L), t = static_cast<int>(rem - static_cast<long long>(b) * L);
        const int s = slot0 + sl;
        const long long v = tab.p[sl][rem];
        const long long dst = s < H ? (static_cast<long long>(b) * H + s) * L + t
                                    : static_cast<long long>(B) * H * L + (static_cast<long long>(b) * C + (s - H)) * L + t;
        out[dst] = v;
    }
}
// 1 when every pointer is readable by a kernel on the current device (device / managed memory, or page-locked host memory
// whose device alias is the same address)
int slots_device_readable(const void* const* slots, int n) {
    for (int i = 0; i < n; ++i) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, slots[i]) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) continue;
        if (at.type != cudaMemoryTypeHost || at.devicePointer != slots[i]) return 0;
    }
    return 1;
}
int pack_slots(const void* const* slots, int H, int C, int B, int L, long long* out, cudaStream_t stream) {
    const int n_slots = H + C;
    if (n_slots == 0 || B == 0 || L == 0) return 0;
    ProfScope ps("pack_slots", n_slots, B, L, stream);
    for (int s0 = 0; s0 < n_slots; s0 += kSlotTable) {
        SlotTable tab;
        const int n = std::min(kSlotTable, n_slots - s0);
        for (int i = 0; i < kSlotTable; ++i) tab.p[i] = static_cast<const long long*>(slots[s0 + std::min(i, n - 1)]);
        const long long total = static_cast<long long>(n) * B * L;
        const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
        pack_slots_kernel<<<blocks, 256, 0, stream>>>(tab, n, s0, H, C, B, L, out);
        ++g_launches;
        NR_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace nr
