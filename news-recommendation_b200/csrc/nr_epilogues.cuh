// Fused epilogue functors for gemm_nt.  256 epilogue threads: thread (quarter, lane) owns accumulator row
// r = 32*quarter + lane (= TMEM lane) and the two warps of a quarter ("halves") split the row's 32-column chunks
// [ch0, ch1).  Contract for every functor:
//   * the accumulator is read through epi_chunks() (nr_gemm.cuh): one tcgen05.ld per chunk of the thread's range
//     (warp-collective, the range is warp-uniform), software pipelined, acc.release() exactly once per tile
//   * init()/finish() bracket the CTA's whole tile loop (all 256 epilogue threads call them)
//   * kScratchBytes of shared memory belong to the functor (the planner sizes the A ring around it)
//   * per-slice vectors (bias, query vector, dOut rows) are staged in shared memory: with ~220 KB of smem carved out
//     the L1 holds next to nothing and per-element global loads made the epilogue 10x the MMA time (ncu, profiles/).
//     Anything that must come from global memory per tile is fetched with many loads in flight or prefetched one
//     tile ahead with cp.async -- a dependent L2 round trip (~600 cycles) per chunk was the whole epilogue time.
#pragma once
#include "nr_gemm.cuh"

namespace nr {

// Row re-mapping between the compact token layout (segment s, token t -> row s*T + t) and the
// zero-padded CNN layout (row s*(T+2) + 1 + t; rows 0 and T+1 of every segment are zero).
struct RowMap {
    int seg_in;   // 0 => identity
    int in_off;
    int seg_len;
    int seg_out;
    int out_off;
    __device__ __forceinline__ bool map(int grow, long long& orow, int& t_out) const {
        if (seg_in == 0) { orow = grow; t_out = 0; return true; }
        const int s = grow / seg_in;
        const int t = grow - s * seg_in - in_off;
        t_out = t;
        orow = static_cast<long long>(s) * seg_out + t + out_off;
        return t >= 0 && t < seg_len;
    }
};

__device__ __forceinline__ void zero_row_bf16(__nv_bfloat16* row, int ld) {  // ld % 8 == 0, 16B aligned
    for (int i = 0; i < ld; i += 8) *reinterpret_cast<uint4*>(row + i) = make_uint4(0u, 0u, 0u, 0u);
}

struct Dropout {
    float p;          // 0 => off
    float scale;      // 1/(1-p)
    uint32_t thresh;  // keep iff 16-bit lane >= thresh
    uint64_t seed;
    // multiplier for element (row, col) of a matrix with pitch ld; cols are visited in aligned groups of 4
    __device__ __forceinline__ void mask4(long long row, int ld, int col4, float* m) const {
        mask4_group((static_cast<uint64_t>(row) * ld + col4) >> 2, m);
    }
    // the same for a precomputed group index ((row * ld + col) >> 2; callers that walk a row keep row * ld / 4 in a register);
    // 32-bit field tests: the 64-bit shifts / compares of the straightforward form were a quarter of the instructions of the
    // dropout-carrying epilogues (ncu source page, profiles/)
    __device__ __forceinline__ void mask4_group(uint64_t group, float* m) const {
        const uint64_t bits = dropout_bits4(seed, group);
        const uint32_t lo = static_cast<uint32_t>(bits), hi = static_cast<uint32_t>(bits >> 32);
        m[0] = ((lo & 0xffffu) >= thresh) ? scale : 0.f;
        m[1] = ((lo >> 16) >= thresh) ? scale : 0.f;
        m[2] = ((hi & 0xffffu) >= thresh) ? scale : 0.f;
        m[3] = ((hi >> 16) >= thresh) ? scale : 0.f;
    }
};

// 16 bf16 = one full 32-byte sector per thread (STG.256, sm_100): halves the L2 write requests of the row-per-thread
// epilogues compared with two half-sector 16-byte stores.  Needs a 32-byte aligned destination.
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* o, const float* y) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(pack_bf16x2(y[0], y[1])),
                 "r"(pack_bf16x2(y[2], y[3])), "r"(pack_bf16x2(y[4], y[5])), "r"(pack_bf16x2(y[6], y[7])),
                 "r"(pack_bf16x2(y[8], y[9])), "r"(pack_bf16x2(y[10], y[11])), "r"(pack_bf16x2(y[12], y[13])),
                 "r"(pack_bf16x2(y[14], y[15]))
                 : "memory");
}
__device__ __forceinline__ bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* o, const float* y, int nvalid) {
    if (nvalid == 8) {
        *reinterpret_cast<uint4*>(o) =
            make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    } else {  // fully unrolled + predicated: a run-time index into y would push the caller's accumulator registers to local memory
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nvalid) o[j] = __float2bfloat16_rn(y[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// out = act(acc + bias) [* dropout]  ->  bf16 or fp32, optional row re-map, optional "ones" column
// scratch floats: [0,256) bias of the slice | then the WarpTileStore staging region
// ------------------------------------------------------------------------------------------------
struct EpiStore {
    static constexpr int kScratchBytes = 1024 + kTileStoreBytes;
    CUtensorMap tm_out;  // bf16 output as a TMA tensor (32 x 32 boxes, SWIZZLE_64B); valid when use_tma
    CUtensorMap tm_lo;   // low plane of the columns >= lo_col0 (bf16(y - bf16(y))), its column 0 = output column lo_col0
    int accumulate;      // fp32 output only: out += result
    int lo_col0;         // < 0: no low plane
    __nv_bfloat16* lo_out;  // the same plane for the (rare) partial chunks that leave through plain stores; pitch ld_lo
    int ld_lo;
    int use_tma;         // identity row map + bf16 output: full 32-column chunks leave through WarpTileStore
    void* out;
    int ld;
    int out_bf16;
    const float* bias;  // may be null
    int relu;
    int N;              // total valid columns
    RowMap rm;
    int zero_pad_rows;  // compact->padded: the first/last token also zero the neighbouring pad row
    Dropout drop;
    int ones_col;       // >=0: column set to 1.0 (bias-gradient trick for the next weight-grad GEMM); -1 off
    int ones_cols_zero_upto;  // columns (ones_col, upto) are zeroed
    int dbg_skip;       // tuning only (NEWSREC_EPI_DBG=1): release the accumulator untouched -> MMA/TMA pipeline alone

    __device__ void init(const EpiInit& e, int) const {
        for (int i = e.tid; i < 256; i += kEpiThreads) e.scratch[i] = (bias != nullptr && i < e.ncols) ? bias[e.col0 + i] : 0.f;
        epi_bar_sync();
    }
    __device__ void finish(const EpiInit& e) const {
        if (use_tma) WarpTileStore::drain(e.tid & 31);
    }

    template <class Acc>
    __device__ void operator()(const Acc& acc, const EpiCtx& c) const {
        long long orow;
        int t;
        const bool v = rm.map(c.grow, orow, t) && c.valid;
#ifdef NEWSREC_TRIAGE
        if (dbg_skip) {
            acc.release();
            return;
        }
#endif
        const int lane = c.tid & 31;
        WarpTileStore ts;
        ts.attach(c.scratch + 256, c.tid >> 5);
        if (use_tma) ts.begin_tile(lane);
        float b[32];
        epi_chunks(
            acc, c,
            [&](int ch) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = lds_f4(c.scratch + ch * 32 + j);
                    b[j] = b4.x; b[j + 1] = b4.y; b[j + 2] = b4.z; b[j + 3] = b4.w;
                }
            },
            [&](int ch, float* x) {
                const int lc0 = ch * 32;
                const bool whole = use_tma && lc0 + 32 <= c.ncols;  // warp-uniform
                // row-mapped bf16 output (conv forward: padded rows -> compact rows): full chunks leave through the staging tile
                // as coalesced stores (WarpTileStore::put_rows) instead of one 16/32-byte store per lane and row
                const bool coop = !use_tma && out_bf16 && lc0 + 32 <= c.ncols && c.col0 + lc0 + 32 <= N && (ld & 7) == 0 &&
                                  ((c.col0 + lc0) & 7) == 0;  // warp-uniform
                if (!whole && !coop && !v) return;
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] += b[j];
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (drop.p > 0.f) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float m[4];
                        drop.mask4(orow, ld, c.col0 + lc0 + j, m);
                        x[j] *= m[0]; x[j + 1] *= m[1]; x[j + 2] *= m[2]; x[j + 3] *= m[3];
                    }
                }
                if (whole) {  // columns >= N and rows >= M are clipped by the tensor map
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
                    ts.put(&tm_out, w, c.col0 + lc0, c.grow - lane, lane);
                    if (lo_col0 >= 0 && c.col0 + lc0 >= lo_col0) {  // warp-uniform (the host checked the chunk alignment)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 h = unpack_bf16x2(w[j]);
                            w[j] = pack_bf16x2(x[2 * j] - h.x, x[2 * j + 1] - h.y);
                        }
                        ts.put(&tm_lo, w, c.col0 + lc0 - lo_col0, c.grow - lane, lane);
                    }
                    return;
                }
                if (coop) {
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
                    ts.put_rows(static_cast<__nv_bfloat16*>(out), ld, w, c.col0 + lc0, v ? static_cast<int>(orow) : -1, lane);
                    if (lo_col0 >= 0 && c.col0 + lc0 >= lo_col0 && (ld_lo & 7) == 0) {  // low plane, same (mapped) rows
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 h = unpack_bf16x2(w[j]);
                            w[j] = pack_bf16x2(x[2 * j] - h.x, x[2 * j + 1] - h.y);
                        }
                        ts.put_rows(lo_out, ld_lo, w, c.col0 + lc0 - lo_col0, v ? static_cast<int>(orow) : -1, lane);
                    }
                    return;
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int lc = lc0 + g * 16;  // column inside the slice
                    if (lc >= c.ncols) break;
                    const int col = c.col0 + lc;
                    const float* y = x + g * 16;
                    const int nvalid = min(16, min(c.ncols - lc, N - col));
                    if (out_bf16) {
                        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out) + orow * ld + col;
                        if (nvalid == 16 && aligned32(o)) {
                            store_bf16x16(o, y);
                        } else {
                            store_bf16x8(o, y, min(nvalid, 8));
                            if (nvalid > 8) store_bf16x8(o + 8, y + 8, nvalid - 8);
                        }
                        if (lo_col0 >= 0 && col >= lo_col0) {  // low plane of a partial chunk (identity rows)
                            float yl[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) yl[j] = y[j] - bf16_round(y[j]);
                            __nv_bfloat16* ol = lo_out + orow * ld_lo + (col - lo_col0);
                            store_bf16x8(ol, yl, min(nvalid, 8));
                            if (nvalid > 8) store_bf16x8(ol + 8, yl + 8, nvalid - 8);
                        }
                    } else {
                        float* o = static_cast<float*>(out) + orow * ld + col;
                        if (nvalid == 16) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                float4 v4 = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                                if (accumulate) {  // out += : second pass of a split-operand product (x_lo . W^T on top of x_hi . W^T)
                                    const float4 o4 = *reinterpret_cast<const float4*>(o + j);
                                    v4.x += o4.x; v4.y += o4.y; v4.z += o4.z; v4.w += o4.w;
                                }
                                *reinterpret_cast<float4*>(o + j) = v4;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (j < nvalid) o[j] = accumulate ? o[j] + y[j] : y[j];
                        }
                    }
                }
            });
        if (v && c.col0 == 0 && c.half == 0) {
            if (ones_col >= 0 && out_bf16) {
                __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out) + orow * ld;
                o[ones_col] = __float2bfloat16_rn(1.0f);
                for (int j = ones_col + 1; j < ones_cols_zero_upto; ++j) o[j] = __float2bfloat16_rn(0.f);
            }
            if (zero_pad_rows && out_bf16) {  // neighbours of the first / last token of a segment are pad rows
                __nv_bfloat16* base = static_cast<__nv_bfloat16*>(out);
                if (t == 0) zero_row_bf16(base + (orow - 1) * ld, ld);
                if (t == rm.seg_len - 1) zero_row_bf16(base + (orow + 1) * ld, ld);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Additive-attention pooling (reference additive.py:35-53) fused behind  pre = X.Wa^T:
//   score_r = sum_c tanh(pre_rc + ba_c) * qv_c ; w = softmax over the segment ; out_s = sum_r w_r X_r
// Requires n_slices == 1 and rows_per_tile = (segments per tile) * seg_len.
// scratch floats: [0,384) partial scores (part*128 + row) | [384,512) weights | [512,768) bias | [768,1024) query
// ------------------------------------------------------------------------------------------------
struct EpiPool {
    static constexpr int kScratchBytes = 4096;
    const float* bias;
    const float* qv;
    const __nv_bfloat16* X;  // the GEMM's A operand (rows x lda), re-read (L2 hits) for the weighted sum
    const __nv_bfloat16* X_lo;  // optional second plane (same pitch): the pooled rows are X + X_lo (hi/lo bf16 pair)
    int lda;
    int D;        // pooled width (even)
    int seg_len;
    int rows_per_tile;
    int M;
    float* out;   // [segments][ldo] fp32
    int ldo;
    float* w_out; // [rows] fp32 softmax weights (saved for backward); may be null

    __device__ void init(const EpiInit& e, int) const {
        for (int i = e.tid; i < 256; i += kEpiThreads) {
            e.scratch[512 + i] = i < e.ncols ? bias[e.col0 + i] : 0.f;
            e.scratch[768 + i] = i < e.ncols ? qv[e.col0 + i] : 0.f;
        }
        epi_bar_sync();
    }
    __device__ void finish(const EpiInit&) const {}

    // 16-byte column chunk ck of the rows [r0, r0 + seg_len): weighted sum with the weights in s_w, U loads in flight
    template <int U>
    __device__ __forceinline__ void wsum_rows(const uint4* xp, size_t pitch16, const float* s_w, int& t, float* a) const {
        for (; t + U <= seg_len; t += U) {
            uint4 u[U];
#pragma unroll
            for (int k = 0; k < U; ++k) u[k] = __ldg(xp + static_cast<size_t>(t + k) * pitch16);
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float wt = lds_f(s_w + t + k);
                const uint32_t uw[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(uw[j]);
                    a[2 * j] = fmaf(wt, f.x, a[2 * j]);
                    a[2 * j + 1] = fmaf(wt, f.y, a[2 * j + 1]);
                }
            }
        }
    }

    template <class Acc>
    __device__ void operator()(const Acc& acc, const EpiCtx& c) const {
        float score = 0.f;
        epi_chunks(
            acc, c, [](int) {},
            [&](int ch, float* x) {
                const float* sb = c.scratch + 512 + ch * 32;
                const float* sq = c.scratch + 768 + ch * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = lds_f4(sb + j);
                    const float4 q4 = lds_f4(sq + j);  // zero beyond ncols: no contribution
                    score = fmaf(fast_tanh(x[j] + b4.x), q4.x, score);
                    score = fmaf(fast_tanh(x[j + 1] + b4.y), q4.y, score);
                    score = fmaf(fast_tanh(x[j + 2] + b4.z), q4.z, score);
                    score = fmaf(fast_tanh(x[j + 3] + b4.w), q4.w, score);
                }
            });
        float* s_part = c.scratch;
        float* s_w = c.scratch + 384;
        static_assert(kEpiParts <= 3, "EpiPool scratch layout");
        auto row_score = [&](int r) {
            float t = s_part[r];
#pragma unroll
            for (int pp = 1; pp < kEpiParts; ++pp) t += s_part[pp * 128 + r];
            return t;
        };
        s_part[c.half * 128 + c.r] = score;
        epi_bar_sync();
        if (c.half == 0) {
            float w = 0.f;
            if (c.valid) {
                const int s0 = (c.r / seg_len) * seg_len;
                const float mine = row_score(c.r);
                float m = -INFINITY;
                for (int t = 0; t < seg_len; ++t) m = fmaxf(m, row_score(s0 + t));
                float sum = 0.f;
                for (int t = 0; t < seg_len; ++t) sum += __expf(row_score(s0 + t) - m);
                w = __fdividef(__expf(mine - m), sum);
                if (w_out != nullptr) w_out[c.grow] = w;
            }
            s_w[c.r] = w;
        }
        epi_bar_sync();
        // out[segment] = sum_t w_t X_t: work item = (segment, 16-byte column chunk); the rows come back from L2 and
        // the loop keeps 10 (then 4, then 1) independent loads in flight per thread.
        const int row0 = c.tile * rows_per_tile;
        const int nseg = rows_per_tile / seg_len;
        const int nck = (D + 7) >> 3;  // the A pitch is a multiple of 8 elements: the last chunk stays in bounds
        const size_t pitch16 = static_cast<size_t>(lda) >> 3;
        for (int item = c.tid; item < nseg * nck; item += kEpiThreads) {
            const int s = item / nck, ck = item - s * nck;
            const int r0 = row0 + s * seg_len;
            if (r0 >= M) continue;
            const uint4* xp = reinterpret_cast<const uint4*>(X + static_cast<size_t>(r0) * lda) + ck;
            const float* sw = s_w + s * seg_len;
            float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int t = 0;
            wsum_rows<10>(xp, pitch16, sw, t, a);
            wsum_rows<4>(xp, pitch16, sw, t, a);
            wsum_rows<1>(xp, pitch16, sw, t, a);
            if (X_lo != nullptr) {  // low plane of a hi/lo pair: same weights, same accumulators
                const uint4* xl = reinterpret_cast<const uint4*>(X_lo + static_cast<size_t>(r0) * lda) + ck;
                t = 0;
                wsum_rows<10>(xl, pitch16, sw, t, a);
                wsum_rows<4>(xl, pitch16, sw, t, a);
                wsum_rows<1>(xl, pitch16, sw, t, a);
            }
            float* o = out + static_cast<size_t>(r0 / seg_len) * ldo + ck * 8;
            if (ck * 8 + 8 <= D && (ldo & 3) == 0) {
                *reinterpret_cast<float4*>(o) = make_float4(a[0], a[1], a[2], a[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(a[4], a[5], a[6], a[7]);
            } else {
                for (int j = 0; j < 8 && ck * 8 + j < D; ++j) o[j] = a[j];
            }
        }
        epi_bar_sync();
    }
};

// ------------------------------------------------------------------------------------------------
// Backward of the additive scorer, fused behind the recomputed pre = X.Wa^T:
//   T = tanh(pre + ba);  dPre_rc = dscore_r * qv_c * (1 - T^2)  -> bf16;   dqv_c += sum_r dscore_r * T_rc
// scratch floats: [0,256) column sums | [256,512) bias | [512,768) query vector
// ------------------------------------------------------------------------------------------------
struct EpiDPre {
    static constexpr int kScratchBytes = 3072 + kTileStoreBytes;
    CUtensorMap tm_out;       // dpre as a TMA tensor (32 x 32 boxes); valid when use_tma
    int use_tma;
    const float* bias;
    const float* qv;
    const float* dscore;      // [rows]
    __nv_bfloat16* dpre;      // [rows][ld]
    int ld;
    float* dqv;               // [q] fp32, accumulated

    __device__ void init(const EpiInit& e, int) const {
        for (int i = e.tid; i < 256; i += kEpiThreads) {
            e.scratch[i] = 0.f;
            e.scratch[256 + i] = i < e.ncols ? bias[e.col0 + i] : 0.f;
            e.scratch[512 + i] = i < e.ncols ? qv[e.col0 + i] : 0.f;
        }
        epi_bar_sync();
    }
    __device__ void finish(const EpiInit& e) const {
        epi_bar_sync();
        for (int i = e.tid; i < e.ncols; i += kEpiThreads) atomicAdd(dqv + e.col0 + i, e.scratch[i]);
        if (use_tma) WarpTileStore::drain(e.tid & 31);
    }
    template <class Acc>
    __device__ void operator()(const Acc& acc, const EpiCtx& c) const {
        const float ds = c.valid ? __ldg(dscore + c.grow) : 0.f;
        const int lane = c.tid & 31;
        WarpTileStore ts;
        if (use_tma) {
            ts.attach(c.scratch + 768, c.tid >> 5);
            ts.begin_tile(lane);
        }
        epi_chunks(
            acc, c, [](int) {},
            [&](int ch, float* x) {
                // bias and query vector are zero past ncols in shared memory: dp = 0 there without a column test, and
                // the column sums of those columns are never added to dqv
                float dp[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = lds_f4(c.scratch + 256 + ch * 32 + j);
                    const float4 q4 = lds_f4(c.scratch + 512 + ch * 32 + j);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, qq[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float tt = tanh_approx(x[j + i] + bb[i]);
                        dp[j + i] = (ds * qq[i]) * fmaf(-tt, tt, 1.f);
                        x[j + i] = ds * tt;
                    }
                }
                if (use_tma && ch * 32 + 32 <= c.ncols) {  // invalid rows carry ds = 0 and are clipped at M anyway
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(dp[2 * j], dp[2 * j + 1]);
                    ts.put(&tm_out, w, c.col0 + ch * 32, c.grow - lane, lane);
                } else if (c.valid) {
                    __nv_bfloat16* o = dpre + static_cast<size_t>(c.grow) * ld + c.col0 + ch * 32;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int lc = ch * 32 + g * 16;
                        if (lc >= c.ncols) break;  // ld is padded to a multiple of 8: whole 8-groups are in bounds
                        if (lc + 16 <= ld - c.col0 && aligned32(o + g * 16)) {
                            store_bf16x16(o + g * 16, dp + g * 16);
                        } else {
                            store_bf16x8(o + g * 16, dp + g * 16, 8);
                            if (lc + 8 < c.ncols) store_bf16x8(o + g * 16 + 8, dp + g * 16 + 8, 8);
                        }
                    }
                }
                const float colsum = warp_transpose_sum32(x);
                if (ch * 32 + lane < c.ncols) atomicAdd(c.scratch + ch * 32 + lane, colsum);
            });
    }
};

// ------------------------------------------------------------------------------------------------
// dX_rc = acc_rc + w_r * dOut[seg(r)][c]  (pool backward, both paths into X) [* relu mask] [* dropout] -> bf16
// scratch floats: two staging buffers of kStageFloats: the dOut rows (slice columns only) of every segment the tile
// touches, [segment][pitch = slice width]; the buffer of the NEXT tile is filled by cp.async while this one is used.
// ------------------------------------------------------------------------------------------------
struct EpiDPoolIn {
    static constexpr int kStageFloats = 2560;
    static constexpr int kScratchBytes = 2 * kStageFloats * 4 + kTileStoreBytes;
    CUtensorMap tm_out;  // dx as a TMA tensor (32 x 32 boxes); valid when use_tma (identity row map, no ReLU mask)
    int use_tma;
    const float* w;      // [rows]
    const float* dout;   // [segments][ldo]
    int ldo;
    int seg_len;
    __nv_bfloat16* dx;   // [rows(mapped)][ld]
    int ld;
    int N;
    RowMap rm;
    int zero_pad_rows;
    Dropout drop;
    const __nv_bfloat16* relu_src;  // non-null: multiply by (relu_src[r][c] > 0) (ReLU backward of the CNN); pitch relu_ld
    int relu_ld;
    int M;
    int rows_per_tile;

    // all 256 threads: queue the dOut rows of `tile` (columns [col0, col0 + ncols)) into buf
    __device__ __forceinline__ void stage_tile(int tile, int col0, int ncols, int tid, float* buf) const {
        const int row0 = tile * rows_per_tile;
        const int seg_first = row0 / seg_len;
        const int seg_last = min(row0 + 127, M - 1) / seg_len;
        const int total = (seg_last - seg_first + 1) * ncols;
        for (int i = tid; i < total; i += kEpiThreads) {
            const int sgi = i / ncols, j = i - sgi * ncols;
            const bool ok = col0 + j < N;
            cp_async_f32(buf + i, dout + static_cast<size_t>(seg_first + sgi) * ldo + (ok ? col0 + j : 0), ok);
        }
        cp_async_commit();
    }
    __device__ void init(const EpiInit& e, int) const {
        if (e.first_tile < e.num_tiles) stage_tile(e.first_tile, e.col0, e.ncols, e.tid, e.scratch);
    }
    __device__ void finish(const EpiInit& e) const {
        if (use_tma) WarpTileStore::drain(e.tid & 31);
    }

    template <class Acc>
    __device__ void operator()(const Acc& acc, const EpiCtx& c) const {
        long long orow;
        int t;
        const bool v = rm.map(c.grow, orow, t) && c.valid;
        const float wr = c.valid ? __ldg(w + c.grow) : 0.f;
        const int seg_first = (c.tile * rows_per_tile) / seg_len;
        const int myseg = (c.valid ? c.grow / seg_len : seg_first) - seg_first;
        float* cur = c.scratch + (c.it & 1) * kStageFloats;
        cp_async_wait_all();
        epi_bar_sync();  // this tile's dOut rows are visible; everybody is done with the other buffer
        if (c.next_tile >= 0) stage_tile(c.next_tile, c.col0, c.ncols, c.tid, c.scratch + ((c.it + 1) & 1) * kStageFloats);
        const float* sd = cur + myseg * c.ncols;
        const int lane = c.tid & 31;
        WarpTileStore ts;
        if (use_tma) {
            ts.attach(c.scratch + 2 * kStageFloats, c.tid >> 5);
            ts.begin_tile(lane);
        }
        epi_chunks(
            acc, c, [](int) {},
            [&](int ch, float* x) {
                const bool whole = use_tma && ch * 32 + 32 <= c.ncols && (c.ncols & 3) == 0;  // warp-uniform
                if (whole) {  // identity row map, no ReLU mask; rows >= M / columns >= N are clipped by the tensor map
                    const float* sdc = c.valid ? sd + ch * 32 : cur;  // invalid rows: any staged address (wr = 0)
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 d4 = lds_f4(sdc + j);
                        x[j] = fmaf(wr, d4.x, x[j]); x[j + 1] = fmaf(wr, d4.y, x[j + 1]);
                        x[j + 2] = fmaf(wr, d4.z, x[j + 2]); x[j + 3] = fmaf(wr, d4.w, x[j + 3]);
                    }
                    if (drop.p > 0.f) {
                        const uint64_t g0 = (static_cast<uint64_t>(c.grow) * ld + c.col0 + ch * 32) >> 2;  // ld, col0 are multiples of 4
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float m[4];
                            drop.mask4_group(g0 + (j >> 2), m);
                            x[j] *= m[0]; x[j + 1] *= m[1]; x[j + 2] *= m[2]; x[j + 3] *= m[3];
                        }
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
                    ts.put(&tm_out, pk, c.col0 + ch * 32, c.grow - lane, lane);
                    return;
                }
                // Row-mapped destination and/or ReLU mask (the CNN encoders): a full 32-column chunk goes through the warp's
                // two staging tiles so that the mask rows are LOADED and the result rows STORED as 8 rows x 64 contiguous
                // bytes per instruction (row-per-lane 16-byte accesses touch 32 half-used sectors per instruction and
                // queue in the LSU: this epilogue ran at 0.64 ms against 0.26 ms for the identity/TMA form of the same GEMM).
                const int col = c.col0 + ch * 32;
                const bool coop = !use_tma && ch * 32 + 32 <= c.ncols && col + 32 <= N && (col & 7) == 0 && (ld & 7) == 0 &&
                                  (relu_src == nullptr || (relu_ld & 7) == 0);  // warp-uniform
                if (coop) {
                    uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(c.scratch + 2 * kStageFloats) + 1023) & ~uintptr_t(1023)) +
                                     (c.tid >> 5) * (kTileStoreBufs * 2048);
                    uint8_t* rb = stage;          // mask tile  (32 rows x 64 bytes, SWIZZLE_64B like WarpTileStore)
                    uint8_t* ob = stage + 2048;   // result tile
                    const long long row0 = c.grow - lane;
                    if (relu_src != nullptr) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int r = (lane >> 2) + 8 * k, q = lane & 3;
                            const long long gr = row0 + r;
                            const uint4 u = gr < M ? __ldg(reinterpret_cast<const uint4*>(relu_src + gr * relu_ld + col) + q) : make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(rb + r * 64 + ((q ^ (r >> 1)) & 3) * 16) = u;
                        }
                        __syncwarp();
                    }
                    const float* sdc = c.valid ? sd + ch * 32 : cur;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 d4 = lds_f4(sdc + j);
                        x[j] = fmaf(wr, d4.x, x[j]); x[j + 1] = fmaf(wr, d4.y, x[j + 1]);
                        x[j + 2] = fmaf(wr, d4.z, x[j + 2]); x[j + 3] = fmaf(wr, d4.w, x[j + 3]);
                    }
                    if (relu_src != nullptr) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 ru = *reinterpret_cast<const uint4*>(rb + lane * 64 + ((q ^ (lane >> 1)) & 3) * 16);
                            const uint32_t rw[4] = {ru.x, ru.y, ru.z, ru.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = unpack_bf16x2(rw[j]);
                                // relu_src is the STORED activation dropout(relu(.)): positive <=> passed the ReLU and was kept, so the
                                // keep-multiplier is the constant scale and no counter hash is drawn (bit-identical to mask * relu')
                                x[q * 8 + 2 * j] = f.x > 0.f ? x[q * 8 + 2 * j] * drop.scale : 0.f;
                                x[q * 8 + 2 * j + 1] = f.y > 0.f ? x[q * 8 + 2 * j + 1] * drop.scale : 0.f;
                            }
                        }
                    } else if (drop.p > 0.f) {
                        const uint64_t g0 = (static_cast<uint64_t>(c.grow) * ld + col) >> 2;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float m[4];
                            drop.mask4_group(g0 + (j >> 2), m);
                            x[j] *= m[0]; x[j + 1] *= m[1]; x[j + 2] *= m[2]; x[j + 3] *= m[3];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(ob + lane * 64 + ((q ^ (lane >> 1)) & 3) * 16) =
                            make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                                       pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
                    __syncwarp();
                    const int my_orow = v ? static_cast<int>(orow) : -1;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int r = (lane >> 2) + 8 * k, q = lane & 3;
                        const int orr = __shfl_sync(0xffffffffu, my_orow, r);
                        if (orr >= 0)
                            *(reinterpret_cast<uint4*>(dx + static_cast<long long>(orr) * ld + col) + q) =
                                *reinterpret_cast<const uint4*>(ob + r * 64 + ((q ^ (r >> 1)) & 3) * 16);
                    }
                    __syncwarp();
                    return;
                }
                if (!v) return;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int lc = ch * 32 + g * 8;
                    if (lc >= c.ncols) break;
                    const int col = c.col0 + lc;
                    const int nvalid = min(8, min(c.ncols - lc, N - col));
                    float dd[8];
                    if (((c.ncols & 3) == 0) && lc + 8 <= c.ncols) {
                        const float4 d0 = lds_f4(sd + lc), d1 = lds_f4(sd + lc + 4);
                        dd[0] = d0.x; dd[1] = d0.y; dd[2] = d0.z; dd[3] = d0.w;
                        dd[4] = d1.x; dd[5] = d1.y; dd[6] = d1.z; dd[7] = d1.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) dd[j] = (j < nvalid) ? lds_f(sd + lc + j) : 0.f;
                    }
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = (j < nvalid) ? fmaf(wr, dd[j], x[g * 8 + j]) : 0.f;
                    if (relu_src != nullptr) {
                        const uint4 ru = *reinterpret_cast<const uint4*>(relu_src + static_cast<size_t>(c.grow) * relu_ld + col);
                        const uint32_t rw[4] = {ru.x, ru.y, ru.z, ru.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = unpack_bf16x2(rw[j]);
                            if (!(f.x > 0.f)) y[2 * j] = 0.f;
                            if (!(f.y > 0.f)) y[2 * j + 1] = 0.f;
                        }
                    }
                    if (drop.p > 0.f) {
                        float m[8];
                        drop.mask4(c.grow, relu_src != nullptr ? relu_ld : ld, col, m);
                        drop.mask4(c.grow, relu_src != nullptr ? relu_ld : ld, col + 4, m + 4);
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[j] *= m[j];
                    }
                    store_bf16x8(dx + orow * ld + col, y, nvalid);
                }
            });
        if (c.col0 == 0 && c.half == 0 && zero_pad_rows) {  // warp-uniform; the pad rows next to a segment's first / last token
            if ((ld & 7) == 0) {  // the warp zeroes each pad row together (512 contiguous bytes per instruction)
                const int my_orow = static_cast<int>(orow);
                unsigned first = __ballot_sync(0xffffffffu, v && t == 0), last = __ballot_sync(0xffffffffu, v && t == rm.seg_len - 1);
                for (int pass = 0; pass < 2; ++pass) {
                    unsigned m = pass == 0 ? first : last;
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const long long prow = __shfl_sync(0xffffffffu, my_orow, src) + (pass == 0 ? -1 : 1);
                        uint4* z = reinterpret_cast<uint4*>(dx + prow * ld);
                        for (int i = lane; i < (ld >> 3); i += 32) z[i] = make_uint4(0, 0, 0, 0);
                    }
                }
            } else if (v) {
                if (t == 0) zero_row_bf16(dx + (orow - 1) * ld, ld);
                if (t == rm.seg_len - 1) zero_row_bf16(dx + (orow + 1) * ld, ld);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Embedding gradient: dEmb[ids[r]][c] += acc_rc [* dropout of the gathered row]; row 0 (padding_idx) skipped
// ------------------------------------------------------------------------------------------------
struct EpiScatter {
    static constexpr int kScratchBytes = 16;
    const long long* ids;  // [rows] token ids (row-mapped through rm for the padded CNN layout)
    float* demb;           // [V][D] fp32
    int V;                 // rows of demb: ids outside [1, V) contribute nothing (the forward gather flags them, the reference
                           // raises IndexError; an unchecked id here would be an out-of-bounds atomic into a neighbouring gradient)
    int D;
    RowMap rm;             // maps the GEMM row to the token index in ids
    Dropout drop;
    int drop_ld;           // pitch used when the forward mask was drawn

    __device__ void init(const EpiInit&, int) const {}
    __device__ void finish(const EpiInit&) const {}

    template <class Acc>
    __device__ void operator()(const Acc& acc, const EpiCtx& c) const {
        long long trow;
        int t;
        const bool v = rm.map(c.grow, trow, t) && c.valid;
        long long id = v ? ids[trow] : 0;
        if (id < 0 || id >= V) id = 0;  // out of range: skipped like the padding row
        float* dst = demb + static_cast<size_t>(id) * D;
        epi_chunks(
            acc, c, [](int) {},
            [&](int ch, float* x) {
                if (id == 0) return;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int lc = ch * 32 + g * 4;
                    if (lc >= c.ncols) break;
                    const int col = c.col0 + lc;
                    float y[4] = {x[g * 4], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]};
                    if (drop.p > 0.f) {
                        float m[4];
                        drop.mask4(c.grow, drop_ld, col, m);
#pragma unroll
                        for (int j = 0; j < 4; ++j) y[j] *= m[j];
                    }
                    if (col + 4 <= D && lc + 4 <= c.ncols) {
                        red_add_v4_f32(dst + col, y[0], y[1], y[2], y[3]);
                    } else {
                        for (int j = 0; j < 4; ++j)
                            if (col + j < D && lc + j < c.ncols) red_add_f32(dst + col + j, y[j]);
                    }
                }
            });
    }
};

}  // namespace nr
