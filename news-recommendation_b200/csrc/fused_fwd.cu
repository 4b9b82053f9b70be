// Fused front end of the NRMS news encoder, forward (reference src/model/NRMS/news_encoder.py:38-43 and
// src/model/general/attention/multihead_self.py:15-23,46-76):
//
//     ids --gather--> X (dropout) --tcgen05--> Q|K|V per head --tcgen05--> S = Q K^T --softmax--> P --tcgen05--> P V --> C (dropout)
//
// in ONE persistent kernel: the gathered rows, Q|K|V, the scores and the probabilities never leave the SM.  What reaches HBM
// is the context C (as a bf16 hi plane + a bf16 lo plane: C is the operand of the pooling GEMM and the precise input of the
// pooled sum) and, only when the caller asks for it (training: the backward kernels read it), the gathered rows X.
// Q|K|V is never written: the backward recomputes it from X.
//
// One CTA per SM, 14 warps:
//   warps 0-7   epilogue / softmax: two groups of four (one warp per TMEM lane quarter); group g owns the head PAIRS
//               (2j, 2j+1) with j = g (mod 2): a pair is 80 bytes of a context row, the granularity at which every warp
//               hands its 32 rows to TMA stores (row-per-thread 8-byte global stores cost one LSU sector cycle per lane:
//               ~190k cycles per tile, 7x the whole rest of the kernel, in the first version)
//   warps 8-11  gather: table rows -> registers -> (dropout) -> the SWIZZLE_128B A-operand tile X (row-contiguous 16-byte
//               pieces, 512 B of one row per warp instruction)
//   warp  12    TMA producer: streams the per-head weight block W_h = [W_Q[h] | W_K[h] | W_V[h] | 0] (64 rows) through a ring
//   warp  13    tcgen05 issuer: a small dependency-driven scheduler over three kinds of work -- Q|K|V of the next head
//               (in head order, paced by the weight ring), scores and P.V of each group's next head (as soon as the
//               operands that group produces are ready)
//
// A tile is 128 accumulator rows = 6 whole titles of 20 tokens (+8 dead rows), so the per-title attention is the block
// diagonal of ONE 128 x 128 x 32 score MMA per head; each row reads only its own 20 score columns (a 64-column TMEM window per
// lane quarter), runs the reference's exp-softmax with the +1e-8 in fp32 registers, and writes its probabilities back to
// TMEM as the bf16 A operand of the P.V MMA (A-from-TMEM).  V enters that MMA as a hi/lo bf16 pair ([V_hi | V_lo], N = 48):
// together with the hi/lo context planes this removes the two roundings that dominated the bf16 error of the unfused path
// (V and C: every token of a title sees the same rounding error of V_j, so pooling does not average it away).
//
// TMEM (512 columns): [0,128) two Q|K|V accumulators, [128,384) two score tiles (the P operand -- a hi/lo bf16 pair, 64 + 64
// packed columns -- overwrites its score tile), [384,512) two context accumulators.
#include <algorithm>
#include <cstring>

#define NR_WATCHDOG_SYMBOL g_fused_dev_error
#include "nr_fused.cuh"
#include "nr_ops.h"

namespace nr {

extern int g_launches;

int read_gru_device_error(int* out4);
int read_fused_device_error(int* out4) {
    const int rc = static_cast<int>(cudaMemcpyFromSymbol(out4, fused::g_fused_dev_error, sizeof(int) * 4));
    if (rc != 0 || out4[0] != 0) return rc;
    return read_gru_device_error(out4);  // the persistent GRU kernel keeps its own record (gru_persist.cu)
}

namespace fused {

constexpr int kThreads = 14 * 32;
constexpr int kWStages = 4;
constexpr int kNB = 64;                 // weight rows of one head block: 3 * d_k padded to a multiple of 16
constexpr int kNV = 48;                 // N of the P.V product: [V_hi (d_k) | pad to 24 | V_lo (d_k) | pad]
constexpr int kVLo = 24;                // first column of V_lo inside the V tile / the context accumulator
constexpr int kXChunk = 128 * 128;      // one 64-column k-chunk of the X tile
constexpr int kWStage = kNB * 128;
constexpr int kTile = 128 * 128;        // a [128][64] bf16 operand tile
constexpr int kStageBuf = 32 * 80;      // one warp's staging tile of one context plane: 32 rows x (2 heads x 20 columns) bf16

struct FwdParams {
    const long long* ids;
    const uint4* table;   // bf16 [V][ldx], 16-byte pieces
    int V;
    long long M;          // n_seq * T
    int num_tiles;
    int heads;
    int kch;              // 64-column k-chunks of X: ceil(d / 64)
    int ksteps_last;      // 16-column k-steps in the last chunk
    int d, ldx, ld3;
    const float* bias;    // [heads * kNB]
    float sc;             // log2(e) / sqrt(d_k)
    uint32_t thresh;      // dropout: keep iff 16-bit lane >= thresh; 0 = off
    float scale;
    uint64_t seed_x, seed_c;
    __nv_bfloat16* X;     // [M][ldx] or null
    __nv_bfloat16* C_hi;  // [M][ldx]
    __nv_bfloat16* C_lo;  // [M][ldx]
    int* bad_flag;
    long long* timing;    // tuning only (nr_debug_set_fused_timing): per CTA 32 cycle counters, see the roles
};
// TMA maps of the two context planes (dense boxes): [plane][0/1 = head pair / odd last head + ones column][0/1 = 32-row box /
// the shorter box of the last lane quarter]
struct CtxMaps {
    CUtensorMap m[2][2][2];
};

struct Smem {
    uint8_t* x;      // kch chunks of 16 KB
    uint8_t* w;      // kWStages stages of 8 KB
    uint8_t* qk;     // 2 tiles: Q in elements [0,32), K in [32,64) of every row
    uint8_t* v;      // 2 tiles: V_hi in elements [0,d_k), V_lo in [24, 24+d_k)
    uint8_t* stage;  // 8 warps x 2 planes x kStageBuf
    float* bias;
    uint64_t* bars;
    uint32_t* tmem_slot;
};
enum Bar { X_FULL = 0, X_EMPTY = 1, W_FULL = 2, W_EMPTY = 2 + kWStages, QKV_FULL = 2 + 2 * kWStages, QKV_EMPTY = QKV_FULL + 2,
           QK_READY = QKV_EMPTY + 2, S_FULL = QK_READY + 2, P_READY = S_FULL + 2, O_FULL = P_READY + 2, O_EMPTY = O_FULL + 2,
           NUM_BARS = O_EMPTY + 2 };

__host__ __device__ inline size_t smem_bytes(int heads, int kch) {
    return 1024 + static_cast<size_t>(kch) * kXChunk + kWStages * kWStage + 4 * kTile + 16 * kStageBuf + ((heads * kNB * 4 + 1023) & ~1023) + 1024;
}

__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// heads of group g inside a tile, in processing order: pairs (2j, 2j+1), j = g, g+2, ...
__device__ __forceinline__ int first_head(int g) { return 2 * g; }
__device__ __forceinline__ int next_head(int h, int H) { return (!(h & 1) && h + 1 < H) ? h + 1 : ((h >> 1) + 2) * 2; }
__device__ __forceinline__ int group_of(int h) { return (h >> 1) & 1; }

// ---------------------------------------------------------------------------------------------------------------------------
// scores -> exp-softmax (multihead_self.py:16-20) -> P as the bf16 A operand in TMEM, for the 32 rows of lane quarter QD.
// The only piece of the epilogue that needs the lane quarter at compile time (register-array offsets of the score window):
// everything else runs from ONE copy of the code (the fully specialised role was 4 x 25 KB of SASS; ncu: no_instruction stalls).
// ---------------------------------------------------------------------------------------------------------------------------
template <int T, int QD>
__device__ __forceinline__ void softmax_step(uint32_t s_t, int sel, float sc) {
    using W = Win<T, QD>;
    float x[T];
    {
        float v[64];
        tmem_ld32(s_t + W::start, v);
        tmem_ld32(s_t + W::start + 32, v + 32);
        tmem_ld_wait();
        constexpr int o0 = W::tlo * T - W::start;
        constexpr int o1 = W::ncand > 1 ? o0 + T : o0;       // candidates that do not exist alias candidate 0
        constexpr int o2 = W::ncand > 2 ? o0 + 2 * T : o0;
#pragma unroll
        for (int j = 0; j < T; ++j) {
            x[j] = v[o0 + j];
            if (W::ncand > 1 && sel == 1) x[j] = v[o1 + j];
            if (W::ncand > 2 && sel == 2) x[j] = v[o2 + j];
        }
    }
    float m = x[0];
#pragma unroll
    for (int j = 1; j < T; ++j) m = fmaxf(m, x[j]);
    m *= sc;
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < T; ++j) {
        x[j] = exp2f(fmaf(x[j], sc, -m));
        l += x[j];
    }
    const float inv = 1.f / (l + 1e-8f * exp2f(-m));  // == exp(S) / (sum exp(S) + 1e-8)
    // P leaves as a hi/lo bf16 pair (the bf16 rounding of the probabilities alone is 1.1e-3 of the logits): packed columns
    // [0, 64) of the score tile take P_hi, [64, 128) take P_lo; the issuer accumulates P_hi.V and P_lo.V into one accumulator
    uint32_t pk[T / 2], pl[T / 2];
#pragma unroll
    for (int i = 0; i < T / 2; ++i) {
        const float p0 = x[2 * i] * inv, p1 = x[2 * i + 1] * inv;
        pk[i] = pack_bf16x2(p0, p1);
        const float2 f = unpack_bf16x2(pk[i]);
        pl[i] = pack_bf16x2(p0 - f.x, p1 - f.y);
    }
    // all 128 columns of the tile are rewritten (the score MMA overwrote them): zeros off the diagonal
    uint32_t pw[64];
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
            pw[c] = 0u;
#pragma unroll
            for (int i = 0; i < W::ncand; ++i) {
                const int c0 = (W::tlo + i) * (T / 2);
                if (c >= c0 && c < c0 + T / 2) pw[c] = (sel == i) ? (part == 0 ? pk[c - c0] : pl[c - c0]) : 0u;
            }
        }
        tmem_st32(s_t + 64 * part, pw);
        tmem_st32(s_t + 64 * part + 32, pw + 32);
    }
    tmem_st_wait();
}

// ---------------------------------------------------------------------------------------------------------------------------
// epilogue / softmax role of one warp: TMEM lane quarter QD of group g
// ---------------------------------------------------------------------------------------------------------------------------
template <int T, int DK>
__device__ __noinline__ void epilogue_role(const FwdParams& p, const CtxMaps& maps, const Smem& sm, uint32_t tmem_base, int g,
                                           int QD, int lane) {
    using G = Geo<T>;
    static_assert(DK % 4 == 0 && DK <= kVLo && 3 * DK <= kNB && kVLo + DK <= kNV && 2 * DK * 2 * 32 <= kStageBuf, "head width");
    constexpr int H2 = DK / 2;  // packed words per head row
    const int kBoxRows = (G::kRows - 32 * QD) < 32 ? (G::kRows - 32 * QD) : 32;  // used rows of this lane quarter
    const int r = 32 * QD + lane;
    int t = r / T;
    if (t > G::kTPT - 1) t = G::kTPT - 1;  // dead rows ride with the last title (their results are never stored)
    const int sel = t - (32 * QD) / T;  // index among the titles of this lane quarter (Win<T, QD>::tlo)
    const uint32_t lane_base = static_cast<uint32_t>(QD * 32) << 16;
    const uint32_t qk_tile = smem_u32(sm.qk + g * kTile), v_tile = smem_u32(sm.v + g * kTile);
    const uint32_t acc_t = tmem_base + lane_base + g * kNB;
    const uint32_t s_t = tmem_base + lane_base + 128 + g * 128;
    const uint32_t o_t = tmem_base + lane_base + 384 + g * 64;
    const int H = p.heads;
    uint64_t* bars = sm.bars;
    uint8_t* st_hi = sm.stage + (g * 4 + QD) * 2 * kStageBuf;  // this warp's staging tiles
    uint8_t* st_lo = st_hi + kStageBuf;
    const CUtensorMap* m_hi = &maps.m[0][0][kBoxRows < 32 ? 1 : 0];
    const CUtensorMap* m_lo = &maps.m[1][0][kBoxRows < 32 ? 1 : 0];
    const CUtensorMap* m_hi_last = &maps.m[0][1][kBoxRows < 32 ? 1 : 0];
    const CUtensorMap* m_lo_last = &maps.m[1][1][kBoxRows < 32 ? 1 : 0];
    const int tail_cols = p.ldx - p.d;  // ones column + zero pad behind the last head

    // tuning counters (quarter 0 of each group): [g*8 + 0] wait Q|K|V, [1] step a, [2] wait scores, [3] step b, [4] wait context,
    // [5] step c, [6] whole role
    long long* tmr = (p.timing != nullptr && QD == 0 && lane == 0) ? p.timing + blockIdx.x * 32 + g * 8 : nullptr;
    long long tw[6] = {0, 0, 0, 0, 0, 0};
    const long long t_role = tmr != nullptr ? clock64() : 0;
    long long t_mark = t_role;
    auto lap = [&](int slot) {
        if (tmr != nullptr) {
            const long long t = clock64();
            tw[slot] += t - t_mark;
            t_mark = t;
        }
    };
    uint32_t k = 0;  // heads this group has processed (barrier phase)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const long long grow = static_cast<long long>(tile) * G::kRows + r;
        for (int h = first_head(g); h < H; h = next_head(h, H), ++k) {
            const uint32_t par = k & 1u;
            // ---- (a) Q|K|V accumulator -> + bias -> bf16 operand tiles -------------------------------------------------------
            f_wait(&bars[QKV_FULL + g], par, 331);
            tc_fence_after();
            lap(0);
            {
                float acc[kNB];
                tmem_ld32(acc_t, acc);
                tmem_ld32(acc_t + 32, acc + 32);
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive(&bars[QKV_EMPTY + g], lane);  // the issuer may start Q|K|V of this group's next head
                const float* b = sm.bias + h * kNB;
#pragma unroll
                for (int j = 0; j < 3 * DK; j += 4) {
                    const float4 b4 = lds_f4(b + j);
                    acc[j] += b4.x; acc[j + 1] += b4.y; acc[j + 2] += b4.z; acc[j + 3] += b4.w;
                }
                uint32_t qw[H2], kw[H2], vh[H2], vl[H2];
#pragma unroll
                for (int i = 0; i < H2; ++i) {
                    qw[i] = pack_bf16x2(acc[2 * i], acc[2 * i + 1]);
                    kw[i] = pack_bf16x2(acc[DK + 2 * i], acc[DK + 2 * i + 1]);
                    const float v0 = acc[2 * DK + 2 * i], v1 = acc[2 * DK + 2 * i + 1];
                    vh[i] = pack_bf16x2(v0, v1);
                    const float2 f = unpack_bf16x2(vh[i]);
                    vl[i] = pack_bf16x2(v0 - f.x, v1 - f.y);
                }
                // the tiles of this group are free: step (c) of its previous head waited for that head's P.V MMA (and with it
                // the score MMA)
                sts_row20(qk_tile, r, 0, qw);
                sts_row20(qk_tile, r, 32, kw);
                sts_row20(v_tile, r, 0, vh);
                sts_row20(v_tile, r, kVLo, vl);
                fence_proxy_async();
                warp_arrive(&bars[QK_READY + g], lane);
            }
            // ---- (b) scores -> exp-softmax -> P as the bf16 A operand in TMEM --------------------------------------------------
            lap(1);
            f_wait(&bars[S_FULL + g], par, 332);
            tc_fence_after();
            lap(2);
            switch (QD) {
                case 0: softmax_step<T, 0>(s_t, sel, p.sc); break;
                case 1: softmax_step<T, 1>(s_t, sel, p.sc); break;
                case 2: softmax_step<T, 2>(s_t, sel, p.sc); break;
                default: softmax_step<T, 3>(s_t, sel, p.sc); break;
            }
            tc_fence_before();
            warp_arrive(&bars[P_READY + g], lane);
            // ---- (c) context accumulator -> hi + lo parts -> dropout -> staging tiles -> TMA store per head pair -------------
            lap(3);
            f_wait(&bars[O_FULL + g], par, 333);
            tc_fence_after();
            lap(4);
            {
                float o[kNV];
                tmem_ld32(o_t, o);
                tmem_ld16(o_t + 32, o + 32);
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive(&bars[O_EMPTY + g], lane);
                const bool single = !(h & 1) && h + 1 >= H;  // odd head count: the last head travels with the ones column
                const int pitch = single ? (DK + tail_cols) * 2 : 4 * DK;  // bytes per staging row
                if (!(h & 1)) {  // first head of a pair: the TMA stores of the previous pair have read the staging tiles
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                uint8_t* row_hi = st_hi + lane * pitch + (h & 1) * (2 * DK);
                uint8_t* row_lo = st_lo + lane * pitch + (h & 1) * (2 * DK);
#pragma unroll
                for (int q4 = 0; q4 < DK; q4 += 4) {
                    float c4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) c4[j] = o[q4 + j] + o[kVLo + q4 + j];
                    if (p.thresh != 0u) {
                        float mk[4];
                        drop4(p.seed_c, p.thresh, p.scale, grow, p.ldx, h * DK + q4, mk);
#pragma unroll
                        for (int j = 0; j < 4; ++j) c4[j] *= mk[j];
                    }
                    const uint32_t h0 = pack_bf16x2(c4[0], c4[1]), h1 = pack_bf16x2(c4[2], c4[3]);
                    const float2 f0 = unpack_bf16x2(h0), f1 = unpack_bf16x2(h1);
                    *reinterpret_cast<uint2*>(row_hi + 2 * q4) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(row_lo + 2 * q4) =
                        make_uint2(pack_bf16x2(c4[0] - f0.x, c4[1] - f0.y), pack_bf16x2(c4[2] - f1.x, c4[3] - f1.y));
                }
                if (single) {  // ones column (1.0 in the hi plane) + zero pad
                    for (int j = 0; j < tail_cols; ++j) {
                        reinterpret_cast<__nv_bfloat16*>(row_hi)[DK + j] = __float2bfloat16_rn(j == 0 ? 1.0f : 0.f);
                        reinterpret_cast<__nv_bfloat16*>(row_lo)[DK + j] = __float2bfloat16_rn(0.f);
                    }
                }
                if ((h & 1) || single) {  // pair complete: rows beyond M are clipped by the tensor map, dead rows by the box
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && kBoxRows > 0) {
                        const int col0 = (h & ~1) * DK;
                        const int row0 = static_cast<int>(static_cast<long long>(tile) * G::kRows + 32 * QD);
                        tma_store_2d(single ? m_hi_last : m_hi, st_hi, col0, row0);
                        tma_store_2d(single ? m_lo_last : m_lo, st_lo, col0, row0);
                        bulk_commit();
                    }
                }
            }
            lap(5);
        }
    }
    if (lane == 0) bulk_wait_all();
    if (tmr != nullptr) {
        for (int i = 0; i < 6; ++i) tmr[i] = tw[i];
        tmr[6] = clock64() - t_role;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------------
template <int T, int DK>
__global__ void __launch_bounds__(kThreads, 1) mhsa_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmW,
                                                                      const __grid_constant__ CtxMaps maps, const FwdParams p) {
    using G = Geo<T>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Smem sm;
    sm.x = base;
    sm.w = sm.x + p.kch * kXChunk;
    sm.qk = sm.w + kWStages * kWStage;
    sm.v = sm.qk + 2 * kTile;
    sm.stage = sm.v + 2 * kTile;
    sm.bias = reinterpret_cast<float*>(sm.stage + 16 * kStageBuf);
    sm.bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm.bias) + ((p.heads * kNB * 4 + 1023) & ~1023));
    sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.bars + NUM_BARS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.heads;

    // zero every operand tile once: the K padding of Q / K (elements d_k..31), the unused columns of V and the dead rows
    // must be finite zeros for the whole kernel; nothing below ever writes them
    {
        uint4* z = reinterpret_cast<uint4*>(base);
        const int n16 = static_cast<int>(sm.stage - base) >> 4;
        for (int i = threadIdx.x; i < n16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < H * kNB; i += kThreads) sm.bias[i] = p.bias[i];
    }
    fence_proxy_async();
    if (warp == 12 && lane == 0) {
        tma_prefetch_desc(&tmW);
        mbar_init(&sm.bars[X_FULL], 4);
        mbar_init(&sm.bars[X_EMPTY], 1);
        for (int i = 0; i < kWStages; ++i) {
            mbar_init(&sm.bars[W_FULL + i], 1);
            mbar_init(&sm.bars[W_EMPTY + i], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&sm.bars[QKV_FULL + g], 1);
            mbar_init(&sm.bars[QKV_EMPTY + g], 4);
            mbar_init(&sm.bars[QK_READY + g], 4);
            mbar_init(&sm.bars[S_FULL + g], 1);
            mbar_init(&sm.bars[P_READY + g], 4);
            mbar_init(&sm.bars[O_FULL + g], 1);
            mbar_init(&sm.bars[O_EMPTY + g], 4);
        }
        fence_barrier_init();
    } else if (warp == 13) {
        tmem_alloc(sm.tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    uint64_t* bars = sm.bars;

    if (warp < 8) {
        epilogue_role<T, DK>(p, maps, sm, tmem_base, warp >> 2, warp & 3, lane);
    } else if (warp < 12) {
        // ===================== gather: table rows -> (dropout) -> X tile (+ the X rows in HBM) =====================
        const int gw = warp - 8;
        const int chunks = p.ldx >> 3;                 // 16-byte pieces of a table row
        const int smem_pieces = p.kch * 8;             // pieces that exist in the X tile
        const uint32_t x_s = smem_u32(sm.x);
        const bool tail_here = (H & 1) == 0;           // even head count: nobody else writes the ones column of the context
        long long* tmr = (p.timing != nullptr && gw == 0 && lane == 0) ? p.timing + blockIdx.x * 32 + 16 : nullptr;  // [16] wait X free, [17] gather
        long long tg_wait = 0, tg_work = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const long long tg0 = tmr != nullptr ? clock64() : 0;
            f_wait(&bars[X_EMPTY], static_cast<uint32_t>(it & 1) ^ 1u, 321);
            const long long tg1 = tmr != nullptr ? clock64() : 0;
            tg_wait += tg1 - tg0;
            const long long row0 = static_cast<long long>(tile) * G::kRows;
            for (int b = 0; b < 8; ++b) {
                uint4 u[4][2];
                bool ok[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = gw + 4 * (4 * b + i);
                    const long long gr = row0 + r;
                    ok[i] = r < G::kRows && gr < p.M;
                    long long id = ok[i] ? __ldg(p.ids + gr) : 0;
                    if (id < 0 || id >= p.V) {
                        if (lane == 0) atomicExch(p.bad_flag, 1);
                        id = 0;
                    }
                    const uint4* src = p.table + id * chunks;
                    u[i][0] = (ok[i] && lane < chunks) ? __ldg(src + lane) : make_uint4(0u, 0u, 0u, 0u);
                    u[i][1] = (ok[i] && lane + 32 < chunks) ? __ldg(src + 32 + lane) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!ok[i]) continue;
                    const int r = gw + 4 * (4 * b + i);
                    const long long gr = row0 + r;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int pc = lane + 32 * half;
                        if (pc >= chunks) continue;
                        const int col = pc * 8;
                        uint32_t w[4] = {u[i][half].x, u[i][half].y, u[i][half].z, u[i][half].w};
                        if (p.thresh != 0u) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                float mk[4];
                                drop4(p.seed_x, p.thresh, p.scale, gr, p.ldx, col + 4 * hh, mk);
                                float2 f0 = unpack_bf16x2(w[2 * hh]), f1 = unpack_bf16x2(w[2 * hh + 1]);
                                w[2 * hh] = pack_bf16x2(f0.x * mk[0], f0.y * mk[1]);
                                w[2 * hh + 1] = pack_bf16x2(f1.x * mk[2], f1.y * mk[3]);
                            }
                        }
                        if (pc < smem_pieces) sts128(x_s + (pc >> 3) * kXChunk + sw128_off(r, pc & 7), w[0], w[1], w[2], w[3]);
                        if (p.X != nullptr) {
                            if (p.d >= col && p.d < col + 8) {  // ones column of the HBM copy (bias-gradient trick), zeros behind it
                                __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(w);
                                for (int j = p.d - col; j < 8; ++j) e[j] = __float2bfloat16_rn(j == p.d - col ? 1.0f : 0.f);
                            }
                            reinterpret_cast<uint4*>(p.X)[gr * chunks + pc] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                    if (tail_here && lane < p.ldx - p.d) {  // ones column + zero tail of the context planes
                        p.C_hi[gr * p.ldx + p.d + lane] = __float2bfloat16_rn(lane == 0 ? 1.0f : 0.f);
                        p.C_lo[gr * p.ldx + p.d + lane] = __float2bfloat16_rn(0.f);
                    }
                }
            }
            fence_proxy_async();
            warp_arrive(&bars[X_FULL], lane);
            if (tmr != nullptr) tg_work += clock64() - tg1;
        }
        if (tmr != nullptr) { tmr[0] = tg_wait; tmr[1] = tg_work; }
    } else if (warp == 12) {
        // ===================== TMA producer: per-head weight blocks, k-chunk by k-chunk =====================
        int st = 0;
        uint32_t ph = 0;
        long long tp_wait = 0;  // [20] producer waiting for a free weight stage
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (int h = 0; h < H; ++h)
                for (int kc = 0; kc < p.kch; ++kc) {
                    const long long tp0 = p.timing != nullptr ? clock64() : 0;
                    f_wait(&bars[W_EMPTY + st], ph ^ 1u, 301);
                    if (p.timing != nullptr) tp_wait += clock64() - tp0;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&bars[W_FULL + st], kWStage);
                        tma_load_2d(sm.w + st * kWStage, &tmW, &bars[W_FULL + st], kc * 64, h * kNB);
                    }
                    __syncwarp();
                    if (++st == kWStages) { st = 0; ph ^= 1u; }
                }
        }
        if (p.timing != nullptr && lane == 0) p.timing[blockIdx.x * 32 + 20] = tp_wait;
    } else {
        // ===================== tcgen05 issuer: dependency-driven over Q|K|V (head order) and each group's S / P.V =====================
        const uint32_t idesc_qkv = make_idesc_bf16(128, kNB, 0, 0);
        const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
        const uint32_t idesc_pv = make_idesc_bf16(128, kNV, 0, 1);
        constexpr int KS_S = (DK + 15) / 16;  // k-steps of the score product (Q / K are zero padded to 32 elements)
        const uint32_t x_s = smem_u32(sm.x), w_s = smem_u32(sm.w), qk_s = smem_u32(sm.qk), v_s = smem_u32(sm.v);
        auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {  // warp-uniform non-blocking test
            return __shfl_sync(0xffffffffu, mbar_test_wait(bar, parity) ? 1 : 0, 0) != 0;
        };
        int st = 0;
        uint32_t ph = 0;
        uint32_t cq0 = 0u, cq1 = 0u;                 // Q|K|V projections issued per group (barrier phases); scalars: a run-time
        uint32_t cs[2] = {0u, 0u}, cp[2] = {0u, 0u};  // indexed array would live in local memory (cs / cp are indexed by the unrolled g)
        // tuning counters: [24] wait X, [25] Q|K|V issue incl. weight waits, [26] weight waits alone, [27] idle polling, [28] role, [29] idle polls
        long long tm_x = 0, tm_q = 0, tm_w = 0, tm_idle = 0, n_idle = 0;
        const bool tmon = p.timing != nullptr;
        const long long tm_role = tmon ? clock64() : 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const long long tx0 = tmon ? clock64() : 0;
            f_wait(&bars[X_FULL], static_cast<uint32_t>(it & 1), 311);
            tc_fence_after();
            if (tmon) tm_x += clock64() - tx0;
            int q_next = 0;                                  // next head to project
            int s_next[2] = {first_head(0), first_head(1)};  // per group: next head whose scores / P.V are due
            int p_next[2] = {first_head(0), first_head(1)};
            uint32_t idle_polls = 0;
            while (p_next[0] < H || p_next[1] < H) {
                bool progressed = false;
                const long long ti0 = tmon ? clock64() : 0;
                if (q_next < H) {  // ---- Q|K|V of head q_next
                    const int h = q_next, g = group_of(h);
                    if (ready(&bars[QKV_EMPTY + g], ((g ? cq1 : cq0) & 1u) ^ 1u)) {
                        tc_fence_after();
                        const uint32_t d_t = tmem_base + g * kNB;
                        for (int kc = 0; kc < p.kch; ++kc) {
                            const long long tw0 = tmon ? clock64() : 0;
                            f_wait(&bars[W_FULL + st], ph, 313);
                            tc_fence_after();
                            if (tmon) tm_w += clock64() - tw0;
                            if (elect_one()) {
                                const uint64_t da = make_sw128_desc(x_s + kc * kXChunk, 0, 1024);
                                const uint64_t db = make_sw128_desc(w_s + st * kWStage, 0, 1024);
                                const int nk = (kc == p.kch - 1) ? p.ksteps_last : 4;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (k < nk) umma_bf16(d_t, da + 2 * k, db + 2 * k, idesc_qkv, (kc | k) ? 1u : 0u);
                                umma_commit(&bars[W_EMPTY + st]);
                            }
                            __syncwarp();
                            if (++st == kWStages) { st = 0; ph ^= 1u; }
                        }
                        if (elect_one()) {
                            umma_commit(&bars[QKV_FULL + g]);
                            if (h == H - 1) umma_commit(&bars[X_EMPTY]);  // the gather warps may refill X for the next tile
                        }
                        __syncwarp();
                        if (g) ++cq1; else ++cq0;
                        ++q_next;
                        progressed = true;
                        if (tmon) tm_q += clock64() - ti0;
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    // ---- scores of this group's next head (its Q|K|V MMAs were issued: s_next < q_next in head order)
                    if (s_next[g] < H && s_next[g] < q_next && ready(&bars[QK_READY + g], cs[g] & 1u)) {
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t da = make_sw128_desc(qk_s + g * kTile, 0, 1024);
                            const uint64_t db = make_sw128_desc(qk_s + g * kTile + 64, 0, 1024);  // K lives in elements [32, 64)
#pragma unroll
                            for (int k = 0; k < KS_S; ++k)
                                umma_bf16(tmem_base + 128 + g * 128, da + 2 * k, db + 2 * k, idesc_s, k ? 1u : 0u);
                            umma_commit(&bars[S_FULL + g]);
                        }
                        __syncwarp();
                        ++cs[g];
                        s_next[g] = next_head(s_next[g], H);
                        progressed = true;
                    }
                    // ---- P.V of this group's next head: A = P (bf16, TMEM), B = [V_hi | V_lo] (MN-major)
                    if (p_next[g] < H && cp[g] < cs[g] && ready(&bars[P_READY + g], cp[g] & 1u) &&
                        ready(&bars[O_EMPTY + g], (cp[g] & 1u) ^ 1u)) {
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t db = make_sw128_desc(v_s + g * kTile, 8192, 1024);
#pragma unroll
                            for (int k = 0; k < 16; ++k)  // P_hi then P_lo; 16 key rows per k-step: +2048 bytes in B, +8 packed columns in A
                                umma_bf16_ts(tmem_base + 384 + g * 64, tmem_base + 128 + g * 128 + 8 * k, db + 128 * (k & 7), idesc_pv,
                                             k ? 1u : 0u);
                            umma_commit(&bars[O_FULL + g]);
                        }
                        __syncwarp();
                        ++cp[g];
                        p_next[g] = next_head(p_next[g], H);
                        progressed = true;
                    }
                }
                if (progressed) {
                    idle_polls = 0;
                } else {  // nothing was ready: back off (the polls would otherwise take this scheduler's issue slots from the
                          // epilogue warps sharing it); bounded like every other wait of this kernel
                    __nanosleep(64);
                    if (++idle_polls > 20000000u) f_timeout(317, static_cast<uint32_t>(q_next));
                    if (tmon) { tm_idle += clock64() - ti0; ++n_idle; }
                }
            }
        }
        if (tmon && lane == 0) {
            long long* tmr = p.timing + blockIdx.x * 32 + 24;
            tmr[0] = tm_x; tmr[1] = tm_q; tmr[2] = tm_w; tmr[3] = tm_idle; tmr[4] = clock64() - tm_role; tmr[5] = n_idle;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) tmem_dealloc(tmem_base, 512);
}

}  // namespace fused

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
static long long* g_fused_timing = nullptr;
void set_debug_fused_timing(void* dev_buf) { g_fused_timing = static_cast<long long*>(dev_buf); }

int mhsa_fused_supported(int T, int d, int heads) {
    return (T == 20 && heads >= 1 && heads <= 16 && d == heads * 20 && d <= 320) ? 1 : 0;
}

int mhsa_fused_fwd(const long long* ids, long long n_seq, int T, const void* table, int V, int d, int heads, int ldx,
                   const void* w_heads, const float* b_heads, DropoutCfg drop_x, DropoutCfg drop_c, void* X, void* C_hi, void* C_lo,
                   int* bad_id_flag, cudaStream_t stream) {
    using namespace fused;
    NR_REQUIRE(mhsa_fused_supported(T, d, heads), "mhsa_fused_fwd: unsupported shape T=%d d=%d heads=%d", T, d, heads);
    NR_REQUIRE(ldx % 8 == 0 && ldx >= d + 1 && ldx <= 512, "mhsa_fused_fwd: pitch ldx=%d", ldx);
    NR_REQUIRE(n_seq * T < (1ll << 31), "mhsa_fused_fwd: too many tokens");
    if (n_seq == 0) return 0;
    constexpr int TT = 20, DK = 20;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.ids = ids;
    p.table = static_cast<const uint4*>(table);
    p.V = V;
    p.M = n_seq * T;
    p.num_tiles = static_cast<int>((n_seq + Geo<TT>::kTPT - 1) / Geo<TT>::kTPT);
    p.heads = heads;
    p.kch = ceil_div(d, 64);
    p.ksteps_last = ceil_div(d - (p.kch - 1) * 64, 16);
    p.d = d;
    p.ldx = ldx;
    p.bias = b_heads;
    p.sc = 1.4426950408889634f / sqrtf(static_cast<float>(d / heads));
    NR_REQUIRE(drop_x.p == drop_c.p, "mhsa_fused_fwd: one dropout probability for both sites");
    p.thresh = drop_x.p > 0.f ? static_cast<uint32_t>(drop_x.p * 65536.0f + 0.5f) : 0u;
    p.scale = drop_x.p > 0.f ? 1.f / (1.f - drop_x.p) : 1.f;
    p.seed_x = drop_x.seed;
    p.seed_c = drop_c.seed;
    p.X = static_cast<__nv_bfloat16*>(X);
    p.C_hi = static_cast<__nv_bfloat16*>(C_hi);
    p.C_lo = static_cast<__nv_bfloat16*>(C_lo);
    p.bad_flag = bad_id_flag;
    p.timing = g_fused_timing;
    CUtensorMap tmW;
    NR_PROPAGATE(make_tmap_bf16_2d(&tmW, w_heads, static_cast<int64_t>(heads) * kNB, d, ldx, 64, kNB));
    // context planes: dense boxes of one head pair (or the odd last head + ones column) x the rows of one lane quarter
    CtxMaps maps;
    const int last_rows = Geo<TT>::kRows - 96;  // used rows of the last lane quarter
    const int single_cols = DK + (ldx - d);
    NR_REQUIRE((heads & 1) == 0 || (single_cols * 2) % 16 == 0, "mhsa_fused_fwd: odd head count needs (d_k + ldx - d) * 2 %% 16 == 0");
    for (int pl = 0; pl < 2; ++pl) {
        void* base = pl == 0 ? C_hi : C_lo;
        for (int rb = 0; rb < 2; ++rb) {
            const int rows = rb == 0 ? 32 : (last_rows > 0 ? last_rows : 32);
            NR_PROPAGATE(make_tmap_bf16_2d(&maps.m[pl][0][rb], base, p.M, ldx, ldx, 2 * DK, rows, 0));
            NR_PROPAGATE(make_tmap_bf16_2d(&maps.m[pl][1][rb], base, p.M, ldx, ldx, (heads & 1) ? single_cols : 2 * DK, rows, 0));
        }
    }
    const size_t smem = smem_bytes(heads, p.kch);
    NR_REQUIRE(smem <= 232448, "mhsa_fused_fwd: %zu bytes of shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        NR_CHECK_CUDA(cudaFuncSetAttribute(mhsa_fused_fwd_kernel<TT, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    const int grid = std::min(p.num_tiles, num_sms());
    ProfScope ps("mhsa_fused_fwd", static_cast<int>(n_seq), T, d, stream);
    mhsa_fused_fwd_kernel<TT, DK><<<grid, kThreads, smem, stream>>>(tmW, maps, p);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nr
