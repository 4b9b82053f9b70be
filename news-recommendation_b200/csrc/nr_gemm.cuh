// Weight-stationary persistent tcgen05 GEMMs for the news-recommendation hot path (sm_100a).
//
//  gemm_nt : D[M x N] = A[M x K] . B[N x K]^T      (both K-major; "activation x weight^T")
//            * a CTA pair's weight slice (<=256 output columns, all K, all conv taps; half the rows in
//              each CTA) is loaded ONCE by TMA and stays resident in shared memory; 128-row activation
//              tiles stream through a TMA/mbarrier ring; accumulators are double buffered in TMEM
//              (2 x 256 fp32 columns); 8 epilogue warps run a fused epilogue functor on tcgen05.ld'ed rows.
//            * conv taps: tap s re-loads the A tile shifted by (s - taps/2) rows (zero rows separate
//              the segments in the padded layout), accumulating into the same TMEM tile.
//  gemm_tn : D[Ma x Nb] += A[Kr x Ma]^T . B[Kr x Nb]  (both MN-major; weight gradients, Kr = all tokens)
//            split over Kr across CTAs, fp32 red.global.add epilogue.
//
// Warp roles of gemm_nt (kGemmThreads per CTA, CTAs launched as pairs): warps 0..kEpiWarps-1 epilogue (TMEM lane
// quarter = warp & 3; the kEpiParts warps of a quarter split the accumulator columns -- with 4 warps the row-per-thread
// epilogue was latency bound at ~20 % issue utilisation, ncu profiles/), then the TMA producer warp and the warp that
// allocates TMEM and, in the leader CTA, issues the MMAs.
// gemm_tn keeps 192 threads (4 epilogue warps, one-shot epilogue).
#pragma once
#include "nr_common.cuh"

namespace nr {

constexpr int kTnThreads = 192;     // gemm_tn
constexpr int kEpiWarps = 8;  // gemm_nt epilogue warps: 4 TMEM lane quarters x kEpiParts column parts
constexpr int kEpiParts = kEpiWarps / 4;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = kEpiThreads + 64;  // + TMA producer warp + MMA warp
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                     // bf16 elements per 128-byte swizzle row
constexpr int kAStageBytes = kTileM * 128;      // 16 KB
constexpr int kMaxStages = 12;
constexpr int kSmemLimit = 232448;              // 227 KB

struct GemmNTParams {
    int M;              // rows of A that exist
    int rows_per_tile;  // rows OWNED by one M tile (<=128); tile t loads rows [t*rpt, t*rpt+128)
    int num_m_tiles;
    int N;              // output columns
    int n_stride;       // columns per weight slice (multiple of 16)
    int n_slices;
    int n_box;          // rows of one resident weight box (multiple of 16, <=256)
    int K;              // reduction length per tap (elements)
    int k_chunks;       // ceil(K/64)
    int taps;           // 1, or 3 for the window-3 title CNN
    int b_tap_rows;     // row offset between taps inside the weight operand
    int stages;
    float* dbg_acc;     // debug backend only: fp32 accumulators [num_m_tiles*128][dbg_ld]
    int dbg_ld;
    int dbg_flags;      // tuning only (NEWSREC_GEMM_DBG): bit 1 = the producers skip the A loads (MMA on stale data)
    long long* timing;  // tuning only (nr_debug_set_gemm_timing): per CTA 16 cycle counters, see the kernel
};

// What an epilogue functor sees for one (tile,row).
struct EpiCtx {
    int tile;
    int r;        // row inside the tile (0..127) == TMEM lane
    int grow;     // global A row
    bool valid;   // r < rows_per_tile && grow < M
    int col0;     // first output column of this CTA's slice
    int ncols;    // valid output columns in the slice
    int tid;      // 0..255 within the epilogue group
    int half;     // 0 .. kEpiParts-1: which of the warps of this TMEM lane quarter (column part)
    int ch0, ch1; // this thread's range of 32-column chunks
    float* scratch;  // Epi::kScratchBytes of shared memory private to the epilogue group
    int it;          // how many tiles this CTA has finished before this one (double-buffer parity)
    int next_tile;   // the tile this CTA processes next, or -1
};
// What init()/finish() see.
struct EpiInit {
    int col0, ncols, tid;
    float* scratch;
    int first_tile;  // first tile of this CTA (>= num_tiles: the CTA has no work)
    int num_tiles;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
// the two column halves run different chunk counts: barriers inside a chunk loop are per half

__device__ __forceinline__ void epi_chunk_range(int ncols, int part, int& ch0, int& ch1) {
    const int nch = (ncols + 31) >> 5, base = nch / kEpiParts, rem = nch - base * kEpiParts;
    ch0 = part * base + min(part, rem);
    ch1 = ch0 + base + (part < rem ? 1 : 0);
}

struct TmemAcc {
    uint32_t taddr;
    uint32_t release_bar;  // shared::cluster address of the leader CTA's "accumulator free" barrier
    __device__ __forceinline__ void load32(int chunk, float* v) const {
        tmem_ld32(taddr + chunk * 32, v);
        tmem_ld_wait();
    }
    __device__ __forceinline__ void issue32(int chunk, float* v) const { tmem_ld32(taddr + chunk * 32, v); }
    __device__ __forceinline__ void wait32(float* v) const { tmem_ld_wait32(v); }
    // All TMEM reads of this tile by this WARP are done (every call site is warp-uniform): one arrival per warp.
    __device__ __forceinline__ void release() const {
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(release_bar);
    }
};
struct GlobalAcc {  // debug backend: accumulators computed by a plain SIMT kernel
    const float* row;
    __device__ __forceinline__ void load32(int chunk, float* v) const {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = row[chunk * 32 + j];
    }
    __device__ __forceinline__ void issue32(int chunk, float* v) const { load32(chunk, v); }
    __device__ __forceinline__ void wait32(float*) const {}
    __device__ __forceinline__ void release() const {}
};

// The chunk loop every epilogue shares, software pipelined over two register buffers: the tcgen05.ld of chunk i+1 is
// in flight while body(i) runs (a single-buffered loop exposed the TMEM round trip once per chunk).  pre(ch) runs
// before the wait of chunk ch (shared-memory operand loads go there).  Calls acc.release() exactly once, right after
// the last chunk has landed in registers.
template <class Acc, class Pre, class Body>
__device__ __forceinline__ void epi_chunks(const Acc& acc, const EpiCtx& c, Pre&& pre, Body&& body) {
    if (c.ch0 >= c.ch1) {
        acc.release();
        return;
    }
    float xa[32], xb[32];
    acc.issue32(c.ch0, xa);
    for (int ch = c.ch0; ch < c.ch1; ch += 2) {
        pre(ch);
        acc.wait32(xa);
        if (ch + 1 < c.ch1) acc.issue32(ch + 1, xb); else acc.release();
        body(ch, xa);
        if (ch + 1 < c.ch1) {
            pre(ch + 1);
            acc.wait32(xb);
            if (ch + 2 < c.ch1) acc.issue32(ch + 2, xa); else acc.release();
            body(ch + 1, xb);
        }
    }
}

// bf16 output tiles leave through TMA instead of 32 scattered rows per store instruction: a warp packs its 32 rows x 32
// columns into a private staging buffer (SWIZZLE_64B layout: 16-byte chunk q of row r at r*64 + ((q ^ (r>>1)) & 3)*16,
// conflict-free for row-per-lane 16-byte writes) and one lane issues cp.async.bulk.tensor.  Two buffers per warp.
// Measured alone (tools/stbench.cu): 5.7 TB/s vs 2.6 (2 x STG.128) / 5.0 (STG.256); inside the GEMM the row-per-thread
// stores sat in the LSU queue and stalled the warps on their source registers (ncu, profiles/).
constexpr int kTileStoreBufs = 2;                                  // staging tiles per warp (2 KB each)
constexpr int kTileStoreBytes = kEpiWarps * kTileStoreBufs * 2048 + 1024;  // + alignment slack
struct WarpTileStore {
    uint8_t* buf;  // this warp's kTileStoreBufs x 2 KB (1024-byte aligned)
    int nput;
    __device__ __forceinline__ void attach(void* region, int warp_in_group) {
        buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(region) + 1023) & ~uintptr_t(1023)) + warp_in_group * (kTileStoreBufs * 2048);
    }
    __device__ __forceinline__ void begin_tile(int lane) {
        nput = 0;
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    // w: this lane's row, 32 columns as 16 packed bf16 pairs; (col, row0) = global coordinates of the warp's tile
    __device__ __forceinline__ void put(const CUtensorMap* tm, const uint32_t* w, int col, int row0, int lane) {
        uint8_t* b = buf + (nput % kTileStoreBufs) * 2048;
        if (nput >= kTileStoreBufs) {
            if (lane == 0) bulk_wait_read<kTileStoreBufs - 1>();  // the store issued from this buffer has been read out
            __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(b + lane * 64 + ((q ^ (lane >> 1)) & 3) * 16) =
                make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tm, b, col, row0);
            bulk_commit();
        }
        ++nput;
    }
    // Row-mapped destinations (no tensor map fits them): the same staging tile leaves as coalesced 16-byte stores, 8 rows x 64
    // contiguous bytes per instruction.  my_orow = destination row of this lane's accumulator row, -1 for "drop the row".
    __device__ __forceinline__ void put_rows(__nv_bfloat16* base, int ld, const uint32_t* w, int col, int my_orow, int lane) {
        uint8_t* b = buf;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(b + lane * 64 + ((q ^ (lane >> 1)) & 3) * 16) =
                make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = (lane >> 2) + 8 * k, q = lane & 3;
            const int orr = __shfl_sync(0xffffffffu, my_orow, r);
            if (orr >= 0)
                *(reinterpret_cast<uint4*>(base + static_cast<long long>(orr) * ld + col) + q) =
                    *reinterpret_cast<const uint4*>(b + r * 64 + ((q ^ (r >> 1)) & 3) * 16);
        }
        __syncwarp();
    }
    static __device__ __forceinline__ void drain(int lane) {  // before the CTA exits
        if (lane == 0) bulk_wait_all();
    }
};

// ---------------------------------------------------------------------------------------------
// gemm_nt kernel: one CTA PAIR (cluster of 2, tcgen05 cta_group::2) per (256-row block, weight slice)
// ---------------------------------------------------------------------------------------------
// Both CTAs stream their own 128-row activation tiles and hold HALF of the slice's weight rows; the leader (cluster rank
// 0) issues M=256 MMAs that read A and B from both CTAs' shared memory and leave rows 0-127 of D in its own TMEM, rows
// 128-255 in the peer's.  Halving the resident weight bytes is what buys the A ring its depth: with 4 stages the ring
// was latency bound (a stage is refilled only after its MMAs retire; (TMA latency + MMA time) / stages > MMA time).
// mbarrier protocol (every barrier exists in both CTAs at the same offset; "L" = only the leader's copy is used):
//   bfull  L  count 1 + tx of both weight halves        -> MMA issuer may start
//   full[s] L count 1 + tx of both CTAs' A boxes         (leader's producer arrives with expect_tx of 2 boxes; the
//                                                         peer's TMA signals the leader's barrier, cta_group::2)
//   empty[s]  count 1, multicast tcgen05.commit          -> each CTA's producer refills its own stage s
//   tfull[a]  count 1, multicast tcgen05.commit          -> each CTA's epilogue reads its own TMEM rows
//   tempty[a] L count 16 (8 epilogue warps x 2 CTAs, the peer arrives through shared::cluster)
template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmNTParams p,
               const __grid_constant__ Epi epi) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const int b_region = (p.n_box >> 1) * 128;  // bytes of one (tap, k-chunk) box of this CTA's weight half
    uint8_t* sB = smem;
    uint8_t* sA = sB + p.taps * p.k_chunks * b_region;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + p.stages * kAStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* bfull = bars + 2 * kMaxStages;
    uint64_t* tfull = bars + 2 * kMaxStages + 1;
    uint64_t* tempty = bars + 2 * kMaxStages + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
    float* scratch = reinterpret_cast<float*>(bars + 2 * kMaxStages + 8);

    const int pair = blockIdx.x >> 1;
    const int slice = pair % p.n_slices;
    const int ptile0 = pair / p.n_slices;                // pair tile = 256 rows_per_tile-rows: tiles 2*pt and 2*pt + 1
    const int ptile_step = (gridDim.x >> 1) / p.n_slices;
    const int num_ptiles = (p.num_m_tiles + 1) >> 1;
    const int col0 = slice * p.n_stride;
    const int ncols = min(p.n_stride, p.N - col0);
    const int n_mma = (ncols + 15) & ~15;                // MMA N of this slice; each CTA supplies n_mma/2 weight rows
    const int tap_shift = p.taps / 2;
    // tuning counters: [0] producer waits for a free A stage, [1] MMA waits for A data, [2] MMA waits for a free
    // accumulator, [3] epilogue waits for a finished accumulator, [4] epilogue body, [5] kernel, [6] tiles,
    // [7] MMA issue loops, [8] tcgen05.commit
    long long* tmr = p.timing != nullptr ? p.timing + blockIdx.x * 16 : nullptr;
    const long long t_begin = tmr != nullptr ? clock64() : 0;
    long long tw_a = 0, tw_b = 0, tw_c = 0, tw_d = 0;
    auto timed_wait = [&](uint64_t* bar, uint32_t parity, int code, long long& acc_t) {
        if (tmr != nullptr) {
            const long long t = clock64();
            mbar_wait(bar, parity, code);
            acc_t += clock64() - t;
        } else {
            mbar_wait(bar, parity, code);
        }
    };

    if (warp == kEpiWarps && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(bfull, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 2 * (kEpiThreads / 32));
        }
        fence_barrier_init();
    } else if (warp == kEpiWarps + 1) {
        tmem_alloc_pair(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kEpiWarps) {
        // ===================== TMA producer (both CTAs; uniform loops, one elected lane issues) =====================
        const uint32_t bfull_l = mapa_shared(bfull, 0);
        if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(bfull, static_cast<uint32_t>(2 * p.taps * p.k_chunks * b_region));
            for (int s = 0; s < p.taps; ++s)
                for (int kc = 0; kc < p.k_chunks; ++kc)
                    tma_load_2d_pair(sB + (s * p.k_chunks + kc) * b_region, &tmB, bfull_l, kc * kChunkK,
                                     s * p.b_tap_rows + col0 + static_cast<int>(rank) * (n_mma >> 1));
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 0;
        for (int pt = ptile0; pt < num_ptiles; pt += ptile_step) {
            const int row0 = (2 * pt + static_cast<int>(rank)) * p.rows_per_tile;  // past M: zero filled
            for (int s = 0; s < p.taps; ++s)
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    timed_wait(&empty[st], ph ^ 1, 101, tw_a);
                    if (elect_one()) {
#ifdef NEWSREC_TRIAGE
                        if (p.dbg_flags & 2) {
                            if (leader) mbar_arrive(&full[st]);
                        } else
#endif
                        {
                            if (leader) mbar_arrive_expect_tx(&full[st], 2 * kAStageBytes);
                            tma_load_2d_pair(sA + st * kAStageBytes, &tmA, mapa_shared(&full[st], 0), kc * kChunkK,
                                             row0 + s - tap_shift);
                        }
                    }
                    __syncwarp();
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
        }
        if (tmr != nullptr && lane == 0) tmr[0] = tw_a;
    } else if (warp == kEpiWarps + 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        // The whole warp runs the loops (uniform control flow) and one elected lane issues: under an `if (lane == 0)`
        // the compiler wraps every tcgen05 instruction in a uniform-register waterfall loop, and together with the
        // run-time k-step count that made the issuing thread -- not the tensor pipe -- the limiter (~240 cycles per MMA
        // against the pipe's 120 even with loads and epilogue switched off).  Columns past K are zero filled by TMA in
        // both operands, so every k-chunk issues all four k-steps.
        if (leader) {
            const uint32_t idesc = make_idesc_bf16(2 * kTileM, n_mma, 0, 0);
            mbar_wait(bfull, 0, 102);
            tc_fence_after();
            int st = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int pt = ptile0; pt < num_ptiles; pt += ptile_step, ++it) {
                const int as = it & 1;
                timed_wait(&tempty[as], ((it >> 1) & 1) ^ 1, 103, tw_b);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                uint32_t acc = 0;
                for (int s = 0; s < p.taps; ++s)
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        timed_wait(&full[st], ph, 104, tw_a);
                        tc_fence_after();
                        const long long t_i0 = tmr != nullptr ? clock64() : 0;
                        if (elect_one()) {
                            const uint64_t da = make_sw128_desc(smem_u32(sA + st * kAStageBytes), 0, 1024);
                            const uint64_t db = make_sw128_desc(smem_u32(sB + (s * p.k_chunks + kc) * b_region), 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)  // +32 bytes per k-step = +2 in the descriptor's >>4 address field
                                umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (k == 0) ? acc : 1u);
                            umma_commit_pair(&empty[st]);  // frees stage st in both CTAs when these MMAs retire
                        }
                        __syncwarp();
                        acc = 1;
                        if (tmr != nullptr) tw_c += clock64() - t_i0;
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                    }
                if (elect_one()) umma_commit_pair(&tfull[as]);
                __syncwarp();
            }
            if (tmr != nullptr && lane == 0) { tmr[1] = tw_a; tmr[2] = tw_b; tmr[7] = tw_c; tmr[8] = tw_d; }
        }
    } else {
        // ===================== epilogue warps 0..kEpiWarps-1 (both CTAs, own TMEM rows) =====================
        const int tile_step = 2 * ptile_step;
        const EpiInit ei{col0, ncols, static_cast<int>(threadIdx.x), scratch, 2 * ptile0 + static_cast<int>(rank), p.num_m_tiles};
        epi.init(ei, tile_step);
        const int quarter = warp & 3;
        const uint32_t tempty_l0 = mapa_shared(&tempty[0], 0), tempty_l1 = mapa_shared(&tempty[1], 0);
        int it = 0;
        for (int pt = ptile0; pt < num_ptiles; pt += ptile_step, ++it) {
            const int tile = 2 * pt + static_cast<int>(rank);  // may be one past the last tile: every row invalid
            const int as = it & 1;
            timed_wait(&tfull[as], (it >> 1) & 1, 105, tw_a);
            tc_fence_after();
            const long long t_epi = tmr != nullptr ? clock64() : 0;
            EpiCtx c;
            c.tile = tile;
            c.r = quarter * 32 + lane;
            c.grow = tile * p.rows_per_tile + c.r;
            c.valid = (c.r < p.rows_per_tile) && (c.grow < p.M);
            c.col0 = col0;
            c.ncols = ncols;
            c.tid = threadIdx.x;
            c.half = warp >> 2;
            epi_chunk_range(ncols, c.half, c.ch0, c.ch1);
            c.scratch = scratch;
            c.it = it;
            c.next_tile = tile + tile_step < p.num_m_tiles ? tile + tile_step : -1;
            TmemAcc acc{tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * 256, as ? tempty_l1 : tempty_l0};
            epi(acc, c);
            if (tmr != nullptr) tw_b += clock64() - t_epi;
        }
        epi.finish(ei);
        if (tmr != nullptr && threadIdx.x == 0) { tmr[3] = tw_a; tmr[4] = tw_b; tmr[5] = clock64() - t_begin; tmr[6] = it; }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or frees TMEM) while the partner may still read its shared memory / signal it
    if (warp == kEpiWarps + 1) tmem_dealloc_pair(tmem_base, 512);
}

// Debug backend (TRIAGE builds only -- `make TRIAGE=1`, -DNEWSREC_TRIAGE; the release library has no second backend and
// consults no environment switch on the launch path): plain SIMT accumulate + the SAME epilogue functors.
#ifdef NEWSREC_TRIAGE
__global__ void gemm_nt_simt_acc_kernel(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                                        GemmNTParams p);
template <class Epi>
__global__ void __launch_bounds__(kEpiThreads, 1) gemm_nt_simt_epi_kernel(const GemmNTParams p, const __grid_constant__ Epi epi) {
    extern __shared__ __align__(1024) float scratch[];  // Epi::kScratchBytes
    const int slice = blockIdx.x % p.n_slices;
    const int tile0 = blockIdx.x / p.n_slices;
    const int tile_step = gridDim.x / p.n_slices;
    const int col0 = slice * p.n_stride;
    const int ncols = min(p.n_stride, p.N - col0);
    for (int i = threadIdx.x; i < Epi::kScratchBytes / 4; i += kEpiThreads) scratch[i] = 0.f;
    __syncthreads();
    const EpiInit ei{col0, ncols, static_cast<int>(threadIdx.x), scratch, tile0, p.num_m_tiles};
    epi.init(ei, tile_step);
    int it = 0;
    for (int tile = tile0; tile < p.num_m_tiles; tile += tile_step, ++it) {
        EpiCtx c;
        c.tile = tile;
        c.r = threadIdx.x & 127;
        c.grow = tile * p.rows_per_tile + c.r;
        c.valid = (c.r < p.rows_per_tile) && (c.grow < p.M);
        c.col0 = col0;
        c.ncols = ncols;
        c.tid = threadIdx.x;
        c.half = threadIdx.x >> 7;
        epi_chunk_range(ncols, c.half, c.ch0, c.ch1);
        c.scratch = scratch;
        c.it = it;
        c.next_tile = tile + tile_step < p.num_m_tiles ? tile + tile_step : -1;
        GlobalAcc acc{p.dbg_acc + (static_cast<size_t>(tile) * 128 + c.r) * p.dbg_ld + col0};
        epi(acc, c);
    }
    epi.finish(ei);
}
#endif  // NEWSREC_TRIAGE

// ---------------------------------------------------------------------------------------------
// gemm_tn kernel
// ---------------------------------------------------------------------------------------------
struct GemmTNParams {
    int Kr;          // reduction rows (tokens)
    int Ma;          // output rows  = columns of A
    int Nb;          // output cols  = columns of B used (<=512)
    int b_col0;      // first B column
    int b_row_shift; // B row = A row + shift (conv taps)
    int m_tiles;
    int k_slices;
    int chunks_per_slice;  // 64-row chunks per CTA
    int n_boxes;     // ceil(Nb/64)
    int stages;
    float* D;        // fp32 [Ma][ldd], accumulated with red.global.add
    int ldd;
};

__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTNParams p);

// ---------------------------------------------------------------------------------------------
// host-side launch helpers (gemm.cu)
// ---------------------------------------------------------------------------------------------
struct GemmNTPlan {
    GemmNTParams p;
    CUtensorMap tmA, tmB;
    int grid;
    size_t smem;
};
// Fills slices/boxes/stages and encodes the tensor maps.  A: [M rows][K] pitch lda; B: [taps*b_tap_rows][K] pitch ldb.
// scratch_bytes = Epi::kScratchBytes; max_n_stride (0 = none) caps the slice width for epilogues whose staging
// grows with it.
int plan_gemm_nt(GemmNTPlan* plan, const void* A, int M, int lda, const void* B, int N, int ldb, int K, int taps,
                 int b_tap_rows, int rows_per_tile, int num_sms, int max_slices, int scratch_bytes, int max_n_stride);
bool debug_simt_gemm();

template <class Epi>
int launch_gemm_nt(const GemmNTPlan& plan, const Epi& epi, const void* A, int lda, const void* B, int ldb,
                   cudaStream_t stream) {
    if (plan.p.num_m_tiles <= 0) return 0;
#ifdef NEWSREC_TRIAGE
    if (debug_simt_gemm()) {
        GemmNTParams p = plan.p;
        const size_t ld = static_cast<size_t>(round_up(p.N, 32) + 32);
        float* acc = nullptr;
        NR_CHECK_CUDA(cudaMallocAsync(&acc, sizeof(float) * ld * p.num_m_tiles * 128, stream));
        p.dbg_acc = acc;
        p.dbg_ld = static_cast<int>(ld);
        dim3 g(ceil_div(static_cast<int>(ld), 128), p.num_m_tiles);
        gemm_nt_simt_acc_kernel<<<g, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(A), lda,
                                                       static_cast<const __nv_bfloat16*>(B), ldb, p);
        static bool dbg_attr_set = false;  // per Epi instantiation
        if (!dbg_attr_set) {
            NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_simt_epi_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               Epi::kScratchBytes));
            dbg_attr_set = true;
        }
        gemm_nt_simt_epi_kernel<Epi><<<plan.grid, kEpiThreads, Epi::kScratchBytes, stream>>>(p, epi);
        NR_CHECK_CUDA(cudaGetLastError());
        NR_CHECK_CUDA(cudaFreeAsync(acc, stream));
        return 0;
    }
#endif
    static bool attr_set = false;  // per Epi instantiation
    if (!attr_set) {
        NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    gemm_nt_kernel<Epi><<<plan.grid, kGemmThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.p, epi);
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// D[Ma x Nb] (+)= A[:, 0:Ma]^T . B[shifted rows, b_col0 : b_col0+Nb]
int launch_gemm_tn(const void* A, int Kr, int Ma, int lda, const void* B, int b_rows, int b_cols, int ldb, int b_col0,
                   int Nb, int b_row_shift, float* D, int ldd, int num_sms, cudaStream_t stream);

int num_sms();
extern int g_launches;  // kernels launched by this library (bench.py reports it)

}  // namespace nr
