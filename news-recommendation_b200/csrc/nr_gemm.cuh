// Weight-stationary persistent tcgen05 GEMMs for the news-recommendation hot path (sm_100a).
//
//  gemm_nt : D[M x N] = A[M x K] . B[N x K]^T      (both K-major; "activation x weight^T")
//            * the CTA's weight slice (<=256 output columns, all K, all conv taps) is loaded ONCE by
//              TMA and stays resident in shared memory; 128-row activation tiles stream through a
//              TMA/mbarrier ring; accumulators are double buffered in TMEM (2 x 256 fp32 columns);
//              4 epilogue warps run a fused epilogue functor on tcgen05.ld'ed rows.
//            * conv taps: tap s re-loads the A tile shifted by (s - taps/2) rows (zero rows separate
//              the segments in the padded layout), accumulating into the same TMEM tile.
//  gemm_tn : D[Ma x Nb] += A[Kr x Ma]^T . B[Kr x Nb]  (both MN-major; weight gradients, Kr = all tokens)
//            split over Kr across CTAs, fp32 red.global.add epilogue.
//
// Warp roles of gemm_nt (320 threads): warps 0-7 epilogue (TMEM lane quarter = warp & 3; the two warps of a quarter
// split the accumulator columns -- with 4 warps the row-per-thread epilogue was latency bound at ~20 % issue
// utilisation, ncu profiles/), warp 8 TMA producer, warp 9 MMA issuer + TMEM allocator.
// gemm_tn keeps 192 threads (4 epilogue warps, one-shot epilogue).
#pragma once
#include "nr_common.cuh"

namespace nr {

constexpr int kGemmThreads = 320;   // gemm_nt
constexpr int kTnThreads = 192;     // gemm_tn
constexpr int kEpiThreads = 256;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                     // bf16 elements per 128-byte swizzle row
constexpr int kAStageBytes = kTileM * 128;      // 16 KB
constexpr int kMaxStages = 8;
constexpr int kEpiScratchBytes = 10240;
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
    int half;     // 0 / 1: which of the two warps of this TMEM lane quarter
    int ch0, ch1; // this thread's range of 32-column chunks
    float* scratch;  // kEpiScratchBytes of shared memory private to the epilogue group
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// the two column halves run different chunk counts: barriers inside a chunk loop are per half
__device__ __forceinline__ void epi_bar_sync_half(int half) { asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory"); }
__device__ __forceinline__ void epi_chunk_range(int ncols, int half, int& ch0, int& ch1) {
    const int nch = (ncols + 31) >> 5, mid = (nch + 1) >> 1;
    ch0 = half ? mid : 0;
    ch1 = half ? nch : mid;
}

struct TmemAcc {
    uint32_t taddr;
    uint64_t* release_bar;
    __device__ __forceinline__ void load32(int chunk, float* v) const {
        tmem_ld32(taddr + chunk * 32, v);
        tmem_ld_wait();
    }
    __device__ __forceinline__ void release() const {  // all TMEM reads of this tile by this thread are done
        tc_fence_before();
        mbar_arrive(release_bar);
    }
};
struct GlobalAcc {  // debug backend: accumulators computed by a plain SIMT kernel
    const float* row;
    __device__ __forceinline__ void load32(int chunk, float* v) const {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = row[chunk * 32 + j];
    }
    __device__ __forceinline__ void release() const {}
};

// ---------------------------------------------------------------------------------------------
// gemm_nt kernel
// ---------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmNTParams p,
               const Epi epi) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int b_region = p.n_box * 128;  // bytes of one (tap, k-chunk) weight box
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

    const int slice = blockIdx.x % p.n_slices;
    const int tile0 = blockIdx.x / p.n_slices;
    const int tile_step = gridDim.x / p.n_slices;
    const int col0 = slice * p.n_stride;
    const int ncols = min(p.n_stride, p.N - col0);
    const int n_mma = (ncols + 15) & ~15;
    const int tap_shift = p.taps / 2;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(bfull, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], kEpiThreads);
        }
        fence_barrier_init();
    } else if (warp == 9) {
        tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(bfull, static_cast<uint32_t>(p.taps * p.k_chunks * b_region));
            for (int s = 0; s < p.taps; ++s)
                for (int kc = 0; kc < p.k_chunks; ++kc)
                    tma_load_2d(sB + (s * p.k_chunks + kc) * b_region, &tmB, bfull, kc * kChunkK,
                                s * p.b_tap_rows + col0);
            int st = 0;
            uint32_t ph = 0;
            for (int tile = tile0; tile < p.num_m_tiles; tile += tile_step) {
                const int row0 = tile * p.rows_per_tile;
                for (int s = 0; s < p.taps; ++s)
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&empty[st], ph ^ 1, 101);
                        mbar_arrive_expect_tx(&full[st], kAStageBytes);
                        tma_load_2d(sA + st * kAStageBytes, &tmA, &full[st], kc * kChunkK, row0 + s - tap_shift);
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(kTileM, n_mma, 0, 0);
            mbar_wait(bfull, 0, 102);
            tc_fence_after();
            int st = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int tile = tile0; tile < p.num_m_tiles; tile += tile_step, ++it) {
                const int as = it & 1;
                mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1, 103);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                uint32_t acc = 0;
                for (int s = 0; s < p.taps; ++s)
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&full[st], ph, 104);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sA + st * kAStageBytes);
                        const uint32_t b_addr = smem_u32(sB + (s * p.k_chunks + kc) * b_region);
                        const int ksteps = min(4, (p.K - kc * kChunkK + 15) >> 4);
                        for (int k = 0; k < ksteps; ++k) {
                            umma_bf16(d_tmem, make_sw128_desc(a_addr + k * 32, 0, 1024),
                                      make_sw128_desc(b_addr + k * 32, 0, 1024), idesc, acc);
                            acc = 1;
                        }
                        umma_commit(&empty[st]);  // frees the A stage when these MMAs retire
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                    }
                umma_commit(&tfull[as]);
            }
        }
    } else {
        // ===================== epilogue warps 0..7 =====================
        epi.init(col0, ncols, threadIdx.x, scratch);
        const int quarter = warp & 3;
        int it = 0;
        for (int tile = tile0; tile < p.num_m_tiles; tile += tile_step, ++it) {
            const int as = it & 1;
            mbar_wait(&tfull[as], (it >> 1) & 1, 105);
            tc_fence_after();
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
            TmemAcc acc{tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * 256, &tempty[as]};
            epi(acc, c);
        }
        epi.finish(col0, ncols, threadIdx.x, scratch);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// Debug backend (triage only, NR_DEBUG_SIMT_GEMM=1): plain SIMT accumulate + the SAME epilogue functors.
__global__ void gemm_nt_simt_acc_kernel(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                                        GemmNTParams p);
template <class Epi>
__global__ void __launch_bounds__(kEpiThreads, 1) gemm_nt_simt_epi_kernel(const GemmNTParams p, const Epi epi) {
    __shared__ float scratch[kEpiScratchBytes / 4];
    const int slice = blockIdx.x % p.n_slices;
    const int tile0 = blockIdx.x / p.n_slices;
    const int tile_step = gridDim.x / p.n_slices;
    const int col0 = slice * p.n_stride;
    const int ncols = min(p.n_stride, p.N - col0);
    for (int i = threadIdx.x; i < kEpiScratchBytes / 4; i += kEpiThreads) scratch[i] = 0.f;
    __syncthreads();
    epi.init(col0, ncols, threadIdx.x, scratch);
    for (int tile = tile0; tile < p.num_m_tiles; tile += tile_step) {
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
        GlobalAcc acc{p.dbg_acc + (static_cast<size_t>(tile) * 128 + c.r) * p.dbg_ld + col0};
        epi(acc, c);
    }
    epi.finish(col0, ncols, threadIdx.x, scratch);
}

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
int plan_gemm_nt(GemmNTPlan* plan, const void* A, int M, int lda, const void* B, int N, int ldb, int K, int taps,
                 int b_tap_rows, int rows_per_tile, int num_sms, int max_slices);
bool debug_simt_gemm();

template <class Epi>
int launch_gemm_nt(const GemmNTPlan& plan, const Epi& epi, const void* A, int lda, const void* B, int ldb,
                   cudaStream_t stream) {
    if (plan.p.num_m_tiles <= 0) return 0;
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
        gemm_nt_simt_epi_kernel<Epi><<<plan.grid, kEpiThreads, 0, stream>>>(p, epi);
        NR_CHECK_CUDA(cudaGetLastError());
        NR_CHECK_CUDA(cudaFreeAsync(acc, stream));
        return 0;
    }
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
