// tcgen05 GEMM building blocks + their fused-epilogue instantiations.  See nr_gemm.cuh for the design.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#define NR_OWNS_WATCHDOG 1
#include "nr_epilogues.cuh"
#include "nr_ops.h"

namespace nr {

int g_launches = 0;

// ------------------------------------------------------------------------------------------------
// error string (thread local) + device watchdog record
// ------------------------------------------------------------------------------------------------
static thread_local char t_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return t_err; }
int read_device_error(int* out4) {
    const int rc = (int)cudaMemcpyFromSymbol(out4, g_dev_error, sizeof(int) * 4);
    if (rc != 0 || out4[0] != 0) return rc;
    return read_fused_device_error(out4);  // the fused encoder kernels keep their own record (fused_fwd.cu)
}

// ------------------------------------------------------------------------------------------------
// live per-kernel timing
// ------------------------------------------------------------------------------------------------
struct ProfRec {
    std::string name;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::string g_prof_ctx;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_prof_pool.empty()) {
        e = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
void prof_enable(int on) { g_prof_on = on != 0; }
void prof_context(const char* ctx) { g_prof_ctx = ctx; }
// NEWSREC_TRACE=1: print every launch and synchronise after it (pin-points a stuck or faulting kernel).
static int trace_on() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NEWSREC_TRACE");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v;
}
ProfScope::ProfScope(const char* op, int a, int b, int c, cudaStream_t s) : idx(-1), stream(s) {
    if (trace_on()) {
        fprintf(stderr, "[nr] launch %s/%s[%d,%d,%d]\n", g_prof_ctx.c_str(), op, a, b, c);
        fflush(stderr);
    }
    if (!g_prof_on) return;
    char buf[160];
    snprintf(buf, sizeof(buf), "%s/%s[%d,%d,%d]", g_prof_ctx.c_str(), op, a, b, c);
    ProfRec r{buf, prof_event(), prof_event()};
    cudaEventRecord(r.a, s);
    idx = static_cast<int>(g_prof.size());
    g_prof.push_back(r);
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof[idx].b, stream);
    if (trace_on()) {
        const cudaError_t e = cudaStreamSynchronize(stream);
        fprintf(stderr, "[nr]   -> %s\n", cudaGetErrorString(e));
        fflush(stderr);
    }
}
int prof_report(char* buf, int cap) {
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& e = agg[r.name];
            e.first += 1;
            e.second += ms;
        }
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": [%d, %.6f]", first ? "" : ", ", kv.first.c_str(), kv.second.first,
                 kv.second.second);
        out += line;
        first = false;
    }
    out += "}";
    if (static_cast<int>(out.size()) + 1 > cap) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return static_cast<int>(out.size());
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 0;
        if (const char* e = getenv("NEWSREC_NUM_SMS")) n = atoi(e);
    }
    return n;
}

#ifdef NEWSREC_TRIAGE
static int g_debug_simt = -1;
bool debug_simt_gemm() {
    if (g_debug_simt < 0) {
        const char* e = getenv("NEWSREC_DEBUG_SIMT_GEMM");
        g_debug_simt = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return g_debug_simt == 1;
}
void set_debug_simt_gemm(int on) { g_debug_simt = on ? 1 : 0; }
int has_triage_backends() { return 1; }
#else
bool debug_simt_gemm() { return false; }
void set_debug_simt_gemm(int) {}
int has_triage_backends() { return 0; }
#endif
// tuning (tools/kbench.py): device buffer [slots][148][16]; every planned gemm_nt takes the next slot
static long long* g_gemm_timing = nullptr;
static int g_gemm_timing_slots = 0, g_gemm_timing_next = 0;
void set_debug_gemm_timing(void* dev_buf, int slots) {
    g_gemm_timing = static_cast<long long*>(dev_buf);
    g_gemm_timing_slots = slots;
    g_gemm_timing_next = 0;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding via the driver entry point (resolved at run time: the library must load
// on a machine without libcuda so that the CPU-side symbol tests can dlopen it)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld_elems, int box_cols,
                      int box_rows, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    NR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    NR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16B aligned", base);
    NR_REQUIRE((ld_elems * 2) % 16 == 0, "TMA row pitch %lld elements is not a multiple of 16 bytes", (long long)ld_elems);
    NR_REQUIRE(((swizzle_bytes == 128 || swizzle_bytes == 64) ? box_cols * 2 <= swizzle_bytes
                                                               : (swizzle_bytes == 0 && (box_cols * 2) % 16 == 0 && box_cols <= 256)) &&
                   box_rows <= 256 && box_rows >= 1,
               "bad TMA box %d x %d (swizzle %d)", box_cols, box_rows, swizzle_bytes);
    NR_REQUIRE(rows >= 1 && cols >= 1, "empty tensor for TMA (%lld x %lld)", (long long)rows, (long long)cols);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%dx%d)",
               (int)r, (long long)rows, (long long)cols, (long long)ld_elems, box_cols, box_rows);
    return 0;
}

// Untyped 2-D tensor map: rows of `row_bytes` valid bytes at pitch `pitch_bytes`, moved as 8-byte elements (so that a box
// can be up to 2 KB wide), box = box_bytes x box_rows, no swizzle.  box_bytes may exceed row_bytes: the tail is zero
// filled on loads and dropped on stores (it lets the caller choose a bank-conflict-free row pitch in shared memory).
int make_tmap_bytes_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t row_bytes, int64_t pitch_bytes, int box_bytes,
                       int box_rows) {
    EncodeTiledFn enc = get_encode();
    NR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    NR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch_bytes % 16 == 0 && row_bytes % 8 == 0 && row_bytes >= 8,
               "TMA bytes map: base %p / pitch %lld / row %lld misaligned", base, (long long)pitch_bytes, (long long)row_bytes);
    NR_REQUIRE(box_bytes % 16 == 0 && box_bytes >= 16 && box_bytes <= 2048 && box_rows >= 1 && box_rows <= 256 && rows >= 1,
               "TMA bytes map: bad box %d B x %d", box_bytes, box_rows);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(row_bytes / 8), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / 8), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (bytes) failed with CUresult %d (rows=%lld row=%lld pitch=%lld box=%dx%d)", (int)r,
               (long long)rows, (long long)row_bytes, (long long)pitch_bytes, box_bytes, box_rows);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// gemm_nt planning
// ------------------------------------------------------------------------------------------------
int plan_gemm_nt(GemmNTPlan* plan, const void* A, int M, int lda, const void* B, int N, int ldb, int K, int taps,
                 int b_tap_rows, int rows_per_tile, int sms, int max_slices, int scratch_bytes, int max_n_stride) {
    NR_REQUIRE(M >= 0 && N >= 1 && K >= 1 && taps >= 1 && rows_per_tile >= 1 && rows_per_tile <= kTileM,
               "plan_gemm_nt: bad shape M=%d N=%d K=%d taps=%d rpt=%d", M, N, K, taps, rows_per_tile);
    NR_REQUIRE(sms > 0, "no CUDA device (SM count unknown)");
    GemmNTParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    p.M = M;
    p.rows_per_tile = rows_per_tile;
    p.num_m_tiles = ceil_div(M, rows_per_tile);
    p.N = N;
    p.K = K;
    p.k_chunks = ceil_div(K, kChunkK);
    p.taps = taps;
    p.b_tap_rows = b_tap_rows;
#ifdef NEWSREC_TRIAGE
    static const int dbg_flags = [] { const char* v = getenv("NEWSREC_GEMM_DBG"); return v != nullptr ? atoi(v) : 0; }();
    p.dbg_flags = dbg_flags;
#endif
    if (g_gemm_timing != nullptr && g_gemm_timing_next < g_gemm_timing_slots) {
        p.timing = g_gemm_timing + static_cast<size_t>(g_gemm_timing_next) * 148 * 16;
        fprintf(stderr, "[nr] gemm timing slot %d: M=%d N=%d K=%d taps=%d\n", g_gemm_timing_next, M, N, K, taps);
        ++g_gemm_timing_next;
    }
    const int fixed = 1024 + round_up(scratch_bytes, 16) + 512;
    int slices = 1;
    for (;; ++slices) {
        NR_REQUIRE(slices <= 64, "plan_gemm_nt: cannot fit weight slice (N=%d K=%d taps=%d)", N, K, taps);
        p.n_stride = round_up(ceil_div(N, slices), 16);  // 32-byte aligned slice starts (STG.256 epilogues)
        p.n_box = round_up(std::min(p.n_stride, N), 16);
        if (p.n_box > 256 || (max_n_stride > 0 && p.n_stride > max_n_stride)) continue;
        const long bbytes = static_cast<long>(taps) * p.k_chunks * (p.n_box / 2) * 128;  // per CTA: half the slice
        if (bbytes + 4L * kAStageBytes + fixed <= kSmemLimit) break;
        // epilogues that reduce over the whole output row need ONE slice: accept a shallower A ring instead
        if (slices == max_slices && bbytes + 2L * kAStageBytes + fixed <= kSmemLimit) break;
    }
    p.n_slices = ceil_div(N, p.n_stride);
    const long bbytes = static_cast<long>(taps) * p.k_chunks * (p.n_box / 2) * 128;
    p.stages = static_cast<int>(std::min<long>(kMaxStages, (kSmemLimit - fixed - bbytes) / kAStageBytes));
    NR_REQUIRE(p.stages >= 2, "plan_gemm_nt: only %d pipeline stages fit", p.stages);
    plan->smem = static_cast<size_t>(bbytes) + static_cast<size_t>(p.stages) * kAStageBytes + fixed;
    // CTA pairs: a group = n_slices pairs working on the same 256-row blocks
    const int groups = std::max(1, std::min((sms / 2) / p.n_slices, ceil_div(p.num_m_tiles, 2)));
    plan->grid = 2 * groups * p.n_slices;
    if (p.num_m_tiles == 0) return 0;
    NR_PROPAGATE(make_tmap_bf16_2d(&plan->tmA, A, M, K, lda, kChunkK, kTileM));
    const int64_t brows = (taps > 1) ? static_cast<int64_t>(taps) * b_tap_rows : N;
    NR_PROPAGATE(make_tmap_bf16_2d(&plan->tmB, B, brows, K, ldb, kChunkK, p.n_box / 2));
    return 0;
}

#ifdef NEWSREC_TRIAGE
// Debug backend accumulate: thread = one output column of one M tile (slow, obviously correct).
__global__ void gemm_nt_simt_acc_kernel(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                                        GemmNTParams p) {
    const int n = blockIdx.x * 128 + threadIdx.x;
    const int tile = blockIdx.y;
    if (n >= p.dbg_ld) return;
    const int shift = p.taps / 2;
    for (int r = 0; r < 128; ++r) {
        float acc = 0.f;
        if (n < p.N) {
            for (int s = 0; s < p.taps; ++s) {
                const long long row = static_cast<long long>(tile) * p.rows_per_tile + r + s - shift;
                if (row < 0 || row >= p.M) continue;
                const __nv_bfloat16* a = A + row * lda;
                const __nv_bfloat16* b = B + static_cast<long long>(s * p.b_tap_rows + n) * ldb;
                for (int k = 0; k < p.K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(b[k]), acc);
            }
        }
        p.dbg_acc[(static_cast<size_t>(tile) * 128 + r) * p.dbg_ld + n] = acc;
    }
}

#endif  // NEWSREC_TRIAGE

// ------------------------------------------------------------------------------------------------
// gemm_tn kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTNParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stage_bytes = (2 + p.n_boxes) * 8192;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* tfull = bars + 2 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);

    const int mt = blockIdx.x % p.m_tiles;
    const int ks = blockIdx.x / p.m_tiles;
    const int total_chunks = (p.Kr + 63) >> 6;
    const int chunk0 = ks * p.chunks_per_slice;
    const int n_my = max(0, min(p.chunks_per_slice, total_chunks - chunk0));
    const int nb_pad = (p.Nb + 15) & ~15;
    // more than 256 columns take two MMAs per k-step: split them near the middle, on a 64-column box boundary -- a single-CTA
    // tcgen05.mma costs >= 86 cycles whatever its N (tools/mmabench_small.cu), so 256 + 48 is 128 + 86 cycles where 192 + 112
    // is 96 + 86
    const int n0 = nb_pad <= 256 ? nb_pad : (((nb_pad >> 1) + 63) & ~63);
    const int n1 = nb_pad - n0;
    const uint32_t tmem_cols = nb_pad > 256 ? 512u : (nb_pad > 128 ? 256u : (nb_pad > 64 ? 128u : (nb_pad > 32 ? 64u : 32u)));

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    } else if (warp == 5) {
        tmem_alloc(tmem_slot, tmem_cols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (n_my > 0) {
        // producer and issuer loops are warp-uniform, one elected lane issues (see gemm_nt)
        if (warp == 4) {
            int st = 0;
            uint32_t ph = 0;
            for (int c = 0; c < n_my; ++c) {
                const int k0 = (chunk0 + c) * 64;
                mbar_wait(&empty[st], ph ^ 1, 201);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[st], static_cast<uint32_t>(stage_bytes));
                    uint8_t* sa = smem + st * stage_bytes;
                    uint8_t* sb = sa + 2 * 8192;
                    tma_load_2d(sa, &tmA, &full[st], mt * 128, k0);
                    tma_load_2d(sa + 8192, &tmA, &full[st], mt * 128 + 64, k0);
                    for (int j = 0; j < p.n_boxes; ++j)
                        tma_load_2d(sb + j * 8192, &tmB, &full[st], p.b_col0 + j * 64, k0 + p.b_row_shift);
                }
                __syncwarp();
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
        } else if (warp == 5) {
            const uint32_t idesc0 = make_idesc_bf16(kTileM, n0, 1, 1);
            const uint32_t idesc1 = make_idesc_bf16(kTileM, n1 > 0 ? n1 : 16, 1, 1);
            int st = 0;
            uint32_t ph = 0;
            uint32_t acc = 0;
            for (int c = 0; c < n_my; ++c) {
                mbar_wait(&full[st], ph, 202);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + st * stage_bytes);
                    const uint32_t sb = sa + 2 * 8192;
                    const uint64_t da = make_sw128_desc(sa, 8192, 1024);
                    const uint64_t db0 = make_sw128_desc(sb, 8192, 1024);
                    const uint64_t db1 = make_sw128_desc(sb + (n0 >> 6) * 8192, 8192, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // +2048 bytes per k-step = +128 in the descriptor's >>4 address field
                        umma_bf16(tmem_base, da + 128 * k, db0 + 128 * k, idesc0, (k == 0) ? acc : 1u);
                        if (n1 > 0) umma_bf16(tmem_base + n0, da + 128 * k, db1 + 128 * k, idesc1, (k == 0) ? acc : 1u);
                    }
                    umma_commit(&empty[st]);
                }
                __syncwarp();
                acc = 1;
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(tfull);
            __syncwarp();
        } else {
            mbar_wait(tfull, 0, 203);
            tc_fence_after();
            const int grow = mt * 128 + warp * 32 + lane;
            const bool valid = grow < p.Ma;
            float* drow = p.D + static_cast<size_t>(valid ? grow : 0) * p.ldd;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
            const int nch = (p.Nb + 31) >> 5;
            const bool vec4 = (p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.D) & 15) == 0;
            for (int ch = 0; ch < nch; ++ch) {
                float x[32];
                tmem_ld32(taddr + ch * 32, x);
                tmem_ld_wait();
                if (!valid) continue;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {  // 16-byte vector reductions: 4x fewer L2 atomic operations
                    const int col = ch * 32 + j;
                    if (vec4 && col + 4 <= p.Nb) {
                        red_add_v4_f32(drow + col, x[j], x[j + 1], x[j + 2], x[j + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (col + i < p.Nb) red_add_f32(drow + col + i, x[j + i]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, tmem_cols);
}

#ifdef NEWSREC_TRIAGE
__global__ void gemm_tn_simt_kernel(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int b_rows,
                                    GemmTNParams p) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (n >= p.Nb || m >= p.Ma) return;
    float acc = 0.f;
    for (int k = 0; k < p.Kr; ++k) {
        const long long br = static_cast<long long>(k) + p.b_row_shift;
        if (br < 0 || br >= b_rows) continue;
        acc = fmaf(__bfloat162float(A[static_cast<size_t>(k) * lda + m]),
                   __bfloat162float(B[static_cast<size_t>(br) * ldb + p.b_col0 + n]), acc);
    }
    atomicAdd(p.D + static_cast<size_t>(m) * p.ldd + n, acc);
}
#endif

// SMs the split-K weight-gradient GEMM leaves free (set by a data-parallel caller): the all-reduce of the embedding gradient
// is issued on a side stream right before this GEMM, and NCCL's channel CTAs need SMs to run on -- with every SM holding a
// 200 KB gemm_tn CTA the "overlapped" reduction simply waited for the GEMM to finish (2 GPUs: +0.27 ms per step, unchanged).
static int g_comm_reserved_sms = 0;
void set_comm_reserved_sms(int n) { g_comm_reserved_sms = n < 0 ? 0 : n; }

int gemm_tn_accumulate(const void* A, int Kr, int Ma, int lda, const void* B, int b_rows, int b_cols, int ldb,
                       int b_col0, int Nb, int b_row_shift, float* D, int ldd, cudaStream_t stream) {
    NR_REQUIRE(Nb >= 1 && Nb <= 512 && Ma >= 1 && Kr >= 0, "gemm_tn: bad shape Kr=%d Ma=%d Nb=%d", Kr, Ma, Nb);
    if (Kr == 0) return 0;
    GemmTNParams p;
    memset(&p, 0, sizeof(p));
    p.Kr = Kr;
    p.Ma = Ma;
    p.Nb = Nb;
    p.b_col0 = b_col0;
    p.b_row_shift = b_row_shift;
    p.D = D;
    p.ldd = ldd;
    ProfScope ps("gemm_tn", Kr, Ma, Nb, stream);
#ifdef NEWSREC_TRIAGE
    if (debug_simt_gemm()) {
        dim3 g(ceil_div(Nb, 64), Ma);
        gemm_tn_simt_kernel<<<g, 64, 0, stream>>>(static_cast<const __nv_bfloat16*>(A), lda,
                                                  static_cast<const __nv_bfloat16*>(B), ldb, b_rows, p);
        ++g_launches;
        NR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
#endif
    NR_REQUIRE(num_sms() > 0, "no CUDA device");
    const int sms = std::max(num_sms() / 2, num_sms() - g_comm_reserved_sms);
    p.m_tiles = ceil_div(Ma, 128);
    const int total_chunks = ceil_div(Kr, 64);
    int k_slices = std::max(1, std::min(sms / p.m_tiles, total_chunks));
    p.chunks_per_slice = ceil_div(total_chunks, k_slices);
    k_slices = ceil_div(total_chunks, p.chunks_per_slice);
    p.k_slices = k_slices;
    p.n_boxes = ceil_div(Nb, 64);
    const int stage_bytes = (2 + p.n_boxes) * 8192;
    p.stages = std::min(kMaxStages, (kSmemLimit - 1024 - 512) / stage_bytes);
    NR_REQUIRE(p.stages >= 2, "gemm_tn: stage of %d bytes does not double-buffer", stage_bytes);
    p.stages = std::min(p.stages, 6);
    CUtensorMap tmA, tmB;
    NR_PROPAGATE(make_tmap_bf16_2d(&tmA, A, Kr, Ma, lda, 64, 64));
    NR_PROPAGATE(make_tmap_bf16_2d(&tmB, B, b_rows, b_cols, ldb, 64, 64));
    static bool attr_set = false;
    if (!attr_set) {
        NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 1024 + 512;
    gemm_tn_kernel<<<p.m_tiles * k_slices, kTnThreads, smem, stream>>>(tmA, tmB, p);
    ++g_launches;
    NR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// epilogue instantiations
// ------------------------------------------------------------------------------------------------
static RowMap to_rm(const RowMapCfg& c) { return RowMap{c.seg_in, c.in_off, c.seg_len, c.seg_out, c.out_off}; }
static Dropout to_drop(const DropoutCfg& c) {
    Dropout d;
    d.p = c.p;
    d.scale = c.p > 0.f ? 1.f / (1.f - c.p) : 1.f;
    d.thresh = static_cast<uint32_t>(c.p * 65536.0f + 0.5f);
    d.seed = c.seed;
    return d;
}

int gemm_store(const void* A, int M, int lda, const void* W, int N, int ldw, int K, int taps, int w_tap_rows,
               int rows_per_tile, const float* bias, int relu, void* out, int ld_out, int out_bf16, RowMapCfg rm,
               int zero_pad_rows, DropoutCfg drop, int ones_col, int ones_zero_upto, cudaStream_t stream, void* lo_out, int ld_lo,
               int lo_col0, int accumulate) {
    if (M == 0) return 0;
    GemmNTPlan plan;
    NR_PROPAGATE(plan_gemm_nt(&plan, A, M, lda, W, N, ldw, K, taps, w_tap_rows, rows_per_tile, num_sms(), 0, EpiStore::kScratchBytes, 0));
    NR_REQUIRE(out_bf16 ? (ld_out % 8 == 0) : (ld_out % 4 == 0), "gemm_store: output pitch %d breaks vector stores", ld_out);
    EpiStore e;
    memset(&e, 0, sizeof(e));
    e.use_tma = (out_bf16 && rm.seg_in == 0 && rows_per_tile == kTileM && N >= 32) ? 1 : 0;
    if (e.use_tma) NR_PROPAGATE(make_tmap_bf16_2d(&e.tm_out, out, M, N, ld_out, 32, 32, 64));
    e.lo_col0 = -1;
    NR_REQUIRE(!accumulate || !out_bf16, "gemm_store: accumulation needs an fp32 output");
    e.accumulate = accumulate;
    if (lo_out != nullptr) {
        NR_REQUIRE(out_bf16 && lo_col0 >= 0 && lo_col0 < N && ld_lo % 8 == 0 && ld_lo >= N - lo_col0 && (e.use_tma || lo_col0 % 8 == 0),
                   "gemm_store: the low plane needs bf16 output and aligned columns (N=%d lo_col0=%d ld_lo=%d)", N, lo_col0, ld_lo);
        for (int sl = 0; sl < plan.p.n_slices; ++sl) {  // a 32-column chunk never straddles the first low-plane column
            const int c0 = sl * plan.p.n_stride;
            NR_REQUIRE(!(c0 < lo_col0 && lo_col0 < c0 + plan.p.n_stride) || (lo_col0 - c0) % 32 == 0,
                       "gemm_store: low-plane start %d is not chunk aligned in the slice at column %d", lo_col0, c0);
        }
        if (e.use_tma) NR_PROPAGATE(make_tmap_bf16_2d(&e.tm_lo, lo_out, M, N - lo_col0, ld_lo, 32, 32, 64));
        e.lo_col0 = lo_col0;
        e.lo_out = static_cast<__nv_bfloat16*>(lo_out);
        e.ld_lo = ld_lo;
    }
    e.out = out;
    e.ld = ld_out;
    e.out_bf16 = out_bf16;
    e.bias = bias;
    e.relu = relu;
    e.N = N;
    e.rm = to_rm(rm);
    e.zero_pad_rows = zero_pad_rows;
    e.drop = to_drop(drop);
    e.ones_col = ones_col;
    e.ones_cols_zero_upto = ones_zero_upto;
#ifdef NEWSREC_TRIAGE
    static const int dbg_skip = [] { const char* v = getenv("NEWSREC_EPI_DBG"); return v != nullptr && v[0] == '1' ? 1 : 0; }();
    e.dbg_skip = dbg_skip;
#endif
    g_launches += debug_simt_gemm() ? 2 : 1;
    ProfScope ps("gemm_store", M, N, K * taps, stream);
    return launch_gemm_nt(plan, e, A, lda, W, ldw, stream);
}

int gemm_additive_pool(const void* X, int M, int lda, int D, const void* Wa, int q, int ldw, const float* ba,
                       const float* qv, int seg_len, float* out, int ldo, float* w_out, cudaStream_t stream, const void* X_lo) {
    if (M == 0) return 0;
    NR_REQUIRE(seg_len >= 1 && seg_len <= kTileM && M % seg_len == 0, "additive_pool: M=%d seg_len=%d", M, seg_len);
    NR_REQUIRE(q <= 256 && (D % 2) == 0 && (ldo % 2) == 0, "additive_pool: q=%d D=%d ldo=%d unsupported", q, D, ldo);
    const int rpt = (kTileM / seg_len) * seg_len;
    GemmNTPlan plan;
    NR_PROPAGATE(plan_gemm_nt(&plan, X, M, lda, Wa, q, ldw, D, 1, 0, rpt, num_sms(), 1, EpiPool::kScratchBytes, 0));
    NR_REQUIRE(plan.p.n_slices == 1, "additive_pool: the query dimension must fit one weight slice (q=%d D=%d)", q, D);
    EpiPool e;
    e.bias = ba;
    e.qv = qv;
    e.X = static_cast<const __nv_bfloat16*>(X);
    e.X_lo = static_cast<const __nv_bfloat16*>(X_lo);
    e.lda = lda;
    e.D = D;
    e.seg_len = seg_len;
    e.rows_per_tile = rpt;
    e.M = M;
    e.out = out;
    e.ldo = ldo;
    e.w_out = w_out;
    g_launches += debug_simt_gemm() ? 2 : 1;
    ProfScope ps("gemm_additive_pool", M, q, D, stream);
    return launch_gemm_nt(plan, e, X, lda, Wa, ldw, stream);
}

int gemm_additive_dpre(const void* X, int M, int lda, int D, const void* Wa, int q, int ldw, const float* ba,
                       const float* qv, const float* dscore, void* dpre, int ld_dpre, float* dqv,
                       cudaStream_t stream) {
    if (M == 0) return 0;
    NR_REQUIRE(q <= 256 && ld_dpre % 8 == 0 && ld_dpre >= round_up(q, 8), "additive_dpre: q=%d ld=%d", q, ld_dpre);
    GemmNTPlan plan;
    NR_PROPAGATE(plan_gemm_nt(&plan, X, M, lda, Wa, q, ldw, D, 1, 0, kTileM, num_sms(), 1, EpiDPre::kScratchBytes, 0));
    NR_REQUIRE(plan.p.n_slices == 1, "additive_dpre: q=%d D=%d does not fit one weight slice", q, D);
    EpiDPre e;
    memset(&e, 0, sizeof(e));
    e.use_tma = q >= 32 ? 1 : 0;
    if (e.use_tma) NR_PROPAGATE(make_tmap_bf16_2d(&e.tm_out, dpre, M, ld_dpre, ld_dpre, 32, 32, 64));
    e.bias = ba;
    e.qv = qv;
    e.dscore = dscore;
    e.dpre = static_cast<__nv_bfloat16*>(dpre);
    e.ld = ld_dpre;
    e.dqv = dqv;
    g_launches += debug_simt_gemm() ? 2 : 1;
    ProfScope ps("gemm_additive_dpre", M, q, D, stream);
    return launch_gemm_nt(plan, e, X, lda, Wa, ldw, stream);
}

int gemm_pool_dinput(const void* dpre, int M, int ld_dpre, int q, const void* WaT, int D, int ldwT, const float* w,
                     const float* dout, int ldo, int seg_len, void* dx, int ld_dx, RowMapCfg rm, int zero_pad_rows,
                     DropoutCfg drop, const void* relu_src, int relu_ld, cudaStream_t stream) {
    if (M == 0) return 0;
    NR_REQUIRE(seg_len >= 1, "pool_dinput: seg_len=%d", seg_len);
    // the epilogue stages the dOut rows of every segment a tile touches: cap the slice width so that they fit
    const int nseg_max = kTileM / seg_len + 2;
    const int max_stride = (EpiDPoolIn::kStageFloats / nseg_max) & ~15;
    NR_REQUIRE(max_stride >= 16, "pool_dinput: seg_len=%d needs %d staged segments per tile", seg_len, nseg_max);
    GemmNTPlan plan;
    NR_PROPAGATE(plan_gemm_nt(&plan, dpre, M, ld_dpre, WaT, D, ldwT, q, 1, 0, kTileM, num_sms(), 0,
                              EpiDPoolIn::kScratchBytes, max_stride));
    NR_REQUIRE(ld_dx % 8 == 0, "pool_dinput: ld_dx=%d", ld_dx);
    EpiDPoolIn e;
    memset(&e, 0, sizeof(e));
    e.use_tma = (rm.seg_in == 0 && relu_src == nullptr && D >= 32) ? 1 : 0;
    if (e.use_tma) NR_PROPAGATE(make_tmap_bf16_2d(&e.tm_out, dx, M, D, ld_dx, 32, 32, 64));
    e.w = w;
    e.dout = dout;
    e.ldo = ldo;
    e.seg_len = seg_len;
    e.dx = static_cast<__nv_bfloat16*>(dx);
    e.ld = ld_dx;
    e.N = D;
    e.rm = to_rm(rm);
    e.zero_pad_rows = zero_pad_rows;
    e.drop = to_drop(drop);
    e.relu_src = static_cast<const __nv_bfloat16*>(relu_src);
    e.relu_ld = relu_ld;
    e.M = M;
    e.rows_per_tile = kTileM;
    g_launches += debug_simt_gemm() ? 2 : 1;
    ProfScope ps("gemm_pool_dinput", M, D, q, stream);
    return launch_gemm_nt(plan, e, dpre, ld_dpre, WaT, ldwT, stream);
}

int gemm_scatter_emb(const void* A, int M, int lda, const void* W, int N, int ldw, int K, int taps, int w_tap_rows,
                     int rows_per_tile, const long long* ids, float* demb, int V, int D, RowMapCfg rm, DropoutCfg drop,
                     int drop_ld, cudaStream_t stream) {
    if (M == 0) return 0;
    NR_REQUIRE(N == D && D % 4 == 0 && V >= 1, "scatter_emb: N=%d D=%d V=%d", N, D, V);
    GemmNTPlan plan;
    NR_PROPAGATE(plan_gemm_nt(&plan, A, M, lda, W, N, ldw, K, taps, w_tap_rows, rows_per_tile, num_sms(), 0,
                              EpiScatter::kScratchBytes, 0));
    EpiScatter e;
    e.ids = ids;
    e.demb = demb;
    e.V = V;
    e.D = D;
    e.rm = to_rm(rm);
    e.drop = to_drop(drop);
    e.drop_ld = drop_ld;
    g_launches += debug_simt_gemm() ? 2 : 1;
    ProfScope ps("gemm_scatter_emb", M, N, K * taps, stream);
    return launch_gemm_nt(plan, e, A, lda, W, ldw, stream);
}

}  // namespace nr
