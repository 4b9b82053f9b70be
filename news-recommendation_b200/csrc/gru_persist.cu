// Persistent recurrence of the LSTUR user encoder's GRU (reference src/model/LSTUR/user_encoder.py:27-45), forward:
// ONE cooperative launch runs all S time steps instead of 3 launches per step.
//
//   gh_t = h_{t-1} . W_hh^T + b_hh            [B x Hd] x [Hd x 3Hd]   (tcgen05, bf16 operands, fp32 accumulate)
//   r = sig(gi_r + gh_r), z = sig(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h_t = (1 - z) n + z h_{t-1}   for t < len[b]
//
// Decomposition: users in 128-row tiles (m) x hidden units in slices of 32 (s); CTA (m, s) keeps its 96 weight rows
// (32 units x gates r, z, n; 180 KB, SWIZZLE_128B K-major) resident in shared memory for the whole launch and the fp32
// hidden state of its 128 x 32 block in REGISTERS.  Per step it streams the bf16 h_{t-1} rows of its tile (all Hd columns,
// written by the 29 slice CTAs of the same row tile) through a 2-stage TMA ring, issues ceil(Hd/16) MMAs of N = 96 into
// TMEM, and runs the gates in the epilogue (gi of the step is prefetched into registers while the previous step's barrier
// is pending).  Steps are separated by a release/acquire counter barrier per ROW TILE (only the slices of one tile exchange
// data).  The launch is cooperative (all CTAs resident) -- row_tiles * slices <= SM count is checked by the host.
// Saved for the (per-step) backward: gh (fp32), hs (fp32), hb (bf16 operand rows with the ones column), as before.
#include <algorithm>
#include <cstring>

#define NR_WATCHDOG_SYMBOL g_gru_dev_error
#include "nr_fused.cuh"
#include "nr_ops.h"

namespace nr {

extern int g_launches;

int read_attn_device_error(int* out4);
int read_gru_device_error(int* out4) {
    const int rc = static_cast<int>(cudaMemcpyFromSymbol(out4, fused::g_gru_dev_error, sizeof(int) * 4));
    if (rc != 0 || out4[0] != 0) return rc;
    return read_attn_device_error(out4);  // the title-level attention kernel keeps its own record (attn_title.cu)
}

namespace gru {

using namespace fused;

constexpr int kThreads = 6 * 32;   // 4 epilogue warps + TMA producer + tcgen05 issuer
constexpr int kU = 32;             // hidden units per slice
constexpr int kN = 3 * kU;         // MMA N: gates r | z | n of the slice
constexpr int kWChunk = kN * 128;  // one 64-column k-chunk of the resident weight slice
constexpr int kAStage = 128 * 128;
constexpr int kAStages = 2;

struct Params {
    int B, S, Hd, ldh, ldg;
    int slices, k_chunks, ksteps_last;
    const float* gi;         // [B*S][ldg]  rows b*S + t
    const float* bhh;        // [3Hd]
    const float* h0;         // [B][Hd]
    const long long* len;    // [B]
    float* gh;               // [S][B][ldg]
    float* hs;               // [S+1][B][Hd]   hs[0] = h0 (filled by the caller)
    __nv_bfloat16* hb;       // [S+1][B][ldh]  hb[0] = bf16(h0) + ones column (filled by the caller)
    float* out;              // [B][Hd]
    unsigned int* bar;       // [row tiles] step barrier counters, zero on entry
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned int* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) gru_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmH,
                                                                         const __grid_constant__ CUtensorMap tmW, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = base;
    uint8_t* sA = sW + p.k_chunks * kWChunk;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kAStages * kAStage);
    uint64_t* wfull = bars;            // resident weights landed
    uint64_t* afull = bars + 1;        // [kAStages]
    uint64_t* aempty = bars + 1 + kAStages;
    uint64_t* tfull = bars + 1 + 2 * kAStages;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    float* sBias = reinterpret_cast<float*>(bars + 16);      // [3][kU] b_hh of this slice (zero for units that do not exist); 16-byte aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x / p.slices, s = blockIdx.x - m * p.slices;
    const int j0 = s * kU;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmH);
        tma_prefetch_desc(&tmW);
        mbar_init(wfull, 1);
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&afull[i], 1);
            mbar_init(&aempty[i], 1);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        fence_barrier_init();
    } else if (warp == 5) {
        tmem_alloc(tmem_slot, 128);
    }
    if (threadIdx.x < kN) {
        const int g = threadIdx.x / kU, u = threadIdx.x - g * kU;
        sBias[threadIdx.x] = (j0 + u < p.Hd) ? p.bhh[g * p.Hd + j0 + u] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer: the weight slice once, then h_{t-1} tiles step by step =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(p.k_chunks * kWChunk));
            for (int c = 0; c < p.k_chunks; ++c)
                for (int g = 0; g < 3; ++g)  // rows beyond 3*Hd (last slice, gate n) are zero filled; rows of a neighbouring gate are finite and unused
                    tma_load_2d(sW + c * kWChunk + g * (kU * 128), &tmW, wfull, c * 64, g * p.Hd + j0);
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 0;
        for (int t = 0; t < p.S; ++t) {
            if (t > 0) {  // every slice of this row tile has published h_t
                if (lane == 0) {
                    const unsigned int want = static_cast<unsigned int>(t) * static_cast<unsigned int>(p.slices);
                    const uint64_t t0 = globaltimer_ns();
                    uint32_t spins = 0;
                    while (ld_acquire(p.bar + m) < want) {
                        __nanosleep(40);
                        if ((++spins & 0xfff) == 0 && globaltimer_ns() - t0 > 4000000000ull) f_timeout(401, static_cast<uint32_t>(t));
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of the other CTAs -> this CTA's TMA reads
                }
                __syncwarp();
            }
            for (int c = 0; c < p.k_chunks; ++c) {
                f_wait(&aempty[st], ph ^ 1u, 402);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&afull[st], kAStage);
                    tma_load_2d(sA + st * kAStage, &tmH, &afull[st], c * 64, t * p.B + m * 128);
                }
                __syncwarp();
                if (++st == kAStages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ===================== tcgen05 issuer =====================
        const uint32_t idesc = make_idesc_bf16(128, kN, 0, 0);
        const uint32_t a_s = smem_u32(sA), w_s = smem_u32(sW);
        f_wait(wfull, 0, 403);
        tc_fence_after();
        int st = 0;
        uint32_t ph = 0;
        for (int t = 0; t < p.S; ++t) {
            f_wait(tempty, static_cast<uint32_t>(t & 1) ^ 1u, 404);  // the epilogue has read the accumulator of step t - 1
            tc_fence_after();
            for (int c = 0; c < p.k_chunks; ++c) {
                f_wait(&afull[st], ph, 405);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_sw128_desc(a_s + st * kAStage, 0, 1024);
                    const uint64_t db = make_sw128_desc(w_s + c * kWChunk, 0, 1024);
                    const int nk = (c == p.k_chunks - 1) ? p.ksteps_last : 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < nk) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (c | k) ? 1u : 0u);
                    umma_commit(&aempty[st]);
                }
                __syncwarp();
                if (++st == kAStages) { st = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull);
            __syncwarp();
        }
    } else {
        // ===================== epilogue: gates, hidden state in registers =====================
        const int r = warp * 32 + lane;
        const int b = m * 128 + r;
        const bool valid = b < p.B;
        const int nvalid = min(kU, p.Hd - j0);  // units of this slice that exist (multiple of 4)
        const uint32_t acc_t = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        long long L = valid ? p.len[b] : 1;
        if (L < 1) L = 1;  // reference clamps 0 -> 1 (user_encoder.py:27)
        float h[kU];
#pragma unroll
        for (int u = 0; u < kU; u += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && u < nvalid) v = *reinterpret_cast<const float4*>(p.h0 + static_cast<size_t>(b) * p.Hd + j0 + u);
            h[u] = v.x; h[u + 1] = v.y; h[u + 2] = v.z; h[u + 3] = v.w;
        }
        auto add_bias = [&](float* x, int g) {
#pragma unroll
            for (int u = 0; u < kU; u += 4) {
                const float4 b4 = lds_f4(sBias + g * kU + u);
                x[u] += b4.x; x[u + 1] += b4.y; x[u + 2] += b4.z; x[u + 3] += b4.w;
            }
        };
        float gi[3][kU];
        auto load_gi = [&](int t) {  // input projections of step t for this row's units (prefetched under the barrier / the MMAs)
            const float* row = p.gi + (static_cast<size_t>(valid ? b : 0) * p.S + t) * p.ldg + j0;
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int u = 0; u < kU; u += 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u < nvalid) v = __ldg(reinterpret_cast<const float4*>(row + g * p.Hd + u));
                    gi[g][u] = v.x; gi[g][u + 1] = v.y; gi[g][u + 2] = v.z; gi[g][u + 3] = v.w;
                }
        };
        load_gi(0);
        for (int t = 0; t < p.S; ++t) {
            f_wait(tfull, static_cast<uint32_t>(t & 1), 406);
            tc_fence_after();
            float rg[kU], zg[kU], x[kU];
            float* ghrow = p.gh + (static_cast<size_t>(t) * p.B + (valid ? b : 0)) * p.ldg + j0;
            // gate r
            tmem_ld32(acc_t, x);
            tmem_ld_wait();
#pragma unroll
            add_bias(x, 0);
#pragma unroll
            for (int u = 0; u < kU; ++u) rg[u] = fast_sigmoid(gi[0][u] + x[u]);
            if (valid) {
#pragma unroll
                for (int u = 0; u < kU; u += 4)
                    if (u < nvalid) *reinterpret_cast<float4*>(ghrow + u) = make_float4(x[u], x[u + 1], x[u + 2], x[u + 3]);
            }
            // gate z
            tmem_ld32(acc_t + kU, x);
            tmem_ld_wait();
#pragma unroll
            add_bias(x, 1);
#pragma unroll
            for (int u = 0; u < kU; ++u) zg[u] = fast_sigmoid(gi[1][u] + x[u]);
            if (valid) {
#pragma unroll
                for (int u = 0; u < kU; u += 4)
                    if (u < nvalid) *reinterpret_cast<float4*>(ghrow + p.Hd + u) = make_float4(x[u], x[u + 1], x[u + 2], x[u + 3]);
            }
            // gate n and the state update
            tmem_ld32(acc_t + 2 * kU, x);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);  // the issuer may overwrite the accumulator (after the next step's barrier)
            add_bias(x, 2);
            if (valid) {
#pragma unroll
                for (int u = 0; u < kU; u += 4)
                    if (u < nvalid) *reinterpret_cast<float4*>(ghrow + 2 * p.Hd + u) = make_float4(x[u], x[u + 1], x[u + 2], x[u + 3]);
            }
            if (t < L) {
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const float n = fast_tanh(gi[2][u] + rg[u] * x[u]);
                    h[u] = (1.f - zg[u]) * n + zg[u] * h[u];
                }
            }
            if (valid) {
                float* hsrow = p.hs + (static_cast<size_t>(t + 1) * p.B + b) * p.Hd + j0;
                __nv_bfloat16* hbrow = p.hb + (static_cast<size_t>(t + 1) * p.B + b) * p.ldh + j0;
#pragma unroll
                for (int u = 0; u < kU; u += 4) {
                    if (u < nvalid) {
                        *reinterpret_cast<float4*>(hsrow + u) = make_float4(h[u], h[u + 1], h[u + 2], h[u + 3]);
                        *reinterpret_cast<uint2*>(hbrow + u) = make_uint2(pack_bf16x2(h[u], h[u + 1]), pack_bf16x2(h[u + 2], h[u + 3]));
                    }
                }
                if (nvalid < kU || j0 + kU == p.Hd) {  // the slice that ends at Hd also owns the ones column and the zero pad
                    for (int c = p.Hd; c < p.ldh; ++c) p.hb[(static_cast<size_t>(t + 1) * p.B + b) * p.ldh + c] = __float2bfloat16_rn(c == p.Hd ? 1.0f : 0.f);
                }
            }
            if (t + 1 < p.S) {
                load_gi(t + 1);
                __threadfence();                                   // this thread's h_{t+1} rows are visible GPU-wide ...
                asm volatile("bar.sync 1, 128;" ::: "memory");     // ... for all four epilogue warps ...
                if (threadIdx.x == 0) red_release(p.bar + m);      // ... before the slice counts as arrived
            }
        }
        if (valid) {
            float* o = p.out + static_cast<size_t>(b) * p.Hd + j0;
#pragma unroll
            for (int u = 0; u < kU; u += 4)
                if (u < nvalid) *reinterpret_cast<float4*>(o + u) = make_float4(h[u], h[u + 1], h[u + 2], h[u + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

}  // namespace gru

// 1 if the persistent recurrence covers this shape on this device (else the caller runs the per-step sequence)
int gru_persistent_supported(int B, int Hd) {
    using namespace gru;
    if (B < 1 || Hd < 32 || Hd % 4 != 0 || Hd > 960) return 0;
    const int row_tiles = ceil_div(B, 128), slices = ceil_div(Hd, kU);
    return row_tiles * slices <= num_sms() ? 1 : 0;
}

int gru_fwd_persistent(int B, int S, int Hd, int ldh, int ldg, const float* gi, const void* whh, const float* bhh, const float* h0,
                       const long long* len, float* gh, float* hs, void* hb, float* out, cudaStream_t stream) {
    using namespace gru;
    NR_REQUIRE(gru_persistent_supported(B, Hd), "gru_fwd_persistent: unsupported shape B=%d Hd=%d", B, Hd);
    NR_REQUIRE(ldh % 8 == 0 && ldh >= Hd + 1 && ldg % 4 == 0 && ldg >= 3 * Hd, "gru_fwd_persistent: pitches ldh=%d ldg=%d", ldh, ldg);
    Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.S = S; p.Hd = Hd; p.ldh = ldh; p.ldg = ldg;
    p.slices = ceil_div(Hd, kU);
    p.k_chunks = ceil_div(Hd, 64);
    p.ksteps_last = ceil_div(Hd - (p.k_chunks - 1) * 64, 16);
    p.gi = gi; p.bhh = bhh; p.h0 = h0; p.len = len; p.gh = gh; p.hs = hs;
    p.hb = static_cast<__nv_bfloat16*>(hb);
    p.out = out;
    const int row_tiles = ceil_div(B, 128);
    unsigned int* bar = nullptr;
    NR_CHECK_CUDA(cudaMallocAsync(&bar, sizeof(unsigned int) * row_tiles, stream));
    NR_CHECK_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned int) * row_tiles, stream));
    p.bar = bar;
    CUtensorMap tmH, tmW;
    NR_PROPAGATE(make_tmap_bf16_2d(&tmH, hb, static_cast<int64_t>(S + 1) * B, Hd, ldh, 64, 128));
    NR_PROPAGATE(make_tmap_bf16_2d(&tmW, whh, 3 * static_cast<int64_t>(Hd), Hd, ldh, 64, kU));
    const size_t smem = 1024 + static_cast<size_t>(p.k_chunks) * kWChunk + kAStages * kAStage + 1024;
    NR_REQUIRE(smem <= 232448, "gru_fwd_persistent: %zu bytes of shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        NR_CHECK_CUDA(cudaFuncSetAttribute(gru_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    ProfScope ps("gru_fwd_persistent", B, S, Hd, stream);
    void* args[] = {&tmH, &tmW, &p};
    NR_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gru_fwd_persistent_kernel), dim3(row_tiles * p.slices), dim3(kThreads),
                                              args, smem, stream));
    ++g_launches;
    NR_CHECK_CUDA(cudaFreeAsync(bar, stream));
    return 0;
}

}  // namespace nr
