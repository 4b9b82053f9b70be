// LSTUR user encoder: nn.GRU over the (left-padded) history, packed-sequence semantics, last hidden state.
// reference: src/model/LSTUR/user_encoder.py:16-45  (pack_padded_sequence(first len[b] steps) -> nn.GRU).
//
// The input projection is ONE tcgen05 GEMM over all (user, step) rows; the recurrent projection is a chain of
// S sequentially dependent [B x Hd] x [Hd x 3Hd] tcgen05 GEMMs (weight slices stay resident per launch, rows of
// one step are one or a few M tiles), each followed by a fused gate kernel.  Gate order r, z, n (torch):
//   r = sig(gi_r + gh_r), z = sig(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) n + z h ;
// user b stops updating after len[b] steps.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/newsrec_b200.h"
#include "nr_common.cuh"
#include "nr_ops.h"

namespace nr {
extern int g_launches;

// gi: fp32 [B*S][ldg] rows (b*S + t);  gh: fp32 [B][ldg];  h: fp32 [B][Hd] (in/out);
// hb_next: bf16 [B][ldh] operand of the next step (ones column at Hd)
__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int ldg, float* __restrict__ h,
                                    __nv_bfloat16* __restrict__ hb_next, int ldh, const long long* __restrict__ len, int B, int S,
                                    int Hd, int t) {
    const long long total = static_cast<long long>(B) * ldh;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / ldh), j = static_cast<int>(i - static_cast<long long>(b) * ldh);
        if (j > Hd) { hb_next[i] = __float2bfloat16_rn(0.f); continue; }
        if (j == Hd) { hb_next[i] = __float2bfloat16_rn(1.0f); continue; }
        float hv = h[static_cast<size_t>(b) * Hd + j];
        const long long L = len[b] < 1 ? 1 : len[b];  // reference clamps 0 -> 1 (user_encoder.py:27)
        if (t < L) {
            const float* gir = gi + (static_cast<size_t>(b) * S + t) * ldg;
            const float* ghr = gh + static_cast<size_t>(b) * ldg;
            const float r = fast_sigmoid(gir[j] + ghr[j]);
            const float z = fast_sigmoid(gir[Hd + j] + ghr[Hd + j]);
            const float n = fast_tanh(gir[2 * Hd + j] + r * ghr[2 * Hd + j]);
            hv = (1.f - z) * n + z * hv;
            h[static_cast<size_t>(b) * Hd + j] = hv;
        }
        hb_next[i] = __float2bfloat16_rn(hv);
    }
}

// Backward of one step.  dh_in = dha + dhb (direct path + recurrent GEMM path of the step after).
// Writes dgi (bf16, rows b*S+t of [B*S][ldb]), dgh (bf16 [B][ldb]), dh_direct (fp32 [B][Hd]).
__global__ void gru_gate_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int ldg, const float* __restrict__ hprev,
                                    const float* __restrict__ dha, int pa, const float* __restrict__ dhb, int pb,
                                    const long long* __restrict__ len, int B, int S, int Hd, int t, __nv_bfloat16* __restrict__ dgi,
                                    __nv_bfloat16* __restrict__ dgh, int ldb, float* __restrict__ dh_direct, int pd) {
    const long long total = static_cast<long long>(B) * Hd;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / Hd), j = static_cast<int>(i - static_cast<long long>(b) * Hd);
        const float dh = dha[static_cast<size_t>(b) * pa + j] + (dhb != nullptr ? dhb[static_cast<size_t>(b) * pb + j] : 0.f);
        const long long L = len[b] < 1 ? 1 : len[b];
        __nv_bfloat16* dgir = dgi + (static_cast<size_t>(b) * S + t) * ldb;
        __nv_bfloat16* dghr = dgh + static_cast<size_t>(b) * ldb;
        if (t >= L) {
            dgir[j] = dgir[Hd + j] = dgir[2 * Hd + j] = __float2bfloat16_rn(0.f);
            dghr[j] = dghr[Hd + j] = dghr[2 * Hd + j] = __float2bfloat16_rn(0.f);
            dh_direct[static_cast<size_t>(b) * pd + j] = dh;
            continue;
        }
        const float* gir = gi + (static_cast<size_t>(b) * S + t) * ldg;
        const float* ghr = gh + static_cast<size_t>(b) * ldg;
        const float r = fast_sigmoid(gir[j] + ghr[j]);
        const float z = fast_sigmoid(gir[Hd + j] + ghr[Hd + j]);
        const float ghn = ghr[2 * Hd + j];
        const float n = fast_tanh(gir[2 * Hd + j] + r * ghn);
        const float hp = hprev[i];
        const float dn = dh * (1.f - z);
        const float dz = dh * (hp - n);
        const float dpn = dn * (1.f - n * n);
        const float dpz = dz * z * (1.f - z);
        const float dpr = dpn * ghn * r * (1.f - r);
        dgir[j] = __float2bfloat16_rn(dpr);
        dgir[Hd + j] = __float2bfloat16_rn(dpz);
        dgir[2 * Hd + j] = __float2bfloat16_rn(dpn);
        dghr[j] = __float2bfloat16_rn(dpr);
        dghr[Hd + j] = __float2bfloat16_rn(dpz);
        dghr[2 * Hd + j] = __float2bfloat16_rn(dpn * r);
        dh_direct[static_cast<size_t>(b) * pd + j] = dh * z;
    }
}

__global__ void add2_kernel(const float* __restrict__ a, int pa, const float* __restrict__ b, int pb, float* __restrict__ out, int B, int Hd) {
    const long long n = static_cast<long long>(B) * Hd;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / Hd, j = i - r * Hd;
        out[i] = a[r * pa + j] + (b != nullptr ? b[r * pb + j] : 0.f);
    }
}
__global__ void zero_pad_cols_kernel(__nv_bfloat16* __restrict__ m, long long rows, int from, int ld) {
    const int w = ld - from;
    const long long total = rows * w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / w;
        m[r * ld + from + (i - r * w)] = __float2bfloat16_rn(0.f);
    }
}

}  // namespace nr

using namespace nr;
static inline cudaStream_t S_(void* s) { return static_cast<cudaStream_t>(s); }
static inline long long align256(long long x) { return (x + 255) & ~255ll; }
static inline int ru8(int x) { return (x + 7) & ~7; }
static inline int ru4(int x) { return (x + 3) & ~3; }
static const RowMapCfg kIdentity = {0, 0, 0, 0, 0};
static const DropoutCfg kNoDrop = {0.f, 0};
static inline int blocks_for(long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8)); }

extern "C" {

// Saved-state sizes (elements): gi fp32 [B*S][ldg], gh fp32 [S][B][ldg], hs fp32 [S+1][B][Hd], hb bf16 [S+1][B][ldh],
// xb bf16 [B*S][ldd]   with ldg = round_up(3Hd, 4), ldh = round_up(Hd+1, 8), ldd = round_up(D+1, 8)
int nr_gru_fwd(const nr_gru_fwd_args* a, void* stream) {
    NR_REQUIRE(a != nullptr, "nr_gru_fwd: null args");
    const int B = a->B, S = a->S, D = a->D, Hd = a->Hd;
    NR_REQUIRE(B >= 0 && S >= 1 && D >= 8 && Hd >= 8 && D % 4 == 0 && Hd % 2 == 0, "nr_gru_fwd: bad shape B=%d S=%d D=%d Hd=%d", B, S, D, Hd);
    NR_REQUIRE(a->x && a->len && a->h0 && a->wih_bf16 && a->whh_bf16 && a->bih && a->bhh && a->xb && a->gi && a->gh && a->hs && a->hb && a->out,
               "nr_gru_fwd: null operand");
    if (B == 0) return 0;
    const cudaStream_t st = S_(stream);
    const int ldg = ru4(3 * Hd), ldh = ru8(Hd + 1), ldd = ru8(D + 1);
    const long long BH = static_cast<long long>(B) * Hd;
    prof_context("gru.fwd");
    // gi = X Wih^T + bih over all (user, step) rows
    NR_PROPAGATE(rows_to_bf16(a->x, B, S, D, a->x_s_b, a->x_s_t, a->x_s_c, a->xb, ldd, st));
    NR_PROPAGATE(gemm_store(a->xb, B * S, ldd, a->wih_bf16, 3 * Hd, ldd, D, 1, 0, 128, a->bih, 0, a->gi, ldg, 0, kIdentity, 0, kNoDrop, -1, 0, st));
    if (a->x_lo_bf16 != nullptr) {
        // accurate mode: the news vectors enter as a hi/lo bf16 pair, gi = x_hi . W_ih^T + b + x_lo . W_ih^T.  Two passes over the
        // SAME resident weights with fp32 accumulation into gi: a K-concatenated single pass doubles K, which shrinks the weight-
        // stationary slices to N = 80 and costs 0.86 ms instead of 2 x 0.23
        NR_PROPAGATE(rows_to_bf16_lo(a->x, B, S, D, a->x_s_b, a->x_s_t, a->x_s_c, a->x_lo_bf16, ldd, st));
        NR_PROPAGATE(gemm_store(a->x_lo_bf16, B * S, ldd, a->wih_bf16, 3 * Hd, ldd, D, 1, 0, 128, nullptr, 0, a->gi, ldg, 0, kIdentity, 0, kNoDrop,
                                -1, 0, st, nullptr, 0, 0, 1));
    }
    // h_0
    NR_CHECK_CUDA(cudaMemcpyAsync(a->hs, a->h0, sizeof(float) * BH, cudaMemcpyDeviceToDevice, st));
    NR_PROPAGATE(rows_to_bf16(a->h0, B, 1, Hd, Hd, 0, 1, a->hb, ldh, st));
    static const bool no_persist = getenv("NEWSREC_GRU_STEPWISE") != nullptr;  // tests compare the two paths
    if (!no_persist && gru_persistent_supported(B, Hd)) {
        // one cooperative launch for the whole recurrence (gru_persist.cu); same saved state as the per-step sequence below
        return gru_fwd_persistent(B, S, Hd, ldh, ldg, a->gi, a->whh_bf16, a->bhh, a->h0, a->len, a->gh, a->hs, a->hb, a->out, st);
    }
    for (int t = 0; t < S; ++t) {
        float* gh_t = a->gh + static_cast<size_t>(t) * B * ldg;
        const void* hb_t = static_cast<const __nv_bfloat16*>(a->hb) + static_cast<size_t>(t) * B * ldh;
        __nv_bfloat16* hb_n = static_cast<__nv_bfloat16*>(a->hb) + static_cast<size_t>(t + 1) * B * ldh;
        float* h_n = a->hs + static_cast<size_t>(t + 1) * BH;
        NR_PROPAGATE(gemm_store(hb_t, B, ldh, a->whh_bf16, 3 * Hd, ldh, Hd, 1, 0, 128, a->bhh, 0, gh_t, ldg, 0, kIdentity, 0, kNoDrop, -1, 0, st));
        NR_CHECK_CUDA(cudaMemcpyAsync(h_n, a->hs + static_cast<size_t>(t) * BH, sizeof(float) * BH, cudaMemcpyDeviceToDevice, st));
        {
            ProfScope ps("gru_gate_fwd", B, Hd, t, st);
            gru_gate_fwd_kernel<<<blocks_for(static_cast<long long>(B) * ldh), 256, 0, st>>>(a->gi, gh_t, ldg, h_n, hb_n, ldh, a->len, B, S, Hd, t);
            ++g_launches;
        }
        NR_CHECK_CUDA(cudaGetLastError());
    }
    NR_CHECK_CUDA(cudaMemcpyAsync(a->out, a->hs + static_cast<size_t>(S) * BH, sizeof(float) * BH, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int nr_gru_persistent_supported(int B, int Hd) { return gru_persistent_supported(B, Hd); }

long long nr_gru_bwd_workspace(int B, int S, int D, int Hd) {
    const long long ldb = ru8(3 * Hd + 1), BP = static_cast<long long>(B) * ru4(Hd);
    return align256(static_cast<long long>(B) * S * ldb * 2) * 2 + align256(BP * 4) * 2 + 256;
}

int nr_gru_bwd(const nr_gru_bwd_args* a, void* stream) {
    NR_REQUIRE(a != nullptr, "nr_gru_bwd: null args");
    const int B = a->B, S = a->S, D = a->D, Hd = a->Hd;
    NR_REQUIRE(B >= 0 && S >= 1 && D % 4 == 0 && Hd % 2 == 0, "nr_gru_bwd: bad shape B=%d S=%d D=%d Hd=%d", B, S, D, Hd);
    NR_REQUIRE(a->len && a->wihT_bf16 && a->whhT_bf16 && a->xb && a->gi && a->gh && a->hs && a->hb && a->dout && a->dWih_ext && a->dWhh_ext &&
                   a->dx && a->dh0 && a->workspace, "nr_gru_bwd: null operand");
    NR_REQUIRE(a->workspace_bytes >= nr_gru_bwd_workspace(B, S, D, Hd), "nr_gru_bwd: workspace too small");
    if (B == 0) return 0;
    const cudaStream_t st = S_(stream);
    const int ldg = ru4(3 * Hd), ldh = ru8(Hd + 1), ldd = ru8(D + 1), ldb = ru8(3 * Hd + 1);
    const long long BH = static_cast<long long>(B) * Hd;
    char* ws = static_cast<char*>(a->workspace);
    __nv_bfloat16* dgi = reinterpret_cast<__nv_bfloat16*>(ws);            // [B*S][ldb]  rows b*S+t
    ws += align256(static_cast<long long>(B) * S * ldb * 2);
    __nv_bfloat16* dgh = reinterpret_cast<__nv_bfloat16*>(ws);            // [S][B][ldb]
    ws += align256(static_cast<long long>(B) * S * ldb * 2);
    const int P = ru4(Hd);  // pitch of the internal dh buffers (fp32 vector stores of the GEMM epilogue)
    float* dh_direct = reinterpret_cast<float*>(ws);
    ws += align256(static_cast<long long>(B) * P * 4);
    float* dh_rec = reinterpret_cast<float*>(ws);
    prof_context("gru.bwd");
    {   // pad columns [3Hd, ldb) of both gradient matrices are K-extent-masked by TMA but read by nothing else: keep them clean
        zero_pad_cols_kernel<<<blocks_for(static_cast<long long>(B) * S * (ldb - 3 * Hd)), 256, 0, st>>>(dgi, static_cast<long long>(B) * S, 3 * Hd, ldb);
        zero_pad_cols_kernel<<<blocks_for(static_cast<long long>(B) * S * (ldb - 3 * Hd)), 256, 0, st>>>(dgh, static_cast<long long>(B) * S, 3 * Hd, ldb);
        g_launches += 2;
    }
    const float* dha = a->dout;
    const float* dhb = nullptr;
    int pa = Hd;
    for (int t = S - 1; t >= 0; --t) {
        const float* gh_t = a->gh + static_cast<size_t>(t) * B * ldg;
        __nv_bfloat16* dgh_t = dgh + static_cast<size_t>(t) * B * ldb;
        {
            ProfScope ps("gru_gate_bwd", B, Hd, t, st);
            gru_gate_bwd_kernel<<<blocks_for(BH), 256, 0, st>>>(a->gi, gh_t, ldg, a->hs + static_cast<size_t>(t) * BH, dha, pa, dhb, P, a->len, B, S,
                                                               Hd, t, dgi, dgh_t, ldb, dh_direct, P);
            ++g_launches;
        }
        NR_CHECK_CUDA(cudaGetLastError());
        // recurrent path: dh_{t-1} += dgh_t . Whh
        NR_PROPAGATE(gemm_store(dgh_t, B, ldb, a->whhT_bf16, Hd, ldb, 3 * Hd, 1, 0, 128, nullptr, 0, dh_rec, P, 0, kIdentity, 0, kNoDrop, -1, 0, st));
        dha = dh_direct;
        dhb = dh_rec;
        pa = P;
    }
    add2_kernel<<<blocks_for(BH), 256, 0, st>>>(dha, pa, dhb, P, a->dh0, B, Hd);
    ++g_launches;
    // weight gradients over all (step, user) rows; the ones column of the saved operands yields the bias gradients
    for (int c0 = 0; c0 < Hd + 1; c0 += 512) {
        const int nb = std::min(512, Hd + 1 - c0);
        NR_PROPAGATE(gemm_tn_accumulate(dgh, B * S, 3 * Hd, ldb, a->hb, B * S, Hd + 1, ldh, c0, nb, 0, a->dWhh_ext + c0, ldh, st));
    }
    for (int c0 = 0; c0 < D + 1; c0 += 512) {
        const int nb = std::min(512, D + 1 - c0);
        NR_PROPAGATE(gemm_tn_accumulate(dgi, B * S, 3 * Hd, ldb, a->xb, B * S, D + 1, ldd, c0, nb, 0, a->dWih_ext + c0, ldd, st));
    }
    // input gradient
    NR_PROPAGATE(gemm_store(dgi, B * S, ldb, a->wihT_bf16, D, ldb, 3 * Hd, 1, 0, 128, nullptr, 0, a->dx, D, 0, kIdentity, 0, kNoDrop, -1, 0, st));
    return 0;
}

}  // extern "C"
