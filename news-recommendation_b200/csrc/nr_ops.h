// Internal (C++) operator layer between the kernels and the C-ABI (abi.cu).  All pointers are device
// pointers, all work is enqueued on `stream`, nothing is allocated except in the debug GEMM backend.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nr {

struct DropoutCfg {
    float p;        // 0 => off
    uint64_t seed;
};

struct RowMapCfg {  // see RowMap in nr_epilogues.cuh; seg_in == 0 => identity
    int seg_in, in_off, seg_len, seg_out, out_off;
};

// ---- tcgen05 GEMMs with fused epilogues (gemm.cu) ------------------------------------------------
// out[rows x N] = act(A . W^T + bias) (bf16 or fp32).  A bf16 [M x K] pitch lda (taps>1: padded CNN layout),
// W bf16 [taps*w_tap_rows x K] pitch ldw.
int gemm_store(const void* A, int M, int lda, const void* W, int N, int ldw, int K, int taps, int w_tap_rows,
               int rows_per_tile, const float* bias, int relu, void* out, int ld_out, int out_bf16, RowMapCfg rm,
               int zero_pad_rows, DropoutCfg drop, int ones_col, int ones_zero_upto, cudaStream_t stream,
               void* lo_out = nullptr, int ld_lo = 0, int lo_col0 = 0, int accumulate = 0);
// accumulate (fp32 output only): out += A . W^T (+ bias) instead of out =
// lo_out (bf16 [M][ld_lo], identity rows, bf16 output only): columns [lo_col0, N) additionally leave as a LOW plane,
// lo[r][c - lo_col0] = bf16(y - bf16(y)), so that a consumer can read y as a hi/lo bf16 pair (~16 mantissa bits)

// additive-attention pooling: out[seg][D] = sum_r softmax_seg(tanh(X Wa^T + ba) . qv)_r X_r ; w_out[rows]
// X_lo (may be null): a second bf16 plane with X = X_hi + X_lo; the scores use X_hi, the pooled sum both planes.
int gemm_additive_pool(const void* X, int M, int lda, int D, const void* Wa, int q, int ldw, const float* ba,
                       const float* qv, int seg_len, float* out, int ldo, float* w_out, cudaStream_t stream,
                       const void* X_lo = nullptr);

// dPre = dscore * qv * (1 - tanh^2(X Wa^T + ba)) -> bf16 [M x ld_dpre]; dqv += sum_r dscore_r tanh(..)
int gemm_additive_dpre(const void* X, int M, int lda, int D, const void* Wa, int q, int ldw, const float* ba,
                       const float* qv, const float* dscore, void* dpre, int ld_dpre, float* dqv,
                       cudaStream_t stream);

// dX = dPre . Wa + w (x) dOut  [* relu mask] [* dropout]  -> bf16 (optionally re-mapped to the padded layout)
int gemm_pool_dinput(const void* dpre, int M, int ld_dpre, int q, const void* WaT, int D, int ldwT, const float* w,
                     const float* dout, int ldo, int seg_len, void* dx, int ld_dx, RowMapCfg rm, int zero_pad_rows,
                     DropoutCfg drop, const void* relu_src, int relu_ld, cudaStream_t stream);

// dEmb[ids[row]] += A . W^T  (embedding gradient; padding row 0 skipped) [* dropout of the gathered rows]
int gemm_scatter_emb(const void* A, int M, int lda, const void* W, int N, int ldw, int K, int taps, int w_tap_rows,
                     int rows_per_tile, const long long* ids, float* demb, int V, int D, RowMapCfg rm, DropoutCfg drop,
                     int drop_ld, cudaStream_t stream);

// D[Ma x Nb] += A[:, :Ma]^T . B[rows + shift, b_col0 : b_col0 + Nb]   (fp32 accumulate into D, pitch ldd)
int gemm_tn_accumulate(const void* A, int Kr, int Ma, int lda, const void* B, int b_rows, int b_cols, int ldb,
                       int b_col0, int Nb, int b_row_shift, float* D, int ldd, cudaStream_t stream);

// ---- memory-bound companions (aux.cu) --------------------------------------------------------------
// fp32 [R x C] (pitch lds) -> bf16 [R x ld] zero padded; transpose: out[c][r] = in[r][c] (out is [C x ld])
int cast_pad_bf16(const float* src, int R, int C, int lds, void* dst, int ld, int transpose, cudaStream_t stream);
// up to 8 such casts in one launch
int cast_pad_bf16_many(int n, const float* const* src, const int* R, const int* C, const int* lds, void* const* dst, const int* ld,
                       const int* transpose, cudaStream_t stream);
// fp32 rows [n_seq][T][D] with element strides -> bf16 [n_seq*T x ld], ones column at D, zeros after
int rows_to_bf16(const float* src, long long n_seq, int T, int D, long long s_seq, long long s_tok, long long s_col,
                 void* dst, int ld, cudaStream_t stream);
// X[row(seg,t)] = table_bf16[ids[seg*T+t]] (bit-exact copy), ones column at D, optional dropout, optional padded layout
int gather_rows(const long long* ids, long long n_tok, int T, const void* table, int V, int D, int ld_table, void* X,
                int ld_x, int padded, DropoutCfg drop, int* bad_id_flag, cudaStream_t stream);
// multi-head self attention core on packed Q|K|V bf16 [n_seq*T x ld_qkv]: sections start at columns 0, sec, 2*sec
// (sec >= d = heads*dk; dQKV uses the same sections)
int mhsa_core_fwd(const void* qkv, int ld_qkv, int sec, long long n_seq, int T, int heads, int dk, void* ctx, int ld_ctx,
                  DropoutCfg drop, cudaStream_t stream);
int mhsa_core_bwd(const void* qkv, int ld_qkv, int sec, const void* dctx, int ld_dctx, long long n_seq, int T, int heads, int dk,
                  void* dqkv, int ld_dqkv, cudaStream_t stream);
// title-level backward (attn_title.cu): T = 20, d_k = 20, <= 15 heads, sections with a 16-byte phase (sec % 8 == 0)
bool mhsa_title_fwd_supported(int T, int dk, int heads, int sec, int ld_qkv, int ld_ctx);
int mhsa_title_fwd(const void* qkv, int ld_qkv, int sec, long long n_seq, int heads, void* ctx, int ld_ctx, DropoutCfg drop,
                   cudaStream_t stream, const void* v_lo = nullptr, int ld_vlo = 0, void* ctx_lo = nullptr);
bool mhsa_title_bwd_supported(int T, int dk, int heads, int sec, int ld_qkv, int ld_dctx, int ld_dqkv);
int mhsa_title_bwd(const void* qkv, int ld_qkv, int sec, const void* dctx, int ld_dctx, long long n_seq, int heads, void* dqkv,
                   int ld_dqkv, cudaStream_t stream);
// dscore_r = w_r (dw_r - sum_seg w dw), dw_r = dOut[seg] . X_r
int pool_dscore(const void* X, int lda, int D, long long n_seg, int seg_len, const float* w, const float* dout, int ldo,
                float* dscore, cudaStream_t stream);
// logits[b][c] = cand[b][c] . user[b]
int dot_score_fwd(const float* cand, const float* user, int B, int C, int D, float* logits, cudaStream_t stream);
int dot_score_bwd(const float* cand, const float* user, const float* dlogits, int B, int C, int D, float* dcand,
                  float* duser, cudaStream_t stream);

// dst bf16 [n][ldn] = dy[n][N] (pitch ld_dy) masked by (relu_out > 0) when relu_out != null (same pitch); zero padded
int accumulate_ext_grad(float* ext, int rows, int ld, int D, float* dW, float* db, cudaStream_t stream);
int relu_bwd_to_bf16(const float* dy, const float* relu_out, long long n, int N, int ld_dy, void* dst, int ldn, cudaStream_t stream);
// fp32 embedding lookup / scatter-add (padding row 0: value read as-is, gradient skipped)
int embedding_f32_fwd(const long long* ids, long long n, const float* table, int V, int D, float* out, int* bad_id_flag,
                      cudaStream_t stream);
int embedding_f32_bwd(const long long* ids, long long n, const float* dout, int V, int D, float* dtable, cudaStream_t stream);

// ---- fused NRMS news-encoder front end (fused_fwd.cu): ids -> gather -> Q|K|V -> attention -> context hi/lo planes -----
// w_heads bf16 [heads*64][ldx]: per head the rows W_Q[h] | W_K[h] | W_V[h] | zero rows up to 64; b_heads fp32 [heads*64].
// X may be null (inference): it is only written for the backward kernels.  Q|K|V never reaches HBM.
int mhsa_fused_supported(int T, int d, int heads);
int mhsa_fused_fwd(const long long* ids, long long n_seq, int T, const void* table, int V, int d, int heads, int ldx,
                   const void* w_heads, const float* b_heads, DropoutCfg drop_x, DropoutCfg drop_c, void* X, void* C_hi, void* C_lo,
                   int* bad_id_flag, cudaStream_t stream);
int read_fused_device_error(int* out4);

// ---- persistent GRU recurrence (gru_persist.cu): all S steps of h_t = GRU(gi_t, h_{t-1}) in one cooperative launch --------
int gru_persistent_supported(int B, int Hd);
int gru_fwd_persistent(int B, int S, int Hd, int ldh, int ldg, const float* gi, const void* whh, const float* bhh, const float* h0,
                       const long long* len, float* gh, float* hs, void* hb, float* out, cudaStream_t stream);

// precise user encoder (NRMS precise mode): hi/lo K-concatenated operand rows, fp32 attention with hi/lo context planes
int rows_to_bf16_lo(const float* src, long long n_seq, int T, int D, long long s_seq, long long s_tok, long long s_col, void* dst, int ld,
                    cudaStream_t stream);
int rows_to_bf16_hilo(const float* src, long long n_seq, int T, int D, long long s_seq, long long s_tok, long long s_col, void* dst,
                      int ld, cudaStream_t stream);
int mhsa_f32_fwd(const float* qkv, int ld, int sec, long long n_seq, int T, int heads, int dk, void* c_hi, void* c_lo, int ldc,
                 cudaStream_t stream);
// scores[i] = news[cand[i]] . user[s] for seg_offsets[s] <= i < seg_offsets[s+1]  (batched evaluate.py:245-265)
int segment_dot(const float* news, long long n_news, int D, const long long* cand, long long n_cand, const long long* seg_offsets,
                long long n_seg, const float* user, float* scores, int* bad_flag, cudaStream_t stream);

int slots_device_readable(const void* const* slots, int n);
int pack_slots(const void* const* slots, int H, int C, int B, int L, long long* out, cudaStream_t stream);
int num_sms();

// ---- live per-kernel timing (bench.py): CUDA events on the launching stream around every kernel ------
// Off by default.  A ProfScope brackets one kernel launch; names are "<context>/<op>[shape]".
void prof_enable(int on);
void prof_context(const char* ctx);  // prefix set by the composites ("news.fwd", "user.bwd", ...)
int prof_report(char* buf, int cap); // JSON {"name": [launches, total_ms], ...}; clears the records
struct ProfScope {
    ProfScope(const char* op, int a, int b, int c, cudaStream_t s);
    ~ProfScope();
    int idx;
    cudaStream_t stream;
};

}  // namespace nr
