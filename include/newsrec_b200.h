/* newsrec_b200 -- C ABI of the Blackwell-native (sm_100a) NRMS / NAML / LSTUR / TANR hot path.
 *
 * Drop-in boundary for the reference's Python modules (yusanshi/news-recommendation @ 8323a4f).  The
 * reference has no FFI of its own (pure PyTorch); these are the entry points a maintainer binds with
 * ctypes from src/model/general/**, src/model/<NAME>/{news,user}_encoder.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns every buffer
 *     (PyTorch caching allocator); the library allocates nothing persistent;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no implicit syncs;
 *   - return 0 on success, a positive cudaError_t on a CUDA failure, -1 on an argument/shape violation
 *     (detected before any launch); nr_last_error() returns the message for the calling thread;
 *   - bf16 operand matrices are row-major with a pitch ("ld", in elements) that is a multiple of 8;
 *     an activation matrix of logical width D carries a constant 1.0 in column D (it turns the next
 *     weight-gradient GEMM's extra column into the bias gradient) and zeros behind it;
 *   - token ids are int64 exactly as the reference's DataLoader produces them; indexing is bit exact;
 *     id 0 is padding_idx: its row is READ like any other (reference behaviour) and its gradient skipped.
 *   - re-entrant; the only global state is a launch counter and the device watchdog record.
 */
#ifndef NEWSREC_B200_H
#define NEWSREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ------------------------------------------------------------------------------ */
int nr_version(void);                       /* ABI version, currently 1 */
const char* nr_last_error(void);            /* message of the last failing call on this thread */
int nr_device_error(int out4[4]);           /* watchdog record {code, block, thread, aux}; code 0 = none */
long long nr_launch_count(void);            /* kernels this library has launched so far */
int nr_num_sms(void);
/* TRIAGE ONLY (tests): route the GEMMs through a plain SIMT accumulate + the same epilogue functors, to
 * tell a tcgen05/TMA pipeline bug from an epilogue bug.  Never enabled by the product path. */
void nr_debug_set_simt_gemm(int on);
/* 1 in a triage build (`make TRIAGE=1`), 0 in the release library, where nr_debug_set_simt_gemm is a no-op, the SIMT
 * kernels are not compiled and no environment switch is consulted on a launch path */
int nr_has_triage_backends(void);
/* TUNING ONLY (tools/kbench.py): dev_buf holds slots x 148 x 16 int64; the k-th gemm_nt planned after this call
 * writes, per CTA, cycle counters into slot k: [0] TMA producer waiting for a free A stage, [1] MMA issuer waiting
 * for A data, [2] MMA issuer waiting for a free TMEM accumulator, [3] epilogue waiting for a finished accumulator,
 * [4] epilogue body, [5] kernel, [6] tiles, [7] MMA issue loops, [8] tcgen05.commit.  Null (default) switches the counters off. */
void nr_debug_set_gemm_timing(void* dev_buf, int slots);
/* TUNING ONLY (tools/fused_timing.py): dev_buf holds 148 x 32 int64; every fused front-end launch after this call writes,
 * per CTA, cycle counters of its warp roles (epilogue groups [0..6], [8..14]; gather [16,17]; weight producer [20];
 * tcgen05 issuer [24..29]).  Null (default) switches the counters off. */
void nr_debug_set_fused_timing(void* dev_buf);
/* Data parallel: the weight-gradient GEMMs (nr_gemm_tn and the composites' internal calls) leave n SMs free, so that the
 * channel CTAs of a gradient all-reduce running on a side stream have somewhere to run (0 = use every SM, the default). */
void nr_reserve_sms_for_comm(int n);
/* Live per-kernel timing for bench.py: CUDA events on the launching stream around every kernel of this
 * library.  nr_profile_report writes JSON {"<context>/<op>[shape]": [launches, total_ms], ...}, returns its
 * length (or -1 if cap is too small) and clears the records.  Off by default. */
void nr_profile_enable(int on);
void nr_profile_context(const char* ctx);
int nr_profile_report(char* buf, int cap);

/* ---- operand preparation ------------------------------------------------------------------------- */
/* fp32 [R][C] (pitch lds) -> zero padded bf16 [R][ld]; transpose!=0: dst is [C][ld] with dst[c][r]=src[r][c] */
int nr_cast_pad_bf16(const float* src, int R, int C, int lds, void* dst_bf16, int ld, int transpose, void* stream);
/* the same for up to 8 matrices in ONE launch (host arrays of length n; the operands of an encoder are rebuilt together) */
int nr_cast_pad_bf16_many(int n, const float* const* src, const int* R, const int* C, const int* lds, void* const* dst_bf16,
                          const int* ld, const int* transpose, void* stream);
/* fp32 rows [n][D] with element strides -> bf16 [n][ld] + ones column */
int nr_rows_to_bf16(const float* src, long long n, int D, long long s_row, long long s_col, void* dst_bf16, int ld,
                    void* stream);

/* ---- reference: nn.Embedding lookup (src/model/NRMS/news_encoder.py:38 etc.) --------------------- */
/* X[row(seg,t)] = table[ids[seg*T+t]]; padded!=0 writes the zero-padded CNN layout (T+2 rows per segment).
 * *bad_id_flag (device int) is set to 1 if an id is outside [0,V) (the reference would raise IndexError). */
int nr_gather_rows(const long long* ids, long long n_tok, int T, const void* table_bf16, int V, int D, int ld,
                   void* X_bf16, int padded, float p_drop, unsigned long long seed, int* bad_id_flag, void* stream);

/* ---- reference: nn.Linear / nn.Conv2d(1,F,(3,d)) as a tcgen05 GEMM with fused bias/ReLU/dropout --- */
/* out[M][N] = act(A . W^T + bias); taps==3: window-3 conv over the padded layout (rows_per_tile = k*(T+2)) */
int nr_linear(const void* A_bf16, int M, int lda, const void* W_bf16, int N, int ldw, int K, int taps, int w_tap_rows,
              int rows_per_tile, const float* bias, int relu, void* out, int ld_out, int out_is_bf16, void* stream);

/* D[Ma][Nb] += A[:, :Ma]^T . B[rows+shift, b_col0:b_col0+Nb]   (weight gradients; fp32 accumulate) */
int nr_gemm_tn(const void* A_bf16, int Kr, int Ma, int lda, const void* B_bf16, int b_rows, int b_cols, int ldb,
               int b_col0, int Nb, int b_row_shift, float* D, int ldd, void* stream);

/* ---- reference: MultiHeadSelfAttention core (src/model/general/attention/multihead_self.py:15-23) -- */
/* Q | K | V sections of a row start at columns 0, sec, 2*sec (sec >= heads*dk; dQ|dK|dV likewise, padding written as zeros) */
int nr_mhsa_core_fwd(const void* qkv_bf16, int ld_qkv, int sec, long long n_seq, int T, int heads, int dk, void* ctx_bf16,
                     int ld_ctx, float p_drop, unsigned long long seed, void* stream);
int nr_mhsa_core_bwd(const void* qkv_bf16, int ld_qkv, int sec, const void* dctx_bf16, int ld_dctx, long long n_seq, int T,
                     int heads, int dk, void* dqkv_bf16, int ld_dqkv, void* stream);

/* ---- reference: AdditiveAttention.forward (src/model/general/attention/additive.py:27-53) ----------- */
/* X bf16 [n_seg*seg_len][ldx] (ones column at D) -> out fp32 [n_seg][ldo]; w_out [rows] saved for backward */
int nr_additive_attention_fwd(const void* X_bf16, long long n_seg, int seg_len, int D, int ldx, const void* Wa_bf16,
                              int q, int ldw, const float* ba, const float* qv, float* out, int ldo, float* w_out,
                              void* stream);
/* backward.  dX bf16 [rows][ld_dx] (=), dWa_ext fp32 [q][ldx] (+=, column D is d(bias)), dqv [q] (+=).
 * workspace: nr_additive_attention_bwd_workspace(...) bytes. */
long long nr_additive_attention_bwd_workspace(long long n_seg, int seg_len, int q);
int nr_additive_attention_bwd(const void* X_bf16, long long n_seg, int seg_len, int D, int ldx, const void* Wa_bf16,
                              const void* WaT_bf16, int q, int ldw, int ldwT, const float* ba, const float* qv,
                              const float* w, const float* dout, int ldo, void* dX_bf16, int ld_dx, float* dWa_ext,
                              float* dqv, void* workspace, long long workspace_bytes, void* stream);

/* ---- reference: DotProductClickPredictor.forward (src/model/general/click_predictor/dot_product.py) - */
int nr_dot_score_fwd(const float* cand, const float* user, int B, int C, int D, float* logits, void* stream);
int nr_dot_score_bwd(const float* cand, const float* user, const float* dlogits, int B, int C, int D, float* dcand,
                     float* duser, void* stream);

/* ---- reference: src/dataset.py:64-85 + default_collate (the slot-major batch the model receives, src/train.py:183-190) --
   slots[0 .. n_clicked) are the browsed-news tensors, then n_candidates candidate tensors, each int64 [B][L] contiguous and
   readable by the device (device memory or page-locked host memory, nr_slots_device_readable == 1).  One launch writes the
   impression-major block out[(b*n_clicked + h)*L + t], then out[B*n_clicked*L + (b*n_candidates + c)*L + t]. */
int nr_slots_device_readable(const void* const* slots, int n);
int nr_pack_slots(const void* const* slots, int n_clicked, int n_candidates, int B, int L, long long* out, void* stream);

/* Batched form of the evaluator's scoring loop (src/evaluate.py:245-265 calls get_prediction once per impression and
 * synchronises on .tolist() each time): the news vectors live in ONE device matrix news[n_news][D]; the candidates of
 * impression s are cand[seg_offsets[s] .. seg_offsets[s+1]) (indices into news), user[s] its user vector;
 * scores[i] = news[cand[i]] . user[s].  seg_offsets has n_seg + 1 entries (int64, device), seg_offsets[0] = 0.
 * *bad_id_flag is set if a candidate index is outside [0, n_news). */
int nr_segment_dot(const float* news, long long n_news, int D, const long long* cand, long long n_cand,
                   const long long* seg_offsets, long long n_seg, const float* user, float* scores, int* bad_id_flag,
                   void* stream);

/* Host-side glue of the weight-gradient GEMMs (nr_gemm_tn with the ones column): ext is [rows][ld] fp32 whose columns
 * [0,D) hold dW and column D holds db.  Adds them into the parameters' own gradient storage (dW [rows][D] contiguous,
 * db [rows] or null) and CLEARS ext, so the caller can keep it as a persistent accumulator across steps. */
int nr_accumulate_ext_grad(float* ext, int rows, int ld, int D, float* dW, float* db, void* stream);

/* ---- reference: NRMS NewsEncoder.forward / UserEncoder.forward -------------------------------------
 *   news  (src/model/NRMS/news_encoder.py:27-48): embedding -> dropout -> MHSA -> dropout -> additive pool
 *   user  (src/model/NRMS/user_encoder.py:15-26): MHSA -> additive pool over dense fp32 news vectors     */
typedef struct {
    long long n_seq;          /* titles (news) or users                                             */
    int T;                    /* tokens per title / history length                                  */
    int d;                    /* model width (word_embedding_dim)                                   */
    int heads;                /* num_attention_heads, d % heads == 0                                */
    int q;                    /* query_vector_dim                                                   */
    int ldx;                  /* pitch of X / C / weight operands: multiple of 8, >= d+1            */
    int ld3;                  /* pitch of Q|K|V rows: round_up(3*sec, 16), sec = round_up(d, 8): sections at columns 0, sec, 2*sec;
                                 packed weights / biases carry zero rows at the section padding */
    /* input: ids+table (news encoder) or dense (user encoder) */
    const long long* ids;     /* [n_seq*T] or NULL                                                  */
    const void* table_bf16;   /* [V][ldx]                                                           */
    int V;
    const float* dense;       /* fp32 [n_seq][T][d] with element strides below, or NULL             */
    long long dense_s_seq, dense_s_tok, dense_s_col;
    /* parameters as prepared operands */
    const void* wqkv_bf16;    /* [3d][ldx]  rows = W_Q | W_K | W_V                                  */
    const float* bqkv;        /* [3d]                                                               */
    const void* wa_bf16;      /* [q][ldx]                                                           */
    const float* ba;          /* [q]                                                                */
    const float* qv;          /* [q]                                                                */
    float p_drop;             /* dropout_probability when training, else 0                          */
    unsigned long long seed;
    /* outputs; X/QKV/C/w are what backward needs (the caller keeps them alive) */
    void* X_bf16;             /* [n_seq*T][ldx]                                                     */
    void* QKV_bf16;           /* [n_seq*T][ld3]                                                     */
    void* C_bf16;             /* [n_seq*T][ldx]                                                     */
    float* w;                 /* [n_seq*T] additive-attention weights                               */
    float* out;               /* [n_seq][d] fp32                                                    */
    int* bad_id_flag;         /* device int, set if an id is out of range                           */
    /* fused front end (ids variant, shapes nr_mhsa_fused_supported() accepts; all three NULL = unfused kernel sequence):
     * gather -> Q|K|V -> attention run as ONE kernel; the gathered rows are written to X_bf16 only when that pointer is
     * non-NULL (a backward pass will read them), Q|K|V never leaves the chip (QKV_bf16 must be NULL; the backward
     * recomputes it from X), and the context leaves as a bf16 hi plane (C_bf16) plus a bf16 lo plane (C_lo_bf16): the
     * pooled sum uses hi + lo. */
    const void* wqkv_heads_bf16; /* [heads*64][ldx]: per head the rows W_Q[h] | W_K[h] | W_V[h] | zero rows up to 64 */
    const float* bqkv_heads;     /* [heads*64] biases in the same order                                */
    void* C_lo_bf16;             /* [n_seq*T][ldx]                                                      */
    /* precise DENSE variant (user encoder of the precise mode; selected by dense != NULL and C_lo_bf16 != NULL): the fp32
     * input enters the projection as a hi/lo bf16 pair against K-concatenated weights, Q|K|V stays fp32, the attention runs
     * in fp32 on the CUDA cores, the context leaves as hi (C_bf16) + lo (C_lo_bf16) planes.  QKV_bf16 must be NULL. */
    const void* wqkv_kcat_bf16;  /* [3d][2*ldx]: columns [0,d) = W, [ldx, ldx+d) = W again, zeros elsewhere            */
    void* X_kcat_bf16;           /* [n_seq*T][2*ldx] workspace: hi | lo operand rows                                   */
    float* QKV_f32;              /* [n_seq*T][3*sec] workspace                                                           */
    /* accurate NEWS variant on the unfused kernels (selected by ids != NULL and V_lo_bf16 != NULL; needs C_lo_bf16 and
     * nr_mhsa_accurate_supported): V, the attention probabilities and the context are hi/lo bf16 pairs; X_bf16 and QKV_bf16
     * (the hi planes) are written as usual and saved for the backward. */
    void* V_lo_bf16;             /* [n_seq*T][sec] low plane of the V section                                            */
} nr_mhsa_encoder_fwd_args;
int nr_mhsa_accurate_supported(int T, int d, int heads); /* 1: the accurate news variant exists for this shape */
int nr_mhsa_encoder_fwd(const nr_mhsa_encoder_fwd_args* a, void* stream);
/* 1 if the fused front end handles (tokens per title, model width, heads): the reference's news level, T = 20, d_k = 20 */
int nr_mhsa_fused_supported(int T, int d, int heads);

typedef struct {
    long long n_seq;
    int T, d, heads, q, ldx, ld3, ldq;   /* ldq: pitch of dPre / WaT = round_up(q, 16)                   */
    const long long* ids;                /* NULL for the dense (user) variant                            */
    int V;
    const void* wqkvT_bf16;              /* [d][ld3]  = (W_Q|0|W_K|0|W_V|0)^T                               */
    const void* wa_bf16;                 /* [q][ldx]                                                     */
    const void* waT_bf16;                /* [d][ldq]                                                     */
    const float* ba;
    const float* qv;
    float p_drop;
    unsigned long long seed;
    const void* X_bf16;
    const void* QKV_bf16;
    const void* C_bf16;
    const float* w;
    const float* dout;                   /* [n_seq][d] fp32                                              */
    /* gradients */
    float* dWqkv_ext;                    /* [3d][ldx] (+=)  column d = d(bias)                           */
    float* dWa_ext;                      /* [q][ldx]  (+=)  column d = d(bias)                           */
    float* dqv;                          /* [q] (+=)                                                     */
    float* demb;                         /* [V][d] (+=) embedding gradient (ids variant)                 */
    float* ddense;                       /* [n_seq*T][d] (=) input gradient (dense variant)              */
    void* workspace;
    long long workspace_bytes;
    /* QKV_bf16 == NULL (the fused forward keeps Q|K|V on chip): recomputed here from X_bf16 with these operands */
    const void* wqkv_bf16;               /* [3d][ldx]                                                    */
    const float* bqkv;                   /* [3d]                                                         */
    /* optional cudaEvent_t recorded on `stream` as soon as demb is complete (before the weight-gradient GEMM): a data-
     * parallel caller starts the embedding-gradient all-reduce on a side stream that waits for it */
    void* emb_grad_ready_event;
} nr_mhsa_encoder_bwd_args;
long long nr_mhsa_encoder_bwd_workspace(long long n_seq, int T, int d, int q);
int nr_mhsa_encoder_bwd(const nr_mhsa_encoder_bwd_args* a, void* stream);

/* ---- reference: title / abstract CNN encoder ------------------------------------------------------------
 *   NAML  TextEncoder      src/model/NAML/news_encoder.py:21-37
 *   LSTUR title branch     src/model/LSTUR/news_encoder.py:56-72
 *   TANR  NewsEncoder      src/model/TANR/news_encoder.py:40-52
 * embedding -> dropout -> Conv2d(1, F, (3, d), padding (1, 0)) -> ReLU -> dropout -> additive pooling.
 * The conv is three row-shifted tcgen05 GEMM taps over a zero-padded layout (T+2 rows per segment). */
typedef struct {
    long long n_seq;
    int T, d, F, q, ldx, ldf;       /* ldx = round_up(d+1, 8), ldf = round_up(F+1, 8)                       */
    const long long* ids;           /* [n_seq*T]                                                          */
    const void* table_bf16;         /* [V][ldx]                                                           */
    int V;
    const void* wconv_bf16;         /* [3*F][ldx]; rows s*F..(s+1)*F hold tap s = weight[:, 0, s, :]        */
    const float* bconv;             /* [F]                                                                */
    const void* wa_bf16;            /* [q][ldf]                                                           */
    const float* ba;
    const float* qv;
    float p_drop;
    unsigned long long seed;
    void* Xp_bf16;                  /* [n_seq*(T+2)][ldx]  gathered rows, zero-padded layout (saved)        */
    void* Y_bf16;                   /* [n_seq*T][ldf]      relu(conv) rows (saved)                          */
    float* w;                       /* [n_seq*T]                                                          */
    float* out;                     /* [n_seq][F]                                                         */
    int* bad_id_flag;
    void* Y_lo_bf16;          /* optional [n_seq*T][ldf]: low plane of the conv output (accurate mode: the pooled sum reads Y + Y_lo) */
} nr_cnn_encoder_fwd_args;
int nr_cnn_encoder_fwd(const nr_cnn_encoder_fwd_args* a, void* stream);

typedef struct {
    long long n_seq;
    int T, d, F, q, ldx, ldf, ldq;
    const long long* ids;
    int V;
    const void* wconvT_bf16;        /* [3*d][ldf]; rows s*d..(s+1)*d hold (weight[:, 0, 2-s, :])^T          */
    const void* wa_bf16;            /* [q][ldf]                                                           */
    const void* waT_bf16;           /* [F][ldq]                                                           */
    const float* ba;
    const float* qv;
    float p_drop;
    unsigned long long seed;
    const void* Xp_bf16;
    const void* Y_bf16;
    const float* w;
    const float* dout;              /* [n_seq][F]                                                         */
    float* dWconv_ext;              /* [3][F][ldx] (+=); column d of tap 1 is d(bias)                       */
    float* dWa_ext;                 /* [q][ldf] (+=); column F is d(bias)                                   */
    float* dqv;                     /* [q] (+=)                                                           */
    float* demb;                    /* [V][d] (+=)                                                        */
    void* workspace;
    long long workspace_bytes;
} nr_cnn_encoder_bwd_args;
long long nr_cnn_encoder_bwd_workspace(long long n_seq, int T, int F, int q);
int nr_cnn_encoder_bwd(const nr_cnn_encoder_bwd_args* a, void* stream);

/* ---- generic Linear over dense fp32 rows (TANR topic predictor src/model/TANR/__init__.py:58-61, GRU
 * projections).  fwd: X_bf16 = bf16(x | 1) (saved), out = act(X W^T + b).  bwd: dW_ext[N][ldx] += dY^T [X|1]
 * (column K = d(bias)), dx = dY W (optional).  relu_out masks dy with (relu_out > 0). */
int nr_linear_rows_fwd(const float* x, long long n, int K, long long s_row, long long s_col, void* X_bf16, int ldx,
                       const void* W_bf16, int N, int ldw, const float* bias, int relu, float* out, int ld_out,
                       void* stream);
int nr_linear_rows_bwd(const float* dy, const float* relu_out, long long n, int N, int ld_dy, void* dY_bf16, int ldn,
                       const void* X_bf16, int K, int ldx, const void* WT_bf16, int ldwT, float* dW_ext, float* dx,
                       int ld_dx, void* stream);

/* ---- fp32 embedding lookups (LSTUR category / user embeddings, src/model/LSTUR/news_encoder.py:47-53) ---- */
int nr_embedding_f32_fwd(const long long* ids, long long n, const float* table, int V, int D, float* out,
                         int* bad_id_flag, void* stream);
/* ids outside [1, V) contribute nothing (row 0 = padding_idx; out-of-range ids are flagged by the forward lookup) */
int nr_embedding_f32_bwd(const long long* ids, long long n, const float* dout, int V, int D, float* dtable, void* stream);

/* ---- reference: NAML ElementEncoder  relu(Linear(embedding(id)))  (src/model/NAML/news_encoder.py:40-47) --- */
int nr_element_encoder_fwd(const long long* ids, long long n, const void* table_bf16, int V, int E, int lde,
                           void* E_bf16, const void* W_bf16, int F, const float* bias, float* out, int* bad_id_flag,
                           void* stream);
int nr_element_encoder_bwd(const long long* ids, long long n, const float* dout, const float* out, int F, void* dY_bf16,
                           int ldf, const void* E_bf16, int E, int lde, const void* WT_bf16, float* dW_ext,
                           float* dtable, int V, void* stream);

/* ---- reference: LSTUR UserEncoder -- pack_padded_sequence + nn.GRU, last hidden state -----------------
 * (src/model/LSTUR/user_encoder.py:16-45).  Gate order r, z, n; user b consumes the FIRST len[b] positions of
 * its (left-padded) history (reference quirk kept as-is); len 0 is clamped to 1 (user_encoder.py:27).
 * ldd = round_up(D+1, 8), ldh = round_up(Hd+1, 8), ldg = round_up(3Hd, 4), ldb = round_up(3Hd+1, 8). */
typedef struct {
    int B, S, D, Hd;
    const float* x;                 /* fp32 [B][S][D] clicked-news vectors with element strides below     */
    long long x_s_b, x_s_t, x_s_c;
    const long long* len;           /* [B] int64 (device)                                                */
    const float* h0;                /* [B][Hd] initial hidden state (user embedding for 'ini', zeros for 'con') */
    const void* wih_bf16;           /* [3Hd][ldd]                                                        */
    const void* whh_bf16;           /* [3Hd][ldh]                                                        */
    const float* bih;
    const float* bhh;
    /* saved for backward (caller-allocated) */
    void* xb;                       /* bf16 [B*S][ldd]                                                   */
    float* gi;                      /* fp32 [B*S][ldg]   input projections, rows b*S+t                    */
    float* gh;                      /* fp32 [S][B][ldg]  recurrent projections                            */
    float* hs;                      /* fp32 [S+1][B][Hd] hidden states                                    */
    void* hb;                       /* bf16 [S+1][B][ldh]                                                 */
    float* out;                     /* fp32 [B][Hd] last hidden state                                     */
    /* accurate mode (non-NULL): the input enters the projection as a hi/lo bf16 pair, gi = x_hi.W^T + b + x_lo.W^T (two passes) */
    void* x_lo_bf16;                /* [B*S][ldd] workspace: bf16(x - bf16(x))                                         */
} nr_gru_fwd_args;
int nr_gru_fwd(const nr_gru_fwd_args* a, void* stream);
/* 1 if nr_gru_fwd runs the whole recurrence as ONE cooperative launch for this shape on this device (users in 128-row tiles x
 * hidden units in slices of 32, one CTA each, all resident: tiles * slices <= SM count; Hd % 4 == 0, Hd <= 960); else it
 * runs three launches per step */
int nr_gru_persistent_supported(int B, int Hd);

typedef struct {
    int B, S, D, Hd;
    const long long* len;
    const void* wihT_bf16;          /* [D][ldb]  = W_ih^T                                                */
    const void* whhT_bf16;          /* [Hd][ldb] = W_hh^T                                                */
    const void* xb;
    const float* gi;
    const float* gh;
    const float* hs;
    const void* hb;
    const float* dout;              /* [B][Hd]                                                           */
    float* dWih_ext;                /* [3Hd][ldd] (+=), column D  = d(bias_ih)                            */
    float* dWhh_ext;                /* [3Hd][ldh] (+=), column Hd = d(bias_hh)                            */
    float* dx;                      /* [B*S][D] (=)                                                      */
    float* dh0;                     /* [B][Hd] (=)                                                       */
    void* workspace;
    long long workspace_bytes;
} nr_gru_bwd_args;
long long nr_gru_bwd_workspace(int B, int S, int D, int Hd);
int nr_gru_bwd(const nr_gru_bwd_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEWSREC_B200_H */
