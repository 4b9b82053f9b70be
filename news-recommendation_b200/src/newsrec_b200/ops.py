"""torch.autograd.Function wrappers over the C ABI.

Each Function allocates its buffers with torch (the caller owns every buffer, see the header), passes raw
device pointers + the current CUDA stream to one composite C call, and returns torch tensors.  No
arithmetic happens here; `torch.cat`/slicing of parameters is layout plumbing only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import (MhsaEncoderBwdArgs, MhsaEncoderFwdArgs, NewsrecError, check, load_library, require_cuda)


def ru8(x: int) -> int:
    return (x + 7) // 8 * 8


def ru16(x: int) -> int:
    return (x + 15) // 16 * 16


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Data-parallel hook (ddp.FlatGradients): a torch.cuda.Event the news-encoder backward records as soon as the embedding
# gradient is complete, so that its all-reduce can start under the remaining backward kernels.  None = not armed.
grad_ready_hook = {"event": None, "recorded": False}

_seed_counter = [0x243F6A8885A308D3]


def _seed_from(counter: int) -> int:
    return (counter ^ (torch.initial_seed() & 0xFFFFFFFFFFFFFFFF) ^ (int(os.environ.get("RANK", "0")) << 48)) & 0xFFFFFFFFFFFFFFFF


def next_seed() -> int:
    """Per-call dropout seed (counter based, mixed with torch's CPU seed so manual_seed() reproduces runs)."""
    _seed_counter[0] = (_seed_counter[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return _seed_from(_seed_counter[0])


def peek_seeds(n: int = 1):
    """The next `n` values next_seed() will return, without advancing the counter (tests hand them to the oracle so
    that it applies exactly the masks the kernels are going to draw)."""
    c, out = _seed_counter[0], []
    for _ in range(n):
        c = (c * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out.append(_seed_from(c))
    return out


# ---------------------------------------------------------------------------------------------------
# bf16 operand cache: fp32 nn.Parameters -> padded bf16 tensor-core operands, rebuilt only when a
# parameter's version counter (bumped by optimizer.step / load_state_dict) or storage changes.
# ---------------------------------------------------------------------------------------------------
class OperandCache:
    def __init__(self):
        self._store = {}

    def get(self, name, params, builder):
        key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params)
        hit = self._store.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            val = builder(*[p.detach() for p in params])
        self._store[name] = (key, val)
        return val

    def clear(self):
        self._store.clear()

    def invalidate_operands(self):
        """Drop every entry derived from parameters (what an optimizer step does implicitly by bumping their version
        counters); parameter-independent workspaces stay.  bench.py calls it every step so that the timed step pays for
        the operand refresh of real training."""
        for name in [n for n, (key, _) in self._store.items() if key]:
            del self._store[name]


DIRECT_GRAD_MARK = "_newsrec_direct_grad"


def grad_sink(p):
    """The parameter's own gradient storage if the kernels may accumulate into it directly, else None.
    Direct accumulation bypasses AccumulateGrad (tensor / DDP hooks do not fire, torch.autograd.grad returns nothing for the
    parameter), so it is an explicit opt-in: the owner of the gradient storage marks it (ddp.FlatGradients does); a .grad that
    merely exists -- e.g. after optimizer.zero_grad(set_to_none=False) -- takes the ordinary return path."""
    g = getattr(p, "grad", None)
    if g is None or not getattr(g, DIRECT_GRAD_MARK, False):
        return None
    if not p.requires_grad or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device \
            or g.shape != p.shape or g.data_ptr() % 16 != 0:  # 16-byte vector reductions (red.global.add.v4.f32)
        return None
    return g


def cast_pad(src: torch.Tensor, ld: int, transpose: bool = False) -> torch.Tensor:
    """fp32 [R][C] -> zero padded bf16 [R][ld] (or the transpose [C][ld]) on the device."""
    lib = load_library()
    src = src.contiguous().float()
    R, Cc = src.shape
    rows = Cc if transpose else R
    dst = torch.empty((rows, ld), dtype=torch.bfloat16, device=src.device)
    check(lib.nr_cast_pad_bf16(_p(src), R, Cc, Cc, _p(dst), ld, int(transpose), _stream()), "nr_cast_pad_bf16")
    return dst


def precision_mode(config):
    """"fast" | "accurate" | "fused" from a model config (config.py: precision, fused_news_encoder)."""
    if bool(getattr(config, "fused_news_encoder", False)):
        return "fused"
    mode = str(getattr(config, "precision", "fast"))
    if mode not in ("fast", "accurate", "fused"):
        raise NewsrecError(f"config.precision must be 'fast', 'accurate' or 'fused' (got {mode!r})")
    return mode


def qkv_pitches(d):
    """(sec, ld3) of the projected rows Q | K | V: sections at columns 0, sec, 2*sec with sec = round_up(d, 8) so that every
    section has the same 16-byte phase (abi.cu qkv_section); row pitch ld3 = round_up(3*sec, 16)."""
    sec = ru8(d)
    return sec, ru16(3 * sec)


def stack_qkv(mq, mk, mv):
    """W_Q | W_K | W_V (or the biases) stacked on dim 0 with zero rows up to the section stride after each."""
    d = mq.shape[0]
    pad = qkv_pitches(d)[0] - d
    parts = []
    for m in (mq, mk, mv):
        m = m.float()
        parts.append(m)
        if pad:
            parts.append(m.new_zeros((pad,) + tuple(m.shape[1:])))
    return torch.cat(parts, dim=0)


def cast_pad_many(specs):
    """[(fp32 [R][C] tensor, ld, transpose), ...] (at most 8) -> the zero padded bf16 operands, ONE launch for all of them."""
    lib = load_library()
    n = len(specs)
    srcs = [t.contiguous().float() for t, _, _ in specs]
    dsts = [torch.empty(((s.shape[1] if tr else s.shape[0]), ld), dtype=torch.bfloat16, device=s.device) for s, (_, ld, tr) in zip(srcs, specs)]
    arr_i = lambda vals: (C.c_int * n)(*vals)
    check(lib.nr_cast_pad_bf16_many(n, (C.c_void_p * n)(*[s.data_ptr() for s in srcs]), arr_i([s.shape[0] for s in srcs]),
                                    arr_i([s.shape[1] for s in srcs]), arr_i([s.shape[1] for s in srcs]),
                                    (C.c_void_p * n)(*[t.data_ptr() for t in dsts]), arr_i([ld for _, ld, _ in specs]),
                                    arr_i([int(tr) for _, _, tr in specs]), _stream()), "nr_cast_pad_bf16_many")
    return dsts


def mhsa_operands(cache: OperandCache, prefix, Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv):
    d, q = Wq.shape[0], Wa.shape[0]
    ldx, ldq = ru8(d + 1), ru16(q)
    ld3 = qkv_pitches(d)[1]

    def build(Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv):
        wqkv = stack_qkv(Wq, Wk, Wv)
        w, wT, wa, waT = cast_pad_many([(wqkv, ldx, False), (wqkv, ld3, True), (Wa, ldx, False), (Wa, ldq, True)])
        return dict(wqkv=w, wqkvT=wT, bqkv=stack_qkv(bq, bk, bv).contiguous(), wa=wa, waT=waT,
                    ba=ba.float().contiguous(), qv=qv.float().contiguous())

    return cache.get(prefix, (Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv), build)


def pack_head_blocks(Wq, bq, Wk, bk, Wv, bv, heads, ldx, rows_per_head=64):
    """Per-head weight blocks of the fused front end: rows W_Q[h] | W_K[h] | W_V[h] | zero rows up to `rows_per_head`
    (head h owns output features [h*d_k, (h+1)*d_k) of each projection, multihead_self.py:53-58).  Layout plumbing only."""
    d = Wq.shape[0]
    dk = d // heads
    W = torch.zeros((heads, rows_per_head, d), dtype=torch.float32, device=Wq.device)
    b = torch.zeros((heads, rows_per_head), dtype=torch.float32, device=Wq.device)
    for i, (Wm, bm) in enumerate(((Wq, bq), (Wk, bk), (Wv, bv))):
        W[:, i * dk:(i + 1) * dk] = Wm.float().view(heads, dk, d)
        b[:, i * dk:(i + 1) * dk] = bm.float().view(heads, dk)
    return cast_pad(W.view(heads * rows_per_head, d), ldx), b.view(-1).contiguous()


def table_operand(cache: OperandCache, name, weight):
    ldx = ru8(weight.shape[1] + 1)
    return cache.get(name, (weight,), lambda w: cast_pad(w, ldx))


# ---------------------------------------------------------------------------------------------------
# NRMS NewsEncoder / UserEncoder: gather|dense -> MHSA -> additive pooling in ONE C call each way
# ---------------------------------------------------------------------------------------------------
class MhsaPoolEncoderFn(torch.autograd.Function):
    """forward(ids|None, dense|None, emb_weight|None, Wq,bq,Wk,bk,Wv,bv, Wa,ba,qv, heads, p_drop, cache, prefix)

    ids   : int64 (n_seq, T) device tensor  (news encoder)   -- reference src/model/NRMS/news_encoder.py:27-48
    dense : fp32  (n_seq, T, d) any strides (user encoder)   -- reference src/model/NRMS/user_encoder.py:15-26
    """

    @staticmethod
    def forward(ctx, ids, dense, emb_w, Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv, heads, p_drop, cache, prefix, bad_flag, precise=False):
        lib = load_library()
        dev = require_cuda()
        d, q = Wq.shape[0], Wa.shape[0]
        ldx = ru8(d + 1)
        sec, ld3 = qkv_pitches(d)
        ops = mhsa_operands(cache, prefix, Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv)
        a = MhsaEncoderFwdArgs()
        if ids is not None:
            n_seq, T = ids.shape
            ids = ids.contiguous()
            table = table_operand(cache, prefix + ".table", emb_w)
            a.ids, a.table_bf16, a.V = _p(ids), _p(table), emb_w.shape[0]
            a.dense = None
        else:
            n_seq, T, dd = dense.shape
            if dd != d:
                raise NewsrecError(f"user encoder input width {dd} != model width {d}")
            dense = dense.float()
            a.ids, a.table_bf16, a.V = None, None, 0
            a.dense = _p(dense)
            a.dense_s_seq, a.dense_s_tok, a.dense_s_col = dense.stride()
            table = None
        n_tok = n_seq * T
        need_bwd = any(ctx.needs_input_grad)
        # the fused front end (gather -> Q|K|V -> attention in ONE kernel, V / context as hi/lo bf16 pairs) is the PRECISE
        # mode of the news level: 2.6e-3 instead of 7e-3 against the reference's fp32 logits, at ~3.6x the time of the
        # unfused gather | GEMM | attention sequence -- opt-in (config.fused_news_encoder / NEWSREC_FUSED=1)
        # precision modes (config.precision):
        #   "fast"      bf16 storage of every activation (Q|K|V, probabilities, context): fastest, ~6e-3 from the fp32 result
        #   "accurate"  V / probabilities / context as hi/lo bf16 pairs on the same unfused kernels + fp32-accurate user
        #               encoder: within the blueprint's 1e-3 of the fp32 oracle on bf16-rounded weights
        #   "fused"     the one-kernel news front end (same numerics as "accurate", kept for reference; slower)
        mode = precise if isinstance(precise, str) else ("fused" if precise else "fast")
        if os.environ.get("NEWSREC_FUSED") == "1":
            mode = "fused"
        fused = ids is not None and mode == "fused" and bool(lib.nr_mhsa_fused_supported(T, d, heads))
        accurate = ids is not None and not fused and mode in ("accurate", "fused") and bool(lib.nr_mhsa_accurate_supported(T, d, heads))
        precise_dense = ids is None and mode in ("accurate", "fused")  # user encoder: fp32-accurate forward (abi.cu)
        X = QKV = C_lo = V_lo = None
        if need_bwd or not fused:  # X only exists in HBM when a backward pass (or the unfused sequence) reads it
            X = torch.empty((n_tok, ldx), dtype=torch.bfloat16, device=dev)
        if not fused and not precise_dense:  # the precise paths keep no bf16 Q|K|V; their backward recomputes it from X
            QKV = torch.empty((n_tok, ld3), dtype=torch.bfloat16, device=dev)
        Cx = torch.empty((n_tok, ldx), dtype=torch.bfloat16, device=dev)
        w = torch.empty((n_tok,), dtype=torch.float32, device=dev)
        out = torch.empty((n_seq, d), dtype=torch.float32, device=dev)
        seed = next_seed() if p_drop > 0 else 0
        a.n_seq, a.T, a.d, a.heads, a.q, a.ldx, a.ld3 = n_seq, T, d, heads, q, ldx, ld3
        a.wqkv_bf16, a.bqkv, a.wa_bf16, a.ba, a.qv = _p(ops["wqkv"]), _p(ops["bqkv"]), _p(ops["wa"]), _p(ops["ba"]), _p(ops["qv"])
        a.p_drop, a.seed = float(p_drop), seed
        if fused:
            hb = cache.get(prefix + ".heads", (Wq, bq, Wk, bk, Wv, bv),
                           lambda Wq, bq, Wk, bk, Wv, bv: pack_head_blocks(Wq, bq, Wk, bk, Wv, bv, heads, ldx))
            C_lo = torch.empty((n_tok, ldx), dtype=torch.bfloat16, device=dev)
            a.wqkv_heads_bf16, a.bqkv_heads, a.C_lo_bf16 = _p(hb[0]), _p(hb[1]), _p(C_lo)
        if accurate:
            C_lo = torch.empty((n_tok, ldx), dtype=torch.bfloat16, device=dev)
            V_lo = torch.empty((n_tok, sec), dtype=torch.bfloat16, device=dev)
            a.C_lo_bf16, a.V_lo_bf16 = _p(C_lo), _p(V_lo)
        keep = None
        if precise_dense:
            kcat = cache.get(prefix + ".kcat", (Wq, Wk, Wv), lambda Wq, Wk, Wv: cast_pad(
                torch.cat((torch.nn.functional.pad(stack_qkv(Wq, Wk, Wv), (0, ldx - d)),) * 2, dim=1), 2 * ldx))
            C_lo = torch.empty((n_tok, ldx), dtype=torch.bfloat16, device=dev)
            keep = (torch.empty((n_tok, 2 * ldx), dtype=torch.bfloat16, device=dev), torch.empty((n_tok, 3 * sec), dtype=torch.float32, device=dev))
            a.wqkv_kcat_bf16, a.X_kcat_bf16, a.QKV_f32, a.C_lo_bf16 = _p(kcat), _p(keep[0]), _p(keep[1]), _p(C_lo)
        a.X_bf16, a.QKV_bf16, a.C_bf16, a.w, a.out = _p(X), _p(QKV), _p(Cx), _p(w), _p(out)
        a.bad_id_flag = _p(bad_flag)
        check(lib.nr_mhsa_encoder_fwd(C.byref(a), _stream()), "nr_mhsa_encoder_fwd")
        if need_bwd:
            ctx.save_for_backward(X, QKV if QKV is not None else torch.empty(0, device=dev), Cx, w,
                                  ids if ids is not None else torch.empty(0, device=dev))
        ctx.meta = dict(n_seq=n_seq, T=T, d=d, q=q, heads=heads, p_drop=float(p_drop), seed=seed, ops=ops,
                        has_ids=ids is not None, V=emb_w.shape[0] if ids is not None else 0,
                        dense_shape=None if dense is None else tuple(dense.shape),
                        params=(emb_w, Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv), cache=cache, prefix=prefix)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load_library()
        X, QKV, Cx, w, ids = ctx.saved_tensors
        m = ctx.meta
        dev = X.device
        d, q, T, n_seq = m["d"], m["q"], m["T"], m["n_seq"]
        ldx, ldq = ru8(d + 1), ru16(q)
        sec, ld3 = qkv_pitches(d)
        ops = m["ops"]
        dout = dout.contiguous().float()
        emb_w, Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv = m["params"]
        # Parameters whose .grad storage is marked for direct accumulation (ddp.FlatGradients; see grad_sink) are
        # accumulated IN PLACE by the kernels and get None from this Function: no zero fill, no slice
        # copies, no AccumulateGrad adds (together ~50 small framework kernels and 3 passes over the 85 MB embedding
        # gradient per step).  Anything else takes the allocate-and-return path.
        sinks = [grad_sink(t) for t in (Wq, bq, Wk, bk, Wv, bv, Wa, ba, qv)]
        direct = all(g is not None for g in sinks)
        if direct:
            ws_grads = m["cache"].get(m["prefix"] + ".grad_ws", (), lambda: dict(
                dWqkv=torch.zeros((3 * sec, ldx), dtype=torch.float32, device=dev),
                dWa=torch.zeros((q, ldx), dtype=torch.float32, device=dev)))
            dWqkv, dWa, dqv = ws_grads["dWqkv"], ws_grads["dWa"], sinks[8]
        else:
            dWqkv = torch.zeros((3 * sec, ldx), dtype=torch.float32, device=dev)
            dWa = torch.zeros((q, ldx), dtype=torch.float32, device=dev)
            dqv = torch.zeros((q,), dtype=torch.float32, device=dev)
        demb = ddense = None
        emb_direct = False
        if m["has_ids"]:
            demb = grad_sink(emb_w)
            emb_direct = demb is not None
            if not emb_direct:
                demb = torch.zeros((m["V"], d), dtype=torch.float32, device=dev)
        else:
            ddense = torch.empty((n_seq * T, d), dtype=torch.float32, device=dev)
        ws_bytes = int(lib.nr_mhsa_encoder_bwd_workspace(n_seq, T, d, q))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        a = MhsaEncoderBwdArgs()
        a.n_seq, a.T, a.d, a.heads, a.q, a.ldx, a.ld3, a.ldq = n_seq, T, d, m["heads"], q, ldx, ld3, ldq
        a.ids = _p(ids) if m["has_ids"] else None
        a.V = m["V"]
        a.wqkvT_bf16, a.wa_bf16, a.waT_bf16, a.ba, a.qv = _p(ops["wqkvT"]), _p(ops["wa"]), _p(ops["waT"]), _p(ops["ba"]), _p(ops["qv"])
        a.p_drop, a.seed = m["p_drop"], m["seed"]
        a.X_bf16, a.QKV_bf16, a.C_bf16, a.w, a.dout = _p(X), (_p(QKV) if QKV.numel() else None), _p(Cx), _p(w), _p(dout)
        a.wqkv_bf16, a.bqkv = _p(ops["wqkv"]), _p(ops["bqkv"])
        ev = grad_ready_hook["event"] if (m["has_ids"] and emb_direct) else None
        if ev is not None:
            a.emb_grad_ready_event = C.c_void_p(ev.cuda_event)
            grad_ready_hook["recorded"] = True
        a.dWqkv_ext, a.dWa_ext, a.dqv = _p(dWqkv), _p(dWa), _p(dqv)
        a.demb, a.ddense = _p(demb), _p(ddense)
        a.workspace, a.workspace_bytes = _p(ws), ws_bytes
        check(lib.nr_mhsa_encoder_bwd(C.byref(a), _stream()), "nr_mhsa_encoder_bwd")
        g_dense = ddense.view(m["dense_shape"]) if ddense is not None else None
        g_emb = None if emb_direct else demb
        if direct:
            for i in range(3):
                check(lib.nr_accumulate_ext_grad(_p(dWqkv[i * sec:i * sec + d]), d, ldx, d, _p(sinks[2 * i]), _p(sinks[2 * i + 1]),
                                                 _stream()), "nr_accumulate_ext_grad")
            check(lib.nr_accumulate_ext_grad(_p(dWa), q, ldx, d, _p(sinks[6]), _p(sinks[7]), _stream()), "nr_accumulate_ext_grad")
            return (None, g_dense, g_emb) + (None,) * 15
        gW = [dWqkv[i * sec:i * sec + d, :d].contiguous() for i in range(3)]
        gb = [dWqkv[i * sec:i * sec + d, d].contiguous() for i in range(3)]
        return (None, g_dense, g_emb, gW[0], gb[0], gW[1], gb[1], gW[2], gb[2],
                dWa[:, :d].contiguous(), dWa[:, d].contiguous(), dqv, None, None, None, None, None, None)


# ---------------------------------------------------------------------------------------------------
# AdditiveAttention over dense fp32 rows (NAML / TANR user encoder, NAML 4-view fusion, standalone module)
# ---------------------------------------------------------------------------------------------------
class AdditiveAttentionFn(torch.autograd.Function):
    """reference src/model/general/attention/additive.py:27-53;  x (N, S, D) fp32 -> (N, D)."""

    @staticmethod
    def forward(ctx, x, Wa, ba, qv, cache, prefix):
        lib = load_library()
        dev = require_cuda()
        N, S, D = x.shape
        q = Wa.shape[0]
        ldx, ldq = ru8(D + 1), ru16(q)
        ops = cache.get(prefix, (Wa, ba, qv), lambda Wa, ba, qv: dict(
            wa=cast_pad(Wa, ldx), waT=cast_pad(Wa, ldq, transpose=True), ba=ba.float().contiguous(),
            qv=qv.float().contiguous()))
        xf = x.float()
        X = torch.empty((N * S, ldx), dtype=torch.bfloat16, device=dev)
        xs = xf.reshape(N * S, D) if xf.is_contiguous() else xf.contiguous().view(N * S, D)
        check(lib.nr_rows_to_bf16(_p(xs), N * S, D, xs.stride(0), xs.stride(1), _p(X), ldx, _stream()), "nr_rows_to_bf16")
        out = torch.empty((N, D), dtype=torch.float32, device=dev)
        w = torch.empty((N * S,), dtype=torch.float32, device=dev)
        check(lib.nr_additive_attention_fwd(_p(X), N, S, D, ldx, _p(ops["wa"]), q, ldx, _p(ops["ba"]), _p(ops["qv"]),
                                            _p(out), D, _p(w), _stream()), "nr_additive_attention_fwd")
        ctx.save_for_backward(X, w)
        ctx.meta = dict(N=N, S=S, D=D, q=q, ops=ops)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load_library()
        X, w = ctx.saved_tensors
        m = ctx.meta
        N, S, D, q, ops = m["N"], m["S"], m["D"], m["q"], m["ops"]
        dev = X.device
        ldx, ldq = ru8(D + 1), ru16(q)
        dout = dout.contiguous().float()
        dX = torch.empty((N * S, ldx), dtype=torch.bfloat16, device=dev)
        dWa = torch.zeros((q, ldx), dtype=torch.float32, device=dev)
        dqv = torch.zeros((q,), dtype=torch.float32, device=dev)
        ws_bytes = int(lib.nr_additive_attention_bwd_workspace(N, S, q))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        check(lib.nr_additive_attention_bwd(_p(X), N, S, D, ldx, _p(ops["wa"]), _p(ops["waT"]), q, ldx, ldq, _p(ops["ba"]),
                                            _p(ops["qv"]), _p(w), _p(dout), D, _p(dX), ldx, _p(dWa), _p(dqv), _p(ws),
                                            ws_bytes, _stream()), "nr_additive_attention_bwd")
        gx = dX[:, :D].float().view(N, S, D)
        return gx, dWa[:, :D].contiguous(), dWa[:, D].contiguous(), dqv, None, None


# ---------------------------------------------------------------------------------------------------
# DotProductClickPredictor
# ---------------------------------------------------------------------------------------------------
class DotScoreFn(torch.autograd.Function):
    """reference src/model/general/click_predictor/dot_product.py:8-19;  (B,C,D),(B,D) -> (B,C) logits."""

    @staticmethod
    def forward(ctx, cand, user):
        lib = load_library()
        dev = require_cuda()
        cand = cand.contiguous().float()
        user = user.contiguous().float()
        B, Cn, D = cand.shape
        logits = torch.empty((B, Cn), dtype=torch.float32, device=dev)
        check(lib.nr_dot_score_fwd(_p(cand), _p(user), B, Cn, D, _p(logits), _stream()), "nr_dot_score_fwd")
        ctx.save_for_backward(cand, user)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = load_library()
        cand, user = ctx.saved_tensors
        B, Cn, D = cand.shape
        dlogits = dlogits.contiguous().float()
        dcand = torch.empty_like(cand)
        duser = torch.empty_like(user)
        check(lib.nr_dot_score_bwd(_p(cand), _p(user), _p(dlogits), B, Cn, D, _p(dcand), _p(duser), _stream()),
              "nr_dot_score_bwd")
        return dcand, duser


# ---------------------------------------------------------------------------------------------------
# standalone MultiHeadSelfAttention (projection + attention core, no pooling)
# ---------------------------------------------------------------------------------------------------
class MhsaFn(torch.autograd.Function):
    """reference src/model/general/attention/multihead_self.py:46-76 with Q=K=V=x, length=None."""

    @staticmethod
    def forward(ctx, x, Wq, bq, Wk, bk, Wv, bv, heads, cache, prefix):
        lib = load_library()
        dev = require_cuda()
        N, T, d = x.shape
        ldx, ld3 = ru8(d + 1), ru16(3 * d)

        def build(Wq, bq, Wk, bk, Wv, bv):
            wqkv = torch.cat((Wq, Wk, Wv), dim=0)
            return dict(wqkv=cast_pad(wqkv, ldx), wqkvT=cast_pad(wqkv, ld3, transpose=True),
                        bqkv=torch.cat((bq, bk, bv)).float().contiguous())

        ops = cache.get(prefix, (Wq, bq, Wk, bk, Wv, bv), build)
        xs = x.float().contiguous().view(N * T, d)
        X = torch.empty((N * T, ldx), dtype=torch.bfloat16, device=dev)
        check(lib.nr_rows_to_bf16(_p(xs), N * T, d, d, 1, _p(X), ldx, _stream()), "nr_rows_to_bf16")
        QKV = torch.empty((N * T, ld3), dtype=torch.bfloat16, device=dev)
        check(lib.nr_linear(_p(X), N * T, ldx, _p(ops["wqkv"]), 3 * d, ldx, d, 1, 0, 128, _p(ops["bqkv"]), 0, _p(QKV), ld3, 1,
                            _stream()), "nr_linear")
        Cx = torch.empty((N * T, ldx), dtype=torch.bfloat16, device=dev)
        check(lib.nr_mhsa_core_fwd(_p(QKV), ld3, d, N, T, heads, d // heads, _p(Cx), ldx, 0.0, 0, _stream()), "nr_mhsa_core_fwd")
        ctx.save_for_backward(X, QKV)
        ctx.meta = dict(N=N, T=T, d=d, heads=heads, ops=ops)
        return Cx[:, :d].float().view(N, T, d)

    @staticmethod
    def backward(ctx, dctx):
        lib = load_library()
        X, QKV = ctx.saved_tensors
        m = ctx.meta
        N, T, d, heads, ops = m["N"], m["T"], m["d"], m["heads"], m["ops"]
        dev = X.device
        ldx, ld3 = ru8(d + 1), ru16(3 * d)
        g = dctx.float().contiguous().view(N * T, d)
        dC = torch.empty((N * T, ldx), dtype=torch.bfloat16, device=dev)
        check(lib.nr_rows_to_bf16(_p(g), N * T, d, d, 1, _p(dC), ldx, _stream()), "nr_rows_to_bf16")
        dQKV = torch.empty((N * T, ld3), dtype=torch.bfloat16, device=dev)
        check(lib.nr_mhsa_core_bwd(_p(QKV), ld3, d, _p(dC), ldx, N, T, heads, d // heads, _p(dQKV), ld3, _stream()),
              "nr_mhsa_core_bwd")
        dW = torch.zeros((3 * d, ldx), dtype=torch.float32, device=dev)
        check(lib.nr_gemm_tn(_p(dQKV), N * T, 3 * d, ld3, _p(X), N * T, d + 1, ldx, 0, d + 1, 0, _p(dW), ldx, _stream()),
              "nr_gemm_tn")
        dx = torch.empty((N * T, d), dtype=torch.float32, device=dev)
        check(lib.nr_linear(_p(dQKV), N * T, ld3, _p(ops["wqkvT"]), d, ld3, 3 * d, 1, 0, 128, None, 0, _p(dx), d, 0, _stream()),
              "nr_linear")
        gW = [dW[i * d:(i + 1) * d, :d].contiguous() for i in range(3)]
        gb = [dW[i * d:(i + 1) * d, d].contiguous() for i in range(3)]
        return dx.view(N, T, d), gW[0], gb[0], gW[1], gb[1], gW[2], gb[2], None, None, None


# ---------------------------------------------------------------------------------------------------
# Batched scoring for evaluation (the "next" row N2 of SURVEY.md 8f)
# ---------------------------------------------------------------------------------------------------
def predict_impressions(news_matrix, cand_index, seg_offsets, user_vectors):
    """Scores of MANY impressions in one launch: reference src/evaluate.py:245-265 runs `get_prediction` once per impression
    and synchronises on `.tolist()` after each.  news_matrix (n_news, D) fp32 device matrix of news vectors (row = news
    index, instead of the evaluator's dict of rows), cand_index (n_cand,) int64 rows of the candidates of all impressions
    back to back, seg_offsets (n_impressions + 1,) int64 with seg_offsets[0] = 0, user_vectors (n_impressions, D).
    Returns (n_cand,) fp32 scores; candidate i of impression s is scores[seg_offsets[s] + i]."""
    lib = load_library()
    dev = require_cuda()
    news_matrix = news_matrix.to(dev).float().contiguous()
    user_vectors = user_vectors.to(dev).float().contiguous()
    cand_index = cand_index.to(dev).long().contiguous()
    seg_offsets = seg_offsets.to(dev).long().contiguous()
    n_seg = seg_offsets.numel() - 1
    if user_vectors.shape[0] != n_seg or user_vectors.shape[1] != news_matrix.shape[1]:
        raise NewsrecError("predict_impressions: user_vectors must be (len(seg_offsets) - 1, D)")
    scores = torch.empty((cand_index.numel(),), dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.nr_segment_dot(_p(news_matrix), news_matrix.shape[0], news_matrix.shape[1], _p(cand_index), cand_index.numel(),
                             _p(seg_offsets), n_seg, _p(user_vectors), _p(scores), _p(flag), _stream()), "nr_segment_dot")
    return scores
