"""Kernel-vs-oracle checks, written as plain functions returning error metrics so that both the pytest
wrappers (tests/test_gpu_*.py) and the triage ladder (tools/gpu_ladder.py) can run them.

Everything goes through the C ABI (ctypes) or through the drop-in model package that calls it.
The oracle (oracle/newsrec_oracle.py) runs on the CPU; it is the checker, never the thing measured.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

import newsrec_oracle as O
from newsrec_b200 import check, load_library
from newsrec_b200.ops import _p, _stream, cast_pad, ru8, ru16

DEV = "cuda"


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(b.norm().item(), 1e-30))


def maxabs(a, b) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def bf16r(x):
    return x.to(torch.bfloat16).to(torch.float32)


# ------------------------------------------------------------------------------------------------
def check_prep_and_gather():
    lib = load_library()
    V, D, T, n_seq = 97, 300, 20, 13
    ld = ru8(D + 1)
    w = O.det_uniform((V, D), 5)
    table = cast_pad(w.to(DEV), ld)
    ref = torch.zeros(V, ld)
    ref[:, :D] = bf16r(w)
    out = {"cast_pad_exact": bool(torch.equal(table.float().cpu(), ref))}
    wt = cast_pad(w.to(DEV), ru8(V), transpose=True)
    reft = torch.zeros(D, ru8(V))
    reft[:, :V] = bf16r(w).t()
    out["cast_pad_T_exact"] = bool(torch.equal(wt.float().cpu(), reft))
    ids = O.synth_titles(n_seq, T, V, 77).to(DEV)
    X = torch.full((n_seq * T, ld), 7.0, dtype=torch.bfloat16, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    check(lib.nr_gather_rows(_p(ids), n_seq * T, T, _p(table), V, D, ld, _p(X), 0, 0.0, 0, _p(flag), _stream()), "gather")
    expect = ref[ids.cpu().reshape(-1)]
    expect[:, D] = 1.0
    out["gather_exact"] = bool(torch.equal(X.float().cpu(), expect)) and int(flag.item()) == 0
    # padded CNN layout
    Xp = torch.full((n_seq * (T + 2), ld), 7.0, dtype=torch.bfloat16, device=DEV)
    check(lib.nr_gather_rows(_p(ids), n_seq * T, T, _p(table), V, D, ld, _p(Xp), 1, 0.0, 0, _p(flag), _stream()), "gather")
    Xp3 = Xp.float().cpu().view(n_seq, T + 2, ld)
    out["gather_padded_exact"] = bool(torch.equal(Xp3[:, 1:T + 1].reshape(-1, ld), expect)) and \
        float(Xp3[:, 0].abs().sum() + Xp3[:, T + 1].abs().sum()) == 0.0
    # out-of-range id sets the flag
    bad = ids.clone()
    bad[0, 0] = V + 5
    check(lib.nr_gather_rows(_p(bad), n_seq * T, T, _p(table), V, D, ld, _p(X), 0, 0.0, 0, _p(flag), _stream()), "gather")
    out["bad_id_flag"] = int(flag.item()) == 1
    # dropout: keep-rate and scaling
    big = O.synth_titles(2000, T, V, 78, min_len=T).to(DEV)
    Xd = torch.empty((2000 * T, ld), dtype=torch.bfloat16, device=DEV)
    flag.zero_()
    check(lib.nr_gather_rows(_p(big), 2000 * T, T, _p(table), V, D, ld, _p(Xd), 0, 0.2, 1234, _p(flag), _stream()), "gather")
    e = ref[big.cpu().reshape(-1)][:, :D]
    got = Xd.float().cpu()[:, :D]
    kept = got != 0
    out["dropout_keep_rate"] = float(kept.float().mean() / (e != 0).float().mean())
    ratio = (got[kept] / e[kept])
    out["dropout_scale_err"] = float((ratio - 1.25).abs().max())
    out["dropout_ones_col_intact"] = bool((Xd[:, D].float() == 1).all())
    return out


# ------------------------------------------------------------------------------------------------
def _rand_bf16(shape, seed, scale=1.0):
    return bf16r(O.det_uniform(shape, seed, -scale, scale))


def check_linear(M=300, N=900, K=300, taps=1, seg=0, relu=0, out_bf16=1):
    """nr_linear (tcgen05 gemm_nt + store epilogue) against an fp64 matmul of the same bf16 operands."""
    lib = load_library()
    lda, ldw = ru8(K + 1), ru8(K + 1)
    if taps == 1:
        A = _rand_bf16((M, K), 1)
        W = _rand_bf16((N, K), 2, 0.1)
        bias = O.det_uniform((N,), 3, -0.5, 0.5)
        ref = A.double() @ W.double().t() + bias.double()
        Ad = torch.zeros(M, lda)
        Ad[:, :K] = A
        Wd = torch.zeros(N, ldw)
        Wd[:, :K] = W
        rpt, w_tap_rows = 128, 0
    else:  # window-3 conv over the padded layout: seg tokens per segment, M = n_seg*(seg+2) rows
        n_seg = M
        Mrows = n_seg * (seg + 2)
        X = _rand_bf16((n_seg, seg, K), 1)
        W = _rand_bf16((N, taps, K), 2, 0.1)
        bias = O.det_uniform((N,), 3, -0.5, 0.5)
        Xp = torch.zeros(n_seg, seg + 2, K)
        Xp[:, 1:seg + 1] = X
        ref = torch.zeros(n_seg, seg + 2, N, dtype=torch.float64)
        for s in range(taps):
            sh = torch.zeros_like(Xp)
            lo, hi = max(0, 1 - s), min(seg + 2, seg + 3 - s)
            sh[:, lo:hi] = Xp[:, lo + s - 1:hi + s - 1]
            ref += sh.double() @ W[:, s].double().t()
        ref = (ref + bias.double()).view(Mrows, N)
        Ad = torch.zeros(Mrows, lda)
        Ad[:, :K] = Xp.view(Mrows, K)
        Wd = torch.zeros(taps * N, ldw)
        for s in range(taps):
            Wd[s * N:(s + 1) * N, :K] = W[:, s]
        M = Mrows
        rpt, w_tap_rows = (128 // (seg + 2)) * (seg + 2), N
    if relu:
        ref = ref.clamp(min=0)
    ld_out = ru8(N) if out_bf16 else (N + 3) // 4 * 4
    out = torch.full((M, ld_out), float("nan"), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=DEV)
    Ad, Wd, bd = Ad.to(torch.bfloat16).to(DEV), Wd.to(torch.bfloat16).to(DEV), bias.to(DEV)
    check(lib.nr_linear(_p(Ad), M, lda, _p(Wd), N, ldw, K, taps, w_tap_rows, rpt, _p(bd), relu, _p(out), ld_out, out_bf16,
                        _stream()), "nr_linear")
    torch.cuda.synchronize()
    got = out[:, :N].float().cpu()
    if taps > 1:  # pad rows of the padded layout are not part of the result
        keep = torch.ones(M, dtype=torch.bool).view(-1, seg + 2)
        keep[:, 0] = keep[:, -1] = False
        keep = keep.view(-1)
        got, ref = got[keep], ref[keep]
    return {"rel": relerr(got, ref), "maxabs": maxabs(got, ref), "nan": int(torch.isnan(got).sum())}


def check_gemm_tn(Kr=1000, Ma=900, Nb=301, shift=0):
    lib = load_library()
    lda, ldb = ru8(Ma), ru8(Nb)
    A = _rand_bf16((Kr, Ma), 11, 0.5)
    B = _rand_bf16((Kr, Nb), 12, 0.5)
    Bs = torch.zeros_like(B)
    if shift >= 0:
        Bs[:Kr - shift] = B[shift:]
    else:
        Bs[-shift:] = B[:Kr + shift]
    ref = A.double().t() @ Bs.double()
    Ad = torch.zeros(Kr, lda)
    Ad[:, :Ma] = A
    Bd = torch.zeros(Kr, ldb)
    Bd[:, :Nb] = B
    ldd = ru8(Nb)
    D = torch.full((Ma, ldd), 1.0, dtype=torch.float32, device=DEV)  # accumulate on top of ones
    Ad, Bd = Ad.to(torch.bfloat16).to(DEV), Bd.to(torch.bfloat16).to(DEV)
    check(lib.nr_gemm_tn(_p(Ad), Kr, Ma, lda, _p(Bd), Kr, Nb, ldb, 0, Nb, shift, _p(D), ldd, _stream()), "nr_gemm_tn")
    torch.cuda.synchronize()
    got = D[:, :Nb].cpu() - 1.0
    return {"rel": relerr(got, ref), "maxabs": maxabs(got, ref), "nan": int(torch.isnan(got).sum())}


# ------------------------------------------------------------------------------------------------
def check_mhsa_core(n_seq=7, T=20, heads=15, dk=20, sectioned=False):
    """sectioned: Q | K | V at columns 0, sec, 2*sec with sec = round_up(d, 8) (the encoders' layout; padding columns
    carry garbage on the way in and must come back as zeros in dQ|dK|dV), else dense sections (sec = d)."""
    lib = load_library()
    d = heads * dk
    sec = ru8(d) if sectioned else d
    ld3, ldx = ru16(3 * sec), ru8(d + 1)
    qkv = _rand_bf16((n_seq * T, 3 * d), 21, 1.5).requires_grad_(True)
    Q, K, V = [t.view(n_seq, T, heads, dk).transpose(1, 2) for t in qkv.split(d, dim=1)]
    ctx = O.scaled_dot_product_attention(Q, K, V, O.BF16).transpose(1, 2).reshape(n_seq * T, d)
    g = _rand_bf16((n_seq * T, d), 22)
    ctx.backward(g)
    qd = torch.zeros(n_seq * T, ld3)
    for i in range(3):
        qd[:, i * sec:i * sec + d] = qkv.detach()[:, i * d:(i + 1) * d]
    qd = qd.to(torch.bfloat16).to(DEV)
    cd = torch.full((n_seq * T, ldx), 9.0, dtype=torch.bfloat16, device=DEV)
    check(lib.nr_mhsa_core_fwd(_p(qd), ld3, sec, n_seq, T, heads, dk, _p(cd), ldx, 0.0, 0, _stream()), "mhsa_fwd")
    gd = torch.zeros(n_seq * T, ldx)
    gd[:, :d] = g
    gd = gd.to(torch.bfloat16).to(DEV)
    dq = torch.full((n_seq * T, ld3), 9.0, dtype=torch.bfloat16, device=DEV)
    check(lib.nr_mhsa_core_bwd(_p(qd), ld3, sec, _p(gd), ldx, n_seq, T, heads, dk, _p(dq), ld3, _stream()), "mhsa_bwd")
    torch.cuda.synchronize()
    c, dqc = cd.float().cpu(), dq.float().cpu()
    got = torch.cat([dqc[:, i * sec:i * sec + d] for i in range(3)], dim=1)
    pads = torch.cat([dqc[:, i * sec + d:(i + 1) * sec] for i in range(3)], dim=1)
    return {"fwd_rel": relerr(c[:, :d], bf16r(ctx.detach())), "ones_col": bool((c[:, d] == 1).all()),
            "bwd_rel": relerr(got, bf16r(qkv.grad)), "pad_zero": bool((pads == 0).all())}


def check_additive(N=37, S=20, D=300, q=200):
    from newsrec_b200.ops import AdditiveAttentionFn, OperandCache
    x = _rand_bf16((N, S, D), 31).requires_grad_(True)
    p = {"a.linear.weight": O.det_uniform((q, D), 32, -0.1, 0.1).requires_grad_(True),
         "a.linear.bias": O.det_uniform((q,), 33, -0.05, 0.05).requires_grad_(True),
         "a.attention_query_vector": O.det_uniform((q,), 34, -0.1, 0.1).requires_grad_(True)}
    ref = O.additive_attention(x, p, "a", O.BF16)
    g = O.det_uniform((N, D), 35)
    ref.backward(g)
    xd = x.detach().to(DEV).requires_grad_(True)
    pd = {k: v.detach().to(DEV).requires_grad_(True) for k, v in p.items()}
    out = AdditiveAttentionFn.apply(xd, pd["a.linear.weight"], pd["a.linear.bias"], pd["a.attention_query_vector"],
                                    OperandCache(), "t")
    out.backward(g.to(DEV))
    torch.cuda.synchronize()
    return {"fwd_rel": relerr(out, ref), "dx_rel": relerr(xd.grad, x.grad),
            "dW_rel": relerr(pd["a.linear.weight"].grad, p["a.linear.weight"].grad),
            "db_rel": relerr(pd["a.linear.bias"].grad, p["a.linear.bias"].grad),
            "dq_rel": relerr(pd["a.attention_query_vector"].grad, p["a.attention_query_vector"].grad)}


def check_dot_score(B=9, Cn=5, D=300):
    from newsrec_b200.ops import DotScoreFn
    c = O.det_uniform((B, Cn, D), 41).requires_grad_(True)
    u = O.det_uniform((B, D), 42).requires_grad_(True)
    ref = O.dot_product_click_predictor(c, u)
    g = O.det_uniform((B, Cn), 43)
    ref.backward(g)
    cd, ud = c.detach().to(DEV).requires_grad_(True), u.detach().to(DEV).requires_grad_(True)
    out = DotScoreFn.apply(cd, ud)
    out.backward(g.to(DEV))
    return {"fwd_rel": relerr(out, ref), "dc_rel": relerr(cd.grad, c.grad), "du_rel": relerr(ud.grad, u.grad)}


# ------------------------------------------------------------------------------------------------
def check_gru(B=37, S=50, D=900, Hd=900, seed=3):
    """LSTUR user-encoder GRU (pack_padded_sequence + nn.GRU, last hidden state; LSTUR/user_encoder.py:27-45) at the
    reference's history length, mixed lengths including 0 (clamped to 1) and S: forward and every gradient against the
    oracle under the bf16 operand contract.  B <= 128 * floor(SMs / 29) runs the persistent single-launch recurrence."""
    from newsrec_b200.ops import OperandCache
    from newsrec_b200.ops_gru import GruLastHiddenFn
    a = math.sqrt(1.0 / Hd)
    p = {"g.weight_ih_l0": O.det_uniform((3 * Hd, D), seed, -a, a), "g.weight_hh_l0": O.det_uniform((3 * Hd, Hd), seed + 1, -a, a),
         "g.bias_ih_l0": O.det_uniform((3 * Hd,), seed + 2, -a, a), "g.bias_hh_l0": O.det_uniform((3 * Hd,), seed + 3, -a, a)}
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    x = _rand_bf16((B, S, D), seed + 4, 0.5).requires_grad_(True)
    h0 = O.det_uniform((B, Hd), seed + 5, -0.5, 0.5).requires_grad_(True)
    lengths = O.det_randint((B,), seed + 6, 0, S + 1)
    lengths[0], lengths[1 % B] = 0, S
    ref = O.gru_last_hidden(x, lengths.clamp(min=1), h0, p, "g", O.BF16)
    gout = O.det_uniform((B, Hd), seed + 7)
    ref.backward(gout)
    xd, hd = x.detach().to(DEV).requires_grad_(True), h0.detach().to(DEV).requires_grad_(True)
    pd = {k: v.detach().to(DEV).requires_grad_(True) for k, v in p.items()}
    out = GruLastHiddenFn.apply(xd, lengths.to(DEV), hd, pd["g.weight_ih_l0"], pd["g.weight_hh_l0"], pd["g.bias_ih_l0"],
                                pd["g.bias_hh_l0"], OperandCache(), "gru")
    out.backward(gout.to(DEV))
    torch.cuda.synchronize()
    res = {"fwd_rel": relerr(out, ref), "dx_rel": relerr(xd.grad, x.grad), "dh0_rel": relerr(hd.grad, h0.grad),
           "persistent": bool(load_library().nr_gru_persistent_supported(B, Hd))}
    for k in p:
        res["d" + k.split(".")[1]] = relerr(pd[k].grad, p[k].grad)
    return res


# ------------------------------------------------------------------------------------------------
def nrms_model_and_params(V, seed, heads=15, dropout=0.2, fused=False):
    import config as cfgmod
    from model.NRMS import NRMS
    # fused: False = fast mode, True = the one-kernel front end, "accurate" = hi/lo pairs on the unfused kernels
    cfg = type("Cfg", (cfgmod.NRMSConfig,), dict(num_words=V, num_attention_heads=heads, dropout_probability=dropout,
                                                 fused_news_encoder=fused is True, precision="accurate" if fused == "accurate" else "fast"))
    sd = O.det_state_dict(O.nrms_shapes(V), seed)
    model = NRMS(cfg)
    model.load_state_dict(sd)
    return model.to(DEV), sd


def slots(t):
    return [{"title": t[:, j].contiguous()} for j in range(t.shape[1])]


def check_nrms_golden():
    """The committed golden case (minted from the live reference): CUDA path vs reference fp32 outputs and
    vs the oracle under the bf16 storage contract, forward and all parameter gradients."""
    from golden_util import case_params, load_case, oracle_forward, unique_params
    g = load_case("nrms")
    cand_t, clicked_t = torch.from_numpy(g["cand_title"]), torch.from_numpy(g["clicked_title"])
    p = case_params("nrms", g)
    logits_o, _ = oracle_forward("nrms", g, p, O.BF16)
    O.click_loss(logits_o).backward()
    model, _ = nrms_model_and_params(120, int(g["seed"]))
    model.eval()
    logits = model(slots(cand_t), slots(clicked_t))
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(logits.shape[0], dtype=torch.long, device=DEV))
    loss.backward()
    torch.cuda.synchronize()
    out = {"logits_vs_oracle_bf16": relerr(logits, logits_o), "logits_vs_reference_fp32": relerr(logits, torch.from_numpy(g["logits"])),
           "loss_abs_vs_reference": abs(loss.item() - float(g["loss"]))}
    grads = dict(model.named_parameters())
    worst, worst_key = 0.0, ""
    for k, prm in unique_params(p).items():
        if prm.grad.norm() < 1e-5:  # analytically ~0 gradients (W_K.bias) carry only rounding noise
            continue
        e = relerr(grads[k].grad, prm.grad)
        if e > worst:
            worst, worst_key = e, k
        out["grad:" + k] = e
    out["worst_grad_rel"] = worst
    out["worst_grad_key"] = worst_key
    out["emb_row0_grad_zero"] = bool((grads["news_encoder.word_embedding.weight"].grad[0] == 0).all())
    model.check_ids()
    return out


def check_nrms_random(B=8, Cn=5, H=50, T=20, V=500, seed=5, fused=False):
    """A MIND-shaped batch (K=4, history 50, left padded) vs the oracle under the bf16 contract."""
    cand_t, clicked_t, _ = O.synth_batch(B, Cn, H, T, V, seed * 100)
    model, sd = nrms_model_and_params(V, seed, fused=fused)
    model.eval()
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits_o = O.nrms_forward(cand_t, clicked_t, p, 15, O.WEIGHTS_BF16 if fused else O.BF16, c_news=O.BF16_FUSED if fused else O.BF16)
    O.click_loss(logits_o).backward()
    with torch.no_grad():
        logits_x = O.nrms_forward(cand_t, clicked_t, {k: v.detach() for k, v in p.items()}, 15, O.EXACT)
    logits = model(slots(cand_t), slots(clicked_t))
    torch.nn.functional.cross_entropy(logits, torch.zeros(B, dtype=torch.long, device=DEV)).backward()
    torch.cuda.synchronize()
    out = {"logits_vs_oracle_bf16": relerr(logits, logits_o), "logits_vs_exact_fp32": relerr(logits, logits_x),
           "oracle_bf16_vs_exact": relerr(logits_o, logits_x)}
    grads = dict(model.named_parameters())
    worst = 0.0
    for k, prm in p.items():
        if prm.grad.norm() < 1e-5:
            continue
        e = relerr(grads[k].grad, prm.grad)
        out["grad:" + k] = e
        worst = max(worst, e)
    out["worst_grad_rel"] = worst
    return out


def check_nrms_direct_grad_accumulation(B=8, Cn=5, H=50, T=20, V=500, seed=6):
    """Parameters with pre-allocated .grad storage (ddp.FlatGradients) are accumulated in place by the kernels; the
    result must equal the allocate-and-return path, also when two backward passes accumulate."""
    from newsrec_b200 import ddp
    cand_t, clicked_t, _ = O.synth_batch(B, Cn, H, T, V, seed * 100)
    label = torch.zeros(B, dtype=torch.long, device=DEV)
    model_a, _ = nrms_model_and_params(V, seed)
    model_b, _ = nrms_model_and_params(V, seed)
    model_a.eval()
    model_b.eval()
    flat = ddp.FlatGradients(model_b.parameters(), 1)
    flat.zero()
    for _ in range(2):  # two accumulating backward passes
        torch.nn.functional.cross_entropy(model_a(slots(cand_t), slots(clicked_t)), label).backward()
        torch.nn.functional.cross_entropy(model_b(slots(cand_t), slots(clicked_t)), label).backward()
    torch.cuda.synchronize()
    def worst_rel(ma, mb):
        # W_K.bias has an analytically zero gradient (softmax shift invariance): what both paths hold there is fp32
        # accumulation noise in atomic order.  Every difference is therefore measured against the tensor's own scale
        # floored at 1e-3 of the largest gradient in the model.
        floor = 1e-3 * max(float(p.grad.abs().max()) for p in ma.parameters())
        w = 0.0
        for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            scale = max(float(pa.grad.abs().max()), floor)
            w = max(w, float((pa.grad - pb.grad).abs().max()) / scale)
        return w

    worst = worst_rel(model_a, model_b)
    # a third backward with the persistent workspaces must start from cleared accumulators
    flat.zero()
    model_a.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(model_a(slots(cand_t), slots(clicked_t)), label).backward()
    torch.nn.functional.cross_entropy(model_b(slots(cand_t), slots(clicked_t)), label).backward()
    torch.cuda.synchronize()
    again = worst_rel(model_a, model_b)
    return {"direct_vs_returned_rel_maxabs": worst, "after_zero_rel_maxabs": again,
            "grads_are_flat_views": all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in model_b.parameters())}


def check_nrms_prefetch_equals_direct(B=8, Cn=5, H=50, T=20, V=500, seed=7):
    """NRMS.prefetch (copy stream) + forward(PackedBatch) gives bit-identical logits to forward(lists)."""
    cand_t, clicked_t, _ = O.synth_batch(B, Cn, H, T, V, seed * 100)
    model, _ = nrms_model_and_params(V, seed)
    model.eval()
    with torch.no_grad():
        a = model(slots(cand_t), slots(clicked_t))
        pb = model.prefetch(slots(cand_t), slots(clicked_t))
        b = model(pb)
        pb2 = model.prefetch(slots(cand_t), slots(clicked_t))  # a second staged batch while the first is alive
        c = model(pb2)
    torch.cuda.synchronize()
    return {"maxabs": float((a - b).abs().max()), "maxabs_second": float((a - c).abs().max())}


def check_nrms_eval_api(V=300, seed=9):
    """get_news_vector / get_user_vector (non-contiguous input, evaluate.py:220-224) / get_prediction."""
    model, sd = nrms_model_and_params(V, seed)
    model.eval()
    p = {k: v for k, v in sd.items()}
    titles = O.synth_titles(40, 20, V, 3)
    with torch.no_grad():
        nv = model.get_news_vector({"title": titles, "id": ["N%d" % i for i in range(40)]})
        nv_o = O.nrms_news_encoder(titles, p, 15, O.BF16)
        B, H = 4, 10
        stacked = torch.stack([nv[i * 4:(i + 1) * 4] for i in range(H)], dim=0).transpose(0, 1)  # (B,H,d) non-contiguous
        uv = model.get_user_vector(stacked)
        uv_o = O.nrms_user_encoder(stacked.cpu(), p, 15, O.BF16)
        pred = model.get_prediction(nv[:7], uv[0])
        pred_o = (nv[:7].cpu() @ uv[0].cpu())
    return {"news_vec_rel": relerr(nv, nv_o), "user_vec_rel": relerr(uv, uv_o), "pred_rel": relerr(pred, pred_o),
            "user_input_noncontig": not stacked.is_contiguous(), "pred_tolist_len": len(pred.tolist())}


def check_nrms_train_mode(B=16, V=400, seed=4):
    """Training mode: dropout masks come from the in-kernel counter RNG, so parity is statistical:
    the mean logits over many seeds approach the eval logits, gradients are finite, row 0 grad is zero,
    and the backward regenerates exactly the forward masks (grad check against finite differences of the
    SAME masked function is implied by the eval-mode parity of the identical kernels with p=0)."""
    cand_t, clicked_t, _ = O.synth_batch(B, 5, 50, 20, V, seed * 100)
    model, _ = nrms_model_and_params(V, seed)
    model.eval()
    with torch.no_grad():
        ref = model(slots(cand_t), slots(clicked_t))
    model.train()
    acc = torch.zeros_like(ref)
    n = 24
    for _ in range(n):
        with torch.no_grad():
            acc += model(slots(cand_t), slots(clicked_t))
    mean = acc / n
    logits = model(slots(cand_t), slots(clicked_t))
    torch.nn.functional.cross_entropy(logits, torch.zeros(B, dtype=torch.long, device=DEV)).backward()
    g = model.news_encoder.word_embedding.weight.grad
    return {"mean_train_vs_eval_rel": relerr(mean, ref), "train_differs_from_eval": relerr(logits, ref) > 1e-3,
            "grads_finite": bool(torch.isfinite(g).all()), "emb_row0_grad_zero": bool((g[0] == 0).all())}


def check_nrms_train_masked(B=6, Cn=5, H=50, T=20, V=500, seed=8, p_drop=0.2, fused=False):
    """TRAIN mode -- the configuration bench.py times -- forward AND backward against the oracle under the SAME dropout
    masks: the kernels draw their masks from a counter hash of (seed, row, column); the test reads the seed the next
    forward will use (ops.peek_seeds) and hands it to the oracle, which rebuilds the masks with the NumPy restatement of
    that hash (oracle.dropout_mask) at both dropout sites (after the embedding, news_encoder.py:38, and after the
    self-attention, :43).  A backward that regenerated a different mask than its forward would fail every gradient."""
    from newsrec_b200 import ops
    cand_t, clicked_t, _ = O.synth_batch(B, Cn, H, T, V, seed * 100)
    model, sd = nrms_model_and_params(V, seed, dropout=p_drop, fused=fused)
    model.train()
    kseed = ops.peek_seeds(1)[0]  # the news encoder draws the only seed of a forward pass (the user encoder has no dropout)
    drop = dict(p=p_drop, seed=kseed, ld=ru8(300 + 1))
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits_o = O.nrms_forward(cand_t, clicked_t, p, 15, O.WEIGHTS_BF16 if fused else O.BF16, c_news=O.BF16_FUSED if fused else O.BF16,
                              drop=drop)
    O.click_loss(logits_o).backward()
    px = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits_x = O.nrms_forward(cand_t, clicked_t, px, 15, O.EXACT, drop=drop)  # exact arithmetic, same masks
    O.click_loss(logits_x).backward()
    with torch.no_grad():
        logits_eval = O.nrms_forward(cand_t, clicked_t, {k: v.detach() for k, v in px.items()}, 15, O.EXACT)
    logits = model(slots(cand_t), slots(clicked_t))
    torch.nn.functional.cross_entropy(logits, torch.zeros(B, dtype=torch.long, device=DEV)).backward()
    torch.cuda.synchronize()
    out = {"logits_vs_masked_oracle": relerr(logits, logits_o), "logits_vs_masked_exact_fp32": relerr(logits, logits_x),
           "masked_oracle_vs_masked_exact": relerr(logits_o, logits_x), "masks_matter": relerr(logits_x, logits_eval)}
    grads = dict(model.named_parameters())
    gscale = max(float(v.grad.norm()) for v in px.values())
    worst_ratio, worst_key = 0.0, ""
    for k, prm in px.items():
        if prm.grad.norm() < 1e-4 * gscale:
            continue
        e_kernel = relerr(grads[k].grad, prm.grad)
        e_contract = relerr(p[k].grad, prm.grad)
        out["grad:" + k] = [e_kernel, e_contract]
        ratio = e_kernel / max(e_contract, 2e-3)
        if ratio > worst_ratio:
            worst_ratio, worst_key = ratio, k
    out["worst_grad_ratio_kernel_over_contract"] = worst_ratio
    out["worst_grad_key"] = worst_key
    out["emb_row0_grad_zero"] = bool((grads["news_encoder.word_embedding.weight"].grad[0] == 0).all())
    return out


def check_nrms_full_size_properties(B=512):
    """BASELINE.json config[1] sizes (B=512, K=4, H=50, T=20, d=300, 15 heads, V=70976): size-independent
    properties -- (1) permuting the impressions permutes the logits, (2) the logits of a sub-batch equal the
    corresponding rows of the full batch, (3) sum of the embedding gradient over rows == a probe identity:
    d(sum logits)/d(emb) summed over the vocabulary equals the gradient w.r.t. a shared additive shift."""
    V = 70976
    cand_t, clicked_t, _ = O.synth_batch(B, 5, 50, 20, V, 4242)
    model, _ = nrms_model_and_params(V, 3)
    model.eval()
    with torch.no_grad():
        full = model(slots(cand_t), slots(clicked_t))
        perm = torch.from_numpy(__import__("numpy").random.RandomState(0).permutation(B))
        permd = model(slots(cand_t[perm]), slots(clicked_t[perm]))
        sub = model(slots(cand_t[:16]), slots(clicked_t[:16]))
    return {"perm_equivariance_maxabs": maxabs(permd, full[perm.to(DEV)]), "subbatch_maxabs": maxabs(sub, full[:16]),
            "finite": bool(torch.isfinite(full).all())}


# ------------------------------------------------------------------------------------------------
def check_backend_agreement(M=8800, N=900, K=300):
    """tcgen05 accumulators vs the SIMT triage backend on identical operands (fp32 output): locates
    pipeline bugs (which tile / row / column disagrees) that bf16 output rounding would hide."""
    lib = load_library()
    lda = ldw = ru8(K + 1)
    A = torch.zeros(M, lda)
    A[:, :K] = _rand_bf16((M, K), 1)
    W = torch.zeros(N, ldw)
    W[:, :K] = _rand_bf16((N, K), 2, 0.1)
    bias = O.det_uniform((N,), 3, -0.5, 0.5).to(DEV)
    ref = (A[:, :K].double() @ W[:, :K].double().t() + bias.cpu().double())
    Ad, Wd = A.to(torch.bfloat16).to(DEV), W.to(torch.bfloat16).to(DEV)
    ld_out = (N + 3) // 4 * 4
    outs = []
    for simt in (0, 1, 0):
        lib.nr_debug_set_simt_gemm(simt)
        out = torch.full((M, ld_out), float("nan"), dtype=torch.float32, device=DEV)
        check(lib.nr_linear(_p(Ad), M, lda, _p(Wd), N, ldw, K, 1, 0, 128, _p(bias), 0, _p(out), ld_out, 0, _stream()), "nr_linear")
        torch.cuda.synchronize()
        outs.append(out[:, :N].cpu().double())
    lib.nr_debug_set_simt_gemm(0)
    t, s, t2 = outs
    diff = (t - s).abs()
    bad = (diff > 1e-4).nonzero()
    res = {"tc_vs_ref_rel": relerr(t, ref), "simt_vs_ref_rel": relerr(s, ref), "tc_vs_simt_maxabs": float(diff.max()),
           "tc_rerun_maxabs": float((t - t2).abs().max()), "n_bad": int(bad.shape[0])}
    if bad.shape[0]:
        rows, cols = bad[:, 0], bad[:, 1]
        res["bad_rows_sample"] = rows[:12].tolist()
        res["bad_cols_sample"] = cols[:12].tolist()
        res["bad_tiles"] = sorted(set((rows // 128).tolist()))[:40]
        res["bad_col_range"] = [int(cols.min()), int(cols.max())]
        res["bad_row_in_tile_range"] = [int((rows % 128).min()), int((rows % 128).max())]
    return res


def _encoder_fwd_raw(ids, dense, sd, prefix, heads, V, fused=False, p_drop=0.0, seed=0):
    """Direct nr_mhsa_encoder_fwd call returning every intermediate buffer (for differential triage).
    fused=True passes the head-packed operands + the lo plane, i.e. selects the one-kernel front end."""
    from newsrec_b200 import MhsaEncoderFwdArgs
    from newsrec_b200.ops import pack_head_blocks, qkv_pitches, stack_qkv
    lib = load_library()
    d, q = 300, 200
    ldx, ld3 = ru8(d + 1), qkv_pitches(d)[1]
    g = lambda k: sd[f"{prefix}.{k}"].to(DEV)
    wqkv = stack_qkv(*[g(f"multihead_self_attention.W_{n}.weight") for n in "QKV"])
    ops = dict(wqkv=cast_pad(wqkv, ldx), bqkv=stack_qkv(*[g(f"multihead_self_attention.W_{n}.bias") for n in "QKV"]).contiguous(),
               wa=cast_pad(g("additive_attention.linear.weight"), ldx), ba=g("additive_attention.linear.bias").contiguous(),
               qv=g("additive_attention.attention_query_vector").contiguous())
    a = MhsaEncoderFwdArgs()
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    if ids is not None:
        n_seq, T = ids.shape
        table = cast_pad(sd["news_encoder.word_embedding.weight"].to(DEV), ldx)
        a.ids, a.table_bf16, a.V = _p(ids), _p(table), V
    else:
        n_seq, T, _ = dense.shape
        a.dense = _p(dense)
        a.dense_s_seq, a.dense_s_tok, a.dense_s_col = dense.stride()
    n_tok = n_seq * T
    bufs = dict(X=torch.zeros((n_tok, ldx), dtype=torch.bfloat16, device=DEV),
                QKV=None if fused else torch.zeros((n_tok, ld3), dtype=torch.bfloat16, device=DEV),
                C=torch.zeros((n_tok, ldx), dtype=torch.bfloat16, device=DEV), w=torch.zeros((n_tok,), device=DEV),
                out=torch.zeros((n_seq, d), device=DEV))
    a.n_seq, a.T, a.d, a.heads, a.q, a.ldx, a.ld3 = n_seq, T, d, heads, q, ldx, ld3
    a.wqkv_bf16, a.bqkv, a.wa_bf16, a.ba, a.qv = _p(ops["wqkv"]), _p(ops["bqkv"]), _p(ops["wa"]), _p(ops["ba"]), _p(ops["qv"])
    a.p_drop, a.seed = float(p_drop), int(seed)
    a.X_bf16, a.QKV_bf16, a.C_bf16, a.w, a.out = _p(bufs["X"]), _p(bufs["QKV"]), _p(bufs["C"]), _p(bufs["w"]), _p(bufs["out"])
    a.bad_id_flag = _p(flag)
    if fused:
        hw, hb = pack_head_blocks(*[g(f"multihead_self_attention.W_{n}.{k}") for n in "QKV" for k in ("weight", "bias")], heads, ldx)
        bufs["C_lo"] = torch.zeros((n_tok, ldx), dtype=torch.bfloat16, device=DEV)
        a.wqkv_heads_bf16, a.bqkv_heads, a.C_lo_bf16 = _p(hw), _p(hb), _p(bufs["C_lo"])
    check(lib.nr_mhsa_encoder_fwd(C.byref(a), _stream()), "nr_mhsa_encoder_fwd")
    torch.cuda.synchronize()
    out = {k: v.float().cpu() for k, v in bufs.items() if v is not None}
    out["bad_flag"] = int(flag.item())
    return out


def check_fused_front(n_seq=13, V=97, p_drop=0.0, seed=0x1234567, heads=15):
    """The fused front end (csrc/fused_fwd.cu) against (1) the unfused kernel sequence on identical inputs: the gathered
    rows X bit for bit (same table rows, same dropout masks), Q|K|V to bf16 rounding flips; (2) the oracle: the context
    hi + lo planes against an fp64 attention over the kernel's own X with the fused storage contract (Q, K, P bf16; V and the
    context hi/lo pairs), and the pooled news vector."""
    T, d = 20, 300
    sd = O.det_state_dict(O.nrms_shapes(V), 17)
    ids = O.synth_titles(n_seq, T, V, 91).to(DEV)
    un = _encoder_fwd_raw(ids, None, sd, "news_encoder", heads, V, fused=False, p_drop=p_drop, seed=seed)
    fu = _encoder_fwd_raw(ids, None, sd, "news_encoder", heads, V, fused=True, p_drop=p_drop, seed=seed)
    res = {"x_bit_exact": bool(torch.equal(un["X"], fu["X"])), "bad_flag": fu["bad_flag"]}
    # oracle context from the kernel's own gathered rows (dropout already applied there); context mask injected
    p = {k: v.double() for k, v in sd.items()}
    x = fu["X"][:, :d].double().view(n_seq, T, d)
    cmask = None
    if p_drop > 0:
        cmask = O.dropout_mask(seed ^ 0x5bd1e995, p_drop, n_seq * T, d, ru8(d + 1)).double().view(n_seq, T, d)
        xmask = O.dropout_mask(seed, p_drop, n_seq * T, d, ru8(d + 1))
        tab = bf16r(sd["news_encoder.word_embedding.weight"])[ids.cpu().reshape(-1)]
        res["x_vs_masked_oracle_exact"] = bool(torch.equal(fu["X"][:, :d], bf16r(tab * xmask)))
    with torch.no_grad():
        ctx = O.multihead_self_attention(x, p, "news_encoder.multihead_self_attention", heads, O.BF16_FUSED, cmask).view(n_seq * T, d)
        ctx_exact = O.multihead_self_attention(x, p, "news_encoder.multihead_self_attention", heads, O.EXACT, cmask).view(n_seq * T, d)
        pooled = O.additive_attention(ctx.view(n_seq, T, d), p, "news_encoder.additive_attention", O.BF16_FUSED)
    c = fu["C"][:, :d].double() + fu["C_lo"][:, :d].double()
    res["ctx_vs_oracle_fused_contract"] = relerr(c, ctx)
    res["ctx_vs_fp64_attention"] = relerr(c, ctx_exact)
    res["ctx_hi_ones_col"] = bool((fu["C"][:, d] == 1).all() and (fu["C"][:, d + 1:] == 0).all())
    res["out_vs_oracle"] = relerr(fu["out"], pooled)
    res["out_fused_vs_unfused"] = relerr(fu["out"], un["out"])
    res["w_sums_to_one"] = float((fu["w"].view(n_seq, T).sum(1) - 1).abs().max())
    return res


def check_encoder_backend_diff(B=8, V=500, seed=5):
    """Every intermediate of the news and user encoders, tcgen05 vs SIMT triage backend, same inputs."""
    lib = load_library()
    cand_t, clicked_t, _ = O.synth_batch(B, 5, 50, 20, V, seed * 100)
    sd = O.det_state_dict(O.nrms_shapes(V), seed)
    ids = torch.cat((clicked_t.reshape(-1, 20), cand_t.reshape(-1, 20)), 0).to(DEV)
    res = {}
    runs = {}
    for name, simt in (("tc", 0), ("simt", 1)):
        lib.nr_debug_set_simt_gemm(simt)
        news = _encoder_fwd_raw(ids, None, sd, "news_encoder", 15, V)
        dense = news["out"][:B * 50].view(B, 50, 300).to(DEV).contiguous()
        user = _encoder_fwd_raw(None, dense, sd, "user_encoder", 15, V)
        runs[name] = (news, user)
    lib.nr_debug_set_simt_gemm(0)
    for lvl, i in (("news", 0), ("user", 1)):
        for k in ("X", "QKV", "C", "w", "out"):
            a, b = runs["tc"][i][k], runs["simt"][i][k]
            d = (a - b).abs()
            res[f"{lvl}.{k}.maxabs"] = float(d.max())
            res[f"{lvl}.{k}.n_diff"] = int((d > 0).sum())
            res[f"{lvl}.{k}.rel"] = relerr(a, b)
            if k in ("w", "out") and d.max() > 0:
                idx = d.reshape(d.shape[0], -1).max(dim=1).values.topk(min(5, d.shape[0]))
                res[f"{lvl}.{k}.worst_rows"] = idx.indices.tolist()
                res[f"{lvl}.{k}.worst_vals"] = [float(v) for v in idx.values]
    # and against the oracle
    p = {k: v for k, v in sd.items()}
    with torch.no_grad():
        nv_o = O.nrms_news_encoder(ids.cpu(), p, 15, O.BF16)
    res["news.out.tc_vs_oracle"] = relerr(runs["tc"][0]["out"], nv_o)
    res["news.out.simt_vs_oracle"] = relerr(runs["simt"][0]["out"], nv_o)
    dn = (runs["tc"][0]["out"] - nv_o).abs().max(dim=1).values
    res["news.out.tc_vs_oracle_worst_rows"] = dn.topk(5).indices.tolist()
    res["news.out.tc_vs_oracle_worst_vals"] = [float(v) for v in dn.topk(5).values]
    return res


# ------------------------------------------------------------------------------------------------
def build_model(case, V=120, ncat=15, nusers=40, H=6, dropout=0.2, fused=False):
    """Drop-in model + deterministic state_dict for a golden case name (see oracle/make_golden.py)."""
    import importlib
    import config as cfgmod
    from golden_util import case_shapes
    name = {"nrms": "NRMS", "naml": "NAML", "naml_f400": "NAML", "tanr": "TANR", "lstur_ini": "LSTUR", "lstur_con": "LSTUR"}[case]
    over = dict(num_words=V, num_categories=ncat, num_users=nusers, num_clicked_news_a_user=H, dropout_probability=dropout)
    if name == "NRMS":
        over["fused_news_encoder"] = fused is True
        over["precision"] = "accurate" if fused == "accurate" else "fast"
    if name == "LSTUR":
        over["precision"] = "accurate" if fused else "fast"
    if case == "naml_f400":
        over["num_filters"] = 400
    if case.startswith("lstur"):
        over["long_short_term_method"] = case.split("_")[1]
    cfg = type("Cfg", (getattr(cfgmod, name + "Config"),), over)
    Model = getattr(importlib.import_module("model." + name), name)
    return Model(cfg).to(DEV), cfg


def golden_inputs(case, g):
    """Reference-style slot lists for a golden case."""
    t = lambda k: torch.from_numpy(g[k])
    keys = {"title": "title", "abstract": "abstract", "category": "category", "subcategory": "subcategory"}
    def mk(prefix):
        n = g[prefix + "_title"].shape[1]
        out = []
        for j in range(n):
            dct = {}
            for k in keys:
                if f"{prefix}_{k}" in g:
                    dct[k] = t(f"{prefix}_{k}")[:, j].contiguous()
            out.append(dct)
        return out
    return mk("cand"), mk("clicked")


def default_nrms_mode(name="NRMS"):
    """False ("fast") or "accurate": what the NRMS / LSTUR drop-in does when the config says nothing (config.py / NEWSREC_PRECISION)."""
    import config as cfgmod
    return "accurate" if getattr(getattr(cfgmod, name + "Config"), "precision", "fast") == "accurate" else False


def check_golden(case, fused=None):
    """A committed golden case (minted from the live reference): CUDA drop-in vs the reference's fp32 outputs, vs the
    oracle under the bf16 storage contract, and -- per gradient -- against the exact fp32 oracle next to the error the
    bf16 contract itself has (kernel_err <= ~1.5 x contract_err is the pass criterion)."""
    from golden_util import case_params, load_case, oracle_forward, unique_params
    if fused is None:  # the shipped default
        fused = default_nrms_mode() if case == "nrms" else (default_nrms_mode("LSTUR") if case.startswith("lstur") else False)
    g = load_case(case)
    p_b = case_params(case, g)
    logits_b, topic_b = oracle_forward(case, g, p_b, O.BF16, bool(fused))
    (O.click_loss(logits_b) + (0.1 * topic_b if topic_b is not None else 0.0)).backward()
    p_x = case_params(case, g)
    logits_x, topic_x = oracle_forward(case, g, p_x, O.EXACT)
    (O.click_loss(logits_x) + (0.1 * topic_x if topic_x is not None else 0.0)).backward()
    model, _ = build_model(case, fused=fused)
    sd = O.tie_shared(O.det_state_dict(__import__("golden_util").case_shapes(case), int(g["seed"])))
    model.load_state_dict(sd)
    model.eval()
    cand, clicked = golden_inputs(case, g)
    if case.startswith("lstur"):
        out = model(torch.from_numpy(g["user"]), torch.from_numpy(g["clicked_news_length"]).clone(), cand, clicked)
    else:
        out = model(cand, clicked)
    logits, topic = (out if isinstance(out, tuple) else (out, None))
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(logits.shape[0], dtype=torch.long, device=DEV))
    (loss + (0.1 * topic if topic is not None else 0.0)).backward()
    torch.cuda.synchronize()
    with torch.no_grad():  # the blueprint's tolerance definition: fp32 oracle on bf16-rounded weights / embeddings
        logits_w, _ = oracle_forward(case, g, case_params(case, g, requires_grad=False), O.WEIGHTS_BF16)
    res = {"logits_vs_oracle_bf16": relerr(logits, logits_b), "logits_vs_reference_fp32": relerr(logits, torch.from_numpy(g["logits"])),
           "logits_vs_weights_only_oracle": relerr(logits, logits_w),
           "oracle_bf16_vs_reference_fp32": relerr(logits_b, torch.from_numpy(g["logits"])),
           "loss_abs_vs_reference": abs(loss.item() - float(g["loss"]))}
    if topic is not None:
        res["topic_loss_rel_vs_reference"] = abs(topic.item() - float(g["topic_loss"])) / abs(float(g["topic_loss"]))
    grads = dict(model.named_parameters())
    worst_ratio, worst_key, worst_vs_b = 0.0, "", 0.0
    gscale = max(float(v.grad.norm()) for v in unique_params(p_x).values())
    for k, prm in unique_params(p_x).items():
        gk = grads[k].grad
        if gk is None:
            res["missing_grad:" + k] = True
            continue
        if prm.grad.norm() < 1e-4 * gscale:  # analytically ~0 gradients (W_K.bias): rounding noise only
            continue
        e_kernel = relerr(gk, prm.grad)
        e_contract = relerr(unique_params(p_b)[k].grad, prm.grad)
        e_vs_b = relerr(gk, unique_params(p_b)[k].grad)
        res["grad:" + k] = [e_kernel, e_contract, e_vs_b]
        ratio = e_kernel / max(e_contract, 2e-3)
        if ratio > worst_ratio:
            worst_ratio, worst_key = ratio, k
        worst_vs_b = max(worst_vs_b, e_vs_b)
    res["worst_grad_ratio_kernel_over_contract"] = worst_ratio
    res["worst_grad_key"] = worst_key
    res["worst_grad_vs_oracle_bf16"] = worst_vs_b
    w = grads.get("news_encoder.word_embedding.weight", grads.get("news_encoder.text_encoders.title.word_embedding.weight"))
    res["emb_row0_grad_zero"] = bool((w.grad[0] == 0).all())
    return res


def check_train_masked(case, p_drop=0.2, mask_p=0.5, fused=None):
    """TRAIN mode of a CNN family (NAML / TANR / LSTUR) on its golden inputs, forward AND backward, against the oracle under
    the SAME dropout masks (see check_nrms_train_masked): one seed per text-encoder call (NAML: title, then abstract), masks
    over the zero-padded gather layout and the compact conv-output layout.  LSTUR's user masking (F.dropout2d on the
    (1, B, dim) user embedding, LSTUR/__init__.py:74-77) is drawn by torch.rand on the device generator: the test seeds it,
    reads the draw the model is going to make, re-seeds and hands the same keep-multipliers to the oracle."""
    from golden_util import case_params, case_shapes, load_case, oracle_forward, unique_params
    from newsrec_b200 import ops
    g = load_case(case)
    if fused is None:  # the shipped default of the family (LSTUR: accurate = conv output / GRU input as hi/lo pairs)
        fused = default_nrms_mode("LSTUR") if case.startswith("lstur") else False
    model, cfg = build_model(case, dropout=p_drop, fused=fused)
    sd = O.tie_shared(O.det_state_dict(case_shapes(case), int(g["seed"])))
    model.load_state_dict(sd)
    model.train()
    cand, clicked = golden_inputs(case, g)
    B = g["cand_title"].shape[0]
    user_keep = None
    if case.startswith("lstur"):
        cfg.masking_probability = mask_p
        torch.cuda.manual_seed(1234)
        user_keep = ((torch.rand(B, 1, device=DEV) >= mask_p).float() / (1.0 - mask_p)).cpu()
        torch.cuda.manual_seed(1234)
    if case.startswith("naml"):
        s1, s2 = ops.peek_seeds(2)
        drop = dict(p=p_drop, seeds={"title": s1, "abstract": s2})
    else:
        drop = dict(p=p_drop, seed=ops.peek_seeds(1)[0])
    tw = lambda t: (0.1 * t if t is not None else 0.0)
    p_b = case_params(case, g)
    logits_b, topic_b = oracle_forward(case, g, p_b, O.BF16, bool(fused), drop=drop, user_keep=user_keep)
    (O.click_loss(logits_b) + tw(topic_b)).backward()
    p_x = case_params(case, g)
    logits_x, topic_x = oracle_forward(case, g, p_x, O.EXACT, drop=drop, user_keep=user_keep)
    (O.click_loss(logits_x) + tw(topic_x)).backward()
    with torch.no_grad():
        logits_eval, _ = oracle_forward(case, g, case_params(case, g, requires_grad=False), O.EXACT)
    if case.startswith("lstur"):
        out = model(torch.from_numpy(g["user"]), torch.from_numpy(g["clicked_news_length"]).clone(), cand, clicked)
    else:
        out = model(cand, clicked)
    logits, topic = (out if isinstance(out, tuple) else (out, None))
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(logits.shape[0], dtype=torch.long, device=DEV))
    (loss + tw(topic)).backward()
    torch.cuda.synchronize()
    res = {"logits_vs_masked_oracle": relerr(logits, logits_b), "logits_vs_masked_exact_fp32": relerr(logits, logits_x),
           "masked_oracle_vs_masked_exact": relerr(logits_b, logits_x), "masks_matter": relerr(logits_x, logits_eval)}
    if topic is not None:
        res["topic_loss_rel_vs_masked_exact"] = abs(topic.item() - float(topic_x)) / abs(float(topic_x))
    grads = dict(model.named_parameters())
    gscale = max(float(v.grad.norm()) for v in unique_params(p_x).values())
    worst_ratio, worst_key = 0.0, ""
    for k, prm in unique_params(p_x).items():
        if grads[k].grad is None:
            res["missing_grad:" + k] = True
            continue
        if prm.grad.norm() < 1e-4 * gscale:
            continue
        e_kernel = relerr(grads[k].grad, prm.grad)
        e_contract = relerr(unique_params(p_b)[k].grad, prm.grad)
        res["grad:" + k] = [e_kernel, e_contract]
        ratio = e_kernel / max(e_contract, 2e-3)
        if ratio > worst_ratio:
            worst_ratio, worst_key = ratio, k
    res["worst_grad_ratio_kernel_over_contract"] = worst_ratio
    res["worst_grad_key"] = worst_key
    return res


def check_predict_impressions(n_news=500, D=300, n_imp=200, seed=3):
    """Batched evaluation scoring (ops.predict_impressions) against the evaluator's per-impression get_prediction loop."""
    from newsrec_b200.ops import predict_impressions
    model, _ = nrms_model_and_params(50, 1)
    news = O.det_uniform((n_news, D), seed).to(DEV)
    users = O.det_uniform((n_imp, D), seed + 1).to(DEV)
    counts = O.det_randint((n_imp,), seed + 2, 1, 40)
    offs = torch.zeros(n_imp + 1, dtype=torch.int64)
    offs[1:] = counts.cumsum(0)
    cand = O.det_randint((int(offs[-1]),), seed + 3, 0, n_news)
    got = predict_impressions(news, cand, offs, users)
    ref = []
    for s in range(n_imp):  # evaluate.py:245-260
        idx = cand[offs[s]:offs[s + 1]].to(DEV)
        ref.append(model.get_prediction(news[idx], users[s]))
    ref = torch.cat(ref)
    return {"rel": relerr(got, ref), "n": int(got.numel())}


def check_pack_slots(B=37, H=50, Cn=5, tail=(20,), where="pinned"):
    """SlotPacker.pack through the one-launch batch feed (nr_pack_slots) against the stack / transpose / cat it replaces."""
    from newsrec_b200.pack import SlotPacker
    gen = torch.Generator().manual_seed(B * 131 + H)
    mk = lambda: torch.randint(0, 70000, (B,) + tuple(tail), generator=gen, dtype=torch.int64)
    place = {"pinned": lambda t: t.pin_memory(), "device": lambda t: t.to(DEV), "pageable": lambda t: t}[where]
    clicked = [{"f": place(mk())} for _ in range(H)]
    cand = [{"f": place(mk())} for _ in range(Cn)]
    pk = SlotPacker()
    direct = pk._pack_direct([x["f"] for x in clicked], [x["f"] for x in cand], DEV)
    ids, Bo = pk.pack(clicked, cand, "f", DEV)
    torch.cuda.synchronize()
    ref = torch.cat((torch.stack([x["f"].cpu() for x in clicked], 1).reshape(B * H, *tail),
                     torch.stack([x["f"].cpu() for x in cand], 1).reshape(B * Cn, *tail)), 0)
    return {"equal": bool(torch.equal(ids.cpu(), ref)), "B": Bo, "direct": direct is not None,
            "direct_equal": direct is None or bool(torch.equal(direct[0].cpu(), ref))}
