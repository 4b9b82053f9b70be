"""Generate golden vectors from the LIVE reference modules (build container only).

    PYTHONHASHSEED=0 python oracle/make_golden.py

Imports the unmodified reference from /root/reference/src (read-only), loads a
deterministic state_dict (oracle.det_state_dict -- no RNG-stream dependence), runs
forward + CrossEntropy(label 0) + backward in .eval() mode on CPU fp32, and writes
small fixtures to tests/golden/<case>.npz:

    logits, loss, news/user vectors (full), and for every parameter gradient
    its L2 norm, its dot product with a deterministic probe vector and 256
    deterministically sampled elements (full grads would be MBs per case).

The fixtures record torch version and thread count.  /root/reference does not exist
on the GPU box; tests only read the committed .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import newsrec_oracle as O  # noqa: E402

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

V, NCAT, NUSERS = 120, 15, 40
B, C, H, T, TA = 3, 3, 6, 20, 50


def make_config(name, **kw):
    base = dict(num_words=V, num_categories=NCAT, num_users=NUSERS, word_embedding_dim=300,
                category_embedding_dim=100, query_vector_dim=200, dropout_probability=0.2,
                num_clicked_news_a_user=H, num_words_title=T, num_words_abstract=TA,
                num_attention_heads=15, num_filters=300, window_size=3,
                long_short_term_method="ini", masking_probability=0.5,
                dataset_attributes={"news": ["title"], "record": []})
    base.update(kw)
    return type(f"{name}Config", (), base)


def grad_summary(g: torch.Tensor, key: str):
    flat = g.detach().reshape(-1).double()
    n = flat.numel()
    probe = O.det_uniform((n,), O._str_seed("probe:" + key), -1.0, 1.0, torch.float64)
    idx = O.det_randint((256,), O._str_seed("idx:" + key), 0, n)
    return np.array([flat.norm().item(), (flat * probe).sum().item()]), flat[idx].float().numpy()


def slots(t):  # (B,S,...) -> reference slot-major list of (B,...) tensors
    return [t[:, j].contiguous() for j in range(t.shape[1])]


def run_case(case):
    sys.path.insert(0, REF_SRC)
    import importlib
    seed = {"nrms": 11, "naml": 12, "tanr": 13, "lstur_ini": 14, "lstur_con": 15, "naml_f400": 16}[case]
    cand_t, clicked_t, hist_len = O.synth_batch(B, C, H, T, V, seed * 100)
    extra = {}
    if case == "nrms":
        cfg = make_config("NRMS")
        Model = importlib.import_module("model.NRMS").NRMS
        shapes = O.nrms_shapes(V)
        cand = [{"title": x} for x in slots(cand_t)]
        clicked = [{"title": x} for x in slots(clicked_t)]
        args = (cand, clicked)
    elif case in ("naml", "naml_f400"):
        Fn = 400 if case == "naml_f400" else 300
        cfg = make_config("NAML", num_filters=Fn,
                          dataset_attributes={"news": ["category", "subcategory", "title", "abstract"], "record": []})
        Model = importlib.import_module("model.NAML").NAML
        shapes = O.naml_shapes(V, NCAT, Fn=Fn)
        ca, ha, _ = O.synth_batch(B, C, H, TA, V, seed * 100 + 50)
        ca = ca * (cand_t[..., :1] > 0)
        ha = ha * (clicked_t[..., :1] > 0)
        cc = O.det_randint((B, C), seed * 100 + 60, 1, NCAT)
        cs = O.det_randint((B, C), seed * 100 + 61, 1, NCAT)
        hc = O.det_randint((B, H), seed * 100 + 62, 1, NCAT) * (clicked_t[..., 0] > 0)
        hs = O.det_randint((B, H), seed * 100 + 63, 1, NCAT) * (clicked_t[..., 0] > 0)
        extra = dict(cand_abstract=ca, clicked_abstract=ha, cand_category=cc, cand_subcategory=cs,
                     clicked_category=hc, clicked_subcategory=hs)
        cand = [{"title": a, "abstract": b, "category": c_, "subcategory": d_}
                for a, b, c_, d_ in zip(slots(cand_t), slots(ca), slots(cc), slots(cs))]
        clicked = [{"title": a, "abstract": b, "category": c_, "subcategory": d_}
                   for a, b, c_, d_ in zip(slots(clicked_t), slots(ha), slots(hc), slots(hs))]
        args = (cand, clicked)
    elif case == "tanr":
        cfg = make_config("TANR", dataset_attributes={"news": ["category", "title"], "record": []})
        Model = importlib.import_module("model.TANR").TANR
        shapes = O.tanr_shapes(V, NCAT)
        cc = O.det_randint((B, C), seed * 100 + 60, 1, NCAT)
        hc = O.det_randint((B, H), seed * 100 + 62, 1, NCAT) * (clicked_t[..., 0] > 0)
        extra = dict(cand_category=cc, clicked_category=hc)
        cand = [{"title": a, "category": c_} for a, c_ in zip(slots(cand_t), slots(cc))]
        clicked = [{"title": a, "category": c_} for a, c_ in zip(slots(clicked_t), slots(hc))]
        args = (cand, clicked)
    else:
        method = case.split("_")[1]
        cfg = make_config("LSTUR", long_short_term_method=method,
                          dataset_attributes={"news": ["category", "subcategory", "title"],
                                              "record": ["user", "clicked_news_length"]})
        Model = importlib.import_module("model.LSTUR").LSTUR
        shapes = O.lstur_shapes(V, NCAT, NUSERS, method=method)
        cc = O.det_randint((B, C), seed * 100 + 60, 1, NCAT)
        cs = O.det_randint((B, C), seed * 100 + 61, 1, NCAT)
        hc = O.det_randint((B, H), seed * 100 + 62, 1, NCAT) * (clicked_t[..., 0] > 0)
        hs = O.det_randint((B, H), seed * 100 + 63, 1, NCAT) * (clicked_t[..., 0] > 0)
        user = O.det_randint((B,), seed * 100 + 70, 1, NUSERS)
        lengths = hist_len.clone()
        lengths[0] = 0  # exercise the reference's 0 -> 1 clamp (LSTUR/user_encoder.py:27)
        extra = dict(cand_category=cc, cand_subcategory=cs, clicked_category=hc, clicked_subcategory=hs,
                     user=user, clicked_news_length=lengths)
        cand = [{"title": a, "category": c_, "subcategory": d_} for a, c_, d_ in zip(slots(cand_t), slots(cc), slots(cs))]
        clicked = [{"title": a, "category": c_, "subcategory": d_}
                   for a, c_, d_ in zip(slots(clicked_t), slots(hc), slots(hs))]
        args = (user, lengths.clone(), cand, clicked)

    sd = O.tie_shared(O.det_state_dict(shapes, seed))
    model = Model(cfg)
    missing = set(model.state_dict().keys()) ^ set(sd.keys())
    assert not missing, f"state_dict key mismatch for {case}: {sorted(missing)}"
    model.load_state_dict(sd)
    model.eval()

    # capture encoder outputs through forward hooks (order of calls: C candidates then H clicked)
    news_vecs, user_vecs = [], []
    model.news_encoder.register_forward_hook(lambda m, i, o: news_vecs.append(o.detach()))
    model.user_encoder.register_forward_hook(lambda m, i, o: user_vecs.append(o.detach()))
    out = model(*args)
    topic_loss = None
    if isinstance(out, tuple):
        logits, topic_loss = out
    else:
        logits = out
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(B, dtype=torch.long))
    total = loss + (0.1 * topic_loss if topic_loss is not None else 0.0)
    total.backward()

    rec = dict(cand_title=cand_t.numpy(), clicked_title=clicked_t.numpy(), hist_len=hist_len.numpy(),
               logits=logits.detach().numpy(), loss=np.array(loss.item()),
               cand_vec=torch.stack(news_vecs[:C], dim=1).numpy(),
               clicked_vec=torch.stack(news_vecs[C:C + H], dim=1).numpy(),
               user_vec=user_vecs[0].numpy(), seed=np.array(seed),
               meta=np.array(f"torch={torch.__version__} threads={torch.get_num_threads()} ref=8323a4f"))
    if topic_loss is not None:
        rec["topic_loss"] = np.array(topic_loss.item())
    for k, v in extra.items():
        rec[k] = v.numpy()
    seen = set()
    for k, prm in model.named_parameters():
        if prm.grad is None or id(prm) in seen:
            continue
        seen.add(id(prm))
        s, samp = grad_summary(prm.grad, k)
        rec["gsum:" + k] = s
        rec["gsamp:" + k] = samp
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"{case}.npz"), **rec)
    print(f"{case}: loss={loss.item():.6f} logits[0]={logits[0].tolist()} -> {case}.npz "
          f"({os.path.getsize(os.path.join(OUT, case + '.npz')) / 1024:.0f} KB)")


if __name__ == "__main__":
    assert os.path.isdir(REF_SRC), "the reference is only mounted in the build container"
    torch.manual_seed(0)
    for case in (sys.argv[1:] or ["nrms", "naml", "naml_f400", "tanr", "lstur_ini", "lstur_con"]):
        run_case(case)
