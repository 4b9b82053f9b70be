"""Helpers shared by the oracle-vs-golden and kernel-vs-oracle tests."""
import os

import numpy as np
import torch

import newsrec_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V, NCAT, NUSERS = 120, 15, 40  # must match oracle/make_golden.py


def load_case(case):
    z = np.load(os.path.join(GOLDEN, f"{case}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def case_shapes(case):
    if case == "nrms":
        return O.nrms_shapes(V)
    if case == "naml":
        return O.naml_shapes(V, NCAT)
    if case == "naml_f400":
        return O.naml_shapes(V, NCAT, Fn=400)
    if case == "tanr":
        return O.tanr_shapes(V, NCAT)
    return O.lstur_shapes(V, NCAT, NUSERS, method=case.split("_")[1])


def case_params(case, g, dtype=torch.float32, requires_grad=True):
    sd = O.tie_shared(O.det_state_dict(case_shapes(case), int(g["seed"])))
    out, seen = {}, {}
    for k, v in sd.items():
        if id(v) in seen:               # tied storage -> the same leaf tensor under both keys
            out[k] = out[seen[id(v)]]
            continue
        seen[id(v)] = k
        out[k] = v.to(dtype).clone().requires_grad_(requires_grad)
    return out


def t(g, key):
    return torch.from_numpy(g[key])


def oracle_forward(case, g, p, contract=O.EXACT, fused=False, drop=None, user_keep=None):
    """Runs the oracle on a golden case's inputs.  Returns (logits, topic_loss or None).
    fused: the NRMS news level runs through the one-kernel front end (V / context as hi/lo bf16 pairs)."""
    cand_t, clicked_t = t(g, "cand_title"), t(g, "clicked_title")
    if case == "nrms":
        if contract.bf16 and contract.acts and fused:
            # precise mode: fused news front end (V / context / P as hi/lo pairs, Q / K bf16) + fp32-accurate user encoder
            return O.nrms_forward(cand_t, clicked_t, p, 15, O.WEIGHTS_BF16, c_news=O.BF16_FUSED, drop=drop), None
        return O.nrms_forward(cand_t, clicked_t, p, 15, contract, c_news=contract, drop=drop), None
    if case.startswith("naml"):
        cand = dict(title=cand_t, abstract=t(g, "cand_abstract"), category=t(g, "cand_category"),
                    subcategory=t(g, "cand_subcategory"))
        clicked = dict(title=clicked_t, abstract=t(g, "clicked_abstract"), category=t(g, "clicked_category"),
                       subcategory=t(g, "clicked_subcategory"))
        return O.naml_forward(cand, clicked, p, contract, drop=drop), None
    if case == "tanr":
        cand = dict(title=cand_t, category=t(g, "cand_category"))
        clicked = dict(title=clicked_t, category=t(g, "clicked_category"))
        return O.tanr_forward(cand, clicked, p, contract, drop=drop)
    method = case.split("_")[1]
    if contract.bf16 and contract.acts and fused:  # LSTUR's accurate mode: conv output and GRU input as hi/lo bf16 pairs
        contract = O.BF16_FUSED
    cand = dict(title=cand_t, category=t(g, "cand_category"), subcategory=t(g, "cand_subcategory"))
    clicked = dict(title=clicked_t, category=t(g, "clicked_category"), subcategory=t(g, "clicked_subcategory"))
    return O.lstur_forward(t(g, "user"), t(g, "clicked_news_length"), cand, clicked, p, method, contract, drop=drop,
                           user_keep=user_keep), None


def grad_summary(gt: torch.Tensor, key: str):
    flat = gt.detach().reshape(-1).double().cpu()
    n = flat.numel()
    probe = O.det_uniform((n,), O._str_seed("probe:" + key), -1.0, 1.0, torch.float64)
    idx = O.det_randint((256,), O._str_seed("idx:" + key), 0, n)
    return np.array([flat.norm().item(), (flat * probe).sum().item()]), flat[idx].float().numpy()


def unique_params(p):
    seen, out = set(), {}
    for k, v in p.items():
        if id(v) in seen:
            continue
        seen.add(id(v))
        out[k] = v
    return out
