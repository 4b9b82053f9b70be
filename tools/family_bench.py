"""One training step (forward + backward, train mode) of every model family of the drop-in at MIND shapes
(batch 512, 1+K = 5 candidates, history 50, title 20 / abstract 50 tokens), timed with CUDA events on device-resident
inputs.  A coverage measurement next to bench.py (which is the contract benchmark on NRMS).

    python tools/family_bench.py [batch]
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "news-recommendation_b200", "src"))

import torch  # noqa: E402

import config as cfgmod  # noqa: E402
import newsrec_b200  # noqa: E402
from newsrec_b200 import ddp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
C, H, T, TA = 5, 50, 20, 50


def slots(n, seed, cfg, want):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        d = {}
        if "title" in want:
            d["title"] = torch.randint(1, cfg.num_words, (B, T), generator=g).to(dev)
        if "abstract" in want:
            d["abstract"] = torch.randint(1, cfg.num_words, (B, TA), generator=g).to(dev)
        if "category" in want:
            d["category"] = torch.randint(1, cfg.num_categories, (B,), generator=g).to(dev)
            d["subcategory"] = torch.randint(1, cfg.num_categories, (B,), generator=g).to(dev)
        out.append(d)
    return out


CASES = [
    ("NRMS", {}, ("title",)),
    ("NAML", {}, ("title", "abstract", "category")),
    ("TANR", {}, ("title", "category")),
    ("LSTUR", {"long_short_term_method": "ini"}, ("title", "category")),
    ("LSTUR", {"long_short_term_method": "con"}, ("title", "category")),
]
label = torch.zeros(B, dtype=torch.long, device=dev)
res = {}
for name, over, want in CASES:
    cfg = type("Cfg", (getattr(cfgmod, name + "Config"),), over)
    model = getattr(importlib.import_module("model." + name), name)(cfg).to(dev)
    model.train()
    grads = ddp.FlatGradients(model.parameters(), 1)
    cand, clicked = slots(C, 1, cfg, want), slots(H, 2, cfg, want)
    extra = ()
    if name == "LSTUR":
        user = torch.randint(1, cfg.num_users, (B,)).to(dev)
        length = torch.randint(1, H + 1, (B,))
        extra = (user, length)

    def step():
        grads.zero()
        out = model(*extra, cand, clicked)
        if isinstance(out, tuple):  # TANR: (click logits, topic loss)
            loss = torch.nn.functional.cross_entropy(out[0], label) + cfg.topic_classification_loss_weight * out[1]
        else:
            loss = torch.nn.functional.cross_entropy(out, label)
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = newsrec_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    key = name + ("/" + over["long_short_term_method"] if over else "")
    res[key] = {"ms_per_step": round(ms, 3), "impressions_per_s": round(B / ms * 1e3), "launches_per_step": (newsrec_b200.launch_count() - l0) // 5}
    print(key, res[key], flush=True)
    del model, grads
    torch.cuda.empty_cache()
print(json.dumps({"batch": B, "families": res}))
