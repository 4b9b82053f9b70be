"""Where do the gemm_nt CTAs wait?  One NRMS training step at the bench size with the per-role cycle counters of
nr_debug_set_gemm_timing switched on; prints, per GEMM launch, the share of the kernel each role spent waiting.

    python tools/gemm_timing.py [NRMS|NAML|LSTUR|TANR] [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "news-recommendation_b200", "src"))

import torch  # noqa: E402

import importlib  # noqa: E402

import bench  # noqa: E402
import config as cfgmod  # noqa: E402
import newsrec_b200  # noqa: E402

args = [a for a in sys.argv[1:]]
name = args[0] if args and not args[0].isdigit() else "NRMS"
B = int(args[-1]) if args and args[-1].isdigit() else 512
dev = torch.device("cuda", 0)
lib = newsrec_b200.load_library()
Model = getattr(importlib.import_module("model." + name), name)
over = {"long_short_term_method": "ini"} if name == "LSTUR" else {}
model = Model(type("Cfg", (getattr(cfgmod, name + "Config"),), over)).to(dev)
model.train()
extra, cand, clicked = bench.synth_slots(name, B, 7, device=dev)
label = torch.zeros(B, dtype=torch.long, device=dev)


def step():
    model.zero_grad(set_to_none=True)
    out = model(extra[0], extra[1].clone(), cand, clicked) if name == "LSTUR" else model(cand, clicked)
    loss = torch.nn.functional.cross_entropy(out[0], label) + 0.1 * out[1] if isinstance(out, tuple) else torch.nn.functional.cross_entropy(out, label)
    loss.backward()
    torch.cuda.synchronize()


for _ in range(2):
    step()
SLOTS = 32
buf = torch.zeros(SLOTS, 148, 16, dtype=torch.int64, device=dev)
lib.nr_debug_set_gemm_timing(buf.data_ptr(), SLOTS)
step()
lib.nr_debug_set_gemm_timing(None, 0)
t = buf.cpu().double()
names = ["prod:empty", "mma:full", "mma:tempty", "epi:tfull", "epi:body", "kernel", "tiles"]
for s in range(SLOTS):
    used = t[s, :, 5] > 0
    if used.sum() == 0:
        continue
    m = t[s][used].mean(0)
    k = m[5].item()
    print(f"slot {s:2d} ctas={int(used.sum())} kernel={k / 1e3:8.1f} kcyc tiles/cta={m[6].item():6.1f}  " +
          "  ".join(f"{names[i]}={100 * m[i].item() / k:5.1f}%" for i in range(5)) +
          f"  epi/tile={m[4].item() / max(m[6].item(), 1):7.0f} cyc  mma:issue={100 * m[7].item() / k:5.1f}% mma:commit={100 * m[8].item() / k:5.1f}% mma:fence={100 * m[9].item() / k:5.1f}%",
          flush=True)
