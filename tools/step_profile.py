"""Host-side profile of one training step of a drop-in model (cProfile over a few steps, top functions by cumulative time).
A tuning tool: shows where the Python / launch overhead of a step goes when the GPU is not the limiter.

    python tools/step_profile.py [NRMS|NAML|LSTUR|TANR] [batch] [--host]     (--host: pinned CPU slot lists instead of device tensors)
"""
import cProfile
import importlib
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "news-recommendation_b200", "src"))
import torch  # noqa: E402

import bench  # noqa: E402
import config as cfgmod  # noqa: E402
import newsrec_b200  # noqa: E402
from newsrec_b200 import ddp  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args and not args[0].isdigit() else "NRMS"
B = int(args[-1]) if args and args[-1].isdigit() else 512
host = "--host" in sys.argv
dev = torch.device("cuda", 0)
Model = getattr(importlib.import_module("model." + name), name)
over = {"long_short_term_method": "ini"} if name == "LSTUR" else {}
model = Model(type("Cfg", (getattr(cfgmod, name + "Config"),), over)).to(dev)
model.train()
grads = ddp.FlatGradients(model.parameters(), 1)
extra, cand, clicked = bench.synth_slots(name, B, 7, pin=True) if host else bench.synth_slots(name, B, 7, device=dev)
label = torch.zeros(B, dtype=torch.long, device=dev)


def step():
    grads.zero()
    out = model(extra[0], extra[1].clone(), cand, clicked) if name == "LSTUR" else model(cand, clicked)
    loss = torch.nn.functional.cross_entropy(out[0], label) + 0.1 * out[1] if isinstance(out, tuple) else torch.nn.functional.cross_entropy(out, label)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{name} B={B} host={host}: launch side {1e3 * (t1 - t0) / 10:.3f} ms/step, with GPU drain {1e3 * (t2 - t0) / 10:.3f} ms/step, "
      f"library launches/step {newsrec_b200.launch_count() // 15}")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
