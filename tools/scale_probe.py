"""Forward+backward of the NRMS drop-in at growing batch sizes with a hard Python watchdog
(faulthandler dumps the stack and exits if a step takes too long).  NEWSREC_TRACE=1 prints every kernel.

    NEWSREC_TRACE=1 python tools/scale_probe.py 16 64 256 512
"""
import faulthandler
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "news-recommendation_b200", "src"))

import torch  # noqa: E402

import bench  # noqa: E402
import config as cfgmod  # noqa: E402
import newsrec_b200  # noqa: E402
from model.NRMS import NRMS  # noqa: E402

faulthandler.enable()
dev = torch.device("cuda", 0)
model = NRMS(cfgmod.NRMSConfig).to(dev)
model.train()
for B in [int(x) for x in sys.argv[1:]] or [16, 64, 256, 512]:
    faulthandler.dump_traceback_later(int(__import__("os").environ.get("PROBE_WATCHDOG", "40")), exit=True)
    _, cand, clicked = bench.synth_slots("NRMS", B, 7, device=dev)
    label = torch.zeros(B, dtype=torch.long, device=dev)
    for it in range(2):
        t0 = time.time()
        model.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(cand, clicked), label)
        torch.cuda.synchronize()
        t1 = time.time()
        loss.backward()
        torch.cuda.synchronize()
        t2 = time.time()
        print(f"B={B} it={it} loss={loss.item():.4f} fwd={1e3 * (t1 - t0):.1f}ms bwd={1e3 * (t2 - t1):.1f}ms "
              f"launches={newsrec_b200.launch_count()}", flush=True)
    faulthandler.cancel_dump_traceback_later()
print("scale probe done", flush=True)
