"""Per-role cycle counters of the fused front end at the benchmark shape (one launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "news-recommendation_b200", "src")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import newsrec_oracle as O  # noqa: E402
import gpu_checks as G  # noqa: E402
from newsrec_b200 import load_library  # noqa: E402
from newsrec_b200.ops import _p  # noqa: E402

lib = load_library()
n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 28160
p_drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
V = 70976
sd = O.det_state_dict(O.nrms_shapes(V), 17)
ids = O.synth_titles(n_seq, 20, V, 91).to("cuda")
G._encoder_fwd_raw(ids, None, sd, "news_encoder", 15, V, fused=True, p_drop=p_drop, seed=1)  # warm-up
buf = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")
lib.nr_debug_set_fused_timing(_p(buf))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
G._encoder_fwd_raw(ids, None, sd, "news_encoder", 15, V, fused=True, p_drop=p_drop, seed=1)
e1.record()
torch.cuda.synchronize()
lib.nr_debug_set_fused_timing(None)
t = buf.view(148, 32).double().cpu()
names = {0: "g0 wait QKV", 1: "g0 step a", 2: "g0 wait S", 3: "g0 step b", 4: "g0 wait O", 5: "g0 step c", 6: "g0 role",
         8: "g1 wait QKV", 9: "g1 step a", 10: "g1 wait S", 11: "g1 step b", 12: "g1 wait O", 13: "g1 step c", 14: "g1 role",
         16: "gather wait X free", 17: "gather work", 20: "TMA wait W free", 24: "mma wait X", 25: "mma QKV issue (incl W wait)",
         26: "mma W wait", 27: "mma idle polling", 28: "mma role", 29: "mma idle polls"}
tiles = (n_seq + 5) // 6
print(f"n_seq={n_seq} tiles={tiles} ({tiles / 148:.1f}/CTA) p_drop={p_drop}; both launches incl. pool: {e0.elapsed_time(e1):.3f} ms")
for k, nm in names.items():
    col = t[:, k]
    print(f"  [{k:2d}] {nm:32s} mean {col.mean():12.0f}  min {col.min():12.0f}  max {col.max():12.0f}   per tile {col.mean() / (tiles / 148):10.0f}")
