"""GPU triage: the fused front end against the unfused sequence and the oracle (prints the metric dicts)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "news-recommendation_b200", "src")):
    sys.path.insert(0, p)
import gpu_checks as G  # noqa: E402

for kw in (dict(n_seq=6), dict(n_seq=13), dict(n_seq=1000, V=5000), dict(n_seq=13, p_drop=0.2)):
    try:
        print(kw, json.dumps(G.check_fused_front(**kw)), flush=True)
    except Exception as e:  # noqa: BLE001
        import newsrec_b200
        import ctypes as C
        out = (C.c_int * 4)()
        print("FAILED", kw, repr(e), flush=True)
        try:
            newsrec_b200.load_library().nr_device_error(out)
            print("device error record:", list(out), flush=True)
        except Exception as e2:  # noqa: BLE001
            print("no device record:", e2)
        break
