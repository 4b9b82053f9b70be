"""Run-to-run spread of the NRMS parameter gradients: the in-place accumulation path (b) and a second run of the
return-the-gradients path (c) against a first run (a).  Shows that the only O(1e-5) spread is W_K.bias, whose gradient is
analytically zero (softmax shift invariance) and therefore pure fp32 accumulation noise in atomic order."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for q in ("tests", "oracle", os.path.join("news-recommendation_b200", "src")):
    sys.path.insert(0, os.path.join(R, q))
import torch
import gpu_checks as G
from gpu_checks import O, DEV, slots, nrms_model_and_params
from newsrec_b200 import ddp
B, Cn, H, T, V, seed = 8, 5, 50, 20, 500, 6
cand_t, clicked_t, _ = O.synth_batch(B, Cn, H, T, V, seed * 100)
label = torch.zeros(B, dtype=torch.long, device=DEV)
for trial in range(6):
    ma, _ = nrms_model_and_params(V, seed); mb, _ = nrms_model_and_params(V, seed); mc, _ = nrms_model_and_params(V, seed)
    ma.eval(); mb.eval(); mc.eval()
    flat = ddp.FlatGradients(mb.parameters(), 1); flat.zero()
    for m in (ma, mb, mc):
        torch.nn.functional.cross_entropy(m(slots(cand_t), slots(clicked_t)), label).backward()
    torch.cuda.synchronize()
    worst = []
    for (k, pa), (_, pb), (_, pc) in zip(ma.named_parameters(), mb.named_parameters(), mc.named_parameters()):
        sc = float(pa.grad.abs().max()) + 1e-12
        worst.append((float((pa.grad - pb.grad).abs().max()) / sc, float((pa.grad - pc.grad).abs().max()) / sc, k))
    worst.sort(reverse=True)
    print(trial, "a-vs-b(inplace)", ["%.2e %s" % (w[0], w[2][-40:]) for w in worst[:3]], flush=True)
    worst.sort(key=lambda w: -w[1])
    print(trial, "a-vs-c(same path)", ["%.2e %s" % (w[1], w[2][-40:]) for w in worst[:3]], flush=True)
