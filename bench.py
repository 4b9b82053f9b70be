"""bench.py -- impressions/sec, forward+backward, NRMS on MIND-shaped synthetic batches (BASELINE.json).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port)

One "step" = zero_grad -> forward -> CrossEntropy(label 0) -> backward over one batch of B=512 impressions
per GPU (1+K=5 candidates + 50 browsed titles of 20 tokens each = 55 news encodes + 1 user encode + 5
scores), plus -- for N>1 -- the one NCCL gradient all-reduce.  The optimizer step is excluded: the
metric is forward+backward.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "news-recommendation_b200", "src"))

import torch  # noqa: E402

V_WORDS, T_TITLE, H_HIST, K_NEG, D_MODEL, HEADS, Q_DIM = 70976, 20, 50, 4, 300, 15, 200
C_CAND = 1 + K_NEG
N_NEWS = C_CAND + H_HIST
# algorithmic work per impression (padding to MMA shapes is NOT counted), forward; x3 for fwd+bwd
FLOP_FWD_PER_IMPRESSION = (N_NEWS * T_TITLE * 6 * D_MODEL ** 2 + N_NEWS * HEADS * 4 * T_TITLE ** 2 * (D_MODEL // HEADS)
                           + N_NEWS * T_TITLE * 2 * D_MODEL * Q_DIM + N_NEWS * T_TITLE * (2 * Q_DIM + 2 * D_MODEL)
                           + H_HIST * 6 * D_MODEL ** 2 + HEADS * 4 * H_HIST ** 2 * (D_MODEL // HEADS)
                           + H_HIST * 2 * D_MODEL * Q_DIM + C_CAND * 2 * D_MODEL)


def synth_slots(B, seed, device="cpu", pin=False):
    """MIND-shaped ids exactly as default_collate hands them to the model: slot-major lists of {"title": (B,T)}.
    Title length U{5..20} right-padded with 0; history length U{1..50}, LEFT-padded with all-zero news."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, V_WORDS, (B, N_NEWS, T_TITLE), generator=g)
    tl = torch.randint(5, T_TITLE + 1, (B, N_NEWS, 1), generator=g)
    ids = ids * (torch.arange(T_TITLE).view(1, 1, -1) < tl)
    hl = torch.randint(1, H_HIST + 1, (B, 1), generator=g)
    keep = torch.arange(H_HIST).view(1, -1) >= (H_HIST - hl)
    ids[:, C_CAND:] = ids[:, C_CAND:] * keep.unsqueeze(-1)

    def mk(t):
        t = t.contiguous()
        if device != "cpu":
            return t.to(device)
        return t.pin_memory() if pin else t

    cand = [{"title": mk(ids[:, j])} for j in range(C_CAND)]
    clicked = [{"title": mk(ids[:, C_CAND + j])} for j in range(H_HIST)]
    return cand, clicked


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def kernel_work(name):
    """Algorithmic FLOPs and compulsory HBM bytes of ONE launch, from the '<ctx>/<op>[a,b,c]' profile key."""
    op = name.split("/")[1].split("[")[0]
    a, b, c = [int(x) for x in name.split("[")[1].rstrip("]").split(",")]
    if op in ("gemm_store", "gemm_scatter_emb"):          # [M, N, K]
        return 2.0 * a * b * c, 2.0 * a * c + (2.0 * a * b if op == "gemm_store" else 4.0 * a * b)
    if op in ("gemm_additive_pool", "gemm_additive_dpre"):  # [M, q, D]
        return 2.0 * a * b * c + 4.0 * a * b, 2.0 * a * c * (2 if op == "gemm_additive_pool" else 1) + 2.0 * a * b
    if op == "gemm_pool_dinput":                           # [M, D, q]
        return 2.0 * a * b * c, 2.0 * a * c + 2.0 * a * b
    if op == "gemm_tn":                                    # [Kr, Ma, Nb]
        return 2.0 * a * b * c, 2.0 * a * (b + c)
    if op == "mhsa_core_fwd":                              # [n_seq, T, d]
        return 4.0 * a * b * b * c, 2.0 * a * b * c * 4
    if op == "mhsa_core_bwd":
        return 10.0 * a * b * b * c, 2.0 * a * b * c * 7
    if op == "gather_rows":                                # [n_tok, D, ld]
        return 0.0, 2.0 * a * c * 2
    if op == "pool_dscore":                                # [n_seg, seg_len, D]
        return 2.0 * a * b * c, 2.0 * a * b * c
    return 0.0, 0.0


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures summarised in
# profiles/ncu_r01_kernel_summary.csv (batch 512 shapes only; other shapes report null)
NCU_DRAM_BYTES = {
    "mhsa_core_bwd[28160,20,300]": 1.483482e9 + 1.050252e9,   # 48-byte-tile kernels, profiles/ncu_r01_attention_final.csv
    "mhsa_core_fwd[28160,20,300]": 1.055464e9 + 0.329299e9,
    "gemm_store[563200,900,300]": 0.443694e9 + 0.979511e9,
    "gemm_additive_pool[563200,200,300]": 0.374886e9 + 0.030811e9,
    "gemm_additive_dpre[563200,200,300]": 0.344909e9 + 0.191328e9,
}


def usable_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota (os.cpu_count()
    reports the whole machine inside a container; oversubscribing OpenMP threads stalls the CPU arm)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except (OSError, ValueError):
            pass
    return min(n, 64)


def log(msg):
    print(f"[bench] {msg}", file=sys.stderr, flush=True)


def run_reference(args):
    """The reference's own CPU path (oracle port with the reference's per-slot call structure), all usable host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import newsrec_oracle as O
    cores = usable_cores()
    torch.set_num_threads(cores)
    log(f"reference arm: {cores} threads, batch {args.ref_batch}, {args.warmup}+{args.steps} steps")
    B = args.ref_batch
    model = O.ReferenceStructuredNRMS(V_WORDS, D_MODEL, HEADS, Q_DIM, 0.2, 0)
    model.train()
    batches = [synth_slots(B, 100 + i) for i in range(2)]
    for i in range(args.warmup):
        O.reference_cpu_step(model, *batches[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        O.reference_cpu_step(model, *batches[i % 2])
    dt = time.perf_counter() - t0
    return B * args.steps / dt, dt / max(args.steps, 1) * 1e3, cores, B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="impressions per GPU per step (BASELINE.json configs[1])")
    ap.add_argument("--ref-batch", type=int, default=64, help="impressions per CPU step of the reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = f"NRMS bf16 fwd+bwd: batch={args.batch}/GPU, title_len={T_TITLE}, history={H_HIST}, K={K_NEG}, {HEADS} heads x d_k={D_MODEL // HEADS}, d={D_MODEL}, V={V_WORDS}"

    if args.impl == "reference":
        if rank != 0:
            return
        val, ms, cores, B = run_reference(args)
        print(json.dumps({
            "metric": "impressions/sec (fwd+bwd)", "value": val, "unit": "impressions/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload, "reference_step_batch": B},
            "cpu_baseline": {"value": val, "unit": "impressions/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps of {B} impressions (same shapes), fp32, torch CPU, train mode"},
            "e2e": {"value": val, "unit": "impressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import newsrec_b200
    from newsrec_b200 import ddp
    rank, world, local = ddp.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import config as cfgmod
    from model.NRMS import NRMS
    lib = newsrec_b200.load_library()
    torch.manual_seed(0)
    model = NRMS(cfgmod.NRMSConfig).to(dev)
    model.train()  # dropout active exactly as in the reference's training step
    grads = ddp.FlatGradients(model.parameters(), world)
    B = args.batch
    n_rot = 3  # rotate distinct batches; X/QKV intermediates (>1.5 GB per step) far exceed the 126 MB L2
    dev_batches = [synth_slots(B, 1000 * rank + i, device=dev) for i in range(n_rot)]
    host_batches = [synth_slots(B, 1000 * rank + 50 + i, pin=True) for i in range(n_rot)]
    label = torch.zeros(B, dtype=torch.long, device=dev)

    def step(batch, read_loss=False):
        grads.zero()
        logits = model(batch[0], batch[1])
        loss = torch.nn.functional.cross_entropy(logits, label)
        loss.backward()
        grads.all_reduce_mean()
        return loss.item() if read_loss else None

    def timed_e2e(batches):
        """End to end with HOST batches through the public API: every timed step stages one batch (host stacking into
        pinned memory + H2D + device re-ordering, `NRMS.prefetch`, on a copy stream while the previous step's kernels
        run) and reads its own loss back (D2H, synchronising).  `steps` copies and `steps` reads inside the timed region."""
        def run(n):
            cur = model.prefetch(*batches[0])
            for i in range(n):
                grads.zero()
                loss = torch.nn.functional.cross_entropy(model(cur), label)
                loss.backward()
                grads.all_reduce_mean()
                if i + 1 < n:
                    cur = model.prefetch(*batches[(i + 1) % n_rot])  # overlaps the kernels enqueued above
                loss.item()
        run(args.warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(args.steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(batches, read_loss, profile=False):
        for i in range(args.warmup):
            step(batches[i % n_rot], read_loss)
        barrier()
        if profile:  # per-kernel CUDA events cover the TIMED steps only (first launches pay lazy module loading)
            lib.nr_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = newsrec_b200.launch_count()
        e0.record()
        for i in range(args.steps):
            step(batches[i % n_rot], read_loss)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), newsrec_b200.launch_count() - l0

    log(f"rank {rank}/{world}: model + data ready; timing device-resident steps")
    # ---- device-resident inputs: `value` ----
    with ClockSampler(local) as clk:
        ms_total, launches = timed(dev_batches, read_loss=False)
    log(f"device-resident: {ms_total / args.steps:.3f} ms/step; per-kernel pass")
    # ---- per-kernel durations: a separate pass (the event pairs around every launch are kept out of `value`) ----
    timed(dev_batches, read_loss=False, profile=True)
    prof = newsrec_b200.profile_report()
    lib.nr_profile_enable(0)
    log("timing end-to-end steps")
    # ---- end to end through the public API with HOST buffers (H2D of ids + D2H of the loss inside) ----
    ms_e2e = timed_e2e(host_batches)
    log(f"end-to-end: {ms_e2e / args.steps:.3f} ms/step")

    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    pk = peaks()
    imp = B * world * args.steps
    value = imp / (ms_total / 1e3)
    e2e = imp / (ms_e2e / 1e3)
    # dominant kernel of the step and its roofline (timed inside a long step -> sustained tensor peak)
    steps_profiled = args.steps
    tot_prof = sum(v[1] for v in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    flops, bytes_ = kernel_work(dom[0])
    dur_s = dom[1][1] / dom[1][0] / 1e3
    tf, gbs = flops / dur_s / 1e12, bytes_ / dur_s / 1e9
    if tf / pk["bf16_tflops_sustained"] >= gbs / pk["hbm_gbs"]:
        roof = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops_sustained"]}
    else:
        roof = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}
    roof.update({"kernel": dom[0], "share_of_step": dom[1][1] / tot_prof, "avg_launch_ms": dur_s * 1e3,
                 "traffic": NCU_DRAM_BYTES.get(dom[0].split("/")[1]), "algorithmic_bytes": bytes_,
                 "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)"})
    step_tf = value / world * 3 * FLOP_FWD_PER_IMPRESSION / 1e12
    breakdown = {k: round(v[1] / steps_profiled, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    out = {
        "metric": "impressions/sec (fwd+bwd)", "value": value, "unit": "impressions/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": "3 rotating batches; per-step intermediates (~1.7 GB) exceed the 126 MB L2", "dropout": 0.2},
        "e2e": {"value": e2e, "unit": "impressions/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": B * N_NEWS * T_TITLE * 8, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": roof,
        "roofline_step": {"bound": "tensor", "achieved": step_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                          "frac": step_tf / pk["bf16_tflops_sustained"],
                          "note": "whole step: algorithmic 3 x %.1f MFLOP per impression / step time, per GPU" % (FLOP_FWD_PER_IMPRESSION / 1e6)},
        "kernel_ms_per_step": breakdown,
    }
    if world == 1 and not args.no_cpu_baseline:
        ra = argparse.Namespace(steps=3, warmup=1, ref_batch=args.ref_batch)
        val, ms, cores, rb = run_reference(ra)
        out["cpu_baseline"] = {"value": val, "unit": "impressions/s", "cores": cores, "kind": "port",
                               "sample": f"3 steps of {rb} impressions (same shapes), fp32, torch CPU, train mode, {ms:.0f} ms/step"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
