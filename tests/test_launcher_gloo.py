"""newsrec_b200.launch around a stand-in trainer with the reference loop's shape (DataLoader with
shuffle, Adam, zero_grad/backward/step, evaluate, torch.save), two gloo ranks on CPU: disjoint shards, identical
parameters on both ranks, equal to a single-process replay on the union batches, rank-0-only side effects."""
import json
import os
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "news-recommendation_b200", "src")

FAKE_TRAIN = textwrap.dedent('''
    from torch.utils.data import DataLoader
    from torch.utils.tensorboard import SummaryWriter
    import json, os, torch

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 64
        def __getitem__(self, i):
            x = torch.tensor([i / 64.0, (i % 7) / 7.0, (i % 3) / 3.0])
            return i, x, (x * torch.tensor([1.0, -2.0, 0.5])).sum() + 0.25

    def evaluate(model, directory):
        return 0.5 + float(model.weight.sum()) * 0.0, 0.4, 0.3, 0.2

    def train():
        rank = os.environ.get("RANK", "0")
        writer = SummaryWriter(log_dir=os.path.join(os.environ["FAKE_OUT"], "tb" + rank))
        torch.manual_seed(0)
        model = torch.nn.Linear(3, 1)
        dataset = DS()
        mk = lambda: iter(DataLoader(dataset, batch_size=4, shuffle=True, num_workers=0, drop_last=True, pin_memory=True))
        loader = mk()
        optimizer = torch.optim.Adam(model.parameters(), lr=0.05)
        log = []
        for i in range(12):
            try:
                idx, x, y = next(loader)
            except StopIteration:
                loader = mk()
                idx, x, y = next(loader)
            loss = ((model(x).squeeze(-1) - y) ** 2).mean()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            writer.add_scalar("Train/Loss", loss.item(), i)
            log.append(idx.tolist())
        metrics = evaluate(model, "unused")
        torch.save({"w": model.weight.detach()}, os.path.join(os.environ["FAKE_OUT"], "ckpt" + rank + ".pt"))
        with open(os.path.join(os.environ["FAKE_OUT"], "out" + rank + ".json"), "w") as f:
            json.dump({"w": model.weight.detach().flatten().tolist(), "b": model.bias.detach().tolist(), "log": log,
                       "metrics": list(metrics), "grad_is_flat_view": model.weight.grad.data_ptr() == optimizer._flat.flat.data_ptr()}, f)
''')


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      FAKE_OUT=tmp, CUDA_VISIBLE_DEVICES="")
    sys.path.insert(0, SRC)
    from newsrec_b200 import launch
    launch.main(["--reference-src", tmp, "--no-dropin", "--backend", "gloo", "--seed", "3"])


def test_launcher_shards_and_all_reduces_like_one_process(tmp_path):
    pytest.importorskip("tensorboard")
    tmp = str(tmp_path)
    with open(os.path.join(tmp, "train.py"), "w") as f:
        f.write(FAKE_TRAIN)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    out = [json.load(open(os.path.join(tmp, f"out{r}.json"))) for r in range(2)]
    # both ranks end with bit-identical parameters, gradients lived in the flat all-reduce buffer
    assert out[0]["w"] == out[1]["w"] and out[0]["b"] == out[1]["b"]
    assert out[0]["grad_is_flat_view"] and out[1]["grad_is_flat_view"]
    # shards are disjoint within every epoch (8 steps per epoch) and differ between the two epochs' permutations
    for e in range(2):
        seen0 = {i for step in out[0]["log"][8 * e:8 * e + 8] for i in step}
        seen1 = {i for step in out[1]["log"][8 * e:8 * e + 8] for i in step}
        assert not (seen0 & seen1)
    assert out[0]["log"][:4] != out[0]["log"][8:12]
    # rank 0 alone evaluates (broadcast), writes TensorBoard events and checkpoints
    assert out[0]["metrics"] == out[1]["metrics"] == [0.5, 0.4, 0.3, 0.2]
    assert os.path.exists(os.path.join(tmp, "ckpt0.pt")) and not os.path.exists(os.path.join(tmp, "ckpt1.pt"))
    assert os.path.isdir(os.path.join(tmp, "tb0")) and not os.path.isdir(os.path.join(tmp, "tb1"))
    # single-process replay on the union batches gives the same parameters
    torch.manual_seed(0)
    model = torch.nn.Linear(3, 1)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    feat = lambda i: torch.tensor([i / 64.0, (i % 7) / 7.0, (i % 3) / 3.0])
    for s0, s1 in zip(out[0]["log"], out[1]["log"]):
        x = torch.stack([feat(i) for i in s0 + s1])
        y = (x * torch.tensor([1.0, -2.0, 0.5])).sum(1) + 0.25
        loss = ((model(x).squeeze(-1) - y) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert torch.allclose(model.weight.detach().flatten(), torch.tensor(out[0]["w"]), atol=1e-5)
    assert torch.allclose(model.bias.detach(), torch.tensor(out[0]["b"]), atol=1e-5)


REFERENCE_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE_SRC, "train.py")), reason="reference checkout not present")
def test_launcher_drives_the_unmodified_reference_trainer_on_cpu(tmp_path):
    """The reference's own train.py + dataset.py + CPU model (--no-dropin), two gloo ranks under torchrun, on synthetic
    parsed-MIND files: shards differ per rank, the loader is re-created on exhaustion, only rank 0 writes TensorBoard."""
    import subprocess
    tmp = str(tmp_path)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synth_mind.py"), tmp, "24", "60", "2"], check=True,
                   stdout=subprocess.DEVNULL)
    env = dict(os.environ, PYTHONPATH=SRC, CUDA_VISIBLE_DEVICES="", MODEL_NAME="NRMS", OMP_NUM_THREADS="4")
    port = 29700 + (os.getpid() % 90)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "newsrec_b200.launch", "--reference-src", REFERENCE_SRC, "--no-dropin", "--backend", "gloo",
           "--set", "batch_size=4", "--set", "num_workers=0", "--set", "num_epochs=1", "--set", "num_batches_show_loss=2"]
    r = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "Load training dataset with size 24." in out
    losses = [ln for ln in out.splitlines() if "current loss" in ln]
    assert len(losses) >= 2 * 3  # both ranks report at batches 2, 4, 6
    assert out.count("Training data exhausted") >= 2  # each rank's 12-sample shard runs out after 3 batches of 4
    runs = os.listdir(os.path.join(tmp, "runs", "NRMS"))
    assert len(runs) == 1  # rank 0 alone created a TensorBoard run directory
