"""Data-parallel step on real GPUs: two ranks with different micro-batches must end the step with IDENTICAL parameters, equal to a
single-process step on the averaged per-replica gradients (SURVEY 8e equivalence test).  With >= 2 devices the ranks use NCCL (one GPU
each); on a 1-GPU box both ranks share cuda:0 and exchange through gloo (NCCL refuses duplicate devices), so the test never skips --
the engine, the reduce-then-clip-then-update order and the 1/world scaling are the same code either way."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import crnn_b200 as cb
from oracle import crnn_oracle as N
local = int(os.environ["LOCAL_RANK"]) %% torch.cuda.device_count(); torch.cuda.set_device(local)
cb.parallel.init_distributed(device=torch.device("cuda", local))
r, w = cb.parallel.rank(), cb.parallel.world_size()
cfg = N.Cfg(imgh=100, cell="gru"); B = 4
weights = N.randomize_for_test(N.init_weights(cfg, 5), 5)
def make():
    m = cb.CRNN(cfg.num_classes, cfg.max_len, (cfg.imgh, cfg.imgw, 1), cfg.time_dense, True, cfg.n_units, max_batch=B).get_model()
    m.set_weights(weights); m.compile(optimizer=cb.Adam(lr=1e-4, beta_1=0.5, beta_2=0.999, epsilon=1e-7, clipnorm=5.0)); return m
d = torch.device("cuda", local)
def grads_of(m, seed):
    x, lab, L, il = N.synth_batch(cfg, B, seed)
    m.train_fwd_bwd_device(torch.tensor(x, device=d), torch.tensor(lab, device=d), torch.tensor(L, device=d), torch.tensor(il, device=d), dropout_seed=0)
    return m.tensor("arena/grads").clone()
m = make()
mode = m.enable_native_dp(fused=(os.environ.get("CRNN_TEST_DP_FUSED", "1") == "1"))   # NCCL group: the engine's own communicator; gloo: None
grads_of(m, 100 + r)
if mode == "fused":           # the step already reduced its gradients (two buckets): every rank holds the same SUM
    gsum = m.tensor("arena/grads").clone(); ref_g = gsum.clone(); cb.parallel.broadcast_(ref_g, 0)
    assert torch.equal(ref_g, gsum), "fused all-reduce left different gradients on the ranks"
scale = m.allreduce_grads()
assert scale == 1.0 / w
m.optimizer_step(scale)
mine = m.tensor("arena/params").clone()
# every rank holds the same parameters
ref = mine.clone(); cb.parallel.broadcast_(ref, 0)
assert torch.equal(ref, mine), "replicas diverged"
# single-process reference: average of the per-replica gradients, then clip + Adam
m2 = make()
g = sum(grads_of(m2, 100 + k) for k in range(w))
m2.tensor("arena/grads").copy_(g)
m2.optimizer_step(1.0 / w)
diff = (m2.tensor("arena/params") - mine).abs().max().item()
assert diff < 1e-6, diff     # gradient atomics make the two evaluations differ by rounding only
sys.stdout.write("rank %%d ok %%g mode %%s\n" %% (r, diff, mode)); sys.stdout.flush()
'''


def _torchrun(args, port, extra_env=None, timeout=600):
    import torch
    env = dict(os.environ, PYTHONPATH=ROOT)
    if torch.cuda.device_count() < 2:
        env["CRNN_DIST_BACKEND"] = "gloo"
    env.update(extra_env or {})
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port)] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", ["1", "0"])
def test_dp_two_ranks(tmp_path, fused):
    """fused=1: the step's own bucketed NCCL all-reduce (crnn_set_dp_fused); fused=0: crnn_allreduce_grads called after the step.  On a
    1-GPU box both variants run the torch.distributed/gloo exchange (NCCL cannot put two ranks on one device)."""
    import torch
    script = tmp_path / "dp_gpu.py"
    script.write_text(_SCRIPT % {"root": ROOT})
    r = _torchrun([str(script)], 29544 + int(fused), extra_env={"CRNN_TEST_DP_FUSED": fused})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
    if torch.cuda.device_count() >= 2:
        assert ("mode fused" if fused == "1" else "mode call") in r.stdout, r.stdout[-500:]


@pytest.mark.gpu
def test_dp_train_cli_uneven_shards_and_early_stopping(tmp_path):
    """`train.py` under torchrun with a file count that gives the two ranks shards of 9 and 8 files at batch 8 -- local step counts 2 and 1,
    the case that used to hang in the last gradient all-reduce -- plus --early_stopping, whose decision must be taken on the all-reduced
    loss by both ranks.  Both ranks must finish, and their final parameters must be identical."""
    import cv2
    import numpy as np
    img_dir = tmp_path / "imgs"; img_dir.mkdir()
    rng = np.random.default_rng(0)
    words = ["hello", "world", "ocr", "lite", "b200", "crnn", "text", "line"]
    for i in range(19):                                   # train_portion .9 -> 17 train files -> shards 9 / 8
        w = words[i % len(words)]
        img = np.full((32, 100), 255, np.uint8)
        cv2.putText(img, w, (2, 24), cv2.FONT_HERSHEY_SIMPLEX, 0.8, int(rng.integers(0, 80)), 2)
        cv2.imwrite(str(img_dir / ("%d_%s_%d.png" % (i, w, i))), img)
    r = _torchrun(["train.py", "--path", str(img_dir), "--save_path", str(tmp_path), "--model_name", "m", "--nbepochs", "2", "--batch_size", "8",
                   "--opt", "sgd", "--lr", "0.001", "--imgh", "100", "--imgW", "32", "--train_portion", "0.9", "--early_stopping", "3"],
                  29545, extra_env={"CRNN_DP_DUMP_PARAMS": str(tmp_path / "params")}, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert (tmp_path / "m" / "final_weights.h5").exists()
    a = np.load(str(tmp_path / "params.rank0.npy")); b = np.load(str(tmp_path / "params.rank1.npy"))
    assert np.isfinite(a).all() and np.isfinite(b).all(), "non-finite parameters after training"
    assert a.size > 2_800_000 and np.array_equal(a, b), "replicas diverged"
