"""Data-parallel step on real GPUs (needs >= 2 devices; skipped otherwise): two ranks with different micro-batches must end the
step with IDENTICAL parameters, equal to a single-process step on the averaged per-replica gradients (SURVEY 8e equivalence test)."""
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
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
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
grads_of(m, 100 + r)
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
print("rank", r, "ok", diff)
'''


@pytest.mark.gpu
def test_dp_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    script = tmp_path / "dp_gpu.py"
    script.write_text(_SCRIPT % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29544", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
