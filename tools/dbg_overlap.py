"""debug: RNN-head gradient error vs the fp64 oracle for several seeds, for the library given in CRNN_DBG_LIB."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crnn_b200 as cb
if os.environ.get("CRNN_DBG_LIB"):
    cb._lib.LIB_PATH = os.environ["CRNN_DBG_LIB"]
from oracle import crnn_oracle as N
cfg = N.Cfg(imgh=100, cell="gru")
B = 6
d = "cuda"
out = {}
for seed in (3, 4, 5):
    w = N.randomize_for_test(N.init_weights(cfg, seed), seed)
    x, lab, L, il = N.synth_batch(cfg, B, 30 + seed)
    m = cb.CRNN(cfg.num_classes, cfg.max_len, (cfg.imgh, cfg.imgw, 1), cfg.time_dense, True, cfg.n_units, max_batch=B).get_model()
    m.set_weights(w)
    args = (torch.tensor(x, device=d), torch.tensor(lab, device=d), torch.tensor(L, device=d), torch.tensor(il, device=d))
    per = m.train_fwd_bwd_device(*args, dropout_seed=0).cpu().numpy()
    g = m.get_grads()
    loss64, per64, g64, _, keep = N.loss_and_grads(w, x, lab, L, il, cfg, dtype=torch.float64)
    print("seed", seed, "loss err", np.abs(per - per64).max())
    for k in g64:
        if k.startswith(("dense2", "bidirectional", "dense1")):
            sc = np.abs(g64[k]).max()
            print("  %-52s %.2e" % (k, np.abs(g[k] - g64[k]).max() / sc))
    for nm in ("theta", "block1", "block7", "dense1", "hs2", "softmax", "dlogits"):
        out["%d_%s" % (seed, nm)] = m.activation(nm).copy()
    for layer in (1, 2):
        gt = m.activation("gates%d" % layer).reshape(B, cfg.T, 2, 3, 256)
        for dd in (0, 1):
            zr = gt[:, :, dd, :2]
            near = ((zr > 0) & (zr < 3e-6)) | ((zr < 1) & (zr > 1 - 3e-6))
            print("  layer %d dir %d: gate values within 3e-6 of a hard-sigmoid kink: %d ; saturated %.3f" % (layer, dd, int(near.sum()), float(((zr == 0) | (zr == 1)).mean())))
    for nm in ("theta", "block1", "block7", "dense1"):
        kk = keep[nm].detach().numpy().reshape(-1)
        print("  act %-8s err %.3e" % (nm, np.abs(m.activation(nm)[:kk.size] - kk).max()))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/dbg_%s.npz" % os.environ.get("CRNN_DBG_TAG", "new"), **out)
