#!/usr/bin/env python
"""bench.py -- measurement contract of the CRNN-OCR hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU (oracle port)

A "step" = one full training step of BASELINE.json configs[2]: STN + conv stack + BiRNN forward/backward +
ctc_batch_cost gradient + clipnorm/Adam on one batch of 64 synthetic 128x32 text-line images per GPU (fp32, random-init
weights, dropout on).  N>1: data parallel, batch 64 per GPU (weak scaling), one NCCL all-reduce of the flat gradient
arena per step.  `--scaling strong` fixes the GLOBAL batch at 512 instead (BASELINE configs[4]: 512/N images per GPU); the default weak run
also reports that configuration as `configs4_strong` when N > 1.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "text-line images/sec (fwd+bwd+CTC) per GPU; CTC beam-10 decode lines/sec"
BATCH = 64
GLOBAL_STRONG = 512          # BASELINE configs[4]: global batch of the strong-scaling data-parallel step
IMGH, IMGW, V, MAXLEN = 128, 32, 38, 23
BLOCKS = [(1, 64, 1, 1), (64, 128, 1, 1), (128, 256, 2, 2), (256, 256, 1, 1), (256, 512, 1, 2), (512, 512, 1, 1), (512, 512, 1, 1)]


def conv_stack_flops_per_image():
    """depthwise + pointwise MACs of the 7 blocks on a 128x32 image, x2 FLOP, x3 for forward + dX + dW  (= 4.555 GFLOP)."""
    h, w, mac = IMGH + 4, IMGW + 4, 0
    for cin, cout, ph, pw in BLOCKS:
        mac += h * w * (9 * cin + cin * cout)
        h, w = h // ph, w // pw
    return 3 * 2 * mac


ALGORITHMIC_FUSED_BYTES_PER_IMAGE = 90e6      # HBM bytes fwd+bwd, training, fused estimate (5.8 GB per batch of 64)


def synth_batch(B, seed):
    """SURVEY 8d configs[2]: normalised U{0..255} pixels, L~U{3..23} labels padded with the blank, input_length = T-2."""
    rng = np.random.default_rng(seed)
    u8 = rng.integers(0, 256, (B, IMGH, IMGW, 1))
    synth_batch.last_u8 = u8.astype(np.uint8)               # the same images as 8-bit data (input of the device-side normalisation)
    x = ((u8.astype(np.float32) - np.float32(118.24236953981779)) / np.float32(36.72835353999682))
    L = rng.integers(3, MAXLEN + 1, B).astype(np.int32)
    lab = np.full((B, MAXLEN), V - 1, np.int32)
    for b in range(B):
        lab[b, :L[b]] = rng.integers(0, V - 1, L[b])
    il = np.full(B, (IMGH + 4) // 2 - 2, np.int32)
    return x.astype(np.float32), lab, L, il


def per_gpu_batch(args, world):
    if getattr(args, "scaling", "weak") == "strong":
        if GLOBAL_STRONG % world:
            raise SystemExit("--scaling strong needs a GPU count that divides %d" % GLOBAL_STRONG)
        return GLOBAL_STRONG // world
    return BATCH


def workload_config(args, world):
    pb = per_gpu_batch(args, world)
    cfgname = "configs[4]: data-parallel train step, GLOBAL batch 512 fixed" if getattr(args, "scaling", "weak") == "strong" else "configs[2]: full train step"
    return {"workload": "%s (STN+dw-separable conv stack+Bi%s fwd/bwd + ctc_batch_cost grad + clipnorm5/Adam), "
                        "batch %d per GPU, 128x32 gray, V=38, fp32, random-init, dropout on" % (cfgname, args.cell.upper(), pb),
            "global_batch": pb * world, "per_gpu_batch": pb, "imgh": IMGH, "imgw": IMGW, "num_classes": V, "cell": args.cell,
            "parallelism": "dp%d" % world,
            "gradient_exchange": ("none (1 GPU)" if world == 1 else "2-bucket NCCL all-reduce issued by the step itself (head bucket overlapped with the conv-stack backward), "
                                  "or one torch.distributed all-reduce after the step when CRNN_DP_NATIVE=0"),
            "l2": "per-step working set (~2.5 GB of activations/gradients rewritten every step) >> 126 MB L2; no explicit flush"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


def profiled_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel` from the committed `ncu --set full` summaries
    (profiles/r02_ncu_full_summary.json, else r01): the block-6 launch for the GEMMs (the deep-K shape that dominates their time)."""
    want = {"xw_gemm_tc_v2_kernel": "xw_fwd_b6", "xty_gemm_tc_kernel": "xty_dw_b6", "gru/lstm_{fwd,bwd}_mma_kernel": "gru_fwd_mma",
            "dwconv3x3_rows_*": "dwrows_bwd", "stage:act_pool_bwd": "actbwd", "stage:bn_bwd": "relu6bwd"}.get(kernel)
    if want is None:
        return None
    for name in ("r02_ncu_full_summary.json", "r01_ncu_full_summary.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        recs = [r for r in json.load(open(p)) if r.get("capture") == want and "dram_read_MB" in r]
        if recs:
            r = recs[0]            # the summaries are kept newest capture set first (tools/profiles_merge.py)
            return {"bytes_per_launch": (r["dram_read_MB"] + r["dram_write_MB"]) * 1e6, "launch": "%s (%s), %.1f us under ncu" % (r["kernel"].strip(), want, r["time_us"]),
                    "tensor_pipe_pct_ncu": r.get("tensor_pipe_pct"), "source": "profiles/" + name}
    return None


def step_dram_bytes():
    """Measured DRAM bytes of ONE training step (sum over all its kernels of dram__bytes_read + dram__bytes_write, one ncu pass over an
    eager step, committed as profiles/r02_step_dram.json by tools/ncu_step_dram.py); None when the file is missing."""
    p = os.path.join(ROOT, "profiles", "r02_step_dram.json")
    return json.load(open(p)) if os.path.exists(p) else None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[5 + i].strip().lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ----------------------------------------------------------------------------------------------- CPU arm
def oracle_train_step_fn(cell, sample_batch, threads):
    import torch
    from oracle import crnn_oracle as N
    torch.set_num_threads(threads)
    cfg = N.Cfg(imgh=IMGH, imgw=IMGW, num_classes=V, cell=cell, max_len=MAXLEN)
    w = N.init_weights(cfg, 0)
    x, lab, L, il = synth_batch(sample_batch, 2)
    state = {}

    def step():
        nonlocal w
        loss, per, g, stats, _ = N.loss_and_grads(w, x, lab, L, il, cfg)
        w, _ = N.adam_step(w, g, state, lr=1e-4, b1=0.5, b2=0.999, eps=1e-7, clipnorm=5.0)
        w.update(stats)
        return loss
    return step


def pick_cpu_threads(cell):
    """The CPU arm uses 'all the host threads it can use': the thread count (<= cores) that runs the oracle fastest
    (on many-core hosts the small per-layer ops slow down beyond a few dozen threads)."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32) if c <= cores} | {min(cores, 8)})
    best, best_t = cands[0], None
    for c in cands:
        st = oracle_train_step_fn(cell, 4, c)
        st()
        t0 = time.perf_counter(); st(); dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = pick_cpu_threads(args.cell)
    sample = BATCH                                         # one full batch of the workload per step, no sub-sampling
    step = oracle_train_step_fn(args.cell, sample, cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    desc = ("oracle port (PyTorch-CPU fp32 restatement of the Keras graph + C restatement of TF CTCLoss + Keras Adam): every step is one full "
            "train step on a batch of %d images, %d host threads" % (sample, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, args.gpus), reference_batch_per_step=sample),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
def _emit(line, fd):
    os.write(fd, (json.dumps(line) + "\n").encode())


def run_ours(args):
    # stdout carries exactly ONE JSON line: everything else that libraries print there (e.g. "NCCL version ..." at communicator
    # creation) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run; the line is written to the saved descriptor
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        _emit({"error": "no CUDA device: this path has no CPU fallback"}, out_fd)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line (NCCL_DEBUG=VERSION/INFO prints to stdout)
        dist.init_process_group("nccl", device_id=dev)
    import crnn_b200 as cb
    lib = cb._lib.load()

    PB = per_gpu_batch(args, world)                          # images per GPU per step (64; 512 / N with --scaling strong)
    model = cb.CRNN(V, MAXLEN, (IMGH, IMGW, 1), 128, args.cell == "gru", 256, max_batch=PB, seed=1234).get_model()
    model.compile(optimizer=cb.Adam(lr=1e-4, beta_1=0.5, beta_2=0.999, epsilon=1e-7, clipnorm=5.0))
    dp_mode = None
    if world > 1:   # identical replicas; the engine's own NCCL communicator reduces the gradients inside the step (CRNN_DP_NATIVE=0: torch.distributed)
        dist.broadcast(model.tensor("arena/params"), src=0)
        dp_mode = model.enable_native_dp()
    x, lab, L, il = synth_batch(PB, 2 + rank)
    x_u8 = synth_batch.last_u8
    xd, labd, Ld, ild = (torch.tensor(a, device=dev) for a in (x, lab, L, il))
    seed_base = 0x5EED0000 + rank

    step_no = [0]

    def step_device():
        step_no[0] += 1
        model.train_fwd_bwd_device(xd, labd, Ld, ild, dropout_seed=seed_base + step_no[0])
        model.optimizer_step(model.allreduce_grads())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = lib.crnn_launch_count()
    ms = timed(step_device, args.steps)
    launches = lib.crnn_launch_count() - l0
    value = PB * world * args.steps / (ms / 1e3)

    # ---- e2e: host numpy in -> train_on_batch (pinned staging + H2D, step, D2H of the loss) every step
    host_inputs = {"the_input": x, "the_labels": lab, "input_length": il.reshape(-1, 1), "label_length": L.reshape(-1, 1)}
    for _ in range(2):
        model.train_on_batch(host_inputs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.train_on_batch(host_inputs)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = PB * world * args.steps / float(e2e_s.item())
    # same, with the 8-bit images as host input (what open_img produces): normalisation on the device, a quarter of the H2D bytes
    host_u8 = dict(host_inputs, the_input=x_u8)
    for _ in range(2):
        model.train_on_batch(host_u8)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.train_on_batch(host_u8)
    barrier()
    e2e_u8_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_u8_s, op=dist.ReduceOp.MAX)
    e2e_u8_val = PB * world * args.steps / float(e2e_u8_s.item())
    clocks = sampler.stop() if sampler else None
    h2d = x.nbytes + lab.nbytes + L.nbytes + il.nbytes
    d2h = 4 * PB + 4                                         # per-sample losses + the CTC feasibility status (CRNNModel.train_on_batch)

    # ---- BASELINE configs[4] beside the weak-scaling line: GLOBAL batch 512 fixed, 512 / N images per GPU (per-replica BatchNorm over
    # that local batch), device-timed like `value`.  N = 8 is the main measurement itself (64 per GPU).
    strong = None
    if world > 1 and getattr(args, "scaling", "weak") == "weak" and GLOBAL_STRONG % world == 0:
        sb = GLOBAL_STRONG // world
        if sb == PB:
            strong = {"global_batch": GLOBAL_STRONG, "per_gpu_batch": sb, "value": value, "unit": "images/s", "ms_per_step": ms / args.steps, "note": "same measurement as the weak-scaling line at this N"}
        else:
            ms_model = cb.CRNN(V, MAXLEN, (IMGH, IMGW, 1), 128, args.cell == "gru", 256, max_batch=sb, seed=1234).get_model()
            ms_model.compile(optimizer=cb.Adam(lr=1e-4, beta_1=0.5, beta_2=0.999, epsilon=1e-7, clipnorm=5.0))
            dist.broadcast(ms_model.tensor("arena/params"), src=0)
            ms_model.enable_native_dp()
            sx, slab, sL, sil = synth_batch(sb, 1000 + rank)
            sargs = [torch.tensor(a_, device=dev) for a_ in (sx, slab, sL, sil)]

            def step_strong():
                step_no[0] += 1
                ms_model.train_fwd_bwd_device(*sargs, dropout_seed=seed_base + step_no[0])
                ms_model.optimizer_step(ms_model.allreduce_grads())
            for _ in range(3):
                step_strong()
            ns_ = max(5, args.steps // 2)
            sms = timed(step_strong, ns_)
            strong = {"global_batch": GLOBAL_STRONG, "per_gpu_batch": sb, "value": GLOBAL_STRONG * ns_ / (sms / 1e3), "unit": "images/s", "ms_per_step": sms / ns_, "steps": ns_}
            del ms_model, sargs
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-stage CUDA-event profile of the same step (rank 0), dominant kernel -> roofline
    import ctypes
    ns = lib.crnn_profile_num_stages()
    cb._lib.check(lib.crnn_profile_enable(model.handle, 1))
    nprof = 3
    if world == 1:
        for _ in range(nprof):
            step_device()
    else:   # the other ranks are done: profile forward/backward only, without the in-step gradient exchange (nobody would answer it)
        if dp_mode == "fused":
            cb._lib.check(lib.crnn_set_dp_fused(model.handle, 0))
        for _ in range(nprof):
            model.train_fwd_bwd_device(xd, labd, Ld, ild, dropout_seed=seed_base)
    nf = lib.crnn_profile_num_families()
    msv = (ctypes.c_double * ns)(); wk = (ctypes.c_double * ns)(); ln = (ctypes.c_longlong * ns)()
    fms = (ctypes.c_double * (ns * nf))(); fwk = (ctypes.c_double * (ns * nf))(); fln = (ctypes.c_longlong * (ns * nf))()
    cb._lib.check(lib.crnn_profile_report2(model.handle, msv, wk, ln, fms, fwk, fln))
    cb._lib.check(lib.crnn_profile_enable(model.handle, 0))
    stages = []
    for i in range(ns):
        nm = lib.crnn_profile_stage_name(i).decode()
        stages.append({"stage": nm, "ms_per_step": msv[i] / nprof, "launches_per_step": ln[i] / nprof, "work_per_step": wk[i] / nprof})
    tot = sum(s["ms_per_step"] for s in stages) or 1.0
    for s in stages:
        s["share"] = s["ms_per_step"] / tot
    stages.sort(key=lambda s: -s["ms_per_step"])
    # ---- per KERNEL: a tracked kernel (family) is summed over all the stages it serves -- xw_gemm_tc_v2_kernel runs the pointwise forward,
    # the pointwise dX and most head GEMMs -- so the roofline below is the dominant KERNEL's, not a stage's
    kernels = {}
    for i in range(ns):
        snm = lib.crnn_profile_stage_name(i).decode()
        for f in range(nf):
            k = i * nf + f
            if fms[k] <= 0:
                continue
            name = lib.crnn_profile_family_name(f).decode() if f > 0 else "stage:" + snm
            e = kernels.setdefault(name, {"kernel": name, "ms_per_step": 0.0, "work_per_step": 0.0, "launches_per_step": 0.0, "stages": []})
            e["ms_per_step"] += fms[k] / nprof; e["work_per_step"] += fwk[k] / nprof; e["launches_per_step"] += fln[k] / nprof
            e["stages"].append(snm)
    klist = sorted(kernels.values(), key=lambda e: -e["ms_per_step"])
    for e in klist:
        e["share"] = e["ms_per_step"] / tot
        e["work_unit"] = "flop" if ("gemm" in e["kernel"]) else "byte"
    pk = peaks()
    top = next(e for e in klist if not (e["kernel"].startswith("stage:gemm")))      # (the small exact-fp32 SIMT GEMMs are never the dominant kernel)
    is_gemm = top["work_unit"] == "flop"
    per_launch_work = top["work_per_step"] / max(top["launches_per_step"], 1.0)
    per_launch_s = top["ms_per_step"] / 1e3 / max(top["launches_per_step"], 1.0)
    if is_gemm:
        ach = per_launch_work / per_launch_s / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tensor_sustained"]}
    else:
        ach = per_launch_work / per_launch_s / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    roof.update({"traffic": profiled_traffic(top["kernel"]),
                 "kernel": "%s (%d launches/step in stages %s, %.3f ms/step, %.1f%% of the step; average launch %.1f us, %.3g %s of algorithmic work)" %
                 (top["kernel"], round(top["launches_per_step"]), "+".join(sorted(set(top["stages"]))), top["ms_per_step"], 100 * top["share"], per_launch_s * 1e6, per_launch_work, top["work_unit"]),
                 "peak_source": "%s (MEASURED_PEAKS.json: %s)" % (pk["src"], "bf16_tflops_sustained, kernel timed inside the step" if is_gemm else "hbm_gbs"),
                 "note": ("fp32-faithful tcgen05 GEMM: two tensor-core products (tf32 main term + bf16 cross terms) per mathematical MAC, achieved counts 2*M*N*K once; "
                          "against the bf16 dense peak the ceiling of this number format is 0.43 (tf32 runs at half the bf16 rate, plus the cross-term UMMA)") if is_gemm else ""})
    # ---- step-level figures north_star names: conv-stack arithmetic against the tensor peak, the step's DRAM bytes against the fused floor
    step_s = ms / args.steps / 1e3
    cs_flops = conv_stack_flops_per_image() * PB
    step_level = {"conv_stack_flop_per_step": cs_flops, "conv_stack_tflops": cs_flops / step_s / 1e12,
                  "conv_stack_frac_of_tensor_peak": cs_flops / step_s / 1e12 / pk["tensor_sustained"],
                  "algorithmic_fused_bytes_per_step": ALGORITHMIC_FUSED_BYTES_PER_IMAGE * PB,
                  "algorithmic_gbs": ALGORITHMIC_FUSED_BYTES_PER_IMAGE * PB / step_s / 1e9,
                  "algorithmic_frac_of_hbm_peak": ALGORITHMIC_FUSED_BYTES_PER_IMAGE * PB / step_s / 1e9 / pk["hbm"]}
    sd = step_dram_bytes()
    if sd and PB == BATCH:
        step_level.update({"measured_dram_bytes_per_step": sd["dram_bytes_per_step"], "measured_over_algorithmic": sd["dram_bytes_per_step"] / (ALGORITHMIC_FUSED_BYTES_PER_IMAGE * PB),
                           "measured_dram_gbs": sd["dram_bytes_per_step"] / step_s / 1e9, "measured_frac_of_hbm_peak": sd["dram_bytes_per_step"] / step_s / 1e9 / pk["hbm"],
                           "measured_source": "profiles/r02_step_dram.json (%s)" % sd.get("how", "ncu")})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"stages": stages, "kernels": klist, "ms_per_step_events": tot}, open(os.path.join(ROOT, "gpurun_out", "bench_stages.json"), "w"), indent=1)
    print("[bench] per-stage (CUDA events, rank 0): " + "; ".join("%s %.2fms" % (s["stage"], s["ms_per_step"]) for s in stages[:10]), file=sys.stderr)

    # ---- secondary numbers: configs[1] forward+greedy, configs[3] beam-10 decode
    extra = {}
    try:
        def fwd_greedy():
            sm = model.forward_device(xd)
            cb.ctc_decode_device(sm, greedy=True)
        for _ in range(3):
            fwd_greedy()
        extra["fwd_greedy_images_per_s"] = PB * args.steps / (timed(fwd_greedy, args.steps) / 1e3) if world == 1 else None
    except Exception as e:  # pragma: no cover
        extra["fwd_greedy_error"] = repr(e)
    # ---- the other recurrent cell of the reference (utils.py:77-82: GRU flag off -> LSTM), same workload, device-timed
    if world == 1:
        try:
            other = "lstm" if args.cell == "gru" else "gru"
            m2 = cb.CRNN(V, MAXLEN, (IMGH, IMGW, 1), 128, other == "gru", 256, max_batch=PB, seed=1234).get_model()
            m2.compile(optimizer=cb.Adam(lr=1e-4, beta_1=0.5, beta_2=0.999, epsilon=1e-7, clipnorm=5.0))

            def step_other():
                step_no[0] += 1
                m2.train_fwd_bwd_device(xd, labd, Ld, ild, dropout_seed=seed_base + step_no[0])
                m2.optimizer_step()
            for _ in range(5):
                step_other()
            n2 = max(5, args.steps // 2)
            extra["train_step_%s_images_per_s" % other] = PB * n2 / (timed(step_other, n2) / 1e3)
            del m2
        except Exception as e:  # pragma: no cover
            extra["other_cell_error"] = repr(e)
    beam = None
    if world == 1:
        rng = np.random.default_rng(3)
        z = (rng.standard_normal((4096, 25, 96)).astype(np.float32) * 3)
        pd_ = torch.softmax(torch.tensor(z, device=dev), -1).contiguous()
        ph_pageable = pd_.cpu().numpy()
        ph_pin_t = torch.empty(pd_.shape, dtype=torch.float32).pin_memory()      # host input buffer in pinned memory (the e2e contract)
        ph_pin_t.copy_(pd_)
        ph = ph_pin_t.numpy()
        for _ in range(3):
            cb.ctc_decode_device(pd_, greedy=False, beam_width=10)
        bms = timed(lambda: cb.ctc_decode_device(pd_, greedy=False, beam_width=10), 20) / 20
        cb.ctc_decode_host(ph, greedy=False, beam_width=10)
        t0 = time.perf_counter()
        for _ in range(5):
            cb.ctc_decode_host(ph, greedy=False, beam_width=10)
        bh = (time.perf_counter() - t0) / 5
        cb.ctc_decode_host(ph_pageable, greedy=False, beam_width=10)
        t0 = time.perf_counter()
        for _ in range(3):
            cb.ctc_decode_host(ph_pageable, greedy=False, beam_width=10)
        bh_pageable = (time.perf_counter() - t0) / 3
        from oracle import ctc_oracle as O
        sub = ph[:1024]
        t0 = time.perf_counter(); O.beam(sub, beam_width=10); c1 = 1024 / (time.perf_counter() - t0)
        cores = os.cpu_count() or 1
        t0 = time.perf_counter(); O.beam_threaded(ph, cores, beam_width=10); cN = 4096 / (time.perf_counter() - t0)
        beam = {"workload": "configs[3]: beam-10 decode, (4096,25,96) softmax of N(0,1)*3 logits, merge_repeated",
                "lines_per_s_device": 4096 / (bms / 1e3), "lines_per_s_e2e_host_buffers": 4096 / bh, "lines_per_s_e2e_pageable_host_buffers": 4096 / bh_pageable,
                "e2e_note": "crnn_ctc_beam_host (C ABI): H2D of the 39.3 MB softmax from pinned host memory + decode + D2H of labels, wall clock", "ms_per_call": bms,
                "cpu_oracle_1thread_lines_per_s": c1, "cpu_oracle_all_threads_lines_per_s": cN, "cpu_threads": cores,
                "speedup_e2e_vs_cpu_1thread": (4096 / bh) / c1, "speedup_device_vs_cpu_1thread": (4096 / (bms / 1e3)) / c1,
                "speedup_e2e_vs_cpu_all_threads": (4096 / bh) / cN, "speedup_device_vs_cpu_all_threads": (4096 / (bms / 1e3)) / cN,
                "hbm_frac": (4096 * 25 * 96 * 4 / (bms / 1e3) / 1e9) / pk["hbm"]}
    # ---- CPU baseline: the oracle port, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = pick_cpu_threads(args.cell)
        sample, nrep = BATCH, 3
        st = oracle_train_step_fn(args.cell, sample, cores)
        st()                                                    # warm-up (allocator, oneDNN primitive caches)
        t0 = time.perf_counter()
        for _ in range(nrep):
            st()
        dt = time.perf_counter() - t0
        cpu = {"value": sample * nrep / dt, "unit": "images/s", "cores": cores, "kind": "port", "host_cores": os.cpu_count(),
               "sample": "%d oracle train steps (PyTorch-CPU fp32 restatement + C CTC + Adam) on the full batch of %d images, %.1f s of CPU work, best of 8/16/32 threads" % (nrep, sample, dt)}

    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": getattr(args, "scaling", "weak"), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "CRNNModel.train_on_batch(host numpy dict) == Keras train_on_batch (train.py:201-209); single process: one C-ABI call, crnn_train_on_batch_host (pageable numpy -> pinned staging -> H2D -> step -> optimiser -> D2H of losses + status)",
                    "uint8_input": {"value": e2e_u8_val, "h2d_bytes_per_step": int(x_u8.nbytes + lab.nbytes + L.nbytes + il.nbytes),
                                    "note": "same call with 'the_input' as the raw 8-bit images; utils.py:415 norm() runs on the device (crnn_normalize_u8)"}},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps, "dp_mode": dp_mode,
            "roofline": roof, "step_level": step_level, "configs4_strong": strong, "cpu_baseline": cpu, "beam_decode": beam, "extra": extra,
            "kernels_top": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.items()} for e in klist[:8]],
            "stages_top": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items()} for s in stages[:8]]}
    sys.stdout.flush()
    _emit(line, out_fd)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cell", default="gru", choices=["gru", "lstm"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 64 images per GPU (default); strong: BASELINE configs[4], global batch 512 = 512/N images per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
