#!/usr/bin/env python
"""Benchmark of the V-Net training hot path (BASELINE.json metric: patches/sec, 128^3, 1ch -> 2cls).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one optimiser step (forward + weighted Dice + backward + Adam [+ gradient all-reduce])
over one batch of synthetic patches.  N>1 is launched by torchrun (one rank per GPU, NCCL); whole
patches are sharded over ranks (weak scaling: per-GPU batch fixed).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_PATCH_128 = 3371.7   # BASELINE.md §2 (fprop + dgrad + wgrad, 128^3, M=1, K=2)
METRIC = "patches/sec (128^3, 1ch->2cls) training step"
DROPOUT = 0.01                       # configs/*.json Networks.Dropout

# BASELINE.json configs[i]; [1] is the one the metric is quoted on (the default), the others are parity-test
# configurations that can be timed on request with --config
PRESETS = {
    2: dict(patch=128, batch=2, M=1, K=2, precision="bf16x3", attention=False, weights=(0.1, 1.0), train_gflop=3371.7,
            name="BASELINE configs[1]: 128^3, 1 modality, 2 classes, batch 2, fp32-grade"),
    3: dict(patch=128, batch=2, M=4, K=4, precision="bf16", attention=False, weights=(0.01, 0.1, 0.5, 1.0), train_gflop=3439.0,
            name="BASELINE configs[2]: 128^3, 4 modalities, 4 classes (BraTS-shaped), batch 2, bf16"),
    4: dict(patch=128, batch=2, M=1, K=2, precision="bf16", attention=False, weights=(0.1, 1.0), train_gflop=3371.7,
            name="BASELINE configs[3]: 128^3, 1 modality, 2 classes, 2 per GPU, bf16 (run with --gpus 8)"),
    5: dict(patch=192, batch=1, M=2, K=3, precision="bf16", attention=True, weights=(0.01, 0.1, 1.0), train_gflop=59600.0,
            name="BASELINE configs[4]: 192^3, 2 modalities, 3 classes, attention.py gating, 1 per GPU, bf16 (run with --gpus 8)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(PRESETS), help="BASELINE.json configs[i-1]")
    ap.add_argument("--patch", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="patches per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("VNB_BENCH_PRECISION"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-staged", action="store_true", help="accepted for compatibility: the e2e leg is the staged loop by default")
    ap.add_argument("--per-layer", default=None, metavar="FILE",
                    help="also write the per-layer roofline table (every 5^3 / 3^3 convolution launch of one profiled step: "
                         "layer, pass, ms, algorithmic TFLOP/s, fraction of the measured peak) as JSON to FILE")
    ap.add_argument("--sync-bn", action="store_true", help="N>1: batch-norm statistics of the global batch (off: local statistics)")
    args = ap.parse_args()
    pre = PRESETS[args.config]
    args.preset = pre
    args.patch = args.patch or pre["patch"]
    args.batch = args.batch or pre["batch"]
    args.precision = args.precision or pre["precision"]
    return args


def train_gflop_per_patch(patch: int, preset=None) -> float:
    if preset is not None:
        return preset["train_gflop"] * (patch / float(preset["patch"])) ** 3
    return TRAIN_GFLOP_PER_PATCH_128 * (patch / 128.0) ** 3


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback"}


def roofline_block(precision, peaks, fd_ms, fd_n, fd_fl, wg_ms, wg_n, wg_fl, prof_steps, ms_per_step, step_tflops):
    """Roofline of the dominant kernel class (largest share of the step).  achieved = algorithmic FLOPs
    (2*125*Cin*Cout*voxels per launch; the three MMA passes of bf16x3 are NOT counted) / CUDA-event time of the
    launches on the engine stream; peak = measured cuBLAS bf16 (sustained); traffic = DRAM bytes per launch of
    that class from the committed ncu --set full capture (profiles/r01_traffic.json), else null."""
    classes = {
        "conv5_tc_kernel (5^3 fprop + dgrad)": (fd_ms, fd_n, fd_fl),
        "wgrad5_tc_kernel (5^3 filter gradient)": (wg_ms, wg_n, wg_fl),
    }
    dom = max(classes, key=lambda k: classes[k][0])
    ms, n, fl = classes[dom]
    achieved = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    traffic, traffic_src = None, None
    for fn in ("r02_traffic.json", "r01_traffic.json"):   # committed ncu --set full capture: ONE launch of that class
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                traffic = json.load(f).get(precision, {}).get(dom.split(" ")[0])
            if traffic is not None:
                traffic_src = "profiles/" + fn
                break
        except Exception:
            pass
    return {
        "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tflops"], "traffic": traffic,
        "traffic_source": (traffic_src + ": dram__bytes_read.sum + dram__bytes_write.sum of a single level-1 launch of this "
                           "kernel in a committed ncu --set full capture (not measured in this run, not a class average)")
        if traffic_src else None,
        "launches_per_step": n / max(prof_steps, 1), "avg_launch_ms": ms / max(n, 1),
        "share_of_step": ms / max(prof_steps, 1) / ms_per_step,
        "peak_source": peaks["source"] + " cuBLAS bf16 sustained (MEASURED_PEAKS.json)",
        "mma_passes": 3 if precision == "bf16x3" else 1,
        "fprop_dgrad_tflops": fd_fl / (fd_ms * 1e-3) / 1e12 if fd_ms > 0 else 0.0,
        "wgrad_tflops": wg_fl / (wg_ms * 1e-3) / 1e12 if wg_ms > 0 else 0.0,
        "all_conv5_tflops": (fd_fl + wg_fl) / ((fd_ms + wg_ms) * 1e-3) / 1e12 if fd_ms + wg_ms > 0 else 0.0,
        "conv_share_of_step": (fd_ms + wg_ms) / max(prof_steps, 1) / ms_per_step,
        "whole_step_tflops_per_gpu": step_tflops,
        "method": "CUDA events around every 5^3 convolution launch on the engine stream, %d profiled steps after the "
                  "timed region (filter gradients serialised on the main stream while profiling, so every launch is "
                  "timed alone; in the timed steps they overlap the following units on a side stream)" % prof_steps,
    }


def write_per_layer_table(path, launches, prof_steps, peaks, precision):
    """SURVEY H1's per-layer roofline table: the profiled launches of each (layer, pass) averaged over the profiled
    steps, with the algorithmic rate (one MMA pass counted, whatever the precision mode issues) over the measured peak."""
    rows = {}
    for label, cls, ms, flops in launches:
        r = rows.setdefault(label, {"layer": label, "class": "wgrad" if cls == 1 else "fprop/dgrad", "ms": 0.0, "gflop": 0.0, "n": 0,
                                    "each": []})
        r["each"].append(ms)
        r["ms"] += ms
        r["gflop"] += flops / 1e9
        r["n"] += 1
    table = []
    for r in rows.values():
        ms = float(np.median(r["each"]))     # median over the profiled steps: one launch delayed by the host does not count
        tflops = (r["gflop"] / r["n"]) / ms if ms > 0 else 0.0     # GFLOP / ms = TFLOP/s
        table.append({"layer": r["layer"], "class": r["class"], "launches_per_step": r["n"] / max(prof_steps, 1),
                      "ms_per_launch": ms, "ms_each": r["each"], "gflop_per_launch": r["gflop"] / r["n"],
                      "tflops": tflops, "frac_of_peak": tflops / peaks["tflops"]})
    with open(path, "w") as f:
        json.dump({"precision": precision, "peak_tflops": peaks["tflops"], "peak_source": peaks["source"],
                   "profiled_steps": prof_steps, "rows": table}, f, indent=1)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region: NVML every 20 ms when pynvml is importable,
    else one `nvidia-smi` query per sample."""

    NVML_REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                    0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML indexes the physical ones
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for bit, nm in self.NVML_REASONS.items():
            if mask & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for nm, v in zip(names, f[2:6]):
            if v.lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self._nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sm_mhz_min": float(min(self.samples)) if self.samples else None,
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# CPU reference leg (oracle port of the reference's TF1 graph; TensorFlow itself cannot run here)
# --------------------------------------------------------------------------------------------------
def _oracle_setup():
    import torch
    from oracle import ref_vnet as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return R, cores


def cpu_reference_steps(sample_patch: int, batch: int, steps: int, warmup: int, budget_s: float = 240.0):
    """`warmup` untimed + `steps` timed optimiser steps of the oracle (oracle/ref_vnet.py: the TF1-semantics CPU
    restatement of networks.py / model.py on PyTorch/oneDNN, all host threads) on `batch` synthetic patches of
    `sample_patch`^3 each.  When the first warm-up step shows that warmup + steps would overrun `budget_s`, the sample
    shrinks to 64^3 patches (returned with its voxel ratio, which the caller scales by).  Returns
    (seconds per timed step, patch extent used, cores)."""
    R, cores = _oracle_setup()
    from vnet_tensorflow_b200.synthetic import synth_batch
    spec = R.VNetSpec(num_classes=2, in_channels=1)
    state = R.TrainState(params=R.init_params(spec, 42))
    # thread pool / allocator spin-up on a 32^3 patch (not one of the W warm-up steps)
    R.train_step(state, *synth_batch(0, 1, min(32, sample_patch), 1, 2), spec, "weighted_sorensen", (0.1, 1.0))
    used = sample_patch
    times = []
    i = 0
    while i < warmup + steps:
        img, lab = synth_batch(i, batch, used, 1, 2)
        t0 = time.perf_counter()
        R.train_step(state, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
        dt = time.perf_counter() - t0
        if i == 0 and used > 64 and dt * (warmup + steps) > budget_s:
            used = 64          # too slow for the requested K + W on these cores: bounded sample of 64^3 patches
            continue
        if i >= warmup:
            times.append(dt)
        i += 1
    return times, used, cores


def _sample_text(steps: int, warmup: int, sample: int, patch: int, batch: int, cores: int, seconds: float) -> str:
    txt = ("%d timed + %d warm-up optimiser steps, each on %d synthetic %d^3 patch%s (fwd + weighted Dice + bwd + Adam; "
           "oracle/ref_vnet.py = TF1-semantics CPU restatement on PyTorch/oneDNN, not TensorFlow; %d threads), "
           "%.2f s per step" % (steps, warmup, batch, sample, "" if batch == 1 else "es", cores, seconds))
    if sample != patch:
        txt += "; patches/sec scaled by (%d/%d)^3 voxels to %d^3 patches" % (sample, patch, patch)
    return txt


def workload_config(args, world: int, dropout: float):
    """The `config` object of the JSON line -- identical for both arms (the reference arm describes how it sampled
    this workload in cpu_baseline.sample)."""
    pre = args.preset
    P, B, M, K, att = args.patch, args.batch, pre["M"], pre["K"], pre["attention"]
    return {"workload": "V-Net 3D train step (fwd + weighted Dice%s + bwd + Adam%s), %d^3 patch, %d modalit%s, %d classes, "
                        "batch %d per GPU (%s)" % (" + attention gating / attention loss" if att else "",
                                                   " + hand-rolled all-reduce" if world > 1 else "", P, M,
                                                   "y" if M == 1 else "ies", K, B, pre["name"]),
            "patch": P, "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "batch_norm": "synchronised" if (args.sync_bn and world > 1) else "local statistics",
            "precision": args.precision, "dropout": dropout, "attention": bool(att),
            "l2": "per-step working set (activations %.1f GB) >> 126 MB L2, no explicit flush" % (0.7 * 3 * B * (P / 128) ** 3)}


def run_reference(args, rank: int, world: int):
    """Reference arm: the reference's CPU path for this workload = the oracle port (TensorFlow 1.15 cannot be installed
    here), on all host threads.  Runs exactly W warm-up + K timed steps; each step is a bounded sample of
    the workload: ONE patch of the batch (same FLOPs per patch; its batch norm sees one patch instead of two)."""
    if rank != 0:
        return
    times, used, cores = cpu_reference_steps(min(args.patch, 128), 1, args.steps, args.warmup)
    t = float(np.mean(times))
    scale = (used / args.patch) ** 3
    value = scale / t
    sample = _sample_text(len(times), args.warmup, used, args.patch, 1, cores, t)
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "patches/sec", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1), DROPOUT),
        "cpu_baseline": {"value": value, "unit": "patches/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "patches/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# Parity of the CUDA engine against the oracle on the benchmarked configuration ("parity printed
# next to the timings").  The oracle is the checker here, never the thing measured.
# --------------------------------------------------------------------------------------------------
def parity_and_cpu_baseline(eng, patch: int, batch: int, precision: str, timed_steps: int = 2):
    """Runs the oracle's optimiser step on the workload's own batch (`batch` x `patch`^3, seed-42 weights) -- timed,
    which is the cpu_baseline -- and the engine's forward / loss / backward on the identical tensors.
    Returns (parity dict, cpu_baseline dict)."""
    R, cores = _oracle_setup()
    from vnet_tensorflow_b200.synthetic import synth_batch
    spec = R.VNetSpec(num_classes=2, in_channels=1)
    weights = (0.1, 1.0)
    params = R.init_params(spec, 42)
    state = R.TrainState(params={k: v.copy() for k, v in params.items()})
    R.train_step(state, *synth_batch(0, 1, min(32, patch), 1, 2), spec, "weighted_sorensen", weights)   # spin-up
    state = R.TrainState(params={k: v.copy() for k, v in params.items()})
    img, lab = synth_batch(0, batch, patch, 1, 2)
    times = []
    t0 = time.perf_counter()
    loss_o, logits_o, grads_o = R.train_step(state, img, lab, spec, "weighted_sorensen", weights)
    times.append(time.perf_counter() - t0)
    for i in range(1, timed_steps):
        im2, lb2 = synth_batch(i, batch, patch, 1, 2)
        t0 = time.perf_counter()
        R.train_step(state, im2, lb2, spec, "weighted_sorensen", weights)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    cpu = {"value": batch / t, "unit": "patches/sec", "cores": cores, "kind": "port",
           "sample": _sample_text(len(times), 0, patch, patch, batch, cores, t) + " (after a 32^3 spin-up step)"}
    parity = engine_parity(eng, params, img, lab, loss_o, logits_o, grads_o, precision)
    return parity, cpu


def engine_parity(eng, params, img, lab, loss_o, logits_o, grads_o, precision):
    """Engine (through the C ABI) vs oracle outputs on identical inputs: north_star's bars are logits within 1e-3
    relative, bit-exact argmax label volume and hard Dice."""
    eng.set_params(params)
    logits, _, argmax = eng.forward(img, want_softmax=False)
    loss = eng.forward_backward(img, lab, dropout_rate=0.0)
    ref = np.asarray(logits_o, np.float32)
    scale = float(np.abs(ref).max())
    err = float(np.abs(logits - ref).max() / scale)
    ref_arg = np.argmax(ref, -1)          # first maximum, as tf.argmax (model.py:568)
    flips = argmax != ref_arg
    mism = int(flips.sum())
    # Two fp32 evaluations of one graph that differ in summation order cannot agree on argmax where the two top logits
    # are closer than their own rounding error: a mismatch counts against parity only outside that band
    band = 2.0 * float(np.abs(logits - ref).max())
    top = np.sort(ref, -1)
    outside = int((flips & ((top[..., -1] - top[..., -2]) > band)).sum())
    K = ref.shape[-1]
    hard = {}
    dice_equal = True
    dice_diff = 0
    for c in range(1, K):
        a = [int(((x == c) & (lab == c)).sum()) for x in (argmax, ref_arg)]          # TP
        b = [int(((x == c) & (lab != c)).sum()) for x in (argmax, ref_arg)]          # FP
        d = [int(((x != c) & (lab == c)).sum()) for x in (argmax, ref_arg)]          # FN
        hard["class_%d" % c] = {"tp": a[0], "fp": b[0], "fn": d[0], "oracle_tp": a[1], "oracle_fp": b[1], "oracle_fn": d[1]}
        dice_equal = dice_equal and a[0] == a[1] and b[0] == b[1] and d[0] == d[1]
        dice_diff = max(dice_diff, abs(a[0] - a[1]), abs(b[0] - b[1]), abs(d[0] - d[1]))
    g = eng.get_grads()
    worst, worst_name = 0.0, ""
    top = max(float(np.sqrt((np.asarray(r, np.float64) ** 2).sum())) for r in grads_o.values())
    for k, v in g.items():
        r = np.asarray(grads_o[k], np.float64)
        den = float(np.sqrt((r ** 2).sum()))
        if k.endswith("/biases") or den < 1e-5 * top:
            continue   # analytically zero gradients (conv biases and the dead / re-normalised batch norms):
                       # autodiff leaves rounding noise there, the engine writes exact zeros
        e = float(np.sqrt(((v.astype(np.float64) - r) ** 2).sum()) / den)
        if e > worst:
            worst, worst_name = e, k
    w1 = "vnet/encoder/level_1/conv_1/weights"
    r1 = np.asarray(grads_o[w1], np.float64)
    e1 = float(np.sqrt(((g[w1].astype(np.float64) - r1) ** 2).sum()) / max(float(np.sqrt((r1 ** 2).sum())), 1e-30))
    return {"against": "oracle/ref_vnet.py (fp32, CPU) on identical synthetic tensors and seed-42 weights, dropout 0",
            "shape": list(img.shape), "precision": precision,
            "logits_max_rel_err": err, "logits_tol": 1e-3, "argmax_mismatches": mism,
            "argmax_mismatches_outside_rounding_band": outside, "rounding_band_abs": band, "voxels": int(ref_arg.size),
            "hard_dice_counts_equal": bool(dice_equal), "hard_dice_max_count_diff": int(dice_diff), "hard_dice": hard,
            "loss": float(loss), "oracle_loss": float(loss_o), "abs_loss_diff": abs(float(loss) - float(loss_o)),
            "grad_rel_l2_worst": worst, "grad_rel_l2_worst_tensor": worst_name, "grad_rel_l2_first_conv": e1,
            "pass": bool(err <= 1e-3 and outside == 0 and dice_diff <= mism)}


# --------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from vnet_tensorflow_b200.init import initialize
    from vnet_tensorflow_b200.engine import VNetEngine
    from vnet_tensorflow_b200.synthetic import synth_patch

    # VNB_BENCH_DRYRUN=1 (tests/test_host_mirror.py): the flow of this script against the CPU emulation of the kernels
    # (VNB_LIBRARY points at tests/emul's build) -- test infrastructure; without it a missing GPU is an error here
    dry = os.environ.get("VNB_BENCH_DRYRUN") == "1"
    if not dry:
        torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    P, B = args.patch, args.batch
    pre = args.preset
    M, K, att = pre["M"], pre["K"], pre["attention"]
    eng = VNetEngine(num_classes=K, in_channels=M, patch_shape=(P, P, P), max_batch=B, precision=args.precision,
                     loss="weighted_sorensen", loss_weights=pre["weights"], optimizer="Adam", learning_rate=1e-2,
                     decay_factor=0.99, decay_steps=100.0, device=local_rank,
                     attention=att, attention_loss="l2" if att else None)
    initialize(eng, 42)
    if world > 1:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])
        if args.sync_bn:
            eng.comm_sync_bn(True)

    def barrier():
        if world > 1:
            dist.barrier()
        if not dry:
            torch.cuda.synchronize()
        eng.sync()

    # synthetic batches in pinned host memory (one per step so the e2e leg really moves new data)
    nb = min(args.steps + args.warmup, 4)
    pinned = []
    for i in range(nb):
        samples = [synth_patch(1234 + 1000 * rank + i + 100000 * b, P, M, K) for b in range(B)]
        tens = [torch.from_numpy(np.stack([smp[j] for smp in samples], 0)) for j in range(3)]
        if not dry:
            tens = [x.pin_memory() for x in tens]
        pinned.append((tens, None, tens[0].numpy(), tens[1].numpy(), tens[2].numpy()))
    dropout = DROPOUT

    # ---- leg 1: device-resident inputs ("value") ------------------------------------------------
    eng.upload_batch(pinned[0][2], pinned[0][3])
    if att:
        eng.set_distmap(pinned[0][4])
    for i in range(args.warmup):
        eng.train_step_resident(B, dropout, seed=i)
    barrier()
    l0 = eng.gpu_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.event_record(0)
    for i in range(args.steps):
        eng.train_step_resident(B, dropout, seed=100 + i)
    eng.event_record(1)
    barrier()
    ms_dev = eng.event_elapsed_ms()
    if dry:
        ms_dev = max(ms_dev, 1e-3)   # the emulation has no device clock
    clocks = sampler.stop()
    launches = eng.gpu_launches() - l0

    # ---- leg 2: end to end through the public API with host buffers ("e2e") ---------------------
    # The call a user makes is model.train's loop (vnet_tensorflow_b200/model.py): vnb_stage_batch of batch i+1 from
    # page-locked host memory on the copy stream while vnb_train_step_staged of batch i runs, and the step's loss read back
    # every step.  Every timed step therefore includes its own host->device copy (33.5 MB) and a device->host read of the
    # loss; the copy overlaps the previous step's compute instead of preceding its own.  The plain synchronous
    # vnb_train_step (copy, then compute, then read) is timed as well and reported as e2e_unstaged.
    def plain_loop(steps, seed0):
        last = None
        for i in range(steps):
            if att:
                eng.set_distmap(pinned[i % nb][4])
            last = eng.train_step(pinned[i % nb][2], pinned[i % nb][3], dropout, seed=seed0 + i, want_loss=True)
        return last

    def staged_loop(steps, seed0):
        last = None
        eng.stage_batch(pinned[0][2], pinned[0][3])
        for i in range(steps):
            eng.train_step_staged(dropout, seed=seed0 + i, want_loss=False)
            if i + 1 < steps:
                eng.stage_batch(pinned[(i + 1) % nb][2], pinned[(i + 1) % nb][3])
            last = eng.last_loss()          # the step's loss is read every step
        return last

    def timed(loop, seed0):
        loop(min(args.warmup, 2), seed0)
        barrier()
        t0 = time.perf_counter()
        eng.event_record(0)
        last = loop(args.steps, seed0 + 100)
        eng.event_record(1)
        barrier()
        # host-side copies / read-backs are part of the end-to-end time: the larger of the device and the wall clock
        return max(eng.event_elapsed_ms(), (time.perf_counter() - t0) * 1e3), last

    ms_plain, loss = timed(plain_loop, 200)
    ms_staged = None
    if not att:   # the staged entry points take images + labels; the attention path's distance map goes through the plain call
        ms_staged, loss = timed(staged_loop, 500)
    ms_e2e = ms_staged if ms_staged is not None else ms_plain

    # ---- dominant-kernel roofline: CUDA events around every 5^3 convolution launch ---------------
    eng.profile_enable(True)
    prof_steps = min(args.steps, 3)
    for i in range(prof_steps):
        eng.train_step_resident(B, dropout, seed=300 + i)
    launches_prof = eng.profile_launches()

    def class_totals(cls):
        """(ms, launches, FLOPs) of a kernel class over the profiled steps, every (layer, pass) entering with the median
        of its launches (a launch delayed by the host while every kernel is timed alone does not distort the rate)."""
        per = {}
        for label, c, ms, fl in launches_prof:
            if c == cls:
                per.setdefault(label, []).append((ms, fl))
        ms_t = sum(float(np.median([m for m, _ in v])) * len(v) for v in per.values())
        return ms_t, sum(len(v) for v in per.values()), sum(f for v in per.values() for _, f in v)

    conv_ms, conv_n, conv_fl = class_totals(0)
    wg_ms, wg_n, wg_fl = class_totals(1)
    if args.per_layer and rank == 0:
        write_per_layer_table(args.per_layer, launches_prof, prof_steps, measured_peaks(), args.precision)
    eng.profile_enable(False)

    if world > 1:   # max over ranks of the device-timed regions
        t = torch.tensor([ms_dev, ms_e2e, ms_plain], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_plain = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        peaks = measured_peaks()
        patches = B * world * args.steps
        value = patches / (ms_dev / 1e3)
        e2e_val = patches / (ms_e2e / 1e3)
        img_bytes = B * P ** 3 * 4 * M
        lab_bytes = B * P ** 3 * 4 * (2 if att else 1)   # labels (+ distance map)
        step_tflops = value * train_gflop_per_patch(P, pre) / 1e3 / world
        line = {
            "metric": METRIC, "value": value, "unit": "patches/sec",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "fp32", "bf16x3": "bf16x3 (fp32-grade split, fp32 accumulate)", "bf16": "bf16"}[args.precision],
            "data": "synthetic",
            "config": workload_config(args, world, dropout),
            "e2e": {"value": e2e_val, "unit": "patches/sec", "h2d_bytes_per_step": img_bytes + lab_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "how": ("model.train's loop: vnb_stage_batch (pinned host -> device, copy stream) of batch i+1 under "
                            "vnb_train_step_staged of batch i, loss read back every step") if ms_staged is not None else
                           "vnb_train_step with pinned host buffers (copy, compute, loss read-back in sequence)"},
            "e2e_unstaged": {"value": patches / (ms_plain / 1e3), "unit": "patches/sec", "ms_per_step": ms_plain / args.steps,
                             "how": "vnb_train_step: host -> device copy, step and loss read-back in sequence"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline_block(args.precision, peaks, conv_ms, conv_n, conv_fl, wg_ms, wg_n, wg_fl, prof_steps,
                                       ms_dev / args.steps, step_tflops),
            "final_loss": loss,
        }
        if world == 1 and not args.no_cpu_baseline and args.config == 2:
            # oracle on the workload's own batch, timed (cpu_baseline), and the engine held against it (parity)
            line["parity"], line["cpu_baseline"] = parity_and_cpu_baseline(eng, P, B, args.precision)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on some hosts) out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
