#!/usr/bin/env python
"""Benchmark of the U-ResNet training hot path on B200 (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 2d|3d] [--dtype mix16|f32|bf16]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement on the host cores

A step = one training step of the reference's hot path (lib/ssnet_trainval.py:164-191 with
NUM_MINIBATCHES=1): zero the accumulators, forward, weighted softmax-CE + metrics, backward
into the accumulators, (N>1: bucketed all-reduce overlapped with backward,) Adam apply.
Workloads (BASELINE.json configs[1] / configs[3]; weak scaling: per-GPU batch fixed, so N=8 is
configs[2] / configs[4]):  2d = U-ResNet F=16 S=5, 512x512, batch 32 per GPU;
3d = 192^3, batch 4 per GPU.  Synthetic LArTPC-like data, Xavier-initialised weights.

`value`  : images/s, inputs resident in HBM, CUDA events, max over ranks.
`e2e`    : same metric through the reference-facing API (uresnet.accum_gradients with HOST numpy
           arrays: pinned H2D copy of data/label/weight and D2H read of loss/accuracies inside
           the timed region).
`roofline`: the dominant kernel family (conv forward/dgrad/wgrad launches), algorithmic FLOPs
           of those launches / their CUDA-event time, against MEASURED_PEAKS.json.
           `roofline.top_launch`: the single slowest conv launch group of the step (layer, phase)
           against the roofline that bounds it (HBM or tensor by its arithmetic intensity), from
           the same live CUDA-event profile; `traffic` = DRAM bytes of the committed ncu capture
           when that instance was captured (profiles/), else null.
           `roofline_hbm`: algorithmic bytes of the WHOLE step / step time against the HBM peak.
`cpu_baseline`: the CPU oracle (torch-CPU fp32 restatement of the reference graph -- the
           reference's own TF-1.x path cannot run here) on a bounded sample, rank 0 only.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "2d": dict(dims=[512, 512, 1], batch=32, name="U-ResNet 2D train, 512x512, batch 32/GPU, F=16, S=5, 3 classes, weighted CE"),
    "3d": dict(dims=[192, 192, 192, 1], batch=4, name="U-ResNet 3D train, 192^3, batch 4/GPU, F=16, S=5, 3 classes, weighted CE"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def conv_flops(layer_table):
    """Algorithmic FLOPs per step by phase: 2*|Y|*taps*cx*cy for fwd, dgrad and wgrad (SURVEY 8a
    Table T); no dgrad for conv0 (its input is data)."""
    f = {"conv_fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for i, (scope, d, tr) in enumerate(layer_table):
        fl = 2.0 * d.n * d.yd * d.yh * d.yw * d.kd * d.kh * d.kw * d.cx * d.cy
        f["conv_fwd"] += fl
        f["wgrad"] += fl
        if scope != "conv0":
            f["dgrad"] += fl
    return f


# dram__bytes_read.sum + dram__bytes_write.sum of committed `ncu --set full` captures, per launch
# (profiles/raw_*_r01e_*.summary.txt); keyed by (workload, scope suffix, phase)
NCU_TRAFFIC = {("2d", "16->16 k3 s1 @512", "conv_fwd"): 289.1e6 + 485.8e6}


def top_launch(layer_table, prof_ms, prof_steps, pk, workload=""):
    """The single conv launch group (layer, phase) with the largest measured time per step, with
    the roofline that bounds it: algorithmic FLOPs 2*|Y|*taps*cx*cy and bytes (16-bit operands in,
    fp32 raw / bf16 gradient out) over its CUDA-event time; HBM-bound when its arithmetic
    intensity is under the ridge of the measured peaks, tensor-bound otherwise."""
    phases = {0: "conv_fwd", 4: "wgrad", 5: "dgrad"}
    best = None
    for i, (scope, d, tr) in enumerate(layer_table):
        for col, name in phases.items():
            ms = float(prof_ms[i, col]) / prof_steps
            if best is None or ms > best[0]:
                best = (ms, i, name)
    ms, i, name = best
    scope, d, tr = layer_table[i]
    big = float(d.n) * d.xd * d.xh * d.xw * d.cx
    small = float(d.n) * d.yd * d.yh * d.yw * d.cy
    ein, eout = (small, big) if tr else (big, small)
    flops = 2.0 * d.n * d.yd * d.yh * d.yw * d.kd * d.kh * d.kw * d.cx * d.cy
    byts = {"conv_fwd": ein * 2 + eout * 4, "dgrad": eout * 2 + ein * 2, "wgrad": ein * 2 + eout * 2}[name]
    ridge = pk["tf_sus"] * 1e12 / (pk["hbm"] * 1e9)
    hbm_bound = flops / byts < ridge
    sec = max(ms, 1e-9) * 1e-3
    achieved = byts / sec / 1e9 if hbm_bound else flops / sec / 1e12
    peak = pk["hbm"] if hbm_bound else pk["tf_sus"]
    side = max(d.xh, d.xw)
    shape = "%d->%d k%d s%d @%d" % (((d.cy, d.cx) if tr else (d.cx, d.cy)) + (d.kh, d.s, side))
    return {"kernel": "%s %s (%s%s)" % (scope, name, shape, ", transposed" if tr else ""), "bound": "hbm" if hbm_bound else "tensor",
            "achieved": achieved, "peak": peak, "unit": "GB/s" if hbm_bound else "TFLOP/s", "frac": achieved / peak,
            "ms": ms, "algorithmic_bytes": byts, "algorithmic_flops": flops,
            "traffic": NCU_TRAFFIC.get((workload, shape, name))}


def hbm_bytes(layer_table, elt_act=2, elt_grad=2):
    """Algorithmic HBM bytes of one training step (every logical tensor counted once per
    kernel that must touch it): conv fwd (in + fp32 raw out), BN statistics (raw), BN apply (raw in,
    activation + bf16 shadow out), BN backward (two passes over dout and out/raw, d(raw) out),
    dgrad (d(raw) in, dX out), wgrad (X + d(raw) in).  Residual / concat extras are left out, so
    this is a lower bound of what the step has to move."""
    tot = 0.0
    for scope, d, tr in layer_table:
        big = float(d.n) * d.xd * d.xh * d.xw * d.cx      # X side elements
        small = float(d.n) * d.yd * d.yh * d.yw * d.cy    # Y side elements
        ein, eout = (small, big) if tr else (big, small)
        tot += ein * elt_act + eout * 4                    # conv forward
        tot += eout * 4                                    # BN statistics
        tot += eout * 4 + eout * 2 * elt_act               # BN apply (+ shadow)
        tot += 2 * (eout * elt_grad + eout * 4) + eout * elt_grad   # BN backward, 2 passes
        if scope != "conv0":
            tot += eout * elt_grad + ein * elt_grad        # dgrad
        tot += ein * elt_act + eout * elt_grad             # wgrad
    return tot


# ------------------------------------------------------------------------------------------
# reference arm: the CPU oracle on the host cores
# ------------------------------------------------------------------------------------------
def cpu_sample(workload, steps, warmup, sample_batch):
    """Times the oracle's train step (forward + autograd backward, fp32, all host threads) on a
    bounded sample of the workload: full-size images, batch `sample_batch`."""
    import torch
    from oracle import uresnet_oracle as O            # the one place bench.py executes oracle/
    from uresnet_b200 import synthetic_io as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dims = WORKLOADS[workload]["dims"]
    if workload == "3d":
        dims = [64, 64, 64, 1]                         # full 192^3 autograd does not fit/finish on CPU
    nd = len(dims) - 1
    data, label, weight = S.make_batch(sample_batch, dims, seed=1234)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(nd, seed=4321)
    shp = [sample_batch] + dims[:-1]
    args = (P, data.reshape(shp + [1]), label.reshape(shp), weight.reshape(shp))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.run(*args, dtype=torch.float32, want_grads=True)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = float(np.mean(ts))
    scale = 1.0
    sample = "%d full-size image(s) per step (%s), fwd+autograd bwd, torch-CPU fp32" % (sample_batch, "x".join(map(str, dims[:-1])))
    if workload == "3d":
        scale = (64.0 / 192.0) ** 3                    # extrapolate by voxel count
        sample += "; 64^3 crop, images/s scaled by (64/192)^3"
    return dict(value=sample_batch / sec * scale, unit="images/s", cores=cores, kind="port", sample=sample), sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    cb, sec = cpu_sample(args.workload, steps, min(args.warmup, 1), 1 if args.workload == "2d" else 1)
    line = {"impl": "reference", "metric": "train images/sec", "value": cb["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["name"], "note": "CPU restatement of the reference graph (TF-1.x reference cannot run here)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("URN_BENCH_WORKLOAD", "2d"), choices=list(WORKLOADS))
    ap.add_argument("--dtype", default=os.environ.get("URN_BENCH_DTYPE", "mix16"))
    ap.add_argument("--kimpl", default=os.environ.get("URN_BENCH_IMPL", "auto"), help="kernel impl: auto|ffma|tc")
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-layer", default="", help="write the per-layer/phase timing table to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner) write to
    # fd 1 directly, so fd 1 is pointed at stderr for the run and the line goes to the saved fd
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from uresnet_b200 import _capi, dp, synthetic_io as S
    from uresnet_b200.uresnet import uresnet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the host-side staging copies of the e2e path are
        # multi-threaded torch copies: give every rank its share of the host cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    dims, B = wl["dims"], (args.batch or wl["batch"])
    lib = _capi.load()

    net = uresnet(dims=dims, num_class=3, num_strides=5, base_num_outputs=16)
    net.construct(trainable=True, use_weight=True, learning_rate=1e-3, dtype=args.dtype, impl=args.kimpl, seed=1234)
    ex = None
    if world > 1:
        dp.broadcast_variables(net)
        ex = dp.attach(net)
    # synthetic batch of this rank (seed 1234 + 1000*rank); weights normalised per event
    data, label, weight = S.make_batch(B, dims, seed=1234 + 1000 * rank)
    weight /= weight.sum(axis=1, keepdims=True)
    d_dev = torch.from_numpy(data).cuda()
    l_dev = torch.from_numpy(label).cuda()
    w_dev = torch.from_numpy(weight).cuda()

    def step_device():
        net.zero_gradients(None)
        net.accum_gradients(None, d_dev, l_dev, w_dev, fetch=False)
        net.apply_gradients(None)

    def step_e2e():
        net.zero_gradients(None)
        res, _ = net.accum_gradients(None, data, label, weight)      # host arrays in, metrics out
        net.apply_gradients(None)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    # per-layer / per-phase CUDA-event profile of the same step (roofline numerator times), taken
    # right before the timed region on the same stream: the event pairs cost ~1 % of a step and
    # keep the launch sequence out of a CUDA graph, so they are not left inside it
    prof_steps = max(1, min(args.steps, 3))
    launches0 = lib.urn_launch_count()
    net.profile_enable(B, True)
    for _ in range(prof_steps):
        step_device()
    prof_ms, prof_cnt = net.profile_collect(B)
    net.profile_enable(B, False)
    launches_per_step = (lib.urn_launch_count() - launches0) // prof_steps
    for _ in range(3):          # graph capture happens on the third un-profiled call of each sequence
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    sampler.stop_flag = True
    launches = launches_per_step * args.steps      # kernels executed in the timed region (graph replays included)
    ms = ev0.elapsed_time(ev1) / args.steps

    # end to end through the public API with host buffers
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(3):        # warm-up: staging buffers, graph capture for the staged-input pointers
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = step_e2e()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    barrier()

    if world > 1:
        t = torch.tensor([ms, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1])
    if rank == 0:
        pk = peaks()
        flops = conv_flops(net.layer_table(B))
        ph = {n: float(prof_ms[:, i].sum()) / prof_steps for i, n in enumerate(net.PHASES)}      # ms per step
        conv_ms = ph["conv_fwd"] + ph["dgrad"] + ph["wgrad"]
        conv_fl = flops["conv_fwd"] + flops["dgrad"] + flops["wgrad"]
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        n_conv_launch_groups = int(prof_cnt[:, [0, 4, 5]].sum() / prof_steps)
        clocks = sampler.summary()
        line = {
            "metric": "train images/sec", "value": world * B / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"mix16": "f16 activations/filters, bf16 gradients, f32 accumulate"}.get(args.dtype, args.dtype),
            "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": B, "global_batch": world * B, "kernel_impl": args.kimpl,
                       "l2": "inputs larger than L2 (%.1f GB of activations per step)" % (net.workspace_bytes(B) / 1e9),
                       "parallelism": "dp%d" % world,
                       "cuda_graphs": "forward+backward" if net._use_graphs else "off"},
            "clocks": clocks,
            "e2e": {"value": world * B / t_e2e, "unit": "images/s", "h2d_bytes_per_step": int(data.nbytes + label.nbytes + weight.nbytes),
                    "d2h_bytes_per_step": 32, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv fwd + dgrad + wgrad launches (%d launch groups/step)" % n_conv_launch_groups,
                         "achieved": achieved, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sus"],
                         "peak_source": pk["src"] + " (sustained bf16)", "traffic": None,
                         "share_of_step": conv_ms / ms if ms > 0 else None, "profiled_steps": prof_steps,
                         "top_launch": top_launch(net.layer_table(B), prof_ms, prof_steps, pk, args.workload)},
            "roofline_hbm": {"bound": "hbm", "kernel": "whole step (all kernels), algorithmic bytes lower bound",
                             "achieved": hbm_bytes(net.layer_table(B)) / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": hbm_bytes(net.layer_table(B)) / (ms * 1e-3) / 1e9 / pk["hbm"], "peak_source": pk["src"]},
            "phases_ms_per_step": {k: round(v, 3) for k, v in ph.items() if k != "-"},
            "loss": float(res[1]),
        }
        # HBM-side view of the bandwidth kernels (BN apply/bwd/stats): bytes are reported in DESIGN.md
        if not args.no_cpu_baseline and world == 1:
            try:
                cb, _ = cpu_sample(args.workload, 2, 1, 1)
                line["cpu_baseline"] = cb
            except Exception as e:  # keep the GPU line even if the host runs out of memory
                line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        if args.per_layer:
            tab = net.layer_table(B)
            with open(args.per_layer, "w") as f:
                f.write("layer,scope," + ",".join(net.PHASES[:7]) + "\n")
                for i, (scope, d, tr) in enumerate(tab):
                    f.write("%d,%s,%s\n" % (i, scope, ",".join("%.4f" % (prof_ms[i, j] / prof_steps) for j in range(7))))
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
