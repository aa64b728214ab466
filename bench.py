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


# dram__bytes_read.sum + dram__bytes_write.sum per launch of committed `ncu --set full` captures
# (profiles/): keyed by (workload, "cin->cout kK sS @side", phase)
NCU_TRAFFIC = {
    # profiles/raw_*_r02c_{2d,3d}.summary.txt (ncu --set full, B = 32 / B = 4; dram bytes read + written per launch)
    ("2d", "16->16 k3 s1 @512", "conv_fwd"): 331.9e6 + 224.8e6,    # gconv_halo_kernel<2,0,0,1> (images-as-rows, fp16 raw out)
    ("2d", "16->16 k3 s1 @512", "dgrad"): 332.7e6 + 225.9e6,       # gconv_halo_kernel<1,0,0,0>
    ("2d", "16->16 k3 s1 @512", "wgrad"): 634.5e6 + 3.7e6,         # wgrad_halo_kernel<0> (images-as-rows, TMA)
    ("3d", "16->16 k3 s1 @192", "dgrad"): 1515.7e6 + 871.0e6,      # gconv_halo_kernel<1,0,1,0>
    ("3d", "16->16 k3 s1 @192", "wgrad"): 2153.0e6 + 4.8e6,        # wgrad_halo_kernel<0>
}


def launch_groups(layer_table, prof_ms, prof_steps, pk, workload, elt_act=2, elt_grad=2, elt_raw=4):
    """Conv launch groups of one step keyed by (phase, geometry): the per-launch CUDA-event times of all
    layers with the same shape are pooled.  Every group is rated against the roofline its arithmetic
    intensity selects: algorithmic bytes = each operand tensor once (activations `elt_act` B, gradients
    `elt_grad` B, raw conv output `elt_raw` B as stored), algorithmic FLOPs = 2*|Y|*taps*cx*cy."""
    phases = {0: "conv_fwd", 4: "wgrad", 5: "dgrad"}
    ridge = pk["tf_sus"] * 1e12 / (pk["hbm"] * 1e9)
    groups = {}
    for i, (scope, d, tr) in enumerate(layer_table):
        big = float(d.n) * d.xd * d.xh * d.xw * d.cx
        small = float(d.n) * d.yd * d.yh * d.yw * d.cy
        ein, eout = (small, big) if tr else (big, small)
        flops = 2.0 * d.n * d.yd * d.yh * d.yw * d.kd * d.kh * d.kw * d.cx * d.cy
        side = max(d.xh, d.xw)
        shape = "%d->%d k%d s%d @%d" % (((d.cy, d.cx) if tr else (d.cx, d.cy)) + (d.kh, d.s, side))
        for col, name in phases.items():
            ms = float(prof_ms[i, col]) / prof_steps
            if ms <= 0.0:
                continue
            er = elt_raw if (d.cy % 16 == 0 or tr) and scope not in ("conv0", "conv2") else 4
            byts = {"conv_fwd": ein * elt_act + eout * er, "dgrad": eout * elt_grad + ein * elt_grad,
                    "wgrad": ein * elt_act + eout * elt_grad}[name]
            g = groups.setdefault((name, shape, tr), dict(ms=0.0, bytes=0.0, flops=0.0, launches=0, layers=[]))
            g["ms"] += ms
            g["bytes"] += byts
            g["flops"] += flops
            g["launches"] += 1
            g["layers"].append(scope)
    out = []
    for (name, shape, tr), g in groups.items():
        hbm_bound = g["flops"] / g["bytes"] < ridge
        sec = g["ms"] * 1e-3
        achieved = g["bytes"] / sec / 1e9 if hbm_bound else g["flops"] / sec / 1e12
        peak = pk["hbm"] if hbm_bound else pk["tf_sus"]
        traffic = NCU_TRAFFIC.get((workload, shape, name))
        out.append({"kernel": "%s %s%s" % (name, shape, " (transposed conv)" if tr else ""), "layers": g["layers"],
                    "bound": "hbm" if hbm_bound else "tensor", "achieved": achieved, "peak": peak,
                    "unit": "GB/s" if hbm_bound else "TFLOP/s", "frac": achieved / peak,
                    "launches_per_step": g["launches"], "ms_per_launch": g["ms"] / g["launches"],
                    "ms_per_step": g["ms"], "algorithmic_bytes_per_launch": g["bytes"] / g["launches"],
                    "algorithmic_flops_per_launch": g["flops"] / g["launches"],
                    "traffic": traffic})
    return out


def hbm_bytes(layer_table, elt_act=2, elt_grad=2, elt_raw=4):
    """Algorithmic HBM bytes of one training step by SURVEY 8d's definition (every logical tensor once
    per kernel that must touch it; `e` = stored element size): conv forward (input + raw output, `elt_raw`
    bytes as stored: 4, or 2 in mix16 except stem and head),
    BN statistics 0 (fused into the conv epilogue), BN apply (raw in, activation out: 2 tensors), BN
    backward reductions (dout + raw: 2 tensors) and apply (dout + raw in, d(raw) out: 3 tensors), input
    gradient (d(raw) in, dX out), filter gradient (X + d(raw) in).  Residual / shortcut / concat extras,
    the bf16 shadow copies this implementation also writes, CE, Adam and filter packing are NOT
    counted: a lower bound of what the step has to move."""
    tot = 0.0
    for scope, d, tr in layer_table:
        big = float(d.n) * d.xd * d.xh * d.xw * d.cx      # X side elements
        small = float(d.n) * d.yd * d.yh * d.yw * d.cy    # Y side elements
        ein, eout = (small, big) if tr else (big, small)
        er = elt_raw if scope not in ("conv0", "conv2") else 4
        tot += ein * elt_act + eout * er                   # conv forward
        tot += eout * er + eout * elt_act                  # BN apply
        tot += eout * (elt_grad + er)                      # BN backward reductions
        tot += eout * (elt_grad + er + elt_grad)           # BN backward apply
        if scope != "conv0":
            tot += eout * elt_grad + ein * elt_grad        # input gradient
        tot += ein * elt_act + eout * elt_grad             # filter gradient
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
            "config": {"workload": WORKLOADS[args.workload]["name"], "per_gpu_batch": WORKLOADS[args.workload]["batch"],
                       "global_batch": args.gpus * WORKLOADS[args.workload]["batch"], "parallelism": "dp%d" % args.gpus,
                       "note": "CPU restatement of the reference graph (the TF-1.x reference cannot run here), all host cores; "
                               "every step is a bounded sample of the workload: " + cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class Ctx(object):
    pass


def make_inputs(net, dims, B, rank, pinned):
    """Two synthetic batches of this rank (seeds (1234, rank, 0/1)), weights normalised per
    event (ssnet_trainval.py:173); host copies in page-locked memory when `pinned`."""
    from uresnet_b200 import synthetic_io as S
    sets = []
    for step in range(2):
        out = None
        if pinned:
            P = int(np.prod(dims[:-1]))
            out = tuple(net.pinned_empty((B, P)) for _ in range(3))
        d, l, w = S.make_batch(B, dims, seed=(1234, rank, step), out=out)
        w /= w.sum(axis=1, keepdims=True)
        sets.append((d, l, w))
    return sets


def run_leg(cx, workload, dtype, B, steps, warmup, want_e2e=True, want_profile=True, nmb=1, kimpl="auto"):
    """One workload in one numeric mode: device-resident timed region, live per-layer profile, end-to-end
    timed region.  Returns a dict; the network and its workspace are released before returning."""
    import torch
    import torch.distributed as dist
    from uresnet_b200 import dp
    from uresnet_b200.uresnet import uresnet
    wl = WORKLOADS[workload]
    dims = wl["dims"]
    world, rank = cx.world, cx.rank
    net = uresnet(dims=dims, num_class=3, num_strides=5, base_num_outputs=16)
    net.construct(trainable=True, use_weight=True, learning_rate=1e-3, dtype=dtype, impl=kimpl, seed=1234)
    ex = None
    if world > 1:
        dp.broadcast_variables(net)
        ex = dp.attach(net)
        ex.expected_minibatches = nmb
        ex.time_finish = True
    host = make_inputs(net, dims, B, rank, pinned=True)
    dev = [tuple(torch.from_numpy(a).cuda() for a in hs) for hs in host]

    def step_device():
        net.zero_gradients(None)
        for j in range(nmb):
            net.accum_gradients(None, *dev[j % 2], fetch=False)
        net.apply_gradients(None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    barrier()
    out = {"workload": wl["name"], "mode": dtype, "per_gpu_batch": B * nmb, "global_batch": world * B * nmb}
    prof_ms = None
    launches_per_step = 0
    if want_profile:
        # per-layer / per-phase CUDA-event profile of the same step (roofline numerator times), taken
        # right before the timed region on the same stream: the event pairs cost ~1 % of a step and
        # keep the launch sequence out of a CUDA graph, so they are not left inside it
        prof_steps = max(1, min(steps, 3))
        launches0 = cx.lib.urn_launch_count()
        net.profile_enable(B, True)
        for _ in range(prof_steps):
            step_device()
        prof_ms, prof_cnt = net.profile_collect(B)
        net.profile_enable(B, False)
        launches_per_step = (cx.lib.urn_launch_count() - launches0) // prof_steps
    for _ in range(3 * (2 if nmb > 1 else 1)):   # graph capture happens on the third un-profiled call of each sequence
        step_device()
    if ex is not None:
        ex.exposed_ms()
    sampler = ClockSampler(cx.local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step_device()
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1) / steps
    exposed = ex.exposed_ms() if ex is not None else None

    t_e2e, e2e_steps, res = None, 0, [None, float("nan"), None, None]
    if want_e2e:
        # end to end through the public API: HOST arrays in (page-locked, the two batches alternate), the
        # transfer of the next step's inputs enqueued under the current step's kernels (stage_next), loss and
        # accuracies read back every step
        e2e_steps = max(2, min(steps, 6))

        def step_e2e(i):
            net.zero_gradients(None)
            for j in range(nmb):
                net.accum_gradients(None, *host[(i * nmb + j) % 2], fetch=False)
                net.stage_next(*host[(i * nmb + j + 1) % 2])
            net.apply_gradients(None)
            return net.fetch_metrics(B)[0]

        for i in range(8):        # warm-up: staging slots, graph capture for both sets of staged-input pointers
            step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            res = step_e2e(i)
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        barrier()

    # parameter checksum (data-parallel replicas must stay identical)
    csum = float(net._params.double().sum().item())
    cabs = float(net._params.double().abs().sum().item())
    if world > 1:
        t = torch.tensor([ms, t_e2e or 0.0, exposed or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e_max, exposed = float(t[0]), float(t[1]), float(t[2])
        t_e2e = t_e2e_max if want_e2e else None
        c = torch.tensor([csum, cabs], dtype=torch.float64, device="cuda")
        cmin, cmax = c.clone(), c.clone()
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        out["replicas_identical"] = bool(torch.equal(cmin, cmax))
        out["param_checksum"] = [float(cmin[0]), float(cmin[1])]
        out["exposed_allreduce_ms"] = exposed
    imgs = world * B * nmb
    out.update({"value": imgs / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "steps": steps,
                "clocks": sampler.summary(),
                "workspace_gb": net.workspace_bytes(B) / 1e9,
                "cuda_graphs": "forward+backward" if net._use_graphs else "off"})
    if want_e2e:
        out["e2e"] = {"value": imgs / t_e2e, "unit": "images/s",
                      "h2d_bytes_per_step": int(sum(a.nbytes for a in host[0]) * nmb), "d2h_bytes_per_step": 32,
                      "steps": e2e_steps, "host_buffers": "page-locked, next step's transfer overlapped (stage_next)"}
        out["loss"] = float(res[1])
    if want_profile and rank == 0:
        pk = cx.pk
        tab = net.layer_table(B)
        ea = 4 if dtype == "f32" else 2
        er = 2 if (dtype == "mix16" and not os.environ.get("URN_RAW_F32")) else 4
        flops = conv_flops(tab)
        ph = {n: float(prof_ms[:, i].sum()) / prof_steps for i, n in enumerate(net.PHASES)}      # ms per step
        conv_ms = ph["conv_fwd"] + ph["dgrad"] + ph["wgrad"]
        conv_fl = (flops["conv_fwd"] + flops["dgrad"] + flops["wgrad"]) * nmb
        groups = launch_groups(tab, prof_ms, prof_steps * nmb, pk, workload, ea, ea, er)
        dom = max(groups, key=lambda g: g["ms_per_step"])                 # largest share of the step
        top = max(groups, key=lambda g: g["ms_per_launch"])               # slowest single launch
        for g in (dom, top):
            g["share_of_step"] = g["ms_per_step"] * nmb / ms
        byts = hbm_bytes(tab, ea, ea, er) * nmb
        out["gpu_launches_per_step"] = int(launches_per_step)
        out["roofline"] = dict(dom, peak_source=pk["src"], profiled_steps=prof_steps,
                               note="dominant conv launch group of the step (largest summed CUDA-event time); "
                                    "bound chosen by its arithmetic intensity against the measured ridge")
        out["top_launch"] = top
        out["roofline_conv_tensor"] = {"bound": "tensor", "kernel": "all conv fwd + dgrad + wgrad launches",
                                       "achieved": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
                                       "peak": pk["tf_sus"], "unit": "TFLOP/s",
                                       "frac": (conv_fl / (conv_ms * 1e-3) / 1e12 / pk["tf_sus"]) if conv_ms > 0 else 0.0,
                                       "share_of_step": conv_ms / ms, "peak_source": pk["src"] + " (sustained bf16)"}
        out["roofline_hbm"] = {"bound": "hbm", "kernel": "whole step, algorithmic bytes by SURVEY 8d (lower bound)",
                               "achieved": byts / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                               "frac": byts / (ms * 1e-3) / 1e9 / pk["hbm"], "peak_source": pk["src"]}
        out["phases_ms_per_step"] = {k: round(v, 3) for k, v in ph.items() if k != "-"}
        if cx.per_layer:
            with open(cx.per_layer + "." + workload + "." + dtype + ".csv", "w") as f:
                f.write("layer,scope," + ",".join(net.PHASES[:7]) + "\n")
                for i, (scope, d, tr) in enumerate(tab):
                    f.write("%d,%s,%s\n" % (i, scope, ",".join("%.4f" % (prof_ms[i, j] / prof_steps) for j in range(7))))
    net._exchange = None
    del net, dev, host, ex
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def dp_equivalence_check(cx):
    """R ranks with all-reduce SUM  ==  one replica with NUM_MINIBATCHES = R (lib/ssnet.py:76-79,
    lib/ssnet_trainval.py:164-191: accumulate a SUM of per-minibatch gradients, each minibatch with its own
    batch-norm statistics, then ONE Adam apply).  Tiny f32 net; every rank ALSO accumulates all R batches
    locally and the updated parameters are compared."""
    import torch
    import torch.distributed as dist
    from uresnet_b200 import dp, synthetic_io as S
    from uresnet_b200.uresnet import uresnet
    dims, B, F, St = [64, 64, 1], 2, 8, 3
    batches = []
    for r in range(cx.world):
        d, l, w = S.make_batch(B, dims, seed=(77, r, 0))
        w /= w.sum(axis=1, keepdims=True)
        batches.append((d, l, w))
    res, grads = {}, {}
    for tag in ("dp", "accum"):
        net = uresnet(dims=dims, num_class=3, num_strides=St, base_num_outputs=F)
        net.construct(trainable=True, use_weight=True, learning_rate=1e-3, dtype="f32", seed=99)
        if tag == "dp":
            dp.broadcast_variables(net)
            dp.attach(net)
        for it in range(3):
            net.zero_gradients(None)
            for (d, l, w) in ([batches[cx.rank]] if tag == "dp" else batches):
                net.accum_gradients(None, d, l, w)
            net.apply_gradients(None)
            if it == 0:
                grads[tag] = net._accum.clone()      # the (all-reduced / accumulated) gradient sum of the first step
        res[tag] = net._params.clone()
        net._exchange = None
    gdiff = float((grads["dp"] - grads["accum"]).abs().max() / grads["accum"].abs().max())
    pdiff = float((res["dp"] - res["accum"]).abs().max() / res["accum"].abs().max())
    t = torch.tensor([gdiff, pdiff], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"what": "%d-rank all-reduce SUM vs 1 replica with NUM_MINIBATCHES=%d, f32, 64x64 F=8 S=3: summed gradients of "
                    "step 1 (bound 1e-6: fp32 summation order) and parameters after 3 Adam steps (bound 1e-5)" % (cx.world, cx.world),
            "max_rel_grad_diff": float(t[0]), "max_rel_param_diff": float(t[1]),
            "ok": bool(float(t[0]) < 1e-6 and float(t[1]) < 1e-5)}


def inference_b1(cx):
    """BASELINE configs[0]: U-ResNet 2D inference, batch 1, 512x512 (lib/ssnet.py:130-139: softmax out),
    host array in / host softmax out, GPU (mix16 and f32) and the CPU oracle on the host cores."""
    import torch
    from uresnet_b200 import synthetic_io as S
    from uresnet_b200.uresnet import uresnet
    dims = [512, 512, 1]
    data, label, weight = S.make_batch(1, dims, seed=(1234, 0, 0))
    out = {"workload": "U-ResNet 2D inference, batch 1, 512x512, F=16, S=5, 3 classes (BASELINE configs[0])"}
    for dtype in ("mix16", "f32"):
        net = uresnet(dims=dims, num_class=3, num_strides=5, base_num_outputs=16)
        net.construct(trainable=False, use_weight=False, dtype=dtype, seed=1234)
        d_dev = torch.from_numpy(data).cuda()
        for _ in range(5):
            net.inference(None, data)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 20
        for _ in range(n):
            net.inference(None, data)
        torch.cuda.synchronize()
        out["gpu_ms_e2e_" + dtype] = (time.perf_counter() - t0) / n * 1e3
        feed = {"data": d_dev}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            net._forward(feed)
        e0.record()
        for _ in range(n):
            net._forward(feed)
        e1.record()
        torch.cuda.synchronize()
        out["gpu_ms_device_" + dtype] = e0.elapsed_time(e1) / n
        del net
        torch.cuda.empty_cache()
    try:
        from oracle import uresnet_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        P = O.init_params(2, seed=4321)
        x = torch.tensor(data.reshape([1, 512, 512, 1]), dtype=torch.float32)
        ts = []
        with torch.no_grad():
            for i in range(7):
                t0 = time.perf_counter()
                torch.softmax(O.forward(P, x), dim=-1)
                if i >= 2:
                    ts.append(time.perf_counter() - t0)
        out["cpu_ms"] = float(np.median(ts)) * 1e3
        out["cpu_cores"] = os.cpu_count()
        out["cpu_kind"] = "port (oracle forward, torch-CPU fp32)"
    except Exception as e:
        out["cpu_ms"] = None
        out["cpu_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("URN_BENCH_WORKLOAD", "2d"), choices=list(WORKLOADS),
                    help="primary workload (top-level keys of the JSON line)")
    ap.add_argument("--dtype", default=os.environ.get("URN_BENCH_DTYPE", "mix16"))
    ap.add_argument("--kimpl", default=os.environ.get("URN_BENCH_IMPL", "auto"), help="kernel impl: auto|ffma|tc")
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch of the primary workload")
    ap.add_argument("--only", action="store_true", help="primary workload only (no 3D / f32 / inference / strong legs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-layer", default="", help="write per-layer/phase timing tables to <prefix>.<workload>.<mode>.csv")
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
    from uresnet_b200 import _capi

    cx = Ctx()
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    cx.per_layer = args.per_layer
    if cx.world != args.gpus and cx.world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, cx.world))
    torch.cuda.set_device(cx.local)
    if cx.world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the host-side staging copies of pageable inputs are
        # multi-threaded torch copies: give every rank its share of the host cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // cx.world))
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local))
    cx.lib = _capi.load()
    cx.pk = peaks()
    world, rank = cx.world, cx.rank
    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]

    main_leg = run_leg(cx, args.workload, args.dtype, B, args.steps, args.warmup, kimpl=args.kimpl)
    line = None
    if rank == 0:
        line = {
            "metric": "train images/sec", "value": main_leg["value"], "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_leg["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"mix16": "f16 activations/filters, bf16 gradients, f32 accumulate"}.get(args.dtype, args.dtype),
            "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": B, "global_batch": world * B, "kernel_impl": args.kimpl,
                       "l2": "inputs larger than L2 (%.1f GB of activations per step)" % main_leg["workspace_gb"],
                       "parallelism": "dp%d" % world, "cuda_graphs": main_leg["cuda_graphs"]},
            "clocks": main_leg["clocks"],
            "e2e": main_leg["e2e"],
            "gpu_launches": int(main_leg.get("gpu_launches_per_step", 0) * args.steps),
            "roofline": main_leg.get("roofline"),
            "top_launch": main_leg.get("top_launch"),
            "roofline_conv_tensor": main_leg.get("roofline_conv_tensor"),
            "roofline_hbm": main_leg.get("roofline_hbm"),
            "phases_ms_per_step": main_leg.get("phases_ms_per_step"),
            "loss": main_leg.get("loss"),
        }
        for k in ("replicas_identical", "param_checksum", "exposed_allreduce_ms"):
            if k in main_leg:
                line.setdefault("dp_check", {})[k] = main_leg[k]
    if not args.only:
        legs = {}
        other = "3d" if args.workload == "2d" else "2d"
        k3 = max(3, min(args.steps, 10))
        legs[other] = run_leg(cx, other, args.dtype, WORKLOADS[other]["batch"], k3, 3)
        if world == 1:
            # the 1e-3 parity mode (fp32 storage, CUDA-core FFMA with two-level accumulation), same configs
            f32 = {}
            for w_ in ("2d", "3d"):
                f32[w_] = run_leg(cx, w_, "f32", WORKLOADS[w_]["batch"], 3, 3, want_e2e=False)
        # BASELINE configs[2]: 2D, GLOBAL batch 256 (strong scaling): 256/N per GPU in one minibatch for N >= 2;
        # one GPU holds it as NUM_MINIBATCHES = 8 x 32 (the reference's own way to reach a large batch)
        gb = 256
        if world == 1:
            strong = run_leg(cx, "2d", args.dtype, 32, max(2, min(args.steps, 4)), 3, want_e2e=False, want_profile=False, nmb=8)
        else:
            strong = run_leg(cx, "2d", args.dtype, gb // world, max(3, min(args.steps, 6)), 3, want_e2e=False, want_profile=False)
        dpc = dp_equivalence_check(cx) if world > 1 else None
        if rank == 0:
            line["workloads"] = legs
            if world == 1:
                line["f32"] = f32
                line["inference_b1"] = inference_b1(cx)
            strong["scaling"] = "strong"
            line["strong_2d_global256"] = strong
            if dpc is not None:
                line.setdefault("dp_check", {})["accumulation_equivalence"] = dpc
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                cb, _ = cpu_sample(args.workload, 2, 1, 1)
                line["cpu_baseline"] = cb
            except Exception as e:  # keep the GPU line even if the host runs out of memory
                line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
