#!/usr/bin/env python
"""bench.py -- train images/sec of the CNN layer hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (own arm, B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU arm)

Workload (BASELINE.json metric "train images/sec at 1/2/4/8 B200", configs[4]): ResNet-50
v1.5 224x224 training, batch 256 per GPU, synthetic FP32 images and one-hot labels,
random-init weights, SGD momentum 0.9, built through the bcnn C API (bcnn_b200/configs.py).
One step = bcnn_forward + bcnn_backward + bcnn_update over one batch (the body of
bcnn_train_on_batch). `value` times K steps with inputs resident in HBM; `e2e` times the
same K steps with the per-step host->device upload of the batch from pinned memory and a
device->host read of the loss inside the timed region. N > 1: one process per GPU (torchrun),
weak scaling (256 images per GPU), NCCL all-reduce of weight gradients overlapped with
backward; time = max over ranks of the CUDA-event time on each rank's stream.

Own arm additions: `roofline` (dominant kernel, measured live with CUDA events on the net's
stream), `rooflines` (every kernel class of the path), `cpu_baseline` (the compiled reference
CPU library timed on this box's host cores on a bounded sample; reported, not a target).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "train_images_per_sec"
UNIT = "images/s"


# --------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------

def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"],
                    tc_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# --------------------------------------------------------------------------------
# reference arm / cpu_baseline: the compiled reference CPU library (oracle/_ref)
# --------------------------------------------------------------------------------

def reference_lib():
    """oracle/_ref/libbcnn_ref.so -- the UNMODIFIED reference CPU path. This is the one
    place outside tests/ where bench.py executes oracle/ (as the measured baseline arm)."""
    from bcnn_b200 import capi
    so = ROOT / "oracle" / "_ref" / "libbcnn_ref.so"
    if not so.exists():
        return None
    lib = C.CDLL(str(so), mode=C.RTLD_LOCAL | getattr(os, "RTLD_DEEPBIND", 0))
    capi.bind_bcnn_api(lib, capi.TensorCPU)
    vp, i = C.c_void_p, C.c_int
    for name, (res, args) in {"bcnn_b200_num_nodes": (i, [vp]), "bcnn_b200_num_tensors": (i, [vp]),
                              "bcnn_b200_node_type": (i, [vp, i]),
                              "bcnn_b200_node_src": (i, [vp, i, i]),
                              "bcnn_b200_node_dst": (i, [vp, i, i]),
                              "bcnn_ref_num_threads": (i, [vp])}.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def build_workload(net, workload, batch, res):
    from bcnn_b200 import configs
    if workload in ("mnist", "cifar"):
        return configs.BUILDERS[workload](net, batch=batch)
    return configs.BUILDERS[workload](net, batch=batch, res=res)


def reference_net(lib, workload, batch, threads, res):
    from bcnn_b200 import capi, configs
    net = capi.Net(mode=capi.MODE_TRAIN, lib=lib, flavour="cpu")
    lib.bcnn_set_num_threads(net.handle, threads, None)
    build_workload(net, workload, batch, res)
    net.compile()
    configs.init_params(net, seed=2024)
    net.set("input", configs.synth_input(net.shape("input")))
    net.set("label", configs.synth_labels(net.shape("label")))
    return net


def time_reference(workload, res, steps, warmup, budget_s):
    """Times fwd+bwd+update of the reference CPU library on a bounded sample (batch sized so
    the whole run fits budget_s). Returns images/s and a description."""
    lib = reference_lib()
    if lib is None:
        return None
    threads = max(1, min(os.cpu_count() or 1, 64))
    probe_batch = 2
    net = reference_net(lib, workload, probe_batch, threads, res)
    used = lib.bcnn_ref_num_threads(net.handle)
    t0 = time.perf_counter()
    net.forward(); net.backward(); net.update()
    per_image = (time.perf_counter() - t0) / probe_batch
    net.close()
    per_step_budget = budget_s / max(1, steps + warmup)
    batch = int(max(1, min(16, per_step_budget // max(per_image, 1e-9))))
    net = reference_net(lib, workload, batch, threads, res)
    for _ in range(warmup):
        net.forward(); net.backward(); net.update()
    t0 = time.perf_counter()
    for _ in range(steps):
        net.forward(); net.backward(); net.update()
    dt = time.perf_counter() - t0
    net.close()
    return dict(value=batch * steps / dt, ms_per_step=1e3 * dt / steps, cores=used, batch=batch,
                sample=f"{steps} steps (+{warmup} warm-up) of {workload} at batch {batch} "
                       f"(fwd+bwd+SGD, internal bcnn_gemm, AVX2+OpenMP, {used} threads); "
                       f"probe step {per_image * 1e3:.0f} ms/image")


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cfg = dict(workload=f"{args.workload} {args.res}x{args.res} training (reference CPU arm)",
               per_gpu_batch=args.batch, global_batch=args.batch * args.gpus)
    r = time_reference(args.workload, args.res, args.steps, args.warmup, budget_s=150.0)
    if r is None:
        emit(json.dumps({"impl": "reference", "unavailable":
                         "oracle/_ref/libbcnn_ref.so missing (reference CPU library not built)"}))
        return 0
    cfg["cpu_sample_batch"] = r["batch"]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                             "kind": "reference", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------
# roofline micro-measurements (CUDA events on the net's stream, inputs > L2)
# --------------------------------------------------------------------------------

def event_time_ms(lib, stream, fn, iters):
    e0, e1 = lib.bcnn_b200_event_create(), lib.bcnn_b200_event_create()
    fn()
    lib.bcnn_b200_stream_sync(stream)
    lib.bcnn_b200_event_record(e0, stream)
    for _ in range(iters):
        fn()
    lib.bcnn_b200_event_record(e1, stream)
    ms = lib.bcnn_b200_event_elapsed_ms(e0, e1) / iters
    lib.bcnn_b200_event_destroy(e0); lib.bcnn_b200_event_destroy(e1)
    return ms


# (cin, h, cout, k, stride, pad, operand type of the kernel, share-of-step note)
CONV_ROOFLINE_SHAPES = (
    (64, 56, 256, 1, 1, 0, "tf32"),    # DIRECT 1x1: conv_tma_fwd_kernel<0,0>, the kernel with the largest step share
    (256, 56, 64, 1, 1, 0, "tf32"),
    (256, 14, 1024, 1, 1, 0, "tf32"),
    (64, 56, 64, 3, 1, 1, "bf16"),     # NHWC-shadow route, kind::f16
    (256, 14, 256, 3, 1, 1, "bf16"),
    (512, 7, 512, 3, 1, 1, "bf16"),    # the compute-bound end of the net
)


def ncu_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` captures (profiles/), keyed by
    roofline entry name; made by tools/ncu_rooflines.sh + tools/ncu_traffic.py."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    return json.loads(p.read_text()) if p.exists() else {}


def kernel_rooflines(lib, stream, math, peaks, batch, only=""):
    """Algorithmic bytes / flops per launch over the measured launch time, at the
    bench batch, for the kernel classes of the path. Convolutions are reported against the roof
    that bounds the shape: with FP32 NCHW tensors in HBM most ResNet-50 layers sit left of the
    ridge, so their bound is "hbm" (algorithmic bytes = read x and W once, write y once);
    `tc_frac` is always given beside it. Convolution entries time the whole C-ABI call (weight
    pack, NHWC shadow of the shadow-route shapes, main kernel, split-K reduction)."""
    from bcnn_b200 import capi
    out = []
    n = batch
    traffic = ncu_traffic()

    def buf(elems):
        return capi.DeviceBuffer(nbytes=int(elems) * 4)

    def want(name):
        return only in name

    # --- batchnorm on [n, 256, 56, 56] (largest BN class of the net)
    c, hw = 256, 56 * 56
    E = n * c * hw
    if any(want(k) for k in ("bn_forward_train(stats+apply+relu)", "bn_backward(reduce+apply, relu fused)",
                             "bn_apply+relu (statistics from the conv epilogue)", "eltwise_add_relu")):
        x, y, dy = buf(E), buf(E), buf(E)
        prm = [buf(c) for _ in range(9)]
        scratch = buf(lib.bcnn_b200_bn_scratch_floats(c))
        lib.bcnn_b200_fill_f32(prm[3].ptr, c, 1.0, stream)  # var
        lib.bcnn_b200_fill_f32(prm[4].ptr, c, 1.0, stream)  # gamma
        if want("bn_forward_train(stats+apply+relu)"):
            ms = event_time_ms(lib, stream, lambda: (
                lib.bcnn_b200_bn_stats(x.ptr, n, c, hw, prm[0].ptr, prm[1].ptr, prm[2].ptr, prm[3].ptr,
                                       scratch.ptr, stream),
                lib.bcnn_b200_bn_apply(x.ptr, y.ptr, prm[0].ptr, prm[1].ptr, prm[4].ptr, prm[5].ptr, n,
                                       c, hw, 2, stream)), 5)
            out.append(dict(kernel="bn_forward_train(stats+apply+relu)", shape=[n, c, 56, 56],
                            bound="hbm", bytes=12 * E, ms=ms))
        if want("bn_apply+relu (statistics from the conv epilogue)"):
            ms = event_time_ms(lib, stream, lambda: lib.bcnn_b200_bn_apply(
                x.ptr, y.ptr, prm[0].ptr, prm[1].ptr, prm[4].ptr, prm[5].ptr, n, c, hw, 2, stream), 5)
            out.append(dict(kernel="bn_apply+relu (statistics from the conv epilogue)",
                            shape=[n, c, 56, 56], bound="hbm", bytes=8 * E, ms=ms))
        if want("bn_backward(reduce+apply, relu fused)"):
            ms = event_time_ms(lib, stream, lambda: lib.bcnn_b200_bn_backward(
                x.ptr, y.ptr, dy.ptr, dy.ptr, prm[0].ptr, prm[3].ptr, prm[4].ptr, prm[5].ptr,
                prm[6].ptr, prm[7].ptr, prm[8].ptr, prm[2].ptr, n, c, hw, 2, scratch.ptr, stream), 5)
            out.append(dict(kernel="bn_backward(reduce+apply, relu fused)", shape=[n, c, 56, 56],
                            bound="hbm", bytes=20 * E, ms=ms))
        if want("eltwise_add_relu"):
            ms = event_time_ms(lib, stream, lambda: lib.bcnn_b200_eltwise_forward(
                x.ptr, dy.ptr, y.ptr, E, E, 2, stream), 5)
            out.append(dict(kernel="eltwise_add_relu", shape=[n, c, 56, 56], bound="hbm",
                            bytes=12 * E, ms=ms))
        for b in (x, y, dy, scratch, *prm):
            b.free()
    # --- max pool 3x3 s2 on [n, 64, 112, 112]
    if want("maxpool_forward k3s2") or want("maxpool_backward k3s2"):
        Ei, Eo = n * 64 * 112 * 112, n * 64 * 56 * 56
        px, py, pi = buf(Ei), buf(Eo), buf(Eo)
        ms = event_time_ms(lib, stream, lambda: lib.bcnn_b200_maxpool_forward(
            px.ptr, py.ptr, pi.ptr, n, 64, 112, 112, 3, 2, 56, 56, stream), 5)
        out.append(dict(kernel="maxpool_forward k3s2", shape=[n, 64, 112, 112], bound="hbm",
                        bytes=4 * Ei + 8 * Eo, ms=ms))
        ms = event_time_ms(lib, stream, lambda: lib.bcnn_b200_maxpool_backward(
            px.ptr, py.ptr, pi.ptr, n, 64, 112, 112, 3, 2, 56, 56, stream), 5)
        out.append(dict(kernel="maxpool_backward k3s2", shape=[n, 64, 112, 112], bound="hbm",
                        bytes=8 * Eo + 8 * Ei, ms=ms))
        for b in (px, py, pi):
            b.free()
    # --- convolution
    for (cin, hh, cout, k, s, pad, kind) in CONV_ROOFLINE_SHAPES:
        tag = f"{k}x{k} {cin}->{cout} @{hh}"
        if not any(want(f"conv_{nm} {tag}") for nm in ("fprop", "dgrad", "wgrad")):
            continue
        d = capi.ConvDesc.make(n, cin, hh, hh, cout, k, s, pad, 1)
        ws_bytes = lib.bcnn_b200_conv_workspace_bytes(d, math)
        ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
        ex, ey, ew = n * cin * hh * hh, n * cout * d.ho * d.wo, cout * cin * k * k
        cx, cw, cy, cgw = buf(ex), buf(ew), buf(ey), buf(ew)
        flops = 2.0 * n * cout * d.ho * d.wo * cin * k * k
        for name, call, nbytes in (
            ("fprop", lambda: lib.bcnn_b200_conv_forward(d, cx.ptr, cw.ptr, None, 0, cy.ptr, ws.ptr,
                                                         ws_bytes, math, stream), 4 * (ex + ew + ey)),
            ("dgrad", lambda: lib.bcnn_b200_conv_backward_data(d, cw.ptr, cy.ptr, cx.ptr, 0, ws.ptr,
                                                               ws_bytes, math, stream), 4 * (ex + ew + ey)),
            ("wgrad", lambda: lib.bcnn_b200_conv_backward_weights(d, cx.ptr, cy.ptr, cgw.ptr, ws.ptr,
                                                                  ws_bytes, math, stream),
             4 * (ex + ey + 2 * ew))):
            if not want(f"conv_{name} {tag}"):
                continue
            ms = event_time_ms(lib, stream, call, 3)
            out.append(dict(kernel=f"conv_{name} {tag}", shape=[n, cin, hh, hh], flops=flops,
                            bytes=nbytes, ms=ms, kind=kind))
        for b in (ws, cx, cw, cy, cgw):
            b.free()
    res = []
    for r in out:
        t = traffic.get(r["kernel"], {})
        dram = t.get("dram_bytes")
        if "flops" not in r:
            ach = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            res.append(dict(kernel=r["kernel"], shape=r["shape"], bound="hbm", achieved=ach,
                            peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"],
                            ms_per_launch=r["ms"], traffic=dram))
            continue
        # tensor peak of the operand type: the measured BF16 GEMM peak, halved for kind::tf32
        tc_peak = peaks["tc_burst"] * (0.5 if r["kind"] == "tf32" else 1.0)
        t_hbm = r["bytes"] / (peaks["hbm"] * 1e9)
        t_tc = r["flops"] / (tc_peak * 1e12)
        tf = r["flops"] / (r["ms"] * 1e-3) / 1e12
        gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        e = dict(kernel=r["kernel"], shape=r["shape"], ms_per_launch=r["ms"], traffic=dram,
                 operands=r["kind"], tflops=tf, tc_peak=tc_peak, tc_frac=tf / tc_peak,
                 algorithmic_gbs=gb, hbm_frac=gb / peaks["hbm"])
        if t_hbm >= t_tc:
            e.update(bound="hbm", achieved=gb, peak=peaks["hbm"], unit="GB/s", frac=gb / peaks["hbm"])
        else:
            e.update(bound="tensor", achieved=tf, peak=tc_peak, unit="TFLOP/s", frac=tf / tc_peak)
        if t.get("tensor_pipe_pct") is not None:
            e["ncu_tensor_pipe_pct"] = t["tensor_pipe_pct"]
        res.append(e)
    return res


# --------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------

def run_own_arm(args):
    from bcnn_b200 import capi, configs
    rank, local_rank, world = dist_env()
    lib = capi.b200()
    if lib.bcnn_b200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    lib.bcnn_b200_set_device(local_rank)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    math = {"tc": capi.MATH_TC, "fp32": capi.MATH_FP32, "resident": capi.MATH_TC_BF16}[args.math]
    net = capi.Net(mode=capi.MODE_TRAIN)
    net.set_conv_math(math)
    net.set_reference_quirks(False)  # batch-correct residual adds
    build_workload(net, args.workload, args.batch, args.res)
    net.compile()
    configs.init_params(net, seed=2024)  # identical weights on every rank
    x = configs.synth_input(net.shape("input"), seed=12345 + rank)
    y = configs.synth_labels(net.shape("label"), first_sample=rank * args.batch)
    net.set("input", x)
    net.set("label", y)
    h2d_bytes = x.nbytes + y.nbytes

    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_char * 128)()
            if lib.bcnn_b200_dp_get_unique_id(raw) != 0:
                raise SystemExit("ncclGetUniqueId failed")
            uid.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        raw = (C.c_char * 128).from_buffer_copy(bytes(uid.cpu().numpy().tobytes()))
        if lib.bcnn_b200_dp_init(net.handle, rank, world, raw) != 0:
            raise SystemExit("bcnn_b200_dp_init failed")

    stream = lib.bcnn_b200_get_stream(net.handle)

    def barrier():
        net.sync()
        if dist:
            dist.barrier()
        net.sync()

    for _ in range(args.warmup):
        net.train_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing ----
    e0, e1 = lib.bcnn_b200_event_create(), lib.bcnn_b200_event_create()
    launches0 = lib.bcnn_b200_launch_count()
    barrier()
    lib.bcnn_b200_event_record(e0, stream)
    for _ in range(args.steps):
        net.train_step()
    lib.bcnn_b200_event_record(e1, stream)
    barrier()
    ms_total = lib.bcnn_b200_event_elapsed_ms(e0, e1)
    launches = lib.bcnn_b200_launch_count() - launches0
    # ---- end-to-end timing: + H2D of the batch and D2H of the loss every step ----
    # Input pipeline (bcnn_b200_train_step upload_inputs=2): every timed step uploads one batch
    # (inputs + labels) from the pinned host mirrors on the copy stream while it computes on the
    # batch staged by the previous step, and reads its loss back. The prologue stages batch 0.
    net.prefetch_inputs()
    for _ in range(args.warmup):  # the pipelined mode has its own warm-up: it alternates two input
        net.train_step(upload_inputs=2, fetch_loss=True)  # buffers, each with its own step graph
    barrier()
    lib.bcnn_b200_event_record(e0, stream)
    loss = 0.0
    for _ in range(args.steps):
        loss = net.train_step(upload_inputs=2, fetch_loss=True)
    net.sync()   # the last step's upload is part of the region
    lib.bcnn_b200_event_record(e1, stream)
    barrier()
    ms_e2e = lib.bcnn_b200_event_elapsed_ms(e0, e1)
    clocks = sampler.stop() if rank == 0 else None

    if dist:
        import torch
        t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])

    # per-node breakdown of one step (CUDA events around every node). Every rank runs it: the
    # step carries the gradient all-reduce, which all ranks must enter.
    lib.bcnn_b200_profile(net.handle, 1)
    net.train_step()
    barrier()
    result = None
    if rank == 0:
        peaks = measured_peaks()
        imgs = args.batch * world * args.steps
        value = imgs / (ms_total * 1e-3)
        e2e = imgs / (ms_e2e * 1e-3)
        by_type = {}
        names = {0: "conv(+bn+act)", 2: "depthwise", 3: "activation", 4: "fullc", 5: "maxpool",
                 6: "avgpool", 7: "softmax", 9: "batchnorm", 12: "eltwise", 16: "cost"}
        fwd, bwd = C.c_float(), C.c_float()
        for i in range(net.num_nodes()):
            lib.bcnn_b200_profile_node_ms(net.handle, i, C.byref(fwd), C.byref(bwd))
            k = names.get(net.node_type(i), str(net.node_type(i)))
            a = by_type.setdefault(k, [0.0, 0.0])
            a[0] += fwd.value; a[1] += bwd.value
        lib.bcnn_b200_profile(net.handle, 0)
        breakdown = {k: dict(fwd_ms=round(v[0], 3), bwd_ms=round(v[1], 3)) for k, v in by_type.items()}
        dp_bytes = lib.bcnn_b200_dp_bytes_per_step(net.handle)
        net.close()
        roofs = kernel_rooflines(lib, None, math, peaks, args.batch) if args.rooflines else []
        # headline: conv_tma_fwd_kernel (fprop + dgrad) has the largest share of the step in the
        # ncu launch list (profiles/); its heaviest ResNet-50 shape is the 1x1 64->256 @56 fprop
        dominant = next((r for r in roofs if r["kernel"].startswith("conv_fprop 1x1 64->256")), None)
        cpu = None
        if world == 1 and args.cpu_baseline:
            r = time_reference(args.workload, args.res, steps=1, warmup=0, budget_s=20.0)
            if r:
                cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="reference",
                           sample=r["sample"])
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (tcgen05 kind::tf32 on fp32 tensors, fp32 accumulate in TMEM; fp32 elsewhere)"
                     if math == capi.MATH_TC else "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload} {args.res}x{args.res} training "
                                   f"(fwd+bwd+SGD) via the bcnn C API",
                       "per_gpu_batch": args.batch, "global_batch": args.batch * world,
                       "parallelism": f"dp{world}", "conv_math": args.math,
                       "residual_semantics": "batch-correct (reference_quirks off)",
                       "l2_policy": "inputs larger than L2 (activations >> 126 MB per step)"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "loss": loss,
            "step_breakdown_ms": breakdown,
            "allreduce_bytes_per_step": int(dp_bytes),
            "peaks": peaks,
        }
        if dominant:
            result["roofline"] = {k: dominant[k] for k in ("bound", "achieved", "peak", "unit",
                                                           "frac", "traffic")}
            result["roofline"]["kernel"] = dominant["kernel"]
            result["rooflines"] = roofs
        if cpu:
            result["cpu_baseline"] = cpu
        emit(json.dumps(result))
    else:
        net.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class JsonOnlyStdout:
    """Everything libraries print to file descriptor 1 (NCCL's version banner, ...) goes to stderr;
    the JSON line is the only thing written to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        os.write(self.real, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)


OUT = None


def emit(line: str):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="resnet50", choices=["resnet50", "cifar", "mnist",
                                                                "yolo_tiny", "mobilenet"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--math", default=os.environ.get("BCNN_B200_BENCH_MATH", "tc"),
                    choices=["resident", "tc", "fp32"])
    ap.add_argument("--no-rooflines", dest="rooflines", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    # a wedged collective must not hold the box: dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BCNN_B200_BENCH_WATCHDOG_S", "1500")), exit=True)
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    global OUT
    with JsonOnlyStdout() as OUT:
        if args.impl == "reference":
            return run_reference_arm(args)
        return run_own_arm(args)


if __name__ == "__main__":
    sys.exit(main())
