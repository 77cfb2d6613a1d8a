"""Inference images/s of the BASELINE inference configs (SURVEY.md 8d metric i): MobileNet-v1 224
(C3) and YOLOv3-tiny 416 with both yolo heads (C4), PREDICT mode, at batch 1 (as the reference's
examples run) and batched.   python tools/infer_bench.py [workload batch ...]

Per line: `value` = forward only, input resident in HBM (CUDA events on the net's stream);
`e2e` = bcnn_predict_on_batch: pinned-host upload of the batch + forward + read-back of the output
tensor every step; for yolo_tiny `e2e_detections` adds bcnn_yolo_get_detections (D2H of both heads,
box decode + NMS on the host) for every sample. Synthetic input, seeded weights. With --cpu the
compiled reference (oracle/_ref, TEST INFRASTRUCTURE, used here only as the timed baseline) runs
the same PREDICT net on the host cores: weights saved by a VALID-mode twin and read back with
bcnn_load_weights, which is what arms its Winograd / folded batch-norm inference path."""
import ctypes as C
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from bcnn_b200 import capi, configs

lib = capi.b200()
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]
args = [a for a in sys.argv[1:] if a != "--cpu"]
with_cpu = "--cpu" in sys.argv[1:]


def cpu_reference(workload, batch, res):
    import os
    import tempfile
    sys.path.insert(0, str(ROOT / "tests"))
    import helpers
    threads = os.cpu_count() or 1
    path = Path(tempfile.mkdtemp()) / "w.bcnnmodel"
    twin = helpers.ref_net(mode=capi.MODE_VALID, threads=threads)
    configs.BUILDERS[workload](twin, batch=batch, res=res)
    twin.compile()
    configs.init_params(twin, seed=2024)
    twin.save_weights(path)
    twin.close()
    ref = helpers.ref_net(mode=capi.MODE_PREDICT, threads=threads)
    configs.BUILDERS[workload](ref, batch=batch, res=res)
    ref.compile()
    assert ref.load_weights(path) == 0
    ref.set("input", configs.synth_input(ref.shape("input"), seed=12345))
    ref.forward()
    n, t0 = 0, time.perf_counter()
    while n < 3 or time.perf_counter() - t0 < 3.0:
        ref.forward()
        n += 1
    ms = (time.perf_counter() - t0) * 1e3 / n
    ref.close()
    return dict(value=batch / (ms * 1e-3), unit="images/s", ms_per_forward=ms, cores=threads,
                kind="reference", sample=f"{n} forwards at batch {batch}, PREDICT mode, AVX2+OpenMP")


jobs = [(args[i], int(args[i + 1])) for i in range(0, len(args), 2)] or [
    ("mobilenet", 1), ("mobilenet", 64), ("yolo_tiny", 1), ("yolo_tiny", 32)]
for workload, batch in jobs:
    net = capi.Net(mode=capi.MODE_PREDICT)
    net.set_conv_math(capi.MATH_TC)
    info = configs.BUILDERS[workload](net, batch=batch, res=416 if workload == "yolo_tiny" else 224)
    net.compile()
    configs.init_params(net, seed=2024)
    x = configs.synth_input(net.shape("input"), seed=12345)
    net.set("input", x)
    stream = lib.bcnn_b200_get_stream(net.handle)
    steps = 50 if batch == 1 else 20
    for _ in range(5):
        net.forward()
    net.sync()
    launches0 = lib.bcnn_b200_launch_count()
    ms = bench.event_time_ms(lib, stream, net.forward, steps)
    launches = (lib.bcnn_b200_launch_count() - launches0) // (steps + 1)
    net.set_host("input", x)
    net.predict_on_batch()
    t0 = time.perf_counter()
    for _ in range(steps):
        _, out = net.predict_on_batch()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    line = dict(metric="inference_images_per_sec", workload=workload, batch=batch, unit="images/s",
                value=batch / (ms * 1e-3), ms_per_forward=ms, e2e=batch / (e2e_ms * 1e-3),
                e2e_ms=e2e_ms, h2d_bytes_per_step=int(x.nbytes), d2h_bytes_per_step=int(out.nbytes),
                gpu_launches_per_forward=int(launches), cuda_graph=net.graphs() == 2,
                conv_math="tc", data="synthetic",
                output_finite=bool(np.isfinite(out).all()))
    if workload == "yolo_tiny":
        t0 = time.perf_counter()
        kept = 0
        for _ in range(steps):
            net.predict_on_batch()
            for b in range(batch):  # the C call itself; freeing is the caller's job
                count = C.c_int(0)
                dets = lib.bcnn_yolo_get_detections(net.handle, b, 640, 480, 416, 416, 0.6, 1,
                                                    C.byref(count))
                for k in range(count.value):
                    kept += dets[k].objectness > 0
                    libc.free(C.cast(dets[k].prob, C.c_void_p))
                if count.value:
                    libc.free(C.cast(dets, C.c_void_p))
        det_ms = (time.perf_counter() - t0) * 1e3 / steps
        line.update(e2e_detections=batch / (det_ms * 1e-3), e2e_detections_ms=det_ms,
                    boxes_kept_per_image=kept / (steps * batch))
    if with_cpu and batch == 1:
        line["cpu_baseline"] = cpu_reference(workload, batch, 416 if workload == "yolo_tiny" else 224)
    print(json.dumps(line), flush=True)
    net.close()
