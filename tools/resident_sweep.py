"""Per-layer timing of the resident (BF16 NHWC) convolution passes on every ResNet-50 shape (or one shape):
CUDA-event time per C-ABI call (weight pack + main kernel + split-K reduction), algorithmic FLOPs and
bytes, and the fraction of the roof that bounds the shape (max of FLOPs / BF16 peak and bytes / HBM peak).
    python tools/resident_sweep.py [batch] [iters] [only "cin,h,cout,k,s,pad"] [passes fsdw]
passes: f fprop, s fprop + fused batch-norm statistics, d dgrad, w wgrad.
Buffers are random bits (timing only); correctness lives in tests/test_nhwc_bf16_gpu.py.
"""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from bcnn_b200 import capi
from bcnn_b200.configs import RESNET50_CONV_SHAPES as RESNET50

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
only = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] else ""
passes = sys.argv[4] if len(sys.argv) > 4 else "sdw"
shapes = RESNET50
if only:
    v = [int(t) for t in only.split(",")]
    shapes = [tuple(v) + (1,)]
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0) * 1e9
TC = peaks.get("bf16_tflops", 1590.0) * 1e12
lib = capi.b200()


def timeit(fn):
    e0, e1 = lib.bcnn_b200_event_create(), lib.bcnn_b200_event_create()
    for _ in range(2):
        err = fn()
        if err:
            raise RuntimeError(lib.bcnn_b200_error_string(err).decode())
    lib.bcnn_b200_stream_sync(None)
    lib.bcnn_b200_event_record(e0, None)
    for _ in range(iters):
        fn()
    lib.bcnn_b200_event_record(e1, None)
    lib.bcnn_b200_stream_sync(None)
    ms = lib.bcnn_b200_event_elapsed_ms(e0, e1) / iters
    lib.bcnn_b200_event_destroy(e0)
    lib.bcnn_b200_event_destroy(e1)
    return ms


tot = {}
print(f"batch {batch}, peaks: HBM {HBM / 1e9:.0f} GB/s, BF16 {TC / 1e12:.0f} TFLOP/s (burst)")
print(f"{'shape':28} {'pass':6} {'ms':>8} {'TFLOP/s':>8} {'tc_frac':>8} {'GB/s':>8} {'hbm_frac':>8} {'roof':>6} {'frac':>6}")
for (cin, h, cout, k, s, pad, count) in shapes:
    d = capi.ConvDesc.make(batch, cin, h, h, cout, k, s, pad, 1)
    mask = lib.bcnn_b200_conv_nhwc_supported(d)
    ws_bytes = lib.bcnn_b200_conv_nhwc_workspace_bytes(d)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 256))
    ex, ey, ew = batch * cin * h * h, batch * cout * d.ho * d.wo, cout * cin * k * k
    thin = cin < 16
    x = capi.DeviceBuffer(nbytes=ex * (4 if thin else 2))
    y, dy, dx = (capi.DeviceBuffer(nbytes=n * 2) for n in (ey, ey, ex))
    w, gw = capi.DeviceBuffer(np.zeros(ew, np.float32)), capi.DeviceBuffer(nbytes=ew * 4)
    c = cout
    st = [capi.DeviceBuffer(np.ones(c, np.float32)) for _ in range(4)]
    sc1 = capi.DeviceBuffer(nbytes=4 * lib.bcnn_b200_nhwc_scratch_floats(c))
    sc2 = capi.DeviceBuffer(nbytes=4 * lib.bcnn_b200_bn_scratch_floats(c))
    keep = capi.DeviceBuffer(nbytes=max(lib.bcnn_b200_conv_nhwc_x_keep_bytes(d), 4))
    sh = capi.ConvShadows()
    if thin:
        sh.x, sh.x_bytes = keep.ptr, keep.nbytes
    shp = capi.C.byref(sh) if thin else None
    flops = 2.0 * ey * cin * k * k
    xb = ex * (4 if thin else 2)
    calls = {
        "f": ("fprop", 1, lambda: lib.bcnn_b200_conv_forward_nhwc(d, x.ptr, w.ptr, None, 0, y.ptr, ws.ptr, ws_bytes, shp, None),
              xb + ey * 2 + ew * 4),
        "s": ("fprop+st", 1, lambda: lib.bcnn_b200_conv_forward_bn_stats_nhwc(
            d, x.ptr, w.ptr, y.ptr, ws.ptr, ws_bytes, shp, st[0].ptr, st[1].ptr, st[2].ptr, st[3].ptr, sc1.ptr,
            sc2.ptr, None), xb + ey * 2 + ew * 4),
        "d": ("dgrad", 2, lambda: lib.bcnn_b200_conv_backward_data_nhwc(d, w.ptr, dy.ptr, dx.ptr, 0, ws.ptr, ws_bytes, None),
              ex * 2 + ey * 2 + ew * 4),
        "w": ("wgrad", 4, lambda: lib.bcnn_b200_conv_backward_weights_nhwc(d, x.ptr, dy.ptr, gw.ptr, ws.ptr, ws_bytes, shp, None),
              xb + ey * 2 + ew * 8),
    }
    for key in passes:
        name, bit, fn, nbytes = calls[key]
        if not (mask & bit) or (key == "d" and thin):
            continue
        ms = timeit(fn)
        tf, gb = flops / (ms * 1e-3) / 1e12, nbytes / (ms * 1e-3) / 1e9
        t_tc, t_hbm = flops / TC, nbytes / HBM
        roof = "tensor" if t_tc >= t_hbm else "hbm"
        frac = max(t_tc, t_hbm) / (ms * 1e-3)
        a = tot.setdefault(name, [0.0, 0.0, 0.0])
        a[0] += ms * count; a[1] += max(t_tc, t_hbm) * 1e3 * count; a[2] += flops * count
        print(f"{f'{k}x{k}/{s} {cin}->{cout} @{h} x{count}':28} {name:6} {ms:8.4f} {tf:8.1f} {tf * 1e12 / TC:8.3f} {gb:8.0f} "
              f"{gb * 1e9 / HBM:8.3f} {roof:>6} {frac:6.3f}", flush=True)
    for b in (ws, x, y, dy, dx, w, gw, sc1, sc2, keep, *st):
        b.free()
print()
all_ms = all_ideal = all_flops = 0.0
for name, (ms, ideal, fl) in tot.items():
    print(f"{name}: {ms:.3f} ms per step over the net, ideal {ideal:.3f} ms ({ideal / ms:.3f} of the mixed roof), "
          f"{fl / (ms * 1e-3) / 1e12:.0f} TFLOP/s = {fl / (ms * 1e-3) / TC:.3f} of BF16 burst peak")
    all_ms += ms; all_ideal += ideal; all_flops += fl
if tot:
    print(f"all passes: {all_ms:.3f} ms, ideal {all_ideal:.3f} ms ({all_ideal / all_ms:.3f}), conv_tc_util "
          f"{all_flops / (all_ms * 1e-3) / TC:.3f} of BF16 burst peak")
