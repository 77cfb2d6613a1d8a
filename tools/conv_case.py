"""Debug/bench one convolution shape: tensor-core path vs the FP32 SIMT path on the GPU.
    python tools/conv_case.py batch cin h w cout k stride pad [ops=fdw] [iters]
Prints per-op normalised error (max-abs, L2) and the CUDA-event time per launch."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from bcnn_b200 import capi
from helpers import dev, dev_zeros, rel_err, check, f32

a = sys.argv[1:]
batch, cin, h, w, cout, k, s, pad = [int(v) for v in a[:8]]
ops = a[8] if len(a) > 8 else "fdw"
iters = int(a[9]) if len(a) > 9 else 0
lib = capi.b200()
d = capi.ConvDesc.make(batch, cin, h, w, cout, k, s, pad, 1)
r = np.random.default_rng(1)
x = f32(r.uniform(-1, 1, size=(batch, cin, h, w)))
wt = f32(r.uniform(-1, 1, size=(cout, cin, k, k)) * np.sqrt(3.0 / (cin * k * k)))
dy = f32(r.uniform(-1, 1, size=(batch, cout, d.ho, d.wo)))
ws_bytes = max(lib.bcnn_b200_conv_workspace_bytes(d, capi.MATH_TC), lib.bcnn_b200_conv_workspace_bytes(d, capi.MATH_FP32))
ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
dxv, dwt, ddy = dev(x), dev(wt), dev(dy)
flops = 2.0 * batch * cout * d.ho * d.wo * cin * k * k

def sync():
    check(lib.bcnn_b200_stream_sync(None))

def timeit(fn):
    if not iters: return float("nan")
    e0, e1 = lib.bcnn_b200_event_create(), lib.bcnn_b200_event_create()
    for _ in range(3): fn()
    sync()
    lib.bcnn_b200_event_record(e0, None)
    for _ in range(iters): fn()
    lib.bcnn_b200_event_record(e1, None)
    sync()
    return lib.bcnn_b200_event_elapsed_ms(e0, e1) / iters

def run(name, call, out_shape, n_out):
    res = {}
    for math in (capi.MATH_FP32, capi.MATH_TC):
        out = dev_zeros(n_out)
        check(call(math, out)); sync()
        res[math] = out.download(np.float32, out_shape)
        if math == capi.MATH_TC:
            ms = timeit(lambda: call(math, out))
        out.free()
    e = rel_err(res[capi.MATH_TC], res[capi.MATH_FP32])
    print(f"{name}: uses_tc={lib.bcnn_b200_conv_uses_tensor_cores(d, 'fdw'.index(name[0]))} err max {e[0]:.3e} l2 {e[1]:.3e}"
          f"  {ms:.4f} ms  {flops / (ms * 1e-3) / 1e12 if ms == ms else 0:.1f} TFLOP/s", flush=True)
    if e[0] > 2e-2:
        g = res[capi.MATH_TC]
        print("   stats: zeros", int(np.sum(g == 0)), "of", g.size, "nan", int(np.sum(~np.isfinite(g))), "absmax", float(np.abs(g[np.isfinite(g)]).max(initial=0)))
        print("   got[0,0,0,:8]", g[0, 0].ravel()[:8], "want", res[capi.MATH_FP32][0, 0].ravel()[:8])
        diff = np.abs(res[capi.MATH_TC] - res[capi.MATH_FP32])
        idx = np.unravel_index(np.argmax(diff), diff.shape)
        print("   worst at", idx, "got", res[capi.MATH_TC][idx], "want", res[capi.MATH_FP32][idx],
              "bad frac", float(np.mean(diff > 1e-2 * np.abs(res[capi.MATH_FP32]).max())))

if "f" in ops:
    run("fprop", lambda m, o: lib.bcnn_b200_conv_forward(d, dxv.ptr, dwt.ptr, None, 0, o.ptr, ws.ptr, ws_bytes, m, None),
        dy.shape, dy.size)
if "d" in ops:
    run("dgrad", lambda m, o: lib.bcnn_b200_conv_backward_data(d, dwt.ptr, ddy.ptr, o.ptr, 0, ws.ptr, ws_bytes, m, None),
        x.shape, x.size)
if "w" in ops:
    run("wgrad", lambda m, o: lib.bcnn_b200_conv_backward_weights(d, dxv.ptr, ddy.ptr, o.ptr, ws.ptr, ws_bytes, m, None),
        wt.shape, wt.size)
