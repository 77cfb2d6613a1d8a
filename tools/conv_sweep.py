"""Per-layer convolution sweep: every distinct ResNet-50 conv shape x (fprop, dgrad, wgrad), timed
with CUDA events, against both roofs (algorithmic bytes / measured HBM copy rate, FLOPs / TF32 rate).
    python tools/conv_sweep.py [batch] [iters] [filter]
Columns: count = how many layers of the net have the shape; ms; TFLOP/s; t_hbm, t_tc = ideal times;
frac = max(t_hbm, t_tc) / ms (fraction of the binding roof); step_ms = count * ms."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from bcnn_b200 import capi

# (cin, h, cout, k, stride, pad, count) of ResNet-50 v1.5 at 224
SHAPES = [
    (3, 224, 64, 7, 2, 3, 1),
    (64, 56, 64, 1, 1, 0, 1), (64, 56, 64, 3, 1, 1, 3), (64, 56, 256, 1, 1, 0, 4),
    (256, 56, 64, 1, 1, 0, 2),
    (256, 56, 128, 1, 1, 0, 1), (128, 56, 128, 3, 2, 1, 1), (128, 28, 512, 1, 1, 0, 4),
    (256, 56, 512, 1, 2, 0, 1), (512, 28, 128, 1, 1, 0, 3), (128, 28, 128, 3, 1, 1, 3),
    (512, 28, 256, 1, 1, 0, 1), (256, 28, 256, 3, 2, 1, 1), (256, 14, 1024, 1, 1, 0, 6),
    (512, 28, 1024, 1, 2, 0, 1), (1024, 14, 256, 1, 1, 0, 5), (256, 14, 256, 3, 1, 1, 5),
    (1024, 14, 512, 1, 1, 0, 1), (512, 14, 512, 3, 2, 1, 1), (512, 7, 2048, 1, 1, 0, 3),
    (1024, 14, 2048, 1, 2, 0, 1), (2048, 7, 512, 1, 1, 0, 2), (512, 7, 512, 3, 1, 1, 2),
]


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    flt = sys.argv[3] if len(sys.argv) > 3 else ""
    lib = capi.b200()
    peaks = bench.measured_peaks()
    hbm = peaks["hbm"] * 1e9
    tc = peaks["tc_burst"] * 1e12 / 2  # TF32 runs at half the BF16 rate
    math = capi.MATH_TC
    tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    ideal = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    print(f"batch {batch}  hbm {peaks['hbm']:.0f} GB/s  tf32 {tc / 1e12:.0f} TFLOP/s")
    print(f"{'shape':32s} {'pass':6s} {'cnt':>3s} {'ms':>8s} {'TF/s':>7s} {'t_hbm':>7s} {'t_tc':>7s} {'frac':>5s} {'step_ms':>8s}")
    for (cin, hh, cout, k, s, pad, count) in SHAPES:
        name = f"{k}x{k}s{s} {cin}->{cout} @{hh}"
        if flt and flt not in name:
            continue
        d = capi.ConvDesc.make(batch, cin, hh, hh, cout, k, s, pad, 1)
        ws_bytes = lib.bcnn_b200_conv_workspace_bytes(d, math)
        ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
        ex, ey, ew = batch * cin * hh * hh, batch * cout * d.ho * d.wo, cout * cin * k * k
        cx, cw, cy, cgw = (capi.DeviceBuffer(nbytes=4 * e) for e in (ex, ew, ey, ew))
        flops = 2.0 * batch * cout * d.ho * d.wo * cin * k * k
        calls = [
            ("fprop", 4.0 * (ex + ey + ew), lambda: lib.bcnn_b200_conv_forward(
                d, cx.ptr, cw.ptr, None, 0, cy.ptr, ws.ptr, ws_bytes, math, None)),
            ("dgrad", 4.0 * (ex + ey + ew), lambda: lib.bcnn_b200_conv_backward_data(
                d, cw.ptr, cy.ptr, cx.ptr, 0, ws.ptr, ws_bytes, math, None)),
            ("wgrad", 4.0 * (ex + ey + 2 * ew), lambda: lib.bcnn_b200_conv_backward_weights(
                d, cx.ptr, cy.ptr, cgw.ptr, ws.ptr, ws_bytes, math, None)),
        ]
        for pname, nbytes, call in calls:
            if pname == "dgrad" and cin == 3:
                continue
            ms = bench.event_time_ms(lib, None, call, iters)
            t_hbm, t_tc = nbytes / hbm * 1e3, flops / tc * 1e3
            frac = max(t_hbm, t_tc) / ms
            tot[pname] += count * ms
            ideal[pname] += count * max(t_hbm, t_tc)
            print(f"{name:32s} {pname:6s} {count:3d} {ms:8.4f} {flops / ms / 1e9:7.1f} {t_hbm:7.4f} "
                  f"{t_tc:7.4f} {frac:5.2f} {count * ms:8.3f}", flush=True)
        for b in (ws, cx, cw, cy, cgw):
            b.free()
    for k_ in tot:
        print(f"total {k_}: {tot[k_]:.3f} ms per step (ideal {ideal[k_]:.3f} ms)")


if __name__ == "__main__":
    main()
