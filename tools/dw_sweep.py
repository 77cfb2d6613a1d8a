"""Per-layer depthwise timings at MobileNet-v1 shapes: python tools/dw_sweep.py [batch]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from bcnn_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = capi.b200()
peaks = bench.measured_peaks()
shapes = [(32, 112, 1), (64, 112, 2), (128, 56, 1), (128, 56, 2), (256, 28, 1), (256, 28, 2), (512, 14, 1),
          (512, 14, 2), (1024, 7, 1)]
tot = [0.0, 0.0, 0.0, 0.0]
for c, h, s in shapes:
    ho = (h + 2 - 3) // s + 1
    ei, eo = n * c * h * h, n * c * ho * ho
    x, y, dy, dx = (capi.DeviceBuffer(nbytes=4 * e) for e in (ei, eo, eo, ei))
    w, b, gw = (capi.DeviceBuffer(nbytes=4 * c * 9) for _ in range(3))
    nscr = lib.bcnn_b200_depthwise_scratch_floats(n, c, 3)
    scr = capi.DeviceBuffer(nbytes=4 * nscr)
    f = bench.event_time_ms(lib, None, lambda: lib.bcnn_b200_depthwise_forward(
        x.ptr, w.ptr, b.ptr, 2, y.ptr, n, c, h, h, 3, s, 1, None), 5)
    bw = bench.event_time_ms(lib, None, lambda: lib.bcnn_b200_depthwise_backward(
        x.ptr, w.ptr, dy.ptr, gw.ptr, dx.ptr, n, c, h, h, 3, s, 1, scr.ptr, nscr, None), 5)
    fb, bb = 4 * (ei + eo), 4 * (2 * ei + 2 * eo + ei)   # bwd: wgrad reads x, dy; dgrad reads dy, RMW dx
    cnt = 5 if (c, h, s) == (512, 14, 1) else 1
    tot[0] += cnt * f; tot[1] += cnt * bw; tot[2] += cnt * fb / peaks["hbm"] / 1e6; tot[3] += cnt * bb / peaks["hbm"] / 1e6
    print(f"dw3x3 s{s} {c:4d} @{h:3d}: fwd {f:.4f} ms {fb / f / 1e6:7.0f} GB/s ({fb / f / 1e6 / peaks['hbm']:.2f})   "
          f"bwd {bw:.4f} ms {bb / bw / 1e6:7.0f} GB/s ({bb / bw / 1e6 / peaks['hbm']:.2f})", flush=True)
    for t in (x, y, dy, dx, w, b, gw, scr):
        t.free()
print(f"mobilenet total: fwd {tot[0]:.3f} ms (ideal {tot[2]:.3f})  bwd {tot[1]:.3f} ms (ideal {tot[3]:.3f})")
