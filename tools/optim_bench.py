"""Optimizer kernels against the HBM roofline: python tools/optim_bench.py [elements ...]
SGD moves 16 B/element (read + write of w and g), Adam 32 B/element (w, g, m, v)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from bcnn_b200 import capi
lib = capi.b200()
peaks = bench.measured_peaks()
sizes = [int(a) for a in sys.argv[1:]] or [2_359_296, 25_557_032, 134_217_728]  # 3x3x512x512, ResNet-50, 512 MiB
for n in sizes:
    w, g, m, v = (capi.DeviceBuffer(nbytes=4 * n) for _ in range(4))
    t_sgd = bench.event_time_ms(lib, None, lambda: lib.bcnn_b200_sgd_update(
        w.ptr, g.ptr, n, 0.032, -1e-5, 0.9, None), 10)
    t_adam = bench.event_time_ms(lib, None, lambda: lib.bcnn_b200_adam_update(
        w.ptr, g.ptr, m.ptr, v.ptr, n, 0.032, 0.9, 0.999, -1e-5, None), 10)
    for name, t, b in (("sgd", t_sgd, 16 * n), ("adam", t_adam, 32 * n)):
        print(f"{name:5s} n={n:>11d}: {t:.4f} ms  {b / t / 1e6:7.0f} GB/s  "
              f"({b / t / 1e6 / peaks['hbm']:.2f} of measured HBM peak {peaks['hbm']:.0f})", flush=True)
    for t in (w, g, m, v):
        t.free()
