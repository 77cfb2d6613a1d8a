"""Kernel-class rooflines without the whole-net run.  python tools/rooflines.py [batch] [filter]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from bcnn_b200 import capi
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flt = sys.argv[2] if len(sys.argv) > 2 else ""
lib = capi.b200()
peaks = bench.measured_peaks()
for r in bench.kernel_rooflines(lib, None, capi.MATH_TC, peaks, batch):
    if flt in r["kernel"]:
        print(f'{r["kernel"]:45s} {r["achieved"]:9.1f} {r["unit"]:8s} frac {r["frac"]:.3f}  {r["ms_per_launch"]:.4f} ms')
