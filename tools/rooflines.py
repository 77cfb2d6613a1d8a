"""Kernel-class rooflines without the whole-net run.  python tools/rooflines.py [batch] [filter]
Only the entries whose name contains `filter` are run (tools/ncu_rooflines.sh captures them one
at a time under ncu)."""
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
for r in bench.kernel_rooflines(lib, None, capi.MATH_TC, peaks, batch, only=flt):
    print(f'{r["kernel"]:50s} {r["bound"]:6s} {r["achieved"]:9.1f} {r["unit"]:8s} frac {r["frac"]:.3f}  '
          f'{r["ms_per_launch"]:.4f} ms' + (f'  tc_frac {r["tc_frac"]:.3f}' if "tc_frac" in r else ""))
