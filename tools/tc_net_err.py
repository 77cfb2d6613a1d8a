"""Debug: per-tensor error of the tensor-core path vs the golden, in graph order."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import netcases
from bcnn_b200 import capi
from helpers import GOLDEN, rel_err

name = sys.argv[1] if len(sys.argv) > 1 else "resnet_small_b4"
golden = dict(np.load(GOLDEN / f"{name}.npz"))
for math in (capi.MATH_FP32, capi.MATH_TC):
    net = capi.Net(); net.set_conv_math(math)
    out = netcases.run_case(net, name); net.close()
    print("==== math", math)
    for key in golden:
        if key.startswith("sens:") or "/argmax/" in key or key.endswith("/cost"): continue
        base = key[:-4] if key.endswith("@sub") else key
        got, want = out[base], golden[key]
        if key.endswith("@sub"):
            idx = np.linspace(0, got.size - 1, want.size).astype(np.int64); got = got.ravel()[idx]
        if np.abs(want).max(initial=0) == 0: continue
        e = rel_err(got, want)
        d = np.abs(got.astype(np.float64).ravel() - want.astype(np.float64).ravel())
        out_frac = float(np.mean(d > 2e-2 * np.abs(want).max()))
        print(f"{base:50s} max {e[0]:.3e} l2 {e[1]:.3e} outliers {out_frac:.2e} sens {float(golden.get('sens:'+base, 0)):.2e}")
    for k in ("s0/grad/77:s1b1_out", "s0/grad/73:s1b1_c", "s0/data/77:s1b1_out"):
        g, w = out[k].ravel(), golden[k].ravel() if k in golden else golden[k+"@sub"].ravel()
        print(k, g.shape, w.shape, "got", g[:8], "want", w[:8], "absmax", np.abs(g).max(), np.abs(w).max(),
              "nz", np.count_nonzero(g), np.count_nonzero(w))
