"""Per-tensor error of both conv math modes against the LIVE reference (oracle/_ref), in graph
order, with the reference's own 1-ulp response beside it.  python tools/net_err.py CASE [seed]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import netcases
from bcnn_b200 import capi
from helpers import rel_err

name = sys.argv[1] if len(sys.argv) > 1 else "cifar_b128"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 31
want = netcases.live_reference_case(name, seed)
for math in (capi.MATH_FP32, capi.MATH_TC):
    net = capi.Net(); net.set_conv_math(math); net.set_reference_quirks(True)
    out = netcases.run_case(net, name, seed=seed); net.close()
    print("==== math", "tc" if math else "fp32")
    for key in want:
        if key.startswith("sens:") or "/argmax/" in key: continue
        if np.abs(want[key]).max(initial=0) == 0: continue
        e = rel_err(out[key], want[key])
        print(f"{key:44s} max {e[0]:.3e} l2 {e[1]:.3e}  ref-1ulp {float(want.get('sens:' + key, 0)):.2e}")
    for key in want:
        if "/argmax/" in key:
            print(f"{key:44s} mismatch rate {np.mean(out[key] != want[key]):.3e}")
