"""Experiment: does a K-major SWIZZLE_128B UMMA descriptor work when it starts a non-multiple-of-8 rows
into the tile (with / without the matrix-base-offset field)? BCNN_B200_DBG_ROWSHIFT=1|2 makes the 1x1
resident fprop load its A tile one position early and start the descriptor one row late: outputs must
equal the plain run except for the last row of every 128-position tile."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from bcnn_b200 import capi
from helpers import dev, dev_zeros, f32, check
from test_nhwc_bf16_gpu import nhwc_bits, rounded

lib = capi.b200()
batch, cin, h, cout = 4, 64, 16, 64
d = capi.ConvDesc.make(batch, cin, h, h, cout, 1, 1, 0, 1)
r = np.random.default_rng(0)
x = rounded(f32(r.uniform(-1, 1, (batch, cin, h, h))))
w = f32(r.uniform(-1, 1, (cout, cin, 1, 1)))
ws_bytes = lib.bcnn_b200_conv_nhwc_workspace_bytes(d)
ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 256))
dx, dw = dev(nhwc_bits(x)), dev(w)
def run():
    y = dev_zeros(batch * cout * h * h, 2)
    check(lib.bcnn_b200_conv_forward_nhwc(d, dx.ptr, dw.ptr, None, 0, y.ptr, ws.ptr, ws_bytes, None, None))
    return y.download(np.uint16).reshape(-1, cout)
base = run()
for mode in ("1", "2"):
    os.environ["BCNN_B200_DBG_ROWSHIFT"] = mode
    got = run()
    rows = np.arange(got.shape[0])
    keep = (rows % 128) != 127
    same = np.array_equal(got[keep], base[keep])
    frac = float(np.mean(got[keep] == base[keep]))
    print(f"mode {mode} ({'with' if mode == '1' else 'without'} base offset): identical={same} match fraction {frac:.4f}")
