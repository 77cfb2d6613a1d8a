"""Timing decomposition of conv_tma_fwd_kernel<1,1> (resident fprop with fused statistics) through the
BCNN_B200_DBG_EPI bit mask (results are garbage, only the time matters):
  1 no statistics   8 no staging / bulk store   16 no MMAs   32 no TMA loads   64 (with 16) plain mbarrier arrive instead of tcgen05.commit per k-block
    python tools/epi_decomp.py [batch]
"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from bcnn_b200 import capi

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = capi.b200()
SHAPES = [(64, 56, 256, 1, 1, 0), (64, 56, 64, 3, 1, 1), (128, 28, 512, 1, 1, 0), (256, 14, 256, 3, 1, 1),
          (256, 14, 1024, 1, 1, 0), (1024, 14, 256, 1, 1, 0), (512, 7, 512, 3, 1, 1)]
VARIANTS = [(0, "full"), (1, "no stats"), (8, "no store"), (9, "no store, no stats"), (16, "no mma"), (32, "no loads"), (48, "epilogue only"),
            (25, "loads only"), (41, "mma only"), (57, "nothing"), (121, "nothing, arrive")]


def timeit(fn, iters=5):
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


print("shape".ljust(26) + "".join(f"{name[:12]:>13}" for _, name in VARIANTS))
for (cin, h, cout, k, s, pad) in SHAPES:
    d = capi.ConvDesc.make(batch, cin, h, h, cout, k, s, pad, 1)
    ws_bytes = lib.bcnn_b200_conv_nhwc_workspace_bytes(d)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 256))
    ex, ey, ew = batch * cin * h * h, batch * cout * d.ho * d.wo, cout * cin * k * k
    x, y = capi.DeviceBuffer(nbytes=ex * 2), capi.DeviceBuffer(nbytes=ey * 2)
    w = capi.DeviceBuffer(np.zeros(ew, np.float32))
    st = [capi.DeviceBuffer(np.ones(cout, np.float32)) for _ in range(4)]
    sc1 = capi.DeviceBuffer(nbytes=4 * lib.bcnn_b200_nhwc_scratch_floats(cout))
    sc2 = capi.DeviceBuffer(nbytes=4 * lib.bcnn_b200_bn_scratch_floats(cout))
    fn = lambda: lib.bcnn_b200_conv_forward_bn_stats_nhwc(
        d, x.ptr, w.ptr, y.ptr, ws.ptr, ws_bytes, None, st[0].ptr, st[1].ptr, st[2].ptr, st[3].ptr, sc1.ptr,
        sc2.ptr, None)
    row = f"{k}x{k} {cin}->{cout} @{h}".ljust(26)
    for bits, _ in VARIANTS:
        os.environ["BCNN_B200_DBG_EPI"] = str(bits)
        row += f"{timeit(fn) * 1e3:13.1f}"
    os.environ["BCNN_B200_DBG_EPI"] = "0"
    print(row + "  us", flush=True)
    for b in (ws, x, y, w, sc1, sc2, *st):
        b.free()
