"""Development check of the C host runtime without a GPU: drives tools/hoststub/libbcnn_hoststub.so
(real bcnn_b200/src/**/*.c, host-memory device stub) and the compiled reference side by side.
Covers what needs no kernels: weight files (save bytes, load, PREDICT fold, Darknet layout and
transpose, error statuses) and the optimizer dispatch (SGD / Adam from injected gradients).
Usage: tools/hoststub/build.sh && python tools/hoststub/check_host_logic.py"""
import ctypes as C
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
import netcases  # noqa: E402
from bcnn_b200 import capi, configs  # noqa: E402

stub = C.CDLL(str(Path(__file__).parent / "libbcnn_hoststub.so"), mode=C.RTLD_LOCAL)
capi.bind_bcnn_api(stub, capi.TensorB200)
capi.bind_b200_ext(stub)


def nets(mode, builder=netcases.model_io_net, pre=None):
    out = []
    for net in (capi.Net(mode=mode, lib=stub), helpers.ref_net(mode=mode)):
        if pre:
            pre(net)
        builder(net)
        net.compile()
        out.append(net)
    return out


def params(net):
    return {name: net.get(idx) for idx, name, _ in configs.param_tensors(net)}


def same(a, b, what):
    assert a.keys() == b.keys(), what
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), f"{what}: {k}"


tmp = Path(tempfile.mkdtemp())
# --- save: byte-identical files ---
ours, ref = nets(capi.MODE_VALID)  # VALID: init_params randomises the running statistics
for n in (ours, ref):
    configs.init_params(n, seed=11)
ours.save_weights(tmp / "ours.bcnnmodel")
ref.save_weights(tmp / "ref.bcnnmodel")
a, b = (tmp / "ours.bcnnmodel").read_bytes(), (tmp / "ref.bcnnmodel").read_bytes()
assert a == b, (len(a), len(b))
print("save: byte-identical,", len(a), "bytes")
want = params(ref)

# --- load, TRAIN and PREDICT (fold) ---
for mode in (capi.MODE_TRAIN, capi.MODE_PREDICT):
    ours2, ref2 = nets(mode)
    assert ours2.load_weights(tmp / "ref.bcnnmodel") == 0 and ref2.load_weights(tmp / "ref.bcnnmodel") == 0
    same(params(ours2), params(ref2), f"load mode {mode}")
    if mode == capi.MODE_TRAIN:
        same(params(ours2), want, "load TRAIN round trip")
print("load: TRAIN round trip and PREDICT fold bit-identical to the reference")

# --- Darknet layout ---
sys.path.insert(0, str(ROOT / "oracle"))
import bcnn_model_oracle as mo  # noqa: E402

for major, minor, tag in ((0, 2, "v02"), (0, 1, "v01"), (0, 1001, "transpose")):
    path = tmp / f"{tag}.weights"
    ours3, ref3 = nets(capi.MODE_PREDICT)
    layout = mo.net_layout(ours3)
    rng = np.random.default_rng(5)
    values = {name: rng.uniform(0.5, 1.5, size=size).astype(np.float32)
              for name, size in mo.all_names(layout)}
    mo.write_darknet(path, layout, values, major=major, minor=minor)
    assert ours3.load_weights(path) == 0 and ref3.load_weights(path) == 0
    same(params(ours3), params(ref3), f"darknet {tag}")
    got = {k: v.reshape(-1) for k, v in params(ours3).items()}
    same(got, {k: v for k, v in mo.read(path, layout, predict=True).items() if k in got},
         f"darknet {tag} vs numpy restatement")
print("load: Darknet files (u64 / i32 seen, transposed fc) bit-identical to the reference")

# --- error statuses ---
ours4, ref4 = nets(capi.MODE_TRAIN)
assert ours4.load_weights(tmp / "missing.bcnnmodel") == ref4.load_weights(tmp / "missing.bcnnmodel") == 1
(tmp / "bad.bcnnmodel").write_bytes(b"XXXX" + a[4:])
assert ours4.load_weights(tmp / "bad.bcnnmodel") == ref4.load_weights(tmp / "bad.bcnnmodel") == 3
(tmp / "x.onnx").write_bytes(a)
assert ours4.load_weights(tmp / "x.onnx") == ref4.load_weights(tmp / "x.onnx") == 3
(tmp / "short.bcnnmodel").write_bytes(a[: len(a) // 2])
assert ours4.load_weights(tmp / "short.bcnnmodel") == 3  # the reference returns 0 here (documented)
print("load: error statuses ok")

# --- optimizer dispatch: three updates from injected gradients ---
for opt in ("sgd", "adam"):
    pre = (lambda n: n.adam(0.002, 0.9, 0.999, 0.0005)) if opt == "adam" else None
    ours5, ref5 = nets(capi.MODE_TRAIN, pre=pre)
    for n in (ours5, ref5):
        configs.init_params(n, seed=3)
    worst = 0.0
    for step in range(3):
        rng = np.random.default_rng(100 + step)
        for idx, name, shape in configs.param_tensors(ours5):
            t = ours5._tensor(idx)
            if not t.grad_data:
                continue
            g = rng.normal(0, 1e-2, size=shape).astype(np.float32)
            if step == 1:
                g[...] = 0  # tiny second moments: the vdiv guard of the scalar tail
            for n in (ours5, ref5):
                cur = n.get(idx, grad=True)
                n.set(idx, cur + g, grad=True)
        ours5.update()
        ref5.update()
        po, pr = params(ours5), params(ref5)
        for k in po:
            d = np.abs(po[k].astype(np.float64) - pr[k]).max(initial=0)
            s = np.abs(pr[k]).max(initial=0) + 1e-30
            worst = max(worst, d / s)
            if opt == "sgd":
                assert np.array_equal(po[k], pr[k]), (opt, step, k)
    print(f"update[{opt}]: worst max-abs / max over 3 steps = {worst:.3e}")
    assert worst < 2e-6
print("OK")
