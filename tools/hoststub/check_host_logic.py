"""Development check of the C host runtime without a GPU: drives tools/hoststub/libbcnn_hoststub.so
(real bcnn_b200/src/**/*.c, host-memory device stub) and the compiled reference side by side.
Covers what needs no kernels: weight files (save bytes, load, PREDICT fold, Darknet layout and
transpose, error statuses), the optimizer dispatch (SGD / Adam from injected gradients), config
files (graphs and error statuses in both dialects), bcnn_yolo_get_detections and the host
restatement of the yolo loss on the reference's own head tensors.
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

# --- learning-rate schedules (bcnn_set_learning_rate_policy): five SGD steps per policy ---
policies = {"step": (1, dict(scale=0.5, step=2)), "inv": (2, dict(gamma=0.3, power=0.75)),
            "exp": (3, dict(gamma=0.9)), "poly": (4, dict(power=2.0, max_batches=8)),
            "sigmoid": (5, dict(gamma=0.7, step=3)), "constant": (0, {})}
for name, (kind, kw) in policies.items():
    ours6, ref6 = nets(capi.MODE_TRAIN)
    for n in (ours6, ref6):
        n.lr_policy(kind, **kw)
        configs.init_params(n, seed=8)
    for step in range(5):
        rng = np.random.default_rng(300 + step)
        for idx, pname, shape in configs.param_tensors(ours6):
            if not ours6._tensor(idx).grad_data:
                continue
            g = rng.normal(0, 1e-2, size=shape).astype(np.float32)
            for n in (ours6, ref6):
                n.set(idx, n.get(idx, grad=True) + g, grad=True)
        ours6.update()
        ref6.update()
    same(params(ours6), params(ref6), f"lr policy {name}")
print("update: learning-rate policies (step / inv / exp / poly / sigmoid / constant) bit-identical")

# --- config files: same graph, same status, both dialects ---
import contextlib
import os

CFG = ROOT / "tests" / "golden" / "cfg"


@contextlib.contextmanager
def quiet():  # the reference's yolo code prints to stderr
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(2)
    os.dup2(devnull, 2)
    try:
        yield
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)


def pair(mode):
    return capi.Net(mode=mode, lib=stub), helpers.ref_net(mode=mode)


for mode, cfg, model in ((capi.MODE_TRAIN, "mini_bcnn.conf", None),
                         (capi.MODE_PREDICT, "mini_yolo.cfg", CFG / "mini_yolo.weights"),
                         (capi.MODE_TRAIN, "mini_yolo.cfg", CFG / "mini_yolo.weights"),
                         (capi.MODE_PREDICT, "mini_yolo.cfg", tmp / "absent.weights")):
    a, b = pair(mode)
    sa, sb = a.load_net(CFG / cfg, model), b.load_net(CFG / cfg, model)
    assert sa == sb and a.structure() == b.structure(), (cfg, mode, sa, sb)
    if sa == 0 and model is not None:  # without a model the weights are each library's rand() draws
        same(params(a), params(b), f"{cfg} parameters after load_net")
print("load_net: graphs, statuses and loaded parameters identical to the reference (bcnn + Darknet)")

# --- detections and yolo loss on the reference's own head tensors ---
g = dict(np.load(CFG / "mini_yolo_train.npz"))
heads = dict(np.load(CFG / "mini_yolo.npz"))
a, b = pair(capi.MODE_PREDICT)
for n in (a, b):
    assert n.load_net(CFG / "mini_yolo.cfg", CFG / "mini_yolo.weights") == 0
    n.compile()
    n.set("lid10", heads["lid10"])
    n.set("lid17", heads["lid17"])
for thresh, (w, h), rel in ((0.5, (640, 480), True), (0.3, (300, 500), False), (-1.0, (64, 48), True),
                            (0.9999, (10, 10), True)):
    for sample in range(2):
        with quiet():
            da = a.yolo_detections(sample, w, h, thresh, rel)
            db = b.yolo_detections(sample, w, h, thresh, rel)
        assert da.shape == db.shape and np.array_equal(da.view(np.uint32), db.view(np.uint32))
print("yolo detections: boxes, scores and NMS order bit-identical to the reference")

a = capi.Net(mode=capi.MODE_TRAIN, lib=stub)
assert a.load_net(CFG / "mini_yolo.cfg", CFG / "mini_yolo.weights") == 0
a.compile()
for label in (g["label"], configs.synth_yolo_labels(2, [50, 0], classes=2, seed=5)):
    b = helpers.ref_net(mode=capi.MODE_TRAIN)
    assert b.load_net(CFG / "mini_yolo.cfg", CFG / "mini_yolo.weights") == 0
    b.compile()
    b.set("input", g["input"])
    b.set("label", label)
    with quiet():
        b.forward()
    a.set("label", label)
    nodes = [i for i, (kind, _, _) in enumerate(a.structure()["nodes"]) if kind == 14]
    for node, name in zip(nodes, ("lid10", "lid17")):
        a.set(name, b.get(name))
        cost = stub.bcnn_b200_yolo_loss_on_host(a.handle, node)
        want = b.get(name, grad=True)
        got = a.get(name, grad=True)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
        assert abs(cost - float((want.astype(np.float64) ** 2).sum())) <= 1e-5 * cost
print("yolo loss (host loops): gradient bit-identical to the reference")
print("OK")
