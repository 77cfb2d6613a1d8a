"""Debug: per-tensor errors of the YOLOv3-tiny 416 b8 TRAIN step vs the live reference."""
import sys, tempfile, re
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import test_baseline_parity_gpu as T
from bcnn_b200 import capi, configs
from helpers import ref_net, rel_err
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
quirks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tmp = T._yolo_files(tempfile.mkdtemp())
(tmp / "tiny_bN.cfg").write_text(configs.yolov3_tiny_cfg(batch=batch))
x = configs.synth_input((batch, 3, 416, 416), seed=25)
label = configs.synth_yolo_labels(batch, 3, classes=80, seed=26)
nets = []
for make in (lambda: ref_net(mode=capi.MODE_TRAIN, threads=8), lambda: capi.Net(mode=capi.MODE_TRAIN)):
    net = make()
    assert net.load_net(tmp / "tiny_bN.cfg", tmp / "tiny.weights") == 0
    if net.flavour == "b200":
        net.set_reference_quirks(bool(quirks)); net.set_conv_math(capi.MATH_FP32)
    net.compile(); net.set("input", x); net.set("label", label); net.forward(); net.backward()
    nets.append(net)
ref, net = nets
names = [t[0] for t in ref.structure()["tensors"] if re.fullmatch(r"lid\d+", t[0])]
for name in names:
    g = ref.get(name, grad=True)
    print(f"{name:8s} data {rel_err(net.get(name), ref.get(name))[1]:.2e}  grad {rel_err(net.get(name, grad=True), g)[1]:.2e} |g|max {np.abs(g).max():.3e} ours {np.abs(net.get(name, grad=True)).max():.3e}")
print("loss", ref.lib.bcnn_get_batch_size(ref.handle), net.loss())
