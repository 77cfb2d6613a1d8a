"""Run a few training steps of a workload on cuda:0 for ncu (no timing printed is a bench value).
    python tools/profile_step.py [workload] [batch] [steps] [math]
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bcnn_b200 import capi, configs

workload = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
math = sys.argv[4] if len(sys.argv) > 4 else "resident"
lib = capi.b200()
net = capi.Net(mode=capi.MODE_TRAIN)
net.set_conv_math({"tc": capi.MATH_TC, "fp32": capi.MATH_FP32, "resident": capi.MATH_TC_BF16}[math])
net.set_reference_quirks(False)
if workload in ("mnist", "cifar"):
    configs.BUILDERS[workload](net, batch=batch)
else:
    configs.BUILDERS[workload](net, batch=batch, res=224 if workload != "yolo_tiny" else 416)
net.compile()
configs.init_params(net, seed=2024)
net.set("input", configs.synth_input(net.shape("input")))
net.set("label", configs.synth_labels(net.shape("label")))
l0 = lib.bcnn_b200_launch_count()
for s in range(steps):
    net.train_step()
    net.sync()
    print(f"step {s}: launches so far {lib.bcnn_b200_launch_count() - l0}", flush=True)
net.close()
