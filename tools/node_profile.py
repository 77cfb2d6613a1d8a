"""Per-node CUDA-event times of one training step (forward / backward) with each node's shapes and the
time its algorithmic work would take at the measured peaks: where a step's time goes, node by node.
    python tools/node_profile.py [workload] [batch] [math]
A number printed here is a diagnostic, not a bench value (events around every node serialise the step).
"""
import ctypes as C
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bcnn_b200 import capi, configs

workload = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
math = sys.argv[3] if len(sys.argv) > 3 else "resident"
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0) * 1e9
TC = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)) * 1e12
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
for _ in range(3):
    net.train_step()
net.sync()
lib.bcnn_b200_profile(net.handle, 1)
acc = {}
REPS = 3
for _ in range(REPS):
    net.train_step()
    net.sync()
    fwd, bwd = C.c_float(), C.c_float()
    for i in range(net.num_nodes()):
        lib.bcnn_b200_profile_node_ms(net.handle, i, C.byref(fwd), C.byref(bwd))
        a = acc.setdefault(i, [0.0, 0.0])
        a[0] += fwd.value / REPS
        a[1] += bwd.value / REPS
es = 2 if math == "resident" else 4
names = {0: "conv", 2: "dw", 3: "act", 4: "fc", 5: "maxpool", 6: "avgpool", 7: "softmax", 9: "bn", 12: "eltwise", 16: "cost"}
tot = [0.0, 0.0, 0.0, 0.0]
print(f"{'node':>4} {'type':8} {'src -> dst':44} {'fwd ms':>8} {'ideal':>7} {'bwd ms':>8} {'ideal':>7}")
for i in range(net.num_nodes()):
    t = net.node_type(i)
    s = net.tensor_shape(lib.bcnn_b200_node_src(net.handle, i, 0))
    d = net.tensor_shape(lib.bcnn_b200_node_dst(net.handle, i, 0))
    ein, eout = s[0] * s[1] * s[2] * s[3], d[0] * d[1] * d[2] * d[3]
    fi = bi = 0.0
    desc = f"{s[1]}x{s[2]}x{s[3]} -> {d[1]}x{d[2]}x{d[3]}"
    if t == 0:
        w = net.tensor_shape(lib.bcnn_b200_node_src(net.handle, i, 1))
        k = w[2]
        desc += f" k{k}"
        flops = 2.0 * eout * w[1] * k * k
        # fprop: read x, write raw, (bn apply: read raw, write y); bwd: bn-bwd 5 passes, wgrad reads x, dy; dgrad reads dy writes dx
        fi = max(flops / TC, (ein + eout) * es / HBM) + 2 * eout * es / HBM
        bi = 5 * eout * es / HBM + max(flops / TC, (ein + eout) * es / HBM) * 2
    elif t == 12:
        fi = 3 * eout * es / HBM
        bi = 4 * eout * es / HBM
    elif t == 5:
        fi = (ein * es + eout * (es + 4)) / HBM
        bi = (ein * es + eout * (es + 4)) / HBM
    f, b = acc[i]
    tot[0] += f; tot[1] += fi * 1e3; tot[2] += b; tot[3] += bi * 1e3
    print(f"{i:>4} {names.get(t, str(t)):8} {desc:44} {f:8.3f} {fi * 1e3:7.3f} {b:8.3f} {bi * 1e3:7.3f}")
print(f"total fwd {tot[0]:.2f} ms (ideal {tot[1]:.2f}), bwd {tot[2]:.2f} ms (ideal {tot[3]:.2f})")
net.close()
