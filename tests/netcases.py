"""Whole-net parity cases shared by the golden generator (tests/golden/make_golden.py), the
live reference comparison and the GPU tests: build a BASELINE-style net through the bcnn C
API on either library, inject identical synthetic parameters / inputs / labels, run training
steps and collect every observable tensor."""
from __future__ import annotations

import numpy as np

from bcnn_b200 import capi, configs


def small_resnet(net, batch=4):
    """Three bottleneck blocks (projection and identity shortcuts) at 8x8: every ResNet-50
    layer kind, including the residual add, at oracle-friendly size. The stages keep stride 1:
    the reference mis-reads its operand for 1x1 convolutions with stride > 1
    (bcnn_conv_layer.c:445-446), so a strided projection shortcut has no reference answer;
    3x3 stride-2 convolutions are covered by chain_convnet."""
    return configs.resnet50(net, batch=batch, res=32, classes=10, widths=(8, 16), blocks=(1, 2),
                            stage_strides=(1, 1))


def chain_convnet(net, batch=4):
    """Chain-shaped net covering conv(+BN)(+act) variants, both pools, depthwise, avgpool."""
    net.set_input_shape(20, 20, 3, batch)
    net.conv(8, 3, 1, 1, 1, 1, "lrelu", "input", "c1")
    net.maxpool(2, 2, capi.PAD_SAME, "c1", "p1")
    net.conv(12, 3, 2, 1, 1, 0, "relu", "p1", "c2")
    net.depthwise(3, 1, 1, "relu", "c2", "dw")
    net.conv(16, 1, 1, 0, 2, 1, "none", "dw", "c3")
    net.maxpool(3, 2, capi.PAD_SAME, "c3", "p2")
    net.avgpool("p2", "gap")
    net.fullc(10, "none", "gap", "fc")
    net.softmax("fc", "softmax")
    net.cost("softmax", "cost")
    net.sgd(0.01, 0.9, 0.0005)
    return dict(classes=10, out="softmax")


def yolo_two_heads(net, batch=2):
    """YOLOv3-tiny's head topology at oracle-friendly size (examples/yolo/yolov3-tiny.cfg): a
    trunk of conv+BN+leaky / max-pool stages, a coarse head, and a second head fed by a 1x1
    conv -> upsample x2 -> concat with an earlier trunk tensor. The yolo loss itself stays out of
    scope: the fine head carries the euclidean cost, and the coarse branch
    reaches it through the upsample / concat route, so both glue layers see gradients."""
    net.set_input_shape(32, 32, 3, batch)
    net.conv(8, 3, 1, 1, 1, 1, "lrelu", "input", "c0")
    net.maxpool(2, 2, capi.PAD_SAME, "c0", "p0")          # 16x16
    net.conv(16, 3, 1, 1, 1, 1, "lrelu", "p0", "c1")      # route source (16 ch @16)
    net.maxpool(2, 2, capi.PAD_SAME, "c1", "p1")          # 8x8
    net.conv(32, 3, 1, 1, 1, 1, "lrelu", "p1", "c2")
    net.conv(16, 1, 1, 0, 1, 1, "lrelu", "c2", "c3")      # trunk end @8
    net.conv(8, 1, 1, 0, 1, 1, "lrelu", "c3", "r0")       # second head: 1x1 reduce
    net.upsample(2, "r0", "up")                           # 16x16
    net.concat(["up", "c1"], "cat")                       # 8 + 16 channels
    net.conv(24, 3, 1, 1, 1, 1, "lrelu", "cat", "c4")
    net.conv(18, 1, 1, 0, 1, 0, "none", "c4", "head")
    net.cost("head", "cost", metric=capi.METRIC_SSE)
    net.sgd(0.001, 0.9, 0.0005)
    return dict(classes=None, out="head")


def smooth_convnet(net, batch=64):
    """A net without discrete decisions: tanh activations (no mask bits), stride-2 convolutions
    instead of max pools (no argmax), tensor-core-eligible channel counts. Rounding noise of the
    tensor-core path then propagates smoothly, so its gradients can be held to the kernel
    tolerance through the whole backward chain (fprop, stride-1 and strided dgrad, wgrad, fused
    batch-norm backward, fully-connected)."""
    net.set_input_shape(32, 32, 3, batch)
    net.conv(32, 3, 1, 1, 1, 1, "tanh", "input", "c1")
    net.conv(32, 3, 2, 1, 1, 1, "tanh", "c1", "c2")      # 16x16
    net.conv(64, 3, 1, 1, 1, 1, "tanh", "c2", "c3")
    net.conv(64, 3, 2, 1, 1, 1, "tanh", "c3", "c4")      # 8x8
    net.conv(128, 1, 1, 0, 1, 1, "tanh", "c4", "c5")
    net.avgpool("c5", "gap")
    net.fullc(10, "none", "gap", "fc")
    net.softmax("fc", "softmax")
    net.cost("softmax", "cost")
    net.sgd(0.005, 0.9, 0.0005)
    return dict(classes=10, out="softmax")


def chain_adam(net, batch=4):
    """chain_convnet trained with Adam, reached the only way the reference reaches it (SURVEY.md
    H7): rates through bcnn_set_adam_optimizer, the switch through the cfg key, both BEFORE the
    layers exist (the reference allocates the moments at layer creation). chain_convnet's own
    bcnn_set_sgd_optimizer call then only sets lr 0.01 / bias momentum 0.9; decay 5e-4."""
    net.adam(0.004, 0.9, 0.999)
    return chain_convnet(net, batch)


CASES = {
    # name: (builder, kwargs, steps)
    "mnist_b8": (configs.mnist, dict(batch=8), 3),
    "cifar_b4": (configs.cifar, dict(batch=4), 3),
    "chain_b4": (chain_convnet, dict(batch=4), 2),
    "resnet_small_b4": (small_resnet, dict(batch=4), 2),
    "yolo_two_heads_b2": (yolo_two_heads, dict(batch=2), 2),
    "chain_adam_b4": (chain_adam, dict(batch=4), 3),
    # BASELINE.json sizes (C1, C2): live-reference comparisons only, no committed golden
    "mnist_b64": (configs.mnist, dict(batch=64), 3),
    "cifar_b128": (configs.cifar, dict(batch=128), 3),
    "smooth_b64": (smooth_convnet, dict(batch=64), 2),
}


def observable_tensors(net):
    """Indexes of every tensor a user can fetch (all of net->tensors[])."""
    return list(range(net.lib.bcnn_b200_num_tensors(net.handle)))


def run_case(net, name, seed=7, collect_steps=True):
    """Returns {key: array}: per step the loss-layer metric, all data + grad tensors after
    backward (before update) for step 0, and all parameter tensors after the last update."""
    builder, kwargs, steps = CASES[name]
    builder(net, **kwargs)
    net.compile()
    configs.init_params(net, seed=seed)
    x = configs.synth_input(net.shape("input"), seed=seed + 1)
    y = configs.synth_labels(net.shape("label"))
    out = {}
    pool_nodes = [i for i in range(net.lib.bcnn_b200_num_nodes(net.handle))
                  if net.lib.bcnn_b200_node_type(net.handle, i) == capi.LAYER_MAXPOOL]
    for step in range(steps):
        net.set("input", x)
        net.set("label", y)
        net.forward()
        net.backward()
        if step == 0:
            for idx in observable_tensors(net):
                t = net._tensor(idx)
                nm = t.name.decode()
                if t.data:
                    out[f"s0/data/{idx}:{nm}"] = net.get(idx)
                if t.grad_data:
                    out[f"s0/grad/{idx}:{nm}"] = net.get(idx, grad=True)
            for node in pool_nodes:
                out[f"s0/argmax/{node}"] = _pool_indexes(net, node)
        out[f"s{step}/cost"] = net.get("cost").ravel()[:1].copy()
        net.update()
    for idx, nm, _ in configs.param_tensors(net):
        out[f"final/data/{idx}:{nm}"] = net.get(idx)
        t = net._tensor(idx)
        if t.grad_data:
            out[f"final/grad/{idx}:{nm}"] = net.get(idx, grad=True)
    return out


def live_reference_case(name, seed, threads=4):
    """run_case on the compiled reference (oracle/_ref), plus under "sens:<key>" the reference's
    own response to a 1-ulp (2e-7 relative) input perturbation: the noise floor of each tensor."""
    from helpers import ref_net, rel_err

    ref = ref_net(threads=threads)
    want = run_case(ref, name, seed=seed)
    ref.close()
    clean = configs.synth_input
    configs.synth_input = lambda shape, seed=12345: (
        clean(shape, seed) * np.float32(1 + 2e-7)).astype(np.float32)
    try:
        ref = ref_net(threads=threads)
        pert = run_case(ref, name, seed=seed)
        ref.close()
    finally:
        configs.synth_input = clean
    for k in list(want):
        if "/argmax/" not in k and np.abs(want[k]).max(initial=0.0) > 0:
            want["sens:" + k] = np.float32(max(rel_err(pert[k], want[k])))
    return want


def _pool_indexes(net, node):
    dst = net.lib.bcnn_b200_node_dst(net.handle, node, 0)
    t = net._tensor(dst)
    buf = np.empty(t.n * t.c * t.h * t.w, dtype=np.int32)
    n = net.lib.bcnn_b200_maxpool_indexes(net.handle, node, buf.ctypes.data)
    assert n == buf.size
    return buf


def model_io_net(net, batch=2):
    """Every node kind that owns records in a weight file: conv+BN (3x3/s1, the
    reference's Winograd route in PREDICT), plain conv, standalone batchnorm, depthwise, 1x1
    conv+BN, and a fully-connected layer with a non-square weight matrix (Darknet transpose)."""
    net.set_input_shape(12, 12, 3, batch)
    net.conv(8, 3, 1, 1, 1, 1, "lrelu", "input", "c1")
    net.maxpool(2, 2, capi.PAD_SAME, "c1", "p1")
    net.conv(8, 3, 2, 1, 1, 0, "relu", "p1", "c2")
    net.batchnorm("c2", "bn2")
    net.depthwise(3, 1, 1, "relu", "bn2", "dw")
    net.conv(12, 1, 1, 0, 1, 1, "none", "dw", "c3")
    net.fullc(10, "none", "c3", "fc")
    net.softmax("fc", "softmax")
    if net.mode != capi.MODE_PREDICT:
        net.cost("softmax", "cost")
        net.sgd(0.01, 0.9, 0.0005)
    return dict(classes=10, out="softmax")
