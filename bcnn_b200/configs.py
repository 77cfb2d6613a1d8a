"""The BASELINE.json workloads expressed through the bcnn C API, plus synthetic data.

Each builder takes a `capi.Net` (either flavour) and issues the same bcnn_add_*_layer
calls a bcnn user would write:
  mnist      -- reference examples/mnist/mnist_example.c:30-55   (BASELINE configs[0])
  cifar      -- reference examples/cifar10/cifar10_example.c:32-63 simple_net (configs[1])
  mobilenet  -- MobileNet-v1 224 as bcnn layers (configs[2]; no cfg ships with the
                reference, synthesised per SURVEY.md section 8 / Appendix A)
  yolo_tiny  -- examples/yolo/yolov3-tiny.cfg (configs[3]): trunk, both heads, both yolo layers
                (detection loss against box labels in TRAIN mode)
  resnet50   -- ResNet-50 v1.5 224 as bcnn layers (configs[4], the headline workload)

Synthetic inputs follow SURVEY.md 8d: FP32 uniform [-1,1) images, one-hot labels with
class = sample index mod classes, Xavier-range weights, all from a seeded generator so
both libraries receive identical bits.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .capi import PAD_SAME


def mnist(net, batch=64, act="relu"):
    net.set_input_shape(28, 28, 1, batch)
    net.conv(32, 3, 1, 1, 1, 0, act, "input", "conv1")
    net.batchnorm("conv1", "bn1")
    net.maxpool(2, 2, PAD_SAME, "bn1", "pool1")
    net.conv(32, 3, 1, 1, 1, 0, act, "pool1", "conv2")
    net.batchnorm("conv2", "bn2")
    net.maxpool(2, 2, PAD_SAME, "bn2", "pool2")
    net.fullc(256, act, "pool2", "fc1")
    net.batchnorm("fc1", "bn3")
    net.fullc(10, act, "bn3", "fc2")
    net.softmax("fc2", "softmax")
    net.cost("softmax", "cost")
    net.sgd(0.003, 0.9, 0.0005)  # mnist_example.c:136-139
    return dict(classes=10, out="softmax")


def cifar(net, batch=128, act="relu"):
    net.set_input_shape(32, 32, 3, batch)
    net.conv(32, 3, 1, 1, 1, 1, act, "input", "conv1_1")
    net.conv(32, 3, 1, 1, 1, 1, act, "conv1_1", "conv1_2")
    net.conv(32, 3, 1, 1, 1, 1, act, "conv1_2", "conv1_3")
    net.maxpool(2, 2, PAD_SAME, "conv1_3", "pool1")
    net.conv(64, 3, 1, 1, 1, 1, act, "pool1", "conv2_1")
    net.conv(64, 3, 1, 1, 1, 1, act, "conv2_1", "conv2_2")
    net.conv(64, 3, 1, 1, 1, 1, act, "conv2_2", "conv2_3")
    net.maxpool(2, 2, PAD_SAME, "conv2_3", "pool2")
    net.fullc(512, act, "pool2", "fc1")
    net.batchnorm("fc1", "bn3")
    net.fullc(10, act, "bn3", "fc2")
    net.softmax("fc2", "softmax")
    net.cost("softmax", "cost")
    # the example's "adam" call leaves the optimizer on SGD, momentum 0.9, lr 0.005
    # (cifar10_example.c:234-238)
    net.sgd(0.005, 0.9, 0.0005)
    return dict(classes=10, out="softmax")


def mobilenet(net, batch=1, res=224, classes=1000):
    net.set_input_shape(res, res, 3, batch)
    net.conv(32, 3, 2, 1, 1, 1, "relu", "input", "conv0")
    cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1),
           (512, 1), (512, 1), (512, 1), (1024, 2), (1024, 1)]
    prev = "conv0"
    for i, (cout, stride) in enumerate(cfg, 1):
        net.depthwise(3, stride, 1, "relu", prev, f"dw{i}")
        net.conv(cout, 1, 1, 0, 1, 1, "relu", f"dw{i}", f"pw{i}")
        prev = f"pw{i}"
    net.avgpool(prev, "gap")
    net.fullc(classes, "none", "gap", "fc")
    net.softmax("fc", "softmax")
    if net.mode != capi.MODE_PREDICT:
        net.cost("softmax", "cost")
        net.sgd(0.005, 0.9, 0.0005)
    return dict(classes=classes, out="softmax")


def yolo_tiny(net, batch=1, res=416):
    """YOLOv3-tiny (examples/yolo/yolov3-tiny.cfg) with both heads: trunk, coarse head at res/32,
    route -> 1x1 conv -> upsample x2 -> concat with the 256-channel trunk tensor at res/16 -> fine
    head, each ending in its yolo layer (head activation on the device; in TRAIN mode the detection
    loss on the host, as in the reference). Profiles older than round 1l trained this net with a
    euclidean cost on the fine head instead."""
    net.set_input_shape(res, res, 3, batch)
    prev = "input"
    for i, c in enumerate([16, 32, 64, 128, 256, 512]):
        net.conv(c, 3, 1, 1, 1, 1, "lrelu", prev, f"conv{i}")
        net.maxpool(2, 2 if i < 5 else 1, PAD_SAME, f"conv{i}", f"pool{i}")
        prev = f"pool{i}"
    net.conv(1024, 3, 1, 1, 1, 1, "lrelu", prev, "conv6")
    net.conv(256, 1, 1, 0, 1, 1, "lrelu", "conv6", "conv7")
    net.conv(512, 3, 1, 1, 1, 1, "lrelu", "conv7", "conv8")
    net.conv(255, 1, 1, 0, 1, 0, "none", "conv8", "head1")
    net.conv(128, 1, 1, 0, 1, 1, "lrelu", "conv7", "conv9")     # route -4
    net.upsample(2, "conv9", "up")
    net.concat(["up", "conv4"], "route")                         # route -1, 8
    net.conv(256, 3, 1, 1, 1, 1, "lrelu", "route", "conv10")
    net.conv(255, 1, 1, 0, 1, 0, "none", "conv10", "head")
    # the two [yolo] layers of the cfg (80 classes, 6 anchors, masks 3-5 and 0-2); the fine one is
    # added last, as in the cfg, so bcnn_predict_on_batch returns it. In TRAIN mode they carry the
    # detection loss against [N,1,1,250] box labels (synth_labels makes 3 boxes per image).
    anchors = [10, 14, 23, 27, 37, 58, 81, 82, 135, 169, 344, 319]
    net.yolo([3, 4, 5], anchors, 80, "head1", "yolo1")
    net.yolo([0, 1, 2], anchors, 80, "head", "yolo2")
    if net.mode != capi.MODE_PREDICT:
        net.sgd(0.001, 0.9, 0.0005)
    return dict(classes=None, out="yolo2")


def resnet50(net, batch=256, res=224, classes=1000, widths=(64, 128, 256, 512),
             blocks=(3, 4, 6, 3), stage_strides=(1, 2, 2, 2), metric=None):
    """ResNet-50 v1.5 (stride on the 3x3). Residual adds use bcnn_add_eltwise_layer."""
    net.set_input_shape(res, res, 3, batch)
    net.conv(widths[0], 7, 2, 3, 1, 1, "relu", "input", "conv1")
    net.maxpool(3, 2, PAD_SAME, "conv1", "pool1")
    prev, cin = "pool1", widths[0]
    for s, (mid, nblk) in enumerate(zip(widths, blocks)):
        cout = mid * 4
        for b in range(nblk):
            stride = stage_strides[s] if b == 0 else 1
            tag = f"s{s}b{b}"
            net.conv(mid, 1, 1, 0, 1, 1, "relu", prev, tag + "_a")
            net.conv(mid, 3, stride, 1, 1, 1, "relu", tag + "_a", tag + "_b")
            net.conv(cout, 1, 1, 0, 1, 1, "none", tag + "_b", tag + "_c")
            short = prev
            if cin != cout or stride != 1:
                net.conv(cout, 1, stride, 0, 1, 1, "none", prev, tag + "_sc")
                short = tag + "_sc"
            net.eltwise("relu", tag + "_c", short, tag + "_out")
            prev, cin = tag + "_out", cout
    net.avgpool(prev, "gap")
    net.fullc(classes, "none", "gap", "fc")
    net.softmax("fc", "softmax")
    if net.mode != capi.MODE_PREDICT:
        if metric is None:
            net.cost("softmax", "cost")
        else:
            net.cost("softmax", "cost", metric=metric)
        net.sgd(0.005, 0.9, 0.0005)
    return dict(classes=classes, out="softmax")


# every convolution shape of ResNet-50 v1.5 at 224 x 224:
# (cin, h, cout, k, stride, pad, how many nodes of the net have it)
RESNET50_CONV_SHAPES = [
    (3, 224, 64, 7, 2, 3, 1),
    (64, 56, 64, 1, 1, 0, 1), (64, 56, 64, 3, 1, 1, 3), (64, 56, 256, 1, 1, 0, 4), (256, 56, 64, 1, 1, 0, 2),
    (256, 56, 128, 1, 1, 0, 1), (128, 56, 128, 3, 2, 1, 1), (128, 28, 512, 1, 1, 0, 4), (256, 56, 512, 1, 2, 0, 1),
    (512, 28, 128, 1, 1, 0, 3), (128, 28, 128, 3, 1, 1, 3), (512, 28, 256, 1, 1, 0, 1), (256, 28, 256, 3, 2, 1, 1),
    (256, 14, 1024, 1, 1, 0, 6), (512, 28, 1024, 1, 2, 0, 1), (1024, 14, 256, 1, 1, 0, 5), (256, 14, 256, 3, 1, 1, 5),
    (1024, 14, 512, 1, 1, 0, 1), (512, 14, 512, 3, 2, 1, 1), (512, 7, 2048, 1, 1, 0, 3), (1024, 14, 2048, 1, 2, 0, 1),
    (2048, 7, 512, 1, 1, 0, 2), (512, 7, 512, 3, 1, 1, 2),
]


def yolov3_tiny_cfg(batch=1, width=416, height=416, max_filters=None):
    """The text of a Darknet-dialect YOLOv3-tiny config, generated from the layer list (the
    reference ships the same network as examples/yolo/yolov3-tiny.cfg; tests/test_baseline_parity
    checks on the CPU that both files make the reference build the identical graph). Keys the
    reference's reader ignores (augmentation, burn-in, thresholds) are left out.
    max_filters caps the layer widths: with 384 every backward GEMM of the reference stays inside
    the range its transposed-operand blocking handles (SURVEY hazard H12)."""
    def conv(filters, size, bn=1, act="leaky"):
        if max_filters:
            filters = min(filters, max_filters)
        head = "[convolutional]\n" + ("batch_normalize=1\n" if bn else "")
        return head + f"filters={filters}\nsize={size}\nstride=1\npad=1\nactivation={act}\n"

    def yolo(mask):
        return ("[yolo]\nmask = %s\nanchors = 10,14,  23,27,  37,58,  81,82,  135,169,  344,319\n"
                "classes=80\nnum=6\n" % ",".join(str(m) for m in mask))

    sections = [f"[net]\nbatch={batch}\nsubdivisions=1\nwidth={width}\nheight={height}\nchannels=3\n"
                "momentum=0.9\ndecay=0.0005\nlearning_rate=0.001\nmax_batches = 500200\n"
                "policy=steps\nsteps=400000,450000\nscales=.1,.1\n"]
    for i, c in enumerate([16, 32, 64, 128, 256, 512]):
        sections.append(conv(c, 3))
        sections.append(f"[maxpool]\nsize=2\nstride={2 if i < 5 else 1}\n")
    sections += [conv(1024, 3), conv(256, 1), conv(512, 3), conv(255, 1, bn=0, act="linear"),
                 yolo([3, 4, 5]), "[route]\nlayers = -4\n", conv(128, 1), "[upsample]\nstride=2\n",
                 "[route]\nlayers = -1, 8\n", conv(256, 3), conv(255, 1, bn=0, act="linear"),
                 yolo([0, 1, 2])]
    return "\n".join(sections)


BUILDERS = dict(mnist=mnist, cifar=cifar, mobilenet=mobilenet, yolo_tiny=yolo_tiny,
                resnet50=resnet50)


# --------------------------------------------------------------------------------
# synthetic data
# --------------------------------------------------------------------------------

def synth_input(shape, seed=12345):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=shape).astype(np.float32)


def synth_labels(shape, first_sample=0):
    """One-hot [N, classes, 1, 1] (class = global sample index mod classes); for a dense
    target (yolo head stand-in) small uniform values."""
    n, c, h, w = shape
    if (c, h, w) == (1, 1, 250):  # detection labels of a net that ends in yolo layers
        return synth_yolo_labels(n, 3, classes=80, seed=777 + first_sample)
    if h * w == 1:
        y = np.zeros(shape, dtype=np.float32)
        for i in range(n):
            y[i, (first_sample + i) % c, 0, 0] = 1.0
        return y
    rng = np.random.default_rng(777 + first_sample)
    return rng.uniform(0.0, 1.0, size=shape).astype(np.float32)


def param_tensors(net):
    """[(tensor_index, name, shape)] of every parameter tensor, in node order (src[1:] of
    each node; the label and activation inputs are skipped)."""
    lib, h = net.lib, net.handle
    out, seen = [], set()
    for node in range(lib.bcnn_b200_num_nodes(h)):
        ntype = lib.bcnn_b200_node_type(h, node)
        if ntype in (capi.LAYER_COST, 12):  # cost: src[1] is the label; eltwise: 2 inputs
            continue
        i = 1
        while True:
            idx = lib.bcnn_b200_node_src(h, node, i)
            if idx < 0:
                break
            i += 1
            if idx in seen or idx < 2:
                continue
            seen.add(idx)
            t = net._tensor(idx)
            out.append((idx, t.name.decode(), (t.n, t.c, t.h, t.w)))
    return out


def init_params(net, seed=2024, randomize_bn=True):
    """Deterministic synthetic parameters written into both host mirror and device.
    Same enumeration order on both flavours => identical bits on both libraries."""
    rng = np.random.default_rng(seed)
    for idx, name, shape in param_tensors(net):
        n, c, h, w = shape
        size = n * c * h * w
        if name.endswith("_w"):
            fan_in = c * h * w if n > 1 else 9
            amp = np.sqrt(3.0 / max(1, fan_in))
            val = rng.uniform(-amp, amp, size=size)
        elif name.endswith("_b"):
            val = rng.uniform(-0.1, 0.1, size=size)  # never exactly 1.0f
        elif name.endswith("_scales"):
            val = rng.uniform(0.5, 1.5, size=size) if randomize_bn else np.ones(size)
        elif name.endswith("_run_var"):
            val = rng.uniform(0.5, 1.5, size=size) if net.mode != capi.MODE_TRAIN else np.zeros(size)
        elif name.endswith("_run_mean"):
            val = rng.uniform(-0.2, 0.2, size=size) if net.mode != capi.MODE_TRAIN else np.zeros(size)
        elif "prelu" in name:
            val = rng.uniform(0.05, 0.3, size=size)
        else:
            val = np.zeros(size)
        net.set(idx, val.astype(np.float32))


def synth_yolo_labels(n, boxes_per_image=3, classes=80, seed=12345):
    """Detection labels in the reference's layout (src/layers/bcnn_yolo.c:69-72, 282-290):
    [n, 1, 1, 250] = up to 50 boxes x (x, y, w, h, class), centre and size relative to the image,
    the list ends at the first x == 0. Boxes stay inside (0.05, 0.95): a box centred exactly on
    the right / bottom edge indexes one cell past the grid in the reference."""
    rng = np.random.default_rng(seed)
    counts = boxes_per_image if hasattr(boxes_per_image, "__len__") else [boxes_per_image] * n
    lab = np.zeros((n, 1, 1, 250), np.float32)
    for b in range(n):
        for t in range(min(int(counts[b]), 50)):
            lab[b, 0, 0, 5 * t:5 * t + 5] = [rng.uniform(0.05, 0.95), rng.uniform(0.05, 0.95),
                                             rng.uniform(0.05, 0.6), rng.uniform(0.05, 0.6),
                                             rng.integers(0, classes)]
    return lab
