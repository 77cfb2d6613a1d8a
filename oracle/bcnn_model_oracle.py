"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's weight-file layouts.

Only tests/ (and the development check under tools/hoststub/) import this module; nothing under
bcnn_b200/ does. It restates, independently of bcnn_b200/src/bcnn_model.c:

  * the .bcnnmodel writer   jnbraun/bcnn src/bcnn_net.c:597-681
  * the .bcnnmodel reader   src/bcnn_net.c:1485-1558 with the per-node readers :1220-1466
  * the Darknet *.weights reader (same functions, format == 1) and its header rules :1508-1527
  * the PREDICT-mode batch-norm fold :1278-1289, :1394-1404
  * the fully-connected transpose :1427-1437, :1457-1460

Pinned (tests/test_model_io.py, CPU suite): against tests/golden/model_io.bcnnmodel and
model_io.npz, both produced by the compiled reference (tests/golden/make_golden.py), and, when
oracle/_ref is present, against the reference run live.

A "layout" is the list of nodes that own file records, each a `Node(kind, roles)` where roles maps
role -> (tensor name, element count); roles are bias, weights, mean, var, scales, slopes.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

MAGIC = b"BCNN"
VERSION = (0, 2, 0)  # inc/bcnn/bcnn.h:61-65
EPS = np.float32(0.000001)

# bcnn_layer_type values (inc/bcnn/bcnn.h:155-175)
CONV2D, DECONV2D, DEPTHWISE, ACTIVATION, FULLC, BATCHNORM = 0, 1, 2, 3, 4, 9


@dataclass
class Node:
    kind: str                      # conv | depthwise | batchnorm | fc | prelu
    roles: dict = field(default_factory=dict)
    rows: int = 0                  # fc: size3d of the input  (transpose source rows)
    cols: int = 0                  # fc: size3d of the output


def net_layout(net) -> list[Node]:
    """Walk a built net (either library) through the introspection helpers."""
    lib, h = net.lib, net.handle

    def tensor(idx):
        t = net._tensor(idx)
        return t.name.decode(), t.n * t.c * t.h * t.w, t

    out = []
    for i in range(lib.bcnn_b200_num_nodes(h)):
        kind = lib.bcnn_b200_node_type(h, i)
        src = []
        while lib.bcnn_b200_node_src(h, i, len(src)) >= 0:
            src.append(lib.bcnn_b200_node_src(h, i, len(src)))
        if kind in (CONV2D, DEPTHWISE):
            node = Node("conv" if kind == CONV2D else "depthwise")
            node.roles["weights"] = tensor(src[1])[:2]
            node.roles["bias"] = tensor(src[2])[:2]
            rest = src[3:]
            if kind == CONV2D and len(rest) >= 3:
                for role, idx in zip(("mean", "var", "scales"), rest[:3]):
                    node.roles[role] = tensor(idx)[:2]
                rest = rest[3:]
            if kind == CONV2D and rest:
                node.roles["slopes"] = tensor(rest[0])[:2]
            out.append(node)
        elif kind == ACTIVATION and len(src) > 1:
            out.append(Node("prelu", {"slopes": tensor(src[1])[:2]}))
        elif kind == BATCHNORM:
            c = net._tensor(lib.bcnn_b200_node_dst(h, i, 0)).c
            roles = {role: (tensor(idx)[0], c)
                     for role, idx in zip(("mean", "var", "scales", "bias"), src[1:5])}
            out.append(Node("batchnorm", roles))
        elif kind == FULLC:
            x = net._tensor(src[0])
            y = net._tensor(lib.bcnn_b200_node_dst(h, i, 0))
            out.append(Node("fc", {"weights": tensor(src[1])[:2], "bias": tensor(src[2])[:2]},
                            rows=x.c * x.h * x.w, cols=y.c * y.h * y.w))
    return out


def record_order(node: Node, fmt: str, loading: bool) -> list[str]:
    """Roles of `node` in file order. fmt is "bcnn" or "darknet"."""
    r = node.roles
    if node.kind in ("conv", "depthwise"):
        bn = "scales" in r
        if fmt == "bcnn":
            order = ["bias", "weights"] + (["mean", "var", "scales"] if bn else [])
        else:
            order = ["bias"] + (["scales", "mean", "var"] if bn else []) + ["weights"]
        if loading and "slopes" in r:  # read (:1305-1321) but never written by the saver
            order.append("slopes")
        return order
    if node.kind == "prelu":
        return ["slopes"] if fmt == "bcnn" else []
    if node.kind == "batchnorm":
        return ["mean", "var", "scales", "bias"] if fmt == "bcnn" else ["scales", "mean", "var"]
    if node.kind == "fc":
        return ["bias", "weights"]
    raise ValueError(node.kind)


def all_names(layout) -> list[tuple[str, int]]:
    return [v for node in layout for v in node.roles.values()]


def _body(layout, values, fmt, loading):
    parts = []
    for node in layout:
        for role in record_order(node, fmt, loading):
            name, size = node.roles[role]
            arr = np.ascontiguousarray(values[name], dtype=np.float32).reshape(-1)
            assert arr.size >= size, (name, arr.size, size)
            parts.append(arr[:size].tobytes())
    return b"".join(parts)


def write_bcnn(path, layout, values) -> bytes:
    """What bcnn_save_weights writes for these tensor values."""
    blob = MAGIC + struct.pack("<3I", *VERSION) + _body(layout, values, "bcnn", loading=False)
    Path(path).write_bytes(blob)
    return blob


def write_darknet(path, layout, values, major=0, minor=2, revision=0, seen=0) -> bytes:
    """A Darknet *.weights file with the records the reference's reader expects. Header rule
    (:1508-1522): `seen` is 8 bytes when major*10+minor >= 2 (and both < 1000), else 4."""
    head = struct.pack("<3i", major, minor, revision)
    wide = (major * 10 + minor) >= 2 and major < 1000 and minor < 1000
    head += struct.pack("<Q" if wide else "<i", seen)
    blob = head + _body(layout, values, "darknet", loading=True)
    Path(path).write_bytes(blob)
    return blob


def file_format(path) -> str:
    """:1468-1483 -- by the text after the last '.'."""
    ext = str(path).rsplit(".", 1)[-1]
    return {"weights": "darknet", "onnx": "onnx"}.get(ext, "bcnn")


def fold(bias, scales, mean, var):
    """PREDICT fold in float32, operation by operation as :1281-1288."""
    bias, scales, mean, var = (np.asarray(a, dtype=np.float32) for a in (bias, scales, mean, var))
    sd = np.sqrt(var + EPS, dtype=np.float32)
    new_bias = (bias - (scales * mean).astype(np.float32) / sd).astype(np.float32)
    new_scales = (scales / sd).astype(np.float32)
    return new_bias, new_scales


def read(path, layout, predict=False, initial=None) -> dict:
    """What bcnn_load_weights leaves in the parameter tensors: {tensor name: flat float32}.
    `initial` supplies tensors the file does not carry (a standalone batchnorm's bias in a
    Darknet file); default zeros."""
    fmt = file_format(path)
    data = Path(path).read_bytes()
    transpose = False
    if fmt == "bcnn":
        if data[:4] != MAGIC or len(data) < 16:
            raise ValueError("BCNN_INVALID_MODEL: bad magic")
        pos = 16
    elif fmt == "darknet":
        major, minor, _ = struct.unpack_from("<3i", data, 0)
        pos = 12 + (8 if (major * 10 + minor) >= 2 and major < 1000 and minor < 1000 else 4)
        transpose = major > 1000 or minor > 1000
    else:
        raise ValueError("BCNN_INVALID_MODEL: unsupported format")
    out = {name: np.array(initial[name], dtype=np.float32).reshape(-1) if initial and name in initial
           else np.zeros(size, np.float32) for name, size in all_names(layout)}
    for node in layout:
        for role in record_order(node, fmt, loading=True):
            name, size = node.roles[role]
            if pos + 4 * size > len(data):
                raise ValueError(f"BCNN_INVALID_MODEL: short read in {name}")
            out[name][:size] = np.frombuffer(data, dtype="<f4", count=size, offset=pos)
            pos += 4 * size
        if predict and "scales" in node.roles and node.kind in ("conv", "batchnorm"):
            (b, c), (s, _), (m, _), (v, _) = (node.roles[k] for k in ("bias", "scales", "mean", "var"))
            c = node.roles["scales"][1]
            out[b][:c], out[s][:c] = fold(out[b][:c], out[s][:c], out[m][:c], out[v][:c])
        if node.kind == "fc" and transpose:
            name, size = node.roles["weights"]
            out[name] = out[name].reshape(node.rows, node.cols).T.copy().reshape(-1)
    return out
