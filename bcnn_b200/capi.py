"""ctypes binding of the bcnn C API (include/bcnn/bcnn.h) and a thin `Net` wrapper.

The same binding drives two libraries, because they export the same API:
  * libbcnn_b200.so  -- this repo's B200 path (flavour "b200": bcnn_tensor has the
                        data_gpu / grad_data_gpu fields, host mirrors must be uploaded);
  * a CPU build of the reference (flavour "cpu", used ONLY by tests/ and bench.py's
    reference arm as the oracle; never loaded from here).

Nothing in this module falls back to the CPU: `load_library()` raises if the CUDA
library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libbcnn_b200.so"

# ---- enums (include/bcnn/bcnn.h) -------------------------------------------------
MODE_PREDICT, MODE_TRAIN, MODE_VALID = 0, 1, 2
ACT = dict(none=0, tanh=1, relu=2, ramp=3, softplus=4, lrelu=5, abs=6, clamp=7, prelu=8,
           logistic=9)
PAD_SAME, PAD_VALID, PAD_CAFFE = 0, 1, 2
FILLER_FIXED, FILLER_XAVIER, FILLER_MSRA = 0, 1, 2
LOSS_EUCLIDEAN = 0
METRIC_ERROR_RATE, METRIC_LOGLOSS, METRIC_SSE, METRIC_MSE = 0, 1, 2, 3
LOG_SILENT = 3
LAYER_CONV2D, LAYER_MAXPOOL, LAYER_BATCHNORM, LAYER_COST = 0, 5, 9, 16
MATH_FP32, MATH_TC, MATH_TC_BF16 = 0, 1, 2   # include/bcnn_b200.h; 2 = resident BF16 NHWC activations


class TensorB200(C.Structure):
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("has_grad", C.c_int), ("name", C.c_char_p),
                ("data", C.POINTER(C.c_float)), ("grad_data", C.POINTER(C.c_float)),
                ("data_gpu", C.c_void_p), ("grad_data_gpu", C.c_void_p)]


class Detection(C.Structure):  # bcnn_output_detection
    _fields_ = [("num_classes", C.c_int), ("x", C.c_float), ("y", C.c_float), ("w", C.c_float),
                ("h", C.c_float), ("prob", C.POINTER(C.c_float)), ("mask", C.POINTER(C.c_float)),
                ("objectness", C.c_float)]


class TensorCPU(C.Structure):  # reference built without BCNN_USE_CUDA
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("has_grad", C.c_int), ("name", C.c_char_p),
                ("data", C.POINTER(C.c_float)), ("grad_data", C.POINTER(C.c_float))]


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libbcnn_b200.so. Fails loudly if it is missing: there is no fallback."""
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            f"{p} not found: build the CUDA extension first (python -m bcnn_b200.build). "
            "bcnn_b200 has no CPU or PyTorch fallback path.")
    # RTLD_LOCAL: the library exports the bcnn API names; keep them out of the global
    # scope so another bcnn build in the same process (the tests' CPU reference) is not
    # interposed. The library itself is linked -Bsymbolic.
    return C.CDLL(str(p), mode=C.RTLD_LOCAL)


def bind_bcnn_api(lib: C.CDLL, tensor_type) -> None:
    """Declare argtypes / restypes of the public bcnn API on `lib`."""
    vp, i, f, s = C.c_void_p, C.c_int, C.c_float, C.c_char_p
    sigs = {
        "bcnn_init_net": (i, [C.POINTER(vp), i]),
        "bcnn_end_net": (None, [C.POINTER(vp)]),
        "bcnn_set_log_context": (None, [vp, vp, i]),
        "bcnn_set_num_threads": (i, [vp, i, vp]),
        "bcnn_set_input_shape": (None, [vp, i, i, i, i]),
        "bcnn_compile_net": (i, [vp]),
        "bcnn_set_mode": (i, [vp, i]),
        "bcnn_set_sgd_optimizer": (None, [vp, f, f]),
        "bcnn_set_learning_rate_policy": (None, [vp, i, f, f, f, i, i]),
        "bcnn_set_weight_regularizer": (None, [vp, f]),
        "bcnn_forward": (None, [vp]),
        "bcnn_backward": (None, [vp]),
        "bcnn_update": (None, [vp]),
        "bcnn_get_tensor_index_by_name": (i, [vp, s]),
        "bcnn_get_tensor_by_index": (C.POINTER(tensor_type), [vp, i]),
        "bcnn_get_tensor_by_name": (C.POINTER(tensor_type), [vp, s]),
        "bcnn_add_convolutional_layer": (i, [vp, i, i, i, i, i, i, i, i, i, s, s]),
        "bcnn_add_depthwise_conv_layer": (i, [vp, i, i, i, i, i, i, s, s]),
        "bcnn_add_batchnorm_layer": (i, [vp, s, s]),
        "bcnn_add_activation_layer": (i, [vp, i, s]),
        "bcnn_add_maxpool_layer": (i, [vp, i, i, i, s, s]),
        "bcnn_add_avgpool_layer": (i, [vp, s, s]),
        "bcnn_add_fullc_layer": (i, [vp, i, i, i, i, s, s]),
        "bcnn_add_softmax_layer": (i, [vp, s, s]),
        "bcnn_add_eltwise_layer": (i, [vp, i, s, s, s]),
        "bcnn_add_concat_layer": (i, [vp, i, C.POINTER(C.c_char_p), s]),
        "bcnn_add_upsample_layer": (i, [vp, i, s, s]),
        "bcnn_add_cost_layer": (i, [vp, i, i, f, s, s, s]),
        "bcnn_set_adam_optimizer": (None, [vp, f, f, f]),
        "bcnn_train_on_batch": (f, [vp]),
        "bcnn_predict_on_batch": (f, [vp, C.POINTER(C.POINTER(tensor_type))]),
        "bcnn_add_input": (i, [vp, i, i, i, s]),
        "bcnn_load_net": (i, [vp, s, s]),
        "bcnn_yolo_get_detections": (C.POINTER(Detection), [vp, i, i, i, i, i, f, i, C.POINTER(C.c_int)]),
        "bcnn_add_yolo_layer": (i, [vp, i, i, i, i, C.POINTER(C.c_int), C.POINTER(C.c_float), s, s]),
        "bcnn_get_batch_size": (i, [vp]),
        "bcnn_load_weights": (i, [vp, s]),
        "bcnn_save_weights": (i, [vp, s]),
        # internal but exported by both libraries (reference src/bcnn_net.h:75)
        "bcnn_net_set_param": (None, [vp, s, s]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args


def bind_b200_ext(lib: C.CDLL) -> None:
    """Declare the net-level extensions (include/bcnn_b200_net.h) and the helpers of
    include/bcnn_b200.h that Python-side plumbing needs."""
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sigs = {
        "bcnn_b200_set_conv_math": (None, [vp, i]),
        "bcnn_b200_get_conv_math": (i, [vp]),
        "bcnn_b200_set_reference_quirks": (None, [vp, i]),
        "bcnn_b200_get_reference_quirks": (i, [vp]),
        "bcnn_b200_yolo_loss_on_host": (f, [vp, i]),
        "bcnn_b200_set_graphs": (None, [vp, i]),
        "bcnn_b200_get_graphs": (i, [vp]),
        "bcnn_b200_get_stream": (vp, [vp]),
        "bcnn_b200_sync": (None, [vp]),
        "bcnn_b200_upload_tensor": (i, [vp, i]),
        "bcnn_b200_upload_inputs": (sz, [vp]),
        "bcnn_b200_prefetch_inputs": (sz, [vp]),
        "bcnn_b200_get_loss": (f, [vp]),
        "bcnn_b200_train_step": (f, [vp, i, i]),
        "bcnn_b200_profile": (None, [vp, i]),
        "bcnn_b200_profile_node_ms": (i, [vp, i, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "bcnn_b200_num_nodes": (i, [vp]),
        "bcnn_b200_num_tensors": (i, [vp]),
        "bcnn_b200_node_type": (i, [vp, i]),
        "bcnn_b200_node_src": (i, [vp, i, i]),
        "bcnn_b200_node_dst": (i, [vp, i, i]),
        "bcnn_b200_tensor_dims": (i, [vp, i, C.POINTER(C.c_int)]),
        "bcnn_b200_maxpool_indexes": (i, [vp, i, vp]),
        "bcnn_b200_bn_saved_stats": (i, [vp, i, vp, vp]),
        "bcnn_b200_dp_get_unique_id": (i, [vp]),
        "bcnn_b200_dp_init": (i, [vp, i, i, vp]),
        "bcnn_b200_dp_shutdown": (None, [vp]),
        "bcnn_b200_dp_world": (i, [vp]),
        "bcnn_b200_dp_bytes_per_step": (sz, [vp]),
        "bcnn_b200_dp_groups_per_step": (i, [vp]),
        "bcnn_b200_dp_allreduce_probe_ms": (f, [vp, i]),
        "bcnn_b200_set_device": (i, [i]),
        "bcnn_b200_device_count": (i, []),
        "bcnn_b200_sm_count": (i, []),
        "bcnn_b200_launch_count": (C.c_uint64, []),
        "bcnn_b200_event_create": (vp, []),
        "bcnn_b200_event_destroy": (None, [vp]),
        "bcnn_b200_event_record": (i, [vp, vp]),
        "bcnn_b200_event_elapsed_ms": (f, [vp, vp]),
        "bcnn_b200_stream_sync": (i, [vp]),
        "bcnn_b200_malloc": (vp, [sz]),
        "bcnn_b200_free": (None, [vp]),
        "bcnn_b200_memcpy_h2d": (i, [vp, vp, sz, vp]),
        "bcnn_b200_memcpy_d2h": (i, [vp, vp, sz, vp]),
        "bcnn_b200_fill_f32": (i, [vp, sz, f, vp]),
        "bcnn_b200_error_string": (C.c_char_p, [i]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args


_B200 = None


def b200() -> C.CDLL:
    """The process-wide libbcnn_b200.so handle, bound once."""
    global _B200
    if _B200 is None:
        lib = load_library()
        bind_bcnn_api(lib, TensorB200)
        bind_b200_ext(lib)
        bind_kernel_abi(lib)
        _B200 = lib
    return _B200


def _b(s: str) -> bytes:
    return s.encode()


class Net:
    """A bcnn_net driven through the C API. `flavour` is "b200" (this repo's library) or
    "cpu" (a CPU build of the reference, tests only)."""

    def __init__(self, mode: int = MODE_TRAIN, lib: C.CDLL | None = None,
                 flavour: str = "b200"):
        self.flavour = flavour
        self.lib = lib if lib is not None else b200()
        self.handle = C.c_void_p()
        st = self.lib.bcnn_init_net(C.byref(self.handle), mode)
        if st != 0:
            raise RuntimeError(f"bcnn_init_net failed with status {st} "
                               "(the B200 path needs a CUDA device)")
        self.lib.bcnn_set_log_context(self.handle, None, LOG_SILENT)
        self.mode = mode

    # -- lifetime --
    def close(self):
        if self.handle:
            self.lib.bcnn_end_net(C.byref(self.handle))
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, what):
        if st != 0:
            raise RuntimeError(f"{what} failed with bcnn_status {st}")

    # -- construction (argument order of include/bcnn/bcnn.h) --
    def set_input_shape(self, w, h, c, batch):
        self.lib.bcnn_set_input_shape(self.handle, w, h, c, batch)

    def conv(self, num, size, stride, pad, groups, batch_norm, act, src, dst,
             init=FILLER_XAVIER):
        self._check(self.lib.bcnn_add_convolutional_layer(
            self.handle, num, size, stride, pad, groups, batch_norm, init, ACT[act], 0,
            _b(src), _b(dst)), f"conv {dst}")

    def depthwise(self, size, stride, pad, act, src, dst, init=FILLER_XAVIER):
        self._check(self.lib.bcnn_add_depthwise_conv_layer(
            self.handle, size, stride, pad, 0, init, ACT[act], _b(src), _b(dst)),
            f"depthwise {dst}")

    def batchnorm(self, src, dst):
        self._check(self.lib.bcnn_add_batchnorm_layer(self.handle, _b(src), _b(dst)),
                    f"batchnorm {dst}")

    def activation(self, act, name):
        self._check(self.lib.bcnn_add_activation_layer(self.handle, ACT[act], _b(name)),
                    f"activation {name}")

    def maxpool(self, size, stride, padding, src, dst):
        self._check(self.lib.bcnn_add_maxpool_layer(self.handle, size, stride, padding,
                                                    _b(src), _b(dst)), f"maxpool {dst}")

    def avgpool(self, src, dst):
        self._check(self.lib.bcnn_add_avgpool_layer(self.handle, _b(src), _b(dst)),
                    f"avgpool {dst}")

    def fullc(self, out, act, src, dst, init=FILLER_XAVIER):
        self._check(self.lib.bcnn_add_fullc_layer(self.handle, out, init, ACT[act], 0,
                                                  _b(src), _b(dst)), f"fullc {dst}")

    def softmax(self, src, dst):
        self._check(self.lib.bcnn_add_softmax_layer(self.handle, _b(src), _b(dst)),
                    f"softmax {dst}")

    def eltwise(self, act, src1, src2, dst):
        self._check(self.lib.bcnn_add_eltwise_layer(self.handle, ACT[act], _b(src1),
                                                    _b(src2), _b(dst)), f"eltwise {dst}")

    def concat(self, srcs, dst):
        arr = (C.c_char_p * len(srcs))(*[_b(x) for x in srcs])
        self._check(self.lib.bcnn_add_concat_layer(self.handle, len(srcs), arr, _b(dst)),
                    f"concat {dst}")

    def upsample(self, size, src, dst):
        self._check(self.lib.bcnn_add_upsample_layer(self.handle, size, _b(src), _b(dst)),
                    f"upsample {dst}")

    def cost(self, src, dst="cost", metric=METRIC_ERROR_RATE, scale=1.0):
        self._check(self.lib.bcnn_add_cost_layer(self.handle, LOSS_EUCLIDEAN, metric, scale,
                                                 _b(src), b"label", _b(dst)), f"cost {dst}")

    def sgd(self, lr, momentum, decay=0.0):
        self.lib.bcnn_set_sgd_optimizer(self.handle, lr, momentum)
        self.lib.bcnn_set_weight_regularizer(self.handle, decay)

    def set_param(self, name: str, value) -> None:
        """A solver key as the cfg reader sets it (optimizer=adam, beta1=..., decay=...)."""
        self.lib.bcnn_net_set_param(self.handle, _b(name), _b(str(value)))

    def adam(self, lr, beta1=0.9, beta2=0.999, decay=0.0):
        """Adam the only way the reference reaches it: the setter for the rates,
        the cfg key for the switch. Call before adding layers on the reference (it allocates the
        moments at layer creation)."""
        self.lib.bcnn_set_adam_optimizer(self.handle, lr, beta1, beta2)
        self.lib.bcnn_set_weight_regularizer(self.handle, decay)
        self.set_param("optimizer", "adam")

    def load_net(self, cfg_path, model_path=None) -> int:
        """bcnn_load_net: build the graph from a .cfg / .conf file (Darknet dialect when the model
        is a *.weights file) and load the model. Returns the bcnn_status."""
        model = _b(str(model_path)) if model_path is not None else None
        return int(self.lib.bcnn_load_net(self.handle, _b(str(cfg_path)), model))

    def yolo(self, mask, anchors, classes, src, dst, coords=4):
        m = (C.c_int * len(mask))(*mask)
        a = (C.c_float * len(anchors))(*anchors)
        self._check(self.lib.bcnn_add_yolo_layer(self.handle, len(mask), classes, coords,
                                                 len(anchors) // 2, m, a, _b(src), _b(dst)),
                    f"yolo {dst}")

    def yolo_detections(self, batch, width, height, thresh, relative=True) -> np.ndarray:
        """bcnn_yolo_get_detections as an array [num_dets, 5 + classes]: x, y, w, h, objectness,
        class probabilities (rows in the order the library returns them)."""
        count = C.c_int(0)
        _, _, net_h, net_w = self.shape("input")
        dets = self.lib.bcnn_yolo_get_detections(self.handle, batch, width, height, net_w, net_h,
                                                 thresh, int(relative), C.byref(count))
        rows = []
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        for k in range(count.value):
            d = dets[k]
            rows.append([d.x, d.y, d.w, d.h, d.objectness] + [d.prob[j] for j in range(d.num_classes)])
            libc.free(C.cast(d.prob, C.c_void_p))
            if d.mask:
                libc.free(C.cast(d.mask, C.c_void_p))
        if count.value:
            libc.free(C.cast(dets, C.c_void_p))
        if not rows:
            return np.zeros((0, 5), np.float32)
        return np.array(rows, dtype=np.float32).reshape(count.value, -1)

    def structure(self) -> dict:
        """Graph as plain data: nodes [type, src indices, dst indices], tensors [name, n, c, h, w]."""
        lib, h = self.lib, self.handle
        nodes = []
        for i in range(lib.bcnn_b200_num_nodes(h)):
            src, dst = [], []
            while lib.bcnn_b200_node_src(h, i, len(src)) >= 0:
                src.append(lib.bcnn_b200_node_src(h, i, len(src)))
            while lib.bcnn_b200_node_dst(h, i, len(dst)) >= 0:
                dst.append(lib.bcnn_b200_node_dst(h, i, len(dst)))
            nodes.append([lib.bcnn_b200_node_type(h, i), src, dst])
        tensors = []
        for i in range(lib.bcnn_b200_num_tensors(h)):
            t = self._tensor(i)
            tensors.append([t.name.decode(), t.n, t.c, t.h, t.w])
        return dict(nodes=nodes, tensors=tensors, batch=int(lib.bcnn_get_batch_size(h)))

    def save_weights(self, path) -> None:
        self._check(self.lib.bcnn_save_weights(self.handle, _b(str(path))), "bcnn_save_weights")

    def load_weights(self, path) -> int:
        """Returns the bcnn_status (0 = BCNN_SUCCESS); does not raise, so tests can probe the
        error paths."""
        return int(self.lib.bcnn_load_weights(self.handle, _b(str(path))))

    def lr_policy(self, decay_type: int, gamma=0.0, scale=0.0, power=0.0, max_batches=0, step=0):
        """bcnn_set_learning_rate_policy; decay_type: 0 constant, 1 step, 2 inv, 3 exp, 4 poly,
        5 sigmoid (bcnn_lr_decay)."""
        self.lib.bcnn_set_learning_rate_policy(self.handle, decay_type, gamma, scale, power,
                                               max_batches, step)

    def compile(self):
        self._check(self.lib.bcnn_compile_net(self.handle), "bcnn_compile_net")

    def set_mode(self, mode):
        self.lib.bcnn_set_mode(self.handle, mode)
        self.mode = mode

    def add_input(self, w, h, c, name):
        self._check(self.lib.bcnn_add_input(self.handle, w, h, c, _b(name)), f"add_input {name}")

    def train_on_batch(self) -> float:
        """bcnn_train_on_batch. On the B200 flavour the batch is taken from the host mirrors
        (fill them with set_host); the CPU reference needs a file loader and cannot run this."""
        return float(self.lib.bcnn_train_on_batch(self.handle))

    def predict_on_batch(self):
        """bcnn_predict_on_batch -> (loss, output array)."""
        ttype = TensorB200 if self.flavour == "b200" else TensorCPU
        out = C.POINTER(ttype)()
        loss = float(self.lib.bcnn_predict_on_batch(self.handle, C.byref(out)))
        t = out.contents
        size = t.n * t.c * t.h * t.w
        arr = np.ctypeslib.as_array(t.data, shape=(size,)).reshape(t.n, t.c, t.h, t.w).copy()
        return loss, arr

    # -- the three loops --
    def forward(self):
        self.lib.bcnn_forward(self.handle)

    def backward(self):
        self.lib.bcnn_backward(self.handle)

    def update(self):
        self.lib.bcnn_update(self.handle)

    # -- tensors --
    def tensor_index(self, name: str) -> int:
        return self.lib.bcnn_get_tensor_index_by_name(self.handle, _b(name))

    def _tensor(self, name_or_index):
        if isinstance(name_or_index, int):
            t = self.lib.bcnn_get_tensor_by_index(self.handle, name_or_index)
        else:
            t = self.lib.bcnn_get_tensor_by_name(self.handle, _b(name_or_index))
        if not t:
            raise KeyError(name_or_index)
        return t.contents

    def shape(self, name) -> tuple:
        t = self._tensor(name)
        return (t.n, t.c, t.h, t.w)

    def get(self, name, grad: bool = False) -> np.ndarray:
        """Copy of a tensor's data (or gradient) as an NCHW float32 array. On the B200
        flavour the getter refreshes the host mirrors from the device first."""
        t = self._tensor(name)
        size = t.n * t.c * t.h * t.w
        ptr = t.grad_data if grad else t.data
        if not ptr:
            raise ValueError(f"tensor {name!r} has no {'gradient' if grad else 'data'} buffer")
        return np.ctypeslib.as_array(ptr, shape=(size,)).reshape(t.n, t.c, t.h, t.w).copy()

    def set(self, name, array: np.ndarray, grad: bool = False) -> None:
        """Write a tensor's host mirror (and push it to the device on the B200 flavour)."""
        idx = name if isinstance(name, int) else self.tensor_index(name)
        if idx < 0:
            raise KeyError(name)
        t = self._tensor(idx)
        size = t.n * t.c * t.h * t.w
        arr = np.ascontiguousarray(array, dtype=np.float32).reshape(-1)
        if arr.size != size:
            raise ValueError(f"tensor {name!r}: expected {size} elements, got {arr.size}")
        ptr = t.grad_data if grad else t.data
        if not ptr:
            raise ValueError(f"tensor {name!r} has no host buffer to write")
        C.memmove(ptr, arr.ctypes.data, size * 4)
        if self.flavour == "b200":
            self._check(self.lib.bcnn_b200_upload_tensor(self.handle, idx), "upload")

    def tensor_names(self):
        names = []
        if self.flavour != "b200":
            raise NotImplementedError
        for i in range(self.lib.bcnn_b200_num_tensors(self.handle)):
            names.append(self._tensor(i).name.decode())
        return names

    # -- B200-only helpers --
    def set_conv_math(self, math: int):
        self.lib.bcnn_b200_set_conv_math(self.handle, math)

    def set_graphs(self, on: bool):
        self.lib.bcnn_b200_set_graphs(self.handle, int(on))

    def graphs(self) -> int:
        """0 = off, 1 = on, 2 = on and the forward graph is live."""
        return int(self.lib.bcnn_b200_get_graphs(self.handle))

    def set_reference_quirks(self, on: bool):
        self.lib.bcnn_b200_set_reference_quirks(self.handle, int(on))

    def sync(self):
        self.lib.bcnn_b200_sync(self.handle)

    def loss(self) -> float:
        return float(self.lib.bcnn_b200_get_loss(self.handle))

    def train_step(self, upload_inputs=False, fetch_loss=False) -> float:
        """upload_inputs: False / True (upload, then step) / 2 (input pipeline, see
        bcnn_b200_net.h: consumes the staged batch, uploads the mirrors for the next call)."""
        return float(self.lib.bcnn_b200_train_step(self.handle, int(upload_inputs),
                                                   int(fetch_loss)))

    def set_host(self, name, array: np.ndarray) -> None:
        """Fill a tensor's pinned host mirror only (the input pipeline uploads it)."""
        idx = name if isinstance(name, int) else self.tensor_index(name)
        t = self._tensor(idx)
        arr = np.ascontiguousarray(array, dtype=np.float32).reshape(-1)
        if arr.size != t.n * t.c * t.h * t.w or not t.data:
            raise ValueError(f"tensor {name!r}: bad size or no host mirror")
        C.memmove(t.data, arr.ctypes.data, arr.size * 4)

    def prefetch_inputs(self) -> int:
        return int(self.lib.bcnn_b200_prefetch_inputs(self.handle))

    def num_nodes(self) -> int:
        return self.lib.bcnn_b200_num_nodes(self.handle)

    def node_type(self, i) -> int:
        return self.lib.bcnn_b200_node_type(self.handle, i)

    def tensor_shape(self, index: int) -> tuple:
        """(n, c, h, w) of tensor `index` without a device -> host refresh of its buffers."""
        dims = (C.c_int * 4)()
        if self.lib.bcnn_b200_tensor_dims(self.handle, index, dims) != 0:
            raise KeyError(index)
        return tuple(dims)

    def maxpool_indexes(self, node: int) -> np.ndarray:
        dst = self.lib.bcnn_b200_node_dst(self.handle, node, 0)
        t = self._tensor(dst)
        out = np.empty(t.n * t.c * t.h * t.w, dtype=np.int32)
        n = self.lib.bcnn_b200_maxpool_indexes(self.handle, node, out.ctypes.data)
        if n < 0:
            raise ValueError(f"node {node} is not a maxpool node")
        return out.reshape(t.n, t.c, t.h, t.w)

    def bn_saved_stats(self, node: int):
        src_dst = self.lib.bcnn_b200_node_dst(self.handle, node, 0)
        c = self._tensor(src_dst).c
        mean = np.empty(c, dtype=np.float32)
        var = np.empty(c, dtype=np.float32)
        n = self.lib.bcnn_b200_bn_saved_stats(self.handle, node, mean.ctypes.data,
                                              var.ctypes.data)
        if n < 0:
            raise ValueError(f"node {node} has no batchnorm statistics")
        return mean, var


# --------------------------------------------------------------------------------
# kernel-level C ABI (include/bcnn_b200.h) -- used by the parity tests and bench.py
# --------------------------------------------------------------------------------

class ConvDesc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("batch", "cin", "h", "w", "cout", "ho", "wo", "ksize",
                                       "stride", "pad", "groups")]

    @classmethod
    def make(cls, batch, cin, h, w, cout, ksize, stride, pad, groups=1):
        ho = (h + 2 * pad - ksize) // stride + 1
        wo = (w + 2 * pad - ksize) // stride + 1
        return cls(batch, cin, h, w, cout, ho, wo, ksize, stride, pad, groups)


class ConvShadows(C.Structure):
    """bcnn_b200_conv_shadows: NHWC shadow storage kept between the passes of one layer."""
    _fields_ = [("x", C.c_void_p), ("x_bytes", C.c_size_t), ("x_fmt", C.c_int),
                ("dy", C.c_void_p), ("dy_bytes", C.c_size_t), ("dy_fmt", C.c_int)]


def bind_kernel_abi(lib: C.CDLL) -> None:
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    dp = C.POINTER(ConvDesc)
    shp = C.POINTER(ConvShadows)
    sigs = {
        "bcnn_b200_axpy": (i, [vp, vp, sz, f, vp]),
        "bcnn_b200_maxpool_forward": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, vp]),
        "bcnn_b200_maxpool_backward": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, vp]),
        "bcnn_b200_avgpool_forward": (i, [vp, vp, i, i, vp]),
        "bcnn_b200_avgpool_backward": (i, [vp, vp, i, i, vp]),
        "bcnn_b200_activation_forward": (i, [vp, i, i, vp, i, i, vp]),
        "bcnn_b200_activation_backward": (i, [vp, vp, i, i, vp, vp, i, i, vp]),
        "bcnn_b200_add_bias": (i, [vp, vp, i, i, i, vp]),
        "bcnn_b200_actbwd_grad_bias": (i, [vp, vp, vp, i, i, i, i, vp, vp]),
        "bcnn_b200_bn_scratch_floats": (sz, [i]),
        "bcnn_b200_bn_stats": (i, [vp, i, i, i, vp, vp, vp, vp, vp, vp]),
        "bcnn_b200_bn_apply": (i, [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]),
        "bcnn_b200_scale_bias": (i, [vp, vp, vp, vp, i, i, i, i, vp]),
        "bcnn_b200_bn_backward": (i, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i,
                                      vp, vp]),
        "bcnn_b200_conv_workspace_bytes": (sz, [dp, i]),
        "bcnn_b200_conv_uses_tensor_cores": (i, [dp, i]),
        "bcnn_b200_conv_forward": (i, [dp, vp, vp, vp, i, vp, vp, sz, i, vp]),
        "bcnn_b200_conv_backward_data": (i, [dp, vp, vp, vp, i, vp, sz, i, vp]),
        "bcnn_b200_conv_backward_weights": (i, [dp, vp, vp, vp, vp, sz, i, vp]),
        "bcnn_b200_conv_x_shadow_bytes": (sz, [dp, i]),
        "bcnn_b200_conv_dy_shadow_bytes": (sz, [dp, i]),
        "bcnn_b200_conv_forward_sh": (i, [dp, vp, vp, vp, i, vp, vp, sz, i, shp, vp]),
        "bcnn_b200_conv_forward_bn_stats": (i, [dp, vp, vp, vp, vp, sz, i, shp, vp, vp, vp, vp, vp, vp]),
        "bcnn_b200_conv_backward_data_sh": (i, [dp, vp, vp, vp, i, vp, sz, i, shp, vp]),
        "bcnn_b200_conv_backward_weights_sh": (i, [dp, vp, vp, vp, vp, sz, i, shp, vp]),
        "bcnn_b200_depthwise_forward": (i, [vp, vp, vp, i, vp, i, i, i, i, i, i, i, vp]),
        "bcnn_b200_depthwise_backward": (i, [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp, sz,
                                             vp]),
        "bcnn_b200_depthwise_scratch_floats": (sz, [i, i, i]),
        "bcnn_b200_sgd_update": (i, [vp, vp, sz, f, f, f, vp]),
        "bcnn_b200_adam_update": (i, [vp, vp, vp, vp, sz, f, f, f, f, vp]),
        "bcnn_b200_yolo_activate": (i, [vp, vp, i, i, i, i, i, vp]),
        "bcnn_b200_yolo_cost_scratch_floats": (i, []),
        "bcnn_b200_yolo_loss_forward": (i, [vp, vp, vp, vp, vp, vp] + [i] * 10 + [vp]),
        "bcnn_b200_softmax_forward": (i, [vp, vp, i, i, i, vp]),
        "bcnn_b200_cost_forward": (i, [vp, vp, vp, vp, i, i, i, vp]),
        "bcnn_b200_eltwise_forward": (i, [vp, vp, vp, i, i, i, vp]),
        "bcnn_b200_eltwise_backward": (i, [vp, vp, vp, vp, i, i, i, i, vp]),
        "bcnn_b200_concat_forward": (i, [vp, vp, i, i, i, i, vp]),
        "bcnn_b200_concat_backward": (i, [vp, vp, i, i, i, i, i, vp]),
        "bcnn_b200_upsample_forward": (i, [vp, vp, i, i, i, i, i, vp]),
        "bcnn_b200_upsample_backward": (i, [vp, vp, i, i, i, i, i, i, vp]),
        # BF16 NHWC resident kernels (csrc/nhwc_bf16.cu)
        "bcnn_b200_f32nchw_to_bf16nhwc": (i, [vp, vp, i, i, i, vp]),
        "bcnn_b200_bf16nhwc_to_f32nchw": (i, [vp, vp, i, i, i, vp]),
        "bcnn_b200_nhwc_scratch_floats": (sz, [i]),
        "bcnn_b200_bn_apply_nhwc": (i, [vp, vp, vp, vp, vp, vp, sz, i, i, vp]),
        "bcnn_b200_bn_backward_nhwc": (i, [vp] * 11 + [sz, i, i, vp, vp]),
        "bcnn_b200_actbwd_grad_bias_nhwc": (i, [vp, vp, vp, i, sz, i, vp, vp]),
        "bcnn_b200_eltwise_forward_bf16": (i, [vp, vp, vp, sz, sz, i, vp]),
        "bcnn_b200_eltwise_backward_bf16": (i, [vp, vp, vp, vp, sz, sz, i, i, vp]),
        "bcnn_b200_eltwise_backward_bn_reduce_bf16": (i, [vp, vp, vp, vp, sz, i, i, i, vp, vp, vp, vp, vp, vp,
                                                          C.POINTER(C.c_int), vp]),
        "bcnn_b200_bn_backward_nhwc_partials": (i, [vp] * 11 + [sz, i, vp, i, vp]),
        "bcnn_b200_maxpool_forward_nhwc": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, vp]),
        "bcnn_b200_maxpool_backward_nhwc": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]),
        "bcnn_b200_avgpool_forward_nhwc": (i, [vp, vp, i, i, i, vp]),
        "bcnn_b200_avgpool_backward_nhwc": (i, [vp, vp, i, i, i, i, vp]),
        "bcnn_b200_bn_stats_nhwc": (i, [vp, sz, i, vp, vp, vp, vp, vp, vp, vp]),
        "bcnn_b200_bn_add_act_nhwc": (i, [vp] * 11 + [sz, i, i, vp]),
        # convolution on resident BF16 NHWC tensors (csrc/conv_tma.cu)
        "bcnn_b200_conv_nhwc_supported": (i, [dp]),
        "bcnn_b200_conv_nhwc_workspace_bytes": (sz, [dp]),
        "bcnn_b200_conv_nhwc_x_keep_bytes": (sz, [dp]),
        "bcnn_b200_conv_forward_nhwc": (i, [dp, vp, vp, vp, i, vp, vp, sz, shp, vp]),
        "bcnn_b200_conv_forward_bn_stats_nhwc": (i, [dp, vp, vp, vp, vp, sz, shp, vp, vp, vp, vp, vp, vp, vp]),
        "bcnn_b200_conv_backward_data_nhwc": (i, [dp, vp, vp, vp, i, vp, sz, vp]),
        "bcnn_b200_conv_backward_weights_nhwc": (i, [dp, vp, vp, vp, vp, sz, shp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args


class DeviceBuffer:
    """A cudaMalloc'd buffer owned through the C ABI (tests / bench plumbing)."""

    def __init__(self, array: np.ndarray | None = None, *, nbytes: int | None = None,
                 lib: C.CDLL | None = None):
        self.lib = lib or b200()
        if array is not None:
            array = np.ascontiguousarray(array)
            nbytes = array.nbytes
        self.nbytes = int(nbytes)
        self.ptr = self.lib.bcnn_b200_malloc(max(self.nbytes, 4))
        if not self.ptr:
            raise MemoryError(f"bcnn_b200_malloc({self.nbytes}) failed")
        if array is not None and self.nbytes:
            self.upload(array)

    def upload(self, array: np.ndarray):
        array = np.ascontiguousarray(array)
        err = self.lib.bcnn_b200_memcpy_h2d(self.ptr, array.ctypes.data, array.nbytes, None)
        err = err or self.lib.bcnn_b200_stream_sync(None)
        if err:
            raise RuntimeError(self.lib.bcnn_b200_error_string(err).decode())

    def download(self, dtype=np.float32, shape=None) -> np.ndarray:
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        err = self.lib.bcnn_b200_stream_sync(None)
        err = err or self.lib.bcnn_b200_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, None)
        err = err or self.lib.bcnn_b200_stream_sync(None)
        if err:
            raise RuntimeError(self.lib.bcnn_b200_error_string(err).decode())
        return out.reshape(shape) if shape is not None else out

    def free(self):
        if self.ptr:
            self.lib.bcnn_b200_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
