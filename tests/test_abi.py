"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
declared in include/*.h, keeps the reference's internal layer entry points, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from bcnn_b200 import capi

ROOT = Path(__file__).resolve().parents[1]
DECL = re.compile(r"BCNN(?:_B200)?_API\s+[^;{]*?\b(bcnn\w+)\s*\(", re.S)


def declared_symbols():
    names = []
    for header in sorted((ROOT / "include").rglob("*.h")):
        names += [(header.name, n) for n in DECL.findall(header.read_text())]
    return names


def test_headers_declare_the_expected_surface():
    names = {n for _, n in declared_symbols()}
    for must in ("bcnn_init_net", "bcnn_forward", "bcnn_backward", "bcnn_update",
                 "bcnn_add_convolutional_layer", "bcnn_add_batchnorm_layer",
                 "bcnn_add_maxpool_layer", "bcnn_add_avgpool_layer", "bcnn_add_activation_layer",
                 "bcnn_add_depthwise_conv_layer", "bcnn_b200_conv_forward",
                 "bcnn_b200_conv_backward_data", "bcnn_b200_conv_backward_weights",
                 "bcnn_b200_bn_stats", "bcnn_b200_bn_backward", "bcnn_b200_maxpool_forward",
                 "bcnn_b200_dp_init",
                 # callers and data formats either side of the path
                 "bcnn_load_net", "bcnn_load_weights", "bcnn_save_weights", "bcnn_train_on_batch",
                 "bcnn_predict_on_batch", "bcnn_add_input", "bcnn_add_yolo_layer",
                 "bcnn_yolo_get_detections", "bcnn_add_concat_layer", "bcnn_add_upsample_layer",
                 "bcnn_set_adam_optimizer", "bcnn_net_set_param", "bcnn_b200_adam_update",
                 "bcnn_b200_yolo_activate", "bcnn_b200_yolo_loss_forward", "bcnn_b200_set_graphs"):
        assert must in names
    assert len(names) >= 120


@pytest.mark.parametrize("header,name", declared_symbols())
def test_library_exports_every_declared_symbol(header, name):
    lib = capi.load_library()
    assert hasattr(lib, name), f"{name} (declared in {header}) is not exported"


@pytest.mark.parametrize("name", [
    # internal, non-static layer entry points the reference declares in its layer headers
    # (SURVEY.md 8b): kept so its runtime could link against this layer library
    "bcnn_forward_conv_layer", "bcnn_backward_conv_layer", "bcnn_update_conv_layer",
    "bcnn_release_param_conv_layer", "bcnn_forward_batchnorm_layer",
    "bcnn_backward_batchnorm_layer", "bcnn_release_param_batchnorm_layer",
    "bcnn_forward_batchnorm_gpu", "bcnn_backward_batchnorm_gpu", "bcnn_forward_maxpool_layer",
    "bcnn_backward_maxpool_layer", "bcnn_forward_maxpool_layer_gpu",
    "bcnn_backward_maxpool_layer_gpu", "bcnn_release_param_maxpool_layer",
    "bcnn_forward_avgpool_layer", "bcnn_backward_avgpool_layer", "bcnn_forward_avgpool_layer_gpu",
    "bcnn_backward_avgpool_layer_gpu", "bcnn_forward_activation_layer",
    "bcnn_backward_activation_layer", "bcnn_update_activation_layer",
    "bcnn_forward_activation_gpu", "bcnn_backward_activation_gpu",
    "bcnn_b200_forward_batchnorm", "bcnn_b200_backward_batchnorm",
    "bcnn_forward_depthwise_conv_layer", "bcnn_backward_depthwise_conv_layer",
    "bcnn_update_depthwise_conv_layer", "bcnn_release_param_depthwise_conv_layer",
    "bcnn_sgd_update_gpu", "bcnn_net_add_node", "bcnn_net_add_tensor", "bcnn_node_add_input",
    "bcnn_node_add_output", "bcnn_tensor_create", "bcnn_tensor_size",
])
def test_internal_layer_entry_points_are_exported(name):
    assert hasattr(capi.load_library(), name)


def test_tensor_struct_layout_is_the_cuda_flavour():
    # n,c,h,w,has_grad (5 ints, padded to 24) + name + data + grad + data_gpu + grad_gpu
    assert C.sizeof(capi.TensorB200) == 24 + 5 * 8
    assert capi.TensorB200.data_gpu.offset == 48
    assert C.sizeof(capi.TensorCPU) == 24 + 3 * 8


def test_no_cpu_fallback_without_a_device():
    lib = capi.b200()
    if lib.bcnn_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        capi.Net()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        capi.load_library(tmp_path / "libbcnn_b200.so")


def test_product_does_not_reference_the_oracle():
    """Nothing under bcnn_b200/ may import, link or load anything under oracle/."""
    for path in (ROOT / "bcnn_b200").rglob("*"):
        if path.suffix in {".py", ".c", ".h", ".cu", ".cuh"}:
            text = path.read_text()
            assert "liboracle" not in text and "libbcnn_ref" not in text, path
            assert "bcnn_oracle" not in text and "orc_" not in text, path
