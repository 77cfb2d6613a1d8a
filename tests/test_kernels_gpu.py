"""GPU parity of every kernel-level C-ABI entry point (include/bcnn_b200.h) against the
oracle restatement (oracle/bcnn_oracle.c) on identical seeded inputs.

Bar: bit-exact for max-pool values and int32 argmax; 1e-5 normalised (tests/helpers.py
rel_err) for floating point on the FP32 path; 2e-2 on the tensor-core path.
"""
import ctypes as C

import numpy as np
import pytest

from bcnn_b200 import capi
from helpers import FP32_TOL, TC_TOL, assert_close, check, dev, dev_zeros, f32, oracle, p, rel_err

pytestmark = pytest.mark.gpu
ACT = capi.ACT


def rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------ max pooling
POOL_CASES = [
    # n, c, h, w, k, stride, padding
    (2, 3, 28, 28, 2, 2, capi.PAD_SAME),    # C1/C2 shape class (vector path)
    (2, 4, 13, 13, 2, 1, capi.PAD_SAME),    # yolo-tiny k2 s1: bottom/right -FLT_MAX taps
    (2, 4, 112, 112, 3, 2, capi.PAD_SAME),  # resnet stem pool, overlapping windows
    (1, 2, 15, 17, 2, 2, capi.PAD_SAME),    # odd sizes (ragged right/bottom)
    (1, 2, 15, 17, 3, 2, capi.PAD_VALID),
    (1, 2, 15, 17, 3, 2, capi.PAD_CAFFE),
    (3, 5, 8, 12, 2, 2, capi.PAD_VALID),
    (1, 1, 1, 1, 2, 2, capi.PAD_SAME),      # degenerate 1x1
    (2, 3, 26, 28, 2, 2, capi.PAD_SAME),    # vector path with h even, w % 4 == 0
    (2, 3, 27, 28, 2, 2, capi.PAD_SAME),    # vector path with odd h (missing bottom row)
    (2, 3, 24, 32, 3, 2, capi.PAD_SAME),    # k3 s2 quad kernels (w % 8 == 0), bottom row missing
    (2, 3, 25, 40, 3, 2, capi.PAD_SAME),    # k3 s2 quad kernels, odd h
    (1, 2, 14, 20, 3, 2, capi.PAD_SAME),    # k3 s2: forward generic (w % 8 != 0), backward quads
    (1, 2, 9, 8, 3, 2, capi.PAD_SAME),      # k3 s2: a single quad per output row
]


@pytest.mark.parametrize("case", POOL_CASES)
@pytest.mark.parametrize("ties", ["random", "relu_zeros", "constant"])
def test_maxpool_forward_backward_bit_exact(case, ties):
    n, c, h, w, k, s, pad = case
    lib, orc = capi.b200(), oracle()
    r = rng(hash(case) % 2**31)
    x = f32(r.uniform(-1, 1, size=(n, c, h, w)))
    if ties == "relu_zeros":
        x = np.maximum(x, 0).astype(np.float32)     # many exact-zero ties
    elif ties == "constant":
        x = np.full_like(x, 0.25)                   # every window is a tie
    ho = orc.orc_maxpool_out_dim(h, k, s, pad)
    wo = orc.orc_maxpool_out_dim(w, k, s, pad)
    y_ref = np.zeros((n, c, ho, wo), np.float32)
    i_ref = np.zeros((n, c, ho, wo), np.int32)
    orc.orc_maxpool_forward(p(x), p(y_ref), p(i_ref), n, c, h, w, k, s, ho, wo)

    dx, dy_, di = dev(x), dev_zeros(y_ref.size), dev_zeros(i_ref.size)
    check(lib.bcnn_b200_maxpool_forward(dx.ptr, dy_.ptr, di.ptr, n, c, h, w, k, s, ho, wo, None))
    y = dy_.download(np.float32, y_ref.shape)
    idx = di.download(np.int32, i_ref.shape)
    assert np.array_equal(idx, i_ref), "argmax indices must be bit-exact"
    assert np.array_equal(y.view(np.uint32), y_ref.view(np.uint32)), "max values must be bit-exact"

    # backward: dx += scatter(dy) on a non-zero dx (the += contract)
    g = f32(r.uniform(-1, 1, size=y_ref.shape))
    gx0 = f32(r.uniform(-1, 1, size=x.shape))
    gx_ref = gx0.copy()
    orc.orc_maxpool_backward(p(gx_ref), p(g), p(i_ref), g.size)
    dgx, dg = dev(gx0), dev(g)
    check(lib.bcnn_b200_maxpool_backward(dgx.ptr, dg.ptr, di.ptr, n, c, h, w, k, s, ho, wo, None))
    gx = dgx.download(np.float32, x.shape)
    assert np.array_equal(gx.view(np.uint32), gx_ref.view(np.uint32)), \
        "scatter-add must reproduce the CPU summation order bit for bit"


def test_maxpool_all_minus_flt_max_gives_index_minus_one():
    lib, orc = capi.b200(), oracle()
    x = np.full((1, 1, 4, 4), -np.finfo(np.float32).max, np.float32)
    y_ref = np.zeros((1, 1, 2, 2), np.float32)
    i_ref = np.zeros((1, 1, 2, 2), np.int32)
    orc.orc_maxpool_forward(p(x), p(y_ref), p(i_ref), 1, 1, 4, 4, 2, 2, 2, 2)
    assert (i_ref == -1).all()
    dx, dy_, di = dev(x), dev_zeros(4), dev_zeros(4)
    check(lib.bcnn_b200_maxpool_forward(dx.ptr, dy_.ptr, di.ptr, 1, 1, 4, 4, 2, 2, 2, 2, None))
    assert np.array_equal(di.download(np.int32), i_ref.ravel())
    # backward must skip the -1 entries instead of writing out of bounds
    dgx, dg = dev_zeros(16), dev(np.ones(4, np.float32))
    check(lib.bcnn_b200_maxpool_backward(dgx.ptr, dg.ptr, di.ptr, 1, 1, 4, 4, 2, 2, 2, 2, None))
    assert (dgx.download() == 0).all()


# ------------------------------------------------------------------ avg pooling
@pytest.mark.parametrize("n,c,hw", [(2, 1024, 49), (3, 7, 64), (1, 5, 1), (4, 33, 169)])
def test_avgpool(n, c, hw):
    lib, orc = capi.b200(), oracle()
    r = rng(n * 1000 + c)
    x = f32(r.uniform(-1, 1, size=(n, c, hw)))
    y_ref = np.zeros((n, c), np.float32)
    orc.orc_avgpool_forward(p(x), p(y_ref), n, c, hw)
    dx, dy_ = dev(x), dev_zeros(n * c)
    check(lib.bcnn_b200_avgpool_forward(dx.ptr, dy_.ptr, n * c, hw, None))
    assert_close(dy_.download(np.float32, (n, c)), y_ref, FP32_TOL, "avgpool fwd")
    g = f32(r.uniform(-1, 1, size=(n, c)))
    gx0 = f32(r.uniform(-1, 1, size=x.shape))
    gx_ref = gx0.copy()
    orc.orc_avgpool_backward(p(gx_ref), p(g), n, c, hw)
    dgx, dg = dev(gx0), dev(g)
    check(lib.bcnn_b200_avgpool_backward(dgx.ptr, dg.ptr, n * c, hw, None))
    assert_close(dgx.download(np.float32, x.shape), gx_ref, FP32_TOL, "avgpool bwd")


# ------------------------------------------------------------------ activations
@pytest.mark.parametrize("act", ["tanh", "relu", "ramp", "softplus", "lrelu", "abs", "clamp",
                                 "prelu", "logistic"])
@pytest.mark.parametrize("shape", [(2, 6, 8, 8), (3, 5, 7, 7), (1, 1, 1, 3)])
def test_activation_forward_backward(act, shape):
    lib, orc = capi.b200(), oracle()
    n, c, h, w = shape
    r = rng(len(act) * 97 + n)
    x = f32(r.uniform(-2, 2, size=shape))
    x.ravel()[:: 7] = 0.0  # exact zeros: the x > 0 / x >= 0 edges
    slope = f32(r.uniform(0.05, 0.4, size=c))
    y_ref = x.copy()
    orc.orc_activation_forward(p(y_ref), x.size, p(slope), h * w, c, ACT[act])
    dx, dslope = dev(x), dev(slope)
    check(lib.bcnn_b200_activation_forward(dx.ptr, x.size, ACT[act],
                                           dslope.ptr if act == "prelu" else None, h * w, c, None))
    y = dx.download(np.float32, shape)
    assert_close(y, y_ref, FP32_TOL, f"{act} fwd")

    g = f32(r.uniform(-1, 1, size=shape))
    g_ref = g.copy()
    gs_ref = f32(r.uniform(-1, 1, size=c))
    gs0 = gs_ref.copy()
    orc.orc_activation_backward(p(y_ref), p(g_ref), x.size, p(slope), p(gs_ref), h * w, c, ACT[act])
    dyv, dg, dgs = dev(y_ref), dev(g), dev(gs0)
    check(lib.bcnn_b200_activation_backward(dyv.ptr, dg.ptr, x.size, ACT[act],
                                            dslope.ptr if act == "prelu" else None,
                                            dgs.ptr if act == "prelu" else None, h * w, c, None))
    assert_close(dg.download(np.float32, shape), g_ref, FP32_TOL, f"{act} bwd")
    if act == "prelu":
        assert_close(dgs.download(), gs_ref, FP32_TOL, "prelu slope grad")


# ------------------------------------------------------------------ bias kernels
@pytest.mark.parametrize("n,c,hw", [(4, 32, 196), (3, 7, 49), (2, 255, 169), (64, 10, 1)])
@pytest.mark.parametrize("act", ["none", "relu", "lrelu"])
def test_add_bias_and_fused_actbwd_grad_bias(n, c, hw, act):
    lib, orc = capi.b200(), oracle()
    r = rng(n * 31 + c)
    y0 = f32(r.uniform(-1, 1, size=(n, c, hw)))
    b = f32(r.uniform(-0.5, 0.5, size=c))
    y_ref = y0.copy()
    orc.orc_add_bias(p(y_ref), p(b), n, c, hw)
    dy_, db = dev(y0), dev(b)
    check(lib.bcnn_b200_add_bias(dy_.ptr, db.ptr, n, c, hw, None))
    assert_close(dy_.download(np.float32, y0.shape), y_ref, FP32_TOL, "add_bias")

    # fused: dy *= act'(y); gb += sum dy
    yv = y_ref.copy()
    orc.orc_activation_forward(p(yv), yv.size, None, hw, c, ACT[act])
    g = f32(r.uniform(-1, 1, size=(n, c, hw)))
    g_ref = g.copy()
    orc.orc_activation_backward(p(yv), p(g_ref), g.size, None, None, hw, c, ACT[act])
    gb0 = f32(r.uniform(-1, 1, size=c))
    gb_ref = gb0.copy()
    orc.orc_grad_bias(p(gb_ref), p(g_ref), n, c, hw)
    scratch = dev_zeros(lib.bcnn_b200_bn_scratch_floats(c))
    dg, dyv, dgb = dev(g), dev(yv), dev(gb0)
    for _ in range(2):  # twice: the ticket counters must re-arm
        dgb.upload(gb0)
        dg.upload(g)
        check(lib.bcnn_b200_actbwd_grad_bias(dgb.ptr, dg.ptr, dyv.ptr, ACT[act], n, c, hw,
                                             scratch.ptr, None))
        assert_close(dg.download(np.float32, g.shape), g_ref, FP32_TOL, "act-bwd in place")
        assert_close(dgb.download(), gb_ref, FP32_TOL, "grad_bias")


# ------------------------------------------------------------------ batchnorm
BN_SHAPES = [(8, 32, 28 * 28), (4, 64, 49), (16, 7, 13 * 13), (128, 512, 1), (2, 3, 5),
             (6, 256, 196)]


@pytest.mark.parametrize("n,c,hw", BN_SHAPES)
def test_bn_forward_train_and_valid(n, c, hw):
    lib, orc = capi.b200(), oracle()
    r = rng(n * 7 + c * 3 + hw)
    x = f32(r.normal(0.3, 1.2, size=(n, c, hw)))
    gamma = f32(r.uniform(0.5, 1.5, size=c))
    beta = f32(r.uniform(-0.3, 0.3, size=c))
    rm0 = f32(r.uniform(-0.2, 0.2, size=c))
    rv0 = f32(r.uniform(0.5, 1.5, size=c))
    for mode in (1, 2):  # TRAIN, VALID
        y_ref = x.copy()
        rm, rv = rm0.copy(), rv0.copy()
        sm, sv = np.zeros(c, np.float32), np.zeros(c, np.float32)
        xn, xc = np.zeros_like(x), np.zeros_like(x)
        orc.orc_bn_forward(p(y_ref), n, c, hw, p(rm), p(rv), p(gamma), p(beta), p(sm), p(sv),
                           p(xn), p(xc), mode)
        dx, dyo = dev(x), dev_zeros(x.size)
        dg, db, drm, drv = dev(gamma), dev(beta), dev(rm0), dev(rv0)
        dsm, dsv = dev_zeros(c), dev_zeros(c)
        scratch = dev_zeros(lib.bcnn_b200_bn_scratch_floats(c))
        if mode == 1:
            check(lib.bcnn_b200_bn_stats(dx.ptr, n, c, hw, dsm.ptr, dsv.ptr, drm.ptr, drv.ptr,
                                         scratch.ptr, None))
            check(lib.bcnn_b200_bn_apply(dx.ptr, dyo.ptr, dsm.ptr, dsv.ptr, dg.ptr, db.ptr, n, c,
                                         hw, 0, None))
            assert_close(dsm.download(), sm, FP32_TOL, "saved_mean")
            # var = E[x^2]-mean^2 cancels: normalise against the second moment's scale
            assert np.abs(dsv.download() - sv).max() <= 2e-5 * (np.abs(sv).max() + np.abs(sm).max() ** 2)
            assert_close(drm.download(), rm, FP32_TOL, "running mean")
            assert_close(drv.download(), rv, 2e-5, "running var")
        else:
            check(lib.bcnn_b200_bn_apply(dx.ptr, dyo.ptr, drm.ptr, drv.ptr, dg.ptr, db.ptr, n, c,
                                         hw, 0, None))
        assert_close(dyo.download(np.float32, x.shape), y_ref, 2e-5, f"bn fwd mode {mode}")


@pytest.mark.parametrize("n,c,hw", BN_SHAPES)
@pytest.mark.parametrize("act", ["none", "relu", "lrelu"])
@pytest.mark.parametrize("remask", [False, True])
def test_bn_backward_fused_with_activation(n, c, hw, act, remask):
    lib, orc = capi.b200(), oracle()
    r = rng(n * 5 + c * 11 + hw)
    x = f32(r.normal(0.1, 1.0, size=(n, c, hw)))
    gamma = f32(r.uniform(0.5, 1.5, size=c))
    beta = f32(r.uniform(-0.3, 0.3, size=c))
    y = x.copy()
    rm, rv = np.zeros(c, np.float32), np.zeros(c, np.float32)
    sm, sv = np.zeros(c, np.float32), np.zeros(c, np.float32)
    xn, xc = np.zeros_like(x), np.zeros_like(x)
    orc.orc_bn_forward(p(y), n, c, hw, p(rm), p(rv), p(gamma), p(beta), p(sm), p(sv), p(xn), p(xc), 1)
    orc.orc_activation_forward(p(y), y.size, None, hw, c, ACT[act])
    g = f32(r.uniform(-1, 1, size=x.shape))
    g_ref = g.copy()
    orc.orc_activation_backward(p(y), p(g_ref), g.size, None, None, hw, c, ACT[act])
    gg0, gb0 = f32(r.uniform(-1, 1, size=c)), f32(r.uniform(-1, 1, size=c))
    gg, gb = gg0.copy(), gb0.copy()
    dm, dv = np.zeros(c, np.float32), np.zeros(c, np.float32)
    orc.orc_bn_backward(p(g_ref), n, c, hw, p(gamma), p(gg), p(gb), p(sm), p(sv), p(dm), p(dv),
                        p(xn), p(xc))
    dx, dyv, dg = dev(x), dev(y), dev(g)
    dsm, dsv, dgam, dbeta = dev(sm), dev(sv), dev(gamma), dev(beta)
    dgg, dgb, ddm, ddv = dev(gg0), dev(gb0), dev_zeros(c), dev_zeros(c)
    scratch = dev_zeros(lib.bcnn_b200_bn_scratch_floats(c))
    # remask: the ReLU mask is rebuilt from x and beta, y is never read (pass a NULL y to prove it)
    y_ptr = dyv.ptr if act != "none" and not remask else None
    if remask and act == "none":
        pytest.skip("no activation: nothing to rebuild")
    check(lib.bcnn_b200_bn_backward(dx.ptr, y_ptr, dg.ptr, dg.ptr, dsm.ptr, dsv.ptr, dgam.ptr,
                                    dbeta.ptr if remask else None, dgg.ptr, dgb.ptr, ddm.ptr,
                                    ddv.ptr, n, c, hw, ACT[act], scratch.ptr, None))
    tol = 5e-5  # two chained FP32 reductions with a different (tree) summation order
    assert_close(dgb.download(), gb, tol, "g_beta")
    assert_close(dgg.download(), gg, tol, "g_gamma")
    assert_close(dg.download(np.float32, x.shape), g_ref, tol, "bn dx")


# ------------------------------------------------------------------ convolution
CONV_CASES = [
    # batch, cin, h, w, cout, k, stride, pad, groups
    (4, 1, 28, 28, 32, 3, 1, 1, 1),    # C1 conv1 (K = 9)
    (4, 32, 14, 14, 32, 3, 1, 1, 1),   # C1 conv2
    (2, 3, 32, 32, 32, 3, 1, 1, 1),    # C2 conv1 (K = 27)
    (2, 32, 16, 16, 64, 3, 1, 1, 1),   # C2 conv2_1
    (2, 3, 32, 32, 16, 7, 2, 3, 1),    # resnet stem class (7x7 s2 p3)
    (2, 64, 14, 14, 128, 3, 2, 1, 1),  # resnet 3x3 stride 2
    (2, 64, 14, 14, 256, 1, 1, 0, 1),  # 1x1 expand
    (2, 96, 14, 14, 48, 1, 2, 0, 1),   # 1x1 stride-2 shortcut
    (3, 16, 13, 13, 255, 1, 1, 0, 1),  # yolo head: Cout = 255, HW = 169 (odd tails)
    (2, 8, 9, 11, 12, 3, 1, 0, 2),     # groups = 2, pad 0, ragged
    (1, 5, 7, 7, 3, 5, 1, 2, 1),       # 5x5
    (5, 20, 1, 1, 33, 1, 1, 0, 1),     # FC-shaped
    # TMA-addressable shapes (row pitch % 16 B == 0, stride 1): TF32 tcgen05 + TMA path
    (2, 64, 56, 56, 64, 3, 1, 1, 1),   # resnet stage-1 3x3 (2 column chunks x 2 rows per tile)
    (2, 256, 28, 28, 128, 1, 1, 0, 1), # resnet 1x1 reduce (flat plane view, ragged last tile)
    (3, 48, 32, 32, 40, 3, 1, 1, 1),   # Cin, Cout not multiples of the 32-wide k-block / 16
    (1, 16, 8, 8, 16, 5, 1, 2, 1),     # 5x5 pad 2 on an 8-wide plane (box wider than the row)
    (2, 24, 12, 12, 24, 3, 1, 0, 1),   # pad 0: fprop / wgrad on TMA, dgrad (10-wide dY) gathered
    (1, 32, 104, 104, 64, 3, 1, 1, 1), # yolo-tiny conv2 class: 4 column chunks per tile
    (2, 160, 8, 8, 300, 1, 1, 0, 1),   # Cout > 128: several N tiles
    # thin first layers through the im2col buffer (fprop / wgrad as a 1x1 problem over Cin*k*k)
    (4, 3, 64, 64, 64, 7, 2, 3, 1),    # resnet stem: K = 147 -> 148 columns, two wgrad N tiles
    (8, 3, 32, 32, 32, 3, 1, 1, 1),    # cifar conv1: K = 27 -> 28 columns
    # strided dgrad as stride^2 sub-sampled stride-1 problems scattered into dX
    (2, 32, 15, 15, 32, 3, 2, 1, 1),   # odd extents: classes of different sizes
    (2, 32, 16, 16, 32, 3, 2, 0, 1),   # pad 0: trailing input row / column get zero gradient
    (1, 16, 20, 20, 16, 5, 3, 2, 1),   # stride 3, 5x5: nine classes, 1..4 taps each
    (2, 32, 12, 12, 32, 2, 2, 0, 1),   # k2 s2: one tap per class
    (2, 32, 9, 9, 16, 1, 2, 0, 1),     # 1x1 s2 on an odd plane: three empty classes
    (2, 64, 28, 28, 64, 3, 2, 1, 1),   # resnet stride-2 3x3 at a size with several tiles
    # fully-connected layers (1x1 plane): a handful of tiles with a long K -> split-K fprop / dgrad
    (64, 1024, 1, 1, 100, 1, 1, 0, 1),
    (16, 640, 2, 2, 24, 1, 1, 0, 1),
]


def _conv_reference(case, seed):
    batch, cin, h, w, cout, k, s, pad, groups = case
    orc = oracle()
    r = rng(seed)
    d = capi.ConvDesc.make(batch, cin, h, w, cout, k, s, pad, groups)
    x = f32(r.uniform(-1, 1, size=(batch, cin, h, w)))
    wt = f32(r.uniform(-1, 1, size=(cout, cin // groups, k, k)) * np.sqrt(3.0 / (cin // groups * k * k)))
    bias = f32(r.uniform(-0.2, 0.2, size=cout))
    y = np.zeros((batch, cout, d.ho, d.wo), np.float32)
    orc.orc_conv_forward(p(x), p(wt), p(y), batch, cin, h, w, cout, k, s, pad, groups)
    dy = f32(r.uniform(-1, 1, size=y.shape))
    gw0 = f32(r.uniform(-0.1, 0.1, size=wt.shape))
    gw = gw0.copy()
    dx = np.zeros_like(x)
    orc.orc_conv_backward(p(x), p(wt), p(dy), p(gw), p(dx), batch, cin, h, w, cout, k, s, pad, groups)
    return d, x, wt, bias, y, dy, gw0, gw, dx


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("math,tol", [(capi.MATH_FP32, FP32_TOL), (capi.MATH_TC, TC_TOL)])
def test_conv_fprop_dgrad_wgrad(case, math, tol):
    lib, orc = capi.b200(), oracle()
    d, x, wt, bias, y_ref, dy, gw0, gw_ref, dx_ref = _conv_reference(case, hash(case) % 2**31)
    ws_bytes = lib.bcnn_b200_conv_workspace_bytes(d, math)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
    dxv, dwt, dyo = dev(x), dev(wt), dev_zeros(y_ref.size)
    # raw forward
    check(lib.bcnn_b200_conv_forward(d, dxv.ptr, dwt.ptr, None, 0, dyo.ptr, ws.ptr, ws_bytes, math, None))
    y_gpu = dyo.download(np.float32, y_ref.shape)
    assert_close(y_gpu, y_ref, tol, "fprop")
    if math == capi.MATH_TC and lib.bcnn_b200_conv_uses_tensor_cores(d, 0):
        # the tensor-core path really ran: BF16 operands leave a ~1e-3 signature
        assert rel_err(y_gpu, y_ref)[1] > 1e-5, "TC path produced FP32-exact results?"
    # fused bias + relu epilogue
    yb = y_ref.copy()
    orc.orc_add_bias(p(yb), p(bias), d.batch, d.cout, d.ho * d.wo)
    orc.orc_activation_forward(p(yb), yb.size, None, d.ho * d.wo, d.cout, ACT["relu"])
    db = dev(bias)
    check(lib.bcnn_b200_conv_forward(d, dxv.ptr, dwt.ptr, db.ptr, ACT["relu"], dyo.ptr, ws.ptr,
                                     ws_bytes, math, None))
    assert_close(dyo.download(np.float32, y_ref.shape), yb, tol, "fprop + bias + relu")
    # wgrad accumulates on top of gw0
    ddy, dgw = dev(dy), dev(gw0)
    check(lib.bcnn_b200_conv_backward_weights(d, dxv.ptr, ddy.ptr, dgw.ptr, ws.ptr, ws_bytes, math, None))
    assert_close(dgw.download(np.float32, wt.shape), gw_ref, tol, "wgrad (+=)")
    # dgrad overwrites garbage
    ddx = dev(np.full(x.shape, 7.0, np.float32))
    check(lib.bcnn_b200_conv_backward_data(d, dwt.ptr, ddy.ptr, ddx.ptr, 0, ws.ptr, ws_bytes, math, None))
    assert_close(ddx.download(np.float32, x.shape), dx_ref, tol, "dgrad (overwrite)")
    # dgrad accumulate flavour (fully-connected layer contract)
    base = np.full(x.shape, 0.5, np.float32)
    ddx.upload(base)
    check(lib.bcnn_b200_conv_backward_data(d, dwt.ptr, ddy.ptr, ddx.ptr, 1, ws.ptr, ws_bytes, math, None))
    assert_close(ddx.download(np.float32, x.shape), dx_ref + base, tol, "dgrad (+=)")


def test_conv_linearity_at_resnet_size():
    """Size-independent property at a BASELINE-size layer (too slow for the scalar oracle):
    conv(a*x1 + x2) == a*conv(x1) + conv(x2), and dgrad/wgrad adjointness
    <conv(x), dy> == <x, dgrad(dy)> == <W, wgrad(x, dy)>."""
    lib = capi.b200()
    case = (8, 64, 56, 56, 64, 3, 1, 1, 1)
    batch, cin, h, w, cout, k, s, pad, groups = case
    d = capi.ConvDesc.make(*case)
    r = rng(99)
    x1 = f32(r.uniform(-1, 1, size=(batch, cin, h, w)))
    x2 = f32(r.uniform(-1, 1, size=(batch, cin, h, w)))
    wt = f32(r.uniform(-1, 1, size=(cout, cin, k, k)) / np.sqrt(cin * k * k))
    dy = f32(r.uniform(-1, 1, size=(batch, cout, d.ho, d.wo)))
    ws_bytes = lib.bcnn_b200_conv_workspace_bytes(d, capi.MATH_FP32)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
    dw = dev(wt)

    def fwd(x):
        dxv, dyo = dev(x), dev_zeros(dy.size)
        check(lib.bcnn_b200_conv_forward(d, dxv.ptr, dw.ptr, None, 0, dyo.ptr, ws.ptr, ws_bytes,
                                         capi.MATH_FP32, None))
        return dyo.download(np.float32, dy.shape)

    y1, y2, y12 = fwd(x1), fwd(x2), fwd(f32(0.5 * x1 + x2))
    assert_close(y12, 0.5 * y1 + y2, 2e-5, "linearity")
    ddy, ddx, dgw = dev(dy), dev_zeros(x1.size), dev_zeros(wt.size)
    check(lib.bcnn_b200_conv_backward_data(d, dw.ptr, ddy.ptr, ddx.ptr, 0, ws.ptr, ws_bytes,
                                           capi.MATH_FP32, None))
    dxv = dev(x1)
    check(lib.bcnn_b200_conv_backward_weights(d, dxv.ptr, ddy.ptr, dgw.ptr, ws.ptr, ws_bytes,
                                              capi.MATH_FP32, None))
    lhs = float(np.sum(y1.astype(np.float64) * dy))
    via_dx = float(np.sum(x1.astype(np.float64) * ddx.download(np.float32, x1.shape)))
    via_gw = float(np.sum(wt.astype(np.float64) * dgw.download(np.float32, wt.shape)))
    assert abs(lhs - via_dx) <= 1e-4 * abs(lhs) and abs(lhs - via_gw) <= 1e-4 * abs(lhs)


@pytest.mark.parametrize("case", [
    (4, 32, 14, 14, 48, 3, 1, 1, 1),    # 3x3: x shadow fprop -> wgrad, dy shadow wgrad -> dgrad
    (8, 32, 16, 16, 64, 3, 2, 1, 1),    # strided dgrad classes read the kept dy shadow
    (16, 64, 7, 7, 32, 1, 1, 0, 1),     # 1x1 on an odd plane (49 positions): shadow route
    (8, 64, 8, 8, 32, 1, 1, 0, 1),      # DIRECT route: shadows must stay untouched
    (32, 3, 32, 32, 32, 7, 2, 3, 1),    # thin first layer: the im2col buffer is the x shadow
])
def test_conv_shadow_storage_matches_plain_entry_points(case):
    """The *_sh entry points keep NHWC shadows between passes (one transpose per operand and
    step); their results are bit-identical to the plain entry points, which transpose per pass."""
    lib = capi.b200()
    batch, cin, h, w, cout, k, s, pad, groups = case
    d = capi.ConvDesc.make(*case)
    r = rng(sum(case))
    x = f32(r.uniform(-1, 1, size=(batch, cin, h, w)))
    wt = f32(r.uniform(-1, 1, size=(cout, cin, k, k)) / np.sqrt(cin * k * k))
    dy = f32(r.uniform(-1, 1, size=(batch, cout, d.ho, d.wo)))
    math = capi.MATH_TC
    ws_bytes = lib.bcnn_b200_conv_workspace_bytes(d, math)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
    xb = lib.bcnn_b200_conv_x_shadow_bytes(d, math)
    yb = lib.bcnn_b200_conv_dy_shadow_bytes(d, math)
    xs, ys = capi.DeviceBuffer(nbytes=max(xb, 4)), capi.DeviceBuffer(nbytes=max(yb, 4))
    sh = capi.ConvShadows(xs.ptr if xb else None, xb, 0, ys.ptr if yb else None, yb, 0)
    dxv, dwt, ddy = dev(x), dev(wt), dev(dy)

    def run(use_sh):
        shp = C.byref(sh) if use_sh else None
        y, gw, gx = dev_zeros(dy.size), dev_zeros(wt.size), dev_zeros(x.size)
        if use_sh:
            check(lib.bcnn_b200_conv_forward_sh(d, dxv.ptr, dwt.ptr, None, 0, y.ptr, ws.ptr, ws_bytes,
                                                math, shp, None))
            fmt_after_fwd = sh.x_fmt
            check(lib.bcnn_b200_conv_backward_weights_sh(d, dxv.ptr, ddy.ptr, gw.ptr, ws.ptr,
                                                         ws_bytes, math, shp, None))
            fmt_after_wgrad = sh.dy_fmt
            check(lib.bcnn_b200_conv_backward_data_sh(d, dwt.ptr, ddy.ptr, gx.ptr, 0, ws.ptr,
                                                      ws_bytes, math, shp, None))
            assert (fmt_after_fwd != 0) == (xb > 0), "fprop must leave its shadow when storage exists"
            assert (fmt_after_wgrad != 0) == (yb > 0), "wgrad must leave the dy shadow for dgrad"
        else:
            check(lib.bcnn_b200_conv_forward(d, dxv.ptr, dwt.ptr, None, 0, y.ptr, ws.ptr, ws_bytes,
                                             math, None))
            check(lib.bcnn_b200_conv_backward_weights(d, dxv.ptr, ddy.ptr, gw.ptr, ws.ptr, ws_bytes,
                                                      math, None))
            check(lib.bcnn_b200_conv_backward_data(d, dwt.ptr, ddy.ptr, gx.ptr, 0, ws.ptr, ws_bytes,
                                                   math, None))
        return (y.download(np.float32, dy.shape), gw.download(np.float32, wt.shape),
                gx.download(np.float32, x.shape))

    plain, kept = run(False), run(True)
    for a, b, what in zip(plain, kept, ("fprop", "wgrad", "dgrad")):
        assert np.array_equal(a, b), f"{what}: shadow storage changed the result"
    if k == 1 and s == 1 and (h * w) % 4 == 0:
        assert xb == 0 and sh.x_fmt == 0 and sh.dy_fmt == 0
    # a stale x shadow (format NONE) must be ignored, not read
    sh.x_fmt = 0
    gw2 = dev_zeros(wt.size)
    check(lib.bcnn_b200_conv_backward_weights_sh(d, dxv.ptr, ddy.ptr, gw2.ptr, ws.ptr, ws_bytes, math,
                                                 C.byref(sh), None))
    assert np.array_equal(gw2.download(np.float32, wt.shape), plain[1])


@pytest.mark.parametrize("math", [capi.MATH_TC, capi.MATH_FP32])
@pytest.mark.parametrize("case", [
    (8, 64, 14, 14, 256, 1, 1, 0, 1),    # DIRECT 1x1, 196 positions: ragged 128-row tiles
    (8, 64, 28, 28, 320, 1, 1, 0, 1),    # two channel tiles of 160
    (16, 32, 7, 7, 48, 3, 1, 1, 1),      # several images per tile, 48 channels (ragged chunk)
    (6, 32, 30, 30, 64, 3, 2, 1, 1),     # stride 2, tiles past the right / bottom edge
    (32, 3, 32, 32, 32, 7, 2, 3, 1),     # thin first layer (im2col route)
    (4, 32, 12, 12, 32, 3, 1, 1, 2),     # groups: not on the TMA kernel -> unfused statistics
])
def test_conv_forward_bn_stats_matches_separate_kernels(case, math):
    """Batch-norm statistics fused into the convolution epilogue == statistics of the stored
    output (float64 reference over the very tensor the call wrote), output == plain fprop."""
    lib = capi.b200()
    batch, cin, h, w, cout, k, s, pad, groups = case
    d = capi.ConvDesc.make(*case)
    r = rng(sum(case) + math)
    x = f32(r.uniform(-1, 1, size=(batch, cin, h, w)) + 0.25)
    wt = f32(r.uniform(-1, 1, size=(cout, cin // groups, k, k)) / np.sqrt(cin * k * k))
    ws_bytes = lib.bcnn_b200_conv_workspace_bytes(d, math)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 4))
    dxv, dwt = dev(x), dev(wt)
    ysz = batch * cout * d.ho * d.wo
    y_plain, y_fused = dev_zeros(ysz), dev_zeros(ysz)
    check(lib.bcnn_b200_conv_forward(d, dxv.ptr, dwt.ptr, None, 0, y_plain.ptr, ws.ptr, ws_bytes,
                                     math, None))
    run0 = f32(r.uniform(-1, 1, size=cout)), f32(r.uniform(0.5, 1.5, size=cout))
    mean, var, rmean, rvar = dev_zeros(cout), dev_zeros(cout), dev(run0[0]), dev(run0[1])
    scratch = dev_zeros(lib.bcnn_b200_bn_scratch_floats(cout))
    for _ in range(2):  # twice: the tickets of the reduction must re-arm
        rmean, rvar = dev(run0[0]), dev(run0[1])
        check(lib.bcnn_b200_conv_forward_bn_stats(d, dxv.ptr, dwt.ptr, y_fused.ptr, ws.ptr, ws_bytes,
                                                  math, None, mean.ptr, var.ptr, rmean.ptr, rvar.ptr,
                                                  scratch.ptr, None))
    y = y_fused.download(np.float32, (batch, cout, d.ho, d.wo))
    assert np.array_equal(y, y_plain.download(np.float32, y.shape))
    y64 = y.astype(np.float64)
    m_ref = y64.mean(axis=(0, 2, 3))
    v_ref = (y64 ** 2).mean(axis=(0, 2, 3)) - m_ref ** 2
    got_m, got_v = mean.download(np.float32, (cout,)), var.download(np.float32, (cout,))
    assert_close(got_m, f32(m_ref), 1e-5, "fused mean")
    assert np.abs(got_v - v_ref).max() <= 2e-5 * np.abs((y64 ** 2).mean(axis=(0, 2, 3))).max()
    assert_close(rmean.download(np.float32, (cout,)), f32(0.9 * run0[0] + 0.1 * got_m), 1e-6, "run mean")
    assert_close(rvar.download(np.float32, (cout,)), f32(0.9 * run0[1] + 0.1 * got_v), 1e-6, "run var")


# ------------------------------------------------------------------ concat / upsample
@pytest.mark.parametrize("n,c1,c2,h,w", [(3, 5, 4, 10, 10), (2, 3, 7, 13, 13), (1, 256, 128, 26, 26)])
def test_concat_bit_exact(n, c1, c2, h, w):
    lib, orc = capi.b200(), oracle()
    r = rng(n + c1 + h)
    a, b = f32(r.uniform(-1, 1, size=(n, c1, h, w))), f32(r.uniform(-1, 1, size=(n, c2, h, w)))
    sa, sb, sd = c1 * h * w, c2 * h * w, (c1 + c2) * h * w
    want = np.zeros((n, c1 + c2, h, w), np.float32)
    orc.orc_concat_forward(p(a), p(want), n, sa, sd, 0)
    orc.orc_concat_forward(p(b), p(want), n, sb, sd, sa)
    da, db, dd = dev(a), dev(b), dev_zeros(want.size)
    check(lib.bcnn_b200_concat_forward(da.ptr, dd.ptr, n, sa, sd, 0, None))
    check(lib.bcnn_b200_concat_forward(db.ptr, dd.ptr, n, sb, sd, sa, None))
    assert np.array_equal(dd.download(np.float32, want.shape), want)
    assert np.array_equal(want, np.concatenate([a, b], axis=1))
    g = f32(r.uniform(-1, 1, size=want.shape))
    g0 = f32(r.uniform(-1, 1, size=b.shape))
    ref_acc, ref_new = g0.copy(), np.zeros_like(b)
    orc.orc_concat_backward(p(g), p(ref_acc), n, sb, sd, sa)
    orc.orc_concat_backward(p(g), p(ref_new), n, sb, sd, sa)
    dg, dgb = dev(g), dev(g0)
    check(lib.bcnn_b200_concat_backward(dg.ptr, dgb.ptr, n, sb, sd, sa, 1, None))
    assert np.array_equal(dgb.download(np.float32, b.shape), ref_acc)
    check(lib.bcnn_b200_concat_backward(dg.ptr, dgb.ptr, n, sb, sd, sa, 0, None))
    assert np.array_equal(dgb.download(np.float32, b.shape), ref_new)


@pytest.mark.parametrize("n,c,h,w,size", [(3, 5, 5, 5, 2), (2, 4, 7, 9, 3), (1, 128, 13, 13, 2), (2, 3, 6, 6, 1)])
def test_upsample_bit_exact(n, c, h, w, size):
    lib, orc = capi.b200(), oracle()
    r = rng(n * 7 + c + size)
    x = f32(r.uniform(-1, 1, size=(n, c, h, w)))
    want = np.zeros((n, c, h * size, w * size), np.float32)
    orc.orc_upsample_forward(p(x), p(want), n, c, h, w, size)
    dx, dy = dev(x), dev_zeros(want.size)
    check(lib.bcnn_b200_upsample_forward(dx.ptr, dy.ptr, n, c, h, w, size, None))
    assert np.array_equal(dy.download(np.float32, want.shape), want)
    g = f32(r.uniform(-1, 1, size=want.shape))
    g0 = f32(r.uniform(-1, 1, size=x.shape))
    ref_acc, ref_new = g0.copy(), np.zeros_like(x)
    orc.orc_upsample_backward(p(g), p(ref_acc), n, c, h, w, size)
    orc.orc_upsample_backward(p(g), p(ref_new), n, c, h, w, size)
    dg, dgx = dev(g), dev(g0)
    check(lib.bcnn_b200_upsample_backward(dg.ptr, dgx.ptr, n, c, h, w, size, 1, None))
    assert np.array_equal(dgx.download(np.float32, x.shape), ref_acc)   # same summation order
    check(lib.bcnn_b200_upsample_backward(dg.ptr, dgx.ptr, n, c, h, w, size, 0, None))
    assert np.array_equal(dgx.download(np.float32, x.shape), ref_new)


# ------------------------------------------------------------------ depthwise
@pytest.mark.parametrize("n,c,h,w,k,s,pad", [(2, 32, 28, 28, 3, 1, 1), (2, 16, 28, 28, 3, 2, 1),
                                             (1, 7, 9, 11, 3, 1, 0), (2, 4, 12, 12, 5, 1, 2),
                                             # 3x3 fast paths: ragged groups of four, odd planes
                                             (2, 8, 15, 13, 3, 2, 1), (1, 5, 10, 10, 3, 1, 1),
                                             (2, 4, 12, 12, 3, 1, 2), (3, 6, 14, 14, 3, 2, 1),
                                             (2, 8, 7, 7, 3, 1, 1), (2, 6, 11, 12, 3, 2, 0)])
def test_depthwise(n, c, h, w, k, s, pad):
    lib, orc = capi.b200(), oracle()
    r = rng(n + c * 13 + h)
    x = f32(r.uniform(-1, 1, size=(n, c, h, w)))
    wt = f32(r.uniform(-0.5, 0.5, size=(c, k, k)))
    bias = f32(r.uniform(-0.2, 0.2, size=c))
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    y_ref = np.zeros((n, c, ho, wo), np.float32)
    orc.orc_depthwise_forward(p(x), p(wt), p(y_ref), n, c, h, w, k, s, pad)
    orc.orc_add_bias(p(y_ref), p(bias), n, c, ho * wo)
    orc.orc_activation_forward(p(y_ref), y_ref.size, None, ho * wo, c, ACT["relu"])
    dx, dw, db, dyo = dev(x), dev(wt), dev(bias), dev_zeros(y_ref.size)
    check(lib.bcnn_b200_depthwise_forward(dx.ptr, dw.ptr, db.ptr, ACT["relu"], dyo.ptr, n, c, h, w,
                                          k, s, pad, None))
    assert_close(dyo.download(np.float32, y_ref.shape), y_ref, FP32_TOL, "depthwise fwd")
    dy = f32(r.uniform(-1, 1, size=y_ref.shape))
    gw0 = f32(r.uniform(-0.1, 0.1, size=wt.shape))
    gx0 = f32(r.uniform(-0.1, 0.1, size=x.shape))
    gw, gx = gw0.copy(), gx0.copy()
    orc.orc_depthwise_backward(p(x), p(wt), p(dy), p(gw), p(gx), n, c, h, w, k, s, pad)
    nscr = lib.bcnn_b200_depthwise_scratch_floats(n, c, k)
    scratch = dev_zeros(nscr)
    ddy, dgw, dgx = dev(dy), dev(gw0), dev(gx0)
    check(lib.bcnn_b200_depthwise_backward(dx.ptr, dw.ptr, ddy.ptr, dgw.ptr, dgx.ptr, n, c, h, w, k,
                                           s, pad, scratch.ptr, nscr, None))
    assert_close(dgw.download(np.float32, wt.shape), gw, 2e-5, "depthwise wgrad")
    assert_close(dgx.download(np.float32, x.shape), gx, FP32_TOL, "depthwise dgrad")


# ------------------------------------------------------------------ optimizer + glue
@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_sgd_update_matches_reference_sequence(n):
    lib, orc = capi.b200(), oracle()
    r = rng(n)
    w0, g0 = f32(r.uniform(-1, 1, size=n)), f32(r.uniform(-1, 1, size=n))
    w, g = w0.copy(), g0.copy()
    batch, lr, mom, decay = 64, 0.003, 0.9, 0.0005
    orc.orc_sgd_update(p(w), None, p(g), None, n, 0, batch, lr, mom, decay)
    dw, dg = dev(w0), dev(g0)
    wd = float(np.float32(decay) * np.float32(batch))      # float arithmetic, as the C host
    step = float(-np.float32(lr) / np.float32(batch))
    check(lib.bcnn_b200_sgd_update(dw.ptr, dg.ptr, n, wd, step, mom, None))
    # same mul/add sequence without contraction => bit-exact
    assert np.array_equal(dw.download().view(np.uint32), w.view(np.uint32))
    assert np.array_equal(dg.download().view(np.uint32), g.view(np.uint32))


def test_sgd_update_multi_is_the_single_tensor_kernel():
    """One launch over a table of tensors (odd sizes, a tensor shorter than a vector, one longer than a
    CTA's 4096 elements, different weight-decay factors) against bcnn_b200_sgd_update per tensor (itself
    bit-exact against the reference sequence above): bit-identical weights and gradients."""
    lib = capi.b200()
    r = rng(77)
    sizes = [1, 3, 64, 4095, 4096, 4097, 100003, 12, 9 * 64 * 64]
    wds = [0.0, 0.032, 0.032, 0.0, 0.032, 0.016, 0.032, 0.0, 0.032]
    step, g_scale = -0.003 / 64, 0.9

    class Batch(capi.C.Structure):   # include/bcnn_b200.h: bcnn_b200_sgd_batch
        _fields_ = [("w", capi.C.c_void_p * 96), ("g", capi.C.c_void_p * 96), ("n", capi.C.c_uint * 96),
                    ("wd_scale", capi.C.c_float * 96), ("first_block", capi.C.c_uint * 97),
                    ("count", capi.C.c_int), ("step", capi.C.c_float), ("g_scale", capi.C.c_float)]

    batch = Batch()
    single, multi = [], []
    for i, (n, wd) in enumerate(zip(sizes, wds)):
        w0, g0 = f32(r.uniform(-1, 1, size=n)), f32(r.uniform(-1, 1, size=n))
        a, b = (dev(w0), dev(g0)), (dev(w0), dev(g0))
        check(lib.bcnn_b200_sgd_update(a[0].ptr, a[1].ptr, n, wd, step, g_scale, None))
        single.append(a)
        multi.append(b)
        batch.w[i], batch.g[i], batch.n[i], batch.wd_scale[i] = b[0].ptr, b[1].ptr, n, wd
        batch.first_block[i + 1] = batch.first_block[i] + (n + 4095) // 4096
    batch.count, batch.step, batch.g_scale = len(sizes), step, g_scale
    lib.bcnn_b200_sgd_update_multi.argtypes = [capi.C.c_void_p, capi.C.c_void_p]
    lib.bcnn_b200_sgd_update_multi.restype = capi.C.c_int
    check(lib.bcnn_b200_sgd_update_multi(capi.C.byref(batch), None))
    for (sw, sg), (mw, mg), n in zip(single, multi, sizes):
        assert np.array_equal(mw.download().view(np.uint32), sw.download().view(np.uint32)), n
        assert np.array_equal(mg.download().view(np.uint32), sg.download().view(np.uint32)), n


@pytest.mark.parametrize("n", [1, 7, 13, 4096, 100003])
def test_adam_update_matches_reference_sequence(n):
    """Three chained Adam steps of the fused kernel against orc_adam_update (pinned bit-exact
    against the reference's bcnn_adam_update_cpu). Moments and the zeroed gradient must be
    bit-identical; the weights may differ where glibc's powf(v, 0.5f) is not the correctly
    rounded square root the kernel uses (measured: 5.5e-4 of all inputs, always 1 ulp)."""
    lib, orc = capi.b200(), oracle()
    r = rng(n + 1)
    batch, lr, b1, b2, decay = 16, 0.002, 0.9, 0.999, 0.0005
    w = f32(r.uniform(-1, 1, size=n))
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    dw, dm, dv = dev(w), dev(m), dev(v)
    for step in range(3):
        g = f32(r.normal(scale=1e-2, size=n))
        if step == 0:
            g[-3:] = 0.0   # vanishing moments in bcnn_vdiv's scalar tail: guarded quotient
            g[0] = 1e-9
        dg = dev(g)
        it = batch * (step + 1)
        orc.orc_adam_update(p(w), None, p(g), None, p(m), p(v), n, 0, batch, it, b1, b2, lr, 0.9,
                            decay)
        f = np.float32
        mu = f(np.sqrt(f(1) - f(b2) ** f(it + 1), dtype=f)) / (f(1) - f(b1) ** f(it + 1))
        alpha = float(f(-f(lr) / f(batch)) * f(mu))
        check(lib.bcnn_b200_adam_update(dw.ptr, dg.ptr, dm.ptr, dv.ptr, n,
                                        float(f(decay) * f(batch)), b1, b2, alpha, None))
        assert np.array_equal(dg.download(), np.zeros(n, np.float32))
        assert_close(dm.download(), m, 1e-6, "adam m")   # alpha / mu come from numpy's pow here
        assert_close(dv.download(), v, 1e-6, "adam v")
        got = dw.download()
        assert_close(got, w, 1e-6, f"adam w step {step}")
    assert np.array_equal(dm.download().view(np.uint32), m.view(np.uint32))
    assert np.array_equal(dv.download().view(np.uint32), v.view(np.uint32))


@pytest.mark.parametrize("n,c,hw", [(64, 10, 1), (3, 1000, 1), (2, 5, 9)])
def test_softmax(n, c, hw):
    lib, orc = capi.b200(), oracle()
    x = f32(rng(c).uniform(-5, 5, size=(n, c, hw)))
    y_ref = np.zeros_like(x)
    orc.orc_softmax_forward(p(x), p(y_ref), n, c, hw)
    dx, dyo = dev(x), dev_zeros(x.size)
    check(lib.bcnn_b200_softmax_forward(dx.ptr, dyo.ptr, n, c, hw, None))
    assert_close(dyo.download(np.float32, x.shape), y_ref, FP32_TOL, "softmax")


def test_cost_forward_error_rate_and_grad():
    lib = capi.b200()
    n, c = 16, 10
    r = rng(5)
    pred = f32(r.uniform(0, 1, size=(n, c)))
    label = np.zeros((n, c), np.float32)
    label[np.arange(n), np.arange(n) % c] = 1
    dp_, dl, dg, dm = dev(pred), dev(label), dev_zeros(n * c), dev_zeros(1)
    check(lib.bcnn_b200_cost_forward(dp_.ptr, dl.ptr, dg.ptr, dm.ptr, n, c, capi.METRIC_ERROR_RATE, None))
    assert np.array_equal(dg.download(np.float32, pred.shape), pred - label)
    wrong = float(np.sum(label[np.arange(n), pred.argmax(1)] == 0))
    assert dm.download()[0] == wrong
    check(lib.bcnn_b200_cost_forward(dp_.ptr, dl.ptr, dg.ptr, dm.ptr, n, c, capi.METRIC_SSE, None))
    assert abs(dm.download()[0] - float(((pred - label) ** 2).sum())) < 1e-3


def test_eltwise_forward_backward():
    lib = capi.b200()
    r = rng(17)
    a, b = f32(r.uniform(-1, 1, size=4099)), f32(r.uniform(-1, 1, size=4099))
    da, db, dyo = dev(a), dev(b), dev_zeros(a.size)
    check(lib.bcnn_b200_eltwise_forward(da.ptr, db.ptr, dyo.ptr, a.size, a.size, ACT["relu"], None))
    y = dyo.download()
    assert np.array_equal(y, np.maximum(a + b, 0) * 1.0)
    g = f32(r.uniform(-1, 1, size=a.size))
    ga0, gb0 = f32(r.uniform(-1, 1, size=a.size)), f32(r.uniform(-1, 1, size=a.size))
    dg, dga, dgb = dev(g), dev(ga0), dev(gb0)
    check(lib.bcnn_b200_eltwise_backward(dyo.ptr, dg.ptr, dga.ptr, dgb.ptr, a.size, a.size, ACT["relu"], 3, None))
    gm = g * (y > 0)
    assert np.array_equal(dga.download(), ga0 + gm) and np.array_equal(dgb.download(), gb0 + gm)
    # stale-buffer flavour: da overwritten, db accumulated; then both overwritten with the
    # reference quirk (only the first n_add elements of db receive a gradient)
    dg.upload(g)
    check(lib.bcnn_b200_eltwise_backward(dyo.ptr, dg.ptr, dga.ptr, dgb.ptr, a.size, a.size, ACT["relu"], 2, None))
    assert np.array_equal(dga.download(), gm) and np.array_equal(dgb.download(), gb0 + gm + gm)
    dg.upload(g)
    check(lib.bcnn_b200_eltwise_backward(dyo.ptr, dg.ptr, dga.ptr, dgb.ptr, a.size, 1024, ACT["relu"], 0, None))
    want_b = np.where(np.arange(a.size) < 1024, gm, 0).astype(np.float32)
    assert np.array_equal(dga.download(), gm) and np.array_equal(dgb.download(), want_b)
