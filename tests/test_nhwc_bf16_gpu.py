"""Kernel-level parity of the BF16 NHWC resident kernels (csrc/nhwc_bf16.cu) against the oracle
restatement evaluated on the SAME BF16-rounded inputs (transposed to the oracle's NCHW FP32).
Bars: integer work (max-pool values and argmax, the converters) bit-exact; FP32 outputs
(parameter gradients, pooled means) 1e-4; BF16 outputs 1e-2 = 2.5 ulp of the output format,
inside the 2e-2 tensor-core tolerance of BASELINE.json."""
import numpy as np
import pytest

from bcnn_b200 import capi
from bcnn_b200.capi import ACT
from helpers import assert_close, check, dev, dev_zeros, f32, oracle, p

pytestmark = pytest.mark.gpu
BF16_OUT_TOL = 1e-2


def to_bf16_bits(a):
    """float32 -> bfloat16 bits, round to nearest even (what cvt.rn.bf16x2.f32 does)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def from_bf16_bits(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def rounded(a):
    return from_bf16_bits(to_bf16_bits(a)).reshape(np.shape(a))


def nhwc_bits(x_nchw):
    """NCHW float32 (already BF16-representable) -> NHWC bf16 bit array."""
    return to_bf16_bits(np.ascontiguousarray(x_nchw.transpose(0, 2, 3, 1)))


def nchw_from_bits(bits, shape):
    n, c, h, w = shape
    return np.ascontiguousarray(from_bf16_bits(bits).reshape(n, h, w, c).transpose(0, 3, 1, 2))


def rng(seed):
    return np.random.default_rng(seed)


SHAPES = [(2, 8, 5, 7), (3, 64, 14, 14), (2, 24, 9, 9), (4, 256, 7, 7), (2, 2048, 3, 3)]


@pytest.mark.parametrize("shape", SHAPES + [(2, 6, 5, 5), (1, 130, 33, 3)])
def test_converters_round_trip_bit_exact(shape):
    lib = capi.b200()
    n, c, h, w = shape
    x = f32(rng(1).normal(size=shape))
    dx, db, dback = dev(x), dev_zeros(x.size, 2), dev_zeros(x.size)
    check(lib.bcnn_b200_f32nchw_to_bf16nhwc(dx.ptr, db.ptr, n, c, h * w, None))
    bits = db.download(np.uint16)
    assert np.array_equal(bits, nhwc_bits(x).ravel())
    check(lib.bcnn_b200_bf16nhwc_to_f32nchw(db.ptr, dback.ptr, n, c, h * w, None))
    assert np.array_equal(dback.download(np.float32, shape), rounded(x))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("act", ["none", "relu", "lrelu"])
def test_bn_apply_and_backward_nhwc(shape, act):
    lib, orc = capi.b200(), oracle()
    n, c, h, w = shape
    hw = h * w
    r = rng(sum(shape))
    x = rounded(f32(r.normal(0.2, 1.0, size=shape)))
    gamma, beta = f32(r.uniform(0.5, 1.5, size=c)), f32(r.uniform(-0.3, 0.3, size=c))
    # reference forward (TRAIN statistics of the rounded x)
    y = x.copy()
    rm, rv, sm, sv = (np.zeros(c, np.float32) for _ in range(4))
    xn, xc = np.zeros_like(x), np.zeros_like(x)
    orc.orc_bn_forward(p(y), n, c, hw, p(rm), p(rv), p(gamma), p(beta), p(sm), p(sv), p(xn), p(xc), 1)
    pre = y.copy()
    orc.orc_activation_forward(p(y), y.size, None, hw, c, ACT[act])
    dxb, dyb = dev(nhwc_bits(x)), dev_zeros(x.size, 2)
    dsm, dsv, dg, dbt = dev(sm), dev(sv), dev(gamma), dev(beta)
    check(lib.bcnn_b200_bn_apply_nhwc(dxb.ptr, dyb.ptr, dsm.ptr, dsv.ptr, dg.ptr, dbt.ptr, n * hw, c,
                                      ACT[act], None))
    y_gpu = nchw_from_bits(dyb.download(np.uint16), shape)
    assert_close(y_gpu, y, BF16_OUT_TOL, "bn apply")
    # PREDICT flavour: y = act(gamma x + beta)
    check(lib.bcnn_b200_bn_apply_nhwc(dxb.ptr, dyb.ptr, None, None, dg.ptr, dbt.ptr, n * hw, c,
                                      ACT[act], None))
    yp = x * gamma[None, :, None, None] + beta[None, :, None, None]
    orc.orc_activation_forward(p(yp), yp.size, None, hw, c, ACT[act])
    assert_close(nchw_from_bits(dyb.download(np.uint16), shape), yp, BF16_OUT_TOL, "scale / bias")
    # backward; keep pre-activations away from zero so the (bit-consistent) mask matches the oracle's
    g = rounded(f32(r.uniform(-1, 1, size=shape)))
    g[np.abs(pre) < 1e-3] = 0.0
    g_ref = g.copy()
    orc.orc_activation_backward(p(y), p(g_ref), g.size, None, None, hw, c, ACT[act])
    gg0, gb0 = f32(r.uniform(-1, 1, size=c)), f32(r.uniform(-1, 1, size=c))
    gg, gb = gg0.copy(), gb0.copy()
    dm, dv = np.zeros(c, np.float32), np.zeros(c, np.float32)
    orc.orc_bn_backward(p(g_ref), n, c, hw, p(gamma), p(gg), p(gb), p(sm), p(sv), p(dm), p(dv), p(xn), p(xc))
    dgr = dev(nhwc_bits(g))
    dgg, dgb, ddm, ddv = dev(gg0), dev(gb0), dev_zeros(c), dev_zeros(c)
    scratch = dev_zeros(lib.bcnn_b200_nhwc_scratch_floats(c))
    check(lib.bcnn_b200_bn_backward_nhwc(dxb.ptr, dgr.ptr, dgr.ptr, dsm.ptr, dsv.ptr, dg.ptr, dbt.ptr,
                                         dgg.ptr, dgb.ptr, ddm.ptr, ddv.ptr, n * hw, c, ACT[act],
                                         scratch.ptr, None))
    assert_close(dgb.download(), gb, 1e-4, "g_beta")
    assert_close(dgg.download(), gg, 1e-4, "g_gamma")
    assert_close(nchw_from_bits(dgr.download(np.uint16), shape), g_ref, BF16_OUT_TOL, "bn dx")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("act", ["none", "relu", "lrelu"])
def test_actbwd_grad_bias_nhwc(shape, act):
    lib, orc = capi.b200(), oracle()
    n, c, h, w = shape
    r = rng(7 + sum(shape))
    y = rounded(f32(r.normal(size=shape)))
    orc.orc_activation_forward(p(y), y.size, None, h * w, c, ACT[act])
    y = rounded(y)
    g = rounded(f32(r.uniform(-1, 1, size=shape)))
    g_ref = g.copy()
    orc.orc_activation_backward(p(y), p(g_ref), g.size, None, None, h * w, c, ACT[act])
    gb0 = f32(r.uniform(-1, 1, size=c))
    gb = gb0.copy()
    orc.orc_grad_bias(p(gb), p(g_ref), n, c, h * w)
    dy, dg, dgb = dev(nhwc_bits(y)), dev(nhwc_bits(g)), dev(gb0)
    scratch = dev_zeros(lib.bcnn_b200_nhwc_scratch_floats(c))
    check(lib.bcnn_b200_actbwd_grad_bias_nhwc(dgb.ptr, dg.ptr, dy.ptr, ACT[act], n * h * w, c, scratch.ptr, None))
    assert_close(dgb.download(), gb, 1e-4, "g_bias")
    assert_close(nchw_from_bits(dg.download(np.uint16), shape), g_ref, BF16_OUT_TOL, "dy * act'")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("quirk", [False, True])
def test_eltwise_bf16(shape, quirk):
    lib = capi.b200()
    n, c, h, w = shape
    r = rng(3 + sum(shape))
    a, b = rounded(f32(r.normal(size=shape))), rounded(f32(r.normal(size=shape)))
    sz = a.size
    n_add = sz // n if quirk else sz      # the reference adds only sample 0 (H3)
    bb = b.copy()
    if quirk:
        bb[1:] = 0
    y_ref = np.maximum(a + bb, 0)
    da_, db_, dy_ = dev(nhwc_bits(a)), dev(nhwc_bits(b)), dev_zeros(sz, 2)
    check(lib.bcnn_b200_eltwise_forward_bf16(da_.ptr, db_.ptr, dy_.ptr, sz, n_add, ACT["relu"], None))
    y_gpu = nchw_from_bits(dy_.download(np.uint16), shape)
    assert np.array_equal(y_gpu, rounded(y_ref))
    g = rounded(f32(r.uniform(-1, 1, size=shape)))
    gm = g * (y_gpu > 0)
    old = rounded(f32(r.uniform(-1, 1, size=shape)))
    dg, dga, dgb = dev(nhwc_bits(g)), dev(nhwc_bits(old)), dev(nhwc_bits(old))
    # da accumulates onto `old`, db overwrites
    check(lib.bcnn_b200_eltwise_backward_bf16(dy_.ptr, dg.ptr, dga.ptr, dgb.ptr, sz, n_add, ACT["relu"], 1, None))
    assert np.array_equal(nchw_from_bits(dg.download(np.uint16), shape), gm)
    assert np.array_equal(nchw_from_bits(dga.download(np.uint16), shape), rounded(old + gm))
    want_b = gm.copy()
    if quirk:
        want_b[1:] = 0
    assert np.array_equal(nchw_from_bits(dgb.download(np.uint16), shape), want_b)


POOLS = [(2, 8, 28, 28, 2, 2), (2, 64, 112 // 4, 112 // 4, 3, 2), (1, 16, 13, 13, 2, 1), (3, 24, 9, 11, 3, 2),
         (2, 8, 7, 7, 3, 3)]


@pytest.mark.parametrize("case", POOLS)
@pytest.mark.parametrize("ties", ["none", "relu"])
def test_maxpool_nhwc_is_bit_exact_on_the_same_bf16_values(case, ties):
    lib, orc = capi.b200(), oracle()
    n, c, h, w, k, s = case
    r = rng(11 + sum(case))
    x = rounded(f32(r.normal(size=(n, c, h, w))))
    if ties == "relu":
        x = np.maximum(x, 0)    # many exact-zero ties: the first one must win
    ho, wo = orc.orc_maxpool_out_dim(h, k, s, capi.PAD_SAME), orc.orc_maxpool_out_dim(w, k, s, capi.PAD_SAME)
    y_ref = np.zeros((n, c, ho, wo), np.float32)
    i_ref = np.zeros((n, c, ho, wo), np.int32)
    orc.orc_maxpool_forward(p(x), p(y_ref), p(i_ref), n, c, h, w, k, s, ho, wo)
    dxb, dyb, dib = dev(nhwc_bits(x)), dev_zeros(y_ref.size, 2), dev_zeros(i_ref.size)
    check(lib.bcnn_b200_maxpool_forward_nhwc(dxb.ptr, dyb.ptr, dib.ptr, n, c, h, w, k, s, ho, wo, None))
    assert np.array_equal(nchw_from_bits(dyb.download(np.uint16), y_ref.shape), y_ref)
    idx = dib.download(np.int32).reshape(n, ho, wo, c).transpose(0, 3, 1, 2)
    assert np.array_equal(idx, i_ref), "argmax mismatch"
    g = rounded(f32(r.uniform(-1, 1, size=y_ref.shape)))
    dx_ref = np.zeros_like(x)
    orc.orc_maxpool_backward(p(dx_ref), p(g), p(i_ref), g.size)
    dgb, ddx = dev(nhwc_bits(g)), dev(nhwc_bits(np.full_like(x, 3.0)))
    check(lib.bcnn_b200_maxpool_backward_nhwc(ddx.ptr, dgb.ptr, dib.ptr, n, c, h, w, k, s, ho, wo, 0, None))
    assert_close(nchw_from_bits(ddx.download(np.uint16), x.shape), dx_ref, BF16_OUT_TOL, "maxpool dx")
    check(lib.bcnn_b200_maxpool_backward_nhwc(ddx.ptr, dgb.ptr, dib.ptr, n, c, h, w, k, s, ho, wo, 1, None))
    assert_close(nchw_from_bits(ddx.download(np.uint16), x.shape), 2 * dx_ref, BF16_OUT_TOL, "maxpool dx +=")


@pytest.mark.parametrize("shape", [(2, 2048, 7, 7), (3, 1024, 7, 7), (2, 8, 5, 3), (1, 264, 4, 4)])
def test_avgpool_nhwc(shape):
    lib, orc = capi.b200(), oracle()
    n, c, h, w = shape
    r = rng(5 + sum(shape))
    x = rounded(f32(r.normal(size=shape)))
    y_ref = np.zeros((n, c), np.float32)
    orc.orc_avgpool_forward(p(x), p(y_ref), n, c, h * w)
    dxb, dyv = dev(nhwc_bits(x)), dev_zeros(n * c)
    check(lib.bcnn_b200_avgpool_forward_nhwc(dxb.ptr, dyv.ptr, n, c, h * w, None))
    assert_close(dyv.download(np.float32, (n, c)), y_ref, 1e-5, "avgpool")
    g = f32(r.uniform(-1, 1, size=(n, c)))
    dx_ref = np.zeros_like(x)
    orc.orc_avgpool_backward(p(dx_ref), p(g), n, c, h * w)
    dg, ddx = dev(g), dev(nhwc_bits(np.full_like(x, 5.0)))
    check(lib.bcnn_b200_avgpool_backward_nhwc(ddx.ptr, dg.ptr, n, c, h * w, 0, None))
    assert_close(nchw_from_bits(ddx.download(np.uint16), shape), dx_ref, BF16_OUT_TOL, "avgpool dx")


# ------------------------------------------------------------------ convolution on resident tensors
# (cin, h, cout, k, stride, pad): every ResNet-50 class plus YOLO / odd shapes
RESIDENT_CONVS = [
    (64, 56, 64, 1, 1, 0), (64, 56, 64, 3, 1, 1), (64, 56, 256, 1, 1, 0), (256, 56, 64, 1, 1, 0),
    (128, 56, 128, 3, 2, 1), (256, 56, 512, 1, 2, 0), (256, 28, 256, 3, 2, 1), (512, 28, 128, 1, 1, 0),
    (256, 14, 256, 3, 1, 1), (1024, 14, 256, 1, 1, 0), (512, 14, 512, 3, 2, 1), (512, 7, 2048, 1, 1, 0),
    (2048, 7, 512, 1, 1, 0), (512, 7, 512, 3, 1, 1), (1024, 14, 2048, 1, 2, 0),
    (16, 26, 32, 3, 1, 1), (32, 13, 72, 3, 1, 1), (24, 9, 40, 1, 1, 0), (128, 13, 256, 3, 1, 1),
    (64, 20, 384, 3, 1, 1),
]


def _conv_case(shape, batch, seed_extra=0):
    cin, h, cout, k, s, pad = shape
    orc = oracle()
    r = rng(hash(shape) % 2**31 + seed_extra)
    d = capi.ConvDesc.make(batch, cin, h, h, cout, k, s, pad, 1)
    x = rounded(f32(r.uniform(-1, 1, size=(batch, cin, h, h))))
    wt = f32(r.uniform(-1, 1, size=(cout, cin, k, k)) * np.sqrt(3.0 / (cin * k * k)))
    y = np.zeros((batch, cout, d.ho, d.wo), np.float32)
    orc.orc_conv_forward(p(x), p(wt), p(y), batch, cin, h, h, cout, k, s, pad, 1)
    dy = rounded(f32(r.uniform(-1, 1, size=y.shape)))
    gw0 = f32(r.uniform(-0.1, 0.1, size=wt.shape))
    gw = gw0.copy()
    dx = np.zeros_like(x)
    orc.orc_conv_backward(p(x), p(wt), p(dy), p(gw), p(dx), batch, cin, h, h, cout, k, s, pad, 1)
    return d, x, wt, y, dy, gw0, gw, dx


@pytest.mark.parametrize("shape", RESIDENT_CONVS, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("batch", [2, 5])
def test_resident_conv_three_passes_vs_oracle(shape, batch):
    """fprop (+bias +ReLU), dgrad (overwrite and accumulate), wgrad (+=) on BF16 NHWC tensors against
    the oracle on the same BF16-rounded inputs: 2e-2 tensor-core class (weights are rounded to BF16 by
    the packing kernel, results to BF16 by the epilogue)."""
    lib = capi.b200()
    d, x, wt, y_ref, dy, gw0, gw_ref, dx_ref = _conv_case(shape, batch)
    mask = lib.bcnn_b200_conv_nhwc_supported(d)
    assert mask & 1 and mask & 2, f"{shape}: fprop / dgrad not on the resident kernels ({mask})"
    ws_bytes = lib.bcnn_b200_conv_nhwc_workspace_bytes(d)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 256))
    dxb, dwt = dev(nhwc_bits(x)), dev(wt)
    dyo = dev(np.full(y_ref.size, 0x7fc0, np.uint16))           # NaN-filled: every element must be written
    check(lib.bcnn_b200_conv_forward_nhwc(d, dxb.ptr, dwt.ptr, None, 0, dyo.ptr, ws.ptr, ws_bytes, None, None))
    got = nchw_from_bits(dyo.download(np.uint16), y_ref.shape)
    assert_close(got, y_ref, 2e-2, "fprop")
    # bias + ReLU in the epilogue
    bias = f32(rng(3).uniform(-0.5, 0.5, size=d.cout))
    dbias = dev(bias)
    check(lib.bcnn_b200_conv_forward_nhwc(d, dxb.ptr, dwt.ptr, dbias.ptr, ACT["relu"], dyo.ptr, ws.ptr,
                                          ws_bytes, None, None))
    got = nchw_from_bits(dyo.download(np.uint16), y_ref.shape)
    assert_close(got, np.maximum(y_ref + bias.reshape(1, -1, 1, 1), 0), 2e-2, "fprop + bias + relu")
    # dgrad
    ddy = dev(nhwc_bits(dy))
    ddx = dev(np.full(x.size, 0x7fc0, np.uint16))
    check(lib.bcnn_b200_conv_backward_data_nhwc(d, dwt.ptr, ddy.ptr, ddx.ptr, 0, ws.ptr, ws_bytes, None))
    got = nchw_from_bits(ddx.download(np.uint16), x.shape)
    assert_close(got, dx_ref, 2e-2, "dgrad (overwrite)")
    base = rounded(f32(rng(4).uniform(-1, 1, size=x.shape)))
    ddx = dev(nhwc_bits(base))
    check(lib.bcnn_b200_conv_backward_data_nhwc(d, dwt.ptr, ddy.ptr, ddx.ptr, 1, ws.ptr, ws_bytes, None))
    got = nchw_from_bits(ddx.download(np.uint16), x.shape)
    assert_close(got, dx_ref + base, 2e-2, "dgrad (accumulate)")
    # wgrad
    if mask & 4:
        dgw = dev(gw0)
        check(lib.bcnn_b200_conv_backward_weights_nhwc(d, dxb.ptr, ddy.ptr, dgw.ptr, ws.ptr, ws_bytes, None, None))
        assert_close(dgw.download(np.float32, wt.shape), gw_ref, 2e-2, "wgrad (+=)")
    else:
        assert batch * d.ho * d.wo < 512 or d.cout < 32 or d.cin < 16, f"{shape}: wgrad not covered"


@pytest.mark.parametrize("shape", [(64, 56, 64, 3, 1, 1), (256, 14, 1024, 1, 1, 0), (512, 7, 512, 3, 1, 1),
                                   (128, 28, 512, 1, 1, 0)], ids=lambda s: "x".join(map(str, s)))
def test_resident_conv_fused_bn_statistics(shape):
    """Mean / biased variance / running statistics from the epilogue's FP32 accumulators against the
    oracle's batch-norm forward over the oracle's convolution result."""
    lib, orc = capi.b200(), oracle()
    batch = 4
    d, x, wt, y_ref, *_ = _conv_case(shape, batch, 7)
    c, hw = d.cout, d.ho * d.wo
    rm, rv, sm, sv = (np.zeros(c, np.float32) for _ in range(4))
    gamma, beta = np.ones(c, np.float32), np.zeros(c, np.float32)
    yy = y_ref.copy()
    xn, xc = np.zeros_like(yy), np.zeros_like(yy)
    orc.orc_bn_forward(p(yy), batch, c, hw, p(rm), p(rv), p(gamma), p(beta), p(sm), p(sv), p(xn), p(xc), 1)
    ws_bytes = lib.bcnn_b200_conv_nhwc_workspace_bytes(d)
    ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 256))
    dxb, dwt, dyo = dev(nhwc_bits(x)), dev(wt), dev_zeros(y_ref.size, 2)
    dsm, dsv, drm, drv = (dev(np.zeros(c, np.float32)) for _ in range(4))
    sc1 = capi.DeviceBuffer(nbytes=4 * lib.bcnn_b200_nhwc_scratch_floats(c))
    sc2 = capi.DeviceBuffer(nbytes=4 * lib.bcnn_b200_bn_scratch_floats(c))
    check(lib.bcnn_b200_conv_forward_bn_stats_nhwc(d, dxb.ptr, dwt.ptr, dyo.ptr, ws.ptr, ws_bytes, None,
                                                   dsm.ptr, dsv.ptr, drm.ptr, drv.ptr, sc1.ptr, sc2.ptr, None))
    assert_close(nchw_from_bits(dyo.download(np.uint16), y_ref.shape), y_ref, 2e-2, "raw result")
    assert_close(dsm.download(), sm, 2e-2, "saved mean")
    assert_close(dsv.download(), sv, 2e-2, "saved variance")
    assert_close(drm.download(), rm, 2e-2, "running mean")
    assert_close(drv.download(), rv, 2e-2, "running variance")
    # the stand-alone statistics kernel over the BF16 result agrees with the fused ones
    dsm2, dsv2, drm2, drv2 = (dev(np.zeros(c, np.float32)) for _ in range(4))
    check(lib.bcnn_b200_bn_stats_nhwc(dyo.ptr, batch * hw, c, dsm2.ptr, dsv2.ptr, drm2.ptr, drv2.ptr,
                                      sc1.ptr, sc2.ptr, None))
    assert_close(dsm2.download(), sm, 2e-2, "stand-alone mean")
    assert_close(dsv2.download(), sv, 2e-2, "stand-alone variance")


def test_resident_thin_first_layer_reads_fp32_nchw():
    """ResNet-50's stem (3 -> 64, 7x7 / 2) and YOLOv3-tiny's first layer: FP32 NCHW input gathered into a
    BF16 im2col buffer, BF16 NHWC result; wgrad from the kept buffer."""
    lib = capi.b200()
    for shape, batch in (((3, 224, 64, 7, 2, 3), 2), ((3, 64, 16, 3, 1, 1), 2)):
        d, x, wt, y_ref, dy, gw0, gw_ref, _ = _conv_case(shape, batch, 11)
        mask = lib.bcnn_b200_conv_nhwc_supported(d)
        assert mask & 1, shape
        ws_bytes = lib.bcnn_b200_conv_nhwc_workspace_bytes(d)
        ws = capi.DeviceBuffer(nbytes=max(ws_bytes, 256))
        keep = capi.DeviceBuffer(nbytes=lib.bcnn_b200_conv_nhwc_x_keep_bytes(d))
        sh = capi.ConvShadows()
        sh.x, sh.x_bytes = keep.ptr, keep.nbytes
        dxf, dwt, dyo = dev(x), dev(wt), dev_zeros(y_ref.size, 2)
        check(lib.bcnn_b200_conv_forward_nhwc(d, dxf.ptr, dwt.ptr, None, 0, dyo.ptr, ws.ptr, ws_bytes,
                                              capi.C.byref(sh), None))
        assert_close(nchw_from_bits(dyo.download(np.uint16), y_ref.shape), y_ref, 2e-2, f"fprop {shape}")
        if mask & 4:
            ddy, dgw = dev(nhwc_bits(dy)), dev(gw0)
            check(lib.bcnn_b200_conv_backward_weights_nhwc(d, dxf.ptr, ddy.ptr, dgw.ptr, ws.ptr, ws_bytes,
                                                           capi.C.byref(sh), None))
            assert_close(dgw.download(np.float32, wt.shape), gw_ref, 2e-2, f"wgrad {shape}")
            # and without the kept buffer (rebuilt in the workspace)
            dgw = dev(gw0)
            check(lib.bcnn_b200_conv_backward_weights_nhwc(d, dxf.ptr, ddy.ptr, dgw.ptr, ws.ptr, ws_bytes,
                                                           None, None))
            assert_close(dgw.download(np.float32, wt.shape), gw_ref, 2e-2, f"wgrad rebuilt {shape}")


@pytest.mark.parametrize("shape", [(3, 64, 14, 14), (2, 256, 7, 7), (4, 8, 5, 3)])
@pytest.mark.parametrize("kinds", ["bn+plain", "bn+bn", "plain+bn", "fold+plain"])
def test_fused_bn_residual_add_matches_the_unfused_kernels(shape, kinds):
    """bn_add_act (normalise one or both operands, add, ReLU in one pass) against bn_apply_nhwc on each
    operand followed by eltwise_forward_bf16: within the rounding the unfused path adds by storing
    each normalised branch as BF16 (1e-2 = 2.5 ulp of the output format)."""
    lib = capi.b200()
    n, c, h, w = shape
    pos = n * h * w
    r = rng(len(kinds) + sum(shape))
    ka, kb = kinds.split("+")
    ops = []
    for kind in (ka, kb):
        x = dev(nhwc_bits(rounded(f32(r.normal(0.1, 1.0, size=shape)))))
        prm = dict(mean=dev(f32(r.normal(0, 0.3, size=c))), var=dev(f32(r.uniform(0.5, 2.0, size=c))),
                   gamma=dev(f32(r.uniform(0.5, 1.5, size=c))), beta=dev(f32(r.uniform(-0.3, 0.3, size=c))))
        ops.append((kind, x, prm))
    args, branches = [], []
    for kind, x, prm in ops:
        if kind == "plain":
            args += [x.ptr, None, None, None, None]
            branches.append(x)
        else:
            mean, var = (prm["mean"].ptr, prm["var"].ptr) if kind == "bn" else (None, None)
            args += [x.ptr, mean, var, prm["gamma"].ptr, prm["beta"].ptr]
            y = dev_zeros(x.nbytes // 2, 2)
            check(lib.bcnn_b200_bn_apply_nhwc(x.ptr, y.ptr, mean, var, prm["gamma"].ptr, prm["beta"].ptr, pos, c,
                                              ACT["none"], None))
            branches.append(y)
    want = dev_zeros(pos * c, 2)
    check(lib.bcnn_b200_eltwise_forward_bf16(branches[0].ptr, branches[1].ptr, want.ptr, pos * c, pos * c,
                                             ACT["relu"], None))
    got = dev_zeros(pos * c, 2)
    check(lib.bcnn_b200_bn_add_act_nhwc(*args, got.ptr, pos, c, ACT["relu"], None))
    assert_close(nchw_from_bits(got.download(np.uint16), shape), nchw_from_bits(want.download(np.uint16), shape),
                 BF16_OUT_TOL, "fused vs unfused")


@pytest.mark.parametrize("shape", [(3, 64, 14, 14), (2, 256, 7, 7), (4, 8, 5, 3), (2, 2048, 3, 3)])
@pytest.mark.parametrize("branches", ["a", "b", "ab"])
@pytest.mark.parametrize("act", ["relu", "none"])
def test_fused_residual_add_backward_bn_reduce_matches_the_unfused_kernels(shape, branches, act):
    """eltwise_backward_bn_reduce (mask the add's gradient, copy / accumulate it for a plain branch and
    reduce S1, S2 for the batch-normed branches in one pass) + bn_backward_nhwc_partials against
    eltwise_backward_bf16 + bn_backward_nhwc (itself held to the oracle above): the masked gradient and
    the plain branch bit-identical, the parameter gradients 1e-4, dx within the BF16 output bar."""
    lib = capi.b200()
    n, c, h, w = shape
    pos = n * h * w
    r = rng(len(branches) + sum(shape))
    y = dev(nhwc_bits(rounded(f32(np.maximum(r.normal(0.1, 1.0, size=shape), 0.0)))))
    dy0 = nhwc_bits(rounded(f32(r.uniform(-1, 1, size=shape))))
    xs = [dev(nhwc_bits(rounded(f32(r.normal(0.2, 1.0, size=shape))))) for _ in range(2)]
    prm = [dict(mean=dev(f32(r.normal(0.2, 0.1, size=c))), var=dev(f32(r.uniform(0.5, 2.0, size=c))),
                gamma=dev(f32(r.uniform(0.5, 1.5, size=c))), beta=dev(f32(r.uniform(-0.3, 0.3, size=c))))
           for _ in range(2)]
    plain0 = nhwc_bits(rounded(f32(r.uniform(-1, 1, size=shape))))   # existing gradient of a plain branch
    fused = [("a" in branches), ("b" in branches)]

    def bn_bwd(i, dy_buf, partial=None, rows=0):
        gg, gb, dm, dv = dev(np.ones(c, np.float32)), dev(np.ones(c, np.float32)), dev_zeros(c), dev_zeros(c)
        dx = dev_zeros(pos * c, 2)
        q = prm[i]
        if partial is None:
            scratch = dev_zeros(lib.bcnn_b200_nhwc_scratch_floats(c))
            check(lib.bcnn_b200_bn_backward_nhwc(xs[i].ptr, dy_buf.ptr, dx.ptr, q["mean"].ptr, q["var"].ptr,
                                                 q["gamma"].ptr, q["beta"].ptr, gg.ptr, gb.ptr, dm.ptr, dv.ptr,
                                                 pos, c, ACT["none"], scratch.ptr, None))
        else:
            check(lib.bcnn_b200_bn_backward_nhwc_partials(xs[i].ptr, dy_buf.ptr, dx.ptr, q["mean"].ptr,
                                                          q["var"].ptr, q["gamma"].ptr, q["beta"].ptr, gg.ptr,
                                                          gb.ptr, dm.ptr, dv.ptr, pos, c, partial.ptr, rows, None))
        return gg.download(), gb.download(), dx.download(np.uint16)

    # unfused: the add masks dy in place and accumulates into the plain branch (flag bit of that operand)
    dy_u, plain_u = dev(dy0), dev(plain0)
    da = None if fused[0] else plain_u.ptr
    db = None if fused[1] else plain_u.ptr
    flags = (0 if fused[0] else 1) | (0 if fused[1] else 2)
    check(lib.bcnn_b200_eltwise_backward_bf16(y.ptr, dy_u.ptr, da, db, pos * c, pos * c, ACT[act], flags, None))
    want = [bn_bwd(i, dy_u) if fused[i] else None for i in range(2)]
    # fused
    dy_f, plain_f = dev(dy0), dev(plain0)
    da = None if fused[0] else plain_f.ptr
    db = None if fused[1] else plain_f.ptr
    partials = [dev_zeros(lib.bcnn_b200_nhwc_scratch_floats(c)) for _ in range(2)]
    rows = capi.C.c_int(0)
    check(lib.bcnn_b200_eltwise_backward_bn_reduce_bf16(
        y.ptr, dy_f.ptr, da, db, pos, c, ACT[act], flags,
        xs[0].ptr if fused[0] else None, prm[0]["mean"].ptr if fused[0] else None,
        partials[0].ptr if fused[0] else None, xs[1].ptr if fused[1] else None,
        prm[1]["mean"].ptr if fused[1] else None, partials[1].ptr if fused[1] else None, capi.C.byref(rows), None))
    assert rows.value > 0
    assert np.array_equal(dy_f.download(np.uint16), dy_u.download(np.uint16)), "masked gradient"
    assert np.array_equal(plain_f.download(np.uint16), plain_u.download(np.uint16)), "plain branch"
    for i in range(2):
        if not fused[i]:
            continue
        gg, gb, dx = bn_bwd(i, dy_f, partials[i], rows.value)
        assert_close(gb, want[i][1], 1e-4, f"g_beta {i}")
        assert_close(gg, want[i][0], 1e-4, f"g_gamma {i}")
        assert_close(nchw_from_bits(dx, shape), nchw_from_bits(want[i][2], shape), BF16_OUT_TOL, f"bn dx {i}")
