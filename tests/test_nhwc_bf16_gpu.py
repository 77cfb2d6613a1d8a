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
