"""Pins the oracle restatement (oracle/bcnn_oracle.c) against the UNMODIFIED reference CPU
library compiled from /root/reference (oracle/_ref/libbcnn_ref.so): same inputs through the
reference's own functions / public API vs the restatement. Skipped when _ref is absent."""
import ctypes as C

import numpy as np
import pytest

from bcnn_b200 import capi, configs
from helpers import assert_close, f32, oracle, p, ref_available, ref_lib, ref_net

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
ACT = capi.ACT


def rng(seed):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("n", [1, 7, 8, 49, 196, 784, 1000])
def test_sse_lane_reductions_are_bit_exact(n):
    """bcnn_vsum / bcnn_dot / bcnn_shiftdot (bcnn_mat.c:413-475, 652-690)."""
    lib, orc = ref_lib(), oracle()
    r = rng(n)
    x, y = f32(r.normal(size=n)), f32(r.normal(size=n))
    lib.bcnn_dot.restype = C.c_float
    lib.bcnn_dot.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    lib.bcnn_shiftdot.restype = C.c_float
    lib.bcnn_shiftdot.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_float]
    lib.bcnn_vsum.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    s = C.c_float()
    lib.bcnn_vsum(n, p(x), C.byref(s))
    assert np.float32(s.value) == np.float32(orc.orc_vsum(n, p(x)))
    assert np.float32(lib.bcnn_dot(n, p(x), p(y))) == np.float32(orc.orc_dot(n, p(x), p(y)))
    assert np.float32(lib.bcnn_shiftdot(n, p(x), 0.25, p(y), 0.0)) == \
        np.float32(orc.orc_shiftdot(n, p(x), 0.25, p(y), 0.0))


@pytest.mark.parametrize("act", list(ACT)[1:])
def test_activations(act):
    lib, orc = ref_lib(), oracle()
    lib.bcnn_forward_activation_cpu.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.bcnn_backward_activation_cpu.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_int, C.c_int, C.c_int]
    r = rng(ACT[act])
    c, hw = 6, 35
    x = f32(r.uniform(-2, 2, size=(3, c, hw)))
    x.ravel()[::5] = 0
    slope = f32(r.uniform(0.05, 0.3, size=c))
    a, b = x.copy(), x.copy()
    lib.bcnn_forward_activation_cpu(p(a), a.size, p(slope), hw, c, ACT[act])
    orc.orc_activation_forward(p(b), b.size, p(slope), hw, c, ACT[act])
    assert np.array_equal(a, b)
    g = f32(r.uniform(-1, 1, size=x.shape))
    ga, gb = g.copy(), g.copy()
    gsa, gsb = np.zeros(c, np.float32), np.zeros(c, np.float32)
    lib.bcnn_backward_activation_cpu(p(a), p(ga), a.size, p(slope), p(gsa), hw, c, ACT[act])
    orc.orc_activation_backward(p(b), p(gb), b.size, p(slope), p(gsb), hw, c, ACT[act])
    assert np.array_equal(ga, gb) and np.array_equal(gsa, gsb)


def _single_layer(build, batch_shape, seed, train=True):
    """Run one layer on the reference through its API; returns the net (caller closes)."""
    net = ref_net(capi.MODE_TRAIN if train else capi.MODE_PREDICT)
    w, h, c, n = batch_shape
    net.set_input_shape(w, h, c, n)
    # a leading identity-like conv gives the layer under test a source WITH a gradient buffer
    net.conv(c, 1, 1, 0, 1, 0, "none", "input", "pre")
    build(net)
    net.compile()
    configs.init_params(net, seed=seed)
    net.set("input", configs.synth_input(net.shape("input"), seed=seed + 1))
    net.forward()
    return net


CONV_CASES = [(12, 12, 3, 2, 8, 3, 1, 1, 1), (11, 9, 4, 2, 6, 3, 2, 1, 2), (8, 8, 6, 3, 5, 1, 1, 0, 1),
              (14, 14, 3, 1, 4, 7, 2, 3, 1), (9, 9, 4, 2, 4, 5, 1, 2, 1), (7, 7, 8, 2, 16, 1, 1, 0, 1)]
# NOTE: 1x1 convolutions with stride > 1 are deliberately absent: the reference feeds the
# un-strided source straight to its GEMM when size == 1 (bcnn_conv_layer.c:445-446, 567) and so
# reads a mis-strided view; oracle and B200 path compute the true strided convolution
# (deviations).


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_backward(case):
    w, h, c, n, cout, k, s, pad, groups = case
    orc = oracle()
    net = _single_layer(lambda m: m.conv(cout, k, s, pad, groups, 0, "lrelu", "pre", "out"),
                        (w, h, c, n), seed=case[4])
    x, y = net.get("pre"), net.get("out")
    wt, b = net.get("pre_w"), net.get("pre_b")
    ho, wo = y.shape[2:]
    mine = np.zeros_like(y)
    orc.orc_conv_forward(p(x), p(wt), p(mine), n, c, h, w, cout, k, s, pad, groups)
    orc.orc_add_bias(p(mine), p(f32(b.ravel())), n, cout, ho * wo)
    orc.orc_activation_forward(p(mine), mine.size, None, ho * wo, cout, ACT["lrelu"])
    assert_close(mine, y, 1e-5, "conv fwd")
    # backward with an injected output gradient
    dy = f32(rng(1).uniform(-1, 1, size=y.shape))
    net.set("out", dy, grad=True)
    net.backward()
    g = dy.copy()
    orc.orc_activation_backward(p(y), p(g), g.size, None, None, ho * wo, cout, ACT["lrelu"])
    gb = np.zeros(cout, np.float32)
    orc.orc_grad_bias(p(gb), p(g), n, cout, ho * wo)
    gw, dx = np.zeros_like(wt), np.zeros_like(x)
    orc.orc_conv_backward(p(x), p(wt), p(g), p(gw), p(dx), n, c, h, w, cout, k, s, pad, groups)
    assert_close(gb, net.get("pre_b", grad=True).ravel(), 1e-5, "grad bias")
    assert_close(gw, net.get("pre_w", grad=True), 1e-5, "wgrad")
    assert_close(dx, net.get("pre", grad=True), 1e-5, "dgrad")
    net.close()


@pytest.mark.parametrize("shape", [(10, 10, 5, 4), (7, 7, 8, 3), (1, 1, 16, 8)])
def test_fused_conv_batchnorm_train(shape):
    w, h, c, n = shape
    orc = oracle()
    net = _single_layer(lambda m: m.conv(6, 3 if w > 1 else 1, 1, 1 if w > 1 else 0, 1, 1, "relu", "pre", "out"),
                        shape, seed=w)
    k, pad = (3, 1) if w > 1 else (1, 0)
    x, y = net.get("pre"), net.get("out")
    wt, beta, gamma = net.get("pre_w"), net.get("pre_b").ravel(), net.get("pre_scales").ravel()
    raw = np.zeros_like(y)
    orc.orc_conv_forward(p(x), p(wt), p(raw), n, c, h, w, 6, k, 1, pad, 1)
    hw = y.shape[2] * y.shape[3]
    rm, rv = np.zeros(6, np.float32), np.zeros(6, np.float32)
    sm, sv = np.zeros(6, np.float32), np.zeros(6, np.float32)
    xn, xc = np.zeros_like(raw), np.zeros_like(raw)
    mine = raw.copy()
    orc.orc_bn_forward(p(mine), n, 6, hw, p(rm), p(rv), p(f32(gamma)), p(f32(beta)), p(sm), p(sv),
                       p(xn), p(xc), 1)
    orc.orc_activation_forward(p(mine), mine.size, None, hw, 6, ACT["relu"])
    assert_close(mine, y, 1e-5, "conv+bn+relu fwd")
    node = net.lib.bcnn_b200_num_nodes(net.handle) - 1
    ref_mean, ref_var = net.bn_saved_stats(node)
    assert_close(sm, ref_mean, 1e-5, "saved mean")
    assert np.abs(sv - ref_var).max() <= 1e-5 * (np.abs(ref_var).max() + np.abs(ref_mean).max() ** 2)
    assert_close(rm, net.get("pre_run_mean").ravel(), 1e-5, "running mean")
    dy = f32(rng(2).uniform(-1, 1, size=y.shape))
    net.set("out", dy, grad=True)
    net.backward()
    g = dy.copy()
    orc.orc_activation_backward(p(y), p(g), g.size, None, None, hw, 6, ACT["relu"])
    gg, gb = np.zeros(6, np.float32), np.zeros(6, np.float32)
    dm, dv = np.zeros(6, np.float32), np.zeros(6, np.float32)
    orc.orc_bn_backward(p(g), n, 6, hw, p(f32(gamma)), p(gg), p(gb), p(sm), p(sv), p(dm), p(dv),
                        p(xn), p(xc))
    assert_close(gb, net.get("pre_b", grad=True).ravel(), 2e-5, "g_beta")
    assert_close(gg, net.get("pre_scales", grad=True).ravel(), 2e-5, "g_gamma")
    assert_close(g, net.get("out", grad=True), 2e-5, "bn dx")
    net.close()


@pytest.mark.parametrize("k,s,pad", [(2, 2, capi.PAD_SAME), (3, 2, capi.PAD_SAME), (2, 1, capi.PAD_SAME),
                                     (3, 2, capi.PAD_VALID), (3, 2, capi.PAD_CAFFE)])
@pytest.mark.parametrize("hw", [(13, 13), (8, 10)])
def test_maxpool_bit_exact(k, s, pad, hw):
    h, w = hw
    orc = oracle()
    net = _single_layer(lambda m: m.maxpool(k, s, pad, "pre", "out"), (w, h, 3, 2), seed=k * 7 + s)
    x, y = net.get("pre"), net.get("out")
    node = net.lib.bcnn_b200_num_nodes(net.handle) - 1
    idx_ref = net.maxpool_indexes(node)
    ho, wo = y.shape[2:]
    assert (ho, wo) == (orc.orc_maxpool_out_dim(h, k, s, pad), orc.orc_maxpool_out_dim(w, k, s, pad))
    mine, idx = np.zeros_like(y), np.zeros(y.shape, np.int32)
    orc.orc_maxpool_forward(p(x), p(mine), p(idx), 2, 3, h, w, k, s, ho, wo)
    assert np.array_equal(idx, idx_ref) and np.array_equal(mine, y)
    dy = f32(rng(3).uniform(-1, 1, size=y.shape))
    net.set("out", dy, grad=True)
    net.backward()
    dx = np.zeros_like(x)
    orc.orc_maxpool_backward(p(dx), p(dy), p(idx), dy.size)
    assert np.array_equal(dx, net.get("pre", grad=True))
    net.close()


def test_avgpool_depthwise_fc_softmax():
    orc = oracle()
    # depthwise
    net = _single_layer(lambda m: m.depthwise(3, 2, 1, "relu", "pre", "out"), (9, 9, 4, 2), seed=9)
    x, y, wt, b = net.get("pre"), net.get("out"), net.get("pre_w"), net.get("pre_b")
    mine = np.zeros_like(y)
    orc.orc_depthwise_forward(p(x), p(f32(wt.ravel())), p(mine), 2, 4, 9, 9, 3, 2, 1)
    orc.orc_add_bias(p(mine), p(f32(b.ravel())), 2, 4, 25)
    orc.orc_activation_forward(p(mine), mine.size, None, 25, 4, ACT["relu"])
    assert_close(mine, y, 1e-6, "depthwise fwd")
    dy = f32(rng(4).uniform(-1, 1, size=y.shape))
    net.set("out", dy, grad=True)
    net.backward()
    g = dy.copy()
    orc.orc_activation_backward(p(y), p(g), g.size, None, None, 25, 4, ACT["relu"])
    gw, dx = np.zeros(36, np.float32), np.zeros_like(x)
    orc.orc_depthwise_backward(p(x), p(f32(wt.ravel())), p(g), p(gw), p(dx), 2, 4, 9, 9, 3, 2, 1)
    assert_close(gw, net.get("pre_w", grad=True).ravel(), 1e-6, "depthwise wgrad")
    assert_close(dx, net.get("pre", grad=True), 1e-6, "depthwise dgrad")
    net.close()

    # avgpool -> fc -> softmax
    def build(m):
        m.avgpool("pre", "gap")
        m.fullc(5, "none", "gap", "fc")
        m.softmax("fc", "out")
    net = _single_layer(build, (6, 6, 7, 3), seed=11)
    x, gap, fc, sm = net.get("pre"), net.get("gap"), net.get("fc"), net.get("out")
    a = np.zeros_like(gap)
    orc.orc_avgpool_forward(p(x), p(a), 3, 7, 36)
    assert_close(a, gap, 1e-6, "avgpool")
    wt, b = net.get("gap_w"), net.get("gap_b")
    f = np.zeros_like(fc)
    orc.orc_fc_forward(p(gap), p(wt), p(f32(b.ravel())), p(f), 3, 7, 1, 5)
    assert_close(f, fc, 1e-6, "fc")
    s = np.zeros_like(sm)
    orc.orc_softmax_forward(p(fc), p(s), 3, 5, 1)
    assert_close(s, sm, 1e-6, "softmax")
    net.close()


def test_concat_upsample_bit_exact():
    """YOLO second-head glue: upsample x2 and channel concat, forward and backward."""
    orc = oracle()

    def build(m):
        m.conv(5, 3, 2, 1, 1, 0, "lrelu", "pre", "small")      # 10x10 -> 5x5
        m.upsample(2, "small", "up")                             # back to 10x10
        m.concat(["up", "pre"], "out")                           # 5 + 4 channels
    net = _single_layer(build, (10, 10, 4, 3), seed=21)
    pre, small, up, out = (net.get(k) for k in ("pre", "small", "up", "out"))
    mine = np.zeros_like(up)
    orc.orc_upsample_forward(p(small), p(mine), 3, 5, 5, 5, 2)
    assert np.array_equal(mine, up)
    cat = np.zeros_like(out)
    orc.orc_concat_forward(p(up), p(cat), 3, 5 * 100, 9 * 100, 0)
    orc.orc_concat_forward(p(pre), p(cat), 3, 4 * 100, 9 * 100, 5 * 100)
    assert np.array_equal(cat, out)
    dy = f32(rng(6).uniform(-1, 1, size=out.shape))
    net.set("out", dy, grad=True)
    # run only the two glue nodes backward: their source gradients start from zero
    net.set("up", np.zeros_like(up), grad=True)
    net.set("pre", np.zeros_like(pre), grad=True)
    net.set("small", np.zeros_like(small), grad=True)
    net.backward()
    g_up = np.zeros_like(up)
    orc.orc_concat_backward(p(dy), p(g_up), 3, 5 * 100, 9 * 100, 0)
    assert np.array_equal(g_up, net.get("up", grad=True))
    g_small = np.zeros_like(small)
    orc.orc_upsample_backward(p(g_up), p(g_small), 3, 5, 5, 5, 2)
    # small.grad is then run through the conv's leaky-ReLU backward in place by the reference
    fac = np.where(small > 0, 1.0, 0.1).astype(np.float32)
    assert_close(g_small * fac, net.get("small", grad=True), 1e-6, "upsample bwd (+ lrelu')")
    net.close()


def test_sgd_update_bit_exact():
    lib, orc = ref_lib(), oracle()
    lib.bcnn_sgd_update_cpu.argtypes = [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_float] * 3
    r = rng(8)
    for n in (5, 16, 1000):
        w0, g0 = f32(r.normal(size=n)), f32(r.normal(size=n))
        b0, gb0 = f32(r.normal(size=7)), f32(r.normal(size=7))
        a = [w0.copy(), b0.copy(), g0.copy(), gb0.copy()]
        b = [w0.copy(), b0.copy(), g0.copy(), gb0.copy()]
        lib.bcnn_sgd_update_cpu(p(a[0]), p(a[1]), p(a[2]), p(a[3]), n, 7, 64, 0.003, 0.9, 0.0005)
        orc.orc_sgd_update(p(b[0]), p(b[1]), p(b[2]), p(b[3]), n, 7, 64, 0.003, 0.9, 0.0005)
        for u, v in zip(a, b):
            assert np.array_equal(u, v)


def test_adam_update_bit_exact():
    """orc_adam_update against the reference's bcnn_adam_update_cpu (src/bcnn_learner.c:106-132):
    three consecutive steps (moments carried over), sizes with and without an n % 8 tail, and
    vanishing second moments so the tail's division guard (bcnn_vdiv) is taken."""
    lib, orc = ref_lib(), oracle()
    lib.bcnn_adam_update_cpu.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_float] * 5
    r = rng(9)
    for n in (5, 16, 1003):
        state = [f32(r.normal(size=n)), f32(r.normal(size=7)), None, None,
                 np.zeros(n, np.float32), np.zeros(n, np.float32)]
        a = [None if x is None else x.copy() for x in state]
        b = [None if x is None else x.copy() for x in state]
        for step in range(3):
            g = f32(r.normal(scale=1e-2, size=n))
            if step == 0:
                g[-3:] = 0.0  # zero moments in the scalar tail: guarded quotient
                g[0] = 1e-9   # tiny moments in the vector body: unguarded quotient
            gb = f32(r.normal(size=7))
            for side, fn in ((a, lib.bcnn_adam_update_cpu), (b, orc.orc_adam_update)):
                side[2], side[3] = g.copy(), gb.copy()
                fn(p(side[0]), p(side[1]), p(side[2]), p(side[3]), p(side[4]), p(side[5]), n, 7,
                   16, 16 * (step + 1), 0.9, 0.999, 0.002, 0.9, 0.0005)
            for u, v in zip(a, b):
                assert np.array_equal(u.view(np.uint32), v.view(np.uint32)), (n, step)


def _ref_conv_backward(cin, cout, k, hw, batch=2, seed=1):
    """dgrad / wgrad of one convolution through the reference's public API (a 1x1 convolution in
    front gives its source tensor a gradient buffer), and the oracle's on the same inputs."""
    orc = oracle()
    net = ref_net(mode=capi.MODE_TRAIN, threads=1)
    net.set_input_shape(hw, hw, 3, batch)
    net.conv(cin, 1, 1, 0, 1, 0, "none", "input", "c0")
    net.conv(cout, k, 1, k // 2, 1, 0, "none", "c0", "c1")
    net.compile()
    r = rng(seed)
    net.set("input", f32(r.uniform(-1, 1, (batch, 3, hw, hw))))
    net.forward()
    c0, w = net.get("c0").copy(), net.get("c0_w").copy()
    dy = f32(r.uniform(-1, 1, (batch, cout, hw, hw)))
    gw0 = net.get("c0_w", grad=True).copy()
    net.set("c1", dy, grad=True)
    net.backward()
    dx_ref, gw_ref = net.get("c0", grad=True).copy(), net.get("c0_w", grad=True).copy()
    net.close()
    gw, dx = gw0.copy(), np.zeros_like(c0)
    orc.orc_conv_backward(p(c0), p(w), p(dy), p(gw), p(dx), batch, cin, hw, hw, cout, k, 1, k // 2, 1)
    # float64 evaluation of the definition (1x1 only): dx[n,ci,p] = sum_co W[co,ci] dy[n,co,p]
    exact = None
    if k == 1:
        exact = np.einsum("oi,nop->nip", w.reshape(cout, cin).astype(np.float64),
                          dy.reshape(batch, cout, -1).astype(np.float64)).reshape(c0.shape)
    return dx_ref, gw_ref, dx, gw, exact


def test_reference_transposed_gemm_blocking_h12():
    """Hazard H12: `sgemm` (the transposed-operand driver behind bcnn_gemm, src/kernels/bcnn_mat.c
    :2588-2625) addresses K block l of a transposed A as `&A[... + l * KC]` and N block j of a
    transposed B as `&B[... + j * NC]`, without the column increment. Inside the first block
    (K <= KC = 384, N <= NC = 4096) the reference is right and the oracle matches it bit for bit;
    past it the reference's convolution backward is wrong (not merely different): its data gradient
    for Cout > 384, its weight gradient for Cin * k * k > 4096. The oracle restates the intended
    arithmetic; the float64 evaluation of the definition says which side is correct."""
    def err(a, b):
        return float(np.abs(a - b).max() / np.abs(b).max())
    dx_ref, gw_ref, dx, gw, exact = _ref_conv_backward(16, 384, 1, 8)
    assert np.array_equal(dx_ref, dx) and np.array_equal(gw_ref, gw)
    assert err(dx, exact) < 1e-6
    dx_ref, gw_ref, dx, gw, exact = _ref_conv_backward(16, 385, 1, 8)     # K block 1 has one row
    assert np.array_equal(gw_ref, gw)
    assert err(dx, exact) < 1e-6 and err(dx_ref, exact) > 1e-3
    dx_ref, gw_ref, dx, gw, exact = _ref_conv_backward(16, 512, 1, 8)     # YOLOv3-tiny's 256 -> 512 class
    assert err(dx, exact) < 1e-6 and err(dx_ref, exact) > 0.3
    dx_ref, gw_ref, dx, gw, _ = _ref_conv_backward(455, 32, 3, 8)         # Cin k k = 4095: in range
    assert np.array_equal(dx_ref, dx) and err(gw, gw_ref) < 1e-5
    dx_ref, gw_ref, dx, gw, _ = _ref_conv_backward(500, 32, 3, 8)         # Cin k k = 4500 > NC
    assert np.array_equal(dx_ref, dx) and err(gw, gw_ref) > 0.3
