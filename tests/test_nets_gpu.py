"""Whole-net parity on the GPU: the BASELINE-style nets of tests/netcases.py are built
through the SAME bcnn C API calls on libbcnn_b200.so and compared with
  (1) the committed golden fixtures recorded from the reference CPU library, and
  (2) the reference library itself, live, when oracle/_ref travelled to this box.

Tolerances (normalised per tensor, helpers.rel_err): 1e-5 per kernel is the bar asserted in
tests/test_kernels_gpu.py. Across a whole net, rounding noise compounds through 10-20
chained layers and, with tiny-batch batch-norm, is amplified chaotically over SGD steps: the
reference ITSELF moves by up to 1e-4 (step 0) and 3e-2 (after 3 steps, cifar batch 4) when
its input is perturbed by one ulp. Each tensor is therefore held to
max(floor, 8 x the reference's own 1-ulp response recorded beside the golden), with floors
2e-5 at step 0 and 1e-4 after the SGD steps. Tensor-core path floors: 2e-2 / 5e-2.
Arg-max indices are compared as a mismatch RATE here (a 1-ulp upstream difference can
legitimately flip a tie); bit-exactness on identical input bits is asserted at kernel level.
"""
import numpy as np
import pytest

import netcases
from bcnn_b200 import capi, configs
from helpers import GOLDEN, assert_close, ref_available, ref_net, rel_err

pytestmark = pytest.mark.gpu


def _compare(out, golden, tol_s0, tol_final, what, noise=None, grad_l2_tol=None):
    """noise: expected relative operand perturbation of the path under test (BF16 rounding
    for the tensor-core path); the per-tensor tolerance then scales with the case's measured
    amplification (reference response / 2e-7 input perturbation)."""
    worst = (0.0, "")
    checked = 0
    for key in golden:
        if key.startswith("sens:"):
            continue
        base = key[:-4] if key.endswith("@sub") else key
        assert base in out, f"{what}: tensor {base} missing on the B200 path"
        got, want = out[base], golden[key]
        if key.endswith("@sub"):
            idx = np.linspace(0, got.size - 1, want.size).astype(np.int64)
            got = got.ravel()[idx]
        if "/argmax/" in base:
            mismatch = float(np.mean(got.ravel() != want.ravel()))
            assert mismatch <= 2e-3, f"{what}: {base} argmax mismatch rate {mismatch:.2e}"
            continue
        tol = tol_final if base.startswith("final/") or not base.startswith("s0/") else tol_s0
        if base.endswith("/cost"):
            assert abs(float(got.ravel()[0]) - float(want.ravel()[0])) <= 1.0, f"{what}: {base}"
            continue
        if np.abs(want).max(initial=0.0) == 0.0:
            assert np.abs(got).max(initial=0.0) <= 1e-6, f"{what}: {base} should be zero"
            continue
        # floor, or 8x the reference's own response to a 1-ulp input perturbation
        sens = float(golden.get("sens:" + base, 0.0))
        tol = max(tol, 8.0 * sens if noise is None else (sens / 2e-7) * noise)
        e = max(rel_err(got, want))
        if grad_l2_tol is not None and "/grad/" in base:
            # ReLU-mask flips (a pre-activation within rounding noise of zero changes sign) put
            # a few full-size errors into a gradient tensor and tiny-batch batchnorm spreads
            # them: the max-abs metric is meaningless there, the L2 metric is not
            e, tol = rel_err(got, want)[1], max(tol, grad_l2_tol)
        if e / tol > worst[0]:
            worst = (e / tol, f"{base} (err {e:.2e}, tol {tol:.2e})")
        assert e <= tol, f"{what}: {base} rel err {e:.3e} > {tol:.3e}"
        checked += 1
    print(f"[{what}] {checked} tensors checked, worst err/tol {worst[0]:.2f} at {worst[1]}")


@pytest.mark.parametrize("name", [n for n in netcases.CASES if (GOLDEN / f"{n}.npz").exists()])
def test_net_matches_golden_fp32(name):
    golden = dict(np.load(GOLDEN / f"{name}.npz"))
    net = capi.Net()
    out = netcases.run_case(net, name)
    net.close()
    _compare(out, golden, 2e-5, 1e-4, f"{name} fp32 vs golden")


@pytest.mark.parametrize("math", [capi.MATH_TC, capi.MATH_TC_BF16], ids=["tc", "resident"])
@pytest.mark.parametrize("name", ["chain_b4", "resnet_small_b4"])
def test_net_matches_golden_tensor_core(name, math):
    golden = dict(np.load(GOLDEN / f"{name}.npz"))
    net = capi.Net()
    net.set_conv_math(math)
    out = netcases.run_case(net, name)
    net.close()
    golden = {k: v for k, v in golden.items() if "/argmax/" not in k}
    # TF32 / BF16 operands: 2^-11 .. 2^-9 per element, <= 1e-3 effective after the dot products
    # average it. Forward tensors and the loss are held to 2e-2; gradients of the ReLU nets to
    # an L2-relative 0.15 (see _compare: mask flips at batch 4)
    # Resident mode additionally rounds every activation and gradient tensor to BF16 (2^-9), which
    # flips more ReLU masks: at batch 4 gradients reach 0.25 .. 0.4, i.e. this case only checks that
    # they are sane (a missing or doubled branch contribution is >= 1). The decision-free twin at
    # batch 64 (test_baseline_parity_gpu.py) holds the same chained arithmetic to 2e-2 in both modes.
    _compare(out, golden, 2e-2, 5e-2, f"{name} tc vs golden", noise=1e-3 if math == capi.MATH_TC else 4e-3,
             grad_l2_tol=0.15 if math == capi.MATH_TC else 0.6)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref was not built / did not travel")
@pytest.mark.parametrize("name", ["mnist_b8", "chain_b4"])
def test_net_matches_live_reference(name):
    """Same calls, same seed (different from the golden's), both libraries in one process."""
    want = netcases.live_reference_case(name, 31, threads=1)
    net = capi.Net()
    out = netcases.run_case(net, name, seed=31)
    net.close()
    _compare(out, want, 2e-5, 1e-4, f"{name} fp32 vs live reference")


def test_library_defaults(monkeypatch):
    """Without the test session's environment a net computes on the tensor cores with resident BF16
    NHWC activations and batch-correct residual semantics; the FP32-tensor tensor-core mode, the FP32
    path and the reference quirks are opt-in."""
    monkeypatch.delenv("BCNN_B200_CONV_MATH")
    monkeypatch.delenv("BCNN_B200_REFERENCE_QUIRKS")
    net = capi.Net()
    assert net.lib.bcnn_b200_get_conv_math(net.handle) == capi.MATH_TC_BF16
    # ... which bcnn_compile_net keeps for nets whose layers between convolutions know the format
    # (ResNet-style: conv, pooling, residual add) and drops for nets that would convert around
    # format-unaware nodes (here: stand-alone batch norm)
    netcases.small_resnet(net, batch=2)
    net.compile()
    assert net.lib.bcnn_b200_get_conv_math(net.handle) == capi.MATH_TC_BF16
    net.close()
    net = capi.Net()
    configs.mnist(net, batch=2)
    net.compile()
    assert net.lib.bcnn_b200_get_conv_math(net.handle) == capi.MATH_TC
    net.close()
    monkeypatch.setenv("BCNN_B200_CONV_MATH", "tc")
    net = capi.Net()
    assert net.lib.bcnn_b200_get_conv_math(net.handle) == capi.MATH_TC
    assert net.lib.bcnn_b200_get_reference_quirks(net.handle) == 0
    net.close()
    monkeypatch.setenv("BCNN_B200_CONV_MATH", "fp32")
    monkeypatch.setenv("BCNN_B200_REFERENCE_QUIRKS", "1")
    net = capi.Net()
    assert net.lib.bcnn_b200_get_conv_math(net.handle) == capi.MATH_FP32
    assert net.lib.bcnn_b200_get_reference_quirks(net.handle) == 1
    net.close()


def test_residual_fix_mode_accumulates_branch_gradients():
    """reference_quirks OFF: the block input's gradient is the sum of both branches.
    Checked on the last (identity-shortcut) block at batch 1, where everything downstream
    of the block is identical in both modes: prev.grad(off) == prev.grad(on) + out.grad."""
    grads = {}
    for quirks in (True, False):
        net = capi.Net()
        net.set_reference_quirks(quirks)
        netcases.small_resnet(net, batch=1)
        net.compile()
        configs.init_params(net, seed=5)
        net.set("input", configs.synth_input(net.shape("input"), seed=6))
        net.set("label", configs.synth_labels(net.shape("label")))
        net.forward()
        net.backward()
        grads[quirks] = (net.get("s1b0_out", grad=True), net.get("s1b1_out", grad=True),
                         net.get("s1b1_out"), net.get("s1b0_out"))
        net.close()
    prev_on, out_on, y_on, prev_data = grads[True]
    prev_off, out_off, y_off, _ = grads[False]
    assert_close(y_off, y_on, 1e-6, "forward identical at batch 1")
    assert_close(out_off, out_on, 1e-6, "block output gradient identical")
    # what is fetched is prev.grad after the PREVIOUS block's eltwise backward has applied
    # its ReLU derivative in place, hence the mask
    mask = (prev_data > 0).astype(np.float32)
    assert_close(prev_off, prev_on + out_on * mask, 1e-5, "branch gradients accumulate")


def test_input_pipeline_matches_upload_then_step():
    """bcnn_b200_train_step(upload_inputs=2): batch i+1 travels on the copy stream while step i
    computes; losses and final parameters are bit-identical to upload-then-step."""
    batches = [(configs.synth_input((4, 3, 20, 20), seed=40 + i),
                configs.synth_labels((4, 10, 1, 1), first_sample=3 * i)) for i in range(4)]
    runs = {}
    for mode in ("sequential", "pipelined"):
        net = capi.Net()
        netcases.chain_convnet(net, batch=4)
        net.compile()
        configs.init_params(net, seed=9)
        losses = []
        if mode == "pipelined":
            net.set_host("input", batches[0][0])
            net.set_host("label", batches[0][1])
            net.prefetch_inputs()
        for i in range(len(batches)):
            nxt = batches[min(i + 1, len(batches) - 1)] if mode == "pipelined" else batches[i]
            net.set_host("input", nxt[0])
            net.set_host("label", nxt[1])
            losses.append(net.train_step(upload_inputs=2 if mode == "pipelined" else True,
                                         fetch_loss=True))
        net.sync()
        params = {name: net.get(idx) for idx, name, _ in configs.param_tensors(net)}
        net.close()
        runs[mode] = (losses, params)
    assert runs["sequential"][0] == runs["pipelined"][0], "per-step losses differ"
    assert len(set(runs["sequential"][0])) > 1, "batches must differ for the test to bite"
    for name, w in runs["sequential"][1].items():
        assert np.array_equal(w, runs["pipelined"][1][name]), name


def test_valid_and_predict_modes_follow_reference_semantics():
    """VALID normalises with the running statistics; PREDICT applies y = gamma*x + beta."""
    results = {}
    for mode in (capi.MODE_VALID, capi.MODE_PREDICT):
        net = capi.Net(mode=mode)
        net.set_input_shape(12, 12, 3, 2)
        net.conv(8, 3, 1, 1, 1, 1, "relu", "input", "c1")
        net.compile()
        configs.init_params(net, seed=3)
        x = configs.synth_input(net.shape("input"), seed=4)
        net.set("input", x)
        net.forward()
        idx = {n: net.tensor_index(n) for n in ("input_w", "input_b", "input_scales",
                                                "input_run_mean", "input_run_var")}
        results[mode] = (net.get("c1"), {k: net.get(v) for k, v in idx.items()}, x)
        net.close()
    from helpers import f32, oracle, p
    orc = oracle()
    for mode, (y, prm, x) in results.items():
        raw = np.zeros_like(y)
        orc.orc_conv_forward(p(x), p(f32(prm["input_w"])), p(raw), 2, 3, 12, 12, 8, 3, 1, 1, 1)
        rm, rv = f32(prm["input_run_mean"]).ravel().copy(), f32(prm["input_run_var"]).ravel().copy()
        sm, sv = np.zeros(8, np.float32), np.zeros(8, np.float32)
        # keep the arrays alive across the call: p() hands out a raw pointer
        gamma, beta = f32(prm["input_scales"]).ravel().copy(), f32(prm["input_b"]).ravel().copy()
        orc.orc_bn_forward(p(raw), 2, 8, 144, p(rm), p(rv), p(gamma), p(beta), p(sm), p(sv), None,
                           None, mode)
        orc.orc_activation_forward(p(raw), raw.size, None, 144, 8, capi.ACT["relu"])
        assert_close(y, raw, 2e-5, f"mode {mode}")


def test_train_and_predict_on_batch_are_the_three_loops():
    """bcnn_train_on_batch = upload + forward + backward + update + loss; bcnn_predict_on_batch =
    upload + forward + output (reference src/bcnn_net.c:452-483, minus the file loader)."""
    nets = []
    for _ in range(2):
        net = capi.Net()
        info = netcases.chain_convnet(net, batch=4)
        net.compile()
        configs.init_params(net, seed=9)
        nets.append(net)
    a, b = nets
    x = configs.synth_input(a.shape("input"), seed=10)
    y = configs.synth_labels(a.shape("label"))
    for step in range(2):
        a.set_host("input", x + step)
        a.set_host("label", y)
        loss_a = a.train_on_batch()
        b.set("input", x + step)
        b.set("label", y)
        b.forward()
        b.backward()
        b.update()
        assert loss_a == b.loss()
    for idx, name, _ in configs.param_tensors(a):
        assert np.array_equal(a.get(idx), b.get(idx)), name
    for n in (a, b):
        n.set_mode(capi.MODE_VALID)
    a.set_host("input", x)
    loss, out = a.predict_on_batch()
    b.set("input", x)
    b.forward()
    assert np.array_equal(out, b.get(info["out"])) and loss == b.loss()
    assert out.shape == a.shape(info["out"])
    a.close()
    b.close()


def _two_input_net(net, batch=2):
    """A second input tensor (bcnn_add_input) joining the trunk through a concat."""
    net.set_input_shape(8, 8, 3, batch)
    net.add_input(8, 8, 2, "aux")
    net.conv(4, 3, 1, 1, 1, 0, "relu", "input", "c_main")
    net.conv(4, 3, 1, 1, 1, 0, "relu", "aux", "c_aux")
    net.concat(["c_main", "c_aux"], "cat")
    net.conv(6, 1, 1, 0, 1, 0, "none", "cat", "head")
    net.cost("head", "cost", metric=capi.METRIC_SSE)
    net.sgd(0.01, 0.9, 0.0005)
    net.compile()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref was not built / did not travel")
def test_second_input_tensor_matches_live_reference():
    outs = []
    for net in (capi.Net(), ref_net()):
        _two_input_net(net)
        configs.init_params(net, seed=4)
        net.set("input", configs.synth_input(net.shape("input"), seed=5))
        net.set("aux", configs.synth_input(net.shape("aux"), seed=6))
        net.set("label", configs.synth_input(net.shape("label"), seed=7))
        net.forward()
        net.backward()
        net.update()
        outs.append({k: net.get(k) for k in ("head", "c_aux")} |
                    {"g:" + k: net.get(k, grad=True) for k in ("cat", "aux_w")} |
                    {"w:" + k: net.get(k) for k in ("aux_w", "input_w", "cat_w")})
        assert net.tensor_index("aux") == 2  # right after input (0) and label (1), as upstream
        net.close()
    for k in outs[1]:
        assert_close(outs[0][k], outs[1][k], 2e-5, f"two-input net {k}")


@pytest.mark.parametrize("math", [capi.MATH_FP32, capi.MATH_TC])
def test_predict_forward_graph_replay_is_the_eager_forward(math):
    """PREDICT-mode bcnn_forward: 1st call eager, 2nd captured into a CUDA graph, later calls one
    graph launch (bcnn_b200_set_graphs). Same kernels in the same order, so every output must be
    bit-identical to a graph-free twin, for fresh inputs on every call; changing the conv math or
    recompiling rebuilds the graph."""
    nets = []
    for graphs in (True, False):
        net = capi.Net(mode=capi.MODE_PREDICT)
        net.set_graphs(graphs)
        net.set_conv_math(math)
        info = configs.yolo_tiny(net, batch=2, res=64)
        net.compile()
        configs.init_params(net, seed=3)
        nets.append(net)
    fast, eager = nets
    assert fast.graphs() == 1 and eager.graphs() == 0
    for step in range(5):
        x = configs.synth_input(fast.shape("input"), seed=50 + step)
        for net in nets:
            net.set("input", x)
            net.forward()
        assert fast.graphs() == (2 if step >= 1 else 1)
        for name in ("conv0", "route", "yolo1", info["out"]):
            assert np.array_equal(fast.get(name), eager.get(name)), (step, name)
    before = fast.get(info["out"])
    fast.set_conv_math(capi.MATH_FP32 if math == capi.MATH_TC else capi.MATH_TC)
    fast.forward()                      # new configuration: eager again, graph dropped
    assert fast.graphs() == 1
    fast.set_conv_math(math)
    for _ in range(3):
        fast.forward()
    assert fast.graphs() == 2 and np.array_equal(fast.get(info["out"]), before)
    fast.compile()                      # reallocates the input and the workspace
    assert fast.graphs() == 1
    for net in nets:
        net.close()


@pytest.mark.parametrize("case,math", [("mnist", capi.MATH_FP32), ("chain", capi.MATH_FP32),
                                       ("chain", capi.MATH_TC), ("resnet", capi.MATH_TC)])
def test_train_step_graph_replay_is_the_eager_step(case, math):
    """bcnn_b200_train_step: 1st step eager, then forward + backward are captured per input
    buffer (the input pipeline alternates two) and replayed; the update stays eager. Losses and
    every parameter must be bit-identical to a graph-free twin over upload-then-step and
    pipelined steps, with a fresh batch every step."""
    def build(net):
        if case == "mnist":
            return configs.mnist(net, batch=8)
        if case == "chain":
            return netcases.chain_convnet(net, batch=4)
        net.set_reference_quirks(False)
        return netcases.small_resnet(net, batch=4)

    nets = []
    for graphs in (True, False):
        net = capi.Net()
        net.set_graphs(graphs)
        net.set_conv_math(math)
        build(net)
        net.compile()
        configs.init_params(net, seed=21)
        nets.append(net)
    fast, eager = nets
    y = configs.synth_labels(fast.shape("label"))
    batches = [configs.synth_input(fast.shape("input"), seed=70 + i) for i in range(9)]
    for step in range(4):                      # upload, then step (one input buffer)
        losses = []
        for net in nets:
            net.set_host("input", batches[step])
            net.set_host("label", y)
            losses.append(net.train_step(upload_inputs=True, fetch_loss=True))
        assert losses[0] == losses[1], (step, losses)
    assert fast.graphs() == 2 and eager.graphs() == 0
    for net in nets:                           # pipelined: two input buffers alternate
        net.set_host("input", batches[4])
        net.prefetch_inputs()
    for step in range(4, 8):
        losses = []
        for net in nets:
            net.set_host("input", batches[step + 1])
            losses.append(net.train_step(upload_inputs=2, fetch_loss=True))
        assert losses[0] == losses[1], (step, losses)
    for idx, name, _ in configs.param_tensors(fast):
        assert np.array_equal(fast.get(idx), eager.get(idx)), name
    # the SGD update kernels ride in the graph while the learning rate is constant: a new rate
    # must re-record them, a schedule must take them out again
    for phase, change in enumerate((lambda n: n.sgd(0.02, 0.8, 0.001),
                                    lambda n: n.lr_policy(1, scale=0.5, step=1))):
        for net in nets:
            change(net)
        for step in range(3):
            losses = []
            for net in nets:
                net.set_host("input", batches[step])
                losses.append(net.train_step(upload_inputs=True, fetch_loss=True))
            assert losses[0] == losses[1], (phase, step, losses)
        for idx, name, _ in configs.param_tensors(fast):
            assert np.array_equal(fast.get(idx), eager.get(idx)), (phase, name)
    fast.forward()                             # the plain loops still work beside the graphs
    eager.forward()
    assert np.array_equal(fast.get("cost"), eager.get("cost"))
    for net in nets:
        net.close()


def test_pack_table_step_is_the_per_call_packing_step(monkeypatch):
    """Packed weight images kept by the net and rebuilt by one launch per TRAIN forward
    (bcnn_net.c:packs_prepare) against every convolution call packing its own (BCNN_B200_PACK_TABLE=0):
    same bytes in the images, so losses and every parameter after three steps with weight updates in
    between must be bit-identical -- stale images would show from the second step on."""
    nets = []
    for table in ("1", "0"):
        monkeypatch.setenv("BCNN_B200_PACK_TABLE", table)
        net = capi.Net()
        net.set_conv_math(capi.MATH_TC_BF16)
        net.set_reference_quirks(False)
        netcases.small_resnet(net, batch=4)
        net.compile()
        configs.init_params(net, seed=33)
        nets.append(net)
    y = configs.synth_labels(nets[0].shape("label"))
    for step in range(3):
        losses = []
        for table, net in zip(("1", "0"), nets):
            monkeypatch.setenv("BCNN_B200_PACK_TABLE", table)   # read when the table is first built
            net.set_host("input", configs.synth_input(net.shape("input"), seed=90 + step))
            net.set_host("label", y)
            losses.append(net.train_step(upload_inputs=True, fetch_loss=True))
        assert losses[0] == losses[1], (step, losses)
    for idx, name, _ in configs.param_tensors(nets[0]):
        assert np.array_equal(nets[0].get(idx), nets[1].get(idx)), name
    for net in nets:
        net.close()
