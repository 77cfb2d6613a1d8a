import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from bcnn_b200 import capi, configs
from helpers import f32, oracle, p, rel_err
orc = oracle()
for mode in (capi.MODE_VALID, capi.MODE_PREDICT):
    net = capi.Net(mode=mode)
    net.set_input_shape(12, 12, 3, 2)
    net.conv(8, 3, 1, 1, 1, 1, "relu", "input", "c1")
    net.compile()
    configs.init_params(net, seed=3)
    x = configs.synth_input(net.shape("input"), seed=4)
    net.set("input", x)
    net.forward()
    y = net.get("c1")
    prm = {k: net.get(k) for k in ("input_w", "input_b", "input_scales", "input_run_mean", "input_run_var")}
    raw = np.zeros_like(y)
    orc.orc_conv_forward(p(x), p(f32(prm["input_w"])), p(raw), 2, 3, 12, 12, 8, 3, 1, 1, 1)
    conv = raw.copy()
    rm, rv = f32(prm["input_run_mean"]).ravel().copy(), f32(prm["input_run_var"]).ravel().copy()
    sm, sv = np.zeros(8, np.float32), np.zeros(8, np.float32)
    orc.orc_bn_forward(p(raw), 2, 8, 144, p(rm), p(rv), p(f32(prm["input_scales"]).ravel().copy()),
                       p(f32(prm["input_b"]).ravel().copy()), p(sm), p(sv), None, None, mode)
    orc.orc_activation_forward(p(raw), raw.size, None, 144, 8, capi.ACT["relu"])
    print("mode", mode, rel_err(y, raw))
    print(" y", y[0,0,0,:6], "want", raw[0,0,0,:6], "conv", conv[0,0,0,:6], "scale", prm["input_scales"].ravel()[:2], "b", prm["input_b"].ravel()[:2])
    net.close()
