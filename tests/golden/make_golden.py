"""Regenerate tests/golden/*.npz by running the UNMODIFIED reference CPU library
(oracle/_ref/libbcnn_ref.so, built from /root/reference by oracle/Makefile) through its own
public API on the seeded synthetic cases of tests/netcases.py.

    python tests/golden/make_golden.py [case ...]     (default: every case of netcases.CASES)
    python tests/golden/make_golden.py model_io       (weight-file fixtures, see make_model_io)
    python tests/golden/make_golden.py cfg            (config-file fixtures, see make_cfg)

The reference ships no golden vectors of its own, so these files are
the pin: they record what the reference computes, never hand-edited. Large tensors are
stored as a deterministic 4096-element subsample (key suffix '@sub') to keep fixtures small.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import netcases  # noqa: E402
from helpers import GOLDEN, ref_net  # noqa: E402

SUB = 4096
FULL_CASES = {"chain_b4", "resnet_small_b4", "yolo_two_heads_b2", "chain_adam_b4"}  # small enough to keep every tensor whole


def subsample(a: np.ndarray) -> np.ndarray:
    flat = a.ravel()
    if flat.size <= SUB:
        return flat.copy()
    idx = np.linspace(0, flat.size - 1, SUB).astype(np.int64)
    return flat[idx].copy()


def make_model_io():
    """Weight-file fixtures, all written or read by the reference itself:
      model_io.bcnnmodel  bcnn_save_weights of netcases.model_io_net with seeded parameters
      model_io.weights    a Darknet-layout file of the same net (written by the numpy restatement,
                          minor version 1001 so the reader transposes the fc matrix)
      model_io.npz        layout (json); saved/<tensor> the values that were saved;
                          train/<tensor>, predict/<tensor> what bcnn_load_weights leaves in a
                          TRAIN / PREDICT net (PREDICT: batch-norm folded); darknet/<tensor> the
                          same for the Darknet file in PREDICT mode; input + predict_out: one
                          PREDICT forward of the reference after loading model_io.bcnnmodel."""
    import dataclasses
    import json
    sys.path.insert(0, str(ROOT / "oracle"))
    import bcnn_model_oracle as mo
    from bcnn_b200 import capi, configs

    def built(mode):
        net = ref_net(mode=mode, threads=1)
        info = netcases.model_io_net(net)
        net.compile()
        return net, info

    def params(net, prefix):
        return {f"{prefix}/{name}": net.get(idx).ravel()
                for idx, name, _ in configs.param_tensors(net)}

    packed = {}
    src, _ = built(capi.MODE_VALID)  # VALID: init_params randomises the running statistics
    configs.init_params(src, seed=11)
    src.save_weights(GOLDEN / "model_io.bcnnmodel")
    packed.update(params(src, "saved"))
    layout = mo.net_layout(src)
    packed["layout"] = np.array(json.dumps([dataclasses.asdict(n) for n in layout]))
    rng = np.random.default_rng(12)
    values = {name: rng.uniform(0.5, 1.5, size=size).astype(np.float32)
              for name, size in mo.all_names(layout)}
    mo.write_darknet(GOLDEN / "model_io.weights", layout, values, major=0, minor=1001)
    for mode, tag, path in ((capi.MODE_TRAIN, "train", "model_io.bcnnmodel"),
                            (capi.MODE_PREDICT, "predict", "model_io.bcnnmodel"),
                            (capi.MODE_PREDICT, "darknet", "model_io.weights")):
        net, info = built(mode)
        assert net.load_weights(GOLDEN / path) == 0
        packed.update(params(net, tag))
        if tag == "predict":
            x = configs.synth_input(net.shape("input"), seed=21)
            net.set("input", x)
            net.forward()
            packed["input"] = x
            packed["predict_out"] = net.get(info["out"])
        net.close()
    np.savez_compressed(GOLDEN / "model_io.npz", **packed)
    print(f"model_io: {len(packed)} arrays, "
          f"{(GOLDEN / 'model_io.bcnnmodel').stat().st_size} + "
          f"{(GOLDEN / 'model_io.weights').stat().st_size} file bytes")


def make_cfg():
    """Config-file fixtures under tests/golden/cfg/ (the two cfg files are hand-written test
    inputs; everything else is what the reference makes of them):
      mini_bcnn.json / mini_yolo.json  graph built by the reference's bcnn_load_net
                                       (capi.Net.structure: node types, src / dst indices,
                                       tensor names and shapes, batch size)
      mini_yolo.weights                Darknet weights for mini_yolo.cfg, seeded values written by
                                       the numpy restatement in the order the reference reads
      mini_yolo.npz                    input and the tensors of one PREDICT forward of the
                                       reference after bcnn_load_net(cfg, weights), and its
                                       bcnn_yolo_get_detections per sample (thresh 0.5, 640x480
                                       frame, relative) as [num_dets, x y w h objectness p0 p1].
      mini_yolo_train.npz              the same files in TRAIN mode: box labels, the heads, the
                                       gradient its yolo loss leaves on them, gradients after
                                       bcnn_backward and weights after one bcnn_update."""
    import json
    sys.path.insert(0, str(ROOT / "oracle"))
    import bcnn_model_oracle as mo
    from bcnn_b200 import capi, configs
    cfg_dir = GOLDEN / "cfg"
    weights = cfg_dir / "mini_yolo.weights"

    net = ref_net(mode=capi.MODE_TRAIN, threads=1)
    assert net.load_net(cfg_dir / "mini_bcnn.conf") == 0
    (cfg_dir / "mini_bcnn.json").write_text(json.dumps(net.structure()))
    net.close()

    # the layout comes from a graph-only load (a model path that does not exist yet selects the
    # Darknet dialect and fails after the graph is built)
    probe = ref_net(mode=capi.MODE_PREDICT, threads=1)
    if weights.exists():
        weights.unlink()
    assert probe.load_net(cfg_dir / "mini_yolo.cfg", weights) == 1
    layout = mo.net_layout(probe)
    rng = np.random.default_rng(17)
    span = {"weights": (-0.25, 0.25), "bias": (-0.1, 0.1), "scales": (0.5, 1.5),
            "mean": (-0.2, 0.2), "var": (0.5, 1.5), "slopes": (0.05, 0.3)}
    values = {}
    for node in layout:
        for role, (name, size) in node.roles.items():
            lo, hi = span[role]
            values[name] = rng.uniform(lo, hi, size=size).astype(np.float32)
    mo.write_darknet(weights, layout, values, major=0, minor=2, seen=12800)
    probe.close()

    net = ref_net(mode=capi.MODE_PREDICT, threads=1)
    assert net.load_net(cfg_dir / "mini_yolo.cfg", weights) == 0
    (cfg_dir / "mini_yolo.json").write_text(json.dumps(net.structure()))
    net.compile()
    x = configs.synth_input(net.shape("input"), seed=23)
    net.set("input", x)
    net.forward()
    packed = {"input": x}
    for name in ("lid1", "lid7", "lid9", "lid10", "lid14", "lid16", "lid17"):
        packed[name] = net.get(name)
    for b in range(2):  # bcnn_yolo_get_detections of a 640x480 frame (chatty on stderr)
        packed[f"dets_b{b}"] = net.yolo_detections(b, 640, 480, 0.5, relative=True)
    np.savez_compressed(cfg_dir / "mini_yolo.npz", **packed)
    net.close()
    # TRAIN mode on the same files: one step with box labels through the reference's yolo loss
    net = ref_net(mode=capi.MODE_TRAIN, threads=1)
    assert net.load_net(cfg_dir / "mini_yolo.cfg", weights) == 0
    net.compile()
    x = configs.synth_input(net.shape("input"), seed=29)
    lab = configs.synth_yolo_labels(2, [4, 7], classes=2, seed=31)
    lab[0, 0, 0, 5:9] = lab[0, 0, 0, 0:4]          # two truths in one cell: the "claimed" branch
    net.set("input", x)
    net.set("label", lab)
    net.forward()
    train = {"input": x, "label": lab}
    for name in ("lid9", "lid10", "lid16", "lid17"):
        train["data/" + name] = net.get(name)
    for name in ("lid10", "lid17"):
        train["loss_grad/" + name] = net.get(name, grad=True)
    net.backward()
    for name in ("lid9", "lid16", "lid14", "lid7", "lid1", "lid8_w", "lid15_w", "lid11_w", "lid0_w",
                 "lid8_b", "lid4_scales"):
        train["grad/" + name] = net.get(name, grad=True)
    net.update()
    for name in ("lid8_w", "lid15_w", "lid0_w", "lid8_b"):
        train["updated/" + name] = net.get(name)
    np.savez_compressed(cfg_dir / "mini_yolo_train.npz", **train)
    net.close()
    print("cfg:", sorted(p.name for p in cfg_dir.iterdir()))


def main():
    if sys.argv[1:] == ["cfg"]:
        return make_cfg()
    if sys.argv[1:] == ["model_io"]:
        return make_model_io()
    from bcnn_b200 import configs
    from helpers import rel_err
    for name in (sys.argv[1:] or netcases.CASES):
        net = ref_net(threads=1)
        out = netcases.run_case(net, name)
        net.close()
        # Conditioning of the case: the reference's OWN response to a 1-ulp (2e-7 relative)
        # perturbation of the input. Tiny-batch batch-norm nets amplify rounding noise by
        # orders of magnitude over a few SGD steps; the parity tolerance of a tensor can
        # not be tighter than that (tests/test_nets_gpu.py uses max(floor, 8 * sens)).
        clean = configs.synth_input
        configs.synth_input = lambda shape, seed=12345: (
            clean(shape, seed) * np.float32(1 + 2e-7)).astype(np.float32)
        try:
            net = ref_net(threads=1)
            pert = netcases.run_case(net, name)
            net.close()
        finally:
            configs.synth_input = clean
        packed = {}
        for k, v in out.items():
            if "/argmax/" not in k and np.abs(v).max(initial=0.0) > 0:
                packed["sens:" + k] = np.float32(max(rel_err(pert[k], v)))
            if name in FULL_CASES or v.size <= SUB:
                packed[k] = v
            else:
                packed[k + "@sub"] = subsample(v)
        path = GOLDEN / f"{name}.npz"
        np.savez_compressed(path, **packed)
        print(f"{path.name}: {len(packed)} arrays, {path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
