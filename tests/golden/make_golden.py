"""Regenerate tests/golden/*.npz by running the UNMODIFIED reference CPU library
(oracle/_ref/libbcnn_ref.so, built from /root/reference by oracle/Makefile) through its own
public API on the seeded synthetic cases of tests/netcases.py.

    python tests/golden/make_golden.py [case ...]     (default: every case of netcases.CASES)

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
FULL_CASES = {"chain_b4", "resnet_small_b4", "yolo_two_heads_b2"}  # small enough to keep every tensor whole


def subsample(a: np.ndarray) -> np.ndarray:
    flat = a.ravel()
    if flat.size <= SUB:
        return flat.copy()
    idx = np.linspace(0, flat.size - 1, SUB).astype(np.int64)
    return flat[idx].copy()


def main():
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
