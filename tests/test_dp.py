"""Data-parallel path.

CPU (gloo, world_size 2): the algebra the B200 path relies on -- momentum kept in the gradient
buffer scaled by m / world so that a plain SUM all-reduce restores m * v_prev + sum_r g_r, the
global-batch divisor, disjoint contiguous shards -- is executed with the REFERENCE CPU library
as each rank's replica and torch.distributed(gloo) as the collective, and compared with the
reference itself training on the whole global batch in one process.

GPU (needs >= 2 devices, skipped otherwise): two B200 replicas with NCCL (bcnn_b200_dp_init)
against one replica on the concatenated batch.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from bcnn_b200 import capi, configs  # noqa: E402
from helpers import oracle, p, ref_available, ref_net, rel_err  # noqa: E402

WORLD, LOCAL_BATCH, STEPS = 2, 3, 3
LR, MOMENTUM, DECAY = 0.01, 0.9, 0.0005


def build_net(net, batch):
    """conv-pool-conv-fc chain WITHOUT batchnorm: sharded and whole-batch runs are then the
    same function of the data (per-replica BN statistics would differ by design)."""
    net.set_input_shape(12, 12, 3, batch)
    net.conv(8, 3, 1, 1, 1, 0, "relu", "input", "c1")
    net.maxpool(2, 2, capi.PAD_SAME, "c1", "p1")
    net.conv(12, 3, 1, 1, 1, 0, "lrelu", "p1", "c2")
    net.avgpool("c2", "gap")
    net.fullc(5, "none", "gap", "fc")
    net.softmax("fc", "softmax")
    net.cost("softmax", "cost")
    net.sgd(LR, MOMENTUM, DECAY)
    net.compile()
    configs.init_params(net, seed=11)


def global_data():
    x = configs.synth_input((WORLD * LOCAL_BATCH, 3, 12, 12), seed=5)
    y = configs.synth_labels((WORLD * LOCAL_BATCH, 5, 1, 1))
    return x, y


def shard(a, rank):
    return a[rank * LOCAL_BATCH:(rank + 1) * LOCAL_BATCH]


def params_of(net):
    return {name: net.get(idx) for idx, name, _ in configs.param_tensors(net)}


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    orc = oracle()
    net = ref_net()
    build_net(net, LOCAL_BATCH)
    x, y = global_data()
    net.set("input", shard(x, rank))
    net.set("label", shard(y, rank))
    plist = configs.param_tensors(net)
    for _ in range(STEPS):
        net.forward()
        net.backward()
        for idx, name, _shape in plist:
            t = net._tensor(idx)
            if not t.grad_data:
                continue
            g = torch.from_numpy(net.get(idx, grad=True))
            dist.all_reduce(g, op=dist.ReduceOp.SUM)          # the NCCL all-reduce of bcnn_dp.c
            w = net.get(idx).ravel().copy()
            gn = g.numpy().ravel().copy()
            # bcnn_b200_sgd_update: global batch, velocity left behind as (m / world) * v
            if name.endswith("_w"):
                orc.orc_sgd_update(p(w), None, p(gn), None, w.size, 0, WORLD * LOCAL_BATCH, LR,
                                   MOMENTUM / WORLD, DECAY)
            elif name.endswith("_b"):
                orc.orc_sgd_update(None, p(w), None, p(gn), 0, w.size, WORLD * LOCAL_BATCH, LR,
                                   MOMENTUM / WORLD, DECAY)
            else:
                continue
            net.set(idx, w)
            net.set(idx, gn, grad=True)
    np.savez(Path(out_dir) / f"rank{rank}.npz", **params_of(net))
    net.close()
    dist.destroy_process_group()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_sharded_replicas_match_whole_batch_reference_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(WORLD)]
    # whole global batch through the unmodified reference, its own bcnn_update
    net = ref_net()
    build_net(net, WORLD * LOCAL_BATCH)
    x, y = global_data()
    net.set("input", x)
    net.set("label", y)
    for _ in range(STEPS):
        net.forward()
        net.backward()
        net.update()
    want = params_of(net)
    net.close()
    for name, w in want.items():
        assert np.array_equal(ranks[0][name], ranks[1][name]), f"replicas diverged on {name}"
        e = max(rel_err(ranks[0][name], w))
        assert e <= 2e-5, f"{name}: sharded vs whole-batch rel err {e:.2e}"


# ----------------------------------------------------------------------------- GPU, 2 devices

def _nccl_worker(rank, port, out_dir):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    lib = capi.b200()
    lib.bcnn_b200_set_device(rank)
    net = capi.Net()
    build_net(net, LOCAL_BATCH)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_char * 128)()
        assert lib.bcnn_b200_dp_get_unique_id(raw) == 0
        uid.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    raw = (C.c_char * 128).from_buffer_copy(bytes(uid.cpu().numpy().tobytes()))
    assert lib.bcnn_b200_dp_init(net.handle, rank, WORLD, raw) == 0
    x, y = global_data()
    net.set("input", shard(x, rank))
    net.set("label", shard(y, rank))
    for _ in range(STEPS):
        net.train_step()
    net.sync()
    np.savez(Path(out_dir) / f"rank{rank}.npz", **params_of(net))
    net.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_replicas_match_single_gpu_whole_batch(tmp_path):
    lib = capi.b200()
    if lib.bcnn_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(WORLD)]
    net = capi.Net()
    build_net(net, WORLD * LOCAL_BATCH)
    x, y = global_data()
    net.set("input", x)
    net.set("label", y)
    for _ in range(STEPS):
        net.train_step()
    want = params_of(net)
    net.close()
    for name, w in want.items():
        assert np.array_equal(ranks[0][name], ranks[1][name]), f"replicas diverged on {name}"
        e = max(rel_err(ranks[0][name], w))
        assert e <= 2e-5, f"{name}: 2-GPU vs 1-GPU rel err {e:.2e}"
