"""The multi-GPU path on real GPUs: images shard across one process per GPU, the ONLY collective is the NCCL
all-gather of the fixed-size pose records, and what every rank gathers is bit-identical to a single-GPU run over the same
image list.  Needs >= 2 visible GPUs (skipped otherwise; the host-side logic is covered on gloo by test_sharding_cpu.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import posenet  # noqa: E402
import sharding_worker as sw  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip("needs %d GPUs, %d visible" % (n, torch.cuda.device_count()))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gathered_records_equal_the_single_gpu_run(tmp_path, world):
    _need_gpus(world)
    n_total = 16
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharding_worker.py"), str(tmp_path), str(n_total)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, NCCL_DEBUG="WARN"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    g = np.load(tmp_path / "gathered.npz")
    assert int(g["world"]) == world
    # the same image list on ONE GPU, in this process
    dev = torch.device("cuda", 0)
    model = sw.build_model(dev)
    imgs = sw.images(n_total).to(dev)
    heads = model.forward_u8(imgs)
    ref = posenet.decode_multiple_poses_batch(*heads, output_stride=16, **sw.DECODE_KW)[:4]
    assert float(ref[0].max()) > 0                                      # the comparison is not vacuous: poses were found
    for j in range(4):
        one = ref[j].cpu().numpy()
        assert np.array_equal(g["s%d" % j], one), "infer_sharded: tensor %d differs from the single-GPU run" % j
        assert np.array_equal(g["p%d" % j], one), "BatchPipeline(gather=True): tensor %d differs from the single-GPU run" % j
    assert float(g["gather_ms"]) > 0


def test_second_device_after_the_first():
    """Per-device state (dynamic shared-memory attributes, SM counts, plans) must not leak from cuda:0 to cuda:1: the same
    model and images give bit-identical records on both, cuda:1 used AFTER cuda:0 in one process, without changing the
    current device."""
    _need_gpus(2)
    imgs = sw.images(4, 161, 129)
    out = []
    for d in (0, 1, 0):
        dev = torch.device("cuda", d)
        model = sw.build_model(dev)
        heads = model.forward_u8(imgs.to(dev))
        assert all(t.device == dev for t in heads)
        rec = posenet.decode_multiple_poses_batch(*heads, output_stride=16, **sw.DECODE_KW)
        x, _ = posenet.process_input_gpu(imgs.to(dev), 1.0, 16)
        heads32 = model.set_compute_dtype("fp32")(x)
        torch.cuda.synchronize(dev)
        out.append([t.cpu() for t in rec[:4]] + [t.cpu() for t in heads32])
    assert torch.cuda.current_device() == 0
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)
    for a, b in zip(out[0], out[2]):
        assert torch.equal(a, b)
