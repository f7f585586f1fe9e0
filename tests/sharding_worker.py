"""Worker of tests/test_gpu_sharding.py: launched once per GPU by torch.distributed.run with the NCCL backend.
Every rank builds the same seeded model and image list, runs its shard through ``posenet.sharding.infer_sharded`` and through
``posenet.BatchPipeline(gather=True)``; rank 0 writes what it gathered to ``<out_dir>/gathered.npz``."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200"), ROOT]
import posenet  # noqa: E402
from oracle import synth  # noqa: E402


def build_model(device):
    torch.manual_seed(11)
    return posenet.MobileNetV1(50, output_stride=16).to(device).set_compute_dtype("bf16")


def images(n, h=129, w=161):
    return torch.from_numpy(np.stack([synth.smooth_image(h, w, seed=300 + i) for i in range(n)]))


DECODE_KW = dict(max_pose_detections=6, min_pose_score=0.1)


def main():
    out_dir, n_total = sys.argv[1], int(sys.argv[2])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    try:
        model = build_model(dev)
        imgs = images(n_total)
        got = posenet.sharding.infer_sharded(model, imgs, **DECODE_KW)          # every rank ends up with all n_total records
        # the streaming front end: each rank submits its own shard, the step ends with the NCCL all-gather
        per = n_total // world
        assert per * world == n_total
        pipe = posenet.BatchPipeline(model, per, imgs.shape[1], imgs.shape[2], depth=2, gather=True, **DECODE_KW)
        mine = imgs[rank * per:(rank + 1) * per].contiguous().pin_memory()
        piped = [pipe.result(pipe.submit(mine), gathered=(rank == 0)) for _ in range(2)][-1]
        gather_ms = pipe.time_gather(reps=5)
        torch.cuda.synchronize()
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), world=world, gather_ms=gather_ms,
                     **{"s%d" % j: t.cpu().numpy() for j, t in enumerate(got)}, **{"p%d" % j: a for j, a in enumerate(piped)})
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
