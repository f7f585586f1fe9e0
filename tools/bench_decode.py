"""BASELINE configs[4] -- decode_multiple_poses stress: synthetic head tensors of 10-50 people at 1280x720 / OS8 (91x161
maps), max_pose_detections=50; decode-only latency on the GPU (candidates + greedy decode kernels, inputs resident,
CUDA events) next to the reference's algorithm on the host (the oracle port of decode_multi.py:61-148, one core --
it is a sequential numpy loop).  Prints one JSON line per (people, batch).

    python tools/bench_decode.py [--reps 50] [--skip-cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "posenet-pytorch_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

KW = dict(max_pose_detections=50, score_threshold=0.5, nms_radius=20, min_pose_score=0.25)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    import posenet
    from oracle import decode as odec, synth          # input generator + the CPU leg (checker / baseline only)
    torch.cuda.set_device(0)
    h, w, stride = 91, 161, 8
    for people in (10, 20, 30, 50):
        sets = [synth.people_heads(h, w, stride, people, seed=100 + s)[:4] for s in range(32)]
        cpu_ms = None
        if not args.skip_cpu:
            t0 = time.perf_counter()
            ref = [odec.decode_multiple_poses(*sets[s], stride, **KW) for s in range(4)]
            cpu_ms = (time.perf_counter() - t0) / 4 * 1e3
        for batch in (1, 32):
            dev = [torch.from_numpy(np.stack([sets[i][t] for i in range(batch)])).cuda() for t in range(4)]
            ws = {}
            out = posenet.decode_multiple_poses_batch(*dev, output_stride=stride, workspace=ws, **KW)
            torch.cuda.synchronize()
            found = float((out[0] > 0).sum()) / batch
            if not args.skip_cpu:                         # same answers as the CPU leg, bit for bit
                for i in range(min(batch, 4)):
                    for a, b in zip(out[:4], ref[i]):
                        assert np.array_equal(a[i].cpu().numpy(), b), "decode mismatch"
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    posenet.decode_multiple_poses_batch(*dev, output_stride=stride, workspace=ws, **KW)
            torch.cuda.synchronize()
            for _ in range(5):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            print(json.dumps({"metric": "decode-only latency", "workload": "configs[4]: %d people, 91x161 map, OS8, P=50" % people,
                              "batch": batch, "gpu_ms_per_batch": round(ms, 4), "gpu_us_per_image": round(ms / batch * 1e3, 2),
                              "poses_found_per_image": found, "cpu_ms_per_image": None if cpu_ms is None else round(cpu_ms, 2),
                              "cpu_kind": "port (numpy float64, 1 core)", "checked_bit_exact": not args.skip_cpu}), flush=True)


if __name__ == "__main__":
    main()
