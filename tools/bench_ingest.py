"""Host ingest rate: ImageStream alone and ImageStream -> BatchPipeline at C2 shapes, on synthetic JPEG files."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "posenet-pytorch_b200"), ROOT):
    sys.path.insert(0, p)
import cv2
import numpy as np
import torch
import posenet
from oracle import synth            # input generator only

n_files, batch, H, W = 512, 64, 513, 513
tmp = tempfile.mkdtemp()
base = [synth.smooth_image(H, W, s) for s in range(16)]
paths = []
for i in range(n_files):
    p = os.path.join(tmp, "im%04d.jpg" % i)
    cv2.imwrite(p, np.roll(base[i % 16], i, axis=1), [cv2.IMWRITE_JPEG_QUALITY, 90])
    paths.append(p)
stream = posenet.ImageStream(paths, batch=batch)
t0 = time.perf_counter()
n = sum(nv for _, nv in stream.batches())
dt_decode = time.perf_counter() - t0
torch.cuda.set_device(0)
torch.manual_seed(0)
model = posenet.MobileNetV1(101, output_stride=16).cuda().set_compute_dtype("bf16")
pipe = posenet.BatchPipeline(model, batch, H, W, depth=2, max_pose_detections=10, min_pose_score=0.25)
list(pipe.run(b for b, _ in posenet.ImageStream(paths[:128], batch=batch).batches()))
t0 = time.perf_counter()
m = sum(1 for _ in pipe.run((b for b, _ in stream.batches()), copy=False))
torch.cuda.synchronize()
dt_e2e = time.perf_counter() - t0
print(json.dumps({"metric": "host ingest", "files": n_files, "frame": [H, W], "jpeg_quality": 90, "workers": stream.workers,
                  "host_cores": os.cpu_count(), "decode_only_images_per_s": round(n / dt_decode, 1),
                  "files_to_pose_records_images_per_s": round(m * batch / dt_e2e, 1)}))
