import sys, time, os
sys.path[:0]=[os.path.join(os.getcwd(),'posenet-pytorch_b200'), os.getcwd()]
import numpy as np, torch, posenet
torch.manual_seed(0)
m = posenet.MobileNetV1(101, output_stride=16).cuda().set_compute_dtype("bf16")
rng = np.random.default_rng(1)
host = [torch.from_numpy(rng.integers(0,256,(64,513,513,3),dtype=np.uint8)).pin_memory() for _ in range(4)]
kw = dict(max_pose_detections=10, score_threshold=0.5, nms_radius=20, min_pose_score=0.25)
for depth in (1,2,3,4):
    pipe = posenet.BatchPipeline(m, 64, 513, 513, depth=depth, output_stride=16, **kw)
    for _ in pipe.run(host[i%4] for i in range(6)): pass
    torch.cuda.synchronize()
    best=0
    for rep in range(3):
        t0=time.perf_counter()
        for _ in pipe.run((host[i%4] for i in range(40)), copy=False): pass
        torch.cuda.synchronize()
        best=max(best, 40*64/(time.perf_counter()-t0))
    print("depth", depth, "e2e img/s %.0f"%best)
    del pipe
