"""Phase breakdown of decode_kernel (PN_DEC_TRACE build of decode.cu) on the bench workloads' head tensors.
Run on the GPU box with POSENET_B200_LIB=posenet-pytorch_b200/lib/libposenet_b200_trace.so (built by hand:
nvcc ... -DPN_DEC_TRACE -c csrc/decode.cu, linked with the other objects)."""
import os, sys
sys.path[:0] = [os.path.join(os.getcwd(), "posenet-pytorch_b200"), os.getcwd()]
import numpy as np, torch, posenet
KW = dict(max_pose_detections=10, score_threshold=0.5, nms_radius=20, min_pose_score=0.25)
for name, mid, h, w, os_, batch in (("c2", 101, 513, 513, 16, 64), ("c3", 50, 721, 1281, 8, 32), ("c4", 75, 257, 257, 32, 512)):
    torch.manual_seed(0)
    m = posenet.MobileNetV1(mid, output_stride=os_).cuda().set_compute_dtype("bf16")
    x = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (batch, h, w, 3), dtype=np.uint8)).cuda()
    heads = m.forward_u8(x)
    torch.cuda.synchronize()
    print("==", name, flush=True)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = posenet.decode_multiple_poses_batch(*heads, output_stride=os_, **KW)
        e1.record()
        torch.cuda.synchronize()
        print("rep", rep, "candidates+decode ms %.3f" % e0.elapsed_time(e1), "counts", r[4][:8].tolist(), flush=True)
