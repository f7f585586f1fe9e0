"""Streaming front end of the hot path: host uint8 batches in, host pose records out, with the
host->device copy of batch i+1 and the device->host copy of batch i-1 overlapping the kernels of
batch i (two copy engines + one compute stream, one CUDA graph per in-flight slot).

This is the batched form of what the reference's ``benchmark.py:32-44`` does per image
(``model(input)`` then ``decode_multiple_poses``): same results, no per-image synchronisation.

    pipe = posenet.BatchPipeline(model, batch=64, height=513, width=513, min_pose_score=0.25)
    for pose_scores, keypoint_scores, keypoint_coords, pose_offsets in pipe.run(batches):   # pinned uint8 [64,513,513,3]
        ...
"""
import numpy as np
import torch

from posenet import _native as nat
from posenet.constants import NUM_KEYPOINTS
from posenet.decode_multi import decode_multiple_poses_batch, split_pose_records


class BatchPipeline:
    def __init__(self, model, batch, height, width, depth=2, use_graph=True, output_stride=None, scale_factor=1.0, **decode_kw):
        """``height`` x ``width``: size of the uint8 frames handed to ``submit``.  The network runs at
        ``valid_resolution(width * scale_factor, height * scale_factor)`` (utils.py:7-10); when that differs from the frame size
        the bit-exact cv2 resize (``pn_resize_u8``) runs on the GPU in front of the stem, inside the same graph."""
        nat.require_device()
        assert depth >= 1
        self.model, self.batch, self.h, self.w = model, int(batch), int(height), int(width)
        from posenet.utils import valid_resolution
        self.os = output_stride or model.output_stride
        self.tw, self.th = valid_resolution(self.w * scale_factor, self.h * scale_factor, output_stride=self.os)
        self.resize = (self.th, self.tw) != (self.h, self.w)
        self.scale = np.array([self.h / self.th, self.w / self.tw])           # utils.py:19, for mapping coordinates back
        self.decode_kw = dict(decode_kw)
        self.P = int(self.decode_kw.get("max_pose_detections", 10))
        dev = model._device()
        self.dev = dev
        self.h2d, self.d2h, self.compute = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        nrec = self.batch * self.P * (1 + 5 * NUM_KEYPOINTS)
        self.slots = []
        self._ws = {}
        for _ in range(depth):
            s = dict(x=torch.empty((self.batch, self.h, self.w, 3), dtype=torch.uint8, device=dev),
                     xr=torch.empty((self.batch, self.th, self.tw, 3), dtype=torch.uint8, device=dev) if self.resize else None,
                     rec=torch.zeros(nrec, dtype=torch.float64, device=dev),
                     rec_host=torch.zeros(nrec, dtype=torch.float64).pin_memory(),
                     copied=torch.cuda.Event(), done=torch.cuda.Event(), out=torch.cuda.Event(), graph=None, busy=False)
            self.slots.append(s)
        self.h2d_bytes_per_batch = self.slots[0]["x"].numel()
        self.d2h_bytes_per_batch = nrec * 8
        # warm up (plans, workspaces) and capture one graph per slot on the compute stream
        self.compute.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.compute):
            for s in self.slots:
                s["x"].zero_()
                self._enqueue(s)
            self.compute.synchronize()
            if use_graph:
                for s in self.slots:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.compute):
                        self._enqueue(s)
                    s["graph"] = g
        self.compute.synchronize()
        self._next = 0

    def _enqueue(self, s):
        x = s["x"]
        if self.resize:
            nat.check(nat.load().pn_resize_u8(x.data_ptr(), self.batch, self.h, self.w, self.th, self.tw, s["xr"].data_ptr(),
                                              nat.stream_ptr()), "pn_resize_u8")
            x = s["xr"]
        heads = self.model.forward_u8(x)
        decode_multiple_poses_batch(*heads, output_stride=self.os, workspace=self._ws, out=s["rec"], **self.decode_kw)

    def submit(self, host_batch):
        """Enqueue one batch (uint8 [batch,h,w,3], ideally pinned).  Returns a ticket for ``result``; if the slot is
        still in flight its previous result must have been collected."""
        s = self.slots[self._next]
        assert not s["busy"], "pipeline full: collect result() of the oldest ticket first"
        assert tuple(host_batch.shape) == tuple(s["x"].shape) and host_batch.dtype == torch.uint8
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(s["done"])              # the previous use of this slot's input buffer has finished
            s["x"].copy_(host_batch, non_blocking=True)
            s["copied"].record(self.h2d)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(s["copied"])
            self.compute.wait_event(s["out"])           # the previous records of this slot have left the device
            if s["graph"] is not None:
                s["graph"].replay()
            else:
                self._enqueue(s)
            s["done"].record(self.compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s["done"])
            s["rec_host"].copy_(s["rec"], non_blocking=True)
            s["out"].record(self.d2h)
        s["busy"] = True
        ticket = self._next
        self._next = (self._next + 1) % len(self.slots)
        return ticket

    def result(self, ticket, copy=True, source_coords=False):
        """Block until the ticket's records are on the host; returns the reference's 4-tuple for the whole batch
        (numpy float64: [batch,P], [batch,P,17], [batch,P,17,2], [batch,P,17,2]).  ``source_coords``: keypoint coordinates
        mapped back to the submitted frames, ``keypoint_coords *= output_scale`` of image_demo.py:50 for the whole batch
        (implies a copy)."""
        s = self.slots[ticket]
        assert s["busy"], "no batch in flight for this ticket"
        s["out"].synchronize()
        s["busy"] = False
        flat = s["rec_host"].numpy()
        if copy or source_coords:
            flat = flat.copy()
        ps, ks, kc, ko = split_pose_records(flat, self.batch, self.P)
        if source_coords:
            kc *= self.scale                                     # (y, x) * (src_h / target_h, src_w / target_w), utils.py:19
        return ps, ks, kc, ko

    def run(self, batches, copy=True, source_coords=False):
        """Generator: yields the pose records of every batch of ``batches`` in order, keeping ``depth`` batches in flight."""
        pending = []
        for hb in batches:
            if len(pending) == len(self.slots):
                yield self.result(pending.pop(0), copy=copy, source_coords=source_coords)
            pending.append(self.submit(hb))
        while pending:
            yield self.result(pending.pop(0), copy=copy, source_coords=source_coords)
