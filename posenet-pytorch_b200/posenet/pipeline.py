"""Streaming front end of the hot path: host uint8 batches in, host pose records out, with the
host->device copy of batch i+1 and the device->host copy of batch i-1 overlapping the kernels of
batch i (two copy engines + one compute stream, one CUDA graph per in-flight slot).

This is the batched form of what the reference's ``benchmark.py:32-44`` does per image
(``model(input)`` then ``decode_multiple_poses``): same results, no per-image synchronisation.

    pipe = posenet.BatchPipeline(model, batch=64, height=513, width=513, min_pose_score=0.25)
    for pose_scores, keypoint_scores, keypoint_coords, pose_offsets in pipe.run(batches):   # pinned uint8 [64,513,513,3]
        ...

Options on top of that: ``source_coords=True`` maps the keypoint coordinates back to the submitted frames ON THE DEVICE
(``image_demo.py:50``), ``mixed=True`` accepts frames of different sizes in one batch (every file of a directory
pre-processed at its own size, ``benchmark.py:24-29``), ``gather=True`` all-gathers the records of every rank's shard over
NCCL inside the step (one process per GPU, ``posenet.sharding``).
"""
import ctypes as C

import numpy as np
import torch

from posenet import _native as nat
from posenet.constants import NUM_KEYPOINTS
from posenet.decode_multi import decode_multiple_poses_batch, split_pose_records


class BatchPipeline:
    def __init__(self, model, batch, height, width, depth=2, use_graph=True, output_stride=None, scale_factor=1.0,
                 source_coords=False, mixed=False, gather=False, group=None, **decode_kw):
        """``height`` x ``width``: size of the uint8 frames handed to ``submit`` (``mixed=True``: the LARGEST frame; each
        frame of a batch then comes with its own size).  The network runs at ``valid_resolution(width * scale_factor,
        height * scale_factor)`` (utils.py:7-10); frames of any other size go through the bit-exact cv2 resize
        (``pn_resize_u8``) on the GPU in front of the stem -- inside the step's CUDA graph when all frames share one size.

        ``source_coords``: ``keypoint_coords *= (src_h / target_h, src_w / target_w)`` (image_demo.py:50, utils.py:19) applied to
        the record buffer on the device (``pn_scale_keypoint_coords``), per frame in mixed mode.
        ``gather``: with an initialised process group, every step ends with an all-gather of the ranks' record buffers (NCCL over
        NVLink) on its own stream -- the kernels of the next batch do not wait for it -- and rank 0 reads ALL ranks' records back
        (``result(..., gathered=True)``); the other ranks read back their own, as without ``gather``."""
        nat.require_device()
        assert depth >= 1
        self.model, self.batch, self.h, self.w = model, int(batch), int(height), int(width)
        from posenet.utils import valid_resolution
        self.os = output_stride or model.output_stride
        self.tw, self.th = valid_resolution(self.w * scale_factor, self.h * scale_factor, output_stride=self.os)
        self.mixed = bool(mixed)
        self.resize = self.mixed or (self.th, self.tw) != (self.h, self.w)
        self.scale = np.array([self.h / self.th, self.w / self.tw])           # utils.py:19, for mapping coordinates back
        self.source_coords = bool(source_coords)
        self.decode_kw = dict(decode_kw)
        self.P = int(self.decode_kw.get("max_pose_detections", 10))
        dev = model._device()
        self.dev = dev
        self.world, self.rank, self.group = 1, 0, group
        if gather:
            import torch.distributed as dist
            assert dist.is_available() and dist.is_initialized(), "gather=True needs an initialised process group"
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.gather = bool(gather) and self.world > 1
        with torch.cuda.device(dev):
            self.h2d, self.d2h, self.compute = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.collect_all = self.gather and self.rank == 0      # this rank's D2H copy carries every rank's records
            self.nrec = self.batch * self.P * (1 + 5 * NUM_KEYPOINTS)
            nrec = self.nrec
            self.slots = []
            self._ws = {}
            for _ in range(depth):
                s = dict(x=torch.empty((self.batch, self.h, self.w, 3), dtype=torch.uint8, device=dev),
                         xr=torch.empty((self.batch, self.th, self.tw, 3), dtype=torch.uint8, device=dev) if self.resize else None,
                         rec=torch.zeros(nrec, dtype=torch.float64, device=dev),
                         rec_all=torch.zeros(self.world * nrec, dtype=torch.float64, device=dev) if self.gather else None,
                         rec_host=torch.zeros((self.world if self.collect_all else 1) * nrec, dtype=torch.float64).pin_memory(),
                         scales=torch.ones((self.batch, 2), dtype=torch.float64, device=dev) if self.mixed else None,
                         scales_host=torch.ones((self.batch, 2), dtype=torch.float64).pin_memory() if self.mixed else None,
                         copied=torch.cuda.Event(), done=torch.cuda.Event(), out=torch.cuda.Event(), graph=None, busy=False)
                self.slots.append(s)
            self.h2d_bytes_per_batch = self.slots[0]["x"].numel()
            self.d2h_bytes_per_batch = (self.world if self.collect_all else 1) * nrec * 8
            # warm up (plans, workspaces) and capture one graph per slot on the compute stream
            self.compute.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.compute):
                for s in self.slots:
                    s["x"].zero_()
                    if self.mixed:
                        s["xr"].zero_()
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
        """The captured part of a step: (uniform resize ->) stem .. heads -> candidates -> greedy decode (-> coordinate scaling)."""
        x = s["x"]
        lib = nat.load()
        if self.resize and not self.mixed:
            nat.check(lib.pn_resize_u8(x.data_ptr(), self.batch, self.h, self.w, self.th, self.tw, s["xr"].data_ptr(),
                                       nat.stream_ptr()), "pn_resize_u8")
        if self.resize:
            x = s["xr"]                                              # mixed: filled per frame by submit(), outside the graph
        heads = self.model.forward_u8(x)
        ps, ks, kc, ko, _ = decode_multiple_poses_batch(*heads, output_stride=self.os, workspace=self._ws, out=s["rec"], **self.decode_kw)
        if self.source_coords and (self.mixed or self.resize):
            nat.check(lib.pn_scale_keypoint_coords(C.c_void_p(kc.data_ptr()), self.batch, self.P * NUM_KEYPOINTS,
                                                   C.c_void_p(s["scales"].data_ptr()) if self.mixed else None,
                                                   float(self.scale[0]), float(self.scale[1]), nat.stream_ptr()),
                      "pn_scale_keypoint_coords")

    def _resize_mixed(self, s, shapes):
        """Per-frame cv2-exact resize of frames stored at their own size (row i of the slot's input buffer holds frame i
        contiguously: h_i * w_i * 3 bytes) into the network-sized batch; runs of equal-sized frames share a launch when
        they are also full-sized (then the rows are dense)."""
        lib, st = nat.load(), nat.stream_ptr()
        row = self.h * self.w * 3
        base_in, base_out = s["x"].data_ptr(), s["xr"].data_ptr()
        out_row = self.th * self.tw * 3
        i = 0
        while i < len(shapes):
            h, w = shapes[i]
            n = 1
            if (h, w) == (self.h, self.w):                           # dense rows: take the whole run at once
                while i + n < len(shapes) and tuple(shapes[i + n]) == (h, w):
                    n += 1
            nat.check(lib.pn_resize_u8(base_in + i * row, n, h, w, self.th, self.tw, base_out + i * out_row, st), "pn_resize_u8")
            i += n

    def submit(self, host_batch, shapes=None):
        """Enqueue one batch (uint8 [batch,h,w,3], ideally pinned).  Returns a ticket for ``result``; if the slot is
        still in flight its previous result must have been collected.

        ``mixed=True``: ``host_batch`` is uint8 ``[batch, h*w*3]`` (or ``[batch,h,w,3]``) whose row i starts with frame i
        stored contiguously at its own size ``shapes[i] = (h_i, w_i)`` (``h_i <= h``, ``w_i <= w``); rows without a frame
        (``len(shapes) < batch``) decode as black frames."""
        s = self.slots[self._next]
        assert not s["busy"], "pipeline full: collect result() of the oldest ticket first"
        assert host_batch.numel() == s["x"].numel() and host_batch.dtype == torch.uint8
        if self.mixed:
            assert shapes is not None and len(shapes) <= self.batch, "mixed=True: pass the (h, w) of every frame"
            shapes = [(int(h), int(w)) for h, w in shapes]
            assert all(0 < h <= self.h and 0 < w <= self.w for h, w in shapes), "a frame exceeds the pipeline's %dx%d" % (self.h, self.w)
            sc = s["scales_host"]
            sc.fill_(1.0)
            for i, (h, w) in enumerate(shapes):
                sc[i, 0], sc[i, 1] = h / self.th, w / self.tw                     # utils.py:19, per frame
        else:
            assert shapes is None, "shapes are for mixed=True pipelines"
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.h2d):
                self.h2d.wait_event(s["done"])              # the previous use of this slot's input buffer has finished
                if self.mixed and shapes:
                    used = max(h * w * 3 for h, w in shapes)                      # one strided copy of the bytes in use
                    s["x"].view(self.batch, -1)[:len(shapes), :used].copy_(host_batch.view(self.batch, -1)[:len(shapes), :used],
                                                                          non_blocking=True)
                    s["scales"].copy_(sc, non_blocking=True)
                elif not self.mixed:
                    s["x"].copy_(host_batch.view(s["x"].shape), non_blocking=True)
                s["copied"].record(self.h2d)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(s["copied"])
                self.compute.wait_event(s["out"])           # the previous records of this slot have left the device
                if self.mixed:
                    if len(shapes) < self.batch:
                        s["xr"][len(shapes):].zero_()
                    self._resize_mixed(s, shapes)
                if s["graph"] is not None:
                    s["graph"].replay()
                else:
                    self._enqueue(s)
                s["done"].record(self.compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(s["done"])
                if self.gather:                             # off the compute stream: the next batch's kernels do not wait for the peers
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(s["rec_all"], s["rec"], group=self.group)
                s["rec_host"].copy_(s["rec_all"] if self.collect_all else s["rec"], non_blocking=True)
                s["out"].record(self.d2h)
        s["busy"] = True
        ticket = self._next
        self._next = (self._next + 1) % len(self.slots)
        return ticket

    def result(self, ticket, copy=True, source_coords=None, gathered=False):
        """Block until the ticket's records are on the host; returns the reference's 4-tuple for this rank's batch
        (numpy float64: [batch,P], [batch,P,17], [batch,P,17,2], [batch,P,17,2]).  ``gathered=True`` (rank 0 of a
        ``gather=True`` pipeline): the records of ALL ranks' batches instead, leading dimension ``world * batch`` (rank-major).
        ``source_coords`` is a construction-time option (the scaling runs on the device); passing a different value here is an
        error."""
        assert source_coords is None or bool(source_coords) == self.source_coords, \
            "source_coords is fixed when the pipeline is built (the scaling runs on the device, before the D2H copy)"
        s = self.slots[ticket]
        assert s["busy"], "no batch in flight for this ticket"
        s["out"].synchronize()
        s["busy"] = False
        flat = s["rec_host"].numpy()
        if copy:
            flat = flat.copy()
        if not gathered:
            own = flat[self.rank * self.nrec:(self.rank + 1) * self.nrec] if self.collect_all else flat
            return split_pose_records(own, self.batch, self.P)
        assert self.collect_all, "gathered=True: only rank 0 of a gather=True pipeline reads every rank's records back"
        parts = [split_pose_records(flat[r * self.nrec:(r + 1) * self.nrec], self.batch, self.P) for r in range(self.world)]
        return tuple(np.concatenate([p[j] for p in parts], axis=0) for j in range(4))

    def time_gather(self, reps=20):
        """Device time (ms) of one all-gather of a step's record buffer, measured alone with CUDA events."""
        import torch.distributed as dist
        assert self.gather
        s = self.slots[0]
        with torch.cuda.device(self.dev), torch.cuda.stream(self.d2h):
            for _ in range(3):
                dist.all_gather_into_tensor(s["rec_all"], s["rec"], group=self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.d2h)
            for _ in range(reps):
                dist.all_gather_into_tensor(s["rec_all"], s["rec"], group=self.group)
            e1.record(self.d2h)
            self.d2h.synchronize()
        return e0.elapsed_time(e1) / reps

    def run(self, batches, copy=True, source_coords=None):
        """Generator: yields the pose records of every batch of ``batches`` in order, keeping ``depth`` batches in flight.
        Items are uint8 batches, or ``(batch, shapes)`` pairs for a ``mixed=True`` pipeline."""
        pending = []
        for hb in batches:
            if len(pending) == len(self.slots):
                yield self.result(pending.pop(0), copy=copy, source_coords=source_coords)
            pending.append(self.submit(*hb) if isinstance(hb, (tuple, list)) else self.submit(hb))
        while pending:
            yield self.result(pending.pop(0), copy=copy, source_coords=source_coords)
