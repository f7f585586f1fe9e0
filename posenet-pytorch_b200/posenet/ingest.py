"""Host image ingest in front of ``BatchPipeline``: decode image files on a thread pool straight into pinned
uint8 batches, so that file decoding, the host->device copy and the kernels of consecutive batches overlap.

The pixels are exactly what the reference feeds its network: ``cv2.imread`` (BGR uint8 HWC, ``posenet/utils.py:34-38``
``read_imgfile``); resizing / normalisation happen on the GPU (``pn_resize_u8`` + the stem), bit-exact with ``_process_input``.

    stream = posenet.ImageStream(paths, batch=64)                 # files sharing one frame size (camera / video frames)
    pipe = posenet.BatchPipeline(model, 64, stream.height, stream.width, min_pose_score=0.25)
    for (scores, kp_scores, kp_coords, offsets), n_valid in zip(pipe.run(b for b, _ in stream.batches()), stream.valid_counts()):
        ...

A directory of images of DIFFERENT sizes (``benchmark.py:24-29`` pre-processes every file at its own size):

    stream = posenet.ImageStream(paths, batch=64, height=720, width=1280, mixed=True)      # height x width: the largest frame
    pipe = posenet.BatchPipeline(model, 64, 720, 1280, mixed=True, source_coords=True, min_pose_score=0.25)
    for records in pipe.run((b, shapes) for b, _, shapes in stream.batches()):
        ...                                                       # coordinates already in each file's own pixel grid
"""
import concurrent.futures
import os

import cv2
import numpy as np
import torch


class ImageStream:
    def __init__(self, paths, batch, height=None, width=None, workers=None, slots=4, keep=2, pinned=None, mixed=False):
        """``mixed``: the files may have different sizes (each at most ``height`` x ``width``); a batch row then holds its frame
        contiguously at the frame's own size and ``batches()`` also yields the list of ``(h, w)`` -- the input format of a
        ``BatchPipeline(..., mixed=True)``, which resizes every frame on the GPU like ``_process_input`` (utils.py:13-26) does.

        ``slots`` batch buffers rotate.  A buffer handed out by ``batches()`` stays valid while the consumer takes ``keep``
        further batches and is decoded into again when it asks for the one after that; the remaining ``slots - 1 - keep``
        batches decode ahead on the thread pool.  ``BatchPipeline.run`` (depth d) copies batch i to the device asynchronously
        and only blocks on it after taking batch i + d, so ``keep >= d`` is what makes the reuse safe (defaults: d = 2)."""
        self.paths = list(paths)
        assert self.paths, "no image files"
        self.batch = int(batch)
        assert not (mixed and (height is None or width is None)), "mixed=True needs the largest frame size (height, width)"
        if height is None or width is None:
            first = cv2.imread(self.paths[0])
            if first is None:
                raise IOError("Image file not found or unreadable: %s" % self.paths[0])     # utils.py:36-37
            height, width = first.shape[:2]
        self.height, self.width = int(height), int(width)
        self.workers = workers or min(32, os.cpu_count() or 1)
        pinned = torch.cuda.is_available() if pinned is None else pinned
        self.mixed = bool(mixed)
        self.keep = int(keep)
        assert int(slots) >= self.keep + 2, "slots must be at least keep + 2 (one buffer in use, one decoding ahead)"
        self._bufs = []
        for _ in range(int(slots)):
            t = torch.zeros((self.batch, self.height, self.width, 3), dtype=torch.uint8)
            self._bufs.append(t.pin_memory() if pinned else t)
        self._views = [b.numpy() for b in self._bufs]
        self._shapes = [[None] * self.batch for _ in self._bufs]

    def __len__(self):
        return (len(self.paths) + self.batch - 1) // self.batch

    def valid_counts(self):
        """Number of real images in every batch (the last one may be partial; its tail rows are zero images)."""
        n = len(self.paths)
        return [min(self.batch, n - i) for i in range(0, n, self.batch)]

    def _load(self, path, dst, shapes, row):
        img = cv2.imread(path)
        if img is None:
            raise IOError("Image file not found or unreadable: %s" % path)
        if self.mixed:
            if img.shape[0] > self.height or img.shape[1] > self.width:
                raise ValueError("%s is %dx%d, larger than the stream's %dx%d frames" % (path, img.shape[1], img.shape[0], self.width, self.height))
            dst.reshape(-1)[:img.size] = img.reshape(-1)             # the frame at its own size, contiguous at the start of the row
            shapes[row] = (img.shape[0], img.shape[1])
            return
        if img.shape != dst.shape:
            raise ValueError("%s is %dx%d, the stream carries %dx%d frames (mixed=True accepts different sizes)" % (
                path, img.shape[1], img.shape[0], self.width, self.height))
        dst[...] = img

    def batches(self):
        """Yields ``(uint8 tensor [batch, h, w, 3] (pinned when CUDA is available), n_valid)`` -- ``(tensor, n_valid, shapes)``
        for a mixed stream; decoding of the next ``slots - 1 - keep`` batches runs ahead on the thread pool."""
        n = len(self.paths)
        starts = list(range(0, n, self.batch))
        with concurrent.futures.ThreadPoolExecutor(max_workers=self.workers) as ex:
            def launch(bi):
                buf, shapes = self._views[bi % len(self._views)], self._shapes[bi % len(self._views)]
                lo = starts[bi]
                hi = min(n, lo + self.batch)
                if hi - lo < self.batch:
                    buf[hi - lo:] = 0
                return [ex.submit(self._load, self.paths[i], buf[i - lo], shapes, i - lo) for i in range(lo, hi)]
            ahead = len(self._bufs) - 1 - self.keep
            pending = {bi: launch(bi) for bi in range(min(ahead, len(starts)))}
            for bi in range(len(starts)):
                for f in pending.pop(bi):
                    f.result()                                   # re-raises decode errors
                nxt = bi + ahead
                # batch `nxt` decodes into the buffer handed out `keep + 1` batches ago: the consumer asking for batch `bi` is what
                # releases it, so its decode starts BEFORE the yield and overlaps the consumer's work on batch `bi`
                if nxt < len(starts):
                    pending[nxt] = launch(nxt)
                nv = min(self.batch, n - starts[bi])
                if self.mixed:
                    yield self._bufs[bi % len(self._bufs)], nv, list(self._shapes[bi % len(self._bufs)][:nv])
                else:
                    yield self._bufs[bi % len(self._bufs)], nv
