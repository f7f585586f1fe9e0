"""Host image ingest in front of ``BatchPipeline``: decode image files on a thread pool straight into pinned
uint8 batches, so that file decoding, the host->device copy and the kernels of consecutive batches overlap.

The pixels are exactly what the reference feeds its network: ``cv2.imread`` (BGR uint8 HWC, ``posenet/utils.py:34-38``
``read_imgfile``); resizing / normalisation happen on the GPU (``pn_resize_u8`` + the stem), bit-exact with ``_process_input``.

    stream = posenet.ImageStream(paths, batch=64)                 # all files must share one frame size (camera / video frames)
    pipe = posenet.BatchPipeline(model, 64, stream.height, stream.width, min_pose_score=0.25)
    for (scores, kp_scores, kp_coords, offsets), n_valid in zip(pipe.run(b for b, _ in stream.batches()), stream.valid_counts()):
        ...
"""
import concurrent.futures
import os

import cv2
import numpy as np
import torch


class ImageStream:
    def __init__(self, paths, batch, height=None, width=None, workers=None, slots=4, keep=2, pinned=None):
        """``slots`` batch buffers rotate.  A buffer handed out by ``batches()`` stays valid while the consumer takes ``keep``
        further batches and is decoded into again when it asks for the one after that; the remaining ``slots - 1 - keep``
        batches decode ahead on the thread pool.  ``BatchPipeline.run`` (depth d) copies batch i to the device asynchronously
        and only blocks on it after taking batch i + d, so ``keep >= d`` is what makes the reuse safe (defaults: d = 2)."""
        self.paths = list(paths)
        assert self.paths, "no image files"
        self.batch = int(batch)
        if height is None or width is None:
            first = cv2.imread(self.paths[0])
            if first is None:
                raise IOError("Image file not found or unreadable: %s" % self.paths[0])     # utils.py:36-37
            height, width = first.shape[:2]
        self.height, self.width = int(height), int(width)
        self.workers = workers or min(32, os.cpu_count() or 1)
        pinned = torch.cuda.is_available() if pinned is None else pinned
        self.keep = int(keep)
        assert int(slots) >= self.keep + 2, "slots must be at least keep + 2 (one buffer in use, one decoding ahead)"
        self._bufs = []
        for _ in range(int(slots)):
            t = torch.zeros((self.batch, self.height, self.width, 3), dtype=torch.uint8)
            self._bufs.append(t.pin_memory() if pinned else t)
        self._views = [b.numpy() for b in self._bufs]

    def __len__(self):
        return (len(self.paths) + self.batch - 1) // self.batch

    def valid_counts(self):
        """Number of real images in every batch (the last one may be partial; its tail rows are zero images)."""
        n = len(self.paths)
        return [min(self.batch, n - i) for i in range(0, n, self.batch)]

    def _load(self, path, dst):
        img = cv2.imread(path)
        if img is None:
            raise IOError("Image file not found or unreadable: %s" % path)
        if img.shape != dst.shape:
            raise ValueError("%s is %dx%d, the stream carries %dx%d frames" % (path, img.shape[1], img.shape[0], self.width, self.height))
        dst[...] = img

    def batches(self):
        """Yields ``(uint8 tensor [batch, h, w, 3] (pinned when CUDA is available), n_valid)``; decoding of the next
        ``slots - 1 - keep`` batches runs ahead on the thread pool."""
        n = len(self.paths)
        starts = list(range(0, n, self.batch))
        with concurrent.futures.ThreadPoolExecutor(max_workers=self.workers) as ex:
            def launch(bi):
                buf = self._views[bi % len(self._views)]
                lo = starts[bi]
                hi = min(n, lo + self.batch)
                if hi - lo < self.batch:
                    buf[hi - lo:] = 0
                return [ex.submit(self._load, self.paths[i], buf[i - lo]) for i in range(lo, hi)]
            ahead = len(self._bufs) - 1 - self.keep
            pending = {bi: launch(bi) for bi in range(min(ahead, len(starts)))}
            for bi in range(len(starts)):
                for f in pending.pop(bi):
                    f.result()                                   # re-raises decode errors
                nxt = bi + ahead
                # batch `nxt` decodes into the buffer handed out `keep + 1` batches ago, now that the consumer is back for the next one
                yield self._bufs[bi % len(self._bufs)], min(self.batch, n - starts[bi])
                if nxt < len(starts):
                    pending[nxt] = launch(nxt)
