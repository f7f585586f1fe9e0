"""SURVEY 8(f) N2 -- host image ingest: files decoded on a thread pool into (pinned) uint8 batches hold exactly the pixels the
reference's read_imgfile (utils.py:34-38: cv2.imread) would feed its preprocessing."""
import os

import cv2
import numpy as np
import pytest

import posenet


def _write(tmp_path, n, h, w, ext):
    rng = np.random.default_rng(3)
    paths = []
    for i in range(n):
        p = str(tmp_path / ("img%03d.%s" % (i, ext)))
        assert cv2.imwrite(p, rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        paths.append(p)
    return paths


@pytest.mark.parametrize("ext", ["png", "jpg"])
def test_batches_equal_cv2_imread(tmp_path, ext):
    paths = _write(tmp_path, 22, 37, 53, ext)
    stream = posenet.ImageStream(paths, batch=4, workers=3, slots=4, keep=2)
    assert (stream.height, stream.width, len(stream)) == (37, 53, 6) and stream.valid_counts() == [4, 4, 4, 4, 4, 2]
    seen, held = 0, []
    for bi, (buf, n_valid) in enumerate(stream.batches()):
        assert tuple(buf.shape) == (4, 37, 53, 3) and n_valid == stream.valid_counts()[bi]
        held.append((buf, seen, n_valid))
        for b, first, nv in held[-3:]:                            # this batch and the `keep` = 2 before it are still intact
            got = b.numpy()
            for j in range(nv):
                assert np.array_equal(got[j], cv2.imread(paths[first + j]))
            assert not got[nv:].any()                             # the tail of a partial batch is zero images
        seen += n_valid
    assert seen == 22


def test_errors(tmp_path):
    paths = _write(tmp_path, 3, 20, 20, "png")
    with pytest.raises(IOError):
        posenet.ImageStream([str(tmp_path / "missing.png")], batch=2)
    odd = str(tmp_path / "odd.png")
    cv2.imwrite(odd, np.zeros((21, 20, 3), np.uint8))
    with pytest.raises(ValueError):
        list(posenet.ImageStream(paths + [odd], batch=2).batches())
    with pytest.raises(IOError):
        list(posenet.ImageStream(paths + [str(tmp_path / "gone.png")], batch=2, height=20, width=20).batches())


def test_mixed_sizes_and_prefetch(tmp_path):
    """mixed=True: files of different sizes land at their own size at the start of their batch row, with their shapes; the
    decode of the batch after next is already running when a batch is handed out (launch before the yield)."""
    import cv2
    import numpy as np
    import posenet
    sizes = [(40, 60), (64, 64), (30, 20), (64, 50), (10, 64)]
    paths, imgs = [], []
    rng = np.random.default_rng(0)
    for i, (h, w) in enumerate(sizes):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = str(tmp_path / ("m%d.png" % i))
        assert cv2.imwrite(p, img)
        paths.append(p); imgs.append(img)
    st = posenet.ImageStream(paths, batch=2, height=64, width=64, mixed=True, pinned=False)
    k = 0
    for buf, nv, shapes in st.batches():
        assert tuple(buf.shape) == (2, 64, 64, 3) and len(shapes) == nv
        flat = buf.numpy().reshape(2, -1)
        for j in range(nv):
            assert tuple(shapes[j]) == sizes[k]
            assert np.array_equal(flat[j, :imgs[k].size], imgs[k].reshape(-1))
            k += 1
    assert k == len(paths)
    with pytest.raises(ValueError):
        list(posenet.ImageStream(paths, batch=2, height=32, width=64, mixed=True, pinned=False).batches())
