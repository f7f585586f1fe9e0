"""The reference's own caller scripts, UNCHANGED, against the product package on the GPU (north_star: "image_demo.py and
benchmark.py run unchanged").

``oracle/make_ref.py`` byte-compiles ``/root/reference/benchmark.py`` (:16-46) and ``/root/reference/image_demo.py`` (:20-69)
-- sha256-pinned sources -- into ``oracle/_ref/scripts.zip``; here they are unpacked and run as subprocesses with ``PYTHONPATH`` pointing at
``posenet-pytorch_b200`` so that their ``import posenet`` is the product, on a directory of synthetic images of DIFFERENT
sizes (benchmark.py:24-29 / image_demo.py:33-35 pre-process every file at its own size) and a random-init checkpoint written
where ``load_model`` looks for it (``./_models``)."""
import os
import re
import subprocess
import sys

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import posenet  # noqa: E402
from oracle import make_ref, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "posenet-pytorch_b200")
SIZES = [(513, 513, "jpg"), (300, 400, "png"), (720, 1280, "jpg"), (257, 350, "jpg"), (129, 97, "png")]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not make_ref.available():
        pytest.skip("oracle/_ref is not built (python oracle/make_ref.py needs /root/reference)")
    m = make_ref.manifest()
    assert m and m["scripts_sha256"] == make_ref.SCRIPT_SHA256            # the bytecode is that of the pinned, unmodified sources
    d = tmp_path_factory.mktemp("ref_scripts")
    os.makedirs(d / "images")
    for i, (h, w, ext) in enumerate(SIZES):
        assert cv2.imwrite(str(d / "images" / ("img%d.%s" % (i, ext))), synth.smooth_image(h, w, seed=40 + i))
    for mid in (101, 50):
        posenet.write_random_checkpoint(mid, str(d / "_models"), seed=mid)
    make_ref.extract_scripts(str(d / "ref_scripts"))
    return d


def _run(script, workdir, *args):
    env = dict(os.environ, PYTHONPATH=PKG, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, str(workdir / "ref_scripts" / (script + "c"))] + list(args), cwd=str(workdir),
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "%s failed:\n%s\n%s" % (script, r.stdout[-2000:], r.stderr[-3000:])
    return r.stdout


def test_benchmark_py_runs_unchanged(workdir):
    # benchmark.py:16-46 with its default model (101): load_model -> .cuda() -> read_imgfile per file -> model -> decode
    out = _run("benchmark.py", workdir, "--image_dir", "images", "--num_images", "12")
    m = re.search(r"Average FPS: ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) > 0, out


def test_image_demo_py_runs_unchanged(workdir):
    # image_demo.py:20-69: 4-tuple decode, in-place coordinate scaling, overlay drawing, cv2.imwrite, text report
    out = _run("image_demo.py", workdir, "--model", "50", "--scale_factor", "0.75", "--image_dir", "images", "--output_dir", "out")
    assert re.search(r"Average FPS: ([0-9.eE+-]+)", out), out
    assert out.count("Results for image:") == len(SIZES)
    for i, (h, w, ext) in enumerate(SIZES):
        drawn = cv2.imread(str(workdir / "out" / ("img%d.%s" % (i, ext))))
        assert drawn is not None and drawn.shape == (h, w, 3)
    # every reported pose has 17 keypoint lines, each naming a part and a finite (y, x) coordinate
    poses = re.findall(r"Pose #(\d+), score = ([0-9.]+)", out)
    kp_lines = [l for l in out.splitlines() if l.startswith("Keypoint ")]
    assert poses and len(kp_lines) == 17 * len(poses)
    for l in kp_lines:
        m = re.match(r"Keypoint (\w+), score = ([-0-9.eE+]+), coord = \[(.*)\]$", l)
        assert m and m.group(1) in posenet.PART_NAMES, l
        yx = np.array(m.group(3).split(), dtype=np.float64)
        assert yx.shape == (2,) and np.isfinite(yx).all(), l
