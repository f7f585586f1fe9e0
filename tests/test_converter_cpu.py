"""SURVEY 8(f) N3 -- TF.js checkpoint conversion (posenet/converter/tfjs2pytorch.py of the reference).  The name mapping is
pinned by a golden table generated from the reference's own ``to_torch_name`` (tests/golden/make_golden.py ->
tests/golden/converter_names.json); the layout transposes by a round trip through a synthetic TF.js checkpoint."""
import json
import os

import numpy as np
import pytest
import torch

import posenet
from posenet.converter import tfjs2pytorch as conv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "converter_names.json")


def test_variable_names_match_the_reference():
    table = json.load(open(GOLDEN))
    assert len(table) == 68 and sum(1 for v in table.values() if v) == 63        # incl. an activation name the reference maps to a prefix
    for tf_name, want in table.items():
        assert conv.to_torch_name(tf_name) == want, tf_name


def test_converted_tensors_match_the_reference(tmp_path):
    """Same synthetic TF.js checkpoint as tests/golden/make_golden.py::make_converter_names (seeded values, TF layouts):
    every converted tensor must hash to what the reference's load_variables produced."""
    import hashlib
    table = json.load(open(GOLDEN))
    want = json.load(open(os.path.join(os.path.dirname(GOLDEN), "converter_tensors.json")))
    inv = {v: k for k, v in table.items() if v}
    rng = np.random.default_rng(7)
    ck = tmp_path / "mobilenet_v1_050"
    ck.mkdir()
    manifest = {}
    for key, v in posenet.MobileNetV1(50).state_dict().items():
        shp, tfn = tuple(v.shape), inv[key]
        if len(shp) == 4:
            tf_shape = [shp[2], shp[3], shp[0], shp[1]] if "depthwise" in tfn else [shp[2], shp[3], shp[1], shp[0]]
        else:
            tf_shape = list(shp)
        fn = tfn.replace("/", "_")
        rng.standard_normal(tf_shape).astype("<f4").tofile(str(ck / fn))
        manifest[tfn] = {"filename": fn, "shape": tf_shape}
    json.dump(manifest, open(str(ck / "manifest.json"), "w"))
    got = conv.load_variables("mobilenet_v1_050", str(tmp_path))
    assert set(got) == set(want)
    for k, (shape, digest) in want.items():
        assert list(got[k].shape) == shape, k
        assert hashlib.sha256(np.ascontiguousarray(got[k].numpy()).tobytes()).hexdigest() == digest, k


@pytest.mark.parametrize("mid", [50, 101])
def test_round_trip_through_a_tfjs_checkpoint(mid, tmp_path):
    torch.manual_seed(mid)
    sd = posenet.MobileNetV1(mid).state_dict()
    sd = {k: torch.randn(v.shape) for k, v in sd.items()}
    name = posenet.MOBILENET_V1_CHECKPOINTS[mid]
    out = conv.write_tfjs_checkpoint(sd, name, str(tmp_path))
    manifest = json.load(open(os.path.join(out, "manifest.json")))
    assert len(manifest) == 62
    # TF layouts: conv HWIO, depthwise HWC1
    assert manifest["MobilenetV1/Conv2d_0/weights"]["shape"] == [3, 3, 3, sd["features.conv0.conv.weight"].shape[0]]
    c1 = sd["features.conv1.depthwise.weight"].shape[0]
    assert manifest["MobilenetV1/Conv2d_1_depthwise/depthwise_weights"]["shape"] == [3, 3, c1, 1]
    raw = np.fromfile(os.path.join(out, manifest["MobilenetV1/Conv2d_1_depthwise/depthwise_weights"]["filename"]), dtype="<f4")
    assert raw.reshape(3, 3, c1, 1)[1, 2, 5, 0] == sd["features.conv1.depthwise.weight"][5, 0, 1, 2].item()
    back = conv.load_variables(name, str(tmp_path))
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    path = conv.convert(mid, str(tmp_path / "models"), base_dir=str(tmp_path))
    loaded = torch.load(path)
    assert list(loaded) == list(posenet.MobileNetV1(mid).state_dict()) and all(torch.equal(loaded[k], sd[k]) for k in sd)


def test_load_model_converts_when_the_tfjs_files_are_present(tmp_path, monkeypatch):
    monkeypatch.setattr(conv, "BASE_DIR", str(tmp_path / "w"))
    sd = {k: torch.randn(v.shape) for k, v in posenet.MobileNetV1(50).state_dict().items()}
    conv.write_tfjs_checkpoint(sd, posenet.MOBILENET_V1_CHECKPOINTS[50], str(tmp_path / "w"))
    m = posenet.load_model(50, model_dir=str(tmp_path / "m"))
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    with pytest.raises(FileNotFoundError):
        posenet.load_model(75, model_dir=str(tmp_path / "m"))          # nothing on disk, no network
    with pytest.raises(FileNotFoundError):
        conv.load_variables("mobilenet_v1_075", str(tmp_path / "w"))
