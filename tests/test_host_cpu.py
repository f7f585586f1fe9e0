"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol,
the drop-in package mirrors the reference's API surface, and nothing silently falls back."""
import os
import re

import numpy as np
import pytest
import torch

import posenet
from oracle import decode as odec
from oracle import net as onet
from posenet import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "posenet_b200.h")).read()
    declared = set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    lib = nat.load()                                       # raises if missing / symbol absent / ABI mismatch
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pn_abi_version() == nat.ABI_VERSION == 4


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(nat.Map) == 40 and C.sizeof(nat.DecodeParams) == 24 and C.sizeof(nat.Layer) == 48
    assert C.sizeof(nat.NetDesc) == 24 + 16 * 48 + 16 + 8          # + flags, padded to pointer alignment


def test_sepconv_tile_geometry_is_host_computable():
    """pn_sepconv_describe is pure host arithmetic: every block of every model/stride has a tile shape."""
    import ctypes as C
    lib, buf = nat.load(), C.create_string_buffer(256)
    for mid, os_, hw in ((101, 16, 513), (101, 8, 257), (50, 8, 721), (75, 32, 257), (100, 32, 129)):
        m = posenet.MobileNetV1(mid, output_stride=os_)
        h = w = hw
        for L in m._layers:
            s, d = L["stride"], L["rate"]
            if L["block_id"]:
                assert lib.pn_sepconv_describe(2, h, w, L["inp"], L["outp"], s, d, buf, 256) == 0, lib.pn_last_error_string()
                assert b"tile" in buf.value or b"warp-autonomous strips" in buf.value or b"tensor-pipe" in buf.value
            pad = ((s - 1) + 2 * d) // 2
            h = w = (h + 2 * pad - 2 * d - 1) // s + 1
    assert lib.pn_sepconv_describe(2, 9, 9, 64, 64, 2, 2, buf, 256) != 0            # never produced by the tables


def test_tensor_pipe_depthwise_policy_and_geometry(monkeypatch):
    """PN_SEP_TC=1 routes stride-1 blocks with 64-multiple widths (<= 512) to csrc/septc.cu; band layout is host arithmetic."""
    import ctypes as C
    lib, buf = nat.load(), C.create_string_buffer(512)
    monkeypatch.delenv("PN_SEP_TC", raising=False)                  # default policy: the 256 -> 256 blocks only
    assert lib.pn_sepconv_describe(64, 33, 33, 512, 512, 1, 1, buf, 512) == 0 and b"tensor-pipe" not in buf.value
    assert lib.pn_sepconv_describe(32, 91, 161, 256, 256, 1, 2, buf, 512) == 0 and b"tensor-pipe" in buf.value
    monkeypatch.setenv("PN_SEP_TC", "0")
    assert lib.pn_sepconv_describe(32, 91, 161, 256, 256, 1, 2, buf, 512) == 0 and b"tensor-pipe" not in buf.value
    monkeypatch.setenv("PN_SEP_TC", "1")
    for shp, want in (((64, 33, 33, 512, 512, 1, 1), b"bands 1 x 33 cols pitch 34"), ((64, 129, 129, 128, 128, 1, 1), b"A ring x4"),
                      ((32, 91, 161, 256, 256, 1, 2), b"tensor-pipe"), ((512, 17, 17, 384, 384, 1, 1), b"A cache x6")):
        assert lib.pn_sepconv_describe(*shp, buf, 512) == 0, lib.pn_last_error_string()
        assert b"tensor-pipe depthwise" in buf.value and want in buf.value, buf.value
    for shp in ((64, 65, 65, 256, 512, 2, 1), (64, 257, 257, 32, 64, 1, 1), (2, 33, 33, 1024, 1024, 1, 2), (3, 65, 65, 96, 96, 1, 1)):
        assert lib.pn_sepconv_describe(*shp, buf, 512) == 0 and b"tensor-pipe" not in buf.value     # stride 2, narrow, wide, ragged


@pytest.mark.parametrize("mid", [50, 75, 100, 101])
@pytest.mark.parametrize("os_", [8, 16, 32])
def test_layer_table_and_state_dict_mirror_reference(mid, os_):
    m = posenet.MobileNetV1(mid, output_stride=os_)
    assert [(L["inp"], L["outp"], L["stride"], L["rate"]) for L in m._layers] == \
           [(L["cin"], L["cout"], L["stride"], L["dilation"]) for L in onet.layer_table(mid, os_)]
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == onet.param_shapes(mid)
    assert list(sd) == list(onet.param_shapes(mid))
    m.load_state_dict(onet.init_params(mid, 0), strict=True)
    assert m.output_stride == os_


def test_api_surface():
    for name in ("load_model", "MobileNetV1", "MOBILENET_V1_CHECKPOINTS", "decode_multiple_poses", "read_imgfile", "read_cap",
                 "valid_resolution", "draw_skel_and_kp", "draw_skeleton", "draw_keypoints", "get_adjacent_keypoints",
                 "PART_NAMES", "NUM_KEYPOINTS", "PARENT_CHILD_TUPLES", "CONNECTED_PART_INDICES", "LOCAL_MAXIMUM_RADIUS",
                 "POSE_CHAIN", "PART_IDS", "decode", "decode_multi"):
        assert hasattr(posenet, name), name
    assert posenet.decode_multi.decode_multiple_poses is posenet.decode_multiple_poses
    assert tuple(posenet.PARENT_CHILD_TUPLES) == tuple(odec.EDGES)
    assert posenet.NUM_KEYPOINTS == 17 and len(posenet.CONNECTED_PART_INDICES) == 12
    assert posenet.valid_resolution(1280 * 0.7125, 720 * 0.7125, 16) == (913, 513)
    import inspect
    sig = inspect.signature(posenet.decode_multiple_poses)
    assert [(p.name, p.default) for p in sig.parameters.values()][4:] == [
        ("output_stride", inspect.Parameter.empty), ("max_pose_detections", 10), ("score_threshold", 0.5),
        ("nms_radius", 20), ("min_pose_score", 0.5)]


def test_load_model_needs_checkpoint_and_writes_random(tmp_path):
    with pytest.raises(FileNotFoundError):
        posenet.load_model(75, model_dir=str(tmp_path))
    posenet.write_random_checkpoint(75, str(tmp_path), seed=3)
    m = posenet.load_model(75, model_dir=str(tmp_path))
    assert sum(p.numel() for p in m.parameters()) == 1258195


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    m = posenet.MobileNetV1(50)
    with pytest.raises(nat.NativeError):
        m(torch.zeros(1, 3, 33, 33))
    with pytest.raises(nat.NativeError):
        posenet.decode_multiple_poses(torch.zeros(17, 3, 3), torch.zeros(34, 3, 3), torch.zeros(32, 3, 3),
                                      torch.zeros(32, 3, 3), 16)
    with pytest.raises(nat.NativeError):
        posenet._process_input(np.zeros((17, 17, 3), np.uint8))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "posenet-pytorch_b200")
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle\b)|oracle/|oracle\.(net|decode|preprocess|synth)", re.M)
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(d, f)).read()), os.path.join(d, f)
