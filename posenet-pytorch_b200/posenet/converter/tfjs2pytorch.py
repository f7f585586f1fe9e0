"""TF.js PoseNet checkpoint (manifest.json + one raw little-endian float32 file per variable) -> the state_dict of
``posenet.MobileNetV1`` -- SURVEY 8(f) N3, the reference's ``posenet/converter/tfjs2pytorch.py:15-72``.

Same names and results as the reference: ``to_torch_name`` (variable name mapping, tfjs2pytorch.py:15-43),
``load_variables`` (layout transposes: conv HWIO -> OIHW, depthwise HWC1 -> C1HW, tfjs2pytorch.py:46-72) and ``convert``
(write ``<model_dir>/<checkpoint>.pth``, tfjs2pytorch.py:85-97).  There is no network in this environment, so a missing
manifest is an error that says where the files must be placed (the reference would download them, converter/wget.py:28-41);
``write_tfjs_checkpoint`` is the inverse transform, used by the tests to build a manifest from a seeded state_dict.
"""
import json
import os
import tempfile

import numpy as np
import torch

from posenet.models.mobilenet_v1 import MOBILENET_V1_CHECKPOINTS, MobileNetV1

BASE_DIR = os.path.join(tempfile.gettempdir(), '_posenet_weights')


def to_torch_name(tf_name):
    """'MobilenetV1/Conv2d_3_depthwise/depthwise_weights' -> 'features.conv3.depthwise.weight'; head variables
    ('.../heatmap_2/weights', 'offset_2', 'displacement_fwd_2', 'displacement_bwd_2') -> 'heatmap.weight', ...;
    anything else -> '' (skipped)."""
    parts = tf_name.lower().split('/')
    if len(parts) < 3:
        return ''
    layer, kind = parts[1].split('_'), parts[2]
    postfix = {'weights': '.weight', 'depthwise_weights': '.weight', 'biases': '.bias'}.get(kind, '')
    if layer[0] == 'conv2d':
        sub = layer[2] if len(layer) > 2 else 'conv'          # conv2d_0 is the stem (InputConv.conv)
        return 'features.conv%s.%s%s' % (layer[1], sub, postfix)
    if layer[0] in ('offset', 'displacement', 'heatmap') and layer[-1] == '2':
        return '_'.join(layer[:-1]) + postfix
    return ''


def load_variables(chkpoint, base_dir=None):
    base_dir = BASE_DIR if base_dir is None else base_dir
    manifest_path = os.path.join(base_dir, chkpoint, "manifest.json")
    if not os.path.exists(manifest_path):
        raise FileNotFoundError(
            "TF.js weights for checkpoint %s are not at %s and cannot be downloaded here (no network). Place manifest.json and "
            "the variable files of https://storage.googleapis.com/tfjs-models/weights/posenet/%s/ in that directory."
            % (chkpoint, os.path.dirname(manifest_path), chkpoint))
    with open(manifest_path) as f:
        variables = json.load(f)
    state_dict = {}
    for name, meta in variables.items():
        torch_name = to_torch_name(name)
        if not torch_name:
            continue
        d = np.fromfile(os.path.join(base_dir, chkpoint, meta["filename"]), dtype='<f4')
        shape = meta["shape"]
        if len(shape) == 4:
            tpt = (2, 3, 0, 1) if 'depthwise' in meta["filename"] else (3, 2, 0, 1)
            d = np.reshape(d, shape).transpose(tpt)
        state_dict[torch_name] = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
    return state_dict


def convert(model_id, model_dir, output_stride=16, image_size=513, check=True, base_dir=None):
    """Write ``<model_dir>/<checkpoint>.pth``.  ``check`` / ``image_size`` are accepted for signature compatibility; the
    reference's check only prints a few head values of ./images/tennis_in_crowd.jpg when that file exists."""
    checkpoint_name = MOBILENET_V1_CHECKPOINTS[model_id]
    os.makedirs(model_dir, exist_ok=True)
    state_dict = load_variables(checkpoint_name, base_dir)
    m = MobileNetV1(model_id, output_stride=output_stride)
    m.load_state_dict(state_dict)                              # strict: every one of the 62 tensors, right shapes
    path = os.path.join(model_dir, checkpoint_name) + '.pth'
    torch.save(m.state_dict(), path)
    return path


def _tf_name(torch_name):
    base, kind = torch_name.rsplit('.', 1)
    if base.startswith('features.conv'):
        idx, sub = base[len('features.conv'):].split('.')
        layer = 'Conv2d_%s' % idx if sub == 'conv' else 'Conv2d_%s_%s' % (idx, sub)
        var = 'biases' if kind == 'bias' else ('depthwise_weights' if sub == 'depthwise' else 'weights')
    else:
        layer, var = base + '_2', 'biases' if kind == 'bias' else 'weights'
    return 'MobilenetV1/%s/%s' % (layer, var)


def write_tfjs_checkpoint(state_dict, chkpoint, base_dir=None):
    """Inverse of ``load_variables``: write ``state_dict`` as a TF.js checkpoint directory (manifest + raw files)."""
    out = os.path.join(BASE_DIR if base_dir is None else base_dir, chkpoint)
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for key, t in state_dict.items():
        a = t.detach().cpu().numpy().astype('<f4')
        name = _tf_name(key)
        if a.ndim == 4:
            a = a.transpose((2, 3, 0, 1)) if 'depthwise' in name else a.transpose((2, 3, 1, 0))     # C1HW -> HWC1, OIHW -> HWIO
        filename = name.replace('/', '_')
        np.ascontiguousarray(a).tofile(os.path.join(out, filename))
        manifest[name] = {"filename": filename, "shape": list(a.shape)}
    with open(os.path.join(out, "manifest.json"), "w") as f:
        json.dump(manifest, f)
    return out
