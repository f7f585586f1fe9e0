"""``load_model`` with the reference's signature (posenet/models/model_factory.py:11-23).

The reference downloads and converts TF.js checkpoints when the ``.pth`` is missing; here the conversion
(``posenet.converter.tfjs2pytorch``) runs when the TF.js files are already on disk, and otherwise a missing
file is an error that names the helper which writes seeded random-init weights
(``write_random_checkpoint``) -- the weights every parity test uses.
"""
import math
import os

import torch

from posenet.models.mobilenet_v1 import MOBILENET_V1_CHECKPOINTS, MobileNetV1

MODEL_DIR = './_models'
DEBUG_OUTPUT = False


def load_model(model_id, output_stride=16, model_dir=MODEL_DIR):
    path = os.path.join(model_dir, MOBILENET_V1_CHECKPOINTS[model_id] + '.pth')
    if not os.path.exists(path):
        from posenet.converter import tfjs2pytorch
        manifest = os.path.join(tfjs2pytorch.BASE_DIR, MOBILENET_V1_CHECKPOINTS[model_id], "manifest.json")
        if os.path.exists(manifest):                               # the reference's fallback (model_factory.py:13-17), minus the download
            print('Cannot find models file %s, converting from tfjs...' % path)
            tfjs2pytorch.convert(model_id, model_dir, check=False)
    if not os.path.exists(path):
        raise FileNotFoundError(
            "Cannot find model file %s. TF.js checkpoint conversion needs network access; write seeded "
            "random-init weights with posenet.models.model_factory.write_random_checkpoint(%d, %r) or "
            "place a converted state_dict there." % (path, model_id, model_dir))
    model = MobileNetV1(model_id, output_stride=output_stride)
    model.load_state_dict(torch.load(path, map_location="cpu"))
    return model


def write_random_checkpoint(model_id, model_dir=MODEL_DIR, seed=0):
    """Write ``<model_dir>/mobilenet_v1_XXX.pth`` holding a seeded state_dict with the reference's
    key names and torch's default Conv2d init distribution (U(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""
    os.makedirs(model_dir, exist_ok=True)
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for key, ref in MobileNetV1(model_id).state_dict().items():
        w = ref if key.endswith(".weight") else sd[key[:-len("bias")] + "weight"]
        bound = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
        sd[key] = ((torch.rand(ref.shape, generator=g) * 2 - 1) * bound).float()
    path = os.path.join(model_dir, MOBILENET_V1_CHECKPOINTS[model_id] + '.pth')
    torch.save(sd, path)
    return path
