"""Oracle for the MobileNetV1 backbone + 4 heads (TEST INFRASTRUCTURE).

A functional torch-CPU fp32 restatement of ``posenet/models/mobilenet_v1.py`` of the
reference: the architecture tables (:78-127), the stride->dilation conversion
``_to_output_strided_layers`` (:8-39), the padding rule ``_get_padding`` (:42-44),
``InputConv`` (:47-54), ``SeperableConv`` (:57-68) and ``MobileNetV1.forward``
(:156-162).  Pinned against the reference module by ``tests/golden/net_*.npz``.
"""
import math

import torch
import torch.nn.functional as F

# (cout, declared stride) per block; block 0 is the 3x3 stem, 1..13 are separable blocks.
_WIDTHS_100 = [(32, 2), (64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2),
               (512, 1), (512, 1), (512, 1), (512, 1), (512, 1), (1024, 2), (1024, 1)]
# mobilenet_v1.py:95-127 -- the 0.75 / 0.50 tables declare stride 1 for block 12
_WIDTHS_75 = [(24, 2), (48, 1), (96, 2), (96, 1), (192, 2), (192, 1), (384, 2),
              (384, 1), (384, 1), (384, 1), (384, 1), (384, 1), (384, 1), (384, 1)]
_WIDTHS_50 = [(16, 2), (32, 1), (64, 2), (64, 1), (128, 2), (128, 1), (256, 2),
              (256, 1), (256, 1), (256, 1), (256, 1), (256, 1), (256, 1), (256, 1)]

HEADS = (("heatmap", 17), ("offset", 34), ("displacement_fwd", 32), ("displacement_bwd", 32))


def arch(model_id):
    if model_id == 50:
        return _WIDTHS_50
    if model_id == 75:
        return _WIDTHS_75
    if model_id in (100, 101):
        return _WIDTHS_100          # mobilenet_v1.py:141-142: 100 and 101 share a table
    raise AssertionError("unknown model id %r" % (model_id,))


def layer_table(model_id, output_stride):
    """mobilenet_v1.py:8-39: once the running stride reaches ``output_stride`` every later
    layer runs at stride 1 with the accumulated rate as its dilation."""
    table = []
    cur, rate, cin = 1, 1, 3
    for blk, (cout, s) in enumerate(arch(model_id)):
        if cur == output_stride:
            ls, ld = 1, rate
            rate *= s
        else:
            ls, ld = s, 1
            cur *= s
        table.append(dict(block=blk, cin=cin, cout=cout, stride=ls, dilation=ld,
                          padding=((ls - 1) + ld * 2) // 2))     # :42-44 with k=3
        cin = cout
    return table


def param_shapes(model_id):
    """state_dict key -> shape, in the reference's ``named_parameters`` order."""
    shapes = {}
    cin = 3
    for blk, (cout, _) in enumerate(arch(model_id)):
        if blk == 0:
            shapes["features.conv0.conv.weight"] = (cout, 3, 3, 3)
            shapes["features.conv0.conv.bias"] = (cout,)
        else:
            shapes["features.conv%d.depthwise.weight" % blk] = (cin, 1, 3, 3)
            shapes["features.conv%d.depthwise.bias" % blk] = (cin,)
            shapes["features.conv%d.pointwise.weight" % blk] = (cout, cin, 1, 1)
            shapes["features.conv%d.pointwise.bias" % blk] = (cout,)
        cin = cout
    for name, ch in HEADS:
        shapes[name + ".weight"] = (ch, cin, 1, 1)
        shapes[name + ".bias"] = (ch,)
    return shapes


def init_params(model_id, seed=0, scheme="default", gain=1.3):
    """Seeded random-init state_dict (fp32, reference key names).

    ``default``: the distribution torch gives ``nn.Conv2d`` (weight and bias both
    U(-1/sqrt(fan_in), 1/sqrt(fan_in))) -- north_star's "random-init weights".
    ``scaled``: SURVEY Appendix B -- He-normal x gain backbone, head gains 3/12/40, so the
    ReLU6 clamp, the full sigmoid range and multi-cell displacements are exercised.
    """
    g = torch.Generator().manual_seed(int(seed))
    head_gain = {"heatmap": 3.0, "offset": 12.0, "displacement_fwd": 40.0, "displacement_bwd": 40.0}
    sd = {}
    for key, shape in param_shapes(model_id).items():
        mod = key.rsplit(".", 1)[0]
        if key.endswith(".weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            if scheme == "default":
                b = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * b
            else:
                std = (head_gain[mod] if mod in head_gain else gain * math.sqrt(2.0)) / math.sqrt(fan_in)
                t = torch.randn(shape, generator=g) * std
        else:
            w = sd[mod + ".weight"]
            fan_in = w.shape[1] * w.shape[2] * w.shape[3]
            if scheme == "default":
                b = 1.0 / math.sqrt(fan_in)
            else:
                b = 0.5 if mod in head_gain else 0.1
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        sd[key] = t.float().contiguous()
    return sd


def forward(sd, model_id, output_stride, x, return_features=False):
    """mobilenet_v1.py:156-162 on a state_dict; x is f32 [N,3,H,W].  Returns the four NCHW
    head tensors (heatmap already through sigmoid)."""
    feats = []
    with torch.no_grad():
        for L in layer_table(model_id, output_stride):
            b = L["block"]
            if b == 0:
                x = F.relu6(F.conv2d(x, sd["features.conv0.conv.weight"], sd["features.conv0.conv.bias"],
                                     stride=L["stride"], padding=L["padding"], dilation=L["dilation"]))
            else:
                p = "features.conv%d." % b
                x = F.relu6(F.conv2d(x, sd[p + "depthwise.weight"], sd[p + "depthwise.bias"],
                                     stride=L["stride"], padding=L["padding"], dilation=L["dilation"],
                                     groups=L["cin"]))
                if return_features:
                    feats.append(x)
                x = F.relu6(F.conv2d(x, sd[p + "pointwise.weight"], sd[p + "pointwise.bias"]))
            if return_features:
                feats.append(x)
        heat = torch.sigmoid(F.conv2d(x, sd["heatmap.weight"], sd["heatmap.bias"]))
        off = F.conv2d(x, sd["offset.weight"], sd["offset.bias"])
        fwd = F.conv2d(x, sd["displacement_fwd.weight"], sd["displacement_fwd.bias"])
        bwd = F.conv2d(x, sd["displacement_bwd.weight"], sd["displacement_bwd.bias"])
    if return_features:
        return (heat, off, fwd, bwd), feats
    return heat, off, fwd, bwd


def out_hw(model_id, output_stride, h, w):
    for L in layer_table(model_id, output_stride):
        k = 2 * L["dilation"] + 1
        h = (h + 2 * L["padding"] - k) // L["stride"] + 1
        w = (w + 2 * L["padding"] - k) // L["stride"] + 1
    return h, w
