"""MobileNetV1 PoseNet backbone + heads, executed by libposenet_b200 (sm_100a CUDA).

Host-side mirror of the reference's ``posenet/models/mobilenet_v1.py``: the same constructor
(``MobileNetV1(model_id, output_stride=16)``), the same ``state_dict`` keys and shapes (62 tensors:
``features.conv0.conv.*``, ``features.conv{i}.depthwise|pointwise.*``, ``heatmap|offset|
displacement_fwd|displacement_bwd.*``), ``.output_stride`` and a ``forward`` that returns
``(heatmap, offset, displacement_fwd, displacement_bwd)`` as NCHW fp32 tensors with the sigmoid
already applied to the heatmap (reference :156-162).

The ``nn.Conv2d`` children only *hold* parameters -- they are never called, so no cuDNN / ATen
convolution ever runs.  ``forward`` enqueues the hand-written kernels through the C ABI:
stem -> 13 x (depthwise 3x3, pointwise GEMM) -> fused 4-head GEMM.  Without a CUDA device or
without the built library it raises; there is no CPU path.
"""
import ctypes as C
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from posenet import _native as nat

MOBILENET_V1_CHECKPOINTS = {
    50: 'mobilenet_v1_050',
    75: 'mobilenet_v1_075',
    100: 'mobilenet_v1_100',
    101: 'mobilenet_v1_101'
}

# Output width and declared stride of block 0 (3x3 stem) and blocks 1..13 (depthwise-separable).
# The 0.50 / 0.75 variants declare stride 1 for block 12, so they top out at stride 16 (reference
# tables :95-127); width 1.00 serves both ids 100 and 101 (:141-142).
_DECLARED_STRIDES = {100: (2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1), 75: (2, 1, 2, 1, 2, 1, 2) + (1,) * 7,
                     50: (2, 1, 2, 1, 2, 1, 2) + (1,) * 7}
_BASE_WIDTHS = (32, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024, 1024)
_HEAD_CHANNELS = OrderedDict(heatmap=17, offset=34, displacement_fwd=32, displacement_bwd=32)


def _arch(model_id):
    if model_id == 50:
        widths = [min(c // 2, 256) for c in _BASE_WIDTHS]
    elif model_id == 75:
        widths = [min(c * 3 // 4, 384) for c in _BASE_WIDTHS]
    else:
        widths, model_id = list(_BASE_WIDTHS), 100
    return list(zip(widths, _DECLARED_STRIDES[model_id]))


def _to_output_strided_layers(arch, output_stride):
    """Stride -> dilation conversion of the reference (:8-39): layers keep their declared stride until
    the accumulated stride reaches ``output_stride``; from then on they run at stride 1 and inherit
    the product of the strides they swallowed as their dilation ("atrous" path)."""
    layers, reached, rate, cin = [], 1, 1, 3
    for block_id, (cout, declared) in enumerate(arch):
        if reached == output_stride:
            stride, dilation = 1, rate
            rate *= declared
        else:
            stride, dilation = declared, 1
            reached *= declared
        layers.append(dict(block_id=block_id, inp=cin, outp=cout, stride=stride, rate=dilation,
                           output_stride=reached))
        cin = cout
    return layers


def _get_padding(kernel_size, stride, dilation):
    return ((stride - 1) + dilation * (kernel_size - 1)) // 2       # reference :42-44


class InputConv(nn.Module):
    """Parameter holder for the stem (reference :47-54); executed by csrc/stem.cu."""

    def __init__(self, inp, outp, k=3, stride=1, dilation=1):
        super().__init__()
        self.conv = nn.Conv2d(inp, outp, k, stride, padding=_get_padding(k, stride, dilation), dilation=dilation)


class SeperableConv(nn.Module):
    """Parameter holder for one depthwise-separable block (reference :57-68); executed by
    csrc/dwconv.cu + csrc/gemm_tc.cu (bf16) or csrc/gemm_simt.cu (fp32)."""

    def __init__(self, inp, outp, k=3, stride=1, dilation=1):
        super().__init__()
        self.depthwise = nn.Conv2d(inp, inp, k, stride, padding=_get_padding(k, stride, dilation),
                                   dilation=dilation, groups=inp)
        self.pointwise = nn.Conv2d(inp, outp, 1, 1)


class _Plan:
    """One compiled (batch, H, W, dtype, input kind) instance: arena + pn_plan handle."""

    def __init__(self, model, n, h, w, dtype, input_u8, fused=True):
        lib = nat.load()
        pk = model._packed(dtype)
        desc = nat.NetDesc()
        desc.dtype = dtype
        desc.n, desc.h, desc.w = n, h, w
        desc.input_u8 = int(input_u8)
        desc.num_layers = len(model._layers)
        for i, L in enumerate(model._layers):
            d = desc.layers[i]
            d.cin, d.cout, d.stride, d.dilation = L["inp"], L["outp"], L["stride"], L["rate"]
            if i:
                d.dw_w, d.dw_b = pk["dw_w"][i].data_ptr(), pk["dw_b"][i].data_ptr()
            d.pw_w, d.pw_b = pk["pw_w"][i].data_ptr(), pk["pw_b"][i].data_ptr()
        desc.head_w, desc.head_b = pk["head_w"].data_ptr(), pk["head_b"].data_ptr()
        desc.flags = 0 if fused else nat.PLAN_UNFUSED
        arena_bytes, oh, ow = C.c_size_t(), C.c_int(), C.c_int()
        nat.check(lib.pn_plan_query(C.byref(desc), C.byref(arena_bytes), C.byref(oh), C.byref(ow)), "pn_plan_query")
        self.desc, self.packed = desc, pk                      # keep the weight tensors alive
        self.out_h, self.out_w, self.n = oh.value, ow.value, n
        self.arena = torch.empty(arena_bytes.value + 1024, dtype=torch.uint8, device=model._device())
        base = (self.arena.data_ptr() + 1023) & ~1023
        handle = C.c_void_p()
        nat.check(lib.pn_plan_create(C.byref(desc), C.c_void_p(base), arena_bytes, C.byref(handle)), "pn_plan_create")
        self.handle = handle
        self.launches = lib.pn_plan_num_launches(handle)
        self._lib = lib

    def run(self, x, outs):
        nat.check(self._lib.pn_plan_forward(self.handle, C.c_void_p(x.data_ptr()), *[C.c_void_p(o.data_ptr()) for o in outs],
                                            nat.stream_ptr()), "pn_plan_forward")

    def profile(self, x, outs):
        """Per-launch device times (ms) of one forward, in launch order, with their names."""
        ms = (C.c_float * self.launches)()
        nat.check(self._lib.pn_plan_profile(self.handle, C.c_void_p(x.data_ptr()), *[C.c_void_p(o.data_ptr()) for o in outs],
                                            ms, self.launches, nat.stream_ptr()), "pn_plan_profile")
        return list(zip(self.launch_names(), list(ms)))

    def launch_names(self):
        return [self._lib.pn_plan_launch_name(self.handle, i).decode() for i in range(self.launches)]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.pn_plan_destroy(self.handle)
        except Exception:
            pass


class MobileNetV1(nn.Module):

    def __init__(self, model_id, output_stride=16):
        super().__init__()
        assert model_id in MOBILENET_V1_CHECKPOINTS.keys()
        self.model_id = model_id
        self.output_stride = output_stride
        self._layers = _to_output_strided_layers(_arch(model_id), output_stride)
        blocks = []
        for L in self._layers:
            kind = InputConv if L["block_id"] == 0 else SeperableConv
            blocks.append(('conv%d' % L["block_id"], kind(L["inp"], L["outp"], 3, stride=L["stride"], dilation=L["rate"])))
        last_depth = self._layers[-1]["outp"]
        self.features = nn.Sequential(OrderedDict(blocks))
        for name, ch in _HEAD_CHANNELS.items():
            setattr(self, name, nn.Conv2d(last_depth, ch, 1, 1))
        self.compute_dtype = os.environ.get("POSENET_B200_DTYPE", "bf16")
        self.fused_blocks = os.environ.get("POSENET_B200_UNFUSED", "0") != "1"
        self._pack_cache = {}
        self._plans = {}
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    # ------------------------------------------------------------------ configuration
    def set_compute_dtype(self, name):
        """'bf16' (tcgen05 tensor cores, the default) or 'fp32' (FFMA parity mode)."""
        assert name in ("bf16", "fp32"), name
        self.compute_dtype = name
        return self

    def set_fused(self, fused=True):
        """bf16 only: run each SeperableConv block as one fused kernel (default) or as depthwise + pointwise."""
        self.fused_blocks = bool(fused)
        return self

    def _dtype_code(self):
        assert self.compute_dtype in ("bf16", "fp32"), self.compute_dtype
        return nat.PN_BF16 if self.compute_dtype == "bf16" else nat.PN_F32

    def _device(self):
        return self.heatmap.weight.device

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._pack_cache.clear()
        self._plans.clear()
        return out

    def _apply(self, fn, *args, **kwargs):           # .cuda() / .to(): cached packs point at old storage
        out = super()._apply(fn, *args, **kwargs)
        self._pack_cache.clear()
        self._plans.clear()
        return out

    # ------------------------------------------------------------------ weight packing (one-time)
    def _packed(self, dtype):
        """Repack the OIHW fp32 parameters into the layouts the kernels read (include/posenet_b200.h):
        stem [27, C0] f32 (tap-major), depthwise [9, C] f32, pointwise [Cout, Cin] in the activation
        dtype (already K-major), heads concatenated to [128, C] + bias [128]."""
        key = (dtype, self._weights_version())
        if key in self._pack_cache:
            return self._pack_cache[key]
        self._pack_cache.clear()
        tdt = torch.bfloat16 if dtype == nat.PN_BF16 else torch.float32
        pk = dict(dw_w={}, dw_b={}, pw_w={}, pw_b={})
        with torch.no_grad():
            for L in self._layers:
                i = L["block_id"]
                blk = self.features[i]
                if i == 0:
                    pk["pw_w"][0] = blk.conv.weight.float().permute(2, 3, 1, 0).reshape(27, L["outp"]).contiguous()
                    pk["pw_b"][0] = blk.conv.bias.float().contiguous()
                else:
                    pk["dw_w"][i] = blk.depthwise.weight.float().reshape(L["inp"], 9).t().contiguous()
                    pk["dw_b"][i] = blk.depthwise.bias.float().contiguous()
                    pk["pw_w"][i] = blk.pointwise.weight.reshape(L["outp"], L["inp"]).to(tdt).contiguous()
                    pk["pw_b"][i] = blk.pointwise.bias.float().contiguous()
            c = self._layers[-1]["outp"]
            hw = torch.zeros(nat.HEAD_ROWS, c, dtype=torch.float32, device=self._device())
            hb = torch.zeros(nat.HEAD_ROWS, dtype=torch.float32, device=self._device())
            row = 0
            for name, ch in _HEAD_CHANNELS.items():
                conv = getattr(self, name)
                hw[row:row + ch] = conv.weight.reshape(ch, c).float()
                hb[row:row + ch] = conv.bias.float()
                row += ch
            pk["head_w"], pk["head_b"] = hw.to(tdt).contiguous(), hb
        self._pack_cache[key] = pk
        return pk

    def _plan(self, n, h, w, input_u8):
        key = (n, h, w, self._dtype_code(), bool(input_u8), self.fused_blocks, self._weights_version())
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:
                self._plans.clear()
            plan = self._plans[key] = _Plan(self, n, h, w, self._dtype_code(), input_u8, self.fused_blocks)
        return plan

    # ------------------------------------------------------------------ execution
    def _run(self, x, n, h, w, input_u8, out=None):
        nat.require_device()
        if not x.is_cuda or self._device().type != "cuda":
            raise nat.NativeError("MobileNetV1 runs on a CUDA device only: move the model and the input with "
                                  ".cuda() (there is no CPU fallback)")
        if x.device != self._device():
            raise nat.NativeError("input on %s, model on %s" % (x.device, self._device()))
        with torch.cuda.device(x.device):                        # plans, launches and the stream are those of the model's GPU
            plan = self._plan(n, h, w, input_u8)
            if out is None:
                oh, ow = plan.out_h, plan.out_w
                out = tuple(torch.empty((n, ch, oh, ow), dtype=torch.float32, device=x.device)
                            for ch in _HEAD_CHANNELS.values())
            plan.run(x, out)
        return out

    def forward(self, x, out=None):
        """x: fp32 [N,3,H,W] RGB in [-1,1] on the GPU (what ``read_imgfile`` + ``.cuda()`` produce)."""
        assert x.dim() == 4 and x.shape[1] == 3, "expected [N,3,H,W], got %s" % (tuple(x.shape),)
        x = x.contiguous().float()
        return self._run(x, x.shape[0], x.shape[2], x.shape[3], False, out)

    def forward_u8(self, images, out=None):
        """Fused preprocess + forward for already-sized uint8 BGR HWC images [N,H,W,3] on the GPU
        (identity resize: H, W already valid resolutions) -- normalisation happens inside the stem."""
        assert images.dim() == 4 and images.shape[3] == 3 and images.dtype == torch.uint8
        images = images.contiguous()
        return self._run(images, images.shape[0], images.shape[1], images.shape[2], True, out)

    def num_launches(self, n, h, w, input_u8=False):
        return self._plan(n, h, w, input_u8).launches
