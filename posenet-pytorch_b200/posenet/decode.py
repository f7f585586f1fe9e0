"""Placeholder for the reference's ``posenet.decode`` module (posenet/decode.py).

``traverse_to_targ_keypoint`` (decode.py:9-63) and ``decode_pose`` (decode.py:131-182) are internal
steps of ``decode_multiple_poses``; here they are fused into one CUDA kernel
(csrc/decode.cu: ``hop`` and the greedy loop) and are not callable on their own.  The helpers the
reference also keeps in that file (``find_root``, ``print_decoded_heatmap``, ...) are dead code
there and are out of scope.
"""


def _fused(name):
    def stub(*_a, **_k):
        raise NotImplementedError(
            "posenet.decode.%s is fused into the CUDA decoder; call posenet.decode_multiple_poses" % name)
    stub.__name__ = name
    return stub


traverse_to_targ_keypoint = _fused("traverse_to_targ_keypoint")
decode_pose = _fused("decode_pose")
