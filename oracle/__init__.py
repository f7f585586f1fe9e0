"""CPU oracle for the PoseNet hot path (TEST INFRASTRUCTURE - not product code).

This package restates, in numpy (integer / float64 work) and plain torch-CPU fp32
(the convolution stack), the algorithm of the reference path

    preprocess -> MobileNetV1 backbone + 4 heads -> decode_multiple_poses

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package (``posenet-pytorch_b200/``) never does.

Pinning status: the reference ships no tests, golden vectors or fixtures
("parity unpinned" by reference tests).  The oracle is instead
pinned against outputs of the reference code itself, executed in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference`` and
writes the fixtures under ``tests/golden/`` that ``tests/test_oracle_*.py``
replay on any machine.
"""
