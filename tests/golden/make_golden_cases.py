"""Case tables shared by ``make_golden.py`` (which needs the reference) and the tests (which
do not).  Editing a table requires regenerating the fixtures."""
from oracle import synth

PRE_CASES = [  # (h, w, scale_factor, output_stride, seed, keep_full_output)
    (513, 513, 1.0, 16, 1, False), (720, 1280, 1.0, 8, 2, False), (720, 1280, 0.7125, 16, 3, False),
    (257, 257, 1.0, 32, 4, False), (1026, 1026, 0.5, 16, 5, False), (1026, 1026, 0.5001, 16, 6, False),
    (480, 640, 0.5, 16, 7, False), (300, 400, 1.3, 8, 8, False), (100, 37, 2.0, 8, 9, True),
    (37, 100, 1.0, 16, 10, True), (64, 64, 3.7, 16, 11, False), (2000, 31, 0.33, 8, 12, False),
    (17, 17, 1.0, 16, 13, True), (33, 70, 0.9, 8, 14, True), (99, 99, 1 / 3, 8, 15, True),
    (130, 66, 0.5, 8, 16, True), (66, 130, 0.25, 8, 17, True), (1080, 1920, 0.4, 32, 18, False),
]


NET_CASES = [  # (model, output_stride, H, W, batch, init scheme, gain, seed)
    (50, 8, 65, 97, 1, "default", 0.0, 0), (50, 16, 97, 65, 2, "scaled", 1.3, 1),
    (50, 32, 65, 65, 1, "scaled", 0.8, 2), (75, 8, 49, 65, 1, "scaled", 1.3, 3),
    (75, 16, 65, 65, 2, "default", 0.0, 4), (75, 32, 97, 97, 1, "scaled", 0.8, 5),
    (101, 8, 49, 49, 1, "scaled", 1.3, 6), (101, 16, 97, 129, 1, "default", 0.0, 7),
    (101, 32, 129, 97, 1, "scaled", 0.8, 8), (100, 16, 65, 65, 1, "scaled", 1.0, 9),
]


DEC_CASES = [  # (kind, h, w, stride, people|0, seed, P, thr, radius, min_pose, stable_patch, extra)
    ("people", 33, 33, 16, 3, 0, 10, 0.5, 20, 0.25, False, {}),
    ("people", 33, 33, 16, 8, 1, 10, 0.5, 20, 0.5, False, {}),
    ("people", 91, 161, 8, 10, 2, 50, 0.5, 20, 0.25, False, {}),
    ("people", 91, 161, 8, 30, 3, 50, 0.5, 20, 0.25, False, {}),
    ("people", 91, 161, 8, 50, 4, 50, 0.5, 20, 0.25, False, {}),
    ("people", 17, 17, 32, 2, 5, 10, 0.3, 20, 0.0, False, {}),
    ("people", 33, 57, 16, 6, 6, 4, 0.5, 35.5, 0.3, False, {}),
    ("random", 33, 33, 16, 0, 7, 10, 0.9, 20, 0.25, False, {}),
    ("random", 23, 41, 8, 0, 8, 20, 0.7, 10, 0.4, False, {"disp_scale": 90.0}),
    ("random", 17, 17, 16, 0, 9, 10, 0.5, 20, 0.0, False, {"zero_frac": 0.3}),
    ("random", 33, 33, 16, 0, 10, 10, 0.5, 20, 0.25, True, {"tie_levels": 16}),
    ("random", 9, 9, 32, 0, 11, 30, 0.0, 5, 0.0, True, {"tie_levels": 4, "zero_frac": 0.2}),
    ("random", 1, 1, 16, 0, 12, 10, 0.0, 20, 0.0, False, {}),
    ("random", 5, 3, 16, 0, 13, 10, 1.5, 20, 0.5, False, {}),          # empty: threshold above all scores
    # more accepted poses than the decoder caches in shared memory (DEC_ACC = 64 in csrc/decode.cu): 100 of 120 people
    # (the cap is reached), and 90 people with the cap out of reach
    ("people", 91, 161, 8, 120, 20, 100, 0.5, 20, 0.25, False, {}),
    ("people", 91, 161, 8, 90, 21, 100, 0.5, 20, 0.25, False, {}),
]

# decode_pose / traverse_to_targ_keypoint called on their own (posenet/decode.py:9-63,131-182): (index into DEC_CASES,
# number of best candidates used as roots).  make_golden.py adds, per case, roots with a zero and a negative score.
POSE_CASES = [(0, 6), (3, 8), (8, 8), (9, 12), (11, 6)]


def heads_for(kind, h, w, stride, people, seed, extra):
    if kind == "people":
        return synth.people_heads(h, w, stride, people, seed)[:4]
    return synth.random_heads(h, w, seed, **extra)


