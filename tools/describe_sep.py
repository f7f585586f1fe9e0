"""Print the fused-block geometry the library picks for block shapes (host arithmetic only: runs without a GPU).
usage: python tools/describe_sep.py n,h,w,cin,cout,stride,dil [...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "posenet-pytorch_b200")]
from posenet import _native as nat
lib = nat.load()
for a in sys.argv[1:]:
    v = [int(x) for x in a.split(",")]
    d = C.create_string_buffer(512)
    rc = lib.pn_sepconv_describe(*v, d, 512)
    print(a, "->", d.value.decode() if rc == 0 else "rc %d %s" % (rc, lib.pn_last_error_string().decode()))
