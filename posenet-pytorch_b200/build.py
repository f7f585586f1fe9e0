"""Build libposenet_b200.so (sm_100a only) in-tree with nvcc.

    python posenet-pytorch_b200/build.py [--force] [--verbose]

Objects are compiled in parallel into ``build/`` and linked into ``lib/libposenet_b200.so`` (the product: the hot path only)
and ``lib/libposenet_b200_diag.so`` (hardware probes of ``csrc/diag/``, include/posenet_b200_diag.h; tests and tools only).  The
CUDA runtime is linked statically, so the library loads (and exports its symbols) on a machine
without a GPU or driver; only calling a compute entry point needs a B200.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libposenet_b200.so")
DIAG_LIB = os.path.join(LIB_DIR, "libposenet_b200_diag.so")
DIAG_SRC = os.path.join(CSRC, "diag")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("PN_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _diag_sources():
    return sorted(os.path.join(DIAG_SRC, f) for f in os.listdir(DIAG_SRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    root = os.path.dirname(HERE)
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if os.path.isfile(os.path.join(CSRC, f))] + _diag_sources() + \
        [os.path.join(root, "include", "posenet_b200.h"), os.path.join(root, "include", "posenet_b200_diag.h")]
    for p in deps:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp_file = os.path.join(BUILD, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(DIAG_LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(BUILD, ("diag_" if os.path.dirname(src) == DIAG_SRC else "") + os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, _sources() + _diag_sources()))
    n_prod = len(_sources())
    tmap_obj = os.path.join(BUILD, "tmap.o")                     # the probes encode tensor maps too
    for lib, members in ((LIB, objs[:n_prod]), (DIAG_LIB, objs[n_prod:] + [tmap_obj])):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + members
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
