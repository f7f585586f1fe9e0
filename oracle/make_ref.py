"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference, compiled from the read-only checkout (TEST INFRASTRUCTURE).

    python oracle/make_ref.py            # run in the build container; /root/reference must exist

The reference is pure Python; "building" it means byte-compiling its sources WHERE THEY LIE under /root/reference into
CPython bytecode files (no reference source text is copied into this repository):

    /root/reference/posenet/**/*.py   -> oracle/_ref/reference.zip : posenet/**/*.pyc   (sourceless package, imported through zipimport)
    /root/reference/benchmark.py      -> oracle/_ref/scripts.zip   : benchmark.pyc
    /root/reference/image_demo.py     -> oracle/_ref/scripts.zip   : image_demo.pyc

(two uncompressed zip archives rather than loose ``.pyc`` files: the gpurun snapshot leaves ``*.pyc`` behind).  ``oracle/_ref/`` is
git-ignored but NOT gpurun-ignored, so it travels to the GPU box with the working tree like the built ``.so`` files (same
image, same interpreter: the bytecode loads there).  Two consumers, both test / measurement infrastructure:

* ``bench.py --impl reference`` and the ``cpu_baseline`` leg put ``oracle/_ref/reference.zip`` on ``sys.path`` (``IMPORT_PATH``), import
  the reference's ``posenet`` from it (``kind: "reference"``) and time
  the reference's own code on the host cores; without ``oracle/_ref`` they fall back to the oracle port (``kind: "port"``).
* ``tests/test_gpu_ref_scripts.py`` runs the two scripts UNCHANGED (their bytecode, extracted with ``extract_scripts`` and run as
  ``python benchmark.pyc ...``) as subprocesses against the product package (``PYTHONPATH=posenet-pytorch_b200``).  The scripts
  live in their own archive so that ``import posenet`` inside them resolves to the product package, not to the reference.

``SCRIPT_SHA256`` pins the script SOURCES the bytecode was compiled from: the recipe refuses to compile anything else and
records the digests it saw in ``oracle/_ref/MANIFEST.json``; ``tests/test_host_cpu.py`` checks them against ``/root/reference``
(when present) and against the manifest, so "unchanged" is checked and not assumed.
"""
import hashlib
import json
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SCRIPTS = ("benchmark.py", "image_demo.py")
SCRIPT_SHA256 = {
    "benchmark.py": "9daad25cee3639efd0463afb751dd134c6021a95efe3f31a8518623ac4fc4750",
    "image_demo.py": "5f723d26a23bb0be65a51a34ccf888ee5d8535707e05c5f721da989ff4f2d110",
}


def sha256_file(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def tree_digest(root):
    """sha256 over (relative path, content) of every .py file under ``root``, in sorted order."""
    h = hashlib.sha256()
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if x != "__pycache__")
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


IMPORT_PATH = os.path.join(REF_DIR, "reference.zip")          # put this on sys.path to import the reference's `posenet`
SCRIPTS_ZIP = os.path.join(REF_DIR, "scripts.zip")


def available():
    return os.path.exists(IMPORT_PATH) and os.path.exists(SCRIPTS_ZIP) and os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def manifest():
    path = os.path.join(REF_DIR, "MANIFEST.json")
    return json.load(open(path)) if os.path.exists(path) else None


def extract_scripts(dst_dir):
    """Unpack benchmark.pyc / image_demo.pyc into ``dst_dir`` (a scratch directory) and return their paths by script name."""
    os.makedirs(dst_dir, exist_ok=True)
    out = {}
    with zipfile.ZipFile(SCRIPTS_ZIP) as z:
        for s in SCRIPTS:
            z.extract(s + "c", dst_dir)
            out[s] = os.path.join(dst_dir, s + "c")
    return out


def _compile(src, shown_as):
    """Bytecode of ``src`` as the bytes of a .pyc file (unchecked-hash invalidation: independent of source mtimes)."""
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "x.pyc")
        py_compile.compile(src, cfile=dst, dfile=shown_as, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        with open(dst, "rb") as f:
            return f.read()


def make(reference="/root/reference", quiet=False):
    """Byte-compile the reference package and the two caller scripts into oracle/_ref/.  Returns True when the build exists
    afterwards (False: no reference checkout here and nothing built earlier)."""
    if not os.path.isdir(os.path.join(reference, "posenet")):
        return available()
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    os.makedirs(REF_DIR)
    n = 0
    src_pkg = os.path.join(reference, "posenet")
    with zipfile.ZipFile(IMPORT_PATH, "w", zipfile.ZIP_STORED) as z:
        for d, dirs, files in os.walk(src_pkg):
            dirs[:] = sorted(x for x in dirs if x != "__pycache__")
            for f in sorted(files):
                if f.endswith(".py"):
                    rel = os.path.relpath(os.path.join(d, f), reference)
                    z.writestr(rel + "c", _compile(os.path.join(d, f), "<reference>/" + rel))
                    n += 1
    seen = {}
    with zipfile.ZipFile(SCRIPTS_ZIP, "w", zipfile.ZIP_STORED) as z:
        for s in SCRIPTS:
            seen[s] = sha256_file(os.path.join(reference, s))
            assert seen[s] == SCRIPT_SHA256[s], "%s changed upstream: %s" % (s, seen[s])
            z.writestr(s + "c", _compile(os.path.join(reference, s), "<reference>/" + s))
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"scripts_sha256": seen, "package_sha256": tree_digest(src_pkg), "package_files": n,
                   "python": "%d.%d" % sys.version_info[:2]}, f, indent=1, sort_keys=True)
    if not quiet:
        print("oracle/_ref: %d package files + %s byte-compiled from %s" % (n, ", ".join(SCRIPTS), reference))
    return True


if __name__ == "__main__":
    sys.exit(0 if make(*(sys.argv[1:2])) else 1)
