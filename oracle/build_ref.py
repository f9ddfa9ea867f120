"""Byte-compile the reference's network/memory.py into oracle/_ref/ (build container only).

TEST INFRASTRUCTURE. The reference is pure Python, so its "compiled from its own sources where they lie" form is a
.pyc: ``py_compile`` of /root/reference/network/memory.py, unmodified, written to oracle/_ref/reference_memory.bytecode
together with a manifest (source path, SHA-256 of the source, interpreter). oracle/_ref/ is git-ignored (an output, like
a built .so; no reference source text enters the repository) but not gpurun-ignored, so it travels to the GPU box, where
``bench.py --impl reference`` and ``cpu_baseline`` time THE REFERENCE'S OWN MODULE on the host cores instead of the
oracle port (``kind: "reference"``), and ``reference_eager_same_gpu`` times it eagerly on the B200.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is mounted)
"""
import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PINMEM_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REFERENCE_ROOT, "network", "memory.py")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "reference_memory.bytecode")


def build():
    """Returns the path of the .pyc, or None when the reference tree is not mounted."""
    if not os.path.isfile(SRC):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    py_compile.compile(SRC, cfile=OUT, dfile="reference/network/memory.py", doraise=True)
    with open(SRC, "rb") as fh:
        digest = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(OUT_DIR, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": digest, "python": sys.version.split()[0],
                   "what": "py_compile of the unmodified reference file; output only, git-ignored"}, fh, indent=1)
    return OUT


if __name__ == "__main__":
    print(build())
