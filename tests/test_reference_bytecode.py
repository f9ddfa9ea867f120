"""CPU: oracle/_ref/reference_memory.bytecode -- the byte-compiled, unmodified reference module that `bench.py --impl
reference` times on the GPU box (oracle/build_ref.py) -- loads without the reference tree and behaves like the oracle.

Skipped when the file has not been built (a checkout that never saw /root/reference)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

from golden_util import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYC = os.path.join(ROOT, "oracle", "_ref", "reference_memory.bytecode")
pytestmark = pytest.mark.skipif(not os.path.isfile(PYC), reason="oracle/_ref not built (reference tree never mounted)")


def test_manifest_matches_the_mounted_reference_source():
    with open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")) as fh:
        man = json.load(fh)
    assert man["python"] == sys.version.split()[0]
    if os.path.isfile(man["source"]):   # build container: the bytecode is of the file as it lies in the reference tree
        with open(man["source"], "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == man["sha256"]


def test_bytecode_module_runs_without_the_reference_tree_and_matches_the_oracle():
    """In a child process with the reference root pointed at nothing (as on the GPU box)."""
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import ref_loader, memory_oracle as mo
assert not ref_loader.reference_available() and ref_loader.compiled_reference_available()
from pinthememory_b200 import synth
ref = ref_loader.build_reference_memory(19, 32, 0.8, 1.0, False, force_cpu=True)
assert ref_loader.load_reference_module().__pinmem_origin__ == "bytecode"
ora = mo.OracleMemorySup(19, 32, 32, 0.8, 1.0, False)
ora.load_state_dict(ref.state_dict()); ora.m_items = ref.m_items.clone()
x = synth.make_features(2, 32, 9, 11, seed=5); labels = synth.make_labels(2, 40, 52, 19, "iid", seed=6)
xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
with ref_loader.cuda_identity(force=True):
    r = ref(xr, labels, True, False)
    (r[0].square().sum() + r[3] + r[4][0] + r[4][1]).backward()
o = ora(xo, labels, True, False)
(o[0].square().sum() + o[3] + o[4][0] + o[4][1]).backward()
torch.save({"uq": (o[0].detach(), r[0].detach()), "rl": (o[3].detach(), r[3].detach()),
            "mem": (ora.m_items.detach(), ref.m_items.detach()), "dx": (xo.grad, xr.grad)}, sys.argv[1])
''' % ROOT
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), "pinmem_ref_bytecode_%d.pt" % os.getpid())
    env = dict(os.environ, PINMEM_REFERENCE_ROOT="/nonexistent-reference-root", PINMEM_B200_DEVICE="cpu")
    r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = torch.load(out)
    os.remove(out)
    for k, (a, b) in got.items():
        assert_close(a, b, 2e-6, k)
