"""Load the UNMODIFIED reference ``network/memory.py`` (container-only helper).

TEST INFRASTRUCTURE. ``/root/reference`` exists only in the build container, never on
the GPU box, so this loader is used by ``oracle/make_golden.py`` (fixture generation)
and by the CPU tests that pin the oracle against the live reference; everything that
runs on the GPU box uses the committed fixtures in ``tests/golden/`` instead -- except
``bench.py --impl reference`` / ``cpu_baseline``, which TIME the reference's own module from
the byte-compiled ``oracle/_ref/reference_memory.bytecode`` (oracle/build_ref.py) when it is there.

Recipe: the reference file imports one unused foreign symbol
(``transforms.transforms.HideAndSeek``, memory.py:7) and calls ``.cuda()``
unconditionally in its constructor (memory.py:111,121); both are shimmed here without
touching the reference tree.
"""
import contextlib
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PINMEM_REFERENCE_ROOT", "/root/reference")
_REF_FILE = os.path.join(REFERENCE_ROOT, "network", "memory.py")
# Byte-compiled copy of the same file, made by oracle/build_ref.py where the reference tree is mounted (the build
# container). Like a .so built from C sources it is an OUTPUT: git-ignored, never edited, no source text -- but it travels
# to the GPU box with the snapshot, so `bench.py --impl reference` can time the reference's own module there.
_REF_PYC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_memory.bytecode")


def reference_available() -> bool:
    return os.path.isfile(_REF_FILE)


def compiled_reference_available() -> bool:
    return os.path.isfile(_REF_PYC)


@contextlib.contextmanager
def cuda_identity(force=False):
    """Make ``Tensor.cuda`` / ``Module.cuda`` the identity: always on a CPU host (the reference calls ``.cuda()``
    unconditionally, memory.py:111,121,246), and on request on a GPU host (``force``: keep the module on the CPU for the
    CPU timing arm)."""
    if torch.cuda.is_available() and not force:
        yield
        return
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


@contextlib.contextmanager
def _cuda_identity_on_cpu():
    """Make ``Tensor.cuda`` / ``Module.cuda`` the identity while building on a CPU host."""
    if torch.cuda.is_available():
        yield
        return
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


_cached = None


def load_reference_module():
    """Return the reference ``network.memory`` module object, loaded by file path."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available() and not compiled_reference_available():
        raise FileNotFoundError(_REF_FILE)
    if "transforms.transforms" not in sys.modules:
        pkg = types.ModuleType("transforms")
        pkg.__path__ = []
        sub = types.ModuleType("transforms.transforms")
        sub.HideAndSeek = type("HideAndSeek", (), {})
        sys.modules.setdefault("transforms", pkg)
        sys.modules["transforms.transforms"] = sub
    if reference_available():
        spec = importlib.util.spec_from_file_location("pinmem_reference_memory", _REF_FILE)
    else:   # the GPU box: the byte-compiled module (same interpreter, same image)
        loader = importlib.machinery.SourcelessFileLoader("pinmem_reference_memory", _REF_PYC)
        spec = importlib.util.spec_from_loader("pinmem_reference_memory", loader)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.__pinmem_origin__ = "source" if reference_available() else "bytecode"
    _cached = mod
    return mod


def build_reference_memory(memory_size=19, dim=256, momentum=0.8, temperature=1.0, gumbel_read=False, force_cpu=False):
    """Construct the reference ``Memory_sup`` (on CPU when no GPU is present, or when ``force_cpu``)."""
    mod = load_reference_module()
    with cuda_identity(force_cpu):
        return mod.Memory_sup(memory_size, dim, dim, momentum, temperature, gumbel_read)
