"""Load the UNMODIFIED reference ``network/memory.py`` (container-only helper).

TEST INFRASTRUCTURE. ``/root/reference`` exists only in the build container, never on
the GPU box, so this loader is used by ``oracle/make_golden.py`` (fixture generation)
and by the CPU tests that pin the oracle against the live reference; everything that
runs on the GPU box uses the committed fixtures in ``tests/golden/`` instead.

Recipe: the reference file imports one unused foreign symbol
(``transforms.transforms.HideAndSeek``, memory.py:7) and calls ``.cuda()``
unconditionally in its constructor (memory.py:111,121); both are shimmed here without
touching the reference tree.
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PINMEM_REFERENCE_ROOT", "/root/reference")
_REF_FILE = os.path.join(REFERENCE_ROOT, "network", "memory.py")


def reference_available() -> bool:
    return os.path.isfile(_REF_FILE)


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
    if not reference_available():
        raise FileNotFoundError(_REF_FILE)
    if "transforms.transforms" not in sys.modules:
        pkg = types.ModuleType("transforms")
        pkg.__path__ = []
        sub = types.ModuleType("transforms.transforms")
        sub.HideAndSeek = type("HideAndSeek", (), {})
        sys.modules.setdefault("transforms", pkg)
        sys.modules["transforms.transforms"] = sub
    spec = importlib.util.spec_from_file_location("pinmem_reference_memory", _REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


def build_reference_memory(memory_size=19, dim=256, momentum=0.8, temperature=1.0, gumbel_read=False):
    """Construct the reference ``Memory_sup`` (on CPU when no GPU is present)."""
    mod = load_reference_module()
    with _cuda_identity_on_cpu():
        return mod.Memory_sup(memory_size, dim, dim, momentum, temperature, gumbel_read)
