"""CPU: the C-ABI library builds/loads here and exports every symbol include/pinmem_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pinmem_b200.h")


def _declared():
    """name -> number of parameters, parsed from the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int|const char\*)\s+(pm_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_header_declares_the_documented_entry_points():
    d = _declared()
    for name in ("pm_read_fwd", "pm_read_bwd", "pm_read_bwd_dM", "pm_colsoftmax", "pm_readloss_fwd",
                 "pm_write_reduce_fwd", "pm_update_fwd", "pm_update_bwd", "pm_write_bwd", "pm_score_nhwc",
                 "pm_rowsoftmax", "pm_colsoftmax_apply", "pm_status_string", "pm_version"):
        assert name in d, name


def test_library_exports_every_declared_symbol():
    from pinthememory_b200 import capi

    lib = capi.load()
    assert os.path.exists(capi.library_path())
    raw = ctypes.CDLL(capi.library_path())
    for name in _declared():
        assert hasattr(raw, name), "missing export: " + name
    assert lib.pm_version() >= 100
    assert lib.pm_score_stride(19) == 20 and lib.pm_score_stride(20) == 32 and lib.pm_score_stride(31) == 32
    assert b"mem_dim" in lib.pm_status_string(-3)


def test_binding_prototypes_match_the_header():
    from pinthememory_b200 import capi

    d = _declared()
    assert set(capi.EXPORTED_SYMBOLS) == set(d)
    for name, argtypes in capi.PROTOTYPES.items():
        assert len(argtypes) == d[name], "%s: binding has %d args, header %d" % (name, len(argtypes), d[name])


def test_argument_checks_run_without_a_gpu():
    """Rejected arguments return negative status codes before any CUDA call."""
    from pinthememory_b200 import capi

    lib = capi.load()
    one = ctypes.c_void_p(16)
    assert lib.pm_read_fwd(None, one, None, None, one, one, one, None, 1, 256, 4, 4, 19, 0, None) == -1
    assert lib.pm_read_fwd(one, one, None, None, one, one, one, None, 1, 48, 4, 4, 19, 0, None) == -3
    assert lib.pm_read_fwd(one, one, None, None, one, one, one, None, 1, 256, 4, 4, 0, 0, None) == -4
    assert lib.pm_read_fwd(one, one, None, None, one, one, one, None, 1, 256, 4, 4, 19, 7, None) == -2
    assert lib.pm_read_fwd(one, one, None, None, one, one, one, None, 0, 256, 4, 4, 19, 0, None) == -5
    assert lib.pm_write_reduce_fwd(one, one, ctypes.c_void_p(4), 1, 256, 4, 4, 8, 8, 19, 0, None) == -6
