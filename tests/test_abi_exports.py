"""CPU: the C-ABI library loads and exports every symbol include/achelous_b200.h declares."""
import os
import re

from achelous_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "achelous_b200.h")).read()
    declared = set(re.findall(r"ACH_API\s+[\w\s\*]+?\b(ach_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name          # dlsym succeeds
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.ach_version() >= 100
    assert lib.ach_last_error() is not None


def test_bad_arguments_return_status_not_crash():
    lib = _lib.load()
    s = _lib.AchPwConv()   # all null
    import ctypes as C
    st = lib.ach_pw_conv(C.byref(s), None)
    assert st != 0 and b"ach_pw_conv" in lib.ach_last_error()


def test_build_entry_point_is_idempotent():
    """__graft_entry__.build() is the driver's "does it build" check and runs every round on a tree where oracle/_ref was already
    staged once (round 2: the second call tripped over the recipe's own STAGED_FROM marker)."""
    import __graft_entry__ as g
    g.build()
    g.build()
