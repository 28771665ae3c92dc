"""CPU: the C-ABI library builds, loads and exports every symbol include/i2p_b200.h declares."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "i2p_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(i2p_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from i2pnet_b200 import _build, _cabi
    path = _build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == _cabi.exported_symbols()
    lib.i2p_abi_version.restype = ctypes.c_int
    assert lib.i2p_abi_version() == 1


def test_kernels_are_sm_100a_only():
    """One arch, no PTX-JIT fallback to other GPUs, no spills in any kernel."""
    import glob
    from i2pnet_b200 import _build
    _build.build_library()
    logs = glob.glob(os.path.join(ROOT, "i2pnet_b200", "build", "*.ptxas.txt"))
    assert logs
    text = "".join(open(p).read() for p in logs)
    assert "sm_100a" in text and not re.search(r"for 'sm_(?!100a)", text)
    assert not re.search(r"[1-9]\d* bytes spill", text)
