"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/mvae_b200.h declares (no compute)."""
import ctypes
import os
import re

from multiscale_variational_autoencoder_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mvae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mvae_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    for n in _lib.PROTOTYPES:
        assert n in names, f"{n} bound in _lib.py but missing from the header"


def test_version_and_error_text():
    lib = _lib.load()
    assert lib.mvae_version() == 100
    # argument errors are reported, not thrown: null descriptor
    rc = lib.mvae_conv2d_fwd(None, None, None, None, None, None, 0, None, None)
    assert rc == -1 and "descriptor" in _lib.last_error()


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_conv_desc_layout_matches_header():
    assert ctypes.sizeof(_lib.ConvDesc) == 11 * 4
