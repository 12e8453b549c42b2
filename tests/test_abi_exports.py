"""The C-ABI library loads and exports every symbol include/dracob200.h declares; without a GPU it refuses to
decode (there is no CPU fallback)."""
import ctypes as C
import os
import re

import draco_sharp_b200 as D
from draco_sharp_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dracob200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb_[a-z_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    lib = C.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    bound = {e[0] for e in N.EXPORTS}
    assert set(names) == bound


def test_version_and_error_strings():
    assert N.lib().dcb_version() == 100
    for code in list(range(0, -16, -1)) + [-100, -101, -102, -103, -104]:
        s = N.lib().dcb_error_string(code)
        assert s and s != b"unknown error"


def test_struct_sizes_match_the_header():
    # blittable layouts the C# P/Invoke declarations mirror field for field
    assert C.sizeof(N.BufferInfo) == 56
    assert C.sizeof(N.AttrInfo) == 112
    assert C.sizeof(N.LaunchStats) == 152


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    try:
        D.DracoBatchDecoder()
    except D.DracoError as e:
        assert e.code == N.DCB_ERR_NO_DEVICE
    else:
        raise AssertionError("decoder construction must fail without a B200")
    assert N.lib().dcb_device_count() == 0


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", N.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "orc_" not in out
    needed = subprocess.run(["ldd", N.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "liboracle" not in needed
