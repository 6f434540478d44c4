"""CPU-side checks of the boundary: the shared library loads and exports every symbol include/pisces_b200.h declares, the C++ wrapper
compiles against the header, and — without a GPU — the library fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "pisces_b200.h")).read()
    return sorted(set(re.findall(r"\b(pb2_[a-z_0-9]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from pisces_b200 import _native
    lib = _native.load()
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_native.EXPORTS) == names


def test_record_layout_matches_header():
    from pisces_b200 import _native
    assert C.sizeof(_native.CallRecord) == 96
    import numpy as np
    assert np.dtype(_native.RECORD_DTYPE).itemsize == 96


def test_cpp_wrapper_compiles(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "pisces_b200.hpp"\nint main() { pb2_config c; pb2_default_config(&c); return c.min_base_call_quality == 20 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "pisces_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib, "-lpisces_b200",
                           f"-Wl,-rpath,{lib}"])
    assert subprocess.call([str(exe)]) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pisces_b200 as pb
    with pytest.raises(pb.PiscesB200Error) as e:
        pb.GpuStateManager()
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_integration_doc_binds_every_entry_point():
    """INTEGRATION.md shows the reference-side binding (P/Invoke stubs): every function include/pisces_b200.h declares appears there."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "pisces_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    funcs = sorted(set(re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(funcs) > 40
    assert [f for f in funcs if f not in doc] == []
