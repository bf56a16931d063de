"""CPU checks of the C-ABI boundary: the library builds/loads and exports every symbol include/bhs.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "bhs.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bhs_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from biem_helmholtz_sphere_b200 import _lib

    decl = _declared_symbols()
    assert decl, "no symbols parsed from include/bhs.h"
    assert sorted(_lib.SIGNATURES) == decl


def test_library_loads_and_exports_all_symbols():
    from biem_helmholtz_sphere_b200 import _lib, build

    build.build()
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"libbhs.so does not export {name}"
    assert lib.bhs_version() >= 100


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    import ctypes as C

    from biem_helmholtz_sphere_b200 import _lib

    lib = _lib.load()
    assert lib.bhs_plan_create(1, 4, C.byref(C.c_void_p())) == -1
    assert lib.bhs_plan_create(3, 0, C.byref(C.c_void_p())) == -1
    assert lib.bhs_bessel(3, 7, 0, 4, None, 0, None, None) == -1
    assert lib.bhs_zgesv_workspace(0, 1) == -1
    assert lib.bhs_zgesv_workspace(4096, 1) > 0
    assert lib.bhs_zgemm_workspace(64, 64, 0) == -1
    assert lib.bhs_uscat(None, 1, None, None, 1.0, 0.0, 1.0, None, None, 0, 0, None, None, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1)
    with pytest.raises(NotImplementedError):
        _lib.check(-2)


def test_header_is_valid_c_and_cxx(tmp_path):
    """include/bhs.h is the drop-in boundary: it must compile as plain C (gcc) and as C++ (g++) on its own."""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "bhs.h")
    for cc, lang, std in (("gcc", "c", "-std=c99"), ("g++", "c++", "-std=c++11")):
        if shutil.which(cc) is None:
            pytest.skip(f"{cc} not available")
        src = tmp_path / ("t." + ("c" if lang == "c" else "cpp"))
        src.write_text('#include "bhs.h"\nint main(void) { int (*f)(void) = bhs_version; (void)f; return 0; }\n')
        r = subprocess.run([cc, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", f"-I{os.path.dirname(hdr)}", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
