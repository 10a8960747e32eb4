"""The C-ABI libraries load and export every symbol declared in include/*.h (no compute calls:
there is no GPU on the build box), and the product refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(lbmk?_\w+)\s*\(", text, flags=re.M)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_runtime_exports_every_declared_symbol():
    from pylbm_b200 import build, runtime

    lib = ctypes.CDLL(build.build_runtime())
    declared = _declared("lbm_b200.h")
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), "liblbm_b200.so does not export %s" % name
    # the ctypes binding covers exactly the header
    assert sorted(runtime.EXPORTED_SYMBOLS) == declared
    lib.lbm_abi_version.restype = ctypes.c_int
    assert lib.lbm_abi_version() == 1


def test_generated_library_exports_lbmk_symbols():
    from pylbm_b200 import cases
    from pylbm_b200.scheme import Scheme
    from pylbm_b200.simulation import build_kernel_library

    _, path, source = build_kernel_library(Scheme(cases.karman_d2q9(nx=128, ny=32)), need_source=True)
    lib = ctypes.CDLL(path)
    in_place_only = {"lbmk_one_time_step_aa", "lbmk_one_time_step_aa_walls", "lbmk_f2m_sw", "lbmk_f2m_consm_sw"}
    declared = [n for n in _declared("lbmk.h") if n != "lbmk_source_term"]   # only with source terms
    for name in declared:
        assert hasattr(lib, name) == (name not in in_place_only), "%s: export of %s" % (path, name)
    # a library generated for in-place streaming exports everything
    _, path_aa, _ = build_kernel_library(Scheme(cases.karman_d2q9(nx=128, ny=32)), aa=True)
    lib_aa = ctypes.CDLL(path_aa)
    for name in declared:
        assert hasattr(lib_aa, name), "%s does not export %s" % (path_aa, name)
    lib.lbmk_describe.restype = ctypes.c_char_p
    assert b'"one_time_step"' in lib.lbmk_describe()
    # the struct of the header and of the generated source agree field by field
    hdr = open(os.path.join(ROOT, "include", "lbmk.h")).read()
    fields = lambda text: re.findall(r"^\s*(?:int64_t|int)\s+(\w+)(?:\[\d\])?;", text[text.index("typedef struct"):text.index("} lbmk_grid;")], flags=re.M)
    assert fields(hdr) == fields(source)


def test_struct_layouts_match_the_headers():
    from pylbm_b200.runtime import LbmkGrid, LbmSimDesc

    assert ctypes.sizeof(LbmkGrid) == 4 * (3 + 3 + 3 + 1 + 3 + 1) + 4 + 3 * 8 - 4 + 0 or ctypes.sizeof(LbmkGrid) == 80
    assert LbmkGrid.pitch.offset == 56 and LbmkGrid.pstride.offset == 72
    assert LbmSimDesc.grid.offset == 8
    assert LbmSimDesc.vel.size == 192


def test_no_cpu_fallback():
    """on a box without a GPU the product path must fail loudly."""
    from pylbm_b200 import runtime

    if runtime.lib().lbm_device_count() > 0:
        pytest.skip("a GPU is present")
    import pylbm_b200
    from pylbm_b200 import cases

    with pytest.raises(runtime.LbmError):
        pylbm_b200.Simulation(cases.karman_d2q9(nx=64, ny=32))
    with pytest.raises(ValueError):
        pylbm_b200.Simulation(dict(cases.karman_d2q9(nx=64, ny=32), generator="numpy"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pylbm_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("# oracle", ""), "%s mentions the oracle" % fn
