"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, and exports every symbol include/gode.h
declares.  No compute calls (there is no GPU in the build container)."""
import ctypes

import pytest

from gan_ode_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_every_declared_symbol_has_a_python_signature():
    assert sorted(_lib._SIGS) == _lib.declared_symbols()


def test_introspection(lib):
    assert b"sm_100a" in lib.gode_version()
    assert lib.gode_param_count(16, 16) == 544 and lib.gode_param_count(64, 256) == 33088
    assert lib.gode_supported(16, 16, 0) == 1
    assert lib.gode_supported(17, 16, 0) == 0
    assert lib.gode_strerror(0) == b"ok" and b"workspace" in lib.gode_strerror(-4)


def test_argument_errors_do_not_touch_the_gpu(lib):
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.gode_rk4_fwd(None, None, None, None, None, None, 0, 4, 16, 16, 16, 0, 0, None, None) == -2
    assert ctypes.sizeof(_lib.GodeAdaptiveOpts) == 88 and ctypes.sizeof(_lib.GodeStepLog) == 32


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """include/gode.h is the drop-in boundary: it must compile as C99 without any CUDA or torch header, and the ctypes
    mirrors in gan_ode_b200/_lib.py must have the C compiler's struct sizes."""
    import os
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include "gode.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu\\n", sizeof(GodeAdaptiveOpts), sizeof(GodeStepLog), '
                   'sizeof(GodeWorld), (size_t)GODE_GRAD_SLOT_WORDS(8, 544)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(_lib.GodeAdaptiveOpts), ctypes.sizeof(_lib.GodeStepLog),
                                     ctypes.sizeof(_lib.GodeWorld), 2 * 8 * 544]


def test_new_entry_points_reject_bad_arguments_before_any_cuda_call(lib):
    o = _lib.GodeAdaptiveOpts()
    w = _lib.GodeWorld()
    # continuous dopri5 adjoint: null tensors, bad parameter mask
    assert lib.gode_dopri5_adjoint_bwd(None, None, None, None, None, None, None, 4, 16, 16, 2, 0, ctypes.byref(o), 15,
                                       None, None, None, None, None, None, None, 0, None) == -2
    # world-scope norm / fused exchange: a null or inconsistent GodeWorld
    assert lib.gode_dopri5_fwd_world(None, None, None, None, None, None, 4, 16, 16, 2, ctypes.byref(o), 0, None, None, None,
                                     None, None, None, None, None, None, None, 0, ctypes.byref(w), None) == -2
    assert lib.gode_dopri5_backprop_bwd_world(None, None, None, None, None, None, 4, 16, 16, 2, 0, None, None, None, None, 8,
                                              ctypes.c_float(1.0), None, None, None, 0, ctypes.byref(w), None) == -2
    assert lib.gode_rk4_bwd_world(1, None, None, None, None, None, None, None, 0, 4, 16, 16, 2, 0, None, None, None, 0,
                                  ctypes.byref(w), None) == -2
    assert lib.gode_dopri5_adjoint_workspace_bytes(8192, 16, 16) > 0
