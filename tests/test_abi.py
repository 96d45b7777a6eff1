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
