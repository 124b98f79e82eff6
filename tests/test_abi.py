"""CPU checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every
symbol that include/mimo_b200.h declares; argument validation works without touching the device."""
import ctypes as C
import os
import re

import pytest

from mimo_unet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mimo_b200.h")).read()
    names = set(re.findall(r"\b(mimo_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 40
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in mimo_b200.h but not exported"
    assert names == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_version(lib):
    assert lib.mimo_version() == 100


def test_plan_geometry_and_errors(lib):
    cfg = _lib.UnetConfig(3, 2, 2, 21, 4, 128, 160)
    h = C.c_void_p()
    assert lib.mimo_unet_plan_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.mimo_unet_num_state(h) == 172          # 172 state_dict entries at M=2
    assert lib.mimo_unet_num_double_convs(h) == 12    # 3*S + 6
    assert lib.mimo_unet_workspace_bytes(h) > 0
    assert [lib.mimo_unet_dropout_channels(h, i) for i in range(12)] == [21, 21, 42, 42, 168, 336, 336, 168, 84, 42, 21, 21]
    # unbound plan -> state error, message available
    rc = lib.mimo_unet_forward(h, None, None, 1, None, None, None)
    assert rc == -5 and b"not bound" in lib.mimo_last_error()
    lib.mimo_unet_plan_destroy(h)
    # too-small images are rejected like the reference (reflect padding needs >= 2 px at 1/16 scale)
    bad = _lib.UnetConfig(3, 2, 2, 8, 1, 24, 40)
    assert lib.mimo_unet_plan_create(C.byref(bad), C.byref(h)) == -1
    assert b">= 32" in lib.mimo_last_error()


def test_m4_state_count(lib):
    cfg = _lib.UnetConfig(3, 2, 4, 21, 2, 64, 64)
    h = C.c_void_p()
    assert lib.mimo_unet_plan_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.mimo_unet_num_state(h) == 7 * 2 * (3 * 4 + 6) + 2 * 4
    lib.mimo_unet_plan_destroy(h)


def test_null_pointer_rejected(lib):
    a = _lib.Act(None, 1, 8, 8, 1, 8, 0, 3)
    assert lib.mimo_pack_input(None, 0, 0, None, a, None) == -1
    assert lib.mimo_lossbuffer_init(None, 2, 10, 0.3, None) == -1
