"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/ripcurrents_b200.h declares (no compute calls -- there is no GPU here), and it refuses to run
without a CUDA device instead of falling back to a CPU path."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ripcurrents_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from ripcurrents_b200 import build, capi
    build.build()
    return capi.load()


def test_exports_every_declared_symbol(lib):
    from ripcurrents_b200 import capi
    hdr = _header_symbols()
    assert len(hdr) >= 30
    assert sorted(capi.SYMBOLS) == hdr, (set(hdr) ^ set(capi.SYMBOLS))
    for s in hdr:
        assert hasattr(lib, s), "library does not export %s" % s


def test_version_and_error_strings(lib):
    assert lib.rc_version() >= 100
    assert lib.rc_error_string(0) == b"ok"
    assert lib.rc_error_string(-2) == b"CUDA error"
    assert lib.rc_error_string(-99) == b"unknown error"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ripcurrents_b200 import capi
    with pytest.raises(capi.RcError):
        capi.Context(0)


def test_product_does_not_reference_oracle():
    # the oracle is test infrastructure: nothing under ripcurrents_b200/ may import, link or call it
    for dp, _, files in os.walk(os.path.join(ROOT, "ripcurrents_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cpp", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "rc_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_header_is_plain_c_and_cpp(tmp_path):
    """include/ripcurrents_b200.h must compile as C99 and as C++11 on its own (no torch / CUDA / OpenCV types in the ABI)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "ripcurrents_b200.h")
    src_c = tmp_path / "t.c"; src_c.write_text('#include "ripcurrents_b200.h"\nint main(void) { return rc_version() < 0; }\n')
    src_cpp = tmp_path / "t.cpp"; src_cpp.write_text('#include "ripcurrents_b200.h"\nint main() { rc_ctx* c = nullptr; return c != nullptr; }\n')
    inc = os.path.dirname(hdr)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src_c)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src_cpp)])


def test_new_entry_points_reject_null_context_without_cuda():
    """The round-2 entry points (multi-GPU, sharded stream, mask format, averageVector) validate their arguments before any
    CUDA call: a NULL context is RC_ERR_INVALID on a machine without a GPU."""
    import ctypes as C
    from ripcurrents_b200 import capi
    lib = capi.load()
    null = C.c_void_p(0)
    buf = C.create_string_buffer(128)
    assert lib.rc_comm_init(null, buf, 0, 1) == -1
    assert lib.rc_comm_attach(null, null, 0, 1) == -1
    assert lib.rc_comm_destroy(null) == -1
    assert lib.rc_allreduce_accumulators(null, null, null, null) == -1
    assert lib.rc_shard_configure(null, 10, 0) == -1
    assert lib.rc_shard_step(null, null, C.c_size_t(0), C.c_size_t(0), 0, 0, null, null) == -1
    assert lib.rc_shard_report(null, 0, null, null, null) == -1
    assert lib.rc_shard_window_get(null, null) == -1
    assert lib.rc_set_mask_format(null, 1) == -1
    assert lib.rc_average_vector(null, null, null, C.c_size_t(0), 0, 0, null, null, 300, C.c_float(2.0), C.c_float(0.0)) == -1
    assert lib.rc_comm_unique_id(None) == -1
