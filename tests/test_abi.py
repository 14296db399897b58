"""The C-ABI library: it loads without a GPU, exports every symbol include/fra.h
declares, the Python binding lists the same set, and on a box with no CUDA device
it refuses to work instead of falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from fpga_real_time_fft_analyzer_b200 import build
    return build.build()


def header_functions():
    text = open(os.path.join(ROOT, "include", "fra.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fra_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    fns = header_functions()
    for must in ("fra_create", "fra_destroy", "fra_command", "fra_process", "fra_process_host", "fra_load_bank1",
                 "fra_set_mode", "fra_reset", "fra_get_state", "fra_set_state", "fra_iir_stream", "fra_fft_only",
                 "fra_strerror"):
        assert must in fns


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in header_functions():
        assert hasattr(lib, name), f"libfra.so does not export {name}"


def test_binding_covers_header_exactly():
    from fpga_real_time_fft_analyzer_b200 import _abi
    assert sorted(_abi.SIGNATURES) == header_functions()


def test_sm100a_sass_present(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_window_rom_export_matches_reference(built_lib, rom):
    from fpga_real_time_fft_analyzer_b200 import FraContext
    assert np.array_equal(FraContext.window_rom(), rom)


def test_strerror_and_version(built_lib):
    from fpga_real_time_fft_analyzer_b200 import _abi
    from fpga_real_time_fft_analyzer_b200._lib import lib
    L = lib()
    assert L.fra_abi_version() == _abi.FRA_ABI_VERSION
    assert L.fra_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.fra_strerror(_abi.FRA_ERR_NO_DEVICE)


def test_no_cpu_fallback_without_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on CPU-only boxes")
    from fpga_real_time_fft_analyzer_b200 import FraContext, FraError
    with pytest.raises(FraError) as e:
        FraContext(4)
    assert e.value.status == -2


def test_argument_validation_needs_no_device(built_lib):
    from fpga_real_time_fft_analyzer_b200._lib import lib
    L = lib()
    h = ctypes.c_void_p()
    assert L.fra_create(ctypes.byref(h), 0, 0, 16384, 0) == -1        # no channels
    assert L.fra_create(ctypes.byref(h), 0, 4, 12345, 0) == -1        # not a power of two
    assert L.fra_create(ctypes.byref(h), 0, 4, 131072, 0) == -5       # outside 1K..64K
    assert L.fra_destroy(None) == -1
    assert L.fra_process(None, None, 0, 0, None, None) == -1


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "fpga_real_time_fft_analyzer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f
                assert "cusim" not in text or f == "fra_common.cuh", f
