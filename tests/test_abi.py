"""CPU: the C-ABI shared library loads and exports every symbol include/isochrones_b200.h declares; no compute
call can succeed without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "isochrones_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(iso_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_path():
    names = declared_symbols()
    for required in ("iso_ctx_create", "iso_grid_stage", "iso_grid_repack", "iso_interp_values", "iso_interp_mags",
                     "iso_prior_eval", "iso_models_stage", "iso_lnpost_batch", "iso_lnpost_batch_device",
                     "iso_mnest_prior", "iso_sampler_create", "iso_sampler_run", "iso_nccl_init", "iso_allgather_f64"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from isochrones_b200 import _lib

    L = _lib.lib()
    for name in declared_symbols():
        assert hasattr(L, name), "libisochrones_b200.so does not export %s" % name
    assert set(declared_symbols()) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert L.iso_abi_version() == 3


def test_struct_layout_matches_library():
    from isochrones_b200 import _lib

    L = _lib.lib()
    assert L.iso_struct_size(0) == C.sizeof(_lib.IsoPriorLeaf)
    assert L.iso_struct_size(1) == C.sizeof(_lib.IsoPrior)
    assert L.iso_struct_size(2) == C.sizeof(_lib.IsoModel)
    assert L.iso_struct_size(99) == -1


def test_library_is_sm100a_cuda_code():
    """The product is CUDA for sm_100a: the .so must embed an sm_100a cubin with the fused kernel."""
    import subprocess

    from isochrones_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    sym = subprocess.run(["cuobjdump", "-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "iso_lnpost_kernel" in sym and "iso_interp_values_kernel" in sym


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the context cannot be created and nothing computes (on a GPU box this is skipped)."""
    from isochrones_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.IsoError) as e:
        _lib.Context(0)
    assert e.value.code == -2      # ISO_E_CUDA
    from isochrones_b200 import DFInterpolator

    it = DFInterpolator.from_arrays(np.zeros((2, 2, 1)), (np.arange(2.0), np.arange(2.0)), ["a"])
    with pytest.raises(_lib.IsoError):
        it([0.5, 0.5])


def test_product_does_not_import_oracle():
    """Nothing under isochrones_b200/ may import, load or call the oracle."""
    pkg = os.path.join(ROOT, "isochrones_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "iso_oracle" not in txt and "/root/reference" not in txt, f
