"""The C-ABI library loads and exports every symbol include/vsgpu.h declares; without a GPU it fails
loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import vs_testlib as T

HEADER = os.path.join(T.ROOT, "include", "vsgpu.h")


def header_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vsgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from variantstore_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(_lib.PROTOTYPES) == syms                       # the ctypes binding covers the header exactly
    lib = _lib.load()
    for s in syms:
        assert getattr(lib, s) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", T.VSGPU_SO], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vsgpu_[a-z0-9_]+)", out))
    assert set(syms) <= exported


def test_product_does_not_link_the_oracle():
    out = subprocess.run(["ldd", T.VSGPU_SO], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hostsim" not in out
    nm = subprocess.run(["nm", "-D", T.VSGPU_SO], capture_output=True, text=True).stdout
    assert "vso_" not in nm and "3vso" not in nm
    for root, _, files in os.walk(os.path.join(T.ROOT, "variantstore_b200")):
        for f in files:
            if f.endswith((".cc", ".cu", ".h", ".cuh", ".py")):
                src = open(os.path.join(root, f)).read()
                assert "oracle/" not in src and "liboracle" not in src and "vso.h" not in src, f


@pytest.mark.skipif(T.has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_no_answer():
    from variantstore_b200 import VariantStoreIndex, VsgpuError
    with pytest.raises(VsgpuError) as ei:
        VariantStoreIndex(os.path.join(T.GOLDEN, "x_ser"), device=0)
    assert ei.value.code == -4 and "no CPU path" in str(ei.value)


def test_bad_directory_is_an_io_error():
    from variantstore_b200 import VariantStoreIndex, VsgpuError, load_library
    with pytest.raises(VsgpuError) as ei:
        VariantStoreIndex("/nonexistent/ser", lib=load_library(T.HOSTSIM_SO, subset=True))
    assert ei.value.code == -2


def test_hostsim_is_test_only():
    """The host build of the kernel logic lives under tests/ and exports the query subset only."""
    from variantstore_b200 import _lib
    lib = _lib.load(T.HOSTSIM_SO, subset=True)
    assert not hasattr(lib, "vsgpu_batch_create_does_not_exist")
    with pytest.raises(AttributeError):
        lib.vsgpu_batch_create
    assert T.HOSTSIM_SO.startswith(os.path.join(T.ROOT, "tests"))
