"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/cmdiad_b200.h declares, and fails loudly
(no CPU fallback) when no sm_100 GPU is present.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cmdiad_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmdb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(built):
    from cmdiad_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature"
    assert set(_lib.SYMBOLS) == set(names)
    assert lib.cmdb_version() >= 100


def test_score_out_struct_layout(built):
    from cmdiad_b200 import _lib
    assert ctypes.sizeof(_lib.ScoreOut) == 11 * ctypes.sizeof(ctypes.c_void_p)


def test_argument_validation_without_gpu(built):
    from cmdiad_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.cmdb_bank_create(0, 100, 10, ctypes.byref(h)) == _lib.CMDB_ERR_INVALID  # dim not a multiple of 64
    assert b"multiple of 64" in lib.cmdb_last_error()
    assert lib.cmdb_bank_create(0, 768, 0, ctypes.byref(h)) == _lib.CMDB_ERR_INVALID
    assert lib.cmdb_bank_rows(None, None) == _lib.CMDB_ERR_INVALID
    assert lib.cmdb_score(None, None, 0, 0, 0, 0, 0, None) == _lib.CMDB_ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback(built):
    from cmdiad_b200 import Bank, upsample_blur
    from cmdiad_b200._lib import CmdbError, CMDB_ERR_CUDA
    with pytest.raises(CmdbError) as e:
        Bank(768, 16)
    assert e.value.status == CMDB_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(CmdbError):
        upsample_blur(np.ones((28, 28), np.float32))


def test_product_path_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under cmdiad_b200/ may import, link or call it"""
    pkg = os.path.join(ROOT, "cmdiad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "restate" not in txt.replace("restated", ""), f


def test_methods_mirror_reference_interface():
    """same method / attribute names as feature_extractors/features.py + multiple_features.py"""
    from cmdiad_b200 import methods as M
    for cls in (M.RGBFeatures, M.DepthFeatures, M.PointFeatures, M.DoubleRGBPointFeatures,
                M.RGBorXYZWithOneHallucination, M.RGBorXYZWithOneHallucinationFromFeature):
        for name in ("add_sample_to_mem_bank", "run_coreset", "add_sample_to_late_fusion_mem_bank", "run_late_fusion",
                     "predict", "compute_s_s_map", "calculate_dist", "compute_single_s_s_map",
                     "get_coreset_idx_randomp", "calculate_metrics"):
            assert callable(getattr(cls, name)), (cls.__name__, name)
    assert set(M.METHODS) == {"DINO", "Point_MAE", "DINO+Point_MAE", "WithHallucination", "WithHallucinationFromFeature"}
    d = M.DoubleRGBPointFeatures
    assert d.mean_from == {"xyz": "xyz", "rgb": "xyz"} and d.std_from == {"xyz": "rgb", "rgb": "rgb"}  # the quirk
