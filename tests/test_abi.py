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


def test_tile_schedule_covers_every_query_on_every_cta(built):
    """The certificate of the pre-filter relies on the GEMM's tile schedule: tile (n, m) runs on CTA (n*s + m) mod G with
    s >= mt coprime to G.  Checked here on the host: the stride the library picks, exact-once coverage, balance, and that
    every CTA (producer) sees rows of every query tile once nt >= G -- for the M-tile counts of the bench (98), one image
    (7), the re-weighting launch (1), the CTA-pair launches (49 pairs on 74 units) and awkward gcds."""
    from math import gcd
    from cmdiad_b200 import _lib as L
    lib = L.load()
    for mt, nt, G in [(98, 782, 148), (7, 782, 148), (1, 782, 148), (74, 157, 148), (148, 300, 148), (296, 40, 148),
                      (49, 782, 74), (37, 200, 74), (3, 5, 148), (64, 782, 148)]:
        s = lib.cmdb_debug_tile_stride(mt, G)
        assert s >= mt and gcd(s % G or G, G) == 1 and s < mt + G
        owner = {}
        per_cta = [0] * G
        for n in range(nt):
            for m in range(mt):
                c = (n * s + m) % G
                owner[(n, m)] = c
                per_cta[c] += 1
        assert len(owner) == nt * mt
        assert max(per_cta) - min(per_cta) <= mt // G + 1 + (nt % G != 0) * (mt // G + 1)
        if nt >= G:
            for m in (0, mt // 2, mt - 1):
                assert len({owner[(n, m)] for n in range(nt)}) == G, (mt, nt, G, m)
        # the device iterator walks n in order and steps m by G inside a tile row: same tiles, same order
        smod = s % G
        for c in (0, 1, G // 2, G - 1):
            got, n, m, base = [], -1, mt, (c + smod) % G
            while True:
                m += G
                done = False
                while m >= mt:
                    n += 1
                    if n >= nt:
                        done = True
                        break
                    base -= smod
                    if base < 0:
                        base += G
                    m = base
                if done:
                    break
                got.append((n, m))
            assert got == sorted(k for k, v in owner.items() if v == c)


def test_fallback_tier_rule(built):
    """uncertified (query, producer) pairs that fit the work list -> exact rescan (rigorous); beyond it -> 3-term GEMM over
    the uncertified queries"""
    from cmdiad_b200 import _lib as L
    lib = L.load()
    assert lib.cmdb_debug_fallback_use_rescan(75, 75) == 1           # the bench's steady state
    assert lib.cmdb_debug_fallback_use_rescan(0, 0) == 1
    assert lib.cmdb_debug_fallback_use_rescan(600, 700) == 1         # one pair per query: rescan is cheaper per query
    assert lib.cmdb_debug_fallback_use_rescan(784, 40000) == 0       # near-duplicate banks: hundreds of pairs per query
    assert lib.cmdb_debug_fallback_use_rescan(10, 9000) == 1         # many pairs per query, but they fit: stay rigorous
    assert lib.cmdb_debug_fallback_use_rescan(10, 17000) == 0        # beyond the work list
    prev = 1
    for pairs in range(0, 20000, 50):                                # monotone in the number of pairs
        cur = lib.cmdb_debug_fallback_use_rescan(500, pairs)
        assert cur <= prev
        prev = cur
