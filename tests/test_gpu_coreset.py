"""GPU parity tests (call through the C ABI): projection, canonical-order distance pass, greedy coreset loop."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(built):
    from cmdiad_b200 import Bank, coreset_rownorms
    from cmdiad_b200 import _lib as L
    from oracle import restate as O
    assert torch.cuda.is_available()
    return dict(Bank=Bank, rownorms=coreset_rownorms, L=L, O=O)


@pytest.mark.parametrize("d", [32, 64, 100, 127, 128, 129, 130, 131, 198, 221, 267, 301, 329, 341, 375, 768])
def test_rownorms_bit_exact_vs_oracle_and_torch_cuda(env, d):
    """one distance pass of features.py:405: CUDA kernel == C oracle == torch.linalg.norm on CUDA, bit for bit"""
    g = np.random.Generator(np.random.PCG64(1000 + d))
    n = 4099
    z = (g.standard_normal((n, d)) * 1.7).astype(np.float16)
    last = z[17].copy()
    got = env["rownorms"](z, last)
    assert (got.view(np.uint16) == env["O"].rownorms_restated(z, last).view(np.uint16)).all()
    zt = torch.from_numpy(z).cuda()
    ref = torch.linalg.norm(zt - zt[17:18], dim=1, keepdims=True).cpu().numpy()[:, 0]
    assert (got.view(np.uint16) == ref.view(np.uint16)).all(), "ATen CUDA reduction order changed?"
    z64 = g.standard_normal((n, d))
    got64 = env["rownorms"](z64, z64[5])
    assert (got64 == env["O"].rownorms_restated(z64, z64[5])).all()
    # float64 ('TF32' mode): torch-CUDA's own double reduction differs from the canonical order in the last ulp for a
    # fraction of the rows (association not pinned); the selected indices are insensitive to it, which
    # test_coreset_golden_case_all_modes checks free-running against torch
    zt = torch.from_numpy(z64).cuda()
    ref64 = torch.linalg.norm(zt - zt[5:6], dim=1, keepdims=True).cpu().numpy()[:, 0]
    np.testing.assert_allclose(got64, ref64, rtol=4e-16)


@pytest.mark.parametrize("N,D", [(3000, 768), (777, 1152), (500, 1920), (33, 64)])
def test_projection_bit_exact(env, N, D):
    O = env["O"]
    g = np.random.Generator(np.random.PCG64(N + D))
    x = g.standard_normal((N, D), dtype=np.float32)
    eps = 0.9 if D > 64 else 0.99
    try:
        csr = O.sparse_components(N, D, eps, 0)
    except ValueError:
        pytest.skip("d' > D for this shape")
    b = env["Bank"](D, N)
    b.append(x)
    z = b.project(csr)
    assert (z == O.project_restated(x, *csr)).all()
    b.close()


def test_projection_generic_csr(env):
    """non-uniform magnitudes take the generic (float64 data) path"""
    O = env["O"]
    g = np.random.Generator(np.random.PCG64(4))
    x = g.standard_normal((300, 128), dtype=np.float32)
    indptr, indices, data, d = O.sparse_components(300, 128, 0.99, 1) if False else (None, None, None, None)
    dproj = 40
    nnz_per = 9
    indices = np.concatenate([g.choice(128, nnz_per, replace=False) for _ in range(dproj)]).astype(np.int32)
    indptr = (np.arange(dproj + 1) * nnz_per).astype(np.int32)
    data = g.standard_normal(dproj * nnz_per)
    b = env["Bank"](128, 300)
    b.append(x)
    assert (b.project((indptr, indices, data, dproj)) == O.project_restated(x, indptr, indices, data, dproj)).all()
    b.close()


def _bank_with(env, lib):
    b = env["Bank"](lib.shape[1], lib.shape[0])
    b.append(lib)
    return b


def test_coreset_golden_case_all_modes(env, golden):
    """N=7840, d'=221, 784 picks: CUDA == oracle (both modes) == reference golden (float64 mode) == torch-CUDA
    literal restatement of the reference loop (FP16 mode, free-running)"""
    O, L = env["O"], env["L"]
    g = golden["rgb_case"]
    lib = cases.rgb_normalised_lib(golden)
    csr = O.sparse_components(lib.shape[0], lib.shape[1], 0.9, 0)
    n = int(0.1 * lib.shape[0])
    b = _bank_with(env, lib)
    z = O.project_restated(lib, *csr)
    idx64 = b.coreset_select(n, csr, L.CORESET_FP64)
    assert (idx64 == g["coreset_idx_TF32"]).all()
    assert (idx64 == O.coreset_restated(z, n, "TF32")).all()
    idx16, mn = b.coreset_select(n, csr, L.CORESET_FP16, return_min=True)
    ref16, st = O.coreset_restated(z, n, "FP16", return_state=True)
    assert (idx16 == ref16).all()
    assert (mn.view(np.uint16) == st["min_last"].view(np.uint16)).all()
    lit = O.coreset_torch_literal(torch.from_numpy(z), n, "FP16", device="cuda").numpy()
    assert (idx16 == lit).all(), f"first divergence at pick {np.nonzero(idx16 != lit)[0][:1]}"
    lit64 = O.coreset_torch_literal(torch.from_numpy(z), n, "TF32", device="cuda").numpy()
    assert (idx64 == lit64).all()
    # teacher forcing against the CPU-reference FP16 golden: only near-ties may differ
    forced = b.coreset_select(n, csr, L.CORESET_FP16, force_idx=g["coreset_idx_FP16"])
    assert (forced == O.coreset_restated(z, n, "FP16", force_idx=g["coreset_idx_FP16"])).all()
    assert (forced != g["coreset_idx_FP16"]).mean() < 0.02
    b.close()


@pytest.mark.parametrize("N,d,n", [(1001, 100, 64), (5000, 128, 200), (2500, 130, 100), (4097, 131, 150),
                                   (148 * 16 * 4 + 3, 301, 60), (300, 267, 300), (50, 64, 1)])
def test_coreset_shapes_no_projection(env, N, d, n):
    """d_proj = 0 path (features.py:369-370) on raw banks of width d... exercised through a wide identity-free bank:
    the bank itself is the projected matrix, so every alignment class / tail length is covered"""
    O, L = env["O"], env["L"]
    if d % 64 != 0:
        # bank dims must be multiples of 64: embed via an explicit CSR identity projection of width d
        D = (d + 63) // 64 * 64
        g = np.random.Generator(np.random.PCG64(N * 7 + d))
        x = g.standard_normal((N, D), dtype=np.float32)
        indptr = np.arange(d + 1, dtype=np.int32)
        indices = np.arange(d, dtype=np.int32)
        data = np.ones(d)
        csr = (indptr, indices, data, d)
        z = x[:, :d].astype(np.float64)
    else:
        g = np.random.Generator(np.random.PCG64(N * 7 + d))
        x = g.standard_normal((N, d), dtype=np.float32)
        csr = None
        z = x.astype(np.float64)
    b = _bank_with(env, x)
    for mode, name in ((L.CORESET_FP16, "FP16"), (L.CORESET_FP64, "TF32")):
        idx = b.coreset_select(n, csr, mode)
        assert idx[0] == 0
        assert (idx == O.coreset_restated(z, n, name)).all(), (name, N, d)
    b.close()


def test_coreset_duplicate_rows_tie_break(env):
    """exact ties: duplicated rows must resolve to the lowest index, like torch.argmax"""
    O, L = env["O"], env["L"]
    g = np.random.Generator(np.random.PCG64(8))
    base = g.standard_normal((200, 192), dtype=np.float32)
    x = np.concatenate([base, base, base], 0)
    b = _bank_with(env, x)
    idx = b.coreset_select(150, None, L.CORESET_FP16)
    assert (idx == O.coreset_restated(x.astype(np.float64), 150, "FP16")).all()
    assert (idx < 200).all()
    lit = O.coreset_torch_literal(torch.from_numpy(x.astype(np.float64)), 150, "FP16", device="cuda").numpy()
    assert (idx == lit).all()
    b.close()


def test_coreset_errors(env):
    L = env["L"]
    b = env["Bank"](128, 10)
    with pytest.raises(L.CmdbError):
        b.coreset_select(1, None)  # empty bank
    b.append(np.ones((10, 128), np.float32))
    with pytest.raises(L.CmdbError):
        b.coreset_select(11, None)
    with pytest.raises(L.CmdbError):
        b.append(np.ones((1, 128), np.float32))  # capacity
    b.close()


def test_stats_normalize_gather(env):
    O = env["O"]
    g = np.random.Generator(np.random.PCG64(2))
    x = (g.standard_normal((5000, 192)) * 3 + 0.7).astype(np.float32)
    b = _bank_with(env, x)
    mean, std, _, _ = b.stats()
    tm, ts = O.bank_stats_restated(torch.from_numpy(x))
    assert abs(mean - float(tm)) < 1e-6 * max(1, abs(float(tm))) and abs(std - float(ts)) < 1e-6 * float(ts)
    b.normalize(float(tm), float(ts))
    ref = O.normalize_restated(torch.from_numpy(x), tm, ts).numpy()
    assert (b.read().numpy() == ref).all()  # bit-exact float32
    idx = g.choice(5000, 300, replace=False)
    b.gather(idx)
    assert b.rows == 300 and (b.read().numpy() == ref[idx]).all()
    b.close()
