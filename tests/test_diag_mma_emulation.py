"""CPU check of the index arithmetic of the tensor-pipe diagonal-block kernel (csrc/diag.cu:diag64_mma_kernel): the DMMA.8x8x4
fragment layouts, the tile dealing over warps, the parked G^T tiles and the row-wise inverse are replayed lane by lane in
numpy (tools/emulate_diag_mma.py) and compared with numpy's Cholesky / inverse."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emulator():
    spec = importlib.util.spec_from_file_location("emulate_diag_mma", os.path.join(ROOT, "tools", "emulate_diag_mma.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_lane_level_emulation_reproduces_cholesky_and_inverse():
    em = _emulator()
    rng = np.random.default_rng(3)
    X = rng.standard_normal((64, 200))
    A = X @ X.T / 200 + 0.05 * np.eye(64)
    Aout, W, WT, rinv, logdet = em.run(A)
    L = np.linalg.cholesky(A)
    low = np.tril(np.ones((64, 64), bool))
    assert np.all(np.isnan(Aout[~low]))                       # nothing above the diagonal is written back
    assert np.abs(Aout[low] - L[low]).max() < 1e-13
    assert np.abs(W - np.linalg.inv(L)).max() < 1e-11
    assert np.all(W[~low] == 0.0) and np.array_equal(WT, W.T)
    assert abs(logdet - np.linalg.slogdet(A)[1]) < 1e-12 * abs(logdet) + 1e-12


def test_trailing_tile_table_matches_the_kernel_source():
    em = _emulator()
    src = open(os.path.join(ROOT, "nonstationary_multivariate_gaussian_process_b200", "csrc", "diag.cu")).read()
    body = src[src.index("kTrailTile[28] = {") + len("kTrailTile[28] = {"):]
    body = body[:body.index("}")]
    table = [int(t, 16) for t in body.replace("\n", " ").split(",")]
    assert table == em.TRAIL
    for p in range(7):
        m = 7 - p
        off, nt = 8 * p - p * (p + 1) // 2, m * (m + 1) // 2
        tiles = [(e >> 4, e & 15) for e in table[off:off + nt]]
        assert off + nt == 28 and tiles[0] == (p + 1, p + 1)
        assert sorted(tiles) == sorted((i, j) for j in range(p + 1, 8) for i in range(j, 8))
