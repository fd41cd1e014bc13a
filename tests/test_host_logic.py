"""CPU: host-side logic that has no kernel behind it, restated next to the device code it mirrors."""
import numpy as np
import pytest


def k_range(M, K, bk, tile_rows, kb_per_split, sp, tm, band):
    """csrc/gemm.cu `k_range`: k-blocks [kb0, kb1) of split `sp` for the M tile `tm`, clipped to the anti-diagonal band lo <= m + k < hi."""
    kblocks = (K + bk - 1) // bk
    kb0 = sp * kb_per_split
    kb1 = min(kblocks, kb0 + kb_per_split)
    if band is not None:
        lo_b, hi_b = band
        m0 = tm * tile_rows
        m1 = min(M, m0 + tile_rows) - 1
        lo = max(0, lo_b - m1) // bk
        hi = (min(K, hi_b - m0) + bk - 1) // bk
        c0, c1 = max(kb0, lo), min(kb1, hi)
        if c1 > c0:
            kb0, kb1 = c0, c1
        else:
            kb1 = kb0 + 1
    return kb0, kb1


@pytest.mark.parametrize("T,tile_rows,bk,splits", [(1000, 128, 64, 1), (333, 128, 64, 1), (333, 256, 64, 3), (129, 128, 32, 2), (64, 128, 64, 1)])
def test_band_clipped_k_ranges_cover_the_band(T, tile_rows, bk, splits):
    """Both position-gradient products of the rel-pos backward (functional._FlashRelPosAttention.backward): A[m, k] != 0 only for
    T-1 <= m + k <= 2T-2.  Every non-zero element must fall into the k-blocks its M tile visits, over all splits exactly once, and no
    tile may end up with an empty range (the accumulator must be written)."""
    band = (T - 1, 2 * T - 1)
    for (M, K) in ((T, 2 * T - 1), (2 * T - 1, T)):       # d(q+v) = dBD p   and   dp = dBD^T (q+v)
        kblocks = (K + bk - 1) // bk
        per = (kblocks + splits - 1) // splits
        if (kblocks + per - 1) // per != splits:
            continue                                      # the host rejects split counts that leave empty splits
        m = np.arange(M)[:, None]
        k = np.arange(K)[None, :]
        nz = (m + k >= band[0]) & (m + k < band[1])
        visited = np.zeros((M, K), dtype=np.int32)
        for tm in range((M + tile_rows - 1) // tile_rows):
            rows = slice(tm * tile_rows, min(M, (tm + 1) * tile_rows))
            for sp in range(splits):
                kb0, kb1 = k_range(M, K, bk, tile_rows, per, sp, tm, band)
                assert 0 <= kb0 < kb1 <= kblocks
                visited[rows, kb0 * bk:min(K, kb1 * bk)] += 1
        assert (visited[nz] == 1).all()                   # every band element is multiplied exactly once
        assert visited.max() <= 1                          # and nothing twice (splits stay disjoint)
        assert visited.sum() <= 0.75 * M * K or T <= 2 * tile_rows   # the clip is worth something at real sizes
