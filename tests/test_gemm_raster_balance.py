"""Host-side check of the reasoning behind lp_gemm_set_raster_balance (leopard_b200/csrc/lp_gemm.cu, lp_gemm_common.cuh::tile_coords /
gemm_item): persistent units take the tiles `unit + i * n_units` of the raster index; a ragged last column block makes its tiles
cheaper, so the raster decides how evenly the cheap tiles spread over the units.  Restates tile_coords in Python (test infrastructure)."""


def tile_coords(tile, num_m_blk, num_n_blk, group_m):
    per_group = group_m * num_n_blk
    g, r = divmod(tile, per_group)
    m0 = g * group_m
    gsz = min(group_m, num_m_blk - m0)
    n_blk = r // gsz
    return m0 + (r - n_blk * gsz), n_blk


def unit_costs(M, N, block_n, n_units, group_m):
    num_m_blk = -(-M // 256)                      # CTA pairs: 256-row tiles
    num_n_blk = -(-N // block_n)
    last_w = ((N - (num_n_blk - 1) * block_n) + 31) // 32 * 32     # the ragged block runs a narrower MMA
    cost = [0.0] * n_units
    seen = set()
    for t in range(num_m_blk * num_n_blk):
        m, n = tile_coords(t, num_m_blk, num_n_blk, group_m)
        assert 0 <= m < num_m_blk and 0 <= n < num_n_blk and (m, n) not in seen
        seen.add((m, n))
        cost[t % n_units] += (last_w if n == num_n_blk - 1 else block_n) / block_n
    assert len(seen) == num_m_blk * num_n_blk
    return cost


def test_vit_fc2_ragged_tiles_spread_evenly_with_the_row_major_raster():
    # ViT fc2 at cfg3: M = 28 * 676, N = 1152 = 4 x 256 + 128, 74 CTA pairs; W (9.9 MB) fits L2, A (163 MB) does not
    M, N, units = 18928, 1152, 74
    old = unit_costs(M, N, 256, units, group_m=148 // 5)       # groups of SMs / column-blocks row blocks (the earlier choice)
    new = unit_costs(M, N, 256, units, group_m=1)              # row-major: every unit sees each column block equally often
    assert max(old) == 5.0 and abs(sum(old) / units - 4.5) < 1e-9   # 35 pairs got all half-width tiles, the rest ran five full ones
    assert max(new) == min(new) == 4.5
    # in row-major order the five tiles of a row block are neighbours in the raster, i.e. run at the same time on neighbouring units:
    # A is still read from HBM once
    assert [tile_coords(t, 74, 5, 1) for t in range(5)] == [(0, n) for n in range(5)]


def test_exact_column_blocks_are_unaffected():
    for N in (3840, 4096, 6144):
        c = unit_costs(18928 if N == 3840 else 4796, N, 256, 74, group_m=7)
        assert max(c) - min(c) <= 1.0
