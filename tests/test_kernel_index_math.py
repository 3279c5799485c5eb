"""CPU checks of index arithmetic the CUDA kernels rely on, restated in numpy / plain Python.

1. TileWalk (conv_tc.cu): the division-free stepping of a persistent CTA through its tiles.
2. conv3x3_halo_kernel (busca_b200/csrc/conv_tc.cu): one (R+2) x (W+2) halo box per
tile, output rows = FLAT positions f = ho*(W+2) + wo of the halo grid, tap (r, q) = the same box shifted by r*(W+2) + q rows.
The numpy emulation below follows the kernel's loops (including the junk rows and the dense staging tile) and must equal a
direct zero-padded 3x3 convolution."""
import numpy as np
import pytest


def tile_walk(block, grid, tiles_n, h_tiles, total):
    """TileWalk of conv_tc.cu, line for line."""
    tile = block
    n_tile = tile % tiles_n
    m_tile = tile // tiles_n
    hi, ni = m_tile % h_tiles, m_tile // h_tiles
    qn, rn = grid // tiles_n, grid % tiles_n
    qh, qi = qn % h_tiles, qn // h_tiles
    while tile < total:
        yield tile, n_tile, hi, ni
        tile += grid
        n_tile += rn
        e = 0
        if n_tile >= tiles_n:
            n_tile -= tiles_n
            e = 1
        hi += qh + e
        ni += qi
        if hi >= h_tiles:
            hi -= h_tiles
            ni += 1


@pytest.mark.parametrize("grid", [1, 3, 7, 64, 144, 148])
@pytest.mark.parametrize("tiles_n", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("h_tiles", [1, 2, 3, 24, 96])
def test_tile_walk_equals_division(grid, tiles_n, h_tiles):
    tiles_m = 5 * h_tiles + 3                     # ragged: the last image group is partial
    total = tiles_m * tiles_n
    for block in {0, 1, grid // 2, grid - 1}:
        seen = 0
        for tile, n_tile, hi, ni in tile_walk(block, grid, tiles_n, h_tiles, total):
            m_tile = tile // tiles_n
            assert (n_tile, hi, ni) == (tile % tiles_n, m_tile % h_tiles, m_tile // h_tiles), (tile, grid, tiles_n, h_tiles)
            seen += 1
        assert seen == len(range(block, total, grid))


def halo_geometry(H, W):
    P = W + 2
    R = next(r for r in range(128 // P, 0, -1) if H % r == 0)
    return P, R


@pytest.mark.parametrize("H,W", [(96, 32), (48, 16), (24, 8)])
def test_halo_flat_shift_equals_conv3x3(H, W):
    rng = np.random.default_rng(H)
    cin, cout, N = 8, 5, 2
    x = rng.standard_normal((N, H, W, cin))
    w = rng.standard_normal((cout, 3, 3, cin))
    # direct: zero padding 1, stride 1
    xp = np.zeros((N, H + 2, W + 2, cin))
    xp[:, 1:-1, 1:-1] = x
    ref = np.zeros((N, H, W, cout))
    for r in range(3):
        for q in range(3):
            ref += xp[:, r:r + H, q:q + W] @ w[:, r, q].T
    P, R = halo_geometry(H, W)
    halo_rows, stage_rows = (R + 2) * P, 208
    assert R * P <= 128 and halo_rows <= 192 and 2 * P + 2 + 128 <= stage_rows       # the kernel's launch checks
    out = np.full((N, H, W, cout), np.nan)
    for n in range(N):
        for h0 in range(0, H, R):
            stage = rng.standard_normal((stage_rows, cin)) * 1e6                     # stale shared memory beyond the box
            box = xp[n, h0:h0 + R + 2, :, :]                                         # rows h0-1 .. h0+R, columns -1 .. W
            stage[:halo_rows] = box.reshape(halo_rows, cin)
            acc = np.zeros((128, cout))
            for r in range(3):
                for q in range(3):
                    s = r * P + q                                                    # descriptor start = s rows into the box
                    acc += stage[s:s + 128] @ w[:, r, q].T
            dense = np.full((R * W, cout), np.nan)
            for f in range(128):
                ho, wo = divmod(f, P)
                if ho < R and wo < W:
                    dense[ho * W + wo] = acc[f]
            out[n, h0:h0 + R] = dense.reshape(R, W, cout)
    assert np.allclose(out, ref, rtol=1e-9, atol=1e-9)
