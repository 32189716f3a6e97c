"""CPU model of the index logic of the input-row-stationary thin conv (conv_thin_kernel<.., STK = 1>,
csrc/pgk_conv_thin.cu).

The kernel cannot run here, but its addressing can be replayed at the level the hardware sees it: shared memory as an
array of 16-byte units (8 bf16 each), UMMA K-major no-swizzle operand descriptors as (start unit, leading-byte offset,
stride-byte offset), tensor memory as [lane][column].  The loops below use the SAME index expressions as the kernel --
the staging permutation of the packed weights into [k step][K half][ky = 2, 1, 0][Npad], the A start of a k step, the
first filter-row block / first accumulator block / ring wrap of an input row, which output row completes when, and
the epilogue handing every block back zeroed -- over two consecutive work units of one CTA, so that the tile counter
and the accumulator ring run across a unit boundary.  The result is compared with a direct 3x3 convolution."""
import numpy as np
import pytest

from test_atm_model import pack_thin

ROW_PIX = 136          # kRowPix


def umma_operand(mem, start, lbo, sbo, rows):
    """rows x 16 elements of a K-major, un-swizzled operand: 8-row core matrices of 8 x 16 bytes, the two K halves
    `lbo` units apart, consecutive core matrices `sbo` units apart (units of 16 bytes)."""
    out = np.zeros((rows, 16))
    for r in range(rows):
        for h in range(2):
            out[r, 8 * h:8 * h + 8] = mem[start + h * lbo + (r // 8) * sbo + (r % 8)]
    return out


@pytest.mark.parametrize('cin,cout', [(8, 8), (8, 16), (8, 64), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64)])
def test_input_row_stationary_issue_order(cin, cout):
    rng = np.random.default_rng(cin * 100 + cout)
    RC, UNITS, W = 10, 2, 128          # two units of RC output rows each: one 128-pixel strip of a 2*RC-row image
    H = RC * UNITS
    CG, npad = cin // 8, max(cout, 16)
    NACC = 4 if npad == 64 else 8
    x = rng.integers(-3, 4, size=(H, W, cin)).astype(np.float64)
    w = rng.integers(-2, 3, size=(9 * cin, cout)).astype(np.float64)      # [(tap, ci)][co], tap = ky * 3 + kx
    wp = pack_thin(w, cin, cout, npad)                                    # [step][khalf][n][8]
    steps = wp.shape[0]
    ksteps = steps // 3

    # ---- weight staging (the permutation of the one-time setup) ----
    wsm = np.full((steps * 2 * npad, 8), np.nan)
    for i in range(steps * 2 * npad):
        n, rr = i % npad, i // npad
        h, st = rr & 1, rr >> 1
        if cin == 8:
            dy, ks = st >> 1, st & 1
        else:
            tap, cgp = st // (cin // 16), st % (cin // 16)
            dy, ks = tap // 3, (tap % 3) * (cin // 16) + cgp
        wsm[(((ks * 2 + h) * 3 + (2 - dy)) * npad) + n] = wp[st, h, n]
    assert not np.isnan(wsm).any()

    def row_buffer(y):
        """what the producer leaves in a ring slot for image row y: [cg][136 pixels][8], pixel 0 = x0 - 1, zero fill"""
        buf = np.zeros((CG * ROW_PIX, 8))
        if 0 <= y < H:
            for cg in range(CG):
                buf[cg * ROW_PIX + 1:cg * ROW_PIX + 129] = x[y, :, cg * 8:cg * 8 + 8]
        return buf

    tmem = np.zeros((128, NACC * npad))          # zeroed once at kernel start
    out = np.full((H, W, npad), np.nan)
    ti = 0
    for u in range(UNITS):
        ya = u * RC
        for j in range(RC + 2):
            rows = row_buffer(ya - 1 + j)
            i_lo, i_hi = (j - 2 if j >= 2 else 0), (j if j < RC else RC - 1)
            nb = i_hi - i_lo + 1
            kb0 = 2 - j + i_lo
            blk0 = (ti + i_lo) % NACC
            nb_a = nb if blk0 + nb <= NACC else NACC - blk0
            for ks in range(ksteps):
                if cin == 8:
                    aoff, a_lbo = ks * 2, 1
                else:
                    dx, cgp = ks // (cin // 16), ks % (cin // 16)
                    aoff, a_lbo = 2 * cgp * ROW_PIX + dx, ROW_PIX
                a = umma_operand(rows, aoff, a_lbo, 8, 128)
                bstart = ks * 6 * npad + kb0 * npad
                b = umma_operand(wsm, bstart, 3 * npad, 8, nb_a * npad)
                tmem[:, blk0 * npad:(blk0 + nb_a) * npad] += a @ b.T
                if nb_a < nb:
                    b = umma_operand(wsm, bstart + nb_a * npad, 3 * npad, 8, (nb - nb_a) * npad)
                    tmem[:, 0:(nb - nb_a) * npad] += a @ b.T
            if j >= 2:          # output row j - 2 is complete: the epilogue reads its block and hands it back zeroed
                blk = (ti + j - 2) % NACC
                out[ya + j - 2] = tmem[:, blk * npad:(blk + 1) * npad]
                tmem[:, blk * npad:(blk + 1) * npad] = 0.0
        ti += RC
    ref = np.zeros((H, W, cout))
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    for ky in range(3):
        for kx in range(3):
            ref += xp[ky:ky + H, kx:kx + W, :] @ w[(ky * 3 + kx) * cin:(ky * 3 + kx + 1) * cin, :]
    assert np.array_equal(out[:, :, :cout], ref)
    assert not out[:, :, cout:].any()
    assert not tmem.any(), 'every block is handed back zeroed'
