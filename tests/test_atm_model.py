"""CPU model of the index logic of the two tensor-memory flavours written without a GPU (PGK_THIN_ATM, PGK_WTHIN_ATM).

The kernels cannot run here, but their addressing can be replayed: numpy arrays stand for the shared-memory row buffers
([pixel][8 channels] per channel group, one halo pixel on either side), for the tensor-memory slots (128 lanes x 4
columns = 8 bf16 per lane) and for the packed weights (pack_thin_kernel's order), and the loops below use the SAME
index expressions as conv_thin_kernel<.., ATM = 1> / wgrad_thin_kernel<8, .., ATM = 1> (csrc/pgk_conv_thin.cu,
csrc/pgk_wgrad_thin.cu) -- slot of a (dx, channel group), ring row of an input row, K = 16 step -> (dy, slot), lane ->
(kx, ci) of the gather.  The result is compared with a direct 3x3 convolution / weight-gradient correlation.  What the
model cannot check is the hardware's reading of the copy descriptor and of K inside a tensor-memory column; that is the
job of tools/probes/umma_probe.cu (parts 3 and 4)."""
import numpy as np
import pytest


def pack_thin(w, cin, cout, npad):
    """pack_thin_kernel: out[step][khalf][n][e] from w[(tap * cin + ci)][cout]."""
    steps = 6 if cin == 8 else 9 * (cin // 16)
    out = np.zeros((steps, 2, npad, 8))
    for st in range(steps):
        for h in range(2):
            for e in range(8):
                if cin == 8:
                    dy, q = st >> 1, st & 1
                    dx = q * 2 + h
                    k = (dy * 3 + dx) * 8 + e if dx < 3 else -1
                else:
                    per_tap = cin // 16
                    tap, cgp = st // per_tap, st % per_tap
                    k = tap * cin + cgp * 16 + h * 8 + e
                if k >= 0:
                    out[st, h, :cout, e] = w[k, :]
    return out


@pytest.mark.parametrize('cin,cout', [(8, 8), (8, 16), (16, 16), (16, 32), (32, 32), (32, 64)])
def test_thin_conv_tensor_memory_ring_indexing(cin, cout):
    rng = np.random.default_rng(cin * 100 + cout)
    H, W = 7, 128                      # one 128-pixel strip, RC = H output rows
    CG, npad = cin // 8, max(cout, 16)
    x = rng.integers(-3, 4, size=(H, W, cin)).astype(np.float64)
    w = rng.integers(-2, 3, size=(9 * cin, cout)).astype(np.float64)      # [(tap, ci)][co], tap = ky * 3 + kx
    wp = pack_thin(w, cin, cout, npad)
    steps = wp.shape[0]
    SLOTS = 4 if cin == 8 else 3 * CG
    ROW_COLS = 4 * SLOTS
    ring = np.zeros((4, 128, ROW_COLS * 2))      # [ring row][lane][bf16 element] (2 elements per 32-bit column)

    def raw_row(j):
        """shared-memory row buffer of input row (ya - 1 + j): [cg][130 pixels][8], zero outside the image"""
        y = j - 1
        buf = np.zeros((CG, 130, 8))
        if 0 <= y < H:
            buf[:, 1:129, :] = x[y].reshape(W, CG, 8).transpose(1, 0, 2)
        return buf

    def copy_row(g):     # conv_thin_kernel<ATM>: copy_row
        buf = raw_row(g)
        dst = (g & 3)
        for dx in range(3):
            for cg in range(CG):
                col = (dx * CG + cg) * 4
                ring[dst, :, 2 * col:2 * col + 8] = buf[cg, dx:dx + 128, :]     # 128 pixels x 16 bytes from pixel dx on

    out = np.zeros((H, W, npad))
    g = 0
    copy_row(g)
    copy_row(g + 1)
    for i in range(H):
        copy_row(g + 2)
        acc = np.zeros((128, npad))
        for st in range(steps):
            if cin == 8:
                dy, slot = st >> 1, (st & 1) * 2
            else:
                tap, cgp = st // (cin // 16), st % (cin // 16)
                dy, slot = tap // 3, (tap % 3) * CG + 2 * cgp
            at = ((g + dy) & 3, slot * 4)
            a = ring[at[0], :, 2 * at[1]:2 * at[1] + 16]                       # K = 16: 8 columns
            b = np.concatenate([wp[st, 0], wp[st, 1]], axis=1)                   # [n][16]
            acc += a @ b.T
        out[i] = acc
        g += 1
    ref = np.zeros((H, W, cout))
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    for ky in range(3):
        for kx in range(3):
            ref += xp[ky:ky + H, kx:kx + W, :] @ w[(ky * 3 + kx) * cin:(ky * 3 + kx + 1) * cin, :]
    assert np.array_equal(out[:, :, :cout], ref)
    assert not out[:, :, cout:].any()


@pytest.mark.parametrize('cout', [8, 16, 32])
def test_stacked_weight_gradient_tensor_memory_gather(cout):
    rng = np.random.default_rng(cout)
    H, W, cin = 6, 128, 8
    x = rng.integers(-3, 4, size=(H, W, cin)).astype(np.float64)
    gy = rng.integers(-2, 3, size=(H, W, cout)).astype(np.float64)
    A = np.zeros((128, 128))            # tensor-memory A operand: [lane = slot * 32 + (kx * 8 + ci | 24 = ones)][pixel]
    acc = np.zeros((4, 128, cout))      # four rotating accumulators
    used = [False] * 4

    def raw_row(j):
        y = j - 1
        buf = np.zeros((130, 8))
        if 0 <= y < H:
            buf[1:129] = x[y]
        return buf

    def write_slot(gx):     # the owning warp's gather + tcgen05.st
        buf, b = raw_row(gx), gx & 3
        for lane in range(32):
            for p in range(128):
                if lane < 24:
                    A[b * 32 + lane, p] = buf[p + (lane >> 3), lane & 7]
                else:
                    A[b * 32 + lane, p] = 1.0 if lane == 24 else 0.0

    gx = 0
    write_slot(0)
    write_slot(1)
    for i in range(H):
        write_slot(gx + 2)
        c4 = gx & 3
        d = A @ gy[i]                    # one stacked chain: all four slots against G row i
        acc[c4] = d if not used[c4] else acc[c4] + d
        used[c4] = True
        gx += 1
    # flush: accumulator c, rows 32 * j .. = slot j = ky (j - c) & 3; row kx * 8 + ci, row 24 = ones
    dw = np.zeros((9 * cin, cout))
    db = np.zeros(cout)
    for warp in range(4):
        for c4 in range(4):
            ky = (warp - c4) & 3
            for lane in range(32):
                v = acc[c4, warp * 32 + lane]
                if ky < 3 and lane < 24:
                    dw[(ky * 3 + (lane >> 3)) * cin + (lane & 7)] += v
                elif ky == 1 and lane == 24:
                    db += v
    ref = np.zeros((9 * cin, cout))
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    for ky in range(3):
        for kx in range(3):
            ref[(ky * 3 + kx) * cin:(ky * 3 + kx + 1) * cin] = np.einsum('hwc,hwo->co', xp[ky:ky + H, kx:kx + W], gy)
    assert np.array_equal(dw, ref)
    assert np.array_equal(db, gy.sum(axis=(0, 1)))


@pytest.mark.parametrize('cin,cout', [(16, 16), (32, 32), (32, 64)])
def test_wide_channel_weight_gradient_tensor_memory_row_order(cin, cout):
    """Cin = 16 / 32 with PGK_WTHIN_ATM: one 128-lane A tile per input row, rows ordered (channel group, kx, channel) --
    lane quarter cg is written by warp cg from channel-group plane cg -- three accumulators D[ky], flush lane -> (cg, kx,
    ci)."""
    rng = np.random.default_rng(cin + cout)
    H, W, CG = 6, 128, cin // 8
    x = rng.integers(-3, 4, size=(H, W, cin)).astype(np.float64)
    gy = rng.integers(-2, 3, size=(H, W, cout)).astype(np.float64)
    tiles = np.zeros((4, 128, 128))         # [ring slot][lane][pixel], zeroed once
    acc = np.zeros((3, 128, cout))

    def raw_row(j):                          # [cg][130 pixels][8]
        y = j - 1
        buf = np.zeros((CG, 130, 8))
        if 0 <= y < H:
            buf[:, 1:129, :] = x[y].reshape(W, CG, 8).transpose(1, 0, 2)
        return buf

    def write_tile(gx):
        buf, b = raw_row(gx), gx & 3
        for warp in range(CG):
            for lane in range(32):
                for p in range(128):
                    if lane < 24:
                        tiles[b, warp * 32 + lane, p] = buf[warp, p + (lane >> 3), lane & 7]
                    else:
                        tiles[b, warp * 32 + lane, p] = 1.0 if (lane == 24 and warp == 0) else 0.0

    gx = 0
    write_tile(0)
    write_tile(1)
    for i in range(H):
        write_tile(gx + 2)
        for ky in range(3):
            acc[ky] += tiles[(gx + ky) & 3] @ gy[i]
        gx += 1
    dw = np.zeros((9 * cin, cout))
    db = np.zeros(cout)
    for warp in range(4):
        for lane in range(32):
            m = warp * 32 + lane
            valid = warp < CG and lane < 24
            kx, ci = lane >> 3, warp * 8 + (lane & 7)
            for ky in range(3):
                if valid:
                    dw[(ky * 3 + kx) * cin + ci] += acc[ky, m]
                elif ky == 1 and m == 24:
                    db += acc[ky, m]
    ref = np.zeros((9 * cin, cout))
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    for ky in range(3):
        for kx in range(3):
            ref[(ky * 3 + kx) * cin:(ky * 3 + kx + 1) * cin] = np.einsum('hwc,hwo->co', xp[ky:ky + H, kx:kx + W], gy)
    assert np.array_equal(dw, ref)
    assert np.array_equal(db, gy.sum(axis=(0, 1)))
