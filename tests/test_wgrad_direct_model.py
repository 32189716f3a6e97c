"""CPU model of the ring / mirror-slot / release logic of csrc/pgk_wgrad_direct.cu (no GPU): the producer runs as far
ahead of the MMA warp as the barriers allow, and every window the MMA warp reads must hold the three G rows of its
input row in three consecutive physical slots."""
import pytest


def simulate(RC, RG, units, greedy=True):
    phys = {}                      # physical slot -> (unit, seq within unit)
    released = set()               # global G sequence numbers whose slot has been handed back
    loaded = 0                     # global G sequence numbers loaded so far
    per_unit = RC + 2
    total = units * per_unit

    def can_load(t):
        return t < total and (t < RG or (t - RG) in released)

    def load(t):
        s = t % RG
        tag = (t // per_unit, t % per_unit)
        phys[s] = tag
        if s < 2:
            phys[RG + s] = tag

    checked = 0
    for u in range(units):
        g0 = u * per_unit
        for j in range(RC):
            need = g0 + j + 2
            # the producer: greedy = everything the barriers allow; lazy = just what this window needs
            while loaded <= (total - 1 if greedy else need) and can_load(loaded):
                load(loaded)
                loaded += 1
            assert loaded > need, 'deadlock: window %d of unit %d waits for a row the producer cannot load' % (j, u)
            ws = (g0 + j) % RG
            for k in range(3):
                assert phys[ws + k] == (u, j + k), (u, j, k, ws, phys)
            checked += 1
            released.add(g0 + j)
            if j == RC - 1:
                released.add(g0 + j + 1)
                released.add(g0 + j + 2)
    return checked


@pytest.mark.parametrize('RC', [8, 16, 32])
@pytest.mark.parametrize('RG', [4, 8])
@pytest.mark.parametrize('greedy', [True, False])
def test_windows_hold_their_rows(RC, RG, greedy):
    assert simulate(RC, RG, units=5, greedy=greedy) == 5 * RC


def test_m64_accumulator_lane_map():
    # tools/probes/mnmajor_probe.cu on B200: row m of an M = 64 accumulator sits in lane (m & 15) + 32 * (m >> 4);
    # the flush inverts it as m = warp * 16 + lane for lane < 16
    for m in range(64):
        lane_abs = (m & 15) + 32 * (m >> 4)
        warp, lane = lane_abs // 32, lane_abs % 32
        assert lane < 16 and warp * 16 + lane == m


# ---------------------------------------------------------------------------------------------------------------
# Descriptor-level replay of one unit of csrc/pgk_wgrad_direct.cu on the CPU: shared memory as a byte array written the
# way TMA writes it (natural [pixel][channel] rows, hardware swizzle on the absolute address), operands fetched the way
# an MN-major tcgen05 descriptor addresses them (the reading tools/probes/mnmajor_probe.cu established on B200), the
# accumulator flushed through the kernel's row / column maps -- against the definition of the weight gradient.
# ---------------------------------------------------------------------------------------------------------------
import numpy as np


def _swz(addr, cb):
    mask = cb // 16 - 1 if cb > 16 else 0
    return addr ^ (((addr >> 7) & mask) << 4)


class _Smem(object):
    def __init__(self, size):
        self.v = np.zeros(size // 2, dtype=np.float64)      # one value per bf16 slot

    def tma_row(self, dst, cb, row):                        # row: [pixels][channels] -> natural layout, swizzled
        npx, c = row.shape
        for p in range(npx):
            for ch in range(c):
                self.v[_swz(dst + p * cb + ch * 2, cb) // 2] = row[p, ch]

    def operand(self, start, cb, mn_stride, groups, ksteps=16):
        """[groups * C][16] values an MN-major descriptor with this start address reads (C = cb / 2 channels per MN
        group, groups mn_stride bytes apart; K = pixels: 8-pixel core matrices cb * 8 bytes apart, rows cb apart)"""
        c = cb // 2
        out = np.zeros((groups * c, ksteps))
        for g in range(groups):
            for k in range(ksteps):
                for ch in range(c):
                    lin = start + g * mn_stride + (k // 8) * 8 * cb + (k % 8) * cb + ch * 2
                    out[g * c + ch, k] = self.v[_swz(lin, cb) // 2]
        return out


@pytest.mark.parametrize('cin,cout', [(8, 8), (16, 16), (32, 16), (16, 32), (64, 32), (32, 64)])
def test_direct_weight_gradient_replayed_at_descriptor_level(cin, cout):
    rng = np.random.RandomState(cin * 100 + cout)
    H, W, RC, RG, RX = 8, 128, 8, 4, 4
    X = rng.randint(-2, 3, size=(H, W, cin)).astype(np.float64)
    G = rng.randint(-2, 3, size=(H, W, cout)).astype(np.float64)
    cbx, cbg = 2 * cin, 2 * cout
    xslot, gslot = 136 * cbx, 128 * cbg
    xr0, gr0 = 0, ((RX * xslot + 1023) // 1024) * 1024
    sm = _Smem(gr0 + (RG + 2 + 8) * gslot)
    N = 3 * cin
    two = cout == 64
    M0 = 64 if cout <= 16 else 128
    acc0, acc1 = np.zeros((M0, N)), np.zeros((64, N))
    # producer order: G sequence j = image row j - 1 (zero outside the image), X row j - 2 after it
    def load_g(seq):
        y = seq - 1
        row = G[y] if 0 <= y < H else np.zeros((W, cout))
        s = seq % RG
        sm.tma_row(gr0 + s * gslot, cbg, row)
        if s < 2:
            sm.tma_row(gr0 + (RG + s) * gslot, cbg, row)

    def load_x(j):
        row = np.zeros((130, cin))
        row[1:129] = X[j]                                   # x halo: pixel p of the buffer is image x = p - 1
        sm.tma_row(xr0 + (j % RX) * xslot, cbx, row)

    load_g(0), load_g(1)
    for j in range(RC):
        load_g(j + 2)
        load_x(j)
        ws, xs = j % RG, j % RX
        for ks in range(8):
            b = sm.operand(xr0 + xs * xslot + ks * 16 * cbx, cbx, cbx, 3)                # N groups = pixel shifts
            a0 = sm.operand(gr0 + ws * gslot + ks * 16 * cbg, cbg, gslot, M0 // cout)    # M groups = ring slots
            acc0 += a0 @ b.T
            if two:
                a1 = sm.operand(gr0 + (ws + 2) * gslot + ks * 16 * cbg, cbg, gslot, 1)
                acc1 += a1 @ b.T
    # flush maps: row m = slot * Cout + co (ky = 2 - slot), column n = kx * Cin + ci
    dw = np.zeros((9 * cin, cout))
    for acc, slot0 in ((acc0, 0),) + (((acc1, 2),) if two else ()):
        for m in range(acc.shape[0]):
            slot, co = slot0 + m // cout, m % cout
            if slot >= 3:
                continue
            for n in range(N):
                kx, ci = n // cin, n % cin
                dw[((2 - slot) * 3 + kx) * cin + ci, co] += acc[m, n]
    ref = np.zeros((9 * cin, cout))
    Xp = np.zeros((H + 2, W + 2, cin))
    Xp[1:-1, 1:-1] = X
    for ky in range(3):
        for kx in range(3):
            patch = Xp[ky:ky + H, kx:kx + W]               # X[y + ky - 1][x + kx - 1]
            ref[(ky * 3 + kx) * cin:(ky * 3 + kx + 1) * cin] = np.einsum('yxc,yxo->co', patch, G)
    assert np.array_equal(dw, ref)
