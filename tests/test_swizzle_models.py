"""CPU models of the shared-memory address maps added in round 2b (no GPU): the staging tiles of the thin conv's bulk
stores / coalesced mask loads (csrc/pgk_conv_thin.cu), the wide weight gradient's staged flush and its bias-gradient
reads (csrc/pgk_conv_tc.cu).  Each map must be a bijection onto its tile, agree with the hardware swizzle the tensor map
declares (Swizzle<B,4,3>: 16-byte chunk index ^= address bits 7..7+B-1), and be free of bank conflicts (a warp's
16-byte accesses complete in the minimum of four 128-byte wavefronts: every aligned group of 8 lanes covers all 32
banks exactly once)."""
import pytest


def banks16(addr):
    """the four 4-byte banks a 16-byte access at `addr` touches"""
    return {((addr >> 2) + k) & 31 for k in range(4)}


def conflict_free(addrs):
    """32 lane addresses of one 16-byte warp access: each quarter-warp (8 lanes = 128 bytes) hits 32 distinct banks"""
    for q in range(4):
        seen = set()
        for a in addrs[8 * q:8 * q + 8]:
            b = banks16(a)
            if seen & b:
                return False
            seen |= b
        if len(seen) != 32:
            return False
    return True


@pytest.mark.parametrize('cout', [16, 32, 64])
@pytest.mark.parametrize('base', [0, 1024, 5 * 1024, 37 * 1024])
def test_thin_conv_staging_tile(cout, base):
    cb = cout * 2                       # bytes of one pixel row of the tile
    nch = cb // 16
    mask = nch - 1                      # SWIZZLE_32B / 64B / 128B: 1 / 2 / 3 address bits
    for warp in range(8):
        stg = base + warp * 32 * cb

        def addr(lane, h):              # emit8: chunk h of the lane's pixel
            row = stg + lane * cb
            return row + ((h ^ ((row >> 7) & mask)) << 4)

        # (1) bijection onto the warp's tile, and the hardware pattern: linear offset with chunk bits ^= bits 7..
        seen = set()
        for lane in range(32):
            for h in range(nch):
                a = addr(lane, h)
                lin = stg + lane * cb + h * 16
                assert a == lin ^ (((lin >> 7) & mask) << 4)
                seen.add(a)
        assert seen == {stg + 16 * i for i in range(32 * nch)}
        # (2) the epilogue's stores (fixed chunk h across lanes) are conflict free
        for h in range(nch):
            assert conflict_free([addr(lane, h) for lane in range(32)])
        # (3) the coalesced mask path: lane writes chunk c = 32 k + lane of the warp's block where its pixel's thread
        # will read it; a pixel's thread then reads its own row
        for k in range(nch):
            ws = []
            for lane in range(32):
                c = 32 * k + lane
                row = stg + (c // nch) * cb
                ws.append(row + (((c % nch) ^ ((row >> 7) & mask)) << 4))
            assert conflict_free(ws)
            assert sorted(ws) == sorted(addr(c // nch, c % nch) for c in range(32 * k, 32 * k + 32))


def test_wgrad_flush_transposition():
    pitch = 144                                   # bytes: 36 floats per staged row
    # write: lane = row, 8 float4 per lane (column chunk j)
    for j in range(8):
        assert conflict_free([lane * pitch + 16 * j for lane in range(32)])
    # read: lane -> (row it*4 + lane//8, column chunk lane%8): 4 rows x 128 contiguous bytes per access
    cells = set()
    for it in range(8):
        addrs = [(it * 4 + (lane >> 3)) * pitch + 16 * (lane & 7) for lane in range(32)]
        assert conflict_free(addrs)
        cells |= {(it * 4 + (lane >> 3), lane & 7) for lane in range(32)}
    assert cells == {(r, q) for r in range(32) for q in range(8)}


@pytest.mark.parametrize('base', [0, 4096, 3 * 4096])
def test_wgrad_bias_reads_of_a_swizzled_g_box(base):
    # one G box: 32 pixel rows of 128 bytes (64 channels), SWIZZLE_128B; thread t of the four bias warps reads chunk
    # c = t & 7 of pixels pg = t >> 3 and pg + 16
    got = set()
    for h in range(2):
        for warp in range(4):
            addrs = []
            for lane in range(32):
                t = warp * 32 + lane
                c, px = t & 7, (t >> 3) + 16 * h
                a = base + px * 128 + ((c ^ (px & 7)) << 4)
                lin = base + px * 128 + c * 16
                assert a == lin ^ (((lin >> 7) & 7) << 4)
                addrs.append(a)
                got.add((px, c))
            assert conflict_free(addrs)
    assert got == {(px, c) for px in range(32) for c in range(8)}
