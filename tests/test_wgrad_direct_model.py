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
