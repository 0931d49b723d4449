"""CPU tests (-m "not gpu") of the HOST LOGIC of the kernel body.

pomcpp_b200/csrc/pom_core.cuh (packed record, explicit-stack explosion machine, loop-form chain
reversion, AoS<->record converters) is compiled for the host by tests/hostsim (test-only) and
compared with the oracle: the reference's known-answer scenarios, the golden transitions and long
random traces where the records stay PACKED across ticks (so slot indirection, stale ring slots
and signed counters have to survive many ticks), every field compared after every tick.
"""
import os

import numpy as np
import pytest

import oracle
import scenarios
from hostsim import HostSim, HostSimBackend

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def hs():
    return HostSim()


@pytest.mark.parametrize("fn", scenarios.STEP_SCENARIOS, ids=lambda f: f.__name__)
def test_scenarios_core(orc, fn):
    fn(HostSimBackend(orc))


def test_golden_scenarios_core(orc, hs):
    g = np.load(os.path.join(GOLD, "scenarios.npz"))
    before = g["before"].copy().view(oracle.STATE_DT).reshape(-1)
    after = g["after"].copy().view(oracle.STATE_DT).reshape(-1)
    recs, bad = hs.pack(before)
    assert not bad.any()
    rt, _ = hs.unpack(recs)
    assert orc.diff_batch(rt, before)[0] == -1, "pack/unpack round trip"
    hs.step_records(recs, np.ascontiguousarray(g["moves"]), raw=True)
    out, _ = hs.unpack(recs)
    e, why = orc.diff_batch(out, after)
    assert e == -1, "transition %d (%s) differs in field group %d" % (e, g["names"][e], why)


def test_pack_rejects_unrepresentable(orc, hs):
    s = orc.zero_state(4)
    s["board"][0, 3, 3] = 4 << 16 | (17 << 3)       # flame id 17 without a queue entry
    s["board"][1, 0, 0] = (2 << 8) + 9               # wood with an impossible flag
    s["agents"][2, 1]["x"] = 11                      # agent out of bounds
    recs, bad = hs.pack(s)
    assert bad[0] and bad[1] and bad[2] and not bad[3]
    _, status = hs.unpack(recs)
    assert (status[:3] & 0x10).all() and status[3] == 0


def _trace(orc, hs, n, ticks, nact, stress, seed, restart=True, ring_start=None, count_invalid=None):
    seeds = oracle.clean_seeds(128)
    B = orc.zero_state(n)
    for e in range(n):
        assert orc.init_state(B[e:e + 1], seeds[e % len(seeds)]) == 0
    if stress == 1:
        B["agents"]["canKick"] = 1
        B["agents"]["maxBombCount"] = 5
        B["agents"]["bombStrength"] = 4
    elif stress == 2:
        # mixed: some agents kick, some do not (they step onto bombs and are bounced back), several bombs each
        rng = np.random.default_rng(seed)
        B["agents"]["canKick"] = rng.integers(0, 2, (n, 4))
        B["agents"]["maxBombCount"] = rng.integers(1, 5, (n, 4))
        B["agents"]["bombStrength"] = rng.integers(1, 4, (n, 4))
    if ring_start is not None:
        # empty rings that start near the end of the physical array: every queue operation wraps
        # (FixedQueue index arithmetic, reference general_test.cpp:41-61 "Index 5 / Index 2")
        B["bombs_index"] = (np.arange(n) + ring_start) % 20
        B["flames_index"] = (np.arange(n) * 7 + ring_start) % 20
    B0 = B.copy()
    recs, bad = hs.pack(B)
    assert not bad.any()
    recs0 = recs.copy()
    sb = np.zeros(n, np.uint8)
    fo = np.zeros(n, np.uint8)
    fc = np.zeros(n, np.uint8)
    steps = 0
    for t in range(ticks):
        mv = orc.rng_moves(seed, 0, n, t, nact)
        orc.env_step_batch(B, sb, mv, fo)
        hs.step_records(recs, mv, raw=False, flags=fc)
        out, sc = hs.unpack(recs)
        e, why = orc.diff_batch(out, B)
        assert e == -1, "tick %d env %d field group %d" % (t, e, why)
        assert (sc == sb).all(), "status differs at tick %d" % t
        assert ((fo & 0x1F) == (fc & 0x1F)).all(), "flags differ at tick %d" % t
        fin = (sb & 0x11) != 0
        steps += int((~fin).sum())
        if count_invalid is not None:
            count_invalid[0] += int(((sb & 0x10) != 0).sum())
            count_invalid[1] += int(((fo & 0x22) != 0).sum())
        if restart:
            idx = np.nonzero(fin)[0]
            B[idx] = B0[idx]
            recs[idx] = recs0[idx]
            sb[idx] = 0
    return steps


def test_core_random(orc, hs):
    assert _trace(orc, hs, 2048, 100, 6, 0, 2001) > 150000


def test_core_harmless(orc, hs):
    assert _trace(orc, hs, 1024, 200, 5, 0, 2002) > 150000


def test_core_stress(orc, hs):
    assert _trace(orc, hs, 2048, 250, 6, 1, 2003) > 300000


def test_core_mixed_kickers(orc, hs):
    assert _trace(orc, hs, 2048, 250, 6, 2, 2005) > 300000


def test_core_ring_wraparound(orc, hs):
    assert _trace(orc, hs, 2048, 150, 6, 1, 2004, ring_start=15) > 200000


@pytest.fixture()
def hs_rays():
    """the host build with the tick decomposed the way the warp-cooperative kernels run it (pomcore::step_by_rays):
    explosions scanned and committed ray by ray (serial machine only when a ray meets a bomb), flame pops arm by arm"""
    h = HostSim()
    h.set_by_rays(True)
    yield h
    h.set_by_rays(False)


def test_core_by_rays_random(orc, hs_rays):
    assert _trace(orc, hs_rays, 2048, 100, 6, 0, 2011) > 150000


def test_core_by_rays_stress(orc, hs_rays):
    assert _trace(orc, hs_rays, 2048, 250, 6, 1, 2013) > 300000


def test_core_by_rays_mixed_kickers(orc, hs_rays):
    assert _trace(orc, hs_rays, 2048, 250, 6, 2, 2015) > 300000


def test_core_by_rays_ring_wraparound(orc, hs_rays):
    assert _trace(orc, hs_rays, 2048, 150, 6, 1, 2014, ring_start=15) > 200000


@pytest.mark.parametrize("fn", scenarios.STEP_SCENARIOS, ids=lambda f: f.__name__)
def test_scenarios_core_by_rays(orc, fn):
    b = HostSimBackend(orc)
    b.h.set_by_rays(True)
    try:
        fn(b)
    finally:
        b.h.set_by_rays(False)


def test_core_continue_undefined(orc, hs):
    """POM_STEP_CONTINUE_UNDEFINED: D3 (kicker onto a BOMB cell without queue entry) and D5 (reversion chain reaches an
    agent that did not move) keep the env running with the canonical result; kernel body == restatement on the stress
    regime, where such ticks are frequent - and now no env is lost to them"""
    hs.set_continue_undefined(True)
    orc.set_continue_undefined(True)
    try:
        inv = [0, 0]
        assert _trace(orc, hs, 4096, 300, 6, 1, 2021, count_invalid=inv) > 500000
        assert inv[1] > 20, "the regime must actually hit D3 / D5 ticks"
        inv2 = [0, 0]
        _trace(orc, hs, 2048, 200, 6, 2, 2022, count_invalid=inv2)
    finally:
        hs.set_continue_undefined(False)
        orc.set_continue_undefined(False)
    # the same run without the flag loses envs to those ticks
    inv3 = [0, 0]
    _trace(orc, hs, 4096, 300, 6, 1, 2021, count_invalid=inv3)
    assert inv3[0] > inv[0]


def test_core_rng_matches_oracle(orc, hs):
    for env in (0, 1, 77, 2 ** 33 + 5):
        for tick in (0, 1, 799):
            for na in (5, 6):
                assert hs.lib.hostsim_rng_moves(99, env, tick, na) == orc.lib.pom_oracle_rng_moves(99, env, tick, na)


def test_time_step_does_not_wrap(orc):
    """timeStep is 16 bits in the record: the tick that would pass 65535 marks the env INVALID instead of wrapping"""
    from hostsim import HostSim
    hs = HostSim()
    s = orc.zero_state()
    orc.init_state(s)
    s["timeStep"] = 65534
    recs, bad = hs.pack(s)
    assert not bad.any()
    idle = np.zeros((1, 4), np.uint8)
    hs.step_records(recs, idle, False)
    out, st = hs.unpack(recs)
    assert out["timeStep"][0] == 65535 and st[0] == 0
    hs.step_records(recs, idle, False)
    out, st = hs.unpack(recs)
    assert out["timeStep"][0] == 65535 and (st[0] & 0x10)
    hs.step_records(recs, idle, False)                 # frozen from now on
    out2, st2 = hs.unpack(recs)
    assert orc.diff_batch(out, out2)[0] == -1 and st2[0] == st[0]


def test_powerup_counters_do_not_wrap(orc):
    """maxBombCount / bombStrength are bytes in the record: the pickup that would pass 255 marks the env INVALID"""
    from hostsim import HostSim
    hs = HostSim()
    for item, field in ((6, "maxBombCount"), (7, "bombStrength")):
        s = orc.zero_state()
        orc.kill(s, 1, 2, 3)
        orc.put_agent(s, 0, 0, 0)
        orc.put_item(s, 1, 0, item)
        s["agents"][field][0, 0] = 255
        recs, bad = hs.pack(s)
        assert not bad.any()
        hs.step_records(recs, np.array([[4, 0, 0, 0]], np.uint8), False)
        _, st = hs.unpack(recs)
        assert st[0] & 0x10
        s["agents"][field][0, 0] = 254
        recs, _ = hs.pack(s)
        hs.step_records(recs, np.array([[4, 0, 0, 0]], np.uint8), False)
        out, st = hs.unpack(recs)
        assert not (st[0] & 0x10) and out["agents"][field][0, 0] == 255


def test_stuck_flame_queue_does_not_wrap(orc):
    """a front flame with a negative timer never pops (reference: TickFlames tests `== 0`); its byte-sized timer must not
    wrap around into a positive one"""
    from hostsim import HostSim
    hs = HostSim()
    s = orc.zero_state()
    orc.kill(s, 2, 3)
    orc.put_agent(s, 10, 10, 0)
    orc.put_agent(s, 0, 10, 1)
    orc.spawn_flame(s, 3, 3, 1)
    s["flames"]["timeLeft"][0, 0] = -10
    ref_copy = s.copy()
    recs, bad = hs.pack(s)
    assert not bad.any()
    idle = np.zeros((1, 4), np.uint8)
    st_o = np.zeros(1, np.uint8)
    for t in range(130):
        hs.step_records(recs, idle, False)
        out, st = hs.unpack(recs)
        if st[0] & 0x10:
            break
        orc.env_step_batch(ref_copy, st_o, idle)
        assert orc.diff_batch(out, ref_copy)[0] == -1, t
    assert (st[0] & 0x10) and out["flames"]["timeLeft"][0, 0] == -128 and t > 100
    # timers outside [-16, 100] are refused at upload
    s["flames"]["timeLeft"][0, 0] = -17
    assert hs.pack(s)[1][0] != 0
