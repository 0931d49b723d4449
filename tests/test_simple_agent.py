"""CPU tests of the SimpleAgent policy.

* the oracle's restatement (oracle/pom_oracle_agent.c) against the reference's own [strategy] known-answer tests
  (unit_test/bboard/strategy_test.cpp) and against games recorded from the compiled reference
  (tests/golden/simple_agent.npz); where oracle/_ref exists, also move by move against the compiled reference;
* the device policy code (pomcpp_b200/csrc/pom_policy.cuh), compiled for the host by tests/hostsim, against the
  restatement on long games where records stay packed and agent memories persist.
"""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IDLE, UP, DOWN, LEFT, RIGHT, BOMB = range(6)
EXTRABOMB = 6


@pytest.fixture(scope="module", params=["restatement", "reference"])
def impl(request, orc):
    if request.param == "restatement":
        return orc
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libpomref.so not built (needs /root/reference)")
    return oracle.reference()


# ---- unit_test/bboard/strategy_test.cpp -------------------------------------------------------------------
def test_is_adjacent_enemy(impl):                     # :9-29
    s = impl.zero_state()
    impl.put_agent(s, 5, 5, 0)
    impl.put_agent(s, 4, 4, 1)
    assert impl.is_adjacent_enemy(s, 0, 2) and impl.is_adjacent_enemy(s, 0, 3)
    s = impl.zero_state()
    impl.put_agent(s, 5, 5, 0)
    impl.put_agent(s, 3, 2, 1)
    # agents 2 and 3 of a zeroed State sit at (0,0), 10 cells away
    for i in range(5):
        assert not impl.is_adjacent_enemy(s, 0, i)


def test_fill_rmap_never_reaches_rigid(impl, orc):     # :31-59 (the reference uses the dirty seed 0x13327; any clean one here)
    for seed in oracle.clean_seeds(6):
        s = impl.zero_state()
        impl.init_board_items(s, seed)
        impl.kill(s, 1, 2, 3)
        impl.put_agent(s, 0, 0, 0)
        m = impl.fill_rmap(s, 0)
        assert ((m & 0xFFFF)[s["board"][0] == 1] == 0).all()
        assert (m & 0xFFFF).any()


def test_move_towards_methods(impl):                   # :61-110
    def fresh():
        s = impl.zero_state()
        impl.init_board_items(s, 0x1337)
        return s
    s = fresh()
    impl.kill(s, 1, 2, 3)
    impl.put_agent(s, 4, 5, 0)
    assert impl.move_towards(s, 0, 0, 4, 1) == UP
    assert impl.move_towards(s, 0, 0, 3, 6) == DOWN
    assert impl.move_towards(s, 0, 0, 0, 10) == DOWN
    s = fresh()
    impl.kill(s, 1, 2, 3)
    impl.put_agent(s, 4, 5, 0)
    impl.put_item(s, 2, 6, EXTRABOMB)
    assert impl.move_towards(s, 0, 1, 2) == IDLE
    assert impl.move_towards(s, 0, 1, 3) == DOWN
    s = fresh()
    impl.kill(s, 2, 3)
    impl.put_agent(s, 4, 5, 0)
    impl.put_agent(s, 2, 6, 1)
    assert impl.move_towards(s, 0, 2, 2) == IDLE
    assert impl.move_towards(s, 0, 2, 3) == DOWN


def test_is_in_danger(impl):
    s = impl.zero_state()
    impl.plant_bomb(s, 5, 5, 0, True)
    assert impl.is_in_danger(s, 5, 5) == 10 and impl.is_in_danger(s, 6, 5) == 10 and impl.is_in_danger(s, 5, 4) == 10
    assert impl.is_in_danger(s, 7, 5) == 0 and impl.is_in_danger(s, 6, 6) == 0
    assert impl.is_in_danger(s, -1, 5) == 0 and impl.is_in_danger(s, 5, 11) == 0


# ---- recorded games of the unmodified reference agent ------------------------------------------------------
def _golden():
    g = np.load(os.path.join(GOLD, "simple_agent.npz"))
    init = g["init"].copy().view(oracle.STATE_DT).reshape(-1)
    return g, init, int(g["seed"]), int(g["ticks"])


def test_restatement_replays_the_reference_agents_games(orc):
    g, S, seed, ticks = _golden()
    n = S.shape[0]
    status = np.zeros(n, np.uint8)
    A = orc.simple_agents(n)
    for t in range(ticks):
        mv = np.zeros((n, 4), np.uint8)
        live = (status & 0x11) == 0
        assert (live == g["live"][t].astype(bool)).all()
        orc.simple_moves_batch(S, status, A, seed, 0, t, 15, mv)
        assert (mv[live] == g["moves"][t][live]).all(), "tick %d" % t
        orc.env_step_batch(S, status, mv)
    final = g["final"].copy().view(oracle.STATE_DT).reshape(-1)
    assert orc.diff_batch(S, final, g["excluded"])[0] == -1
    assert A.tobytes() == g["agents"].tobytes()
    # the games are not trivial: every kind of move occurs, bombs are laid and agents die
    hist = np.bincount(g["moves"][g["live"].astype(bool)].ravel(), minlength=6)
    assert (hist > 1000).all() and (status & 1).sum() > n // 2


@pytest.mark.parametrize("mask,stress", [(15, 0), (0b0101, 0), (15, 1)])
def test_restatement_vs_compiled_reference_agent(orc, ref, mask, stress):
    n, ticks, seed = 256, 160, 17 + mask
    seeds = oracle.clean_seeds(32)
    S = orc.zero_state(n)
    for i in range(n):
        orc.init_state(S[i:i + 1], seeds[i % 32])
    if stress:
        S["agents"]["canKick"] = 1
        S["agents"]["maxBombCount"] = 3
        S["agents"]["bombStrength"] = 4
    S2 = S.copy()
    st, st2 = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    A, B = orc.simple_agents(n), ref.simple_agents(n)
    for t in range(ticks):
        mv = orc.rng_moves(seed, 0, n, t, 6)
        mv2 = mv.copy()
        orc.simple_moves_batch(S, st, A, seed, 0, t, mask, mv)
        B.moves_batch(S2, st2, orc.rng_moves(seed, 0, n, t, 5), mask, mv2)
        live = (st & 0x11) == 0
        assert (mv[live] == mv2[live]).all(), "tick %d" % t
        assert A[live].tobytes() == B.export()[live].tobytes(), "agent memories, tick %d" % t
        fl = np.zeros(n, np.uint8)
        orc.env_step_batch(S, st, mv, fl)
        ref.env_step_batch(S2, st2, mv2, None, ((fl & 0x20) != 0).astype(np.uint8))
        st[(fl & 0x3E) != 0] |= 0x10
        st[(st2 & 0x10) != 0] |= 0x10
        st2[(st & 0x10) != 0] |= 0x10
        assert orc.diff_batch(S, S2, ((st & 0x10) != 0).astype(np.uint8))[0] == -1


# ---- the device policy code, host build -----------------------------------------------------------------------
@pytest.fixture(scope="module")
def hs():
    from hostsim import HostSim
    return HostSim()


@pytest.mark.parametrize("mask,stress,nact", [(15, 0, 6), (0b1001, 0, 6), (15, 1, 6), (0b0011, 1, 5)])
def test_device_policy_code_matches_restatement(orc, hs, mask, stress, nact):
    n, ticks, seed, env0 = 384, 260, 1000 + mask, 77
    seeds = oracle.clean_seeds(48)
    S = orc.zero_state(n)
    for i in range(n):
        orc.init_state(S[i:i + 1], seeds[i % 48])
    if stress:
        S["agents"]["canKick"] = 1
        S["agents"]["maxBombCount"] = 4
        S["agents"]["bombStrength"] = 3
    T = S.copy()
    st = np.zeros(n, np.uint8)
    recs, bad = hs.pack(S, st)
    assert not bad.any()
    A, B = orc.simple_agents(n), orc.simple_agents(n)
    episodes = 0
    for t in range(ticks):
        mv = orc.rng_moves(seed, env0, n, t, nact)
        mv2 = mv.copy()
        orc.simple_moves_batch(S, st, A, seed, env0, t, mask, mv)
        hs.simple_moves(recs, B, seed, env0, t, mask, mv2)
        live = (st & 0x11) == 0
        assert (mv[live] == mv2[live]).all(), "tick %d env %d" % (t, np.nonzero((mv != mv2).any(1) & live)[0][0])
        assert A[live].tobytes() == B[live].tobytes()
        fl = np.zeros(n, np.uint8)
        orc.env_step_batch(S, st, mv, fl)
        st[(fl & 0x3E) != 0] |= 0x10
        hs.step_records(recs, mv2, False)
        done = np.nonzero((st & 0x11) != 0)[0]
        if done.size:                              # new game, new agents
            episodes += done.size
            S[done], st[done], A[done], B[done] = T[done], 0, 0, 0
            recs[done] = hs.pack(S[done], st[done])[0]
    S2, _ = hs.unpack(recs)
    assert orc.diff_batch(S, S2)[0] == -1
    assert episodes > 50


def test_unreachable_enemy_from_the_origin_corner(orc, hs):
    """MoveTowardsPosition on an unreachable target reads predecessor 0 = cell (0,0): only an agent standing ON
    (0,0) is sent towards the target (strategy.cpp:101-124); everybody else idles."""
    for corner, expect in (((0, 0), RIGHT), ((10, 10), None)):
        s = orc.zero_state()
        orc.kill(s, 2, 3)
        orc.put_agent(s, corner[0], corner[1], 0)
        orc.put_agent(s, 5, 0 if corner == (0, 0) else 10, 1)
        for x in range(11):
            for y in range(11):
                if s["board"][0, y, x] == 0:
                    orc.put_item(s, x, y, 1)      # walls everywhere: nothing is reachable
        m = orc.move_towards(s, 0, 2, 7)
        assert m == (expect if expect is not None else IDLE)
        # and the device code agrees through act(): enemy within 7, memories chosen so that _HasRPLoop is false
        A = orc.simple_agents(1)
        A["rp_count"][0, 0] = 2
        A["recent"][0, 0] = [0x11, 0x22, 0x33, 0x44]
        B = A.copy()
        mv, mv2 = np.zeros((1, 4), np.uint8), np.zeros((1, 4), np.uint8)
        orc.simple_moves_batch(s, None, A, 5, 0, 0, 1, mv)
        recs, bad = hs.pack(s)
        assert not bad.any()
        hs.simple_moves(recs, B, 5, 0, 0, 1, mv2)
        assert (mv == mv2).all() and A.tobytes() == B.tobytes()


# ---- the path-finding pieces on arbitrary boards (far more varied than what full games reach) ---------------
def _random_positions(orc, n, seed):
    """mid-game states with extra random walls / passages / powerups / bombs so that reachability is non-trivial"""
    rng = np.random.default_rng(seed)
    seeds = oracle.clean_seeds(48)
    S = orc.zero_state(n)
    for i in range(n):
        orc.init_state(S[i:i + 1], seeds[i % 48])
    S["agents"]["maxBombCount"] = 3
    S["agents"]["bombStrength"] = 2
    st = np.zeros(n, np.uint8)
    for t in range(int(rng.integers(5, 40))):
        orc.env_step_batch(S, st, orc.rng_moves(seed, 0, n, t, 6))
    for i in range(n):                       # open up / close random cells (never an agent's, a bomb's or a flame's cell)
        for _ in range(int(rng.integers(0, 40))):
            x, y = int(rng.integers(0, 11)), int(rng.integers(0, 11))
            c = int(S["board"][i, y, x])
            if c in (0, 1) or (c >> 8) == 2 or c in (6, 7, 8):
                S["board"][i, y, x] = int(rng.choice([0, 0, 0, 1, 2 << 8, 6, 7, 8]))
    return S, rng


def test_fill_rmap_and_move_towards_match_compiled_reference_on_random_boards(orc, ref):
    S, rng = _random_positions(orc, 300, 5)
    for i in range(S.shape[0]):
        s = S[i:i + 1]
        for a in range(4):
            if s["agents"]["dead"][0, a]:
                continue
            assert (orc.fill_rmap(s, a) == ref.fill_rmap(s, a)).all(), (i, a)
            tx, ty = int(rng.integers(0, 11)), int(rng.integers(0, 11))
            if (tx, ty) != (int(s["agents"]["x"][0, a]), int(s["agents"]["y"][0, a])):
                assert orc.move_towards(s, a, 0, tx, ty) == ref.move_towards(s, a, 0, tx, ty)
            for kind in (1, 2, 3):
                radius = int(rng.integers(1, 11))
                assert orc.move_towards(s, a, kind, radius) == ref.move_towards(s, a, kind, radius), (i, a, kind, radius)


def test_device_floods_equal_the_queue_bfs_on_random_boards(orc, hs):
    """the bitboard floods of pom_policy.cuh against the oracle's FillRMap + MoveTowardsPosition / MoveTowardsSafePlace:
    every cell of the board as a target, every radius"""
    S, rng = _random_positions(orc, 160, 6)
    recs, bad = hs.pack(S)
    assert not bad.any()
    checked = 0
    for i in range(S.shape[0]):
        s = S[i:i + 1]
        for a in range(4):
            if s["agents"]["dead"][0, a]:
                continue
            ax, ay = int(s["agents"]["x"][0, a]), int(s["agents"]["y"][0, a])
            for ty in range(11):
                for tx in range(11):
                    if (tx, ty) == (ax, ay):
                        continue
                    assert hs.move_towards(recs[i], a, tx, ty) == orc.move_towards(s, a, 0, tx, ty), (i, a, tx, ty)
                    checked += 1
            for radius in range(1, 11):
                assert hs.move_towards_safe_place(recs[i], a, radius) == orc.move_towards(s, a, 3, radius), (i, a, radius)
    assert checked > 20000


def test_act_on_arbitrary_states_and_memories(orc, hs):
    """SimpleAgent::act on random boards with ARBITRARY agent memories (any ring index / count, off-board remembered
    positions, any stale move-queue content): what pom_batch_policy_upload may hand to the device code"""
    S, rng = _random_positions(orc, 400, 7)
    n = S.shape[0]
    recs, bad = hs.pack(S)
    assert not bad.any()
    for rep in range(6):
        A = orc.simple_agents(n)
        nib = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 15], np.uint8)          # -1 .. 11
        A["recent"] = nib[rng.integers(0, 13, (n, 4, 4))] | (nib[rng.integers(0, 13, (n, 4, 4))] << 4)
        A["rp_index"] = rng.integers(0, 4, (n, 4))
        A["rp_count"] = rng.integers(0, 5, (n, 4))
        A["move_queue"] = (rng.integers(0, 6, (n, 4)) | (rng.integers(0, 6, (n, 4)) << 3) |
                           (rng.integers(0, 6, (n, 4)) << 6) | (rng.integers(0, 6, (n, 4)) << 9)).astype(np.uint16)
        # some agents remember cells next to them, so that SortDirections has something to reorder
        for e in range(0, n, 3):
            for a in range(4):
                x, y = int(S["agents"]["x"][e, a]), int(S["agents"]["y"][e, a])
                A["recent"][e, a, rep % 4] = ((x + 1) & 15) | ((y & 15) << 4)
                A["recent"][e, a, (rep + 1) % 4] = (x & 15) | (((y + 1) & 15) << 4)
        B = A.copy()
        mv = np.zeros((n, 4), np.uint8)
        mv2 = mv.copy()
        orc.simple_moves_batch(S, None, A, 40 + rep, 0, rep, 15, mv)
        hs.simple_moves(recs, B, 40 + rep, 0, rep, 15, mv2)
        assert (mv == mv2).all(), np.argwhere(mv != mv2)[:3]
        assert A.tobytes() == B.tobytes()
