"""CPU fuzz: the tick on states that no game reaches — mid-game states with random, mutually INCONSISTENT edits (items
written over cells, agents' records moved without the board, ghost BOMB cells, extra queue entries, random bomb
directions / timers / stale slots, bombCount and dead flags out of step).  A user may upload such states, so the device
code's fast paths must not rely on invariants that only hold along real games.

  restatement (oracle/pom_oracle.c)  vs  the kernel body compiled for the host (tests/hostsim), always;
  restatement                        vs  the compiled, unmodified reference, where oracle/_ref exists.
Envs on which the restatement detects one of the reference's defects are frozen and excluded, as everywhere."""
import numpy as np
import pytest

import oracle

AGENT0 = 1 << 24


def mutated_states(orc, seed, n):
    rng = np.random.default_rng(seed)
    seeds = oracle.clean_seeds(48)
    S = orc.zero_state(n)
    for i in range(n):
        orc.init_state(S[i:i + 1], seeds[i % 48])
    S["agents"]["maxBombCount"] = rng.integers(1, 6, (n, 4))
    S["agents"]["bombStrength"] = rng.integers(1, 5, (n, 4))
    S["agents"]["canKick"] = rng.integers(0, 2, (n, 4))
    st = np.zeros(n, np.uint8)
    for t in range(int(rng.integers(6, 16))):
        orc.env_step_batch(S, st, orc.rng_moves(seed, 0, n, t, 6))
    S = S[(st & 0x11) == 0].copy()
    for i in range(S.shape[0]):
        for _ in range(int(rng.integers(0, 8))):
            m = int(rng.integers(0, 12))
            x, y = int(rng.integers(0, 11)), int(rng.integers(0, 11))
            cnt, lo = int(S["bombs_count"][i]), int(S["bombs_index"][i])
            if m == 0:
                S["board"][i, y, x] = int(rng.choice([0, 1, 3, 6, 7, 8, (2 << 8) + int(rng.integers(0, 5))]))
            elif m == 1 and cnt > 0:
                sl = (lo + int(rng.integers(0, cnt))) % 20
                b = int(S["bombs"][i, sl])
                S["bombs"][i, sl] = (b & ~0xFF00000) | (int(rng.integers(0, 5)) << 20) | (int(rng.integers(0, 2)) << 24)
            elif m == 2 and cnt > 0:
                sl = (lo + int(rng.integers(0, cnt))) % 20
                S["bombs"][i, sl] = (int(S["bombs"][i, sl]) & ~0xF0000) | (int(rng.integers(1, 11)) << 16)
            elif m == 3:
                a = int(rng.integers(0, 4))
                S["agents"]["x"][i, a], S["agents"]["y"][i, a] = x, y
            elif m == 4:
                S["board"][i, y, x] = AGENT0 + int(rng.integers(0, 4))
            elif m == 5 and cnt < 18:
                S["bombs"][i, (lo + cnt) % 20] = (x | (y << 4) | (int(rng.integers(0, 4)) << 8) | (int(rng.integers(1, 5)) << 12) |
                                                 (int(rng.integers(1, 11)) << 16) | (int(rng.integers(0, 5)) << 20))
                S["bombs_count"][i] += 1
                if rng.integers(0, 2):
                    S["board"][i, y, x] = 3
            elif m == 6:
                sl = int(rng.integers(0, 20))
                if (sl - lo) % 20 >= cnt:
                    S["bombs"][i, sl] = int(rng.integers(0, 1 << 28))
            elif m == 7:
                S["agents"]["bombCount"][i, int(rng.integers(0, 4))] = int(rng.integers(-1, 4))
            elif m == 8:
                S["agents"]["dead"][i, int(rng.integers(0, 4))] = int(rng.integers(0, 2))
                S["aliveAgents"][i] = int(4 - S["agents"]["dead"][i].sum())
            elif m == 9 and S["flames_count"][i] > 0:          # a live flame with another time left / strength
                sl = (int(S["flames_index"][i]) + int(rng.integers(0, S["flames_count"][i]))) % 20
                S["flames"]["timeLeft"][i, sl] = int(rng.integers(1, 6))
                S["flames"]["strength"][i, sl] = int(rng.integers(0, 7))
            elif m == 10:                                      # a flame cell without queue entry (origin id 0, board_logic.cpp:504)
                S["board"][i, y, x] = (4 << 16) + int(rng.integers(0, 4))
            elif m == 11 and S["flames_count"][i] > 0:         # a flame cell pointing at a live flame from anywhere on the board
                sl = (int(S["flames_index"][i]) + int(rng.integers(0, S["flames_count"][i]))) % 20
                fx, fy = int(S["flames"]["x"][i, sl]), int(S["flames"]["y"][i, sl])
                S["board"][i, y, x] = (4 << 16) + ((fx + 11 * fy) << 3) + int(rng.integers(0, 4))
    return S


@pytest.mark.parametrize("by_rays", [False, True], ids=["one-lane", "by-rays"])
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_kernel_body_on_inconsistent_states(orc, seed, by_rays):
    """by_rays: the tick in the decomposition the warp-cooperative kernels use (ray-parallel explosions, arm-parallel pops)"""
    from hostsim import HostSim
    hs = HostSim()
    hs.set_by_rays(by_rays)
    S = mutated_states(orc, seed, 3000)
    recs, bad = hs.pack(S, np.zeros(S.shape[0], np.uint8))
    S, recs = S[bad == 0].copy(), recs[bad == 0].copy()
    n = S.shape[0]
    assert n > 1500
    st = np.zeros(n, np.uint8)
    steps = 0
    for t in range(12):
        mv = orc.rng_moves(seed + 1000, 0, n, t, 6)
        fl = np.zeros(n, np.uint8)
        orc.env_step_batch(S, st, mv, fl)
        st[(fl & 0x3E) != 0] |= 0x10
        hs.step_records(recs, mv, False)
        S2, st2 = hs.unpack(recs)
        e, why = orc.diff_batch(S, S2, ((st & 0x10) != 0).astype(np.uint8))
        assert e == -1, "tick %d env %d field group %d" % (t, e, why)
        assert ((st2 & 0x11) == (st & 0x11)).all()
        steps += int(((st & 0x11) == 0).sum())
    assert steps > 5000


@pytest.mark.parametrize("seed", [201, 202])
def test_restatement_vs_compiled_reference_on_inconsistent_states(orc, ref, seed):
    S = mutated_states(orc, seed, 3000)
    n = S.shape[0]
    S2 = S.copy()
    st, st2 = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    for t in range(12):
        mv = orc.rng_moves(seed + 1000, 0, n, t, 6)
        fl, pre = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        orc.env_step_batch(S, st, mv, fl)
        ref.env_step_batch(S2, st2, mv, pre, ((fl & 0x3E) != 0).astype(np.uint8))
        st[(fl & 0x3E) != 0] |= 0x10
        st[(st2 & 0x10) != 0] |= 0x10
        st2[(st & 0x10) != 0] |= 0x10
        e, why = orc.diff_batch(S, S2, ((st & 0x10) != 0).astype(np.uint8))
        assert e == -1, "tick %d env %d field group %d" % (t, e, why)


def test_simple_agent_on_inconsistent_states(orc):
    """SimpleAgent::act on the same kind of states: restatement vs the device policy code (host build), and vs the compiled
    reference agent where available"""
    from hostsim import HostSim
    hs = HostSim()
    R = oracle.reference() if oracle.have_reference() else None
    S = mutated_states(orc, 401, 3000)
    recs, bad = hs.pack(S, np.zeros(S.shape[0], np.uint8))
    S, recs = S[bad == 0].copy(), recs[bad == 0].copy()
    n = S.shape[0]
    A = orc.simple_agents(n)
    B = A.copy()
    C_ = R.simple_agents(n) if R is not None else None
    for t in range(3):
        mv, mv2, mv3 = (np.zeros((n, 4), np.uint8) for _ in range(3))
        orc.simple_moves_batch(S, None, A, 401, 0, t, 15, mv)
        hs.simple_moves(recs, B, 401, 0, t, 15, mv2)
        assert (mv == mv2).all() and A.tobytes() == B.tobytes(), "device policy code, act %d" % t
        if C_ is not None:
            C_.moves_batch(S, None, orc.rng_moves(401, 0, n, t, 5), 15, mv3)
            assert (mv == mv3).all() and A.tobytes() == C_.export().tobytes(), "compiled reference agent, act %d" % t
