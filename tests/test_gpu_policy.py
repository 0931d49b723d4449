"""GPU tests (-m gpu) of the device-side SimpleAgent policy: pom_batch_policy_moves and the
fused rollout with POM_ROLL_SIMPLE, compared move by move / field by field with the oracle's restatement of
agents::SimpleAgent (oracle/pom_oracle_agent.c) and with games recorded from the compiled reference
(tests/golden/simple_agent.npz)."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pb():
    import pomcpp_b200 as pb
    assert os.path.exists(pb.LIB_PATH), "libpom_b200.so missing on the GPU box"
    assert pb.device_count() > 0, "no CUDA device"
    return pb


def _same_agents(A, B):
    return A.tobytes() == B.tobytes()


@pytest.mark.parametrize("mask,autoreset", [(15, False), (15, True), (0b0110, True), (0b0001, False)])
def test_policy_moves_per_tick(pb, orc, mask, autoreset):
    """act -> Step per tick through the C ABI; the oracle runs the same collection loop (environment.cpp:137-146)."""
    n, ticks, seed, env0 = 1536 + 7, 140, 4242, 300
    b = pb.Batch(n, env_offset=env0, n_templates=48)
    T, _ = b.templates()
    S, _ = b.download()
    status = np.zeros(n, np.uint8)
    episode = np.zeros(n, np.int64)
    A = orc.simple_agents(n)
    out = np.zeros(n, np.uint8)
    acted = 0
    for t in range(ticks):
        mv = orc.rng_moves(seed, env0, n, t, 6)
        gm = mv.copy()
        b.policy_moves_host(gm, seed, t, mask)
        orc.simple_moves_batch(S, status, A, seed, env0, t, mask, mv)
        live = (status & 0x11) == 0
        assert (gm[live] == mv[live]).all(), "tick %d: moves differ in env %d" % (t, np.nonzero((gm != mv).any(1) & live)[0][0])
        acted += int(live.sum())
        b.step_host(gm, out, pb.STEP_AUTORESET if autoreset else 0)
        fl = np.zeros(n, np.uint8)
        orc.env_step_batch(S, status, mv, fl)
        status[(fl & 0x3E) != 0] |= 0x10
        assert (out == status).all()
        if autoreset:
            for e in np.nonzero((status & 0x11) != 0)[0]:
                episode[e] += 1
                S[e] = T[(env0 + e + episode[e]) % T.shape[0]]
                status[e] = 0
                A[e] = 0
        if t % 20 == 19 or t == ticks - 1:
            assert _same_agents(b.policy_download(), A), "agent memories differ after tick %d" % t
    G, gst = b.download()
    assert orc.diff_batch(G, S)[0] == -1 and (gst == status).all()
    assert acted > n * 20
    if autoreset:
        assert episode.sum() > 0
    b.close()


def _oracle_policy_rollout(orc, S, T, A, env0, ticks, seed, nact, max_ticks, mask, tick0=0, no_reset=False):
    n = S.shape[0]
    status = np.zeros(n, np.uint8)
    episode = np.zeros(n, np.int64)
    stats = np.zeros(10, np.int64)
    nT = T.shape[0]
    for k in range(ticks):
        mv = orc.rng_moves(seed, env0, n, tick0 + k, nact)
        live = (status & 0x31) == 0
        orc.simple_moves_batch(S, (status | ((status & 0x20) >> 5)).astype(np.uint8), A, seed, env0, tick0 + k, mask, mv)
        stats[0] += int(live.sum())
        frozen = ~live
        before, sb = S[frozen].copy(), status[frozen].copy()
        fl = np.zeros(n, np.uint8)
        orc.env_step_batch(S, status, mv, fl)
        status[live & ((fl & 0x3E) != 0)] |= 0x10
        S[frozen], status[frozen] = before, sb
        if max_ticks:
            status[live & ((status & 1) == 0) & (S["timeStep"] >= max_ticks)] |= 0x20
        for e in np.nonzero(live & ((status & 0x31) != 0))[0]:
            stats[1] += 1
            if status[e] & 1:
                stats[6 if status[e] & 2 else 2 + ((status[e] >> 2) & 3)] += 1
            elif status[e] & 0x20:
                stats[7] += 1
            else:
                stats[9] += 1
            stats[8] += int(S["timeStep"][e])
            if not no_reset:
                episode[e] += 1
                S[e] = T[(env0 + e + episode[e]) % nT]
                status[e] = 0
                A[e] = 0
    return status, stats


@pytest.mark.parametrize("mask,harmless,max_ticks,no_reset", [(15, 0, 800, False), (0b1010, 0, 120, False),
                                                               (0b0111, 1, 0, True)])
def test_rollout_with_simple_agents_matches_oracle_replay(pb, orc, mask, harmless, max_ticks, no_reset):
    n, ticks, seed, env0 = 1024 + 33, 260, 808, 5000
    b = pb.Batch(n, env_offset=env0, n_templates=40, max_ticks=max_ticks)
    T, _ = b.templates()
    S, _ = b.download()
    A = orc.simple_agents(n)
    flags = pb.ROLL_SIMPLE(mask) | (pb.ROLL_HARMLESS if harmless else 0) | (pb.ROLL_NO_RESET if no_reset else 0)
    b.rollout(100, seed, 0, flags)            # split: the agents' memories must survive between launches
    b.rollout(ticks - 100, seed, 100, flags)
    G, gst = b.download()
    status, stats = _oracle_policy_rollout(orc, S, T, A, env0, ticks, seed, 5 if harmless else 6, max_ticks, mask,
                                           no_reset=no_reset)
    e, why = orc.diff_batch(G, S)
    assert e == -1, "env %d field group %d" % (e, why)
    assert (gst == status).all()
    assert (b.stats().as_array() == stats).all(), (b.stats().as_dict(), stats)
    assert _same_agents(b.policy_download(), A)
    assert stats[1] > 0
    b.close()


def test_rollout_then_per_tick_share_agent_memories(pb, orc):
    """fused rollout and pom_batch_policy_moves are two views of the same agents"""
    n, seed = 700, 31
    b = pb.Batch(n, n_templates=16, max_ticks=800)
    T, _ = b.templates()
    S, _ = b.download()
    A = orc.simple_agents(n)
    b.rollout(60, seed, 0, pb.ROLL_SIMPLE(15))
    status, _ = _oracle_policy_rollout(orc, S, T, A, 0, 60, seed, 6, 800, 15)
    mv = np.zeros((n, 4), np.uint8)
    gm = mv.copy()
    b.policy_moves_host(gm, seed, 60, 15)
    orc.simple_moves_batch(S, status, A, seed, 0, 60, 15, mv)
    assert (gm == mv).all() and _same_agents(b.policy_download(), A)
    # policy_reset = four new agents per env
    b.policy_reset()
    assert not b.policy_download().view(np.uint8).any()
    # upload/download round trip of the memories
    b.policy_upload(A)
    assert _same_agents(b.policy_download(), A)
    b.close()


def test_golden_simple_agent_games(pb, orc):
    """games played by the UNMODIFIED reference SimpleAgent (recorded by tests/golden/make_golden.py)"""
    g = np.load(os.path.join(GOLD, "simple_agent.npz"))
    init = g["init"].copy().view(oracle.STATE_DT).reshape(-1)
    moves, seed, ticks = g["moves"], int(g["seed"]), int(g["ticks"])
    n = init.shape[0]
    b = pb.Batch(n, n_templates=1, empty=True)
    b.upload(init)
    for t in range(ticks):
        gm = np.zeros((n, 4), np.uint8)
        b.policy_moves_host(gm, seed, t, 15)
        live = g["live"][t].astype(bool)
        assert (gm[live] == moves[t][live]).all(), "tick %d" % t
        b.step_host(gm, None, 0)
    G, gst = b.download()
    final = g["final"].copy().view(oracle.STATE_DT).reshape(-1)
    assert orc.diff_batch(G, final, (g["excluded"] != 0).astype(np.uint8))[0] == -1
    A = g["agents"].copy().view(oracle.SIMPLE_DT).reshape(n, 4)
    keep = g["excluded"] == 0
    assert _same_agents(b.policy_download()[keep], A[keep])
    b.close()


def test_policy_act_single_agent(pb, orc):
    """pom_batch_policy_act: the entry point behind the host mirror's agents::SimpleAgent::act"""
    g = np.load(os.path.join(GOLD, "simple_agent.npz"))
    init = g["init"].copy().view(oracle.STATE_DT).reshape(-1)[96:104].copy()     # boosted-regime games
    n = init.shape[0]
    b = pb.Batch(n, n_templates=1, empty=True)
    b.upload(init)
    S, status = init.copy(), np.zeros(n, np.uint8)
    A = orc.simple_agents(n)
    rng = np.random.default_rng(5)
    for t in range(80):
        mv = np.zeros((n, 4), np.uint8)
        for e in range(n):
            if status[e] & 0x11:
                continue
            for a in range(4):
                if S["agents"]["dead"][e, a]:
                    continue
                d = int(rng.integers(0, 5))
                mv[e, a] = orc.simple_act(S[e:e + 1], a, A[e, a:a + 1], d)
                assert b.policy_act(e, a, d) == mv[e, a], (t, e, a)
        b.step_host(mv, None, 0)
        orc.env_step_batch(S, status, mv)
    assert _same_agents(b.policy_download(), A)
    G, _ = b.download()
    assert orc.diff_batch(G, S)[0] == -1
    b.close()


def test_policy_abi_errors(pb):
    b = pb.Batch(64, n_templates=4)
    L = pb.lib()
    assert L.pom_batch_policy_moves(b.h, None, 1, 0, 15) == -1
    dev = b.alloc(256)
    assert L.pom_batch_policy_moves(b.h, dev, 1, 0, 0) == -1
    assert L.pom_batch_policy_moves(b.h, dev, 1, 0, 16) == -1
    assert L.pom_batch_policy_download(b.h, 60, 10, dev) == -4
    assert L.pom_batch_policy_upload(b.h, 0, 65, dev) == -4
    import ctypes as C
    m = C.c_int(0)
    assert L.pom_batch_policy_act(b.h, 64, 0, 0, C.byref(m)) == -4
    assert L.pom_batch_policy_act(b.h, 0, 4, 0, C.byref(m)) == -1
    assert L.pom_batch_policy_act(b.h, 0, 0, 5, C.byref(m)) == -1
    assert L.pom_batch_policy_act(b.h, 0, 0, 0, None) == -1
    b.free(dev)
    b.close()


def test_rollout_with_simple_agents_at_scale(pb, orc):
    """BASELINE config-3 size (1 Mi envs) with four SimpleAgents: counters are consistent, and a contiguous slice of
    envs in the middle of the batch, replayed on the oracle from the same counter RNG, matches field by field."""
    n, ticks, seed = 1 << 20, 120, 2024
    b = pb.Batch(n, n_templates=512, max_ticks=100)
    T, _ = b.templates()
    b.rollout(ticks, seed, 0, pb.ROLL_SIMPLE(15))
    st = b.stats()
    assert st.env_steps == n * ticks
    assert st.episodes == sum(st.wins) + st.draws + st.truncated + st.invalid and st.episodes >= n
    lo, cnt = 777_001, 192
    S = np.stack([T[(lo + e) % 512] for e in range(cnt)]).astype(oracle.STATE_DT)
    A = orc.simple_agents(cnt)
    status, _ = _oracle_policy_rollout(orc, S, T, A, lo, ticks, seed, 6, 100, 15)
    G, gst = b.download(lo, cnt)
    e, why = orc.diff_batch(G, S)
    assert e == -1, "env %d field group %d" % (lo + e, why)
    assert (gst == status).all()
    assert _same_agents(b.policy_download(lo, cnt), A)
    b.close()


@pytest.mark.parametrize("mask,harmless", [(0b1011, 0), (15, 1)])
def test_rollout_with_simple_agents_fused_equals_tick_by_tick(pb, orc, mask, harmless, monkeypatch):
    """pom_batch_rollout with SimpleAgents on a large batch runs k_policy_moves + the per-tick step kernel; with
    POM_ROLL_POLICY=fused (read when the handle is made) the fused kernel: same states, status bytes, counters and agent
    memories.  Envs that carry TRUNCATED without having been reset (a NO_RESET rollout before) stay frozen in both."""
    n, seed = (1 << 16) + 4099, 77
    monkeypatch.setenv("POM_ROLL_POLICY", "fused")
    a = pb.Batch(n, n_templates=128, max_ticks=60)
    monkeypatch.delenv("POM_ROLL_POLICY")
    c = pb.Batch(n, n_templates=128, max_ticks=60)
    flags = pb.ROLL_SIMPLE(mask) | (pb.ROLL_HARMLESS if harmless else 0)
    for x in (a, c):
        x.rollout(70, seed, 0, pb.ROLL_NO_RESET | (pb.ROLL_HARMLESS if harmless else 0))   # leaves truncated / done envs behind
        x.clear_stats()
        l0 = x.launch_count()
        x.rollout(45, seed, 70, flags)
        x.launches = x.launch_count() - l0
    assert a.launches == 1 and c.launches == 90
    A, sa = a.download()
    Cc, sc = c.download()
    assert A.tobytes() == Cc.tobytes() and (sa == sc).all()
    assert a.stats().as_dict() == c.stats().as_dict()
    assert _same_agents(a.policy_download(), c.policy_download())
    a.close(); c.close()
