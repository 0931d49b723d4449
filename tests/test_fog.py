"""Fog of war: pom_batch_observe / pomcore::fog_state against the oracle's definition
(oracle/pom_oracle.c pom_oracle_fog).  The reference only declares the feature (bboard.hpp:62,218-226,529), so the
checks are: device code == independent C definition, plus the properties that define fog."""
import os

import numpy as np
import pytest

import oracle

FOG = 5


def _mid_game_states(orc, n=256, ticks=40, seed=21):
    seeds = oracle.clean_seeds(32)
    S = orc.zero_state(n)
    for i in range(n):
        orc.init_state(S[i:i + 1], seeds[i % 32])
    S["agents"]["maxBombCount"] = 3
    S["agents"]["bombStrength"] = 3
    status = np.zeros(n, np.uint8)
    for t in range(ticks):
        orc.env_step_batch(S, status, orc.rng_moves(seed, 0, n, t, 6))
    return S


def _check_properties(full, obs, agent, view):
    ax, ay = full["agents"]["x"][:, agent], full["agents"]["y"][:, agent]
    yy, xx = np.mgrid[0:11, 0:11]
    vis = (np.abs(xx[None] - ax[:, None, None]) <= view) & (np.abs(yy[None] - ay[:, None, None]) <= view)
    assert (obs["board"][vis] == full["board"][vis]).all()
    assert (obs["board"][~vis] == FOG).all()
    assert (obs["agents"][:, agent] == full["agents"][:, agent]).all()
    assert (obs["agents"]["dead"] == full["agents"]["dead"]).all()
    assert (obs["timeStep"] == full["timeStep"]).all() and (obs["aliveAgents"] == full["aliveAgents"]).all()
    assert (obs["bombs_count"] <= full["bombs_count"]).all() and (obs["flames_count"] <= full["flames_count"]).all()
    for e in range(full.shape[0]):
        for k in range(int(obs["bombs_count"][e])):
            b = int(obs["bombs"][e, k])
            assert vis[e, (b >> 4) & 15, b & 15]
        for i in range(4):
            g = obs["agents"][e, i]
            if i != agent and g["x"] >= 0:
                assert vis[e, g["y"], g["x"]] and not g["dead"]


@pytest.mark.parametrize("agent,view", [(0, 4), (2, 4), (3, 1), (1, 0), (1, 10)])
def test_fog_definition_and_device_code_on_host(orc, agent, view):
    from hostsim import HostSim
    hs = HostSim()
    full = _mid_game_states(orc)
    a = orc.fog_batch(full.copy(), agent, view)
    b = hs.fog_batch(full.copy(), agent, view)
    assert orc.diff_batch(a, b)[0] == -1
    _check_properties(full, a, agent, view)
    if view >= 10:
        assert (a["board"] == full["board"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("agent,view", [(0, 4), (3, 2)])
def test_observe_on_gpu(orc, agent, view):
    import pomcpp_b200 as pb
    n = 5000
    b = pb.Batch(n, n_templates=64)
    b.rollout(45, 3, 0, pb.ROLL_NO_RESET)
    full, st = b.download()
    obs, st2 = b.observe(agent, view)
    assert (st == st2).all()
    assert orc.diff_batch(obs, orc.fog_batch(full.copy(), agent, view))[0] == -1
    _check_properties(full[:300], obs[:300], agent, view)
    part, _ = b.observe(agent, view, first=1234, count=77)
    assert orc.diff_batch(part, obs[1234:1234 + 77])[0] == -1
    L = pb.lib()
    assert L.pom_batch_observe(b.h, 0, 1, 4, 4, obs.ctypes.data, None) == -1
    assert L.pom_batch_observe(b.h, 0, n + 1, 0, 4, obs.ctypes.data, None) == -4
    b.close()


@pytest.mark.parametrize("agent,view", [(0, 4), (3, 4), (1, 2), (2, 10)])
def test_observation_planes_device_code_on_host(orc, agent, view):
    """pomcore::observe_planes on packed records vs the oracle's definition built from the AoS State, on mid-game states
    and on the mutually inconsistent states of tests/test_fuzz_states.py"""
    from hostsim import HostSim
    from test_fuzz_states import mutated_states
    hs = HostSim()
    for S in (mutated_states(orc, 77 + agent, 1500), _mid_game_states(orc)):
        recs, bad = hs.pack(S, np.zeros(S.shape[0], np.uint8))
        S, recs = S[bad == 0].copy(), recs[bad == 0].copy()
        a = orc.observe_planes_batch(S, agent, view)
        b = hs.observe_planes(recs, agent, view)
        assert (a == b).all(), np.argwhere(a != b)[:5]
        # the board plane is the fogged board with the game's item ids
        fog = orc.fog_batch(S.copy(), agent, view)
        assert ((a[:, :121].reshape(-1, 11, 11) == 5) == (fog["board"] == FOG)).all()
        assert (a[:, 484] == S["agents"]["x"][:, agent]).all() and (a[:, 489] == (1 - S["agents"]["dead"]) @ np.array([1, 2, 4, 8])).all()
        assert a[:, 121:242].any()
    assert a[:, 363:484].any() or view < 10     # flame lives show up once bombs have gone off


@pytest.mark.parametrize("agent,view", [(0, 4), (3, 4), (1, 2), (2, 5), (1, 0), (0, 1), (2, 3)])
def test_cropped_observation_device_code_on_host(orc, agent, view):
    """pomcore::observe_cropped on packed records vs the oracle's definition (the window of the full planes), on mid-game
    states and on the mutually inconsistent states of tests/test_fuzz_states.py; and the definition's own properties"""
    from hostsim import HostSim
    from test_fuzz_states import mutated_states
    hs = HostSim()
    W = 2 * view + 1
    for S in (mutated_states(orc, 31 + agent, 1500), _mid_game_states(orc)):
        recs, bad = hs.pack(S, np.zeros(S.shape[0], np.uint8))
        S, recs = S[bad == 0].copy(), recs[bad == 0].copy()
        a = orc.observe_cropped_batch(S, agent, view)
        b = hs.observe_cropped(recs, agent, view)
        assert a.shape == b.shape == (S.shape[0], (4 * W * W + 12 + 31) // 32 * 32)
        assert (a == b).all(), np.argwhere(a != b)[:5]
        full = orc.observe_planes_batch(S, agent, view)
        # the centre of the window is the observer's own cell, the scalars are the full layout's
        centre = view * W + view
        ax, ay = S["agents"]["x"][:, agent], S["agents"]["y"][:, agent]
        assert (a[:, centre] == full[np.arange(S.shape[0]), ax + 11 * ay]).all()
        assert (a[:, 4 * W * W:4 * W * W + 12] == full[:, 484:496]).all()
        assert (a[:, 4 * W * W + 12:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mask,view", [(15, 4), (0b0101, 2)])
def test_observation_planes_on_gpu(orc, mask, view):
    import pomcpp_b200 as pb
    n = 5000 + 17
    b = pb.Batch(n, n_templates=64)
    b.rollout(45, 3, 0, pb.ROLL_NO_RESET)
    full, _ = b.download()
    obs = b.observe_planes(mask, view)
    agents = [a for a in range(4) if (mask >> a) & 1]
    assert obs.shape == (len(agents), n, 512)
    for k, a in enumerate(agents):
        want = orc.observe_planes_batch(full, a, view)
        assert (obs[k] == want).all(), (a, np.argwhere(obs[k] != want)[:4])
    L = pb.lib()
    dev = b.alloc(512 * 256)
    assert L.pom_batch_observe_planes(b.h, None, 1, 4) == -1
    assert L.pom_batch_observe_planes(b.h, dev, 0, 4) == -1
    assert L.pom_batch_observe_planes(b.h, dev, 16, 4) == -1
    assert L.pom_batch_observe_planes(b.h, dev, 1, -1) == -1
    b.free(dev)
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mask,view", [(15, 4), (0b0110, 2), (0b1000, 5), (0b0001, 0)])
def test_cropped_observation_planes_on_gpu(orc, mask, view):
    """pom_batch_observe_planes_cropped against the oracle's definition, ragged batch size, every agent of the mask"""
    import pomcpp_b200 as pb
    n = 5000 + 19
    b = pb.Batch(n, n_templates=64)
    b.rollout(45, 3, 0, pb.ROLL_NO_RESET)
    full, _ = b.download()
    obs = b.observe_planes_cropped(mask, view)
    agents = [a for a in range(4) if (mask >> a) & 1]
    rb = int(pb.lib().pom_obs_cropped_bytes(view))
    assert rb == orc.obs_cropped_bytes(view) and obs.shape == (len(agents), n, rb)
    for k, a in enumerate(agents):
        want = orc.observe_cropped_batch(full, a, view)
        assert (obs[k] == want).all(), (a, np.argwhere(obs[k] != want)[:4])
    L = pb.lib()
    dev = b.alloc(512 * 256)
    assert L.pom_batch_observe_planes_cropped(b.h, None, 1, 4) == -1
    assert L.pom_batch_observe_planes_cropped(b.h, dev, 0, 4) == -1
    assert L.pom_batch_observe_planes_cropped(b.h, dev, 1, 6) == -1
    assert L.pom_batch_observe_planes_cropped(b.h, dev, 1, -1) == -1
    assert L.pom_obs_cropped_bytes(4) == 352 and L.pom_obs_cropped_bytes(6) == 0
    b.free(dev)
    b.close()


@pytest.mark.gpu
def test_observation_planes_throughput_smoke():
    """1 Mi envs x 4 agents (2 GB of planes per call) stays on the device; just check it runs and is not absurdly slow"""
    import time
    import pomcpp_b200 as pb
    n = 1 << 20
    b = pb.Batch(n, n_templates=256)
    stride = int(pb.lib().pom_batch_obs_stride(b.h))
    dev = b.alloc(4 * stride * 512)
    pb._ck(pb.lib().pom_batch_observe_planes(b.h, dev, 15, 4))
    b.sync()
    b.event(0)
    for _ in range(5):
        pb._ck(pb.lib().pom_batch_observe_planes(b.h, dev, 15, 4))
    b.event(1)
    b.sync()
    ms = b.elapsed_ms() / 5
    print("observe_planes: %.3f ms per 1 Mi envs x 4 agents, %.0f GB/s written" % (ms, 4 * n * 512 / ms / 1e6))
    assert ms < 20
    b.free(dev)
    b.close()
