"""GPU tests (-m gpu): the parity tests proper.  Everything goes through the C ABI
(include/pom_batch.h) and is compared bit-exactly with the oracle."""
import os

import numpy as np
import pytest

import oracle
import scenarios

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pb():
    import pomcpp_b200 as pb
    assert os.path.exists(pb.LIB_PATH), "libpom_b200.so missing on the GPU box"
    assert pb.device_count() > 0, "no CUDA device"
    return pb


@pytest.fixture(scope="module")
def gpu_be(pb, orc):
    from gpu_backend import GpuBackend
    return GpuBackend(orc)


@pytest.mark.parametrize("fn", scenarios.STEP_SCENARIOS, ids=lambda f: f.__name__)
def test_reference_scenarios_on_gpu(gpu_be, fn):
    fn(gpu_be)


def test_golden_scenarios(pb, orc):
    g = np.load(os.path.join(GOLD, "scenarios.npz"))
    before = g["before"].copy().view(oracle.STATE_DT).reshape(-1)
    after = g["after"].copy().view(oracle.STATE_DT).reshape(-1)
    n = before.shape[0]
    b = pb.Batch(n, n_templates=1, empty=True)
    b.upload(before)
    rt, _ = b.download()
    assert orc.diff_batch(rt, before)[0] == -1, "upload/download round trip"
    b.step_host(np.ascontiguousarray(g["moves"]), None, pb.STEP_RAW)
    out, st = b.download()
    e, why = orc.diff_batch(out, after)
    assert e == -1, "transition %d (%s) differs in field group %d" % (e, g["names"][e], why)
    b.close()


@pytest.mark.parametrize("name", ["random6", "harmless5", "stress6"])
def test_golden_traces(pb, orc, name):
    g = np.load(os.path.join(GOLD, "traces.npz"))
    n, ticks, nact, stress, rs = [int(v) for v in g[name + "_cfg"]]
    init = g[name + "_init"].copy().view(oracle.STATE_DT).reshape(-1)
    excluded = g[name + "_excluded"]
    b = pb.Batch(n, n_templates=1, empty=True)
    b.upload(init)
    moves_dev = b.alloc(4 * n)
    for t in range(ticks):
        b.generate_moves(moves_dev, rs, t, nact)
        b.step(moves_dev, 0)
        S, st = b.download()
        live = (excluded < 0) | (excluded > t)
        bad = (orc.hash_batch(S) != g[name + "_hash"][t]) & live
        assert not bad.any(), "tick %d env %d" % (t, int(np.nonzero(bad)[0][0]))
        assert (((st ^ g[name + "_status"][t]) & 0x0F)[live] == 0).all()
    final = g[name + "_final"].copy().view(oracle.STATE_DT).reshape(-1)
    assert orc.diff_batch(S, final, (excluded >= 0).astype(np.uint8))[0] == -1
    b.free(moves_dev)
    b.close()


def test_device_init_matches_reference_init(pb, orc):
    """K3: InitState on the device (mt19937_64 + libstdc++ Lemire) vs the oracle, incl. the golden boards."""
    b = pb.Batch(8, n_templates=512, first_seed=0x1337)
    T, seeds = b.templates()
    g = np.load(os.path.join(GOLD, "init.npz"))
    assert (seeds[:64] == g["seeds"]).all(), "clean-seed filter differs"
    for k in range(64):
        brd = g["boards"][k].copy()
        brd[0, 0], brd[0, 10], brd[10, 10], brd[10, 0] = [(1 << 24) + i for i in range(4)]
        assert (T["board"][k] == brd).all()
    for k in range(0, 512, 7):
        s = orc.zero_state()
        assert orc.init_state(s, int(seeds[k])) == 0
        assert orc.diff_batch(T[k:k + 1], s)[0] == -1, "template %d seed %d" % (k, seeds[k])
    # env e starts as template[(offset + e) % n]
    S, st = b.download()
    assert orc.diff_batch(S, T[:8])[0] == -1 and not st.any()
    b.close()


def _run_vs_oracle(pb, orc, n, ticks, nact, stress, seed, every_tick=True, n_templates=256, ring_start=None):
    b = pb.Batch(n, n_templates=n_templates)
    S, st = b.download()
    if stress:
        S["agents"]["canKick"] = 1
        S["agents"]["maxBombCount"] = 5
        S["agents"]["bombStrength"] = 4
    if ring_start is not None:
        S["bombs_index"] = (np.arange(n) + ring_start) % 20
        S["flames_index"] = (np.arange(n) * 7 + ring_start) % 20
    if stress or ring_start is not None:
        b.upload(S)
    status = np.zeros(n, np.uint8)
    moves_dev = b.alloc(4 * n)
    steps = 0
    for t in range(ticks):
        b.generate_moves(moves_dev, seed, t, nact)
        b.step(moves_dev, 0)
        mv = orc.rng_moves(seed, 0, n, t, nact)
        steps += int(((status & 1) == 0).sum())
        orc.env_step_batch(S, status, mv)
        if every_tick or t == ticks - 1:
            G, gst = b.download()
            e, why = orc.diff_batch(G, S)
            assert e == -1, "tick %d env %d field group %d" % (t, e, why)
            assert (gst == status).all(), "status differs at tick %d" % t
    b.free(moves_dev)
    b.close()
    return steps


def test_config2_65536_harmless_800_ticks_every_field_every_tick(pb, orc):
    """BASELINE config 2 at full size."""
    steps = _run_vs_oracle(pb, orc, 65536, 800, 5, 0, 42)
    assert steps == 65536 * 800


def test_random_with_bombs_every_tick(pb, orc):
    _run_vs_oracle(pb, orc, 16384, 96, 6, 0, 43)


def test_stress_kicks_and_chains_every_tick(pb, orc):
    _run_vs_oracle(pb, orc, 16384, 200, 6, 1, 44)


def test_ring_wraparound(pb, orc):
    """FixedQueue index arithmetic (reference general_test.cpp:41-61): rings that start near the end of the array."""
    _run_vs_oracle(pb, orc, 8192, 150, 6, 1, 46, ring_start=15)


def test_ragged_sizes(pb, orc):
    for n in (1, 31, 129, 257, 1000):
        _run_vs_oracle(pb, orc, n, 40, 6, 1, 45 + n, n_templates=7)


def _oracle_rollout(orc, S, T, env0, ticks, seed, nact, max_ticks, tick0=0, no_reset=False):
    """Host replay of pom_batch_rollout's rule: step, truncate, count, reset to template[(env+episode)%nT]."""
    n = S.shape[0]
    status = np.zeros(n, np.uint8)
    episode = np.zeros(n, np.int64)
    stats = np.zeros(10, np.int64)
    nT = T.shape[0]
    for k in range(ticks):
        mv = orc.rng_moves(seed, env0, n, tick0 + k, nact)
        live = (status & 0x31) == 0
        stats[0] += int(live.sum())
        frozen = ~live
        before = S[frozen].copy()
        sb = status[frozen].copy()
        orc.env_step_batch(S, status, mv)
        S[frozen] = before
        status[frozen] = sb
        if max_ticks:
            tr = live & ((status & 1) == 0) & (S["timeStep"] >= max_ticks)
            status[tr] |= 0x20
        fin = live & ((status & 0x31) != 0)
        for e in np.nonzero(fin)[0]:
            stats[1] += 1
            if status[e] & 1:
                if status[e] & 2:
                    stats[6] += 1
                else:
                    stats[2 + ((status[e] >> 2) & 3)] += 1
            elif status[e] & 0x20:
                stats[7] += 1
            else:
                stats[9] += 1          # aborted: left the reference's defined domain
            stats[8] += int(S["timeStep"][e])
            if not no_reset:
                episode[e] += 1
                S[e] = T[(env0 + e + episode[e]) % nT]
                status[e] = 0
    return status, stats


@pytest.mark.parametrize("harmless,max_ticks", [(0, 800), (1, 60)])
def test_rollout_fused_matches_oracle_replay(pb, orc, harmless, max_ticks):
    n, ticks, seed, env0 = 2048, 150, 77, 1000
    b = pb.Batch(n, env_offset=env0, n_templates=64, max_ticks=max_ticks)
    T, _ = b.templates()
    S, _ = b.download()
    # two launches: the tick counter continues across launches
    b.rollout(100, seed, 0, pb.ROLL_HARMLESS if harmless else 0)
    b.rollout(ticks - 100, seed, 100, pb.ROLL_HARMLESS if harmless else 0)
    G, gst = b.download()
    status, stats = _oracle_rollout(orc, S, T, env0, ticks, seed, 5 if harmless else 6, max_ticks)
    e, why = orc.diff_batch(G, S)
    assert e == -1, "env %d field group %d" % (e, why)
    assert (gst == status).all()
    assert (b.stats().as_array() == stats).all(), (b.stats().as_dict(), stats)
    assert stats[1] > 0
    b.close()


def test_rollout_no_reset_freezes(pb, orc):
    n, ticks, seed = 1024, 120, 5
    b = pb.Batch(n, n_templates=32, max_ticks=100)
    T, _ = b.templates()
    S, _ = b.download()
    b.rollout(ticks, seed, 0, pb.ROLL_NO_RESET)
    G, gst = b.download()
    status, stats = _oracle_rollout(orc, S, T, 0, ticks, seed, 6, 100, no_reset=True)
    assert orc.diff_batch(G, S)[0] == -1
    assert (gst == status).all()
    assert (b.stats().as_array() == stats).all()
    b.close()


def test_per_tick_autoreset_matches_rollout_rule(pb, orc):
    n, ticks, seed = 2048, 90, 9
    b = pb.Batch(n, n_templates=16, max_ticks=0)
    T, _ = b.templates()
    S, _ = b.download()
    moves_dev = b.alloc(4 * n)
    for t in range(ticks):
        b.generate_moves(moves_dev, seed, t, 6)
        b.step(moves_dev, pb.STEP_AUTORESET | pb.STEP_COUNT)
    G, gst = b.download()
    status, stats = _oracle_rollout(orc, S, T, 0, ticks, seed, 6, 0)
    assert orc.diff_batch(G, S)[0] == -1
    assert (gst == status).all()
    assert (b.stats().as_array() == stats).all(), (b.stats().as_dict(), stats)
    b.free(moves_dev)
    b.close()


def test_step_host_e2e_path(pb, orc):
    n = 3000
    b = pb.Batch(n, n_templates=50)
    S, _ = b.download()
    status = np.zeros(n, np.uint8)
    out = np.zeros(n, np.uint8)
    for t in range(40):
        mv = orc.rng_moves(3, 0, n, t, 6)
        b.step_host(mv, out, 0)
        orc.env_step_batch(S, status, mv)
        assert (out == status).all()
    G, _ = b.download()
    assert orc.diff_batch(G, S)[0] == -1
    b.close()


def test_clone_and_tree_search_expansion(pb, orc):
    """BASELINE config 5 shape: roots x 6^4 joint actions, one Step each."""
    n_roots_pool = 512
    src = pb.Batch(n_roots_pool, n_templates=64)
    moves_dev = src.alloc(4 * n_roots_pool)
    for t in range(16):
        src.generate_moves(moves_dev, 21, t, 6)
        src.step(moves_dev, 0)
    R, rst = src.download()
    roots = np.nonzero((rst & 1) == 0)[0][:24].astype(np.uint32)
    assert roots.shape[0] == 24
    # plain clone (gather)
    dst = pb.Batch(24 * 1296, n_templates=1, empty=True)
    dst.clone_from(src, roots[::-1].copy(), first_dst=5)
    C_, cst = dst.download(5, 24)
    assert orc.diff_batch(C_, R[roots[::-1]])[0] == -1 and (cst == rst[roots[::-1]]).all()
    # fused fan-out + Step
    dst.expand_step_from(src, roots, 1296, 0)
    G, gst = dst.download()
    E = np.repeat(R[roots], 1296)
    est = np.repeat(rst[roots], 1296)
    j = np.tile(np.arange(1296), 24)
    mv = np.stack([j % 6, (j // 6) % 6, (j // 36) % 6, (j // 216) % 6], axis=1).astype(np.uint8)
    orc.env_step_batch(E, est, np.ascontiguousarray(mv))
    e, why = orc.diff_batch(G, E)
    assert e == -1, "child %d field group %d" % (e, why)
    assert (gst == est).all()
    src.free(moves_dev)
    src.close()
    dst.close()


def test_upload_rejects_unrepresentable_states(pb, orc):
    s = orc.zero_state(3)
    s["board"][0, 3, 3] = (4 << 16) | (17 << 3)
    b = pb.Batch(3, n_templates=1, empty=True)
    with pytest.raises(pb.PomError) as e:
        b.upload(s)
    assert e.value.code == -5
    st = b.status()
    assert st[0] & pb.STATUS_INVALID and not st[1] & pb.STATUS_INVALID
    with pytest.raises(pb.PomError):
        b.download(2, 5)
    b.close()


def test_config3_1M_envs_properties(pb, orc):
    """BASELINE config 3 at full size: 1,048,576 envs, random actions incl. bombs; size-independent
    properties + a strided sample of envs compared with the oracle on every field."""
    n, ticks, seed = 1 << 20, 64, 1234
    b = pb.Batch(n, n_templates=4096, max_ticks=800)
    T, _ = b.templates()
    sample = np.arange(0, n, 509)
    S0 = T[sample % 4096].copy()
    moves_dev = b.alloc(4 * n)
    for t in range(ticks):
        b.generate_moves(moves_dev, seed, t, 6)
        b.step(moves_dev, pb.STEP_AUTORESET | pb.STEP_COUNT)
    s = b.stats()
    assert s.env_steps == n * ticks
    assert s.episodes == sum(s.wins) + s.draws + s.truncated + s.invalid
    assert s.episodes > n and s.invalid < 1e-4 * s.episodes
    assert 15 < s.sum_episode_len / s.episodes < 40        # mean episode length ~26 ticks
    # sampled envs: replay each on the oracle with the same reset rule
    for e in sample[:400]:
        S = T[e % 4096: e % 4096 + 1].copy()
        _oracle_rollout(orc, S, T, int(e), ticks, seed, 6, 800)
        G, _ = b.download(int(e), 1)
        assert orc.diff_batch(G, S)[0] == -1, "env %d" % e
    b.free(moves_dev)
    b.close()


def test_cpp_host_layer_selftest(pb):
    """The reference-style known-answer checks compiled against include/bboard.hpp (the drop-in mirror)."""
    import subprocess
    exe = os.path.join(os.path.dirname(pb.LIB_PATH), "host", "pom_selftest")
    assert os.path.exists(exe), "pomcpp_b200/host/pom_selftest not built"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all passed" in out.stdout


def test_cpp_bench_driver_runs(pb):
    import json
    import subprocess
    exe = os.path.join(os.path.dirname(pb.LIB_PATH), "host", "pom_bench")
    out = subprocess.run([exe, "--gpus", "1", "--envs-per-gpu", "65536", "--steps", "20", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["episode_stats"]["env_steps"] == 65536 * 23 and d["value"] > 1e8
    # the other modes: fused rollout with device-side SimpleAgent opponents, host-buffer (e2e) stepping, expansion
    for extra in (["--mode", "rollout", "--simple", "14", "--ticks", "50", "--steps", "2", "--warmup", "1"],
                  ["--mode", "step", "--simple", "15", "--steps", "10", "--warmup", "2"],
                  ["--mode", "host", "--steps", "10", "--warmup", "2"],
                  ["--mode", "hostc", "--steps", "10", "--warmup", "2"],
                  ["--mode", "expand", "--roots", "64", "--steps", "2", "--warmup", "1"]):
        out = subprocess.run([exe, "--gpus", "1", "--envs-per-gpu", "65536"] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, " ".join(extra) + "\n" + out.stdout + out.stderr
        assert json.loads(out.stdout.strip().splitlines()[-1])["value"] > 1e6


def test_state_primitives_on_device(pb, orc):
    """pom_batch_apply: ExplodeTopBomb / ExplodeBombAt / PopFlame vs the restatement's Step-internal behaviour."""
    s = orc.zero_state()
    orc.put_agents_in_corners(s)
    orc.plant_bomb(s, 5, 5, 0, True)
    orc.plant_bomb(s, 6, 5, 1, True)
    b = pb.Batch(1, n_templates=1, empty=True)
    b.upload(s)
    b.apply(0, 2, 1)                       # ExplodeBombAt(1): chain-explodes bomb 0 as well
    G, _ = b.download()
    assert int(G["bombs_count"][0]) == 0 and int(G["flames_count"][0]) == 2
    b.apply(0, 3)                          # PopFlame
    G, _ = b.download()
    assert int(G["flames_count"][0]) == 1
    brd, dirty = pb.make_board(0x1337)
    ref = orc.zero_state()
    assert orc.init_board_items(ref, 0x1337) == 0 and dirty == 0
    assert (brd["board"] == ref["board"]).all()
    assert pb.make_board(0x13327)[1] == 1
    b.close()


def test_step_host_reports_terminal_status_with_autoreset(pb, orc):
    """With auto-reset the status bytes returned by pom_batch_step_host are those of the episode that just ended."""
    n, ticks, seed = 4096, 60, 12
    b = pb.Batch(n, n_templates=32, max_ticks=0)
    T, _ = b.templates()
    S, _ = b.download()
    status = np.zeros(n, np.uint8)
    episode = np.zeros(n, np.int64)
    out = np.zeros(n, np.uint8)
    seen_done = 0
    for t in range(ticks):
        mv = orc.rng_moves(seed, 0, n, t, 6)
        b.step_host(mv, out, pb.STEP_AUTORESET)
        orc.env_step_batch(S, status, mv)
        assert (out == status).all(), "tick %d" % t
        fin = np.nonzero(status & 0x11)[0]
        seen_done += fin.shape[0]
        for e in fin:
            episode[e] += 1
            S[e] = T[(e + episode[e]) % 32]
            status[e] = 0
    assert seen_done > n
    G, gst = b.download()
    assert orc.diff_batch(G, S)[0] == -1 and not gst.any()
    b.close()


def test_replay_tool_on_golden_trace_file(pb):
    """pom_replay re-runs the reference-generated POMTRC1 trace on the GPU; exit code 0 = all hashes match."""
    import subprocess
    exe = os.path.join(os.path.dirname(pb.LIB_PATH), "host", "pom_replay")
    out = subprocess.run([exe, os.path.join(GOLD, "stress96.pomtrc"), "--print", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr
    assert "0 mismatches" in out.stdout and "agent 0:" in out.stdout


def test_abi_error_codes(pb):
    """Every entry point reports bad input with a negative code and a message instead of crashing."""
    import ctypes as C
    L = pb.lib()
    h = C.c_void_p()
    d = pb.InitDesc()
    d.n_templates = 4
    d.first_seed = 0x1337
    assert L.pom_batch_init(C.byref(h), 0, 0, C.byref(d)) == -1                    # POM_E_ARG: no envs
    assert L.pom_batch_init(C.byref(h), 99, 16, C.byref(d)) == -1                  # no such device
    d.n_templates = 0
    assert L.pom_batch_init(C.byref(h), 0, 16, C.byref(d)) == -1
    assert b"bad argument" in L.pom_last_error()
    b = pb.Batch(64, n_templates=4)
    b2 = pb.Batch(64, n_templates=1, empty=True)
    assert L.pom_batch_step(b.h, None, 0) == -1
    assert L.pom_batch_step_host(b.h, None, None, 0) == -1
    idx = np.array([1, 2, 64], np.uint32)
    assert L.pom_batch_clone(b2.h, 0, b.h, idx.ctypes.data_as(C.c_void_p), 3) == -4          # POM_E_RANGE: source index
    assert L.pom_batch_clone(b2.h, 63, b.h, idx.ctypes.data_as(C.c_void_p), 2) == -4         # destination range
    assert L.pom_batch_expand_step(b2.h, b.h, idx.ctypes.data_as(C.c_void_p), 1, 0, 0) == -1  # fanout 0
    assert L.pom_batch_expand_step(b2.h, b.h, idx.ctypes.data_as(C.c_void_p), 1, 1296, 0) == -4  # destination too small
    assert L.pom_batch_apply(b.h, 64, 0, 1, 1, 1) == -4 and L.pom_batch_apply(b.h, 0, 9, 0, 0, 0) == -1
    assert L.pom_batch_apply(b.h, 0, 0, 11, 0, 1) == -1
    st = np.zeros(8, np.uint8)
    assert L.pom_batch_download(b.h, 60, 8, None, st.ctypes.data_as(C.c_void_p)) == -4
    assert L.pom_batch_generate_moves(b.h, None, 1, 0, 6) == -1
    # the handle is still usable after all these errors
    S, s0 = b.download()
    assert S.shape[0] == 64 and not s0.any()
    b.close()
    b2.close()


def test_config1_single_game_fixed_trace(pb, orc):
    """BASELINE config 1: one 11x11 FFA game on the default board (InitState, seed 0x1337), four RandomAgents
    replaced by a fixed-seed joint-action trace, one Step per tick until the game ends or 800 ticks."""
    for trace_seed in range(8):
        b = pb.Batch(1, n_templates=1, first_seed=0x1337)
        S, _ = b.download()
        ref0 = orc.zero_state()
        assert orc.init_state(ref0, 0x1337) == 0 and orc.diff_batch(S, ref0)[0] == -1
        status = np.zeros(1, np.uint8)
        out = np.zeros(1, np.uint8)
        for t in range(800):
            mv = orc.rng_moves(1000 + trace_seed, 0, 1, t, 6)
            b.step_host(mv, out, 0)
            orc.env_step_batch(S, status, mv)
            G, gst = b.download()
            assert orc.diff_batch(G, S)[0] == -1 and gst[0] == status[0] == out[0], (trace_seed, t)
            if status[0] & 1:
                break
        assert status[0] & 1, "random agents always finish well before 800 ticks"
        b.close()


def test_config4_fused_800_tick_rollout(pb, orc):
    """BASELINE config 4 (one GPU's shard at reduced width): fused 800-tick rollout with in-kernel RNG, truncation at
    800 and auto-reset; counters are consistent and sampled envs replay bit-exactly on the oracle."""
    n, ticks, seed, env0 = 131072, 800, 4242, 3 * 524288
    b = pb.Batch(n, env_offset=env0, n_templates=4096, max_ticks=800)
    T, _ = b.templates()
    b.rollout(500, seed, 0, 0)
    b.rollout(300, seed, 500, 0)
    s = b.stats()
    assert s.env_steps == n * ticks
    assert s.episodes == sum(s.wins) + s.draws + s.truncated + s.invalid
    assert 20 < s.sum_episode_len / s.episodes < 35 and s.invalid < 1e-4 * s.episodes
    for e in (0, 1, 777, 65535, 131071):
        S = T[(env0 + e) % 4096: (env0 + e) % 4096 + 1].copy()
        status, _ = _oracle_rollout(orc, S, T, env0 + e, ticks, seed, 6, 800)
        G, gst = b.download(e, 1)
        assert orc.diff_batch(G, S)[0] == -1 and gst[0] == status[0], "env %d" % e
    b.close()


def test_config5_expansion_full_size(pb, orc):
    """BASELINE config 5 at full size: 4096 root states x 6^4 joint actions (5,308,416 children), clone + one Step fused;
    all children of sampled roots compared with the oracle."""
    n_roots = 4096
    src = pb.Batch(n_roots, n_templates=512)
    src.rollout(16, 99, 0, pb.ROLL_NO_RESET)
    R, rst = src.download()
    dst = pb.Batch(n_roots * 1296, n_templates=1, empty=True)
    dst.expand_step_from(src, np.arange(n_roots, dtype=np.uint32), 1296, 0)
    j = np.arange(1296)
    mv = np.ascontiguousarray(np.stack([j % 6, (j // 6) % 6, (j // 36) % 6, (j // 216) % 6], axis=1).astype(np.uint8))
    for root in (0, 1, 2047, 4095):
        G, gst = dst.download(root * 1296, 1296)
        E = np.repeat(R[root:root + 1], 1296)
        est = np.repeat(rst[root:root + 1], 1296)
        orc.env_step_batch(E, est, mv)
        e, why = orc.diff_batch(G, E)
        assert e == -1, "root %d child %d field group %d" % (root, e, why)
        assert (gst == est).all()
    src.close()
    dst.close()


@pytest.mark.parametrize("n,pinned", [(70_001, True), (300_000, True), (300_000, False)])
def test_step_host_large_batches(pb, orc, n, pinned):
    """pom_batch_step_host at sizes where it pipelines: pinned buffers are read and written by the kernel directly
    (zero-copy), pageable ones go through the copy pipeline (six chunks on two compute streams above 2^18 envs)"""
    b = pb.Batch(n, n_templates=64)
    S, _ = b.download()
    status = np.zeros(n, np.uint8)
    if pinned:
        bufs = [(pb.pinned_array((n, 4), np.uint8), pb.pinned_array((n,), np.uint8)) for _ in range(2)]
    else:
        bufs = [((np.zeros((n, 4), np.uint8), None), (np.zeros(n, np.uint8), None)) for _ in range(2)]
    for t in range(24):
        (mv, _o1), (out, _o2) = bufs[t % 2]
        mv[:] = orc.rng_moves(5, 0, n, t, 6)
        b.step_host(mv, out, 0)
        orc.env_step_batch(S, status, np.ascontiguousarray(mv))
        assert (out == status).all(), "tick %d" % t
    G, gst = b.download()
    assert orc.diff_batch(G, S)[0] == -1 and (gst == status).all()
    b.close()
    if pinned:
        for (_, o1), (_, o2) in bufs:
            pb.pinned_free(o1)
            pb.pinned_free(o2)


@pytest.mark.parametrize("n", [(1 << 18) + 1500, (1 << 19) + 1500])
def test_step_overlap_flag_matches_oracle(pb, orc, n):
    """POM_STEP_OVERLAP: the two halves of the batch are stepped on two internal streams, ticks overlap at their edges;
    results, status bytes and counters must be those of the plain per-tick path (auto-reset rule included).  Up to 0.5 Mi
    envs the halves run on the tile kernel, above on the persistent kernel: one size for each."""
    ticks, seed = 36, 13
    b = pb.Batch(n, n_templates=32, max_ticks=0)
    T, _ = b.templates()
    S, _ = b.download()
    moves_dev = b.alloc(4 * n * ticks)
    for t in range(ticks):
        b.generate_moves(moves_dev.value + 4 * n * t, seed, t, 6)
    for t in range(ticks):
        b.step(moves_dev.value + 4 * n * t, pb.STEP_AUTORESET | pb.STEP_COUNT | pb.STEP_OVERLAP)
        if t == 17:
            mid = b.stats()                 # any other call joins both halves first
            assert mid.env_steps == n * 18
    G, gst = b.download()
    status, stats = _oracle_rollout(orc, S, T, 0, ticks, seed, 6, 0)
    assert orc.diff_batch(G, S)[0] == -1
    assert (gst == status).all()
    assert (b.stats().as_array() == stats).all(), (b.stats().as_dict(), stats)
    b.free(moves_dev)
    b.close()


def test_step_host_async_double_buffered(pb, orc):
    """two batches stepped alternately through pom_batch_step_host_async + sync (the e2e pattern of bench.py)"""
    h, ticks = 40_000, 30
    B = [pb.Batch(h, env_offset=i * h, n_templates=32) for i in range(2)]
    S = [x.download()[0] for x in B]
    status = [np.zeros(h, np.uint8) for _ in range(2)]
    mv = [pb.pinned_array((h, 4), np.uint8) for _ in range(2)]
    out = [pb.pinned_array((h,), np.uint8) for _ in range(2)]

    def launch(i, t):
        mv[i][0][:] = orc.rng_moves(8, i * h, h, t, 6)
        B[i].step_host_async(mv[i][0], out[i][0], 0)

    def check(i, t):
        B[i].sync()
        orc.env_step_batch(S[i], status[i], np.ascontiguousarray(mv[i][0]))
        assert (out[i][0] == status[i]).all(), "half %d tick %d" % (i, t)

    launch(0, 0)
    for t in range(ticks):
        launch(1, t)
        check(0, t)
        if t + 1 < ticks:
            launch(0, t + 1)
        check(1, t)
    for i in range(2):
        G, gst = B[i].download()
        assert orc.diff_batch(G, S[i])[0] == -1 and (gst == status[i]).all()
    # pageable buffers are refused (the call could not be asynchronous)
    assert pb.lib().pom_batch_step_host_async(B[0].h, np.zeros((h, 4), np.uint8).ctypes.data, None, 0) == -1
    for x in B:
        x.close()
    for a, o in mv + out:
        pb.pinned_free(o)


def test_examples_run(pb):
    """examples/: plain C against the C ABI (actor loop with device opponents, double-buffered) and C++ against the bboard
    mirror (Environment + SimpleAgent, then one ply of tree search)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(pb.LIB_PATH))
    out = subprocess.run([os.path.join(root, "examples", "actor_loop"), "4096", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "491520 env-steps" in out.stdout, out.stdout + out.stderr
    out = subprocess.run([os.path.join(root, "examples", "tree_search")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.count("continuations") == 6, out.stdout + out.stderr


def test_inconsistent_uploaded_states_on_gpu(pb, orc):
    """the states of tests/test_fuzz_states.py (random, mutually inconsistent edits of mid-game states) uploaded through
    the C ABI and stepped on the GPU"""
    from hostsim import HostSim
    from test_fuzz_states import mutated_states
    S = mutated_states(orc, 301, 6000)
    _, bad = HostSim().pack(S, np.zeros(S.shape[0], np.uint8))
    S = S[bad == 0].copy()
    n = S.shape[0]
    assert n > 3000
    b = pb.Batch(n, n_templates=1, empty=True)
    b.upload(S)
    status = np.zeros(n, np.uint8)
    out = np.zeros(n, np.uint8)
    for t in range(12):
        mv = orc.rng_moves(1301, 0, n, t, 6)
        b.step_host(mv, out, 0)
        fl = np.zeros(n, np.uint8)
        orc.env_step_batch(S, status, mv, fl)
        status[(fl & 0x3E) != 0] |= 0x10
        assert ((out & 0x11) == (status & 0x11)).all(), "tick %d" % t
        G, _ = b.download()
        e, why = orc.diff_batch(G, S, ((status & 0x10) != 0).astype(np.uint8))
        assert e == -1, "tick %d env %d field group %d" % (t, e, why)
    b.close()


def test_make_board_extreme_seeds(pb, orc):
    """InitBoardItems on the device for negative and extreme seeds (std::mt19937_64(int) sign-extends the seed)"""
    checked = 0
    for seed in list(range(-12, 0)) + [-2 ** 31, -2 ** 31 + 1, 2 ** 31 - 1, 2 ** 31 - 2, 0, 1]:
        s = orc.zero_state()
        dirty = orc.init_board_items(s, seed)
        board, d = pb.make_board(seed)
        assert bool(d) == bool(dirty), seed
        if not dirty:
            assert (board["board"][0] == s["board"][0]).all(), seed
            checked += 1
    assert checked >= 6


@pytest.mark.parametrize("no_reset", [False, True])
def test_step_seq_fused_ticks_with_given_moves(pb, orc, no_reset):
    """pom_batch_step_seq: the fused kernel fed with the caller's moves (tick-major device buffer) = the per-tick path with
    the same moves, episode rule included"""
    n, ticks, seed, env0 = 3000, 70, 19, 40
    b = pb.Batch(n, env_offset=env0, n_templates=24, max_ticks=50)
    T, _ = b.templates()
    S, _ = b.download()
    seq = b.alloc(4 * n * ticks)
    for t in range(ticks):
        b.generate_moves(seq.value + 4 * n * t, seed, t, 6)
    fl = pb.ROLL_NO_RESET if no_reset else 0
    b.step_seq(seq, 30, fl)
    b.step_seq(seq.value + 4 * n * 30, ticks - 30, fl)
    G, gst = b.download()
    status, stats = _oracle_rollout(orc, S, T, env0, ticks, seed, 6, 50, no_reset=no_reset)
    assert orc.diff_batch(G, S)[0] == -1
    assert (gst == status).all()
    assert (b.stats().as_array() == stats).all(), (b.stats().as_dict(), stats)
    assert pb.lib().pom_batch_step_seq(b.h, None, 1, 0) == -1
    assert pb.lib().pom_batch_step_seq(b.h, seq, 1, pb.ROLL_HARMLESS) == -1
    b.free(seq)
    b.close()
