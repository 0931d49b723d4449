"""Generates the golden vectors under tests/golden/ from the COMPILED, UNMODIFIED reference
(oracle/_ref/libpomref.so).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Outputs
  scenarios.npz : every Step executed by the reference's own unit-test scenarios
                  (tests/scenarios.py) as (before, moves, after) triples.
  traces.npz    : seeded random episodes — initial states (InitState on clean seeds, plus the
                  kick/bomb-boosted "stress" variant), per-tick per-env FNV hashes of the full state,
                  per-tick status bytes and the final states.  Moves come from the shared stateless
                  RNG (pom_oracle_rng_moves), so they are regenerated, not stored.
  init.npz      : InitBoardItems boards for the first 64 clean seeds >= 0x1337.
  simple_agent.npz : games played by four of the reference's agents::SimpleAgent per env (unmodified
                  simple_agent.cpp / strategy.cpp; the agents' engines are re-seeded before each act so that
                  their one intDist draw is byte a of pom_oracle_rng_moves(seed, env, tick, 5)): initial
                  states, the moves of every tick, final states and the agents' final memories.
                  `python tests/golden/make_golden.py simple` regenerates only this file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import scenarios  # noqa: E402

TRACE_CONFIGS = [
    # name, n_envs, ticks, n_actions, stress, rng seed
    ("random6", 256, 96, 6, 0, 42),
    ("harmless5", 128, 160, 5, 0, 43),
    ("stress6", 256, 160, 6, 1, 44),
]


def initial_states(R, n, stress, seeds):
    S = R.zero_state(n)
    for e in range(n):
        R.init_state(S[e:e + 1], seeds[e % len(seeds)])
    if stress:
        S["agents"]["canKick"] = 1
        S["agents"]["maxBombCount"] = 5
        S["agents"]["bombStrength"] = 4
    return S


def run_trace(R, O, S, ticks, n_actions, rng_seed):
    """Environment semantics, finished envs freeze; envs leaving the parity domain (ref_precheck)
    freeze too and are reported in `excluded` (tick of exclusion, -1 = never)."""
    n = S.shape[0]
    status = np.zeros(n, np.uint8)
    pre = np.zeros(n, np.uint8)
    fo = np.zeros(n, np.uint8)
    shadow = S.copy()
    shadow_status = np.zeros(n, np.uint8)
    hashes = np.zeros((ticks, n), np.uint64)
    stat = np.zeros((ticks, n), np.uint8)
    excluded = np.full(n, -1, np.int32)
    d1 = 0
    for t in range(ticks):
        mv = O.rng_moves(rng_seed, 0, n, t, n_actions)
        # the restatement runs the same tick first: it detects defect D5 (the reference would hang)
        O.env_step_batch(shadow, shadow_status, mv, fo)
        R.env_step_batch(S, status, mv, pre, ((fo & 0x20) != 0).astype(np.uint8))
        d1 += int(((pre & 7) > 0).sum())
        newly = ((status & 0x10) != 0) & (excluded < 0)
        excluded[newly] = t
        hashes[t] = O.hash_batch(S)
        stat[t] = status
    return hashes, stat, excluded, d1


def write_pomtrc(R, O, path, n, ticks, nact, stress, rs, seeds):
    import struct
    S = initial_states(R, n, stress, seeds)
    init = S.copy()
    status = np.zeros(n, np.uint8)
    shadow, shadow_status, fo = S.copy(), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    moves = np.zeros((ticks, n, 4), np.uint8)
    for t in range(ticks):
        mv = O.rng_moves(rs, 0, n, t, nact)
        moves[t] = mv
        O.env_step_batch(shadow, shadow_status, mv, fo)
        R.env_step_batch(S, status, mv, None, ((fo & 0x20) != 0).astype(np.uint8))
    assert not (status & 0x10).any(), "pick another seed: an env left the parity domain"
    with open(path, "wb") as f:
        f.write(b"POMTRC1\0")
        f.write(struct.pack("<4I", n, ticks, 0, 0))
        f.write(init.tobytes())
        f.write(moves.tobytes())
        f.write(O.hash_batch(S).tobytes())
        f.write(status.tobytes())
    print("%s: %d envs x %d ticks, %d done, %d bytes" % (os.path.basename(path), n, ticks, int((status & 1).sum()), os.path.getsize(path)))


def simple_agent_games(R, O, seeds, n=192, ticks=220, seed=77):
    S = initial_states(R, n, 0, seeds)
    half = n // 2          # second half: the boosted regime (kicks, long flames) so that fleeing logic is exercised
    S["agents"]["canKick"][half:] = 1
    S["agents"]["maxBombCount"][half:] = 3
    S["agents"]["bombStrength"][half:] = 3
    init = S.copy()
    status = np.zeros(n, np.uint8)
    shadow, shadow_status, fo = S.copy(), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    agents = R.simple_agents(n)
    moves = np.zeros((ticks, n, 4), np.uint8)
    live = np.zeros((ticks, n), np.uint8)
    excluded = np.zeros(n, np.uint8)
    for t in range(ticks):
        draws = O.rng_moves(seed, 0, n, t, 5)
        mv = np.zeros((n, 4), np.uint8)
        live[t] = (status & 0x11) == 0
        agents.moves_batch(S, status, draws, 15, mv)
        moves[t] = mv
        O.env_step_batch(shadow, shadow_status, mv, fo)
        R.env_step_batch(S, status, mv, None, ((fo & 0x20) != 0).astype(np.uint8))
        excluded |= ((status & 0x10) != 0).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "simple_agent.npz"), init=init.view(np.uint8).reshape(n, 1004),
                        final=S.view(np.uint8).reshape(n, 1004), moves=moves, live=live, excluded=excluded,
                        agents=agents.export().view(np.uint8).reshape(n, 32), seed=np.int64(seed), ticks=np.int64(ticks))
    hist = np.bincount(moves[live.astype(bool)].ravel(), minlength=6)
    print("simple_agent: %d envs x %d ticks, done %d, excluded %d, move histogram %s, %d bytes" %
          (n, ticks, int((status & 1).sum()), int(excluded.sum()), hist.tolist(),
           os.path.getsize(os.path.join(HERE, "simple_agent.npz"))))


def main():
    oracle.build()
    R = oracle.reference()
    O = oracle.restatement()
    if len(sys.argv) > 1 and sys.argv[1] == "simple":
        simple_agent_games(R, O, oracle.clean_seeds(64))
        return

    # 1. unit-test scenarios
    rec = scenarios.Recorder(R)
    names = []
    for fn in scenarios.STEP_SCENARIOS:
        k0 = len(rec.transitions)
        fn(rec)
        names += [fn.__name__] * (len(rec.transitions) - k0)
    before = np.concatenate([t[0] for t in rec.transitions])
    moves = np.array([t[1] for t in rec.transitions], dtype=np.uint8)
    after = np.concatenate([t[2] for t in rec.transitions])
    np.savez_compressed(os.path.join(HERE, "scenarios.npz"), before=before.view(np.uint8).reshape(-1, 1004),
                        moves=moves, after=after.view(np.uint8).reshape(-1, 1004), names=np.array(names))
    print("scenarios: %d transitions from %d scenarios" % (len(names), len(scenarios.STEP_SCENARIOS)))

    # 2. InitBoardItems
    seeds = oracle.clean_seeds(64)
    boards = np.zeros((64, 11, 11), np.int32)
    for i, sd in enumerate(seeds):
        s = R.zero_state()
        R.init_board_items(s, sd)
        boards[i] = s["board"][0]
    np.savez_compressed(os.path.join(HERE, "init.npz"), seeds=np.array(seeds, np.int32), boards=boards)

    # 3. random traces
    out = {}
    for name, n, ticks, nact, stress, rs in TRACE_CONFIGS:
        S = initial_states(R, n, stress, seeds)
        init = S.copy()
        hashes, stat, excluded, d1 = run_trace(R, O, S, ticks, nact, rs)
        out[name + "_init"] = init.view(np.uint8).reshape(n, 1004)
        out[name + "_final"] = S.view(np.uint8).reshape(n, 1004)
        out[name + "_hash"] = hashes
        out[name + "_status"] = stat
        out[name + "_excluded"] = excluded
        out[name + "_cfg"] = np.array([n, ticks, nact, stress, rs], np.int64)
        print("%s: %d envs x %d ticks, D1 steps %d, excluded envs %d, done %d" %
              (name, n, ticks, d1, int((excluded >= 0).sum()), int((stat[-1] & 1).sum())))
    np.savez_compressed(os.path.join(HERE, "traces.npz"), **out)

    # 4. an on-disk trace in the POMTRC1 format (pomcpp_b200/host/pom_trace.hpp) for pom_replay:
    #    96 envs x 120 ticks, boosted kick/bomb regime, produced by the compiled reference
    write_pomtrc(R, O, os.path.join(HERE, "stress96.pomtrc"), 96, 120, 6, 1, 45, seeds)
    # 5. SimpleAgent games
    simple_agent_games(R, O, seeds)
    for f in ("scenarios.npz", "init.npz", "traces.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
