"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the CPU checkers.

``restatement()`` -> oracle/_build/libpom_oracle.so  (plain-C restatement: oracle/pom_oracle.c, the step path;
                                                     oracle/pom_oracle_agent.c, SimpleAgent + strategy)
``reference()``   -> oracle/_ref/libpomref.so        (the UNMODIFIED reference sources bboard / step / step_utility /
                                                     strategy / simple_agent + oracle/ref_shim.cpp)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  The product package ``pomcpp_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# numpy view of include/pom_state.h (== reference bboard::State, 1004 bytes)
AGENT_DT = np.dtype([("x", "<i4"), ("y", "<i4"), ("bombCount", "<i4"), ("maxBombCount", "<i4"),
                     ("bombStrength", "<i4"), ("canKick", "u1"), ("dead", "u1"), ("_pad", "u1", (2,))])
FLAME_DT = np.dtype([("x", "<i4"), ("y", "<i4"), ("timeLeft", "<i4"), ("strength", "<i4")])
STATE_DT = np.dtype([("board", "<i4", (11, 11)), ("timeStep", "<i4"), ("aliveAgents", "<i4"),
                     ("agents", AGENT_DT, (4,)), ("bombs", "<i4", (20,)), ("bombs_index", "<i4"),
                     ("bombs_count", "<i4"), ("flames", FLAME_DT, (20,)), ("flames_index", "<i4"),
                     ("flames_count", "<i4")])
assert STATE_DT.itemsize == 1004
# include/pom_state.h pom_simple_agent (persistent members of agents::SimpleAgent), 8 bytes
SIMPLE_DT = np.dtype([("recent", "u1", (4,)), ("rp_index", "u1"), ("rp_count", "u1"), ("move_queue", "<u2")])
assert SIMPLE_DT.itemsize == 8

_vp = C.c_void_p
_u8p = C.c_void_p


def build(quiet=True):
    """Compile the restatement (always) and oracle/_ref (only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Restatement:
    """oracle/pom_oracle.c"""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libpom_oracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.pom_oracle_state_hash.restype = C.c_uint64
        L.pom_oracle_rng_moves.restype = C.c_uint32
        L.pom_oracle_rng_moves.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.pom_oracle_rng_moves_batch.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.c_uint32, C.c_uint32, _vp]
        L.pom_oracle_diff_batch.restype = C.c_long
        L.pom_oracle_diff_batch.argtypes = [_vp, _vp, C.c_long, _vp, _vp]
        L.pom_oracle_env_step_batch.argtypes = [_vp, _vp, C.c_long, _vp, _vp]
        L.pom_oracle_step_batch.argtypes = [_vp, C.c_long, _vp, _vp]
        L.pom_oracle_hash_batch.argtypes = [_vp, C.c_long, _vp]
        L.pom_oracle_bench_steps.restype = C.c_double
        L.pom_oracle_bench_steps.argtypes = [_vp, _vp, C.c_long, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]
        for f in ("pom_oracle_step", "pom_oracle_env_step"):
            getattr(L, f).restype = C.c_int

    def set_continue_undefined(self, on):
        """D3 / D5 ticks keep the env running with the canonical result (POM_STEP_CONTINUE_UNDEFINED)"""
        self.lib.pom_oracle_set_continue_undefined(int(bool(on)))

    # --- single-state helpers (fixtures) ---
    def zero_state(self, n=None):
        s = np.zeros(1 if n is None else n, dtype=STATE_DT)
        for i in range(s.shape[0]):
            self.lib.pom_oracle_zero_state(_ptr(s[i:i + 1]))
        return s

    def step(self, s, moves):
        m = np.asarray(moves, dtype=np.uint8)
        return self.lib.pom_oracle_step(_ptr(s), _ptr(m))

    def put_agent(self, s, x, y, i): self.lib.pom_oracle_put_agent(_ptr(s), x, y, i)
    def put_agents_in_corners(self, s, a0=0, a1=1, a2=2, a3=3): self.lib.pom_oracle_put_agents_in_corners(_ptr(s), a0, a1, a2, a3)
    def kill(self, s, *ids):
        for i in ids:
            self.lib.pom_oracle_kill(_ptr(s), i)
    def plant_bomb(self, s, x, y, i, set_item=False, life=10): self.lib.pom_oracle_plant_bomb(_ptr(s), x, y, i, life, int(set_item))
    def spawn_flame(self, s, x, y, strength): return self.lib.pom_oracle_spawn_flame(_ptr(s), x, y, strength)
    def put_item(self, s, x, y, item): s["board"][0, y, x] = item

    def set_bomb_direction(self, s, logical, d):
        slot = (int(s["bombs_index"][0]) + logical) % 20
        b = int(s["bombs"][0, slot])
        s["bombs"][0, slot] = (b & ~0xF00000) + (d << 20)

    def init_state(self, s, seed=0x1337, order=(0, 1, 2, 3)):
        return self.lib.pom_oracle_init_state(_ptr(s), C.c_int(seed), *order)

    def init_board_items(self, s, seed=0x1337):
        return self.lib.pom_oracle_init_board_items(_ptr(s), C.c_int(seed))

    def fill_dest_pos(self, s, moves):
        out = np.zeros(8, dtype=np.int32)
        m = np.asarray(moves, dtype=np.uint8)
        self.lib.pom_oracle_fill_dest_pos(_ptr(s), _ptr(m), _ptr(out))
        return out

    def fix_switch_move(self, s, pos8):
        p = np.array(pos8, dtype=np.int32)
        self.lib.pom_oracle_fix_switch_move(_ptr(s), _ptr(p))
        return p

    def resolve_dependencies(self, s, pos8):
        p = np.array(pos8, dtype=np.int32)
        dep = np.zeros(4, dtype=np.int32)
        roots = np.zeros(4, dtype=np.int32)
        n = self.lib.pom_oracle_resolve_dependencies(_ptr(s), _ptr(p), _ptr(dep), _ptr(roots))
        return n, dep, roots

    # --- batch helpers ---
    def env_step_batch(self, S, status, moves, flags=None):
        assert S.dtype == STATE_DT and status.dtype == np.uint8 and moves.dtype == np.uint8
        self.lib.pom_oracle_env_step_batch(_ptr(S), _ptr(status), S.shape[0], _ptr(moves),
                                           None if flags is None else _ptr(flags))

    def step_batch(self, S, moves, flags=None):
        self.lib.pom_oracle_step_batch(_ptr(S), S.shape[0], _ptr(moves), None if flags is None else _ptr(flags))

    def diff_batch(self, A, B, skip=None):
        why = C.c_int(0)
        e = self.lib.pom_oracle_diff_batch(_ptr(A), _ptr(B), A.shape[0], None if skip is None else _ptr(skip), C.byref(why))
        return int(e), why.value

    def hash_batch(self, S):
        out = np.zeros(S.shape[0], dtype=np.uint64)
        self.lib.pom_oracle_hash_batch(_ptr(S), S.shape[0], _ptr(out))
        return out

    def rng_moves(self, seed, env0, n, tick, n_actions=6):
        out = np.zeros((n, 4), dtype=np.uint8)
        self.lib.pom_oracle_rng_moves_batch(seed, env0, n, tick, n_actions, _ptr(out))
        return out

    def fog_batch(self, S, agent, view=4):
        """in place: every State as `agent` observes it (pom_oracle_fog)"""
        self.lib.pom_oracle_fog_batch(_ptr(S), C.c_long(S.shape[0]), agent, view)
        return S

    def observe_planes_batch(self, S, agent, view=4):
        out = np.zeros((S.shape[0], 512), np.uint8)
        self.lib.pom_oracle_observe_planes_batch(_ptr(S), C.c_long(S.shape[0]), agent, view, _ptr(out))
        return out

    def obs_cropped_bytes(self, view=4):
        self.lib.pom_oracle_obs_cropped_bytes.restype = C.c_long
        return int(self.lib.pom_oracle_obs_cropped_bytes(view))

    def observe_cropped_batch(self, S, agent, view=4):
        out = np.zeros((S.shape[0], self.obs_cropped_bytes(view)), np.uint8)
        self.lib.pom_oracle_observe_cropped_batch(_ptr(S), C.c_long(S.shape[0]), agent, view, _ptr(out))
        return out

    # --- agents::SimpleAgent / bboard::strategy (oracle/pom_oracle_agent.c) ---
    def simple_agents(self, n_envs):
        return np.zeros((n_envs, 4), dtype=SIMPLE_DT)

    def simple_act(self, s, agent_id, st, draw):
        """st: one SIMPLE_DT element (array of shape (1,))"""
        return self.lib.pom_oracle_simple_act(_ptr(s), agent_id, _ptr(st), int(draw))

    def simple_moves_batch(self, S, status, A, seed, env0, tick, agent_mask, moves):
        self.lib.pom_oracle_simple_moves_batch(_ptr(S), None if status is None else _ptr(status), C.c_long(S.shape[0]),
                                               _ptr(A), C.c_uint64(seed), C.c_uint64(env0), C.c_uint32(tick),
                                               C.c_uint(agent_mask), _ptr(moves))

    def is_adjacent_enemy(self, s, agent_id, dist): return bool(self.lib.pom_oracle_is_adjacent_enemy(_ptr(s), agent_id, dist))
    def is_in_danger(self, s, x, y): return self.lib.pom_oracle_is_in_danger(_ptr(s), x, y)

    def fill_rmap(self, s, agent_id):
        out = np.zeros((11, 11), dtype=np.int32)
        self.lib.pom_oracle_fill_rmap(_ptr(s), agent_id, _ptr(out))
        return out

    def move_towards(self, s, agent_id, kind, a, b=0): return self.lib.pom_oracle_move_towards(_ptr(s), agent_id, kind, a, b)

    def bench_steps(self, S, status, moves, nthreads, templates=None):
        steps = C.c_ulonglong(0)
        t = self.lib.pom_oracle_bench_steps(_ptr(S), _ptr(status), S.shape[0], _ptr(moves), moves.shape[0], nthreads,
                                            None if templates is None else _ptr(templates),
                                            0 if templates is None else templates.shape[0], C.byref(steps))
        return t, steps.value


class Reference:
    """oracle/_ref/libpomref.so: the unmodified reference behind oracle/ref_shim.cpp."""

    def __init__(self, flavour="", stem="libpomref"):
        path = os.path.join(HERE, "_ref", "%s%s.so" % (stem, flavour))
        if not os.path.exists(path) and os.path.exists("/root/reference/src/bboard/step.cpp"):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        assert L.ref_sizeof_state() == 1004
        L.ref_env_step_batch.argtypes = [_vp, _vp, C.c_long, _vp, _vp, _vp]
        L.ref_step_batch.argtypes = [_vp, C.c_long, _vp]
        L.ref_bench_steps.restype = C.c_double
        L.ref_bench_steps.argtypes = [_vp, _vp, C.c_long, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]

    def zero_state(self, n=None):
        s = np.zeros(1 if n is None else n, dtype=STATE_DT)
        for i in range(s.shape[0]):
            self.lib.ref_zero_state(_ptr(s[i:i + 1]))
        return s

    def step(self, s, moves):
        m = np.asarray(moves, dtype=np.uint8)
        self.lib.ref_step(_ptr(s), _ptr(m))
        return 0

    def precheck(self, s, moves):
        m = np.asarray(moves, dtype=np.uint8)
        return self.lib.ref_precheck(_ptr(s), _ptr(m))

    def put_agent(self, s, x, y, i): self.lib.ref_put_agent(_ptr(s), x, y, i)
    def put_agents_in_corners(self, s, a0=0, a1=1, a2=2, a3=3): self.lib.ref_put_agents_in_corners(_ptr(s), a0, a1, a2, a3)
    def kill(self, s, *ids):
        for i in ids:
            self.lib.ref_kill(_ptr(s), i)
    def plant_bomb(self, s, x, y, i, set_item=False): self.lib.ref_plant_bomb(_ptr(s), x, y, i, int(set_item))
    def spawn_flame(self, s, x, y, strength): self.lib.ref_spawn_flame(_ptr(s), x, y, strength); return 0
    def put_item(self, s, x, y, item): self.lib.ref_put_item(_ptr(s), x, y, item)
    def set_bomb_direction(self, s, logical, d): self.lib.ref_set_bomb_direction(_ptr(s), logical, d)

    def init_state(self, s, seed=0x1337, order=(0, 1, 2, 3)):
        self.lib.ref_init_board_items(_ptr(s), C.c_int(seed))
        self.lib.ref_put_agents_in_corners(_ptr(s), *order)
        return 0

    def init_board_items(self, s, seed=0x1337):
        self.lib.ref_init_board_items(_ptr(s), C.c_int(seed))
        return 0

    def fill_dest_pos(self, s, moves):
        out = np.zeros(8, dtype=np.int32)
        m = np.asarray(moves, dtype=np.uint8)
        self.lib.ref_fill_dest_pos(_ptr(s), _ptr(m), _ptr(out))
        return out

    def fix_switch_move(self, s, pos8):
        p = np.array(pos8, dtype=np.int32)
        self.lib.ref_fix_switch_move(_ptr(s), _ptr(p))
        return p

    def resolve_dependencies(self, s, pos8):
        p = np.array(pos8, dtype=np.int32)
        dep = np.zeros(4, dtype=np.int32)
        roots = np.zeros(4, dtype=np.int32)
        n = self.lib.ref_resolve_dependencies(_ptr(s), _ptr(p), _ptr(dep), _ptr(roots))
        return n, dep, roots

    def env_step_batch(self, S, status, moves, pre=None, exclude=None):
        """exclude: uint8 mask of envs NOT to step this tick (marked invalid instead) — the harness sets it
        where the restatement detected defect D5 (the reference would hang there)."""
        self.lib.ref_env_step_batch(_ptr(S), _ptr(status), S.shape[0], _ptr(moves), None if pre is None else _ptr(pre),
                                    None if exclude is None else _ptr(exclude))

    def step_batch(self, S, moves):
        self.lib.ref_step_batch(_ptr(S), S.shape[0], _ptr(moves))

    def bench_steps(self, S, status, moves, nthreads, templates=None):
        steps = C.c_ulonglong(0)
        t = self.lib.ref_bench_steps(_ptr(S), _ptr(status), S.shape[0], _ptr(moves), moves.shape[0], nthreads,
                                     None if templates is None else _ptr(templates),
                                     0 if templates is None else templates.shape[0], C.byref(steps))
        return t, steps.value

    def hardware_concurrency(self):
        return self.lib.ref_hardware_concurrency()

    def set_fence(self, on):
        """the per-tick FenceTick of the timed loops on/off (only ever off on traces where it cannot fire)"""
        self.lib.ref_set_fence(int(bool(on)))

    # --- agents::SimpleAgent / bboard::strategy, unmodified reference code ---
    class SimpleAgents:
        """n_envs x 4 reference SimpleAgent objects (zero-initialised memory, engine re-seeded per act)."""

        def __init__(self, lib, n_envs):
            self.lib, self.n = lib, 4 * n_envs
            lib.ref_simple_new.restype = C.c_void_p
            lib.ref_simple_new.argtypes = [C.c_long]
            lib.ref_simple_free.argtypes = [C.c_void_p, C.c_long]
            lib.ref_simple_act.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_int]
            lib.ref_simple_export.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
            lib.ref_simple_moves_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
            self.mem = lib.ref_simple_new(self.n)

        def __del__(self):
            if getattr(self, "mem", None):
                self.lib.ref_simple_free(self.mem, self.n)
                self.mem = None

        def act(self, env, agent_id, s, draw):
            return self.lib.ref_simple_act(self.mem, 4 * env + agent_id, _ptr(s), agent_id, int(draw))

        def moves_batch(self, S, status, draws, agent_mask, moves):
            self.lib.ref_simple_moves_batch(_ptr(S), None if status is None else _ptr(status), S.shape[0], self.mem,
                                            _ptr(draws), agent_mask, _ptr(moves))

        def export(self):
            out = np.zeros(self.n, dtype=SIMPLE_DT)
            for i in range(self.n):
                self.lib.ref_simple_export(self.mem, i, _ptr(out[i:i + 1]))
            return out.reshape(-1, 4)

    def simple_agents(self, n_envs):
        return Reference.SimpleAgents(self.lib, n_envs)

    def is_adjacent_enemy(self, s, agent_id, dist): return bool(self.lib.ref_is_adjacent_enemy(_ptr(s), agent_id, dist))
    def is_in_danger(self, s, x, y): return self.lib.ref_is_in_danger(_ptr(s), x, y)

    def fill_rmap(self, s, agent_id):
        out = np.zeros((11, 11), dtype=np.int32)
        self.lib.ref_fill_rmap(_ptr(s), agent_id, _ptr(out))
        return out

    def move_towards(self, s, agent_id, kind, a, b=0): return self.lib.ref_move_towards(_ptr(s), agent_id, kind, a, b)


_cache = {}


def restatement():
    if "r" not in _cache:
        _cache["r"] = Restatement()
    return _cache["r"]


def reference(flavour=""):
    k = "ref" + flavour
    if k not in _cache:
        _cache[k] = Reference(flavour)
    return _cache[k]


def dropin():
    """oracle/_ref/libpomdrop.so: the same shim + the unmodified reference simple_agent.cpp compiled against THIS repo's
    include/ and linked with the product's host layer.  Host-only entry points (agents, strategy, util, field setters)
    work anywhere; the stepping ones run on the GPU through the bboard mirror."""
    if "drop" not in _cache:
        _cache["drop"] = Reference(stem="libpomdrop")
    return _cache["drop"]


def have_dropin():
    try:
        dropin()
        return True
    except (FileNotFoundError, OSError):
        return False


def have_reference():
    try:
        reference()
        return True
    except (FileNotFoundError, OSError):
        return False


def clean_seeds(n, start=0x1337):
    """First n seeds >= start for which InitBoardItems never draws q.count (defect D2)."""
    r = restatement()
    s = r.zero_state()
    out = []
    seed = start
    while len(out) < n:
        s[:] = r.zero_state()
        if r.init_board_items(s, seed) == 0:
            out.append(seed)
        seed += 1
    return out
