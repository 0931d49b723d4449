"""TEST-ONLY: host build of the kernel body (see hostsim.cpp).  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libhostsim.so")
REC = 292


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(HERE, "hostsim.cpp")
    deps = [src] + [os.path.join(ROOT, "pomcpp_b200", "csrc", f) for f in ("pom_core.cuh", "pom_policy.cuh", "pom_record.h")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) > os.path.getmtime(d) for d in deps):
        return
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-x", "c++",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "pomcpp_b200", "csrc"), "-o", SO, src]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError(out.stderr)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostSim:
    def __init__(self):
        build()
        L = self.lib = C.CDLL(SO)
        vp = C.c_void_p
        L.hostsim_pack_batch.argtypes = [vp, vp, C.c_long, vp, vp]
        L.hostsim_unpack_batch.argtypes = [vp, C.c_long, vp, vp]
        L.hostsim_step_records.argtypes = [vp, C.c_long, vp, C.c_int, vp]
        L.hostsim_spawn_flame.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.hostsim_rng_moves.restype = C.c_uint32
        L.hostsim_rng_moves.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.hostsim_simple_moves.argtypes = [vp, C.c_long, vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp]
        L.hostsim_fog_batch.argtypes = [vp, C.c_long, C.c_int, C.c_int]
        L.hostsim_observe_planes.argtypes = [vp, C.c_long, C.c_int, C.c_int, vp]
        L.hostsim_move_towards.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.hostsim_move_towards_safe_place.argtypes = [vp, C.c_int, C.c_int]
        assert L.hostsim_record_bytes() == REC

    def set_continue_undefined(self, on):
        """True: D3 / D5 ticks keep the env running with the canonical result (POM_STEP_CONTINUE_UNDEFINED)"""
        self.lib.hostsim_set_continue_undefined(int(bool(on)))

    def set_by_rays(self, on):
        """True: step_records runs the tick in the warp-cooperative kernels' decomposition (pomcore::step_by_rays)"""
        self.lib.hostsim_set_by_rays(int(bool(on)))

    def pack(self, S, status=None):
        n = S.shape[0]
        recs = np.zeros((n, REC), np.uint8)
        bad = np.zeros(n, np.uint8)
        self.lib.hostsim_pack_batch(_p(S), None if status is None else _p(status), n, _p(recs), _p(bad))
        return recs, bad

    def unpack(self, recs, S=None):
        n = recs.shape[0]
        import oracle
        if S is None:
            S = np.zeros(n, oracle.STATE_DT)
        status = np.zeros(n, np.uint8)
        self.lib.hostsim_unpack_batch(_p(recs), n, _p(S), _p(status))
        return S, status

    def step_records(self, recs, moves, raw=False, flags=None):
        self.lib.hostsim_step_records(_p(recs), recs.shape[0], _p(moves), int(raw), None if flags is None else _p(flags))

    def spawn_flame(self, rec, x, y, s):
        self.lib.hostsim_spawn_flame(_p(rec), x, y, s)

    def move_towards(self, rec, agent, tx, ty):
        return self.lib.hostsim_move_towards(_p(rec), agent, tx, ty)

    def move_towards_safe_place(self, rec, agent, radius):
        return self.lib.hostsim_move_towards_safe_place(_p(rec), agent, radius)

    def observe_planes(self, recs, agent, view):
        out = np.zeros((recs.shape[0], 512), np.uint8)
        self.lib.hostsim_observe_planes(_p(recs), recs.shape[0], agent, view, _p(out))
        return out

    def observe_cropped(self, recs, agent, view):
        self.lib.hostsim_obs_cropped_bytes.restype = C.c_long
        rb = int(self.lib.hostsim_obs_cropped_bytes(view))
        out = np.zeros((recs.shape[0], rb), np.uint8)
        self.lib.hostsim_observe_cropped(_p(recs), C.c_long(recs.shape[0]), agent, view, _p(out))
        return out

    def fog_batch(self, S, agent, view):
        self.lib.hostsim_fog_batch(_p(S), S.shape[0], agent, view)
        return S

    def simple_moves(self, recs, A, seed, env0, tick, mask, moves):
        self.lib.hostsim_simple_moves(_p(recs), recs.shape[0], _p(A), seed, env0, tick, mask, _p(moves))


class HostSimBackend:
    """scenarios.py backend: fixtures on the host (restatement primitives), Step through the packed record."""

    def __init__(self, orc):
        self.o = orc
        self.h = HostSim()

    def __getattr__(self, name):
        return getattr(self.o, name)

    def step(self, s, moves):
        recs, bad = self.h.pack(s)
        assert bad[0] == 0, "unrepresentable fixture (%d)" % bad[0]
        m = np.asarray(moves, np.uint8).reshape(1, 4)
        self.h.step_records(recs, m, raw=True)
        self.h.unpack(recs, s)
        return 0

    def spawn_flame(self, s, x, y, strength):
        recs, bad = self.h.pack(s)
        assert bad[0] == 0
        self.h.spawn_flame(recs, x, y, strength)
        self.h.unpack(recs, s)
        return 0
