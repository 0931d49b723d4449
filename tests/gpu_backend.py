"""scenarios.py backend that runs Step (and SpawnFlame) on the GPU through the C ABI.

Fixtures are assembled on the host with the restatement's trivial setters (PutAgent, PutItem, Kill,
PlantBomb are plain field writes in the reference too); every bboard::Step and State::SpawnFlame
goes upload -> CUDA kernel -> download."""
import numpy as np

import pomcpp_b200 as pb


class GpuBackend:
    def __init__(self, orc):
        self.o = orc
        self.b = pb.Batch(1, n_templates=1, empty=True)
        self.moves_dev = self.b.alloc(4)

    def __getattr__(self, name):
        return getattr(self.o, name)

    def step(self, s, moves):
        self.b.upload(s)
        m = np.asarray(moves, np.uint8).reshape(1, 4)
        self.b.step_host(np.ascontiguousarray(m), None, pb.STEP_RAW)
        out, st = self.b.download()
        assert not (st[0] & pb.STATUS_INVALID)
        s[:] = out
        return 0

    def spawn_flame(self, s, x, y, strength):
        self.b.upload(s)
        self.b.spawn_flame(0, x, y, strength)
        out, _ = self.b.download()
        s[:] = out
        return 0
