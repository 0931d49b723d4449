"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/pom_batch.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import pomcpp_b200 as pb
    return pb.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pom_batch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pom_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libpom_b200.so does not export %s" % n


def test_no_oracle_or_cpu_path_in_product():
    """The product sources never reference the oracle."""
    for d in ("pomcpp_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "libpom_oracle" not in txt and "libpomref" not in txt, f


def test_state_layout_matches_reference_offsets():
    import pomcpp_b200 as pb
    dt = pb.STATE_DT
    assert dt.itemsize == 1004
    assert dt.fields["timeStep"][1] == 484 and dt.fields["aliveAgents"][1] == 488
    assert dt.fields["agents"][1] == 492 and dt.fields["bombs"][1] == 588
    assert dt.fields["bombs_index"][1] == 668 and dt.fields["flames"][1] == 676
    assert dt.fields["flames_index"][1] == 996 and dt.fields["flames_count"][1] == 1000


def test_fails_loudly_without_gpu(lib):
    import pomcpp_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PomError) as e:
        pb.Batch(16)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_rng_is_shared_with_oracle(lib, orc):
    for env in (0, 3, 123456789):
        for tick in (0, 5, 799):
            for na in (5, 6):
                assert lib.pom_rng_moves(42, env, tick, na) == orc.lib.pom_oracle_rng_moves(42, env, tick, na)


def test_examples_build_and_fail_loudly_without_gpu(lib):
    """examples/ compile against include/ and the in-tree library; without a device they stop with the library's error"""
    import subprocess
    import pomcpp_b200 as pb
    ex = os.path.join(ROOT, "examples")
    out = subprocess.run(["make", "-C", ex], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    run = subprocess.run([os.path.join(ex, "actor_loop"), "64", "2"], capture_output=True, text=True, timeout=60)
    assert run.returncode != 0 and "no CPU fallback" in run.stderr
