"""CPU test (-m "not gpu"): the kernel body, the device policy code and the restatements under ASan + UBSan
(tests/hostsim/sanitize_main.cpp).
compute-sanitizer is closed on the GPU pool, so this is the out-of-bounds / UB check of the tick code itself."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_kernel_body_and_restatement_under_asan_ubsan(tmp_path):
    exe = str(tmp_path / "sanitize_main")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "pomcpp_b200", "csrc"),
           "-I", os.path.join(ROOT, "oracle"), "-x", "c++", os.path.join(HERE, "hostsim", "sanitize_main.cpp"),
           "-x", "c", os.path.join(ROOT, "oracle", "pom_oracle.c"), os.path.join(ROOT, "oracle", "pom_oracle_agent.c"),
           "-o", exe, "-lpthread"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr and "cannot find" in b.stderr:
        pytest.skip("libasan/libubsan not installed")
    assert b.returncode == 0, b.stderr[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = subprocess.run([exe, "384", "160"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "clean" in r.stdout
