"""csrc/thread_pool.h (the router's persistent workers): correctness of back-to-back jobs, plain and under ThreadSanitizer."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "tp_test.cc")
INC = os.path.join(ROOT, "variantstore_b200", "csrc")


@pytest.mark.parametrize("flags,workers,reps", [(["-O2"], "31", "2000"), (["-O1", "-g", "-fsanitize=thread"], "7", "300")])
def test_thread_pool_jobs(tmp_path, flags, workers, reps):
    exe = str(tmp_path / "tp_test")
    subprocess.run(["g++", "-std=c++17", "-pthread", *flags, "-I", INC, SRC, "-o", exe], check=True)
    r = subprocess.run([exe, workers, "400000", reps], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok") and "FAIL" not in r.stdout
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-2000:]
