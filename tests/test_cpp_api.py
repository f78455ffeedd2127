"""The C++14 host API (vpp_b200/include/vpp): the reference's own tests, rewritten with device kernels,
are compiled by build.sh into tests/cpp/_build and run on the GPU by tests/test_gpu_parity_late.py (after the Python parity
tests: `pytest -x` then reaches the kernels that did not change since their last hardware run first) — and, without a GPU, compiled by g++ against
the CPU block/warp emulator of tests/emu/ (same headers, same test sources, kernels executed thread by thread)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


def test_cpp_binaries_are_built(built):
    for name in ("core_tests", "algo_tests", "nbh_tests", "extruder_tests"):
        assert os.path.exists(os.path.join(BUILD, name))


@pytest.mark.parametrize("schedule", ["forward", "shuffled"])
@pytest.mark.parametrize("name", ["core_tests", "algo_tests", "nbh_tests", "extruder_tests"])
def test_cpp_binary_on_the_cpu_emulator(built, name, schedule):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    exe = build_emu.build_cpp_test(name)
    env = dict(os.environ, VPPB_EMU_LOG=os.path.join(ROOT, "tests", "emu", "_build", "emu_fail.log"))
    if schedule == "shuffled":  # the threads of every block are visited in a fresh pseudo-random order each scheduler round
        env["VPPB_EMU_SHUFFLE"] = "4242"
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden")], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0 and "ALL OK" in r.stdout and "runtime error" not in r.stderr, r.stdout[-2000:] + r.stderr[-3000:]
