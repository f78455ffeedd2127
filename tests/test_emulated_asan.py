"""The emulated parity suite once more, under AddressSanitizer: the emulated library is rebuilt with -fsanitize=address
and every image / workspace is an exact-size heap block, so ANY global-memory access of ANY kernel that leaves its
allocation (a stencil tap past the border, a bilinear footprint outside the frame, a TMA box beyond the tensor, a
keypoint buffer overrun) is reported - on the GPU such reads inside cudaMalloc's granularity silently return garbage.
Runs in a subprocess because libasan has to be preloaded into the interpreter."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suite_under_address_sanitizer(built, tmp_path):
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan.so not available")
    log = str(tmp_path / "asan")
    env = dict(os.environ, VPPB_EMU_ASAN="1", VPPB_EMU_STACK_KB="96", LD_PRELOAD=libasan,
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:log_path=%s" % log)
    env.pop("UBSAN_OPTIONS", None)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emulated_parity.py"), "-q", "-x", "-k", "forward",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=1500)
    reports = "".join(open(f).read() for f in glob.glob(log + "*"))
    errors = [l for l in reports.splitlines() if "ERROR: AddressSanitizer" in l or "SUMMARY: AddressSanitizer" in l]
    assert r.returncode == 0 and not errors, r.stdout[-3000:] + r.stderr[-2000:] + reports[:4000]
    assert " passed" in r.stdout
