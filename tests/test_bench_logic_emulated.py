"""bench.py's GPU arm executed end to end WITHOUT a GPU (tests/emu/run_bench_emulated.py): the C-ABI calls go to the CPU
emulator and the few torch.cuda pieces the bench uses (streams, events, graphs, device tensors) are replaced by inert
CPU stand-ins; at N = 2 two processes form a gloo group, so the halo pipeline really exchanges rows.  Nothing is
measured here - the point is that every line of the bench's control flow (launch-form selection with its child probe,
graph capture per form, the one-step-ahead exchange, parity checks against the oracle, roofline / e2e bookkeeping, the
JSON line) has run before it meets a GPU box, on a tiny geometry."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "emu", "run_bench_emulated.py")
KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
        "roofline", "e2e", "clocks", "gpu_launches", "parity_checked")


def _run(probe_rc, rows, cols, bench_args, env=None):
    return subprocess.Popen([sys.executable, RUNNER, str(probe_rc), str(rows), str(cols), "--"] + bench_args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                            text=True, cwd=ROOT, env=env or dict(os.environ))


@pytest.mark.parametrize("box_launch,probe_rc", [("auto", 0), ("auto", 1), ("per-frame", 0), ("batch", 0)])
def test_bench_single_gpu_control_flow(built, box_launch, probe_rc):
    p = _run(probe_rc, 48, 400, ["--steps", "3", "--warmup", "3", "--frames", "5", "--no-extras", "--cpu-budget", "0.2", "--box-launch", box_launch])
    out, err = p.communicate(timeout=600)
    assert p.returncode == 0, err[-3000:]
    line = json.loads(out.strip().splitlines()[-1])
    for key in KEYS + ("cpu_baseline",):
        assert key in line, key
    assert line["parity_checked"] is True
    bl = line["config"]["box_launch"]
    probes = [l for l in err.splitlines() if l.startswith("PROBE ")]
    if box_launch == "auto":
        assert probes == ["PROBE 0 48 400 5"]
        if probe_rc:
            assert bl["used"] == "per-frame" and list(bl["ms_per_step_by_mode"]) == ["per-frame"]
        else:
            assert set(bl["ms_per_step_by_mode"]) == {"per-frame", "batch"}
    elif box_launch == "per-frame":
        assert bl["used"] == "per-frame" and probes == ["PROBE 0 48 400 5"]  # the probe still runs: it also vouches for the staged upload
    else:
        assert bl["used"] == box_launch and not probes
    assert line["gpu_launches"] == 3 * (5 if bl["used"] == "per-frame" else 1)
    assert line["roofline"]["kernel"] == ("k_box5_bytes_tma_batch<3>" if bl["used"] == "batch" else "k_box5_bytes_tma<3>")
    assert line["e2e"]["h2d_bytes_per_step"] == 5 * 48 * 400 * 3
    # the staged upload is tried only when the probe ran and vouched for vppb_copy2d_mirror (rc 0; --box-launch batch skips the probe)
    forms = {"direct", "staged"} if (probe_rc == 0 and box_launch != "batch") else {"direct"}
    assert set(line["e2e"]["upload"]["ms_per_step_by_form"]) == forms and line["e2e"]["upload"]["used"] in forms


@pytest.mark.parametrize("box_launch", ["auto", "per-frame"])
def test_bench_two_ranks_control_flow(built, box_launch):
    """world_size 2 over gloo: each rank filters its row tile of every frame after the grouped halo exchange; both ranks
    check their tile against the oracle (parity_checked is the AND over ranks).  --graph 0: the stand-in CUDA graph cannot
    replay work, and at N > 1 the unpack / box pieces must really run every step for the halos to be current."""
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(_run(0, 64, 400, ["--gpus", "2", "--steps", "3", "--warmup", "3", "--frames", "4", "--no-extras", "--graph", "0", "--box-launch", box_launch], env))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, err[-3000:]
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert outs[1][0].strip() == ""  # rank 0 alone prints
    for key in KEYS:
        assert key in line, key
    assert line["n_gpus"] == 2 and line["parity_checked"] is True
    bl = line["config"]["box_launch"]
    assert bl["used"] in ("per-frame", "batch")
    if box_launch == "auto":
        assert set(bl["ms_per_step_by_mode"]) == {"per-frame", "batch"}
        assert any(l.startswith("PROBE 1 32 400 4") for l in outs[1][1].splitlines())  # rank 1 probed its own device and tile


def test_probe_child_on_the_emulator(built):
    """the child process of the launch-form selection (bench.py --probe-batch): batched box == per-frame box, fused copy + mirror ==
    upload + mirror fill, on the emulated library -> exit code 0"""
    p = _run(0, 48, 400, ["--probe-batch", "0", "48", "400", "5"])
    out, err = p.communicate(timeout=600)
    assert p.returncode == 0, err[-3000:]
