"""TEST INFRASTRUCTURE ONLY.  Run bench.py's GPU arm in this process with the C-ABI served by the CPU emulator and inert
stand-ins for the torch.cuda objects it touches (tests/test_bench_logic_emulated.py).  Under RANK / WORLD_SIZE > 1 the
NCCL process group becomes a gloo group on CPU tensors, so the row-tile pipeline (pack -> grouped send/recv -> unpack ->
boxes, issued one step ahead) really exchanges halos between the processes.
    python tests/emu/run_bench_emulated.py <probe_rc> <rows> <cols> -- <bench.py arguments>"""
import contextlib
import ctypes as C
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


class Stream:
    def __init__(self, device=None):
        self.cuda_stream = 0

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class Event:  # the emulated kernels run synchronously on the CPU: wall time between two records is their time
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class Graph:  # "capture" remembers nothing: install() makes bench.graph_of fall back to the eager step
    def replay(self):
        raise RuntimeError("the stand-in graph cannot replay work")


def install(probe_rc, rows, cols):
    import torch
    import build_emu

    emu = C.CDLL(build_emu.build())
    import vpp_b200  # noqa: F401
    from vpp_b200 import capi, ops

    for name, (res, args) in capi.PROTOTYPES.items():
        fn = getattr(emu, name)
        fn.restype, fn.argtypes = res, args
    capi.lib = ops.lib = emu
    cur = Stream()
    torch.cuda.set_device = lambda d: None
    torch.cuda.current_stream = lambda device=None: cur
    torch.cuda.synchronize = lambda device=None: None
    torch.cuda.Stream, torch.cuda.Event, torch.cuda.CUDAGraph = Stream, Event, Graph
    def no_capture(g, **kw):
        raise RuntimeError("no graph capture on the emulator")

    torch.cuda.graph = no_capture
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    real_empty, real_tensor = torch.empty, torch.tensor

    def empty_like_cuda(*a, **kw):  # cudaMalloc'ed tensors are 512-byte aligned; vppb_wrap insists on the row alignment
        kw = {k: v for k, v in kw.items() if k != "device"}
        if len(a) == 1 and isinstance(a[0], int) and kw.get("dtype") == torch.uint8:
            raw = real_empty(a[0] + 512, **kw)
            off = (-raw.data_ptr()) % 512
            return raw[off:off + a[0]]
        return real_empty(*a, **kw)

    torch.empty = empty_like_cuda
    torch.tensor = lambda *a, **kw: real_tensor(*a, **{k: v for k, v in kw.items() if k != "device"})
    torch.Tensor.pin_memory = lambda self, *a, **kw: self
    import torch.distributed as dist

    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo", **{k: v for k, v in kw.items() if k != "device_id"})
    bench = importlib.import_module("bench")
    for k in bench.WORKLOADS:
        bench.WORKLOADS[k] = (rows, cols)
    real_run = bench.subprocess.run

    def fake_run(cmd, *a, **kw):  # the child probe of --box-launch auto needs a GPU: answer in its place, remember the request
        if "--probe-batch" in cmd:
            sys.stderr.write("PROBE " + " ".join(cmd[-4:]) + "\n")
            return types.SimpleNamespace(returncode=probe_rc, stdout="", stderr="probe stand-in")
        return real_run(cmd, *a, **kw)

    bench.subprocess.run = fake_run
    return bench


if __name__ == "__main__":
    i = sys.argv.index("--")
    probe_rc, rows, cols = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    bench = install(probe_rc, rows, cols)
    sys.argv = ["bench.py"] + sys.argv[i + 1:]
    sys.exit(bench.main())
