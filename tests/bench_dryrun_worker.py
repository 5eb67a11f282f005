"""Helper of tests/test_bench_dryrun_cpu.py (not a test): one RANK of a multi-process dry run of bench.py on the CPU.
Applies the same stand-ins as the single-process dry run (SIMT-emulated library, inert streams / events), swaps NCCL
for gloo and the CUDA-IPC gathers for their shared-memory host twins, then runs bench.main() with the given flags."""
import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT, os.path.join(ROOT, "epro-pnp_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import simt_native  # noqa: E402
from epropnp_b200 import capi, native, sharded  # noqa: E402
from test_bench_dryrun_cpu import _Event, _Stream, toy_configs  # noqa: E402
from test_push_gather_cpu import _host_class  # noqa: E402


def main():
    torch.multiprocessing.set_sharing_strategy("file_system")
    capi._lib = simt_native.handle(())
    native._need_cuda = lambda t, what: None
    native.stream_ptr = lambda device=None: None
    torch.cuda.device = lambda device=None: contextlib.nullcontext()
    torch.cuda.synchronize = lambda device=None: None
    torch.cuda.set_device = lambda d: None
    torch.cuda.Stream = lambda *a, **k: _Stream()
    torch.cuda.Event = _Event
    torch.cuda.current_stream = lambda device=None: _Stream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.empty = (lambda f: (lambda *a, pin_memory=False, **k: f(*a, **k)))(torch.empty)
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo", **{k: v for k, v in kw.items() if k != "device_id"})
    sharded.PushGather = _host_class(sharded.PushGather)
    os.environ.update(EPNP_BENCH_DEVICE="cpu", EPNP_NO_SAMPLER="1", LOCAL_RANK="0")
    import bench
    toy_configs(bench, lambda d, k, v: d.__setitem__(k, v))
    bench.WARM_SECONDS, bench.L2_BYTES = 0.05, 1.0
    sys.argv = ["bench.py", "--batch", "2", "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"] + sys.argv[1:]
    bench.main()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
