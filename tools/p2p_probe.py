#!/usr/bin/env python
"""2-GPU probe: can a kernel of THIS library, running on GPU `rank`, store into a buffer that lives on the other GPU and
was mapped through CUDA IPC?  Two ways of mapping it are tried, each in a fresh pair of processes (a fault poisons the
CUDA context):
    torch   torch.multiprocessing.reductions.reduce_tensor / rebuild (opened under the EXPORTER's device index) + torch's
            lazy peer-access enabling through a cross-device copy
    raw     the exporter's storage handle (Storage._share_cuda_) opened with cudaIpcOpenMemHandle(LazyEnablePeerAccess)
            while the IMPORTER's own device is current
The store is done by epnp_mc_lse_backward_f32 (grad_logw = g * exp(logw - lse)) with its output pointing at the peer.
    python tools/p2p_probe.py            # prints one JSON line per mode
"""
import ctypes
import json
import os
import socket
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "epro-pnp_b200"))


def worker(rank, world, port, mode, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from epropnp_b200 import capi, native
    from epropnp_b200.sharded import raw_ipc_export, raw_ipc_open
    res = dict(rank=rank, mode=mode)
    try:
        B, M = 8, 256
        mine = torch.zeros(B, M, device=dev)
        if mode == "torch":
            from torch.multiprocessing.reductions import reduce_tensor
            handle = reduce_tensor(mine)
        else:
            handle = raw_ipc_export(mine)
        everyone = [None] * world
        dist.all_gather_object(everyone, handle)
        other = 1 - rank
        if mode == "torch":
            fn, args = everyone[other]
            remote = fn(*args)
            scratch = torch.zeros(B, M, device=dev)
            scratch.copy_(remote); remote.copy_(scratch)
            torch.cuda.synchronize()
            ptr = remote.data_ptr()
            res["remote_device"] = str(remote.device)
        else:
            ptr = raw_ipc_open(everyone[other], dev)
        dist.barrier()
        logw = torch.randn(B, M, device=dev)
        lse = torch.logsumexp(logw, 1)
        g = torch.full((B,), float(rank + 1), device=dev)
        with torch.cuda.device(dev):
            rc = capi.lib().epnp_mc_lse_backward_f32(capi.ptr(logw), capi.ptr(lse), capi.ptr(g), ctypes.c_void_p(ptr), B, M,
                                                     native.stream_ptr(dev))
        torch.cuda.synchronize()
        res["rc"] = rc
        dist.barrier()
        # what the OTHER rank's kernel wrote into my buffer: rows sum to g_other
        res["row_sums"] = [round(float(v), 4) for v in mine.sum(1)[:3].tolist()]
        res["ok"] = bool(torch.allclose(mine.sum(1), torch.full((B,), float(other + 1), device=dev), atol=1e-3))
    except Exception as ex:                                # noqa: BLE001
        res["error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
    q.put(res)
    try:
        dist.destroy_process_group()
    except Exception:
        pass


def main():
    import torch.multiprocessing as mp
    for mode in ("raw", "torch"):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, 2, port, mode, q)) for r in range(2)]
        for p in procs:
            p.start()
        out = []
        for p in procs:
            p.join(90)
            if p.is_alive():
                p.kill()
                out.append(dict(mode=mode, error="timeout"))
        while not q.empty():
            out.append(q.get())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
