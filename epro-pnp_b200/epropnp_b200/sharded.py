"""Multi-GPU execution of the EPro-PnP path: objects are independent, so the batch is split into
contiguous shards (one process per GPU) and solved with NO collective inside the solve.  The only
communication is one gather of the small results afterwards (poses, optional covariances / costs and
the (B, M) log-weights); the (B, M, D) pose samples stay sharded unless explicitly requested
(SURVEY.md section 8e).  The Philox noise is keyed by the GLOBAL object index (`obj_offset`), so a sharded
run reproduces the single-GPU run bit for bit.
"""
import torch
import torch.distributed as dist


def shard_range(num_obj, rank, world_size):
    """Contiguous [begin, end) of `rank`: sizes differ by at most one, earlier ranks take the extra."""
    base, extra = divmod(int(num_obj), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_sizes(num_obj, world_size):
    return [shard_range(num_obj, r, world_size)[1] - shard_range(num_obj, r, world_size)[0]
            for r in range(world_size)]


def gather_objects(local, num_obj, group=None):
    """All-gather a per-object tensor (local shard (b_r, ...)) into the full (num_obj, ...) tensor on
    every rank.  Equal shards use all_gather_into_tensor (one NCCL call); ragged shards pad to the
    largest shard."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(num_obj, world)
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = local.new_empty((num_obj,) + tail)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    biggest = max(sizes)
    padded = local.new_zeros((biggest,) + tail)
    padded[:local.shape[0]] = local
    buf = local.new_empty((world * biggest,) + tail)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * biggest:r * biggest + sizes[r]] for r in range(world)], dim=0)


def gather_results(result, num_obj, keys=("pose_opt", "logw", "pose_cov", "cost"), group=None):
    """result: dict of local object-major tensors (as returned by native.lm_amis_fused)."""
    return {k: gather_objects(result[k], num_obj, group) for k in keys if result.get(k) is not None}


class PendingGather:
    """Handle of an asynchronous gather: `.wait()` makes the current stream wait for it and returns the dict."""

    def __init__(self, tensors, works):
        self.tensors, self.works = tensors, works

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        return self.tensors


def gather_results_async(result, num_obj, keys=("pose_opt", "logw"), group=None):
    """Same gather, issued asynchronously (equal shards only): NCCL runs it on its own stream after the work
    already enqueued on the current stream, so the NEXT batch's solve overlaps this batch's exchange.  Objects of
    different batches are independent; nothing inside a solve ever waits for a collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return PendingGather({k: result[k] for k in keys if result.get(k) is not None}, [])
    world = dist.get_world_size(group)
    if len(set(shard_sizes(num_obj, world))) != 1:
        return PendingGather(gather_results(result, num_obj, keys, group), [])
    outs, works = {}, []
    for k in keys:
        local = result.get(k)
        if local is None:
            continue
        out = local.new_empty((num_obj,) + tuple(local.shape[1:]))
        works.append(dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=True))
        outs[k] = out
    return PendingGather(outs, works)
