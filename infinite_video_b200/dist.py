"""Multi-GPU plumbing: videos are independent, so they are sharded across ranks in contiguous blocks (one
process per GPU, torchrun) and the per-video outputs are gathered with ONE NCCL all_gather after the
consolidation loop -- never inside it (chunks of a video are sequential)."""
import os

import torch
import torch.distributed as dist


def shard_range(n_videos: int, rank: int, world: int):
    """Contiguous block of videos owned by `rank` (first `n % world` ranks get one extra)."""
    base, rem = divmod(n_videos, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def gather_videos(local: torch.Tensor, n_videos: int, group=None, async_op: bool = False):
    """local[n_local, ...] on every rank -> [n_videos, ...] on every rank, in video order.

    Equal shards (n_videos divisible by the world size) are gathered straight into the result; ragged shards go
    through a padded buffer.  async_op=True (equal shards only) returns (result, work): the collective runs on
    NCCL's stream behind the work already queued on the current stream, and the caller `work.wait()`s before it
    reads the result -- the next chunk's kernels are not held up by the gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return (local, None) if async_op else local
    world = dist.get_world_size(group)
    counts = [shard_range(n_videos, r, world) for r in range(world)]
    most = max(e - s for s, e in counts)
    if all(e - s == most for s, e in counts):
        out = torch.empty((world * most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    if async_op:
        raise ValueError("async gather needs equal shards")
    pad = torch.zeros((most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * most: r * most + (e - s)] for r, (s, e) in enumerate(counts)]
    return torch.cat(parts, 0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device=None):
    if dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and torch.device(device).type == "cuda":
            dist.barrier(device_ids=[torch.device(device).index or 0])
        else:
            dist.barrier()


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off (sysfs), so that pinned host buffers
    allocated afterwards are first-touched on that node and H2D copies do not cross the socket interconnect.
    Returns a small report dict; a no-op (with the reason) where sysfs does not say."""
    rep = {"numa_node": None, "cpus": None}
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        rep["numa_node"] = node
        if node < 0:
            rep["note"] = "sysfs reports no NUMA affinity for this GPU"
            return rep
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            rep["cpus"] = len(allowed)
    except Exception as ex:                      # containers without sysfs access, exotic topologies
        rep["note"] = f"{type(ex).__name__}: {ex}"
    return rep

