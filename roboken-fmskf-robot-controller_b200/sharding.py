"""Multi-GPU plumbing: contiguous batch slices, no data-path collective.

Robot instances are independent (the only cross-talk in the firmware is inside one robot:
IMU yaw -> vehicle, VD_task_main.cpp:368), so G GPUs run G disjoint contiguous slices of the
instance index space, one process per GPU.  torch.distributed is used for exactly three
things: the start barrier, the max-over-ranks of the device-timed duration, and the optional
end-of-rollout gather of one cost scalar per instance (NCCL all_gather over NVLink on GPU,
gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_range(n_total, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; the first n_total % world ranks get one more."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if (backend or "nccl") == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend or "nccl", rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers it allocates
    next (first touch) and the threads that drive the copies are local to the GPU's PCIe root.  Returns the node, or
    None when the topology is not exposed (single socket, containers without sysfs): nothing is changed then."""
    try:
        bus = torch.cuda.get_device_properties(local_rank)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (the job's duration is its slowest rank's)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_costs(cost):
    """All ranks' per-instance cost vectors concatenated in rank order.  Slices may differ in size by one
    (shard_range gives the first n_total % world ranks one more instance): every rank pads to the largest
    slice for the collective and the padding is trimmed afterwards."""
    if not dist.is_initialized():
        return cost
    world = dist.get_world_size()
    sizes = torch.zeros(world, dtype=torch.int64, device=cost.device)
    sizes[dist.get_rank()] = cost.numel()
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    sizes = [int(x) for x in sizes.tolist()]
    m = max(sizes)
    if all(x == m for x in sizes):
        out = torch.empty(world * m, dtype=cost.dtype, device=cost.device)
        dist.all_gather_into_tensor(out, cost.contiguous())
        return out
    padded = torch.zeros(m, dtype=cost.dtype, device=cost.device)
    padded[: cost.numel()] = cost
    out = torch.empty(world * m, dtype=cost.dtype, device=cost.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * m : r * m + sizes[r]] for r in range(world)])
