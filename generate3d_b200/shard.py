"""Sharding of independent models / views over ranks: the
reference is a serial `for` over models (CADVoxelization.py:15, raytracing.py:
179-191), so rank r simply owns the contiguous block [r*N/G, (r+1)*N/G). No
collective is on the data path; an optional gather of per-rank summaries uses
torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced block of item indices owned by `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    lo = (n_items * rank) // world_size
    hi = (n_items * (rank + 1)) // world_size
    return lo, hi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _gpu_pci_address(device_index):
    """'dddd:bb:dd.f' of a CUDA device as sysfs spells it, or None."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        return "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except (AttributeError, RuntimeError):
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        return dom[-4:] + ":" + rest
    except Exception:                                     # no NVML, no device: nothing to bind to
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs="/sys"):
    """One process per GPU on a multi-socket host: run (and therefore first-touch and pin host buffers) on the
    CPUs of the NUMA node the GPU hangs off, so that the host<->device copies of the host-buffer front end do not
    cross the socket interconnect. Returns {'numa_node', 'cpus'} or None when the topology is unknown, the node
    has no CPU this process may use, or the host has a single node. Never raises."""
    try:
        addr = _gpu_pci_address(device_index)
        if addr is None:
            return None
        with open(os.path.join(sysfs, "bus/pci/devices", addr, "numa_node")) as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed or allowed == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except (OSError, ValueError):
        return None


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op when single-process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def all_reduce_max(value, device=None):
    """Max over ranks of a python float (used for the max-over-ranks step time)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(value, device=None):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_shards(local, n_total, dim0_counts=None):
    """Single optional gather of result shards (dim 0 = models). Returns the full tensor on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size()
    counts = dim0_counts or [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    cmax = max(counts)                                   # all_gather wants equal shapes: pad, then trim
    padded = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    outs = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(outs, padded)
    return torch.cat([o[:c] for o, c in zip(outs, counts)], 0)
