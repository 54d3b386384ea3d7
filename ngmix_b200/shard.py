"""
Sharding of a batch over ranks (one process per GPU).  Objects are
independent, so the partition is a contiguous index range per rank and there
is NO collective on the hot path; the only communication is an optional
gather of the result records (about 1.2 KB per object) after the fit.
"""
import numpy as np


def shard_range(n, rank, world):
    """contiguous [lo, hi) of rank among world for n objects"""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def shard_arrays(arrays, rank, world):
    """slice every [n, ...] array (None passes through) to this rank's range"""
    n = None
    for a in arrays.values():
        if a is not None:
            n = len(a) if n is None else n
            assert len(a) == n, "all arrays must have the same length"
    lo, hi = shard_range(n, rank, world)
    return {k: (None if a is None else a[lo:hi]) for k, a in arrays.items()}, \
        (lo, hi)


def gather_records(local, n_total, group=None):
    """
    all-gather structured result records from every rank into one [n_total]
    array ordered by global object index.  Works with gloo (CPU tensors) and
    nccl (CUDA tensors); without an initialised process group it returns
    `local` unchanged.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    itemsize = local.dtype.itemsize
    counts = [shard_range(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in counts)
    buf = np.zeros(maxn * itemsize, dtype=np.uint8)
    raw = np.frombuffer(local.tobytes(), dtype=np.uint8)
    buf[:raw.size] = raw
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) \
        if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(buf).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    out = np.zeros(n_total, dtype=local.dtype)
    for r, (lo, hi) in enumerate(counts):
        b = outs[r].cpu().numpy()[:(hi - lo) * itemsize]
        out[lo:hi] = np.frombuffer(b.tobytes(), dtype=local.dtype)
    assert rank >= 0
    return out


def run_sharded(fit_fn, arrays, gather=True, group=None):
    """
    run fit_fn(**shard) on this rank's slice of `arrays` and (optionally)
    gather the structured results of all ranks.  fit_fn is for example
    functools.partial(ngmix_b200.fit_batch, "exp") with keyword arrays
    images=, weights=, jacobians=, psf_pars=, guesses=.
    """
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    shard, (lo, hi) = shard_arrays(arrays, rank, world)
    n_total = max(len(a) for a in arrays.values() if a is not None)
    local = fit_fn(**shard)
    assert len(local) == hi - lo
    if gather:
        return gather_records(local, n_total, group=group)
    return local


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """
    Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so
    that the pinned host buffers it allocates afterwards (first touch) and the
    host side of the H2D pipeline stay local to that GPU's PCIe root.  With
    one process per GPU on a multi-socket host this is what keeps the
    host-buffer calls from sharing one socket's memory bandwidth.  Returns the
    node number, or None when the topology cannot be read (nothing changed).
    """
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        bdf = "%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(
            open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
