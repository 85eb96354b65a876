"""Best-effort NUMA placement of a rank next to its GPU (host side of the end-to-end path).

One process per GPU: its lane threads and, above all, its PINNED staging buffers should live on the socket
the GPU's PCIe root hangs off; otherwise every H2D / D2H copy of every rank crosses the inter-socket link and
all ranks share node 0's memory controllers (round 1: end-to-end efficiency 0.24 at 8 GPUs, every rank on
node 0).  Call `bind_to_gpu(local_rank)` BEFORE allocating pinned memory: cudaHostAlloc places pages on the
calling thread's node under the default (local) policy, so CPU affinity is what steers it.

Nothing here is required for correctness; every step reports what it could and could not do.
"""
import ctypes
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_bus_id(index):
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def gpu_numa_node(index):
    bus = gpu_pci_bus_id(index)
    if not bus:
        return None
    v = _read("/sys/bus/pci/devices/%s/numa_node" % bus.lower())
    try:
        n = int(v)
    except (TypeError, ValueError):
        return None
    return n if n >= 0 else None


def node_cpus(node):
    return _parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % node))


def n_nodes():
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        return 1


def _set_mempolicy_preferred(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) through the raw syscall (x86-64: 238); returns errno or 0"""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        return 0 if r == 0 else ctypes.get_errno()
    except Exception:
        return -1


def bind_to_gpu(local_rank, local_world=1):
    """Pin this process to the CPUs of the NUMA node of GPU `local_rank` (those this process is allowed to use)
    and prefer that node's memory.  If the allowed CPU set does not reach the GPU's node, fall back to an even
    split of the allowed CPUs across the local ranks so that ranks at least do not share cores.
    Returns a dict describing the outcome (goes into bench.py's JSON line)."""
    info = {"gpu": local_rank, "nodes": n_nodes()}
    allowed = os.sched_getaffinity(0)
    info["allowed_cpus"] = len(allowed)
    node = gpu_numa_node(local_rank)
    info["gpu_node"] = node
    target = None
    if node is not None:
        cpus = node_cpus(node) & allowed
        if cpus:
            # ranks whose GPUs share a node split that node's CPUs
            peers = [r for r in range(local_world) if gpu_numa_node(r) == node] or [local_rank]
            mine = sorted(cpus)
            k = peers.index(local_rank) if local_rank in peers else 0
            share = mine[k::len(peers)] if len(mine) >= len(peers) else mine
            target, info["how"] = set(share), "cpus of the GPU's node, split among %d rank(s) on it" % len(peers)
            info["mempolicy_errno"] = _set_mempolicy_preferred(node)
    if target is None and local_world > 1:
        mine = sorted(allowed)
        share = mine[local_rank::local_world] if len(mine) >= local_world else mine
        target, info["how"] = set(share), "GPU's node not reachable from the allowed CPUs: even split of the allowed set"
    if target:
        try:
            os.sched_setaffinity(0, target)
            info["bound_cpus"] = len(target)
        except OSError as e:
            info["error"] = str(e)
    else:
        info["how"] = "single rank, GPU node unknown: left unbound"
    return info
