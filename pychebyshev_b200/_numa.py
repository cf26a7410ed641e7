"""Host-side placement for the host-buffer pipeline: which CPUs / memory node sit next to a GPU.

On a two-socket box half of the GPUs hang off each socket.  A pinned buffer whose pages live on
the other socket is DMA'd across the socket interconnect, and a driving thread on the other socket
pays the same hop for every doorbell.  ``bound_to_device`` binds the calling thread to the CPUs of
the GPU's NUMA node for the duration of an allocation (Linux allocates -- and the driver pins --
pages on the node of the touching thread); ``bind_thread`` does it for good.  Everything here is a
no-op when the topology is not exposed (``numa_node`` = -1, single node, containers without sysfs).
"""

from __future__ import annotations

import contextlib
import os

_cache: dict = {}


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_bus_id(device) -> str | None:
    """``domain:bus:device.function`` of a CUDA ordinal (lower case, sysfs form)."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:  # noqa: BLE001
        return None


def gpu_numa_node(device) -> int:
    """NUMA node of the GPU's PCIe root, or -1 when unknown."""
    key = ("node", device)
    if key not in _cache:
        node = -1
        bdf = pci_bus_id(device)
        if bdf:
            txt = _read(f"/sys/bus/pci/devices/{bdf}/numa_node")
            if txt is not None:
                try:
                    node = int(txt)
                except ValueError:
                    node = -1
        _cache[key] = node
    return _cache[key]


def node_cpus(node: int) -> set:
    if node < 0:
        return set()
    return _parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist"))


def n_nodes() -> int:
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")
                    and d[4:].isdigit()])
    except OSError:
        return 1


def device_cpus(device) -> set:
    """CPUs next to the GPU that this process is allowed to run on (empty = unknown)."""
    if device is None or not hasattr(os, "sched_getaffinity"):
        return set()
    cpus = node_cpus(gpu_numa_node(device))
    if not cpus:
        return set()
    return cpus & os.sched_getaffinity(0)


@contextlib.contextmanager
def bound_to_device(device):
    """Bind the calling thread to the GPU's node while the block runs (no-op if unknown)."""
    cpus = device_cpus(device)
    if not cpus or os.environ.get("PCB_NUMA", "1") == "0":
        yield False
        return
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
        yield True
    finally:
        os.sched_setaffinity(0, old)


def bind_thread(device) -> bool:
    """Bind the calling thread to the GPU's node for good; returns whether anything changed."""
    cpus = device_cpus(device)
    if not cpus or os.environ.get("PCB_NUMA", "1") == "0":
        return False
    os.sched_setaffinity(0, cpus)
    return True


def describe(device=None) -> dict:
    out = {"nodes": n_nodes()}
    if device is not None:
        node = gpu_numa_node(device)
        out.update(gpu_node=node, gpu_pci=pci_bus_id(device),
                   node_cpus=len(node_cpus(node)), usable=len(device_cpus(device)),
                   enabled=os.environ.get("PCB_NUMA", "1") != "0")
    return out
