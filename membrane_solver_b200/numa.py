"""Keep a rank's host buffers next to its GPU.

With one process per GPU every rank moves its own share of the positions and gradients over PCIe.  On a
two-socket box a rank that runs (and first-touches its pinned buffers) on the other socket sends that traffic
through the socket interconnect, and eight ranks that all allocate on node 0 share one memory controller set:
round 1 measured 13.8 GB/s per rank at 8 GPUs against 51 GB/s alone.  ``bind_to_gpu`` restricts the calling
process to the CPUs of the GPU's NUMA node (sysfs ``numa_node`` of its PCI function) BEFORE the buffers are
allocated, so Linux's default first-touch policy places them on that node.  It never fails: on a box without
NUMA information (or a single node) it reports why nothing was bound.
"""

from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def _pci_address(device_index: int) -> str | None:
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        dom, bus, dev = (getattr(p, k, None) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id"))
        if None not in (dom, bus, dev):
            return f"{int(dom):04x}:{int(bus):02x}:{int(dev):02x}.0"
    except Exception:
        pass
    try:  # NVML enumerates in PCI order; honour CUDA_VISIBLE_DEVICES when it is a plain index list
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
        index = device_index
        if visible and all(tok.strip().isdigit() for tok in visible.split(",")):
            index = int(visible.split(",")[device_index])
        bus_id = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
        return bus_id.lower()[-12:]   # NVML prints an 8-digit domain, sysfs a 4-digit one
    except Exception:
        return None


def bind_to_gpu(device_index: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node of CUDA device ``device_index``."""
    info: dict = {"bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
    except OSError:
        nodes = []
    info["nodes"] = len(nodes)
    if len(nodes) < 2:
        info["reason"] = "single NUMA node"
        return info
    addr = _pci_address(device_index)
    if addr is None:
        info["reason"] = "PCI address of the device unknown"
        return info
    try:
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            info["reason"] = "sysfs reports no NUMA node for the device"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            info["reason"] = "no allowed CPU on the device's node"
            return info
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, node=node, cpus=len(allowed), pci=addr)
    except (OSError, ValueError) as exc:
        info["reason"] = f"{type(exc).__name__}: {exc}"
    return info
