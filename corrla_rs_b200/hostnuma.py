"""Host-side placement for the host<->device copies of the RSVD path (no compute here).

A B200 node has two CPU sockets; each GPU hangs off one of them.  A rank whose staging memory (pinned input, output
buffers, the bounce buffers of hostcopy.cu) lives on the other socket pays the inter-socket link on every copy, and with
eight ranks copying 4 GiB each at once that link is what saturates.  `bind_to_gpu_numa_node` pins the calling process to
the CPUs of the NUMA node its GPU is attached to BEFORE the buffers are allocated and first touched, so that first-touch
placement puts them next to the GPU.  Best effort: returns None (and changes nothing) when the topology is not visible
(containers without /sys, single-node machines)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path


def _pci_bus_id(device: int) -> str | None:
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        dom, bus, dev = getattr(p, "pci_domain_id", None), getattr(p, "pci_bus_id", None), getattr(p, "pci_device_id", None)
        if bus is not None and dev is not None:
            return f"{(dom or 0):04x}:{bus:02x}:{dev:02x}.0"
    except Exception:
        pass
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if device < len(ids) and ids[device].strip().isdigit():
                idx = int(ids[device])
        out = subprocess.run(["nvidia-smi", "-i", str(idx), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        if out:
            return out.lower()[-12:]          # nvidia-smi prints an 8-digit domain: keep 0000:xx:yy.z
    except Exception:
        pass
    return None


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


def gpu_numa_node(device: int) -> int | None:
    bdf = _pci_bus_id(device)
    if bdf is None:
        return None
    try:
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text().strip())
    except Exception:
        return None
    return node if node >= 0 else None


def _topo_cpu_affinity(device: int) -> set[int] | None:
    """CPU affinity of the GPU from `nvidia-smi topo -m` (the driver's own view; works where sysfs hides numa_node)."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        lines = [ln for ln in out.splitlines() if ln.strip()]
        header = next(ln for ln in lines if "CPU Affinity" in ln)
        cols = [c.strip() for c in header.split("\t")]
        idx = cols.index("CPU Affinity")
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if device < len(ids) and ids[device].strip().isdigit():
                phys = int(ids[device])
        row = next(ln for ln in lines if ln.startswith(f"GPU{phys}\t") or ln.startswith(f"GPU{phys} "))
        cells = [c.strip() for c in row.split("\t")]
        # the header has one leading empty cell for the row labels
        cell = cells[idx] if len(cells) > idx else ""
        cpus = _parse_cpulist(cell)
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device: int) -> dict | None:
    """Restrict this process to the CPUs of the NUMA node of GPU `device` (sched_setaffinity); memory touched afterwards
    is placed on that node by the kernel's first-touch policy.  Returns {"node", "cpus"} or None when nothing was done."""
    node = gpu_numa_node(device)
    if node is None:
        cpus = _topo_cpu_affinity(device)
        if not cpus:
            return None
        try:
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if not cpus or cpus == allowed:
                return {"node": None, "cpus": len(allowed), "changed": False, "source": "nvidia-smi topo"}
            os.sched_setaffinity(0, cpus)
            return {"node": None, "cpus": len(cpus), "changed": True, "source": "nvidia-smi topo"}
        except Exception:
            return None
    try:
        cpus = _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return {"node": node, "cpus": len(allowed), "changed": False}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "changed": True}
    except Exception:
        return None
