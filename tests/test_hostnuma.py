"""CPU checks of the host-placement helper (corrla_rs_b200/hostnuma.py): parsing, and that it degrades to "nothing done"
where the topology is not visible (this container: no GPU, no PCI device entry)."""
import os

from corrla_rs_b200 import hostnuma


def test_cpulist_parsing():
    assert hostnuma._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostnuma._parse_cpulist("") == set()
    assert hostnuma._parse_cpulist("5") == {5}


def test_binding_is_a_no_op_without_topology(monkeypatch):
    before = os.sched_getaffinity(0)
    monkeypatch.setattr(hostnuma, "_pci_bus_id", lambda device: None)
    monkeypatch.setattr(hostnuma, "_topo_cpu_affinity", lambda device: None)
    assert hostnuma.gpu_numa_node(0) is None
    assert hostnuma.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


def test_binding_from_the_driver_topology_keeps_to_allowed_cpus(monkeypatch):
    allowed = os.sched_getaffinity(0)
    monkeypatch.setattr(hostnuma, "_pci_bus_id", lambda device: None)
    # the GPU's affinity covers every allowed CPU (the 8-GPU test box: 0-31 for every GPU): nothing to change
    monkeypatch.setattr(hostnuma, "_topo_cpu_affinity", lambda device: set(allowed) | {10_000})
    r = hostnuma.bind_to_gpu_numa_node(0)
    assert r is not None and r["changed"] is False and r["cpus"] == len(allowed)
    assert os.sched_getaffinity(0) == allowed
    # an affinity that shares no CPU with the allowed set must not be applied
    monkeypatch.setattr(hostnuma, "_topo_cpu_affinity", lambda device: {10_000, 10_001})
    r = hostnuma.bind_to_gpu_numa_node(0)
    assert r is not None and r["changed"] is False
    assert os.sched_getaffinity(0) == allowed
