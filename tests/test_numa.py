"""CPU: /sys parsing of the NUMA placement helper (no effect on results, only on where the e2e feed runs)."""
from motioncam_decoder_b200 import numa


def test_cpulist_parsing():
    assert numa._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._cpulist("") == set()
    assert numa._cpulist("5") == {5}


def test_unknown_device_is_not_bound():
    assert numa.gpu_numa_node("ffff:ff:1f.7") is None
    assert "not bound" in numa.bind_to_gpu_node("ffff:ff:1f.7")
