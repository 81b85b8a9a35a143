"""Host-side helpers of the ingest path that need no GPU."""
import os

from a2c_ppo_acktr import utils


def test_bind_host_to_device_node_is_a_noop_without_a_device():
    before = os.sched_getaffinity(0)
    assert utils.bind_host_to_device_node(0) is None          # no CUDA device here: nothing is touched
    assert os.sched_getaffinity(0) == before
