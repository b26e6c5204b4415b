"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous sharding covers every model exactly
once and the optional result gather reassembles shards in order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from generate3d_b200 import shard


def test_shard_range_partitions():
    for n in (0, 1, 7, 1024, 100000):
        for w in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = shard.init_process_group("gloo")
    lo, hi = shard.shard_range(n_total, r, w)
    local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1).repeat(1, 3)     # stand-in for per-model grids
    full = shard.gather_shards(local, n_total)
    tmax = shard.all_reduce_max(1.0 + r, device="cpu")
    tsum = shard.all_reduce_sum(hi - lo, device="cpu")
    q.put((r, full[:, 0].tolist(), tmax, tsum))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r, full, tmax, tsum in res:
        assert full == [float(i) for i in range(n_total)]
        assert tmax == 2.0 and tsum == n_total


def test_bind_to_gpu_numa_node_with_fake_sysfs(tmp_path, monkeypatch):
    """Topology lookup of the multi-GPU host path: PCI address -> numa_node -> cpulist -> sched_setaffinity;
    unknown topology must come back as None, never as an exception."""
    from generate3d_b200 import shard
    have = sorted(os.sched_getaffinity(0))
    monkeypatch.setattr(shard, "_gpu_pci_address", lambda i: "0000:1b:00.0")
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    (dev / "numa_node").write_text("-1\n")
    assert shard.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None            # single-node host
    (dev / "numa_node").write_text("1\n")
    (node / "cpulist").write_text("%d\n" % have[0])
    calls = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
    got = shard.bind_to_gpu_numa_node(0, sysfs=str(tmp_path))
    if len(have) > 1:
        assert got == {"numa_node": 1, "cpus": 1} and calls == [{have[0]}]
    else:
        assert got is None and not calls                                          # nothing to narrow down
    (node / "cpulist").write_text("100000-100003\n")                            # no usable CPU on that node
    assert shard.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    monkeypatch.setattr(shard, "_gpu_pci_address", lambda i: None)
    assert shard.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert shard._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
