"""N > 1 host logic on CPU: world_size-2 gloo process group -- shard arithmetic, robot-table broadcast,
result gathering.  (The solve itself needs a GPU; here every rank checks what it WOULD solve.)"""
import os
import socket

import numpy as np
import pytest

from dex_retargeting_b200.parallel import shard_range

torch = pytest.importorskip("torch")


def test_shard_range_partitions():
    for total in (0, 1, 7, 8, 65536, 65537, 100003):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
                assert e0 == b1
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "tests"))
    import torch.distributed as dist

    from helpers import build_product
    from dex_retargeting_b200.parallel import all_gather_qpos, broadcast_table, shard_range
    from dex_retargeting_b200.table import table_bytes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 builds the "authoritative" table; rank 1 starts from a different config on purpose
        key = "teleop/allegro_hand_right" if rank == 0 else "teleop/leap_hand_right"
        seq = build_product(key)
        data = broadcast_table(seq.optimizer, src=0)
        ref = table_bytes(build_product("teleop/allegro_hand_right").optimizer.build_table())
        assert data == ref, "every rank must hold rank 0's table after the broadcast"
        # shards are contiguous, disjoint and cover the batch; gathering restores the order
        total = 1001
        b, e = shard_range(total, rank, world)
        mine = torch.arange(b, e, dtype=torch.float32)[:, None].repeat(1, 3)
        full = all_gather_qpos(mine, total)
        assert full.shape == (total, 3)
        assert torch.equal(full[:, 0], torch.arange(total, dtype=torch.float32))
        (Path(tmpdir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_broadcast_and_shards(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
