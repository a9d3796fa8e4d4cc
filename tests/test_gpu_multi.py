"""Two-GPU path (skipped on single-GPU boxes): NCCL broadcast of the robot table, adoption of the device copy,
contiguous shards solved independently, gathered result identical to the single-GPU result."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


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

    from helpers import build_oracle, build_product, synth_problems
    from dex_retargeting_b200.parallel import all_gather_qpos, broadcast_table, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        key = "teleop/shadow_hand_right"
        seq = build_product(key, device=rank)
        if rank != 0:  # perturb rank 1's local limits: after the broadcast it must solve with rank 0's table
            seq.optimizer.set_joint_limit(seq.joint_limits * 0.5)
        broadcast_table(seq.optimizer, src=0)
        o = build_oracle(key)
        rng = np.random.RandomState(5)
        refs, fixed, x0, _ = synth_problems(o, 1001, rng, init_noise=0.05, target_noise=0.01)
        b, e = shard_range(1001, rank, world)
        q = seq.optimizer.retarget_batch(torch.from_numpy(refs[b:e]).to(dev), None, torch.from_numpy(x0[b:e]).to(dev))
        full = all_gather_qpos(q, 1001)
        torch.cuda.synchronize()
        if rank == 0:
            single = build_product(key, device=0).optimizer.retarget_batch(torch.from_numpy(refs).to(dev), None,
                                                                            torch.from_numpy(x0).to(dev))
            torch.cuda.synchronize()
            assert torch.equal(full, single), "sharded result differs from the single-GPU result"
        (Path(tmpdir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_table_broadcast_and_sharded_solve(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
