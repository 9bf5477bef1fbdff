"""world_size-2 `gloo` test of the multi-rank path on CPU: each rank builds the packed reduced system
[S | b | diagH | g | cost] of ITS factor shard (oracle standing in for the kernels), one all-reduce
sums it, every rank finalises and must hold exactly the single-rank system -- the same buffer layout,
partition rule (landmark owner) and single collective the GPU path uses (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, widened):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from hyperslam_b200 import synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    win = synthetic.make_window(order=4, num_knots=16, num_landmarks=64, num_imu=96, seed=33, constant_knots=2)
    if widened:   # + bearing factors (sharded by landmark owner like the pixel factors) and pose factors (by index)
        win = synthetic.add_bearing_and_pose_factors(win, num_bearing=80, num_pose=17, seed=35)
    shard = win.shard(rank, world)
    packed = torch.from_numpy(ol.OracleWindow(shard).build_packed())
    dist.all_reduce(packed)                                  # the one system collective per iteration
    full = ol.OracleWindow(win)
    S, b = full.finalize_packed(packed.numpy())
    ref = full.iterate(apply=False)
    err_S = float(np.abs(S - ref["S"]).max() / np.abs(ref["S"]).max())
    err_b = float(np.abs(b - ref["b"]).max() / np.abs(ref["b"]).max())
    err_c = abs(float(packed[-2]) - ref["cost"]) / ref["cost"]
    # step acceptance needs the global trial cost: second (4-double) all-reduce
    scal = torch.tensor([float(shard.num_factors), 0.0, 0.0, 0.0], dtype=torch.float64)
    dist.all_reduce(scal)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([err_S, err_b, err_c, float(scal[0]), float(win.num_factors)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("widened", [False, True])
def test_two_rank_allreduce_reproduces_single_rank_system(tmp_path, widened):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), widened), nprocs=world, join=True)
    for r in range(world):
        err_S, err_b, err_c, nf, nf_ref = np.load(tmp_path / f"rank{r}.npy")
        assert err_S < 1e-12 and err_b < 1e-10 and err_c < 1e-12
        assert nf == nf_ref
