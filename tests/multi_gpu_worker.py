"""Worker for tests/test_gpu_multi.py: every rank runs LM iterations on its factor shard (one in-graph
ncclAllReduce per iteration + the peer-memory scalar exchange, or one of the two fallbacks); rank 0 compares
records and final state with the single-process oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hyperslam_b200 import runtime, synthetic  # noqa: E402


def stage(msg):
    if os.environ.get("HB200_DEBUG"):
        print(f"[rank {os.environ.get('RANK')}] {msg}", file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    win = synthetic.make_window(order=4, num_knots=20, num_landmarks=160, num_imu=400, seed=synthetic.SEED_BASE + 700, constant_knots=2)
    stage("process group up")
    mode = sys.argv[1] if len(sys.argv) > 1 else "peer"
    if mode != "peer":   # "peer": barrier + system reduction + scalar exchange all over peer memory, no NCCL call per iteration
        os.environ["HB200_PEER_REDUCE"] = "0"
    ctx = runtime.Context(local, use_graph=(mode != "callback"))
    ctx.load_window(win.shard(rank, world))
    info = {}
    if mode == "callback":   # legacy hook: the host program reduces (torch.distributed here), no graph
        ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
        cache = {}

        def allreduce(ptr, count, stream):
            if (ptr, count) not in cache:
                class _Arr:
                    __cuda_array_interface__ = dict(shape=(count,), typestr="<f8", data=(ptr, False), version=2)
                cache[(ptr, count)] = torch.as_tensor(_Arr(), device=torch.device("cuda", local))
            with torch.cuda.stream(ext):
                dist.all_reduce(cache[(ptr, count)])
            return 0

        ctx.set_allreduce(allreduce)
        beta = torch.tensor([ctx.bandwidth()], device="cuda")   # the packed layout must agree across ranks
        dist.all_reduce(beta, op=dist.ReduceOp.MAX)
        ctx.set_min_bandwidth(int(beta.item()))
    else:                    # the product path: ncclAllReduce enqueued by the library, inside the iteration's CUDA graph
        info = ctx.connect_torch_distributed(dist, peer_mailbox=(mode in ("peer", "nccl+mailbox")), log=stage)
    stage(f"connected {info}")
    recs = ctx.iterate(4)
    stage("iterated")
    state = ctx.state()
    ok, msg = True, ""
    if rank == 0:
        import oracle_lib as ol
        ow = ol.OracleWindow(win)
        for it, rec in enumerate(recs):
            o = ow.iterate(apply=True)
            if not (abs(rec["cost"] - o["cost"]) <= 1e-7 * o["cost"] and abs(rec["cost_new"] - o["cost_new"]) <= 1e-6 * abs(o["cost_new"])
                    and rec["accepted"] == o["accepted"] and rec["spd"] == 1):
                ok, msg = False, f"iteration {it}: {rec} vs oracle cost {o['cost']} -> {o['cost_new']} accepted {o['accepted']}"
        so = ow.state()
        if np.abs(state["knots"] - so["knots"]).max() > 1e-6:
            ok, msg = False, "knots differ"
    # every rank holds the same replicated knots; landmarks are updated by their owner only
    knots = torch.from_numpy(state["knots"]).cuda()
    ref = knots.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(knots, ref))
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        info = ctx.comm_info() if mode != "callback" else info
        print(json.dumps(dict(ok=bool(flag.item() == 1.0), msg=msg, replicas_identical=same, world=world, mode=mode, comm=info)))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
