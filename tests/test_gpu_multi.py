"""Two-GPU checks (skipped on a single-GPU box): the fused peer-memory all-reduce of the trainer
against the NCCL all-reduce path on the same ray shards, rank agreement, and data-parallel ==
whole-batch gradients."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from loma_nerf_b200 import api, sharding
    from oracle import oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = api.Context(rank)
    ctx.set_stream(torch.cuda.current_stream())
    R, S, E = 512, 64, 5
    cases = [O.make_nerf_case(1200 + i, R, S) for i in range(4)]
    dims = [int(v) for v in cases[0]["dims"]]
    ws0, bs0 = cases[0]["ws"], cases[0]["bs"]

    def shard(c):
        b = sharding.shard_rays(dict(X=c["X"], dists=c["dists"], target=c["target"]), world, rank, S)
        return {k: torch.as_tensor(np.ascontiguousarray(v, np.float32)).cuda() for k, v in b.items()}

    # (1) NCCL path: grad -> all_reduce -> apply
    tr_n = api.Trainer(ctx, dims, ws0, bs0)
    for c in cases:
        sharding.data_parallel_step(tr_n, dict(shard(c), path="tc"))
    w_n, b_n, loss_n = tr_n.read()
    # (2) fused peer-memory all-reduce inside the step
    tr_p = api.Trainer(ctx, dims, ws0, bs0)
    tr_p.enable_peer_allreduce()
    for c in cases:
        tr_p.step(**dict(shard(c), path="tc"))
    w_p, b_p, loss_p = tr_p.read()
    status = tr_p.comm_status()
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), w_n=w_n, b_n=b_n, w_p=w_p, b_p=b_p, loss_n=loss_n, loss_p=loss_p,
             status=status)
    tr_n.close(); tr_p.close(); ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_matches_nccl_path_and_ranks_agree(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (dict(np.load(os.path.join(tmp_path, "r%d.npz" % r))) for r in range(2))
    assert int(r0["status"]) == 0 and int(r1["status"]) == 0
    for k in ("w_p", "b_p", "w_n", "b_n"):
        assert np.array_equal(r0[k], r1[k]), k                    # replicas never diverge
    # same per-rank vectors, summed by NCCL or in rank order by the fused kernel
    assert rel_err(r0["w_p"], r0["w_n"]) <= 1e-6 and rel_err(r0["b_p"], r0["b_n"]) <= 1e-6
    assert abs(float(r0["loss_p"]) - float(r0["loss_n"])) <= 1e-5 * abs(float(r0["loss_n"]))
