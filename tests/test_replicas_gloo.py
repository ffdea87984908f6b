"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding + timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xlstm_yolo_clean_b200 import replicas as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, size = R.batch_shard(129, rank, world)
    covered = torch.zeros(129)
    covered[start:start + size] = 1
    dist.all_reduce(covered)
    ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    tput = R.job_throughput(units_this_rank=float(size), ms_this_rank=ms)
    out.put((rank, start, size, bool((covered == 1).all()), R.max_over_ranks(ms), tput, R.env_rank()))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_shard_and_reductions_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, n0, ok0, mx0, t0, e0), (r1, s1, n1, ok1, mx1, t1, e1) = res
    assert (s0, n0, s1, n1) == (0, 65, 65, 64) and ok0 and ok1  # every sample exactly once
    assert mx0 == mx1 == 15.0                                    # slowest rank defines the time
    assert abs(t0 - 129 / 15e-3) < 1e-6 and t0 == t1             # whole-job throughput
    assert e0 == (0, 0, 2) and e1 == (1, 1, 2)


def test_single_process_is_identity():
    assert R.batch_shard(10, 0, 1) == (0, 10)
    assert R.max_over_ranks(3.5) == 3.5
    assert R.job_throughput(8, 2.0) == 4000.0
