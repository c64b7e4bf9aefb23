"""Multi-rank AG-DQN learner over NCCL (BASELINE configs[3]); needs >= 2 GPUs (skipped on one).
Scenes and replay shard per rank; the flat gradient is all-reduced once per update, overlapped with the
next sub-step.  Ranks must hold bit-identical weights after every tick, with and without the overlap."""
import os
import random
import socket

import numpy as np
import pytest

from conftest import PKG_NAME

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, overlap, q):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = importlib.import_module(PKG_NAME)
    M = importlib.import_module(PKG_NAME + ".algorithm.MADQN_structure.MADQN")
    from oracle import oracle as orc
    grid = orc.README_9x9
    E, N = 64, 4
    tasks = np.array([orc.make_tasks(grid, random.Random(1000 * rank + e)) for e in range(E)], dtype=np.uint8)
    sc = pkg.BatchedScene(grid, E, N, tasks=tasks, auto_reset=True, device=rank)
    sc.reset()
    torch.manual_seed(rank)  # different initial weights per rank: the constructor must broadcast rank 0's
    ag = M.BatchedAgent(sc, memory_size=4096, batch_size=32, epsilon=0.8, ddp=True, seed=3)
    ag.overlap_allreduce = overlap
    sums = []
    for t in range(10):
        ag.tick(train=True)
        flat = torch.cat([p.detach().reshape(-1) for p in ag._params]).double()
        sums.append((float(flat.sum()), float((flat * flat).sum()), ag.learn_step_counter))
    q.put((rank, sums, ag.transitions_stored()))
    dist.barrier()
    dist.destroy_process_group()
    sc.close()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_learner_keeps_ranks_identical(overlap):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict()
    for _ in ps:
        r, sums, stored = q.get(timeout=600)
        res[r] = (sums, stored)
    for p in ps:
        p.join(timeout=120)
    assert res[0][0] == res[1][0], (res[0][0][-1], res[1][0][-1])   # weights + update counts identical on both ranks
    assert res[0][0][-1][2] > 20 and res[0][1] > 0 and res[1][1] > 0
    assert res[0][0][0][:2] != res[0][0][-1][:2]                     # and they did move
