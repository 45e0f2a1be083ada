"""Run under torchrun on >= 2 GPUs (not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_gpu_check.py
Checks on real GPUs + NCCL: (1) per-instance tables: each rank's shard equals the oracle window with the same global
offset; (2) shared learner: replicas are identical after a merge and equal the schedule restated with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import multiagent_rlrm_b200 as P
    import oracle as O
    from multiagent_rlrm_b200.dist import ShardedTrainer, shard_range

    # (1) sharded per-instance tables
    sc = P.scenario_config5(shared=False)
    tr = ShardedTrainer(sc, 4096, device=f"cuda:{local}")
    tr.reset(); tr.train(300)
    o = O.Oracle(P.compile_scenario(sc, instance_offset=tr.offset + 7), 16, "f32")
    o.reset(); o.train(0, 300)
    q = tr.engine.q.cpu().numpy().reshape(tr.n_local, -1)[7:23].reshape(o.q.shape)
    assert np.array_equal(q, o.q), "shard window differs from the oracle"
    total = tr.global_counters()

    # (2) shared learner, merge every 16 iterations over NCCL
    sc = P.scenario_config5(shared=True)
    n_total, K, T = 2048, 16, 64
    tr2 = ShardedTrainer(sc, n_total, sync_every=K, device=f"cuda:{local}")
    tr2.reset(); tr2.train(T)
    mine = tr2.engine.q.clone()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "replicas differ after a merge"
    if rank == 0:
        reps = [O.Oracle(P.compile_scenario(sc, instance_offset=shard_range(n_total, r, world)[0]), shard_range(n_total, r, world)[1], "f32")
                for r in range(world)]
        for r_ in reps:
            r_.reset()
        for t0 in range(0, T, K):
            for r_ in reps:
                r_.train(t0, K)
            acc = reps[0].q.copy()
            for r_ in reps[1:]:
                acc = acc + r_.q  # dist.merge_replicas sums in rank order on every rank
            mean = acc / np.float32(world)
            for r_ in reps:
                r_.q[...] = mean
        got = mine.cpu().numpy()
        assert np.array_equal(got, reps[0].q), "shared-learner schedule differs from the oracle restatement"
        print(f"dist_gpu_check ok: world={world} counters={total} syncs={tr2.syncs}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
