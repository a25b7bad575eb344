"""Multi-GPU parity check, launched with torchrun on >= 2 GPUs (see tests/test_dist.py):
every rank trains its own batch through ShardedTrainer (entity rows sharded over the GPUs, remote
rows reached with peer loads / red.add over NVLink, relation gradients all-reduced with NCCL).
The ranks' batches touch disjoint entity and relation ids (rank r uses ids == r mod world), so the
result is order independent and must equal the CPU oracle applying the batches one after another."""
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dgl-ke_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import kge_oracle as ko
    from dglke_b200.engine import Hyper
    from dglke_b200.dist import ShardedTrainer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    th.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = th.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    model = sys.argv[1] if len(sys.argv) > 1 else "TransE_l2"
    n_ent, n_rel, hidden, B, N = 4001, 16, 64, 256, 64
    hp = Hyper(model=model, hidden_dim=hidden, gamma=12.0, lr=0.2, reg_coef=1e-6, adversarial=True)
    ohp = ko.Hyper(model=model, hidden_dim=hidden, gamma=12.0, lr=0.2, reg_coef=1e-6, adversarial=True)
    tr = ShardedTrainer(hp, n_ent, n_rel, dev, seed=1)
    full0 = tr.gather_entity_table().cpu()
    rel0 = tr.rel_emb.cpu().clone()
    steps = 3
    batches = {}
    for r in range(world):
        for s in range(steps):
            rng = np.random.default_rng(1000 * r + s)
            ids = np.arange(r, n_ent, world)
            rels = np.arange(r, n_rel, world)
            h, t, ng = rng.choice(ids, B), rng.choice(ids, B), rng.choice(ids, B)
            rr = rng.choice(rels, B)
            nodes, inv = np.unique(np.concatenate([h, t]), return_inverse=True)
            T = lambda a: th.from_numpy(np.ascontiguousarray(a.astype(np.int64)))
            batches[(r, s)] = [T(nodes), T(inv[:B]), T(inv[B:]), T(rr), T(ng)]
    for s in range(steps):
        b = [x.to(dev) for x in batches[(rank, s)]]
        tr.step(*b, N, N, bool(s % 2))
        tr.barrier()
    got_ent = tr.gather_entity_table().cpu()
    got_rel = tr.rel_emb.cpu()
    # all replicas of the relation table must be identical
    ref_rel = tr.rel_emb.clone()
    dist.broadcast(ref_rel, src=0)
    assert th.equal(ref_rel, tr.rel_emb), "relation replicas diverged"
    if rank == 0:
        ent, rel = full0.clone(), rel0.clone()
        es, rs = th.zeros(n_ent), th.zeros(n_rel)
        for s in range(steps):
            for r in range(world):
                ko.train_step(ohp, ent, es, rel, rs, *batches[(r, s)], B // N, N, N, bool(s % 2))
        np.testing.assert_allclose(got_ent.numpy(), ent.numpy(), rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(got_rel.numpy(), rel.numpy(), rtol=1e-4, atol=2e-6)
        moved = float((got_ent - full0).abs().max())
        assert moved > 1e-4, "tables did not move"
        print("DIST_CHECK_OK model=%s world=%d max|delta|=%.3e" % (model, world, moved), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
