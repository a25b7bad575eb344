"""Multi-GPU parity check, launched with torchrun (see tests/test_dist.py):

    torchrun --nproc-per-node 2 tests/dist_check.py <model> [disjoint|overlap]

Every rank trains its own batch through ShardedTrainer: entity rows sharded over the ranks (CUDA IPC mapped), remote
rows reached with peer loads / system-scope red.add from inside the kernels, relation gradients all-reduced.

  disjoint  the ranks' batches touch disjoint entity and relation ids (rank r uses ids == r mod world), so the result is
            order independent and must equal the CPU oracle applying the batches one after another.
  overlap   all ranks draw from the SAME small id range: the cross-GPU Hogwild case (several GPUs red.add into the same
            rows and state scalars).  The expected tables come from a "synchronous" oracle -- every rank's gradients from
            the common snapshot, then all Adagrad entries applied -- and the comparison is tolerance based: with a large
            initial state_sum the order in which the ranks' state adds land changes an update by < 1e-3 of its size,
            while a lost or doubled atomic changes it by O(1).

  pipelined the disjoint setting with kge_set_next_batch (--async_update): the fused kernels of step s copy the rows of
            step s+1, which therefore reads the entity table as it was BEFORE the update of step s (one step stale; the
            relation table and all updates are current).  The oracle is driven with exactly that lag.

DIST_SAME_GPU=1 puts every rank on cuda:0 with the gloo backend (NCCL refuses two ranks on one device): the IPC mapping,
the sharded TableView and the system-scope atomics are then exercised on a single-GPU box."""
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
    same_gpu = os.environ.get("DIST_SAME_GPU") == "1"
    local = 0 if same_gpu else int(os.environ["LOCAL_RANK"])
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    model = sys.argv[1] if len(sys.argv) > 1 else "TransE_l2"
    mode = sys.argv[2] if len(sys.argv) > 2 else "disjoint"
    n_ent, n_rel, hidden, B, N = 4001, 16, 64, 256, 64
    hp = Hyper(model=model, hidden_dim=hidden, gamma=12.0, lr=0.2, reg_coef=1e-6, adversarial=True)
    ohp = ko.Hyper(model=model, hidden_dim=hidden, gamma=12.0, lr=0.2, reg_coef=1e-6, adversarial=True)
    tr = ShardedTrainer(hp, n_ent, n_rel, dev, seed=1)
    state0 = 0.0
    if mode == "overlap":
        state0 = 0.5
        tr.ent_state_local.fill_(state0)
        tr.rel_state.fill_(state0)
        tr.barrier()
    full0 = tr.gather_entity_table().cpu()
    rel0 = tr.rel_emb.cpu().clone()
    steps = {"disjoint": 3, "pipelined": 5}.get(mode, 1)
    batches = {}
    for r in range(world):
        for s in range(steps):
            rng = np.random.default_rng(1000 * r + s)
            if mode in ("disjoint", "pipelined"):
                ids, rels = np.arange(r, n_ent, world), np.arange(r, n_rel, world)
                if mode == "pipelined":
                    ids = ids[:400]                 # consecutive steps of a rank share most of their rows: the lag matters
            else:                                   # every rank hammers the same 300 entities / 4 relations
                ids, rels = np.arange(0, 300), np.arange(0, 4)
            h, t, ng = rng.choice(ids, B), rng.choice(ids, B), rng.choice(ids, B)
            rr = rng.choice(rels, B)
            nodes, inv = np.unique(np.concatenate([h, t]), return_inverse=True)
            T = lambda a: th.from_numpy(np.ascontiguousarray(a.astype(np.int64)))
            batches[(r, s)] = [T(nodes), T(inv[:B]), T(inv[B:]), T(rr), T(ng)]
    mine = [[x.to(dev) for x in batches[(rank, s)]] for s in range(steps)]
    launches = []
    # pipelined: steps 1 .. steps-2 are announced by their predecessor; the first and the last step gather their own rows
    # (the last one is the launch-count reference: one kernel more than a step fed from staged rows)
    announced = [mode == "pipelined" and 0 < s < steps - 1 for s in range(steps)]      # step s reads staged rows
    for s in range(steps):
        nxt = (mine[s + 1][0], mine[s + 1][4]) if (s + 1 < steps and announced[s + 1]) else None
        n0 = tr.h.lib.kge_launch_count(tr.h.raw)
        tr.step(*mine[s], N, N, bool(s % 2), sync_between=(mode == "overlap"), next_batch=nxt)
        launches.append(tr.h.lib.kge_launch_count(tr.h.raw) - n0)
        tr.barrier()
    if mode == "pipelined":      # steps fed by the previous step's prefetch run without their own node gather
        assert all(launches[s] == launches[-1] - 1 for s in range(steps) if announced[s]), \
            "staged rows were not used: launches per step %r" % (launches,)
    got_ent = tr.gather_entity_table().cpu()
    got_rel = tr.rel_emb.cpu()
    # all replicas of the relation table must be identical
    ref_rel = tr.rel_emb.clone()
    dist.broadcast(ref_rel, src=0)
    assert th.equal(ref_rel, tr.rel_emb), "relation replicas diverged"
    if rank == 0:
        ent, rel = full0.clone(), rel0.clone()
        es, rs = th.full((n_ent,), state0), th.full((n_rel,), state0)
        if mode == "disjoint":
            for s in range(steps):
                for r in range(world):
                    ko.train_step(ohp, ent, es, rel, rs, *batches[(r, s)], B // N, N, N, bool(s % 2))
            np.testing.assert_allclose(got_ent.numpy(), ent.numpy(), rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(got_rel.numpy(), rel.numpy(), rtol=1e-4, atol=2e-6)
        elif mode == "pipelined":
            # ids are disjoint across ranks, so the ranks can be replayed one after the other; within a rank step s reads
            # the entity rows of the table BEFORE update s-1 (what step s-1's kernels copied), relations are current
            for r in range(world):
                snaps = [ent.clone()]
                for s in range(steps):
                    nodes, _, _, rr, ng = batches[(r, s)]
                    fb = ko.forward_backward(ohp, snaps[s - 1] if announced[s] else snaps[s], rel, *batches[(r, s)], B // N, N, N, bool(s % 2))
                    with th.no_grad():
                        ko.adagrad_entry(ent, es, nodes, fb["nodes_grad"], ohp.lr)
                        ko.adagrad_entry(ent, es, ng, fb["negs_grad"], ohp.lr)
                        ko.adagrad_entry(rel, rs, rr, fb["rels_grad"], ohp.lr)
                    snaps.append(ent.clone())
            np.testing.assert_allclose(got_ent.numpy(), ent.numpy(), rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(got_rel.numpy(), rel.numpy(), rtol=1e-4, atol=2e-6)
            # and the lag is real: the synchronous replay differs
            ent2, rel2 = full0.clone(), rel0.clone()
            es2, rs2 = th.full((n_ent,), state0), th.full((n_rel,), state0)
            for s in range(steps):
                for r in range(world):
                    ko.train_step(ohp, ent2, es2, rel2, rs2, *batches[(r, s)], B // N, N, N, bool(s % 2))
            assert float((ent2 - ent).abs().max()) > 1e-4, "test is blind: stale and synchronous replays coincide"
        else:
            # synchronous oracle: all gradients from the snapshot, then every Adagrad entry
            fbs = [ko.forward_backward(ohp, full0, rel0, *batches[(r, 0)], B // N, N, N, False) for r in range(world)]
            with th.no_grad():
                for r, fb in enumerate(fbs):
                    nodes, _, _, rr, ng = batches[(r, 0)]
                    ko.adagrad_entry(ent, es, nodes, fb["nodes_grad"], ohp.lr)
                    ko.adagrad_entry(ent, es, ng, fb["negs_grad"], ohp.lr)
                # relations: the ranks' per-edge rows are summed per relation and applied once (dist.py)
                idx = th.cat([batches[(r, 0)][3] for r in range(world)])
                ko.adagrad_entry(rel, rs, idx, th.cat([fb["rels_grad"] for fb in fbs]), ohp.lr)
            upd = float((ent - full0).abs().max())
            err = float((got_ent - ent).abs().max())
            err_rel = float((got_rel - rel).abs().max())
            upd_rel = float((rel - rel0).abs().max())
            print("overlap: max|update| %.3e, max|got - sync oracle| %.3e (entities); %.3e / %.3e (relations)"
                  % (upd, err, upd_rel, err_rel), flush=True)
            assert err <= 2e-3 * upd and err_rel <= 2e-3 * upd_rel, "cross-GPU atomics lost or doubled updates"
        moved = float((got_ent - full0).abs().max())
        assert moved > (1e-5 if mode == "overlap" else 1e-4), "tables did not move"
        print("DIST_CHECK_OK model=%s mode=%s world=%d max|delta|=%.3e" % (model, mode, world, moved), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
