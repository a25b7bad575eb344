"""N>1 path: (gpu, needs >= 2 devices) the sharded trainer against the oracle via torchrun;
(cpu, gloo world_size 2) the host-side partition logic and the relation all-reduce equivalence."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_calls = [0]


def _torchrun(nproc, script, *args, timeout=600, env=None):
    _calls[0] += 1      # a fresh rendezvous port per launch: the previous launch's store may still hold its port
    port = 29500 + (os.getpid() * 7 + _calls[0] * 131) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["disjoint", "overlap", "pipelined"])
@pytest.mark.parametrize("model", ["TransE_l2", "ComplEx"])
def test_sharded_trainer_matches_oracle(model, mode):
    """2 ranks: on 2 GPUs over NCCL/NVLink when the box has them, else both ranks on cuda:0 (gloo for the relation
    all-reduce): the sharded TableView, the IPC mapping, the system-scope atomics and the fused multi-GPU schedule run
    either way.  `overlap` = every rank updates the SAME rows (cross-GPU Hogwild)."""
    env = {} if th.cuda.device_count() >= 2 else {"DIST_SAME_GPU": "1"}
    out = _torchrun(2, os.path.join(ROOT, "tests", "dist_check.py"), model, mode, env=env)
    assert "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_partition_math():
    from dglke_b200.dist import shard_rows, owner_of
    for n, w in [(10, 3), (14951, 8), (86054151, 8), (7, 8)]:
        per = (n + w - 1) // w
        covered = 0
        for r in range(w):
            p, lo, hi = shard_rows(n, w, r)
            assert p == per and lo == min(n, r * per) and hi == min(n, (r + 1) * per)
            covered += hi - lo
        assert covered == n
        ids = th.tensor([0, n // 2, n - 1])
        own = owner_of(ids, n, w)
        for i, o in zip(ids.tolist(), own.tolist()):
            _, lo, hi = shard_rows(n, w, o)
            assert lo <= i < hi


def test_relation_allreduce_equivalence_gloo():
    """world_size-2 gloo: summing per-relation gradient sums and mean(g^2) sums across ranks and applying
    Adagrad once per replica equals ExternalEmbedding.update over the concatenated per-edge rows."""
    out = _torchrun(2, os.path.join(ROOT, "tests", "gloo_rel_check.py"), timeout=300)
    assert out.stdout.count("GLOO_REL_OK") == 1, out.stdout[-2000:] + out.stderr[-3000:]


def test_descriptor_exchange_gloo():
    """world_size-3 gloo: dist.exchange_fds hands every rank's file descriptors to every peer (what carries the VMM shard
    allocations between the GPU processes)."""
    out = _torchrun(3, os.path.join(ROOT, "tests", "gloo_fd_check.py"), timeout=300)
    assert out.stdout.count("GLOO_FD_OK") == 1, out.stdout[-2000:] + out.stderr[-3000:]


def test_head_owner_edge_partition_covers_every_edge_once():
    from dglke_b200.dist import partition_edges_by_head_owner, shard_rows
    rng = np.random.default_rng(0)
    n_ent, world = 1003, 4
    heads = rng.integers(0, n_ent, 5000)
    seen = np.zeros(5000, dtype=int)
    for r in range(world):
        idx = partition_edges_by_head_owner(heads, n_ent, world, r)
        _, lo, hi = shard_rows(n_ent, world, r)
        assert ((heads[idx] >= lo) & (heads[idx] < hi)).all()
        seen[idx] += 1
    assert (seen == 1).all()
