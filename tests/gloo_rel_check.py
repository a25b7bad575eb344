"""world_size-2 gloo check of the relation-update algebra used by dglke_b200.dist (CPU only)."""
import os
import sys

import torch as th
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kge_oracle as ko  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_rel, D, B, lr = 11, 16, 40, 0.3
    g = th.Generator().manual_seed(5)
    rel0 = th.randn(n_rel, D, generator=g)
    st0 = th.rand(n_rel, generator=g) * 1e-2
    per_rank = []
    for r in range(world):
        gr = th.Generator().manual_seed(100 + r)
        per_rank.append((th.randint(0, n_rel, (B,), generator=gr), th.randn(B, D, generator=gr)))
    idx, grad = per_rank[rank]
    # what every rank does: dense per-relation sums -> all-reduce -> one Adagrad application
    buf = th.zeros(n_rel * D + n_rel)
    rg, rgs = buf[:n_rel * D].view(n_rel, D), buf[n_rel * D:]
    rg.index_add_(0, idx, grad)
    rgs.index_add_(0, idx, (grad * grad).mean(1))
    dist.all_reduce(buf)
    rel, st = rel0.clone(), st0.clone()
    touched = rgs > 0
    st[touched] += rgs[touched]
    rel[touched] += -lr * rg[touched] / (st[touched].sqrt() + 1e-10).unsqueeze(1)
    # reference semantics: one trace entry holding all ranks' per-edge rows
    ref_rel, ref_st = rel0.clone(), st0.clone()
    ko.adagrad_entry(ref_rel, ref_st, th.cat([p[0] for p in per_rank]), th.cat([p[1] for p in per_rank]), lr)
    assert th.allclose(rel, ref_rel, rtol=1e-5, atol=1e-6) and th.allclose(st, ref_st, rtol=1e-6, atol=1e-8)
    gathered = [th.zeros_like(rel) for _ in range(world)]
    dist.all_gather(gathered, rel)
    assert all(th.equal(gathered[0], x) for x in gathered), "replicas diverged"
    if rank == 0:
        print("GLOO_REL_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
