"""Sampler: the numpy restatement's invariants on CPU; on the GPU the device sampler (kge_sampler_*)
against it bit for bit, and a training step fed straight from device-sampled indices against the oracle."""
import numpy as np
import pytest
import torch as th

import kge_oracle as ko


def _graph(n_ent=500, n_rel=7, n_edges=3000, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_ent, n_edges), rng.integers(0, n_rel, n_edges), rng.integers(0, n_ent, n_edges)


def test_host_sampler_invariants():
    from dglke_b200.sampler import HostSampler, feistel_perm, half_bits
    h, r, t = _graph()
    s = HostSampler(h, r, t, 500, 96, 32, seed=11)
    per_epoch = 3000 // 96
    # one epoch = every edge at most once, B * per_epoch distinct edges; the next epoch is a different order
    e0 = np.concatenate([s.edge_ids(k) for k in range(per_epoch)])
    e1 = np.concatenate([s.edge_ids(per_epoch + k) for k in range(per_epoch)])
    assert len(np.unique(e0)) == len(e0) == 96 * per_epoch and e0.min() >= 0 and e0.max() < 3000
    assert len(np.unique(e1)) == len(e1) and not np.array_equal(e0, e1)
    # the Feistel map is a bijection of [0, n) for awkward n
    for n in (1, 2, 5, 97, 1000, 4097):
        p = feistel_perm(np.arange(n, dtype=np.uint64), n, half_bits(n), 12345)
        assert sorted(p.tolist()) == list(range(n))
    for k in (0, 1, 7):
        b = s.sample(k)
        assert b["neg_head"] == bool(k & 1)                        # tail first (sampler.py:853-859)
        assert b["neg"].shape == (3 * 32,) and b["neg"].min() >= 0 and b["neg"].max() < 500
        assert np.array_equal(b["node_ids"][b["head_local"]], b["head"])
        assert np.array_equal(b["node_ids"][b["tail_local"]], b["tail"])
        assert len(np.unique(b["node_ids"])) == len(b["node_ids"])
        # order of first appearance in [heads | tails]
        keys = np.concatenate([b["head"], b["tail"]])
        seen, order = set(), []
        for x in keys.tolist():
            if x not in seen:
                seen.add(x)
                order.append(x)
        assert order == b["node_ids"].tolist()
    # negatives are (close to) uniform: chi-square-ish sanity over many draws
    s2 = HostSampler(h, r, t, 10, 1024, 1024, seed=3)
    cnt = np.bincount(np.concatenate([s2.sample(k)["neg"] for k in range(20)]), minlength=10)
    assert cnt.min() > 0.85 * cnt.mean() and cnt.max() < 1.15 * cnt.mean()
    with pytest.raises(ValueError):
        HostSampler(h, r, t, 500, 100, 32)                          # batch not a multiple of neg_sample_size


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(500, 7, 3000, 96, 32), (14951, 1345, 50000, 1000, 200), (50, 3, 200, 64, 64)])
def test_device_sampler_bit_exact(shape):
    from dglke_b200.sampler import HostSampler, DeviceSampler
    n_ent, n_rel, n_edges, B, Ns = shape
    h, r, t = _graph(n_ent, n_rel, n_edges, seed=5)
    hs = HostSampler(h, r, t, n_ent, B, Ns, seed=77)
    ds = DeviceSampler(h, r, t, n_ent, B, Ns, seed=77)
    per_epoch = n_edges // B
    for k in (0, 1, 2, per_epoch - 1, per_epoch, per_epoch + 1, 3 * per_epoch + 2):
        want = hs.sample(k)
        db = ds.sample(k)
        nodes, hl, tl, rel, neg = (x.cpu().numpy() for x in db.tensors())
        assert db.neg_head == want["neg_head"]
        assert np.array_equal(nodes, want["node_ids"]), "node list, step %d" % k
        assert np.array_equal(hl, want["head_local"]) and np.array_equal(tl, want["tail_local"])
        assert np.array_equal(rel, want["rel"]) and np.array_equal(neg, want["neg"])
    ds.close()


@pytest.mark.gpu
def test_training_step_from_device_sampled_batch():
    """kge_step_fused on indices that never left the GPU (node count known only on the device) vs the oracle fed the
    host mirror's indices."""
    from dglke_b200.sampler import HostSampler, DeviceSampler
    from test_gpu_parity import _engine
    n_ent, n_rel, n_edges, B, Ns = 3000, 20, 20000, 400, 200
    h, r, t = _graph(n_ent, n_rel, n_edges, seed=9)
    hp = ko.Hyper(model="TransE_l2", hidden_dim=400, gamma=19.9, lr=0.25, reg_coef=1e-7, adversarial=True)
    ent, es, rel, rs = ko.init_tables(hp, n_ent, n_rel, seed=1)
    eng, (e, e_s, rr, r_s) = _engine(hp, ent, es, rel, rs)
    hs = HostSampler(h, r, t, n_ent, B, Ns, seed=5)
    ds = DeviceSampler(h, r, t, n_ent, B, Ns, seed=5)
    o = [x.clone() for x in (ent, es, rel, rs)]
    T = lambda a: th.from_numpy(np.ascontiguousarray(a))
    for k in range(4):
        w = hs.sample(k)
        fb = ko.train_step(hp, o[0], o[1], o[2], o[3], T(w["node_ids"]), T(w["head_local"]), T(w["tail_local"]), T(w["rel"]),
                           T(w["neg"]), B // Ns, Ns, Ns, w["neg_head"])
        log4 = eng.step_sampled(ds.sample(k), Ns, Ns).cpu().numpy()
        np.testing.assert_allclose(log4[2], fb["log"]["loss"], rtol=5e-5)
        np.testing.assert_allclose(log4[3], fb["log"]["regularization"], rtol=5e-5)
    th.cuda.synchronize()
    # state_sum starts at zero: the first Adagrad step is a normalised step of size ~lr, so fp32-level gradient differences
    # show up at ~1e-4 of lr = 0.25
    np.testing.assert_allclose(e.cpu().numpy(), o[0].numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(rr.cpu().numpy(), o[2].numpy(), rtol=1e-4, atol=5e-5)
    ds.close()
