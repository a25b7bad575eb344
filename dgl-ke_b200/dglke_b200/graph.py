"""Duck-typed positive / negative batch graphs and a seeded synthetic sampler.

The reference's hot path touches only a handful of members of DGL's sampled subgraphs
(models/general_models.py:376-427,548-568).  PosGraph / NegGraph expose exactly
those, so code written against the reference (`model.forward(pos_g, neg_g, gpu_id)`) runs unchanged.
DGL's C++ EdgeSampler itself is outside this round's scope (section 8f-2): SyntheticSampler draws
uniform ids with numpy instead and reproduces the reference's chunk bookkeeping
(dataloader/sampler.py:459-512) and tail/head alternation (:853-859).
"""
import numpy as np
import torch


class PosGraph:
    """ndata['id'] unique node ids, edata['id'] relation ids, all_edges() -> local (head, tail)."""

    def __init__(self, node_ids, head_local, tail_local, rel_ids, impts=None):
        self.ndata = {"id": node_ids}
        self.edata = {"id": rel_ids}
        if impts is not None:
            self.edata["impts"] = impts
        self._head, self._tail = head_local, tail_local

    def all_edges(self, order="eid"):
        return self._head, self._tail

    def number_of_edges(self):
        return int(self._head.shape[0])

    def number_of_nodes(self):
        return int(self.ndata["id"].shape[0])

    def apply_edges(self, fn):
        class _E:
            pass
        e = _E()
        e.src = {"emb": self.ndata["emb"][self._head]}
        e.dst = {"emb": self.ndata["emb"][self._tail]}
        e.data = self.edata
        self.edata.update(fn(e))

    def to(self, device, non_blocking=True):
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        g = PosGraph(mv(self.ndata["id"]), mv(self._head), mv(self._tail), mv(self.edata["id"]),
                     mv(self.edata["impts"]) if "impts" in self.edata else None)
        return g


class NegGraph:
    """ndata['id'][head_nid | tail_nid] = the C*Ns corrupting entity ids, chunk-major."""

    def __init__(self, neg_ids, num_chunks, chunk_size, neg_sample_size, neg_head):
        self.ndata = {"id": neg_ids}
        loc = torch.arange(neg_ids.shape[0], device=neg_ids.device)
        self.head_nid = loc
        self.tail_nid = loc
        self.num_chunks, self.chunk_size = num_chunks, chunk_size
        self.neg_sample_size, self.neg_head = neg_sample_size, neg_head
        self.edata = {}

    def neg_ids(self):
        return self.ndata["id"][self.head_nid if self.neg_head else self.tail_nid]

    def to(self, device, non_blocking=True):
        return NegGraph(self.ndata["id"].to(device, non_blocking=non_blocking), self.num_chunks, self.chunk_size,
                        self.neg_sample_size, self.neg_head)


def build_pos_graph(head, rel, tail, impts=None):
    """Global (head, rel, tail) id arrays -> PosGraph with unique node ids and local endpoints
    (what DGL's edge-subgraph construction hands to KEModel).  numpy in, CPU tensors out."""
    head, tail = np.asarray(head), np.asarray(tail)
    nodes, inv = np.unique(np.concatenate([head, tail]), return_inverse=True)
    B = head.shape[0]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int64)))
    return PosGraph(T(nodes), T(inv[:B]), T(inv[B:]), T(np.asarray(rel)),
                    None if impts is None else torch.from_numpy(np.asarray(impts, dtype=np.float32)))


def chunk_layout(batch_size, neg_sample_size, neg_chunk_size=None):
    """(num_chunks, chunk_size) the reference derives for a training batch
    (train.py:119-121 passes neg_chunk_size = neg_sample_size; sampler.py:496-512)."""
    cs = neg_sample_size if neg_chunk_size is None else neg_chunk_size
    if batch_size < cs:
        return 1, batch_size
    if batch_size % cs:
        return None          # ragged last batch: skipped by the reference (sampler.py:503-504)
    return batch_size // cs, cs


class SyntheticSampler:
    """Bidirectional one-shot iterator over seeded uniform batches: step 1,3,5.. corrupt tails,
    2,4,6.. corrupt heads (NewBidirectionalOneShotIterator, sampler.py:853-859)."""

    def __init__(self, n_entities, n_relations, batch_size, neg_sample_size, seed=0, rank=0, head_range=None):
        """head_range = (lo, hi): heads are drawn from [lo, hi) only -- the edges of a rank whose edge partition is
        'head owned by this rank' (dist.partition_edges_by_head_owner)"""
        self.head_range = head_range
        lay = chunk_layout(batch_size, neg_sample_size)
        if lay is None:
            raise ValueError("batch_size must be a multiple of neg_sample_size (utils.get_compatible_batch_size)")
        self.num_chunks, self.chunk_size = lay
        self.n_ent, self.n_rel, self.B, self.Ns = n_entities, n_relations, batch_size, neg_sample_size
        self.seed = seed + 100003 * rank
        self.step = 0

    def batch(self, k):
        rng = np.random.default_rng(self.seed + k)
        h, t = rng.integers(0, self.n_ent, self.B), rng.integers(0, self.n_ent, self.B)
        if self.head_range is not None:
            h = self.head_range[0] + h % max(1, self.head_range[1] - self.head_range[0])
        r = rng.integers(0, self.n_rel, self.B)
        ng = rng.integers(0, self.n_ent, self.num_chunks * self.Ns)
        pos_g = build_pos_graph(h, r, t)
        neg_g = NegGraph(torch.from_numpy(ng.astype(np.int64)), self.num_chunks, self.chunk_size, self.Ns,
                         neg_head=bool(k % 2))
        return pos_g, neg_g

    def __iter__(self):
        return self

    def __next__(self):
        k = self.step
        self.step += 1
        return self.batch(k)


class TripleSampler(SyntheticSampler):
    """Same iterator over a given triple list: positives are shuffled edges (one permutation per epoch),
    negatives uniform entity ids with replacement (DGL EdgeSampler semantics as used by
    TrainDataset.create_sampler, dataloader/sampler.py:376-419; exclude_positive=False)."""

    def __init__(self, heads, rels, tails, n_entities, n_relations, batch_size, neg_sample_size, seed=0, rank=0,
                 world=1, impts=None):
        """impts: per-edge importance weights (--has_edge_importance: the 4th column of the triple files), carried into
        pos_g.edata['impts'] (sampler.py:360-364)"""
        super().__init__(n_entities, n_relations, batch_size, neg_sample_size, seed, rank)
        idx = np.arange(len(heads))
        if world > 1:   # RandomPartition (sampler.py:256-290): a fixed random split of the edges over the ranks
            idx = np.random.default_rng(seed).permutation(len(heads))[rank::world]
        self.h, self.r, self.t = np.asarray(heads)[idx], np.asarray(rels)[idx], np.asarray(tails)[idx]
        self.w = None if impts is None else np.asarray(impts, dtype=np.float32)[idx]
        self.n_edges = len(self.h)
        if self.n_edges < batch_size:
            raise ValueError("fewer edges (%d) than batch_size (%d)" % (self.n_edges, batch_size))
        self._perm, self._epoch = None, -1

    def batch(self, k):
        per_epoch = self.n_edges // self.B          # the ragged last batch of an epoch is skipped
        epoch, j = divmod(k, per_epoch)
        if epoch != self._epoch:
            self._perm = np.random.default_rng(self.seed + 7919 * epoch).permutation(self.n_edges)
            self._epoch = epoch
        e = self._perm[j * self.B:(j + 1) * self.B]
        rng = np.random.default_rng(self.seed + 15485863 + k)
        ng = rng.integers(0, self.n_ent, self.num_chunks * self.Ns)
        pos_g = build_pos_graph(self.h[e], self.r[e], self.t[e], None if self.w is None else self.w[e])
        neg_g = NegGraph(torch.from_numpy(ng.astype(np.int64)), self.num_chunks, self.chunk_size, self.Ns,
                         neg_head=bool(k % 2))
        return pos_g, neg_g


class TripleFilter:
    """Which corruptions of a positive are TRUE triples of the graph (the `filter_false_neg` of DGL's EdgeSampler that
    the reference's EvalSampler turns on, dataloader/sampler.py:514-597): those candidates get bias -1 and
    KEModel.forward_test leaves them out of the ranking (general_models.py:463-468) -- the positive's own copy among the
    candidates included, which is what makes the result the usual filtered MRR.

    Two sorted key arrays ((head, rel) -> tails, (tail, rel) -> heads) built once from all known triples; a batch is two
    binary searches and one scatter."""

    def __init__(self, heads, rels, tails, n_relations):
        h, r, t = (np.asarray(a, dtype=np.int64) for a in (heads, rels, tails))
        self.n_rel = int(n_relations)
        kt = h * self.n_rel + r
        o = np.argsort(kt, kind="stable")
        self.key_tail, self.val_tail = kt[o], t[o]
        kh = t * self.n_rel + r
        o = np.argsort(kh, kind="stable")
        self.key_head, self.val_head = kh[o], h[o]

    def bias(self, heads, rels, tails, n_candidates, neg_head, candidates=None):
        """float32 [B, n_candidates]: -1 where replacing the head (neg_head) / tail by that candidate gives a known triple,
        else 0.  `candidates`: sorted entity ids of the columns (default: column j = entity j)."""
        h, r, t = (np.asarray(a, dtype=np.int64) for a in (heads, rels, tails))
        keys, K, V = ((t * self.n_rel + r, self.key_head, self.val_head) if neg_head else
                      (h * self.n_rel + r, self.key_tail, self.val_tail))
        lo, hi = np.searchsorted(K, keys, "left"), np.searchsorted(K, keys, "right")
        cnt = hi - lo
        rows = np.repeat(np.arange(len(keys)), cnt)
        pos = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(lo, cnt)
        ents = V[pos]
        out = np.zeros((len(keys), int(n_candidates)), dtype=np.float32)
        if candidates is None:
            out[rows, ents] = -1.0
        else:
            cand = np.asarray(candidates, dtype=np.int64)
            j = np.searchsorted(cand, ents)
            ok = (j < len(cand)) & (cand[np.minimum(j, len(cand) - 1)] == ents)
            out[rows[ok], j[ok]] = -1.0
        return out


def eval_batches(heads, rels, tails, n_entities, batch_size, neg_head, candidates=None, known=None):
    """Yield (pos_g, neg_g) where every positive is ranked against `candidates` (default: all entities):
    one chunk per batch, chunk_size = batch, neg_sample_size = #candidates (the reference views a
    full-entity negative graph as one chunk, sampler.py:486-490).  known: a TripleFilter -> neg_g.edata['bias']
    marks the candidates that are true triples (filtered evaluation)."""
    cand = np.arange(n_entities) if candidates is None else np.sort(np.asarray(candidates))
    cand_t = torch.from_numpy(cand.astype(np.int64))
    for s in range(0, len(heads), batch_size):
        h, r, t = heads[s:s + batch_size], rels[s:s + batch_size], tails[s:s + batch_size]
        neg_g = NegGraph(cand_t, 1, len(h), len(cand), neg_head)
        if known is not None:
            neg_g.edata["bias"] = torch.from_numpy(known.bias(h, r, t, len(cand), neg_head,
                                                              None if candidates is None else cand))
        yield build_pos_graph(h, r, t), neg_g
