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

    def __init__(self, n_entities, n_relations, batch_size, neg_sample_size, seed=0, rank=0):
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
