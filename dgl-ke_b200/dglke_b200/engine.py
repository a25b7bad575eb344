"""StepEngine: thin, allocation-free driver of the C ABI for one (model, tables) pair.

It owns the ctypes structs (kge_table_t / kge_step_cfg_t) and calls
kge_forward_backward / kge_update / kge_step_fused[_host]; KEModel (general_models.py) and
bench.py are built on it.  Tables are plain torch CUDA tensors (or peer-mapped shards, dist.py).
"""
import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class Hyper:
    """The hyper-parameters KEModel / LossGenerator / ExternalEmbedding read from `args`
    (reference: models/general_models.py:208-236, models/pytorch/loss.py:41-62)."""
    model: str = "TransE_l2"
    hidden_dim: int = 400
    gamma: float = 12.0
    lr: float = 0.01
    reg_coef: float = 2e-6
    reg_norm: int = 3
    adversarial: bool = False
    adv_temperature: float = 1.0
    double_ent: bool = False
    double_rel: bool = False
    loss_genre: str = "Logsigmoid"      # Hinge | Logistic | Logsigmoid | BCE  (loss.py:41-62)
    margin: float = 1.0
    pairwise: bool = False
    neg_deg_sample: bool = False        # general_models.py:396-403,417-424: training steps only

    @property
    def emb_init(self):
        return (self.gamma + 2.0) / self.hidden_dim

    @property
    def entity_dim(self):
        return 2 * self.hidden_dim if self.double_ent else self.hidden_dim

    @property
    def relation_dim(self):
        rd = 2 * self.hidden_dim if self.double_rel else self.hidden_dim
        return rd * self.entity_dim if self.model == "RESCAL" else rd


class DeviceTable:
    """An embedding table + Adagrad state as the C ABI sees it (one or more row-range shards)."""

    def __init__(self, emb_shards, state_shards, num_rows, dim, devices=None):
        self.num_rows, self.dim = num_rows, dim
        self.emb_shards, self.state_shards = emb_shards, state_shards
        self.ctable, self._keep = _lib.make_table(emb_shards, state_shards, num_rows, dim, devices)

    @classmethod
    def from_tensors(cls, emb, state_sum):
        assert emb.is_cuda and emb.dtype == torch.float32 and emb.is_contiguous()
        assert state_sum.is_cuda and state_sum.dtype == torch.float32 and state_sum.numel() == emb.shape[0]
        return cls([emb], [state_sum], emb.shape[0], emb.shape[1], [emb.device.index])

    def ref(self):
        return C.byref(self.ctable)


class StepEngine:
    def __init__(self, hyper, ent, rel, device=None):
        self.hp = hyper
        self.ent, self.rel = ent, rel
        if device is None:
            device = torch.cuda.current_device()
        self.h = _lib.get_handle(device)
        self.device = self.h.device
        self.lib = self.h.lib
        assert ent.dim == hyper.entity_dim and rel.dim == hyper.relation_dim, "table dims do not match hyper"
        self.log4 = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._log4_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self._last = None

    def cfg(self, batch, chunk_size, neg_sample_size, neg_head):
        hp = self.hp
        return _lib.make_cfg(hp.model, hp.entity_dim, hp.relation_dim, hp.gamma, hp.emb_init, hp.lr, hp.reg_coef,
                             hp.reg_norm, hp.adversarial, hp.adv_temperature, neg_head, batch, chunk_size,
                             neg_sample_size, hp.loss_genre, hp.margin, hp.pairwise, hp.neg_deg_sample)

    def check_ids(self, node_ids, head_local, tail_local, rel_ids, neg_ids):
        """KGE_B200_CHECK_IDS=1 (debugging aid, costs a device sync): the kernels index the tables with the ids they are
        given; the reference's tensor indexing raises IndexError on an id outside the table, so does this."""
        if os.environ.get("KGE_B200_CHECK_IDS") != "1":
            return
        for name, t, hi in (("node_ids", node_ids, self.ent.num_rows), ("neg_ids", neg_ids, self.ent.num_rows),
                            ("rel_ids", rel_ids, self.rel.num_rows), ("head_local", head_local, node_ids.numel()),
                            ("tail_local", tail_local, node_ids.numel())):
            if t is not None and t.numel() and (int(t.min()) < 0 or int(t.max()) >= hi):
                raise IndexError("%s: index out of range [0, %d)" % (name, hi))

    # ---- the three-call shape of train_pytorch.py:141-152 -------------------------------------
    def forward_backward(self, node_ids, head_local, tail_local, rel_ids, neg_ids, chunk_size, neg_sample_size,
                         neg_head, edge_weight=None, log4=None):
        self.check_ids(node_ids, head_local, tail_local, rel_ids, neg_ids)
        cfg = self.cfg(head_local.numel(), chunk_size, neg_sample_size, neg_head)
        b, keep = _lib.make_batch(node_ids, head_local, tail_local, rel_ids, neg_ids, edge_weight)
        out = self.log4 if log4 is None else log4
        _lib.check(self.lib.kge_forward_backward(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(b),
                                                 out.data_ptr(), self.h.stream()))
        self._last = (cfg, b, keep)
        return out

    def update(self):
        if self._last is None:
            raise _lib.KgeError("update() without forward_backward()")
        cfg, b, keep = self._last
        cfg.lr = self.hp.lr
        _lib.check(self.lib.kge_update(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(b),
                                       self.h.stream()))
        self._last = None

    # ---- fused ------------------------------------------------------------------------------
    def step(self, node_ids, head_local, tail_local, rel_ids, neg_ids, chunk_size, neg_sample_size, neg_head,
             edge_weight=None, log4=None, head_ids=None, tail_ids=None):
        cfg = self.cfg(head_local.numel(), chunk_size, neg_sample_size, neg_head)
        b, keep = _lib.make_batch(node_ids, head_local, tail_local, rel_ids, neg_ids, edge_weight, head_ids, tail_ids)
        out = self.log4 if log4 is None else log4
        _lib.check(self.lib.kge_step_fused(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(b),
                                           out.data_ptr(), self.h.stream()))
        return out

    def step_sampled(self, batch, chunk_size, neg_sample_size, log4=None):
        """fused step on a DeviceBatch (dglke_b200.sampler.DeviceSampler): the indices never leave the GPU"""
        cfg = self.cfg(batch.B, chunk_size, neg_sample_size, batch.neg_head)
        out = self.log4 if log4 is None else log4
        _lib.check(self.lib.kge_step_fused(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(batch.c),
                                           out.data_ptr(), self.h.stream()))
        return out

    def announce_next(self, next_batch):
        """kge_set_next_batch: `next_batch` is the batch of the NEXT step_begin -- a sampler.DeviceBatch or a
        (node_ids, neg_ids) pair of int64 CUDA tensors that stay alive and unchanged until that step.  Its table rows are
        copied while this step's contraction runs (sharded tables; one-step-stale reads as under --async_update)."""
        if next_batch is None:
            _lib.check(self.lib.kge_set_next_batch(self.h.raw, None, 0))
            return
        if hasattr(next_batch, "c"):                                          # DeviceBatch
            nb, n_neg, keep = next_batch.c, next_batch.Nn, next_batch
        else:
            nodes, negs = next_batch
            assert nodes.is_cuda and negs.is_cuda and nodes.dtype == torch.int64 and negs.dtype == torch.int64
            nb = _lib.Batch(nodes.data_ptr(), nodes.numel(), None, None, None, negs.data_ptr(), None, None, None, None)
            n_neg, keep = negs.numel(), (nodes, negs)
        _lib.check(self.lib.kge_set_next_batch(self.h.raw, C.byref(nb), int(n_neg)))
        self._next_keep = keep

    def step_begin(self, node_ids, head_local=None, tail_local=None, rel_ids=None, neg_ids=None, chunk_size=None,
                   neg_sample_size=None, neg_head=None, edge_weight=None, next_batch=None):
        """first half of step(): everything up to the gradients (kge_step_fused_begin).  `node_ids` may be a
        sampler.DeviceBatch (then only chunk_size / neg_sample_size are read from the other arguments).
        next_batch: see announce_next()."""
        if next_batch is not None:
            self.announce_next(next_batch)
        if hasattr(node_ids, "c") and hasattr(node_ids, "neg_head"):        # DeviceBatch
            batch = node_ids
            cfg = self.cfg(batch.B, chunk_size, neg_sample_size, batch.neg_head)
            b, keep = batch.c, batch
        else:
            cfg = self.cfg(head_local.numel(), chunk_size, neg_sample_size, neg_head)
            b, keep = _lib.make_batch(node_ids, head_local, tail_local, rel_ids, neg_ids, edge_weight)
        _lib.check(self.lib.kge_step_fused_begin(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(b),
                                                 self.h.stream()))
        self._last = (cfg, b, keep)

    def step_end(self, log4=None):
        """second half of step(): the Adagrad update + log scalars (kge_step_fused_end)"""
        if self._last is None:
            raise _lib.KgeError("step_end() without step_begin()")
        cfg, b, keep = self._last
        cfg.lr = self.hp.lr
        out = self.log4 if log4 is None else log4
        _lib.check(self.lib.kge_step_fused_end(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(b),
                                               out.data_ptr(), self.h.stream()))
        self._last = None
        return out

    def step_host(self, node_ids, head_local, tail_local, rel_ids, neg_ids, chunk_size, neg_sample_size, neg_head,
                  edge_weight=None):
        """Index tensors are CPU tensors (as a sampler produces them); returns the pinned host
        log4 buffer, valid after sync()."""
        assert not node_ids.is_cuda
        cfg = self.cfg(head_local.numel(), chunk_size, neg_sample_size, neg_head)
        b, keep = _lib.make_batch(node_ids, head_local, tail_local, rel_ids, neg_ids, edge_weight)
        _lib.check(self.lib.kge_step_fused_host(self.h.raw, C.byref(cfg), self.ent.ref(), self.rel.ref(), C.byref(b),
                                                self._log4_host.data_ptr(), self.h.stream()))
        return self._log4_host

    def sync(self):
        _lib.check(self.lib.kge_sync(self.h.raw, self.h.stream()))

    # ---- introspection ----------------------------------------------------------------------
    def read(self, which, shape):
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.kge_debug_read(self.h.raw, which, out.data_ptr(), out.numel(), self.h.stream()))
        return out


# ---- stand-alone ops (one reference function each) ------------------------------------------
def _cfg_for(hp, batch, chunk_size, neg_sample_size, neg_head):
    return _lib.make_cfg(hp.model, hp.entity_dim, hp.relation_dim, hp.gamma, hp.emb_init, hp.lr, hp.reg_coef,
                         hp.reg_norm, hp.adversarial, hp.adv_temperature, neg_head, batch, chunk_size,
                         neg_sample_size, hp.loss_genre, hp.margin, hp.pairwise)


def gather(table, idx):
    """ExternalEmbedding.__call__ without trace: table rows at idx (bit exact)."""
    h = _lib.get_handle(idx.device.index)
    idx = idx.contiguous()
    out = torch.empty((idx.numel(), table.dim), dtype=torch.float32, device=idx.device)
    _lib.check(h.lib.kge_gather(h.raw, table.ref(), idx.data_ptr(), idx.numel(), out.data_ptr(), h.stream()))
    return out


def score_pos(hp, head, rel, tail):
    h = _lib.get_handle(head.device.index)
    n = head.shape[0]
    cfg = _cfg_for(hp, max(n, 1), 1, 1, False)
    out = torch.empty(n, dtype=torch.float32, device=head.device)
    _lib.check(h.lib.kge_score_pos(h.raw, C.byref(cfg), head.contiguous().data_ptr(), rel.contiguous().data_ptr(),
                                   tail.contiguous().data_ptr(), n, out.data_ptr(), h.stream()))
    return out


def score_neg(hp, heads, rels, tails, num_chunks, chunk_size, neg_sample_size, neg_head):
    h = _lib.get_handle(heads.device.index)
    cfg = _cfg_for(hp, num_chunks * chunk_size, chunk_size, neg_sample_size, neg_head)
    out = torch.empty((num_chunks, chunk_size, neg_sample_size), dtype=torch.float32, device=heads.device)
    heads, rels, tails = heads.contiguous(), rels.contiguous(), tails.contiguous()
    _lib.check(h.lib.kge_score_neg(h.raw, C.byref(cfg), heads.data_ptr(), rels.data_ptr(), tails.data_ptr(),
                                   out.data_ptr(), h.stream()))
    return out


def loss_grad(hp, pos, neg, edge_weight=None):
    """returns (log4 device tensor, dpos, dneg)"""
    h = _lib.get_handle(pos.device.index)
    B, Ns = neg.shape
    cfg = _cfg_for(hp, B, B, Ns, False)
    dpos, dneg = torch.empty_like(pos), torch.empty_like(neg)
    log4 = torch.zeros(4, dtype=torch.float32, device=pos.device)
    _lib.check(h.lib.kge_loss_grad(h.raw, C.byref(cfg), pos.contiguous().data_ptr(), neg.contiguous().data_ptr(),
                                   edge_weight.data_ptr() if edge_weight is not None else None, dpos.data_ptr(),
                                   dneg.data_ptr(), log4.data_ptr(), h.stream()))
    return log4, dpos, dneg


def adagrad(table, idx, grad, lr):
    h = _lib.get_handle(idx.device.index)
    idx, grad = idx.contiguous(), grad.contiguous()
    _lib.check(h.lib.kge_adagrad(h.raw, table.ref(), idx.data_ptr(), grad.data_ptr(), idx.numel(), float(lr),
                                 h.stream()))
