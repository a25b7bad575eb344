"""Edge / negative sampler of the training path (reference: DGL's C++ EdgeSampler as dglke uses it,
dataloader/sampler.py:376-419 create_sampler, :459-512 chunk layout, :823-876 tail/head alternation).

Two implementations of ONE counter-based algorithm (kge_sampler.cu states it):

  DeviceSampler  kge_sampler_* of libkge_b200: the indices are produced in GPU memory and handed to the step without ever
                 visiting the host (what `dglke_b200.train` uses)
  HostSampler    the same integer arithmetic in numpy -- the seeded host sampler the device one is checked against, bit for
                 bit (tests/test_sampler.py), and a drop-in source of host batches

Algorithm: step k of a partition with E edges and batch B -- epoch = k // (E // B), j = k % (E // B); positive i is edge
perm_epoch(j*B + i) where perm_epoch is a 4-round Feistel network on 2^(2*hb) >= E keyed by (seed, epoch) with cycle walking
(a bijection of [0, E): every edge once per epoch, ragged tail dropped); negative j is splitmix64(seed, k, j) mod n_entities
(uniform, with replacement); node list = distinct endpoints of [heads | tails] in order of first appearance; even k corrupt
tails, odd k heads."""
import ctypes as C

import numpy as np
import torch

from . import _lib

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u64(x):
    return np.asarray(x, dtype=np.uint64)


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = _u64(z) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def half_bits(n_edges):
    hb = 1
    while (1 << (2 * hb)) < n_edges:
        hb += 1
    return hb


def feistel_perm(x, n, hb, key):
    """Bijection of [0, n): 4 Feistel rounds on 2*hb bits, cycle walking for values >= n."""
    x = _u64(x).copy()
    mask = np.uint64((1 << hb) - 1)
    hbv = np.uint64(hb)
    todo = np.ones(x.shape, dtype=bool)
    with np.errstate(over="ignore"):
        while todo.any():
            v = x[todo]
            L, R = v >> hbv, v & mask
            for r in range(4):
                f = mix64(R ^ (np.uint64(key) + np.uint64(r) * np.uint64(0xD1B54A32D192ED03))) & mask
                L, R = R, L ^ f
            v = (L << hbv) | R
            x[todo] = v
            todo[todo] = v >= np.uint64(n)
    return x


class HostSampler:
    """numpy restatement of kge_sampler.cu; heads/rels/tails: the partition's edge list (int64 arrays)."""

    def __init__(self, heads, rels, tails, n_entities, batch_size, neg_sample_size, seed=0):
        self.h, self.r, self.t = (np.ascontiguousarray(a, dtype=np.int64) for a in (heads, rels, tails))
        self.n_edges, self.n_ent = len(self.h), int(n_entities)
        self.B, self.Ns = int(batch_size), int(neg_sample_size)
        if self.B >= self.Ns and self.B % self.Ns:
            raise ValueError("batch_size must be a multiple of neg_sample_size (utils.get_compatible_batch_size)")
        if self.n_edges < self.B:
            raise ValueError("fewer edges (%d) than batch_size (%d)" % (self.n_edges, self.B))
        self.num_chunks = self.B // self.Ns if self.B >= self.Ns else 1
        self.chunk_size = self.B // self.num_chunks
        self.seed = np.uint64(seed)
        self.hb = half_bits(self.n_edges)

    def edge_ids(self, step):
        per_epoch = self.n_edges // self.B
        epoch, j = divmod(int(step), per_epoch)
        with np.errstate(over="ignore"):
            key = mix64(self.seed ^ (np.uint64(0xA0761D6478BD642F) * np.uint64(epoch + 1)))
        x = np.arange(j * self.B, (j + 1) * self.B, dtype=np.uint64)
        return feistel_perm(x, self.n_edges, self.hb, key).astype(np.int64)

    def sample(self, step):
        e = self.edge_ids(step)
        h, r, t = self.h[e], self.r[e], self.t[e]
        Nn = self.num_chunks * self.Ns
        with np.errstate(over="ignore"):
            base = mix64(self.seed + np.uint64(0x632BE59BD9B4E019) * np.uint64(step + 1))
            neg = (mix64(base + np.arange(Nn, dtype=np.uint64)) % np.uint64(self.n_ent)).astype(np.int64)
        keys = np.concatenate([h, t])
        uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")           # distinct ids in order of first appearance
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        loc = rank[inv]
        return dict(head=h, rel=r, tail=t, neg=neg, node_ids=uniq[order], head_local=loc[:self.B],
                    tail_local=loc[self.B:], neg_head=bool(step & 1))


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class DeviceBatch:
    """A sampled batch in GPU memory: the kge_batch_t the step consumes + torch views for inspection."""

    def __init__(self, cbatch, B, Nn, device, neg_head):
        self.c, self.B, self.Nn, self.device, self.neg_head = cbatch, B, Nn, device, bool(neg_head)

    def _view(self, ptr, n):
        return torch.as_tensor(_DevArray(ptr, n), device=self.device)

    def n_nodes(self):
        return int(self._view(self.c.n_nodes_dev, 1).item())

    def tensors(self):
        """(node_ids[:n], head_local, tail_local, rel_ids, neg_ids) as int64 CUDA tensors (copies)."""
        n = self.n_nodes()
        c = self.c
        return (self._view(c.node_ids, n).clone(), self._view(c.head_local, self.B).clone(),
                self._view(c.tail_local, self.B).clone(), self._view(c.rel_ids, self.B).clone(),
                self._view(c.neg_ids, self.Nn).clone())


class DeviceSampler:
    def __init__(self, heads, rels, tails, n_entities, batch_size, neg_sample_size, seed=0, device=0):
        self.h = _lib.get_handle(device)
        dev = self.h.device
        self.edges = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64)).to(dev) if not torch.is_tensor(a)
                      else a.to(dev, torch.int64).contiguous() for a in (heads, rels, tails)]
        self.n_edges = self.edges[0].numel()
        self.B, self.Ns = int(batch_size), int(neg_sample_size)
        self.num_chunks = self.B // self.Ns if self.B >= self.Ns else 1
        self.chunk_size = self.B // self.num_chunks
        self._s = C.c_void_p()
        _lib.check(self.h.lib.kge_sampler_create(self.h.raw, self.edges[0].data_ptr(), self.edges[1].data_ptr(),
                                                 self.edges[2].data_ptr(), self.n_edges, int(n_entities), self.B, self.Ns,
                                                 C.c_uint64(int(seed)), C.byref(self._s)))
        self.step = 0

    def sample(self, step=None):
        if step is None:
            step = self.step
            self.step += 1
        b = _lib.Batch()
        nh = C.c_int32()
        _lib.check(self.h.lib.kge_sampler_sample(self._s, int(step), C.byref(b), C.byref(nh), self.h.stream()))
        return DeviceBatch(b, self.B, self.num_chunks * self.Ns, self.h.device, nh.value)

    def close(self):
        if getattr(self, "_s", None):
            self.h.lib.kge_sampler_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
