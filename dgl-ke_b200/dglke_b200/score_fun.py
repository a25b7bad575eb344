"""score_func plugins (reference: models/pytorch/score_fun.py) over the CUDA library.

Each class keeps the reference's plugin surface -- edge_func(edges) -> {'score'}, infer(h, r, t),
prepare, create_neg_prepare(neg_head) -> fn, create_neg(neg_head) -> fn(heads, relations, tails,
num_chunks, chunk_size, neg_sample_size), forward(g), update, reset_parameters, save/load -- and
routes the arithmetic to kge_score_pos / kge_score_neg.  These stand-alone entry points are
forward-only (evaluation, inference, parity tests); training gradients come from the fused
kge_forward_backward that KEModel.forward calls.
"""
import torch as th

from . import engine as E


class _ScoreBase(th.nn.Module):
    model_name = None

    def __init__(self, hyper):
        super().__init__()
        self.hp = hyper

    def edge_func(self, edges):
        return {"score": E.score_pos(self.hp, edges.src["emb"], edges.data["emb"], edges.dst["emb"])}

    def infer(self, head_emb, rel_emb, tail_emb):
        """all (head, rel, tail) combinations -> [n_head, n_rel, n_tail]"""
        nh, nr, nt = head_emb.shape[0], rel_emb.shape[0], tail_emb.shape[0]
        h = head_emb.unsqueeze(1).expand(nh, nr, head_emb.shape[1]).reshape(nh * nr, -1)
        r = rel_emb.unsqueeze(0).expand(nh, nr, rel_emb.shape[1]).reshape(nh * nr, -1)
        s = E.score_neg(self.hp, h, r, tail_emb, 1, nh * nr, nt, False)
        return s.reshape(nh, nr, nt)

    def prepare(self, g, gpu_id, trace=False):
        pass

    def create_neg_prepare(self, neg_head):
        def fn(rel_id, num_chunks, head, tail, gpu_id, trace=False):
            return head, tail
        return fn

    def forward(self, g):
        g.apply_edges(lambda edges: self.edge_func(edges))

    def update(self, gpu_id=-1):
        pass

    def reset_parameters(self):
        pass

    def save(self, path, name):
        pass

    def load(self, path, name):
        pass

    def create_neg(self, neg_head):
        hp = self.hp

        def fn(heads, relations, tails, num_chunks, chunk_size, neg_sample_size):
            return E.score_neg(hp, heads, relations, tails, num_chunks, chunk_size, neg_sample_size, neg_head)
        return fn


def _hyper(model, hidden_dim, gamma=12.0, double_ent=False, double_rel=False):
    return E.Hyper(model=model, hidden_dim=hidden_dim, gamma=gamma, double_ent=double_ent, double_rel=double_rel)


class TransEScore(_ScoreBase):
    """gamma - |h + r - t|_p  (score_fun.py:40-108)"""

    def __init__(self, gamma, dist_func="l2", hidden_dim=None):
        super().__init__(_hyper("TransE_l1" if dist_func == "l1" else "TransE_l2", hidden_dim or 0, gamma))
        self.gamma = gamma
        self.dist_ord = 1 if dist_func == "l1" else 2


class DistMultScore(_ScoreBase):
    """sum h * r * t  (score_fun.py:222-286)"""

    def __init__(self, hidden_dim=None):
        super().__init__(_hyper("DistMult", hidden_dim or 0))


class ComplExScore(_ScoreBase):
    """Re <h, r, conj t>, rows are [re | im]  (score_fun.py:289-376)"""

    def __init__(self, hidden_dim=None):
        super().__init__(_hyper("ComplEx", hidden_dim or 0))


class RESCALScore(_ScoreBase):
    """h^T M_r t with M_r = rel.view(relation_dim, entity_dim)  (score_fun.py:378-449)"""

    def __init__(self, relation_dim, entity_dim):
        super().__init__(_hyper("RESCAL", entity_dim))
        self.relation_dim, self.entity_dim = relation_dim, entity_dim

    def infer(self, head_emb, rel_emb, tail_emb):
        """h^T M_r t for every combination (score_fun.py:397-402).  The tail-mode negative path of the reference
        computes (M_r h).t' = h^T M_r^T t' (score_fun.py:437-447), so the generic route through it would transpose
        M_r; the head-mode path is (M_r t).h', which IS the edge score: positives = all (rel, tail) pairs,
        'negatives' = the heads."""
        nh, nr, nt = head_emb.shape[0], rel_emb.shape[0], tail_emb.shape[0]
        r = rel_emb.unsqueeze(1).expand(nr, nt, rel_emb.shape[1]).reshape(nr * nt, -1)
        t = tail_emb.unsqueeze(0).expand(nr, nt, tail_emb.shape[1]).reshape(nr * nt, -1)
        s = E.score_neg(self.hp, head_emb, r, t, 1, nr * nt, nh, True)
        return s.reshape(nr, nt, nh).permute(2, 0, 1).contiguous()


class RotatEScore(_ScoreBase):
    """gamma - sum_k |h_k e^{i theta_k} - t_k|, theta = r / (emb_init / pi)  (score_fun.py:451-554)"""

    def __init__(self, gamma, emb_init, hidden_dim=None):
        super().__init__(_hyper("RotatE", hidden_dim or 0, gamma, double_ent=True))
        self.gamma, self.emb_init = gamma, emb_init


def _bind_dims(score_func, hyper):
    """KEModel hands the model's real Hyper (dims, gamma) to the plugin."""
    score_func.hp = hyper
    return score_func
