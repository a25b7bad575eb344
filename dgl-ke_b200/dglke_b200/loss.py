"""LossGenerator (reference: models/pytorch/loss.py:41-98, models/base_loss.py) over kge_loss_grad.

The accelerated hot path implements the Logsigmoid criterion with optional self-adversarial
negative weighting and edge-importance weights (the configuration every example script of the
reference uses).  Hinge / Logistic / BCE / pairwise are outside this round's scope:
asking for them raises instead of silently running something else."""
import torch as th

from . import engine as E


class LazyLog(dict):
    """log dict {'pos_loss','neg_loss','loss','regularization'} whose floats are read from the device
    only when somebody looks (the reference pays 3-4 .item() syncs per step, tensor_models.py:55)."""
    KEYS = ("pos_loss", "neg_loss", "loss", "regularization")

    def __init__(self, log4, has_reg=True, lazy=False):
        super().__init__()
        # lazy: the scalars are written by a kernel that has not been enqueued yet (fused step: the update kernel
        # reduces them); an event recorded right after that kernel would be ideal, reading on first use after the
        # caller's update() is what the train loop does
        self._log4 = log4 if lazy else log4.clone()
        self._keys = self.KEYS if has_reg else self.KEYS[:3]
        self._vals = None

    def _load(self):
        if self._vals is None:
            v = self._log4.cpu().tolist()
            self._vals = dict(zip(self.KEYS, v))
            for k in self._keys:
                dict.__setitem__(self, k, self._vals[k])
        return self._vals

    def __getitem__(self, k):
        self._load()
        return dict.__getitem__(self, k)

    def keys(self):
        return list(self._keys)

    def items(self):
        self._load()
        return dict.items(self)

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def __contains__(self, k):
        return k in self._keys


class FusedLoss:
    """What KEModel.forward returns as `loss`: the fused step has already produced every gradient,
    so backward() has nothing left to do (train_pytorch.py:145 keeps working unchanged)."""

    def __init__(self, log4, with_reg, lazy=False):
        self._log4, self._with_reg = log4, with_reg

    def backward(self):
        return None

    def item(self):
        v = self._log4.cpu().tolist()
        return v[2] + (v[3] if self._with_reg else 0.0)

    def detach(self):
        return self

    def __float__(self):
        return self.item()


class LossGenerator:
    def __init__(self, args, loss_genre="Logsigmoid", neg_adversarial_sampling=False, adversarial_temperature=1.0,
                 pairwise=False):
        if loss_genre != "Logsigmoid" or pairwise:
            raise NotImplementedError("the B200 hot path implements loss_genre=Logsigmoid (optionally -adv); "
                                      "%s%s is not accelerated yet" % (loss_genre, " pairwise" if pairwise else ""))
        self.pairwise = False
        self.neg_adversarial_sampling = bool(neg_adversarial_sampling)
        self.adversarial_temperature = adversarial_temperature if neg_adversarial_sampling else 0
        self.neg_label = -1

    def _hyper(self):
        return E.Hyper(model="DistMult", hidden_dim=4, adversarial=self.neg_adversarial_sampling,
                       adv_temperature=float(self.adversarial_temperature or 1.0))

    def get_total_loss(self, pos_score, neg_score, edge_weight=None):
        """-> (loss 0-dim tensor, log).  Forward-only stand-alone op; d loss / d score is available
        through score_gradients()."""
        log4, _, _ = E.loss_grad(self._hyper(), pos_score, neg_score, edge_weight)
        log = LazyLog(log4, has_reg=False)
        return log4[2], log

    def score_gradients(self, pos_score, neg_score, edge_weight=None):
        _, dpos, dneg = E.loss_grad(self._hyper(), pos_score, neg_score, edge_weight)
        return dpos, dneg
