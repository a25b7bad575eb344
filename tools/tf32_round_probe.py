"""Experiment: how does tcgen05 kind::tf32 treat the low 13 mantissa bits of a raw fp32 operand?
Runs the fused step on tables whose values are all exact TF32 ties (low 13 bits = 0x1000) and on generic values, with the
operand split hi = raw fp32, lo = x - round_mode(x) for round_mode in {trunc, rna, rne}; the mode that matches the fp64
oracle on the tie tables is the hardware's.   KGE_B200_SPLIT_TRUNC=<1|2|3> python tools/tf32_round_probe.py"""
import os, sys
import numpy as np, torch as th
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dgl-ke_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import kge_oracle as ko
from test_gpu_parity import _random_step, _engine
from dglke_b200 import _lib

def ties(t):
    b = t.view(th.int32)
    return ((b & ~0x1FFF) | 0x1000).view(th.float32)

for tie in (False, True):
    hp = ko.Hyper(model="DistMult", hidden_dim=128, gamma=12.0, lr=0.1, reg_coef=0.0, adversarial=False)
    ent, es, rel, rs = ko.init_tables(hp, 2000, 10, seed=3)
    rel.fill_(1.0)           # a = h * 1: the A operand carries the raw table bits
    if tie:
        ent = ties(ent)
    si, C = _random_step(hp, 2000, 10, 256, 64, 64, False, seed=1)
    t64 = [x.double() for x in (ent, es, rel, rs)]
    h = t64[0][si["node_ids"]][si["head_local"]]
    S64 = ko.negative_score(hp, h, t64[2][si["rel_ids"]], t64[0][si["neg_ids"]], C, 64, 64, False).reshape(-1, 64)
    eng, _ = _engine(hp, ent, es, rel, rs)
    d = lambda t: t.to(eng.device)
    eng.forward_backward(d(si["node_ids"]), d(si["head_local"]), d(si["tail_local"]), d(si["rel_ids"]), d(si["neg_ids"]), 64, 64, False)
    S = eng.read(_lib.BUF_NEG_SCORE, (256, 64)).double().cpu()
    err = (S - S64).abs().max().item() / S64.abs().max().item()
    print("split mode %s, %s tables: max rel err of the scores %.3e" % (os.environ.get("KGE_B200_SPLIT_TRUNC", "0"), "TIE" if tie else "generic", err))
