"""dglke_train-compatible entry point (reference: train.py:74-380, train_pytorch.py:110-197).

    python -m dglke_b200.train --model_name TransE_l2 --dataset FB15k --batch_size 1000 \
        --neg_sample_size 200 --hidden_dim 400 --gamma 19.9 --lr 0.25 --max_step 24000 -adv --gpu 0

All reference flags parse.  What differs, and why:
  * data: there is no network in this environment, so built-in dataset names select the dataset's SHAPE
    (entities / relations / training edges) and triples are drawn synthetically unless --data_files
    points at udd_hrt-style integer triple files (entity_file relation_file train_file [valid] [test]);
  * --gpu is required (no CPU path).  This CLI drives ONE GPU; the multi-GPU path (entity table row-sharded over the
    GPUs instead of --mix_cpu_gpu's host table, one process per GPU under torchrun) is the
    dglke_b200.dist.ShardedTrainer API that bench.py --gpus N uses -- wiring it into this CLI is on the next list;
  * sampling is numpy-based (DGL's C++ sampler is out of scope).
"""
import os
import sys
import time

import numpy as np
import torch as th

from .utils import ArgParser, get_compatible_batch_size, save_model, prepare_save_path
from .general_models import KEModel
from .graph import SyntheticSampler, TripleSampler, eval_batches

# (entities, relations, training edges): docs/source/benchmarks.rst dataset table
BUILTIN_SHAPES = {
    "FB15k": (14951, 1345, 483142), "FB15k-237": (14541, 237, 272115), "wn18": (40943, 18, 141442),
    "wn18rr": (40943, 11, 86835), "Freebase": (86054151, 14824, 304727650),
    "wikikg2": (2500604, 535, 16109182), "biokg": (93773, 51, 4762678),
}


def _read_udd(args):
    files = args.data_files
    if files is None or len(files) < 3:
        raise SystemExit("--format udd_* needs --data_files entity_file relation_file train_file [valid] [test]")
    path = lambda f: f if os.path.isabs(f) else os.path.join(args.data_path, f)
    count = lambda f: sum(1 for line in open(path(f)) if line.strip())
    n_ent, n_rel = count(files[0]), count(files[1])
    order = args.format.split("_")[-1]            # e.g. hrt
    col = {c: i for i, c in enumerate(order)}

    def triples(f):
        a = np.loadtxt(path(f), dtype=np.int64, delimiter=args.delimiter, ndmin=2)
        if a.size and (a.min() < 0 or a[:, col["h"]].max() >= n_ent or a[:, col["t"]].max() >= n_ent
                       or a[:, col["r"]].max() >= n_rel):
            raise ValueError("triple ids out of range in %s" % f)
        return a[:, col["h"]], a[:, col["r"]], a[:, col["t"]]
    train = triples(files[2])
    valid = triples(files[3]) if len(files) > 3 else None
    test = triples(files[4]) if len(files) > 4 else None
    return n_ent, n_rel, train, valid, test


def train(args, model, train_sampler, valid_batches=None, rank=0, barrier=None):
    """train_pytorch.py:110-197 -- same four phases and the same log lines."""
    gpu_id = args.gpu[rank % len(args.gpu)]
    logs = []
    train_start = start = time.time()
    sample_time = update_time = forward_time = backward_time = 0.0
    for step in range(0, args.max_step):
        t0 = time.time()
        pos_g, neg_g = next(train_sampler)
        sample_time += time.time() - t0
        t0 = time.time()
        loss, log = model.forward(pos_g, neg_g, gpu_id)
        forward_time += time.time() - t0
        t0 = time.time()
        loss.backward()
        backward_time += time.time() - t0
        t0 = time.time()
        model.update(gpu_id)
        update_time += time.time() - t0
        logs.append(log)
        if args.force_sync_interval > 0 and (step + 1) % args.force_sync_interval == 0 and barrier is not None:
            barrier()
        if (step + 1) % args.log_interval == 0:
            th.cuda.synchronize()
            for k in logs[0].keys():
                v = sum(l[k] for l in logs) / len(logs)
                print("[proc {}][Train]({}/{}) average {}: {}".format(rank, (step + 1), args.max_step, k, v))
            logs = []
            print("[proc {}][Train] {} steps take {:.3f} seconds".format(rank, args.log_interval, time.time() - start))
            print("[proc {}]sample: {:.3f}, forward: {:.3f}, backward: {:.3f}, update: {:.3f}".format(
                rank, sample_time, forward_time, backward_time, update_time))
            sample_time = update_time = forward_time = backward_time = 0.0
            start = time.time()
        if args.valid and (step + 1) % args.eval_interval == 0 and step > 1 and valid_batches is not None:
            test(args, model, valid_batches(), rank, mode="Valid")
    th.cuda.synchronize()
    print("proc {} takes {:.3f} seconds".format(rank, time.time() - train_start))


def test(args, model, batches, rank=0, mode="Test"):
    """train_pytorch.py:199-253 (non-wikikg90M branch): average MRR / MR / HITS@k over head and tail ranking."""
    gpu_id = args.gpu[rank % len(args.gpu)]
    logs = []
    with th.no_grad():
        for pos_g, neg_g in batches:
            model.forward_test(pos_g, neg_g, logs, gpu_id)
    metrics = {}
    if logs:
        for m in logs[0].keys():
            metrics[m] = sum(l[m] for l in logs) / len(logs)
    for k, v in metrics.items():
        print("[{}]{} average {}: {}".format(rank, mode, k, v))
    return metrics


def main(argv=None):
    args = ArgParser().parse_args(argv)
    if args.gpu[0] < 0:
        raise SystemExit("dglke_b200 needs --gpu: the hot path is a B200 CUDA library without a CPU fallback")
    prepare_save_path(args)
    args.eval_filter = not args.no_eval_filter
    args.strict_rel_part = args.soft_rel_part = False
    args.batch_size = get_compatible_batch_size(args.batch_size, args.neg_sample_size)
    args.batch_size_eval = get_compatible_batch_size(args.batch_size_eval, args.neg_sample_size_eval)
    if args.format.startswith("udd") or args.format.startswith("raw_udd"):
        if args.format.startswith("raw_udd"):
            raise SystemExit("raw_udd (string ids) is not supported yet: convert to integer udd files first")
        n_ent, n_rel, tr, va, te = _read_udd(args)
    else:
        if args.dataset not in BUILTIN_SHAPES:
            raise SystemExit("unknown built-in dataset %s" % args.dataset)
        n_ent, n_rel, n_edges = BUILTIN_SHAPES[args.dataset]
        print("NOTE: no network -- training on a synthetic graph of %s's shape (%d entities, %d relations)" % (
            args.dataset, n_ent, n_rel))
        tr = va = te = None
    th.cuda.set_device(args.gpu[0])
    model = KEModel(args, args.model_name, n_ent, n_rel, args.hidden_dim, args.gamma,
                    double_entity_emb=args.double_ent, double_relation_emb=args.double_rel)
    if tr is None:
        sampler = SyntheticSampler(n_ent, n_rel, args.batch_size, args.neg_sample_size, seed=0)
    else:
        sampler = TripleSampler(tr[0], tr[1], tr[2], n_ent, n_rel, args.batch_size, args.neg_sample_size, seed=0)

    def split_batches(split):
        def gen():
            for neg_head in (True, False):
                yield from eval_batches(split[0], split[1], split[2], n_ent, args.batch_size_eval, neg_head)
        return gen
    train(args, model, sampler, split_batches(va) if (args.valid and va is not None) else None)
    if not args.no_save_emb:
        save_model(args, model)
    if args.test and te is not None:
        test(args, model, split_batches(te)())
    return model


if __name__ == "__main__":
    main()
