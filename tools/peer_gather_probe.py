"""Experiment (2 GPUs): how fast can a kernel gather random 1600-byte rows from a PEER GPU's HBM, as a function of the
shard size and of how the peer mapping was made (cudaMalloc + CUDA IPC vs torch symmetric memory = CUDA VMM)?
torchrun --nproc-per-node 2 tools/peer_gather_probe.py"""
import ctypes as C, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dgl-ke_b200")):
    sys.path.insert(0, p)
from dglke_b200 import _lib
from dglke_b200.engine import DeviceTable, gather
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
h = _lib.get_handle(rank)
lib = h.lib
D = 400

def time_gather(table, idx, reps=5):
    out = None
    for _ in range(2):
        out = gather(table, idx)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); out = gather(table, idx); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev) * 1e3      # us


def time_gather_cold(table, lo, hi, reps=6):
    """fresh random rows every repetition: cold L2 on the owner, cold TLB on the reader"""
    idxs = [torch.randint(lo, hi, (14800,), device=dev) for _ in range(reps)]
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for (a, b), ix in zip(ev, idxs):
        a.record(); gather(table, ix); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3

class Ext:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}

for rows in (8000, 40_000_000):
    n = rows * world
    res = {}
    # --- A: cudaMalloc + IPC
    p_emb = C.c_void_p()
    _lib.check(lib.kge_device_alloc(h.raw, rows * D * 4, C.byref(p_emb)))
    torch.as_tensor(Ext(p_emb.value, (rows, D)), device=dev).uniform_(-0.05, 0.05)      # touched, incompressible contents
    st = torch.zeros(rows, device=dev)
    hbuf = C.create_string_buffer(64); off = C.c_int64()
    _lib.check(lib.kge_ipc_export(h.raw, p_emb, hbuf, C.byref(off)))
    hs = C.create_string_buffer(64); offs = C.c_int64()
    _lib.check(lib.kge_ipc_export(h.raw, C.c_void_p(st.data_ptr()), hs, C.byref(offs)))
    allh = [None] * world
    dist.all_gather_object(allh, (hbuf.raw, off.value, hs.raw, offs.value))
    emb_ptrs, st_ptrs = [], []
    for r in range(world):
        if r == rank:
            emb_ptrs.append(p_emb.value); st_ptrs.append(st.data_ptr())
        else:
            o1, o2 = C.c_void_p(), C.c_void_p()
            _lib.check(lib.kge_ipc_open(h.raw, allh[r][0], allh[r][1], C.byref(o1)))
            _lib.check(lib.kge_ipc_open(h.raw, allh[r][2], allh[r][3], C.byref(o2)))
            emb_ptrs.append(o1.value); st_ptrs.append(o2.value)
    tabA = DeviceTable(emb_ptrs, st_ptrs, n, D, devices=list(range(world)))
    rng = np.random.default_rng(rank)
    peer = (rank + 1) % world
    idx_peer = torch.from_numpy(rng.integers(peer * rows, (peer + 1) * rows, 14800)).to(dev)
    idx_loc = torch.from_numpy(rng.integers(rank * rows, (rank + 1) * rows, 14800)).to(dev)
    dist.barrier(); torch.cuda.synchronize()
    res["ipc_peer"] = time_gather(tabA, idx_peer); res["ipc_local"] = time_gather(tabA, idx_loc)
    res["ipc_peer_COLD"] = time_gather_cold(tabA, peer * rows, (peer + 1) * rows)
    res["ipc_local_COLD"] = time_gather_cold(tabA, rank * rows, (rank + 1) * rows)
    # sorted ids (np.unique order) and half/half mixes, as the step's node gather sees them
    mix = torch.sort(torch.cat([idx_peer[:7400], idx_loc[:7400]]))[0].contiguous()
    res["ipc_mix_sorted"] = time_gather(tabA, mix)
    # remote Adagrad scatter (state atomics + red.add rows) on the same rows
    from dglke_b200.engine import adagrad
    g = torch.randn(14800, D, device=dev) * 1e-3
    def t_ada(idx):
        adagrad(tabA, idx, g, 0.1); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); adagrad(tabA, idx, g, 0.1); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3
    dist.barrier(); torch.cuda.synchronize()
    res["ipc_adagrad_peer"] = t_ada(idx_peer); res["ipc_adagrad_local"] = t_ada(idx_loc)
    dist.barrier(); torch.cuda.synchronize()
    # --- B: symmetric memory (CUDA VMM)
    try:
        t = symm.empty(rows * D, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        ptrsB = [int(p) for p in hdl.buffer_ptrs]
        tabB = DeviceTable(ptrsB, st_ptrs, n, D, devices=list(range(world)))
        dist.barrier(); torch.cuda.synchronize()
        res["symm_peer"] = time_gather(tabB, idx_peer); res["symm_local"] = time_gather(tabB, idx_loc)
        res["symm_peer_COLD"] = time_gather_cold(tabB, peer * rows, (peer + 1) * rows)
        res["symm_local_COLD"] = time_gather_cold(tabB, rank * rows, (rank + 1) * rows)
        dist.barrier(); torch.cuda.synchronize()
        del tabB, hdl, t
    except Exception as e:
        res["symm_error"] = repr(e)[:200]
    if rank == 0:
        gb = 14800 * D * 4 / 1e9
        print("shard %.2f GB: " % (rows * D * 4 / 1e9) + ", ".join("%s %.1f us (%.0f GB/s)" % (k, v, gb / (v * 1e-6)) if isinstance(v, float) else "%s %s" % (k, v) for k, v in res.items()), flush=True)
    _lib.check(lib.kge_device_free(h.raw, p_emb))
    torch.cuda.empty_cache()
dist.barrier()
os._exit(0)
