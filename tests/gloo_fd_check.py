"""world_size-3 gloo check (CPU only) of the descriptor exchange dglke_b200.dist uses to share VMM shard allocations:
every rank hands two file descriptors to every peer over Unix sockets (SCM_RIGHTS); the received descriptors must open
the SENDER's files."""
import os
import sys
import tempfile

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dgl-ke_b200"))


def main():
    from dglke_b200.dist import exchange_fds
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = []
    for k in range(2):
        f = tempfile.TemporaryFile()
        f.write(b"rank %d file %d" % (rank, k))
        f.flush()
        mine.append(f)
    got = exchange_fds([f.fileno() for f in mine], rank, world)
    assert sorted(got) == [r for r in range(world) if r != rank], got
    for peer, fds in got.items():
        assert len(fds) == 2
        for k, fd in enumerate(fds):
            assert fd not in [f.fileno() for f in mine]
            os.lseek(fd, 0, os.SEEK_SET)
            assert os.read(fd, 64) == b"rank %d file %d" % (peer, k), (rank, peer, k)
            os.close(fd)
    # a second exchange in the same process group works (fresh socket names)
    got2 = exchange_fds([mine[0].fileno()], rank, world)
    assert all(len(v) == 1 for v in got2.values())
    for v in got2.values():
        os.close(v[0])
    dist.barrier()
    if rank == 0:
        print("GLOO_FD_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
