"""CPU tests (gloo, world_size 2) of the multi-GPU plumbing: image sharding and max-over-ranks."""
import os
import socket

import torch
import torch.multiprocessing as mp

from wsovod_b200 import shard


def test_images_of_rank_partition():
    for world in (1, 2, 4, 8):
        for n in (1, 7, 8, 64):
            got = sorted(i for r in range(world) for i in shard.images_of_rank(n, r, world))
            assert got == list(range(n))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = shard.init("gloo")
    shard.barrier()
    mx = shard.max_over_ranks(1.0 + rank)
    sm = shard.sum_over_ranks(10.0 * (rank + 1))
    # every rank scores its own images with the CPU oracle: no data-path collective is needed
    import oracle
    from wsovod_b200 import synth
    g = synth.gen(5)
    C, D = synth.mil_logits(40, 6, g)
    mine = shard.images_of_rank(4, r, w)
    off = [0, 10, 20, 30, 40]
    parts = {i: oracle.mil(C[off[i]:off[i + 1]], D[off[i]:off[i + 1]], [0, 10])[1] for i in mine}
    q.put((rank, mx, sm, {i: v.tolist() for i, v in parts.items()}))
    torch.distributed.destroy_process_group()


def test_gloo_world2_sharding_and_reductions():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert all(abs(r[1] - 2.0) < 1e-12 and abs(r[2] - 30.0) < 1e-12 for r in res)
    merged = {}
    for r in res:
        merged.update(r[3])
    import oracle
    from wsovod_b200 import synth
    g = synth.gen(5)
    C, D = synth.mil_logits(40, 6, g)
    _, img = oracle.mil(C, D, [0, 10, 20, 30, 40])
    for i in range(4):
        torch.testing.assert_close(torch.tensor(merged[i]), img[i:i + 1], rtol=1e-6, atol=1e-9)
