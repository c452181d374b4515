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


def _ddp_worker(rank, world, port, q):
    """the training step's DDP plumbing on CPU (gloo): the FC gradient stand-in of wsovod_b200/steps.py under
    DistributedDataParallel -- zero gradients of the FC layers' shape on every rank, the real layer's gradient averaged
    over ranks, `no_sync()` leaving it local (what bench.py's ms_per_step_no_allreduce run does)"""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    shard.init("gloo")
    from torch.nn.parallel import DistributedDataParallel as DDP
    from wsovod_b200 import steps
    steps.FC_WIDTH, old = 8, steps.FC_WIDTH
    try:
        torch.manual_seed(0)
        head = steps.StandInBoxHead(in_features=4 * 49, width=16, fc_in=4 * 49)
        model = torch.nn.Sequential(head, torch.nn.Linear(16, 3))
        ddp = DDP(model, gradient_as_bucket_view=True)
        torch.manual_seed(100 + rank)
        x = torch.randn(5, 4, 7, 7)
        ddp(x).square().sum().backward()
        g_sync = model[1].weight.grad.clone()
        for p in model.parameters():
            p.grad = None
        with ddp.no_sync():
            ddp(x).square().sum().backward()
        g_local = model[1].weight.grad.clone()
        q.put((rank, g_sync.tolist(), g_local.tolist(), float(head.fc1_standin.grad.abs().sum()),
               tuple(head.fc1_standin.grad.shape), tuple(head.fc2_standin.grad.shape), head.grad_bytes()))
    finally:
        steps.FC_WIDTH = old
    torch.distributed.destroy_process_group()


def test_gloo_world2_ddp_gradient_standin():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, l0, z0, sh1, sh2, nb), (r1, s1, l1, z1, _, _, _) = res
    assert z0 == 0.0 and z1 == 0.0 and sh1 == (4 * 49, 8) and sh2 == (8, 8) and nb == 4 * (4 * 49 * 8 + 64)
    s0, s1, l0, l1 = (torch.tensor(v) for v in (s0, s1, l0, l1))
    torch.testing.assert_close(s0, s1)                                   # all-reduced: identical on both ranks
    torch.testing.assert_close(s0, (l0 + l1) / 2, rtol=1e-5, atol=1e-6)   # = the mean of the local gradients
    assert not torch.allclose(l0, l1)
