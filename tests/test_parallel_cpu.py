"""Host-side multi-rank logic on CPU (gloo, world_size 2): view sharding and the flat gradient
arena all-reduce of sugar_b200.parallel.  No CUDA involved."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sugar_b200 import parallel
    P, M = 50, 16
    g = torch.Generator().manual_seed(100 + rank)
    shapes = dict(means3D=(P, 3), shs=(P, M, 3), opacities=(P, 1), scales=(P, 3), rotations=(P, 4))
    params = {k: torch.zeros(s, requires_grad=True) for k, s in shapes.items()}
    for k, p in params.items():
        p.grad = torch.randn(p.shape, generator=g)
    local = {k: p.grad.clone() for k, p in params.items()}
    arena = parallel.GradArena(P, M, "cpu")
    arena.all_reduce_from(params, 1.0 / world)
    arena.unpack_to(params)
    # numpy, not torch tensors: a tensor travels as a file descriptor that the parent must still be able
    # to open after this process has exited (ConnectionResetError when the worker wins the race)
    q.put((rank, {k: v.numpy().copy() for k, v in local.items()},
           {k: p.grad.detach().numpy().copy() for k, p in params.items()}, parallel.shard_views(8, rank, world)))
    dist.barrier()
    dist.destroy_process_group()
    q.close()
    q.join_thread()  # the feeder thread has written everything before the process exits


def test_arena_allreduce_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:  # a free port for the rendezvous
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    for k in res[0][1]:
        mean = (res[0][1][k] + res[1][1][k]) / world
        for r in res:
            assert torch.allclose(torch.from_numpy(r[2][k]), torch.from_numpy(mean), atol=1e-7), k
    assert res[0][3] == [0, 2, 4, 6] and res[1][3] == [1, 3, 5, 7]
    assert 59 * 50 == sum(n for _, n in __import__("sugar_b200.parallel", fromlist=["x"]).GradArena(50, 16, "cpu").offsets.values())


def test_arena_reduces_backward_buffer_in_place():
    """Gradients handed out by an autograd Function as slices of one flat buffer (what sugar_b200's
    backward does) are recognised and reduced in place: no packing copy."""
    import torch
    from sugar_b200 import parallel
    P, M = 10, 2
    widths = (3, 1, 3, 4, 3 * M, 3, 3, 6)

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *a):
            return sum(x.sum() for x in a).reshape(1)

        @staticmethod
        def backward(ctx, g):
            flat = torch.arange(8 + P * sum(widths) + 20, dtype=torch.float32)[8:]  # nonzero storage offset
            outs, o = [], 0
            for k in widths[:5]:
                outs.append(flat[o:o + P * k])
                o += P * k
            return (outs[0].view(P, 3), outs[1].view(P, 1), outs[4].view(P, M, 3), outs[2].view(P, 3),
                    outs[3].view(P, 4))

    ps = dict(means3D=torch.zeros(P, 3), opacities=torch.zeros(P, 1), shs=torch.zeros(P, M, 3),
              scales=torch.zeros(P, 3), rotations=torch.zeros(P, 4))
    ps = {k: v.requires_grad_(True) for k, v in ps.items()}
    Fn.apply(*ps.values()).sum().backward()
    arena = parallel.GradArena(P, M, "cpu")
    buf = arena._shared_base(ps)
    assert buf is not None and buf.numel() == arena.flat.numel()
    assert buf.data_ptr() == ps["means3D"].grad.data_ptr()
    out = arena.all_reduce_from(ps, scale=2.0)
    assert out.data_ptr() == buf.data_ptr()
    assert float(ps["means3D"].grad[0, 0]) == 16.0  # scaled through the alias
    # separate tensors (e.g. after gradient accumulation) fall back to packing
    ps2 = {k: v.detach().clone().requires_grad_(True) for k, v in ps.items()}
    for v in ps2.values():
        v.grad = torch.ones_like(v)
    assert arena._shared_base(ps2) is None
