"""world_size-2 gloo test of the data-parallel clip sharding (scatter inputs / gather outputs)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from versband_b200.dist import run_data_parallel, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (1, 2, 7, 8, 32):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    midi = torch.randint(0, 130, (n_clips, 1, 16), generator=g) if rank == 0 else None
    cap = torch.randn(n_clips, 4, 8, generator=g) if rank == 0 else None

    def fake_pipeline(local):  # per-clip function: result must not depend on the sharding
        m, c = local
        return m.float().sum((1, 2))[:, None] + c.sum((1, 2))[:, None] * torch.ones(1, 5)

    out = run_data_parallel(fake_pipeline, [midi, cap] if rank == 0 else None, n_clips)
    out = out.clone() if rank == 0 else None  # (the gather buffer is reused by the next step)
    # second step through the cached exchange: no metadata handshake, same buffers
    out2 = run_data_parallel(fake_pipeline, [midi + 1, cap] if rank == 0 else None, n_clips)
    if rank == 0:
        q.put((out.tolist(), out2.tolist()))  # plain lists: tensors travel by fd and need the sender alive
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_scatter_compute_gather_world2():
    ctx = mp.get_context("spawn")
    for n_clips in (5, 2):
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
        for p in procs:
            p.start()
        out, out2 = (torch.tensor(v) for v in q.get())
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        g = torch.Generator().manual_seed(0)
        midi = torch.randint(0, 130, (n_clips, 1, 16), generator=g)
        cap = torch.randn(n_clips, 4, 8, generator=g)
        ref = midi.float().sum((1, 2))[:, None] + cap.sum((1, 2))[:, None] * torch.ones(1, 5)
        assert torch.allclose(out, ref)
        assert torch.allclose(out2, ref + 16.0)


def _host_worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from versband_b200.dist import HostExchange
    specs = [((1, 16), torch.int64), ((4, 8), torch.float32)] if rank == 0 else None
    ex = HostExchange(n_clips, specs, ((5,), torch.float32) if rank == 0 else None, device="cpu", spin_timeout_s=30)
    outs = []
    for step in range(3):
        if rank == 0:
            g = torch.Generator().manual_seed(step)
            ex.inputs[0].copy_(torch.randint(0, 130, (n_clips, 1, 16), generator=g))
            ex.inputs[1].copy_(torch.randn(n_clips, 4, 8, generator=g))
        out = ex.step(lambda loc: loc[0].float().sum((1, 2))[:, None] + loc[1].sum((1, 2))[:, None] * torch.ones(1, 5))
        if rank == 0:
            outs.append(out.clone().tolist())
    if rank == 0:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


def test_host_exchange_shared_memory_world2():
    """Shared-host-memory exchange (each rank moves its own shard; sequence counters instead of collectives)."""
    ctx = mp.get_context("spawn")
    for n_clips in (4, 3):
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_host_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
        for p in procs:
            p.start()
        outs = q.get()
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        for step, out in enumerate(outs):
            g = torch.Generator().manual_seed(step)
            midi = torch.randint(0, 130, (n_clips, 1, 16), generator=g)
            cap = torch.randn(n_clips, 4, 8, generator=g)
            ref = midi.float().sum((1, 2))[:, None] + cap.sum((1, 2))[:, None] * torch.ones(1, 5)
            assert torch.allclose(torch.tensor(out), ref)
