"""World-size-2 gloo tests (CPU) of the N>1 path: prompt sharding and CFG sharding of the denoising loop must return,
on every rank, exactly the latents of the single-process loop (SURVEY §8e determinism check).  The model and the
scheduler step are stand-in deterministic callables: what is under test is the host-side partitioning, the per-step pair
exchange and the single final all-gather of `parallel.sharded_denoise`."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(lat_in, pe, ref, t):
    """Per-sequence deterministic stand-in for the transformer: no cross-sequence interaction, like the real one."""
    B = lat_in.shape[0]
    refb = ref if ref.shape[0] == B else torch.cat([ref, ref], dim=0)       # eval=True doubling (cogvideox_transformer_3d.py:503)
    return torch.tanh(lat_in * 0.9 + pe.mean(dim=(1, 2)).view(B, 1, 1, 1, 1) + 0.1 * refb.mean(dim=(1, 2, 3, 4)).view(B, 1, 1, 1, 1)
                      + t * 1e-3)


def _step(noise2, t, lat, g):
    P = lat.shape[0]
    u, c = noise2[:P], noise2[P:]
    return 0.97 * lat - 0.05 * (u + g * (c - u))


def _inputs(P):
    gen = torch.Generator().manual_seed(7)
    lat = torch.randn(P, 2, 4, 6, 6, generator=gen)
    pe = torch.randn(2 * P, 5, 8, generator=gen)
    ref = torch.randn(P, 1, 4, 6, 6, generator=gen)
    return lat, pe, ref


def _single(P, steps):
    lat, pe, ref = _inputs(P)
    for i, t in enumerate(steps):
        lat = _step(_model(torch.cat([lat, lat]), pe, ref, t), t, lat, 6.0 + i)
    return lat


def _worker(rank, world, port, P, steps, q):
    import s2v_b200
    from s2v_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = parallel.plan(P, world, rank)
        lat, pe, ref = _inputs(P)
        out = parallel.sharded_denoise(sp, P, lat, pe, ref, steps, _model, _step, lambda i: 6.0 + i)
        q.put((rank, sp.mode, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,mode", [(2, "prompt"), (4, "prompt"), (1, "cfg")])
def test_sharded_loop_matches_single_process(P, mode):
    steps = [999, 979, 959]
    want = _single(P, steps)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, m, out in got:
        assert m == mode
        assert out.shape == want.shape
        assert torch.equal(out, want), f"rank {rank}: sharded latents differ from the single-process loop"
