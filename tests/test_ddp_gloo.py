"""world_size-2 gloo (CPU) test of the data-parallel host logic: the flat-gradient sink fires one all-reduce per
bucket as the reverse sweep fills it, `finish()` returns the 1/world scale, and both ranks end with the average of
their two DIFFERENT random gradients.  The kernels themselves need a GPU; this covers the exchange
plumbing.  On hardware the same statement is `PairTrainer.exchange_check` (bench.py prints it as `dp_check` at N > 1)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, deferred, ret):
    sys.path.insert(0, ROOT)
    import vst_b200  # noqa: F401
    from vst_b200.reconet.network import ReCoNetSD2
    from vst_b200.train_core import FlatParams, GradSink

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    flat = FlatParams(ReCoNetSD2(1))
    sink = GradSink(flat, dist.group.WORLD, n_buckets=3)
    sink.defer = deferred
    fired = []

    def grad_of(r, n):                                   # a different random gradient per rank and tensor, reproducible anywhere
        g = torch.Generator().manual_seed(1000 * r + flat.names.index(n))
        return torch.randn(flat.offsets[n][2], generator=g)

    for n in reversed(flat.names):                       # the reverse sweep produces the last layer first
        sink.put(n, grad_of(rank, n))
        fired.append(len(sink.works))
    if deferred:
        assert fired[-1] == 0
        sink.exchange_all()
        scale = 1.0 / sink.world
    else:
        assert fired[-1] == 3 and fired[0] == 0 and sorted(fired) == fired   # buckets fire progressively
        scale = sink.finish()
    g = flat.grad * scale
    want = torch.cat([((grad_of(0, n) + grad_of(1, n)) * 0.5).reshape(-1) for n in flat.names])
    ok = abs(scale - 0.5) < 1e-12 and bool(torch.allclose(g, want, rtol=1e-6, atol=1e-7))
    ret[rank] = ok
    dist.destroy_process_group()


def _run(deferred, port):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, deferred, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_bucketed_allreduce_overlapped():
    _run(False, 29731)


def test_deferred_allreduce_after_graph_replay():
    _run(True, 29733)
