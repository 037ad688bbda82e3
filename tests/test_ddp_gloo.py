"""N>1 host logic on CPU: world_size-2 gloo run of the gradient bucket reducer and scan sharding."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _paths  # noqa: F401


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import _paths  # noqa: F401
    from gcdlss_b200.ddp import GradBucketReducer, shard_scans
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                       # ranks construct DIFFERENT models: the reducer must broadcast rank 0's
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 3))
    reducer = GradBucketReducer(net.parameters(), bucket_bytes=256)      # several small buckets
    assert len(reducer.buckets) > 1
    torch.manual_seed(0)
    rank0_net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 3))
    assert all(torch.equal(a, b) for a, b in zip(net.parameters(), rank0_net.parameters())), "parameters were not broadcast from rank 0"
    gen = torch.Generator().manual_seed(100)
    data = torch.randn(8, 8, generator=gen)
    mine = shard_scans(8, rank, world)
    for step in range(2):
        reducer.reset()
        net(data[mine]).pow(2).sum().backward()
        reducer.finish()
    grads = torch.cat([p.grad.flatten() for p in net.parameters()])
    # single-process reference: mean over ranks of per-shard gradients
    ref_net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 3))
    ref_net.load_state_dict(net.state_dict())
    acc = None
    for r in range(world):
        ref_net.zero_grad()
        ref_net(data[shard_scans(8, r, world)]).pow(2).sum().backward()
        g = torch.cat([p.grad.flatten() for p in ref_net.parameters()])
        acc = g if acc is None else acc + g
    ok = torch.allclose(grads, acc / world, rtol=1e-5, atol=1e-6)
    # zero_grad(set_to_none=True) detaches .grad from the buckets: finish() must say so instead of reducing stale zeros
    net.zero_grad(set_to_none=True)
    net(data[mine]).pow(2).sum().backward()
    try:
        reducer.check_views()
        ok = False
    except RuntimeError:
        pass
    if rank == 0:
        open(out, "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_bucket_reducer_world2_gloo(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_scans_partition():
    from gcdlss_b200.ddp import shard_scans
    for world in (1, 2, 4, 8):
        allidx = sum((shard_scans(16, r, world) for r in range(world)), [])
        assert sorted(allidx) == list(range(16))
    assert [len(shard_scans(10, r, 4)) for r in range(4)] == [3, 3, 2, 2]                      # no trailing scan is dropped
    assert sorted(sum((shard_scans(10, r, 4) for r in range(4)), [])) == list(range(10))


def test_sink_bookkeeping_ignores_the_hooks_of_sunk_parameters():
    """Gradient sink (functional.TrunkFunction.backward -> GradBucketReducer.mark_ready): autograd still runs the empty
    accumulation of a parameter whose gradient went straight into the bucket and fires its post-accumulate hook; counting that
    hook a second time used to send a bucket out before the parameters that really come later (the stem) had arrived
    (found on two GPUs by tools/ddp_check.py).  Single process, the launches recorded instead of issued."""
    from gcdlss_b200.ddp import GradBucketReducer
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (8, 8, 8, 8, 8, 8)]
    red = GradBucketReducer(params, bucket_bytes=3 * 8 * 4)          # two buckets of three parameters, reverse order
    assert [len(pl) for _, pl in red.buckets] == [3, 3]
    launched = []
    red.world = 2
    red._launch = lambda b, gate=None: launched.append((b, gate)) or len(launched)
    red.reset()
    late, sunk = params[0], params[1:]                                # params[0] (registered first) gets its gradient last, by autograd
    assert red.grad_ptr(sunk[0]) is None or red.grad_ptr(sunk[0]) % 16 == 0
    red.mark_ready(sunk[2:], "event A")                               # bucket 0 = params 5, 4, 3: complete -> goes out behind event A
    assert launched == [(0, "event A")]
    red.mark_ready(sunk[:2], "event B")                               # bucket 1 = params 2, 1, 0: two of three
    assert launched == [(0, "event A")]
    for p in sunk:                                                    # autograd's hooks for the sunk parameters: no effect
        red._on_grad_ready(p)
    assert launched == [(0, "event A")] and red._pending == [0, 1]
    red._on_grad_ready(late)                                          # the real late gradient completes bucket 1, ungated
    assert launched == [(0, "event A"), (1, None)]
    red.reset()
    assert not red._sunk and red._pending == [3, 3]
