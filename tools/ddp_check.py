"""Multi-GPU check of the data-parallel gradient path (run under torchrun, N >= 2 GPUs of one node):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py
One training step's reduced gradients with the gradient sink (kernels accumulate into the all-reduce buckets, buckets go out
behind per-stage events while backward continues) against the same step through autograd's accumulation and hooks, and against
the mean of the ranks' local gradients gathered by hand."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import torch
import torch.distributed as dist

import bench
import gcdlss_b200
import MinkowskiEngine as ME
from gcdlss_b200.ddp import GradBucketReducer
from gcdlss_b200.steps import point_cross_entropy
from models.multiheadminkunet import MinkUNetBase

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
gcdlss_b200.set_math_mode("bf16")
torch.manual_seed(1234 + rank)             # different initial weights per rank: the reducer must broadcast rank 0's
model = MinkUNetBase(num_classes=17).to(dev).train()
reducer = GradBucketReducer(model.parameters())
reducer.broadcast_buffers(model)
hb = bench.make_host_batches("kitti", 2, 40000, 17, rank, 1)
bc, f, labels = bench.quantize_batch_on_gpu(hb[0], 0.05, dev)


def step(direct):
    reducer.direct = direct
    reducer.reset()
    out = model(ME.SparseTensor(features=f, coordinates=bc))
    point_cross_entropy(out["logits"], labels).backward()
    reducer.finish()
    torch.cuda.synchronize()
    return torch.cat([p.grad.flatten().clone() for p in model.parameters()])


g_hooks = step(False)
reducer.trace = []
g_sink = step(True)
if rank == 0:
    tr = reducer.trace
    print("trace of the sink step (rank 0): main stream", torch.cuda.current_stream().cuda_stream, flush=True)
    for t in tr:
        if t[0] == "launch" or (t[0] == "hook" and t[1] in (0, len(reducer.buckets) - 1)):
            print("   ", t, flush=True)
reducer.trace = None
g_sink2 = step(True)
# local gradient of this rank without any reduction (through autograd's accumulation, then through the sink), averaged by hand
saved_world = reducer.world


def local(direct):
    reducer.direct = direct
    reducer.world = 1
    reducer.reset()
    out = model(ME.SparseTensor(features=f, coordinates=bc))
    point_cross_entropy(out["logits"], labels).backward()
    torch.cuda.synchronize()
    reducer.world = saved_world
    return torch.cat([p.grad.flatten().clone() for p in model.parameters()])


local_g = local(False)
local_sink = local(True)
print(f"rank {rank}: local gradients, sink vs autograd {float((local_sink - local_g).norm() / local_g.norm()):.2e}", flush=True)
names = [n for n, _ in model.named_parameters()]
off = 0
worst = []
for n, p in model.named_parameters():
    a, b = g_sink[off:off + p.numel()], g_hooks[off:off + p.numel()]
    worst.append((float((a - b).norm() / b.norm().clamp_min(1e-30)), n))
    off += p.numel()
worst.sort(reverse=True)
print(f"rank {rank}: parameters whose reduced gradient differs most (sink vs hooks): {worst[:6]}; {sum(1 for w in worst if w[0] > 1e-2)} of {len(worst)} above 1e-2", flush=True)
dist.all_reduce(local_g)
local_g /= world


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


same_across_ranks = g_sink.clone()
dist.broadcast(same_across_ranks, src=0)
ok = rel(g_sink, g_hooks) < 2e-3 and rel(g_sink, local_g) < 2e-3 and rel(g_sink2, g_sink) < 2e-3 and torch.equal(same_across_ranks, g_sink)
print(f"rank {rank}: sink vs hooks {rel(g_sink, g_hooks):.2e}, sink vs hand-averaged {rel(g_sink, local_g):.2e}, repeat {rel(g_sink2, g_sink):.2e}, "
      f"identical on all ranks {torch.equal(same_across_ranks, g_sink)} -> {'OK' if ok else 'FAILED'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
