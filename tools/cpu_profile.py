"""cProfile of the host side of the training step (which is launch-bound) on the GPU box."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import numpy as np, torch
import gcdlss_b200, MinkowskiEngine as ME
from gcdlss_b200 import synth
from gcdlss_b200.ddp import GradBucketReducer
from models.multiheadminkunet import MinkUNetBase
import bench

gcdlss_b200.set_math_mode(os.environ.get("MODE", "bf16"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MinkUNetBase(num_classes=17).to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused=True)
reducer = GradBucketReducer(model.parameters())
hb = bench.make_host_batches("kitti", 4, None, 17, 0, 2)
res = [bench.quantize_batch_on_gpu(b, 0.05, dev) for b in hb]

def step(i):
    bc, f, l = res[i % 2]
    st = ME.SparseTensor(features=f, coordinates=bc)
    out = model(st)
    loss = torch.nn.functional.cross_entropy(out["logits"], l)
    reducer.reset(); loss.backward(); reducer.finish(); opt.step()

for i in range(5): step(i)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for i in range(10): step(i)
t_cpu = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f"10 steps: host issue time {t_cpu*100:.2f} ms/step, wall {t_all*100:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for i in range(10): step(i)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue()[:6000])
