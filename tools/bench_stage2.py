"""Times the Stage-2 mean-teacher step harness (BASELINE config 4 on one rank): 2 labelled + 2 unlabelled KITTI-like
scans (80 k points each, as the reference's training down-sampling), teacher fwd + student fwd + LaserMix + GPU quantise
+ student fwd #2 + backward + SGD + EMA."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import torch
import gcdlss_b200, MinkowskiEngine as ME
from gcdlss_b200.steps import Stage2Harness
from models.multiheadminkunet import MinkUNetRC
sys.path.insert(0, os.path.join(_paths.ROOT, "tests"))
from test_gpu_stage2 import _half

gcdlss_b200.set_math_mode(os.environ.get("MODE", "bf16"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
student = MinkUNetRC(17).cuda().train()
for name, n_out in (("final2", 3), ("final3", 2)):
    setattr(student.encoder, name, ME.MinkowskiConvolution(96, n_out, kernel_size=1, bias=True, dimension=3).cuda())
teacher = copy.deepcopy(student)
opt = torch.optim.SGD(student.parameters(), lr=0.01, momentum=0.9, fused=True)
h = Stage2Harness(student, teacher, opt, voxel_size=0.05)
sup = _half("kitti", 0, 2, 80000, dev, True)
unsup = _half("kitti", 10, 2, 80000, dev, False)
for _ in range(5): h.step(sup, unsup)
torch.cuda.synchronize(); t = time.perf_counter()
n = 20
for _ in range(n): loss = h.step(sup, unsup)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n
print(f"stage-2 step ({os.environ.get('MODE','bf16')}): {dt*1e3:.2f} ms/step, {4/dt:.1f} scans/s (4 scans + 4 mixed scans per step), loss {float(loss):.4f}")
