"""One native classifier gradient call (B = 8, full-size classifier) bracketed by cudaProfilerStart/Stop:
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv ... python tools/profile_classifier.py"""
import os
import sys
sys.path.insert(0, ".")
os.environ["DFB_NO_CLF_GRAPH"] = "1"
import torch
from diff_foley_b200.classifier import AlignmentClassifierDoubleGuidanceB200
from diff_foley_b200.weights import randomize_parameters_
dev = torch.device("cuda", 0)
clf = AlignmentClassifierDoubleGuidanceB200().to(dev)
randomize_parameters_(clf, seed=9)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, 16, 64, generator=g).to(dev)
f = torch.nn.functional.normalize(torch.randn(B, 32, 512, generator=g), dim=-1).to(dev)
t = torch.full((B,), 500, dtype=torch.long, device=dev)
for _ in range(2):
    clf.loglikelihood_grad(x, t, f, 50.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
clf.loglikelihood_grad(x, t, f, 50.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
