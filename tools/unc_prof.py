"""Two calls of dd_uncertainty_project for ncu (no timing loop).
Usage: ncu --metrics gpu__time_duration.sum ... python tools/unc_prof.py <n_images> <n_vis>"""
import sys, torch, math
sys.path.insert(0, '.')
from dropoutdecoding_b200 import ops
n_img = int(sys.argv[1]); n_vis = int(sys.argv[2])
dev = torch.device('cuda:0')
h = torch.randn(n_img, n_vis, 4096, device=dev).bfloat16()
w = (torch.randn(32064, 4096, device=dev) * 0.02).bfloat16()
for _ in range(2):
    ops.uncertainty_project(h, w, topk=5, round_logits=True)
torch.cuda.synchronize()
