"""Alanine-dipeptide kernels alone (for ncu):  python profiles/run_ad2.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pita_b200 import ops
from pita_b200.egnn_dynamics_ad2_cat import EGNN_dynamics_AD2_cat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
torch.manual_seed(12345)
net = EGNN_dynamics_AD2_cat(22, 3, condition_beta=True).cuda()
w = net.packed_weights("cuda")
x = ops.remove_mean(torch.randn(B, 66, device="cuda") * 1.5, 22)
ht = torch.full((B,), 2.0, device="cuda")
for _ in range(2):
    ops.egnn_energy(w, 64, 5, 22, ht, x, 0.9)
    ops.egnn_score_div(w, 64, 5, 22, ht, x, 0.9)
torch.cuda.synchronize()
print("ok")
