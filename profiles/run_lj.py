"""LJ-55 energy+force kernel alone (for ncu):  PITA_LJ_CFG=<k> python profiles/run_lj.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pita_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
x = torch.randn(B, 165, device="cuda") * 1.5
for _ in range(3):
    ops.lj_energy_force(x, 55)
torch.cuda.synchronize()
print("ok")
