"""Times the default score/divergence engine on LJ-55 (CUDA events, 4 launch pairs):  python profiles/time_scorediv.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pita_b200 import ops
from pita_b200.egnn_temp_conditioned import EGNN_dynamics

n, B = 55, 2368
torch.manual_seed(12345)
net = EGNN_dynamics(n_particles=n, n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(), recurrent=True, tanh=True,
                    attention=True, condition_time=True, condition_temperature=True, agg="sum").cuda()
w = net.packed_weights("cuda")
x = ops.remove_mean(torch.randn(B, 3 * n, device="cuda") * 3.0, n)
ht = torch.full((B,), 9.0, device="cuda")
ops.egnn_score_div(w, 32, 3, n, ht, x, 0.75)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.egnn_score_div(w, 32, 3, n, ht, x, 0.75); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("score+div %d particles: %.3f ms -> %.3f ms per launch pair of 592" % (B, sorted(ts)[1], sorted(ts)[1] / 4))
