"""Launches each hot-path kernel a few times at a moderate size (for ncu captures; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pita_b200 import ops
from pita_b200.egnn_temp_conditioned import EGNN_dynamics, pack_state_dict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
B = int(sys.argv[2]) if len(sys.argv) > 2 else (148 * 16 if n == 13 else 148 * 2)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(12345)
net = EGNN_dynamics(n_particles=n, n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(), recurrent=True, tanh=True,
                    attention=True, condition_time=True, condition_temperature=True, agg="sum")
w = pack_state_dict(net.state_dict(), 32, 3, "cuda")
x = ops.remove_mean(torch.randn(B, 3 * n, device="cuda") * 3.0, n)
ht = torch.full((B,), 9.0, device="cuda")
for _ in range(reps):
    e, g, dh = ops.egnn_energy(w, 32, 3, n, ht, x, 0.75)
    s, d = ops.egnn_score_div(w, 32, 3, n, ht, x, 0.75)
    lp, f = ops.lj_energy_force(x, n)
    xo, araw = ops.sde_fk_step(x, g, s, torch.randn_like(x), d, dh, e, n, g2=2.0, gamma=1.3, dgamma_dt=0.0, dh_dt=1.0, dt=1e-3,
                               sqrt_dt=0.0316, noise_scale=1.4)
    a, _ = ops.fk_quantile_accumulate(araw, torch.zeros_like(araw), 512, 0.9, 1e-3, False)
    wts = ops.softmax_clip(a)
    ids, ch = ops.resample_systematic(wts, 0.3)
    xr = ops.gather_rows([xo.data_ptr()], B, ids, 3 * n)
torch.cuda.synchronize()
print("ok", n, B)
