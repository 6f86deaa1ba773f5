"""Launches every kernel of the library once on a SMALL batch (for compute-sanitizer memcheck / racecheck and for NVTX-ranged
captures; not a benchmark):  python profiles/run_small.py [lj|egnn13|egnn55|ad2|lap|loop|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pita_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(12345)


def lj_net(n):
    from pita_b200.egnn_temp_conditioned import EGNN_dynamics
    return EGNN_dynamics(n_particles=n, n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(), recurrent=True, tanh=True,
                         attention=True, condition_time=True, condition_temperature=True, agg="sum").cuda()


def egnn(n, B):
    net = lj_net(n)
    w = net.packed_weights("cuda")
    x = ops.remove_mean(torch.randn(B, 3 * n, device="cuda") * 2.0, n)
    ht = torch.full((B,), 4.0, device="cuda")
    e, g, dh = ops.egnn_energy(w, 32, 3, n, ht, x, 0.75)
    for mode in ("bilinear", "fp32"):
        s, d = ops.egnn_score_div(w, 32, 3, n, ht, x, 0.75, mode=mode)
    v = ops.egnn_forward(w, 32, 3, n, ht, x, 0.75)
    return x, e, g, dh, s, d


if which in ("lj", "all"):
    for n, B in ((13, 200), (55, 70)):
        x = torch.randn(B, 3 * n, device="cuda") * 1.5
        ops.lj_energy_force(x, n)
        ops.lj_energy_force(x, n, need_force=False)
if which in ("egnn13", "all"):
    egnn(13, 45)
if which in ("egnn55", "all"):
    egnn(55, 5)
if which in ("ad2", "all"):
    from pita_b200.egnn_dynamics_ad2_cat import EGNN_dynamics_AD2_cat
    net = EGNN_dynamics_AD2_cat(22, 3, condition_beta=True).cuda()
    w = net.packed_weights("cuda")
    B = 3
    x = ops.remove_mean(torch.randn(B, 66, device="cuda"), 22)
    ht = torch.full((B,), 2.0, device="cuda")
    ops.egnn_forward(w, 64, 5, 22, ht, x, 0.9)
    ops.egnn_energy(w, 64, 5, 22, ht, x, 0.9)
    ops.egnn_score_div(w, 64, 5, 22, ht, x, 0.9)
if which == "lap13":   # one particle: the racecheck build of the 3n-pass kernel is slow
    net = lj_net(13)
    x = ops.remove_mean(torch.randn(1, 39, device="cuda") * 2.0, 13)
    ops.egnn_energy_laplacian(net.packed_weights("cuda"), 32, 3, 13, torch.full((1,), 3.0, device="cuda"), x, 0.8)
if which in ("lap", "all"):
    for n, B in ((13, 6), (55, 2)):
        net = lj_net(n)
        x = ops.remove_mean(torch.randn(B, 3 * n, device="cuda") * 2.0, n)
        ops.egnn_energy_laplacian(net.packed_weights("cuda"), 32, 3, n, torch.full((B,), 3.0, device="cuda"), x, 0.8)
if which in ("loop", "all"):
    n, B = 13, 300
    x, e, g, dh, s, d = egnn(n, B)
    xo, araw = ops.sde_fk_step(x, g, s, None, d, dh, e, n, g2=2.0, gamma=1.3, dgamma_dt=0.0, dh_dt=1.0, dt=1e-3, sqrt_dt=0.0316,
                               noise_scale=1.4, seed=3, offset=5)
    xo, araw = ops.sde_fk_step(x, g, s, torch.randn_like(x), d, dh, e, n, g2=2.0, gamma=1.3, dgamma_dt=0.0, dh_dt=1.0, dt=1e-3,
                               sqrt_dt=0.0316, noise_scale=1.4)
    a, _ = ops.fk_quantile_accumulate(araw, torch.zeros_like(araw), 128, 0.9, 1e-3, False)
    wts = ops.softmax_clip(a)
    ids, ch = ops.resample_systematic(wts, 0.3, count_changes=True)
    ops.gather_rows([xo.data_ptr()], B, ids, 3 * n)
torch.cuda.synchronize()
print("ok", which)
