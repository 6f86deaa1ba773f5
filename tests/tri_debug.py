"""TEST INFRASTRUCTURE — table-by-table check of the bilinear score/divergence engine (csrc/egnn_tri_*.cu) on a GPU against
oracle/egnn_bilinear.py (fp64).  Run as a script on the B200 box:  python tests/tri_debug.py [n] [B]
Prints, per workspace table, the worst error relative to the table's largest entry, then the divergence itself."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import egnn_analytic as A  # noqa: E402
import egnn_bilinear as BL  # noqa: E402
import pita_oracle as O  # noqa: E402
from helpers import golden, make_net, state_from_golden  # noqa: E402


def layout(n):
    from pita_b200 import _native as N
    out = (ctypes.c_int64 * 16)()
    cnt = N.load().pita_egnn_tri_workspace_layout(n, out, 16)
    names = ["kFloats", "oTS", "oTR", "oOM", "oY", "oX1", "oP1", "oQ1", "oOmg", "oAOm", "oBOm", "oGAgg", "oDirect", "oPartB",
             "kTS", "kTR"]
    assert cnt == len(names)
    return dict(zip(names, list(out)))


def rel(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if not np.isfinite(got).all():
        return float("nan")
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def run(n, B, fixture=None, verbose=True):
    from pita_b200 import ops
    g = golden(fixture or ("fk_n%d_strong.npz" % n))
    sd = state_from_golden(g, "S.")
    x0 = torch.from_numpy(g["x"]).double()
    reps = (B + x0.shape[0] - 1) // x0.shape[0]
    gen = torch.Generator().manual_seed(7)
    x = torch.cat([x0] * reps)[:B] * (1 + 0.02 * torch.randn(B, 1, generator=gen, dtype=torch.float64))
    x = O.centre(x, n)
    sched = O.EDMSchedule(float(g["sigma_min"]))
    t = torch.linspace(0.35, 0.65, B, dtype=torch.float64)
    ht = sched.h(t)
    beta = torch.full((B,), float(g["beta"]), dtype=torch.float64)
    c_in = (1 + ht) ** -0.5
    c_noise = 0.125 * torch.log(ht)
    y = c_in[:, None] * x
    tab = BL.phase_a_tables(sd, c_noise, y, beta, n)
    pb = BL.phase_b_items(tab, n, per_receiver=True)
    s_ref, d_ref = A.score_and_divergence(sd, ht, x, beta[0], n)

    net = make_net(n, sd)
    dev = "cuda"
    sc, dv = ops.egnn_score_div(net.packed_weights(dev), 32, 3, n, ht.float().to(dev), x.float().to(dev), beta.float().to(dev),
                                need_div=True, mode="bilinear")
    torch.cuda.synchronize()
    ws = list(ops._div_ws.values())[-1].view(torch.float32).cpu().numpy().astype(np.float64)
    Lo = layout(n)
    kF, kTS, kTR = Lo["kFloats"], Lo["kTS"], Lo["kTR"]
    t0 = tab["t"]
    lay1 = t0["lay1"]
    res = {}

    def rec(name, got, ref, mask=None):
        got, ref = np.asarray(got), ref.numpy() if torch.is_tensor(ref) else np.asarray(ref)
        if mask is not None:
            got, ref = got[mask], ref[mask]
        res.setdefault(name, 0.0)
        e = rel(got, ref)
        res[name] = e if np.isnan(e) else max(res[name], e)

    offd = ~np.eye(n, dtype=bool)
    for b in range(B):
        w = ws[b * kF:(b + 1) * kF]
        TS = w[Lo["oTS"]:Lo["oTS"] + n * n * kTS].reshape(n, n, kTS)
        TR = w[Lo["oTR"]:Lo["oTR"] + n * n * kTR].reshape(n, n, kTR)
        OM = w[Lo["oOM"]:Lo["oOM"] + n * n * 32].reshape(n, n, 32)
        rec("Y", w[Lo["oY"]:Lo["oY"] + 4 * n].reshape(n, 4)[:, :3], y[b].reshape(n, 3))
        rec("X1", w[Lo["oX1"]:Lo["oX1"] + 4 * n].reshape(n, 4)[:, :3], lay1["x"][b])
        rec("P1", w[Lo["oP1"]:Lo["oP1"] + 32 * n].reshape(n, 32), lay1["p"][b, :, 0])
        rec("Q1", w[Lo["oQ1"]:Lo["oQ1"] + 32 * n].reshape(n, 32), lay1["q"][b, 0])
        rec("Omega", w[Lo["oOmg"]:Lo["oOmg"] + 96 * n].reshape(n, 3, 32), t0["Omega"][b])
        rec("AOm", w[Lo["oAOm"]:Lo["oAOm"] + 96 * n].reshape(n, 3, 32), tab["AOm"][b])
        rec("BOm", w[Lo["oBOm"]:Lo["oBOm"] + 96 * n].reshape(n, 3, 32), tab["BOm"][b])
        rec("GAgg", w[Lo["oGAgg"]:Lo["oGAgg"] + 96 * n].reshape(n, 3, 32), tab["GAgg"][b])
        rec("omega(OM)", OM, t0["omega"][b], offd)
        rec("TS.PB", TS[:, :, :32], tab["PB"][b])
        rec("TS.M", TS[:, :, 32:41].reshape(n, n, 3, 3), tab["Ms"][b])
        rec("TR.PA", TR[:, :, :32], tab["PA"][b])
        rec("TR.gamma", TR[:, :, 32:64], tab["gam"][b])
        rec("TR.w", TR[:, :, 64:67], tab["wv"][b])
        rec("TR.alpha", TR[:, :, 67], tab["alpha_i"][b])
        rec("TR.GXs", TR[:, :, 68:77].reshape(n, n, 3, 3), tab["GXs"][b])
        rec("TR.M", TR[:, :, 77:86].reshape(n, n, 3, 3), tab["Mr"][b])
        rec("direct(node)", w[Lo["oDirect"]:Lo["oDirect"] + n], tab["direct_node"][b])
        npart = 32
        part = w[Lo["oPartB"]:Lo["oPartB"] + npart]
        # phase-B partials: one per (CTA tile, team) = a group of consecutive receivers
        rpt = 2 if n == 55 else 8
        ngrp = (n + 2 * rpt - 1) // (2 * rpt)
        ref_part = np.zeros(ngrp * 2)
        for i in range(n):
            ref_part[i // rpt] += float(pb[b, i])
        rec("phaseB partials", part[:ngrp * 2], ref_part)
        if verbose and b == 0:
            print("  partials gpu", np.array2string(part[:ngrp * 2], precision=5, max_line_width=200))
            print("  partials ref", np.array2string(ref_part, precision=5, max_line_width=200))
    res["score"] = rel(sc.cpu().numpy(), s_ref.numpy())
    d_gpu = dv.cpu().double().numpy()
    res["div (|d|/max(|ref|,1))"] = float((np.abs(d_gpu - d_ref.numpy()) / np.maximum(np.abs(d_ref.numpy()), 1.0)).max())
    if verbose:
        for k, v in res.items():
            print("%-24s %.3e" % (k, v))
        print("div gpu", d_gpu[:4], "ref", d_ref.numpy()[:4])
    return res


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (20 if n == 13 else 5)
    run(n, B)
