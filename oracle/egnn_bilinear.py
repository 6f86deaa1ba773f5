"""TEST INFRASTRUCTURE ONLY — the "bilinear" form of tr(d x_L / d y) for the 3-layer EGNN that the round-2 divergence
kernel (pita_b200/csrc/egnn_tri.cu) evaluates, restated on dense torch tensors without autograd.

Why a second form.  oracle/egnn_analytic.py::trace_dxL_dy pushes three forward-mode tangents per tangent node k through
the dense middle layer: 3 n E dense 32x32 products per particle.  Here the tangent is pushed forward only through layer 0
(sparse: edges touching k), the cotangent of the trace is pulled back through layer 2 (sparse: receiver k only), and the
two meet on the middle layer's edges as a BILINEAR form.  Both sides are rank one over the three directions for every node
other than k, so the pairing on a generic edge (i, j), i != k != j, needs ONE dense product per (edge, k) instead of three,
and every factor that multiplies it is a per-pair table:

    layer 0 (edge (i,k)):  dh1_i[a] = cf_a(i) omega_ik ,  dx1_i[a] = M_ik e_a ,  cf(i) = -2 (y_i - y_k)
    layer 2 (edge (k,i)):  cot on dh2_i[a] = w_a(ki) beta_ki ,  cot on dx2_i[a] = -X_ki e_a

Reference algebra being differentiated: egnn_temp_conditioned.py:56-93,265-356 (E_GCL), utils.py:30-51 (the trace).
Checked against trace_dxL_dy / vmap(jacrev) in tests/test_egnn_analytic.py.
"""
from __future__ import annotations

import torch

import egnn_analytic as A

Tensor = torch.Tensor


def edge_ops(w, lay, rng):
    """Per-edge linear maps of one layer on [B,i,j,*]:  T (tangent of m* w.r.t. the pre-activation of the first edge
    linear), its transpose, v = Wc1^T (wc2 * silu'(zc)), vt = T^T v, and the coordinate-branch geometry."""
    e = lay["e"]
    f1, f2, m, s = e["f1"], e["f2"], e["m"], e["s"]

    def T(z):  # z [...,B,i,j,H]
        dm = f2 * ((f1 * z) @ w["W2"].T)
        ds = s * (1 - s) * (dm * w["wa"]).sum(-1, keepdim=True)
        return dm * s + m * ds

    def Tt(g):
        gg = s * g + s * (1 - s) * (m * g).sum(-1, keepdim=True) * w["wa"]
        return f1 * ((f2 * gg) @ w["W2"])

    v = (w["wc2"] * e["fc"]) @ w["Wc1"]
    vt = Tt(v)
    cphi = rng * (1 - e["th"] ** 2)                    # [B,i,j,1]
    dhat = lay["dlt"] * lay["inv"]                    # [B,i,j,3]
    eye = torch.eye(3, dtype=v.dtype)
    N = lay["inv"][..., None] * eye - (lay["inv"] ** 2 / lay["nrm"])[..., None] * lay["dlt"][..., :, None] * lay["dlt"][..., None, :]
    return dict(T=T, Tt=Tt, v=v, vt=vt, cphi=cphi, dhat=dhat, N=N, phi=lay["phi"])


def pair_tables(sd, tcond, y, beta, n):
    """Everything that is a function of a PAIR of nodes (or of one node): the layer-0 tangent tables, the layer-2 cotangent
    tables and their node-level sums.  Index convention: tangent tables [B,i,k] come from layer-0 edge (i,k) (receiver i,
    sender k = tangent node); cotangent tables [B,k,j] from layer-2 edge (k,j) (receiver k = output node)."""
    B = y.shape[0]
    _, st = A.forward_states(sd, tcond, y, beta, n)
    assert st["L"] == 3
    rng, off = st["rng"], st["off"]
    lay0, lay1, lay2 = st["layers"]
    w0, w1, w2 = lay0["w"], lay1["w"], lay2["w"]
    o0, o1, o2 = edge_ops(w0, lay0, rng), edge_ops(w1, lay1, rng), edge_ops(w2, lay2, rng)
    eye = torch.eye(3, dtype=y.dtype)
    d0 = st["d0"]                                                       # [B,i,j,3]  y_i - y_j

    # ---- layer 0 -> tangent tables
    c01 = w0["c1"] + w0["d1"]
    wvec = o0["T"](c01.expand(B, n, n, -1))                             # [B,i,k,H]
    sigma = (o0["vt"] * c01).sum(-1, keepdim=True)                      # [B,i,k,1]
    cf = -2 * d0                                                        # [B,i,k,3]  cf(i) for tangent node k
    M = (-o0["phi"][..., None] * o0["N"]
         + (o0["cphi"] * sigma)[..., None] * o0["dhat"][..., :, None] * cf[..., None, :])            # [B,i,k,3(b),3(a)]
    omega = ((wvec @ w0["W3a"].T) * lay0["f3"][:, :, None, :]) @ w0["W4"].T                            # [B,i,k,H]
    # the tangent node itself (receiver k of layer 0, all its senders j)
    cfk = 2 * d0                                                        # [B,k,j,3]  cf(j) seen from tangent node k
    S = ((cfk * off)[..., :, None] * wvec[..., None, :]).sum(2)         # [B,k,3,H]
    Omega = ((S @ w0["W3a"].T) * lay0["f3"][:, :, None, :]) @ w0["W4"].T                                # [B,k,3(a),H]
    Dx1 = eye + ((o0["phi"][..., None] * o0["N"]
                  + (o0["cphi"] * sigma)[..., None] * o0["dhat"][..., :, None] * cfk[..., None, :]) * off[..., None]).sum(2)

    # ---- layer 2 -> cotangent tables
    alpha = o2["vt"] @ w2["A"]                                          # A^T vt
    beta_ = o2["vt"] @ w2["B"]
    rho = (o2["vt"] * w2["c1"]).sum(-1, keepdim=True)
    delta = (o2["vt"] * w2["d1"]).sum(-1, keepdim=True)
    wv = o2["cphi"] * o2["dhat"]                                        # [B,k,j,3]
    X = o2["phi"][..., None] * o2["N"] + 2 * rho[..., None] * lay2["dlt"][..., :, None] * wv[..., None, :]   # [B,k,j,3(b),3(a)]
    const = (delta[..., 0] * (wv * cfk).sum(-1) * off[..., 0]).sum(2)   # [B,k]
    Gx = eye + (X * off[..., None]).sum(2)                              # [B,k,3,3]
    Gam = ((wv * off)[..., :, None] * alpha[..., None, :]).sum(2)       # [B,k,3(a),H]

    def node_pullback(g, f3):  # cotangent g on h^2 of a node with silu'(z3) = f3 -> (cot on h^1, cot on agg)
        pb = (g @ w1["W4"]) * f3
        return g + pb @ w1["W3h"], pb @ w1["W3a"]

    betap, gamma = node_pullback(beta_, lay1["f3"][:, None, :, :])      # [B,k,j,H] (node j's f3)
    Gamp, Gamagg = node_pullback(Gam, lay1["f3"][:, :, None, :])        # [B,k,3,H]
    return dict(st=st, o1=o1, w1=w1, lay1=lay1, off=off, d0=d0, wvec=wvec, sigma=sigma, cf=cf, M=M, omega=omega, cfk=cfk,
                Omega=Omega, Dx1=Dx1, alpha=alpha, beta=beta_, rho=rho, delta=delta, wv=wv, X=X, const=const, Gx=Gx, Gam=Gam,
                betap=betap, gamma=gamma, Gamp=Gamp, Gamagg=Gamagg)


def trace_bilinear_dense(sd, tcond, y, beta, n):
    """tr(d x_L / d y) with the middle layer as a bilinear pairing of UNIFIED per-k node tables (no special-casing of the
    edges that touch k: the tables simply hold the full-rank entries at node k).  Dense, O(n^3 * 3 * H) memory: small n."""
    t = pair_tables(sd, tcond, y, beta, n)
    B = y.shape[0]
    o1, w1, lay1, off = t["o1"], t["w1"], t["lay1"], t["off"]
    H = t["omega"].shape[-1]
    ar = torch.arange(n)
    isk = torch.eye(n, dtype=y.dtype)                                   # [k,i]
    # unified tangent tables  [B,k,i,...]
    DH = (t["cf"].transpose(1, 2)[..., :, None] * t["omega"].transpose(1, 2)[..., None, :])      # [B,k,i,3,H]  cf_a(i) omega_ik
    DH[:, ar, ar] = t["Omega"]
    DX = t["M"].transpose(1, 2).clone()                                 # [B,k,i,3,3]
    DX[:, ar, ar] = t["Dx1"]
    # unified cotangent tables [B,k,i,...]
    GA = t["wv"][..., :, None] * t["gamma"][..., None, :]               # [B,k,i,3,H]
    GA[:, ar, ar] = t["Gamagg"]
    GX = -t["X"].clone()
    GX[:, ar, ar] = t["Gx"]
    GHp = t["wv"][..., :, None] * t["betap"][..., None, :]
    GHp[:, ar, ar] = t["Gamp"]
    # d edge_attr_ij[a] = [i == k] cf_a(j)|_k + [j == k] cf_a(i)|_k
    cfk = t["cfk"]                                                      # [B,k,j,3] = cf(j) seen from k
    dea = (isk[None, :, :, None, None] * cfk[:, :, None, :, :]          # i == k: cf(j)
           + isk[None, :, None, :, None] * cfk[:, :, :, None, :])       # j == k: cf(i)          -> [B,k,i,j,3]
    direct = (GX * DX).sum((-1, -2)).sum(2) + (GHp * DH).sum((-1, -2)).sum(2) + t["const"]       # [B,k]
    # layer-1 edges
    dlt, dhat, cphi, phi, N, v = lay1["dlt"], o1["dhat"], o1["cphi"], o1["phi"], o1["N"], o1["v"]
    DXd = DX[:, :, :, None] - DX[:, :, None, :]                         # [B,k,i,j,3(b),3(a)]
    g = 2 * (dlt[:, None, :, :, :, None] * DXd).sum(-2)                 # [B,k,i,j,3(a)]
    dz = (DH @ w1["A"].T)[:, :, :, None] + (DH @ w1["B"].T)[:, :, None, :] + g[..., None] * w1["c1"] + dea[..., None] * w1["d1"]  # [B,k,i,j,3,H]
    lam = cphi[:, None] * (GX[:, :, :, None, :, :] * dhat[:, None, :, :, :, None]).sum(-2)        # [B,k,i,j,3(a)]
    cotin = GA[:, :, :, None] + lam[..., None] * v[:, None, :, :, None, :]                         # [B,k,i,j,3,H]
    # apply T^T per edge: move the (k, a) axes in front so that the per-edge closures broadcast
    cot = o1["Tt"](cotin.permute(1, 4, 0, 2, 3, 5)).permute(2, 0, 3, 4, 1, 5)
    offk = off[:, None, :, :, :]                                        # [B,1,i,j,1]
    E1 = ((cot * dz).sum(-1) * offk).sum((2, 3, 4))
    E2 = (phi[:, None] * (GX[:, :, :, None] * (N[:, None] @ DXd)).sum((-1, -2))[..., None] * offk).sum((2, 3, 4))
    return (direct + E1 + E2).sum(1)


# ---------------------------------------------------------------------------------------------------------------------
# The same trace in the shape the CUDA kernels evaluate it (pita_b200/csrc/egnn_tri.cu):
#   phase A (thread = (particle, node) row): primal forward + the per-pair tables below + the "direct" part of the trace;
#   phase B (thread = middle-layer edge (i, j), loop over the tangent node k): one dense product per (edge, k).
# `rnd` optionally rounds the operand rows of phase B's dense product (emulation of the TF32 operand rounding).
# ---------------------------------------------------------------------------------------------------------------------
def phase_a_tables(sd, tcond, y, beta, n):
    t = pair_tables(sd, tcond, y, beta, n)
    B = y.shape[0]
    w1 = t["w1"]
    ar = torch.arange(n)
    off2 = t["off"][..., 0]                                             # [B,i,j]
    # sender-side table  TS[k][j]:  PB_jk = B1 omega_jk,  M'_jk  (diagonal: Dx1_k; PB diagonal: 0)
    PB = (t["omega"] @ w1["B"].T).transpose(1, 2).clone()               # [B,k,j,H]
    PB[:, ar, ar] = 0
    Ms = t["M"].transpose(1, 2).clone()                                 # [B,k,j,3,3]
    Ms[:, ar, ar] = t["Dx1"]
    # receiver-side table TR[i][k]
    PA = (t["omega"] @ w1["A"].T).clone()                               # [B,i,k,H]
    PA[:, ar, ar] = 0
    gam = t["gamma"].transpose(1, 2).clone()                            # [B,i,k,H]  gamma_ki
    gam[:, ar, ar] = 0
    wv = t["wv"].transpose(1, 2).clone()                                # [B,i,k,3]  w(ki)
    wv[:, ar, ar] = 0
    alpha_i = (wv * t["cf"]).sum(-1)                                    # [B,i,k]    w(ki) . cf(i)|_k
    GXs = -t["X"].transpose(1, 2).clone()                               # [B,i,k,3,3]
    GXs[:, ar, ar] = t["Gx"]
    Mr = t["M"].clone()                                                 # [B,i,k,3,3]
    Mr[:, ar, ar] = t["Dx1"]
    # node tables
    AOm = t["Omega"] @ w1["A"].T                                        # [B,k,3,H]
    BOm = t["Omega"] @ w1["B"].T
    # direct part of the trace (node terms + pair terms)
    pair = (-(t["X"].transpose(1, 2) * t["M"]).sum((-1, -2)) + alpha_i * (t["betap"].transpose(1, 2) * t["omega"]).sum(-1)) * off2
    # per output node k (the layout phase A writes): node terms + the pair terms of its layer-2 edges (k, i)
    direct_node = (t["const"] + (t["Gx"] * t["Dx1"]).sum((-1, -2)) + (t["Gamp"] * t["Omega"]).sum((-1, -2))) + pair.sum(1)
    direct = direct_node.sum(1)
    return dict(t=t, PB=PB, Ms=Ms, PA=PA, gam=gam, wv=wv, alpha_i=alpha_i, GXs=GXs, Mr=Mr, AOm=AOm, BOm=BOm,
                GAgg=t["Gamagg"], direct=direct, direct_node=direct_node)


def phase_b_items(tab, n, rnd=None, per_receiver=False):
    """Sum over the middle layer's edges: generic items (every k != i, the row with j == k included through zeroed table
    diagonals), the S item (row (i,j) at k = j: the full-rank part of the sender's tangent) and the three R items (k = i)."""
    t = tab["t"]
    o1, w1, lay1 = t["o1"], t["w1"], t["lay1"]
    e = lay1["e"]
    f1, f2, m, s, vt = e["f1"], e["f2"], e["m"], e["s"], o1["vt"]       # [B,i,j,*]
    cphi, phi, dhat, dlt = o1["cphi"][..., 0], o1["phi"][..., 0], o1["dhat"], lay1["dlt"]
    inv, nrm = lay1["inv"][..., 0], lay1["nrm"][..., 0]
    duc = (vt * w1["c1"]).sum(-1)                                       # [B,i,j]
    off2 = t["off"][..., 0]
    rnd = rnd or (lambda z: z)
    B = f1.shape[0]
    ar = torch.arange(n)

    def product(u_in, cot, g1):
        """<T^T cot, u_in> evaluated the kernel's way.  u_in, cot: [B,i,j,(k),H];  g1 = <m, cot>."""
        D = rnd(f1e * u_in) @ w1["W2"].T
        tt = f2e * D
        return se * (tt * cot).sum(-1) + se * (1 - se) * g1 * (tt * w1["wa"]).sum(-1)

    # ---- generic items: axes [B,i,j,k]
    f1e, f2e, se = f1[:, :, :, None, :], f2[:, :, :, None, :], s[:, :, :, None, 0]
    PA = tab["PA"][:, :, None, :, :]                                    # [B,i,1,k,H]
    PBk = tab["PB"].permute(0, 2, 1, 3)[:, None]                        # [B,1,j,k,H]  PB_jk
    gam = tab["gam"][:, :, None, :, :]
    wv = tab["wv"][:, :, None, :, :]                                    # [B,i,1,k,3]
    al_i = tab["alpha_i"][:, :, None, :]
    GXs = tab["GXs"][:, :, None]                                        # [B,i,1,k,3,3]
    Mi = tab["Mr"][:, :, None]
    Mj = tab["Ms"].permute(0, 2, 1, 3, 4)[:, None]                      # [B,1,j,k,3,3]  M'_jk
    cf_i = t["cf"][:, :, None, :, :]                                    # [B,i,1,k,3]  -2 (y_i - y_k)
    cf_j = t["cf"][:, None, :, :, :]                                    # [B,1,j,k,3]
    dl, dh = dlt[:, :, :, None, :], dhat[:, :, :, None, :]
    xis = (GXs * dh[..., :, None]).sum(-2)                              # [B,i,j,k,3]  GXs^T dhat
    dM = Mi - Mj
    g = 2 * (dM * dl[..., :, None]).sum(-2)                             # [B,i,j,k,3]
    al_j = (wv * cf_j).sum(-1)
    bet = (wv * g).sum(-1)
    u_in = al_i[..., None] * PA + al_j[..., None] * PBk + bet[..., None] * w1["c1"]
    g1 = (m[:, :, :, None, :] * gam).sum(-1)
    termA = product(u_in, gam, g1)
    vte = vt[:, :, :, None, :]
    termB = cphi[..., None] * ((xis * cf_i).sum(-1) * (vte * PA).sum(-1) + (xis * cf_j).sum(-1) * (vte * PBk).sum(-1)
                               + (xis * g).sum(-1) * duc[..., None])
    E2 = phi[..., None] * (inv[..., None] * (GXs * dM).sum((-1, -2)) - (inv / nrm)[..., None] * 0.5 * (xis * g).sum(-1))
    notk = 1 - torch.eye(n, dtype=f1.dtype)                             # [i,k]
    gen = ((termA + termB + E2) * off2[..., None] * notk[None, :, None, :]).sum((2, 3))

    # ---- S item: row (i,j), k = j:  u_in = sum_a w_a(ji) B Omega_j[a] + alpha_i d1 ;  cot gamma_ji
    f1e, f2e, se = f1, f2, s[..., 0]
    w_s = tab["wv"]                                                     # [B,i,k=j,3]
    BOm = tab["BOm"][:, None]                                           # [B,1,j,3,H]
    gam_s = tab["gam"]                                                  # [B,i,j,H]   gamma_ji
    al_s = tab["alpha_i"]                                               # [B,i,j]
    u_s = (w_s[..., None] * BOm).sum(-2) + al_s[..., None] * w1["d1"]
    g1_s = (m * gam_s).sum(-1)
    sA = product(u_s, gam_s, g1_s)
    xis_s = (tab["GXs"] * dhat[..., :, None]).sum(-2)                   # [B,i,j,3] with k = j
    cf_is = t["cf"]                                                     # [B,i,k=j,3]
    sB = cphi * ((vt[..., None, :] * BOm).sum(-1) * xis_s).sum(-1) + cphi * (xis_s * cf_is).sum(-1) * (vt * w1["d1"]).sum(-1)
    sit = ((sA + sB) * off2).sum(2)

    # ---- R items: row (i,j), k = i, direction a:  tan_a = A Omega_i[a] + cf_a(j)|_i (PB_ji + d1) + g_a c1 ; cot GAgg_i[a]
    GX_r = tab["GXs"][:, ar, ar][:, :, None]                            # [B,i,1,3,3]  Gx_i
    M_i = tab["Mr"][:, ar, ar][:, :, None]                              # Dx1_i
    M_j = tab["Ms"].permute(0, 2, 1, 3, 4)                              # [B,j,k,..] -> need [B,i,j] = M'_{j, k=i}
    M_j = M_j.permute(0, 2, 1, 3, 4)                                    # [B,k=i,j,3,3]
    dM_r = M_i - M_j
    g_r = 2 * (dM_r * dlt[..., :, None]).sum(-2)                        # [B,i,j,3(a)]
    xis_r = (GX_r * dhat[..., :, None]).sum(-2)                         # [B,i,j,3(a)]
    cfj_r = 2 * t["d0"]                                                 # [B,i,j,3]   -2 (y_j - y_i)
    PB_r = tab["PB"]                                                    # [B,k=i,j,H]  PB_ji
    rit = 0
    for a in range(3):
        tan = tab["AOm"][:, :, None, a, :] + cfj_r[..., a:a + 1] * (PB_r + w1["d1"]) + g_r[..., a:a + 1] * w1["c1"]
        cot = tab["GAgg"][:, :, None, a, :].expand_as(tan)
        rA = product(tan, cot, (m * cot).sum(-1))
        rB = cphi * xis_r[..., a] * (vt * tan).sum(-1)
        rit = rit + ((rA + rB) * off2).sum(2)
    E2_r = phi * (inv * (GX_r * dM_r).sum((-1, -2)) - (inv / nrm) * 0.5 * (xis_r * g_r).sum(-1))
    rit = rit + (E2_r * off2).sum(2)
    tot = gen + sit + rit                                              # [B,i]: per receiver
    return tot if per_receiver else tot.sum(1)


def trace_bilinear_kernel_form(sd, tcond, y, beta, n, rnd=None):
    tab = phase_a_tables(sd, tcond, y, beta, n)
    return tab["direct"] + phase_b_items(tab, n, rnd)
