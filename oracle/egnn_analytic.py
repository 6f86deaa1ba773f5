"""TEST INFRASTRUCTURE ONLY — autograd-free restatement of the derivative algebra the CUDA kernels use.

`oracle/pita_oracle.py` obtains grad_x U, dU/dt and div(score) the way the reference does (autograd,
vmap(jacrev)).  The CUDA kernels in pita_b200/csrc/egnn.cu cannot: they run a hand-derived reverse
pass (energy net) and a hand-derived forward-mode tangent pass (score-net divergence).  This file
states that algebra once, on dense tensors with plain torch ops and NO autograd, so the derivation can be
checked against the autograd oracle on the CPU (tests/test_egnn_analytic.py) independently of any
CUDA-specific bug.  Structure facts it relies on (each asserted by a test):
  * the last layer's node update never reaches the velocity output;
  * vel is translation invariant, so  tr(d remove_mean(vel)/dy) == tr(d x_L/dy) - 3n;
  * for the trace only the diagonal blocks d x_L[k] / d y[k] are needed, so the last layer is evaluated
    only on edges (k, j) for the tangent node k, and layer 0 only on edges touching k.
Reference algebra being differentiated: egnn_temp_conditioned.py:56-93,265-356; energy_net.py:14-62;
score_net.py:13-43.
"""
from __future__ import annotations

from typing import Dict

import torch

Tensor = torch.Tensor


def _sig(v):
    return torch.sigmoid(v)


def unpack(sd: Dict[str, Tensor], l: int):
    pre = f"egnn.gcl_{l}."
    W1 = sd[pre + "edge_mlp.0.weight"]
    H = W1.shape[0]
    return dict(
        A=W1[:, :H], B=W1[:, H:2 * H], c1=W1[:, 2 * H], d1=W1[:, 2 * H + 1], b1=sd[pre + "edge_mlp.0.bias"],
        W2=sd[pre + "edge_mlp.2.weight"], b2=sd[pre + "edge_mlp.2.bias"],
        wa=sd[pre + "att_mlp.0.weight"][0], ba=sd[pre + "att_mlp.0.bias"][0],
        Wc1=sd[pre + "coord_mlp.0.weight"], bc1=sd[pre + "coord_mlp.0.bias"], wc2=sd[pre + "coord_mlp.2.weight"][0],
        W3h=sd[pre + "node_mlp.0.weight"][:, :H], W3a=sd[pre + "node_mlp.0.weight"][:, H:],
        b3=sd[pre + "node_mlp.0.bias"], W4=sd[pre + "node_mlp.2.weight"], b4=sd[pre + "node_mlp.2.bias"])


def node_features(tcond: Tensor, beta: Tensor, n: int):
    """[B,n,2] features and d(features)/d(tcond) under the reference's cat/reshape quirk."""
    B = tcond.shape[0]
    flat = torch.cat([tcond[:, None].expand(B, n), beta[:, None].expand(B, n)], dim=-1)
    feat = flat.reshape(B, n, 2)
    is_t = torch.cat([torch.ones(n), torch.zeros(n)]).reshape(n, 2).to(tcond.dtype)  # 1 where the slot holds t
    return feat, is_t


def _edge_primal(w, p, q, r2, ea):
    """p:[B,n,1,H] q:[B,1,n,H] r2,ea:[B,n,n,1] -> dict of primal edge quantities on [B,i,j,*]."""
    z1 = p + q + w["c1"] * r2 + w["d1"] * ea
    s1 = _sig(z1)
    a1 = z1 * s1
    z2 = a1 @ w["W2"].T + w["b2"]
    s2 = _sig(z2)
    m = z2 * s2
    s = _sig((m * w["wa"]).sum(-1, keepdim=True) + w["ba"])
    ms = m * s
    zc = ms @ w["Wc1"].T + w["bc1"]
    sc = _sig(zc)
    ac = zc * sc
    u = (ac * w["wc2"]).sum(-1, keepdim=True)
    th = torch.tanh(u)
    return dict(z1=z1, f1=s1 * (1 + z1 * (1 - s1)), z2=z2, f2=s2 * (1 + z2 * (1 - s2)), m=m, s=s, ms=ms,
                zc=zc, fc=sc * (1 + zc * (1 - sc)), th=th)


def forward_states(sd, tcond, y, beta, n, coords_range=15.0):
    """Primal forward keeping what the derivative passes re-use.  Returns (vel, states)."""
    B = y.shape[0]
    L = 0
    while f"egnn.gcl_{L}.edge_mlp.0.weight" in sd:
        L += 1
    rng = coords_range / L
    x0 = y.reshape(B, n, 3)
    feat, is_t = node_features(tcond, beta, n)
    h = feat @ sd["egnn.embedding.weight"].T + sd["egnn.embedding.bias"]
    off = (~torch.eye(n, dtype=torch.bool)).to(y.dtype)[None, :, :, None]
    d0 = x0[:, :, None, :] - x0[:, None, :, :]
    ea = d0.pow(2).sum(-1, keepdim=True)
    st = dict(L=L, rng=rng, n=n, off=off, d0=d0, ea=ea, is_t=is_t, layers=[])
    x = x0
    for l in range(L):
        w = unpack(sd, l)
        dlt = x[:, :, None, :] - x[:, None, :, :]
        r2 = dlt.pow(2).sum(-1, keepdim=True)
        nrm = torch.sqrt(r2 + 1e-8)
        inv = 1 / (nrm + 1)
        p = (h @ w["A"].T + w["b1"])[:, :, None, :]
        q = (h @ w["B"].T)[:, None, :, :]
        e = _edge_primal(w, p, q, r2, ea)
        phi = e["th"] * rng
        xn = x + (dlt * inv * phi * off).sum(2)
        lay = dict(w=w, x=x, h=h, dlt=dlt, r2=r2, nrm=nrm, inv=inv, p=p, q=q, e=e, phi=phi)
        if l < L - 1:  # the last node update is dead code for vel
            agg = (e["ms"] * off).sum(2)
            z3 = h @ w["W3h"].T + agg @ w["W3a"].T + w["b3"]
            s3 = _sig(z3)
            lay["f3"] = s3 * (1 + z3 * (1 - s3))
            h = h + (z3 * s3) @ w["W4"].T + w["b4"]
        st["layers"].append(lay)
        x = xn
    st["xL"] = x
    vel = x - x0
    vel = vel - vel.mean(1, keepdim=True)
    return vel.reshape(B, 3 * n), st


def u_theta_backward(sd, tcond, y, beta, n):
    """U = <vel(y), y>;  returns U, dU/dy [B,3n], dU/dtcond [B]  by a hand-written reverse pass."""
    B = y.shape[0]
    vel, st = forward_states(sd, tcond, y, beta, n)
    L, off = st["L"], st["off"]
    yv = y.reshape(B, n, 3)
    U = (vel * y).sum(1)
    wv = yv - yv.mean(1, keepdim=True)  # cotangent on x_L (P y)
    gx = wv.clone()  # \bar x^{l+1}
    H = sd["egnn.embedding.weight"].shape[0]
    gh = torch.zeros(B, n, H, dtype=y.dtype)
    gea = torch.zeros(B, n, n, 1, dtype=y.dtype)
    for l in reversed(range(L)):
        lay = st["layers"][l]
        w, e = lay["w"], lay["e"]
        if l < L - 1:
            gz3 = lay["f3"] * (gh @ w["W4"])
            gh_in = gh + gz3 @ w["W3h"]
            gagg = gz3 @ w["W3a"]
        else:
            gh_in = gh.clone()
            gagg = torch.zeros_like(gh)
        dhat = lay["dlt"] * lay["inv"]
        gphi = (gx[:, :, None, :] * dhat).sum(-1, keepdim=True)
        gu = gphi * st["rng"] * (1 - e["th"] ** 2)
        gzc = gu * w["wc2"] * e["fc"]
        gms = gagg[:, :, None, :] + gzc @ w["Wc1"]
        gs = (gms * e["m"]).sum(-1, keepdim=True)
        gm = gms * e["s"] + w["wa"] * (gs * e["s"] * (1 - e["s"]))
        gz2 = gm * e["f2"]
        gz1 = (gz2 @ w["W2"]) * e["f1"] * off
        gp = gz1.sum(2)  # over senders j  -> receiver i
        gq = gz1.sum(1)  # over receivers i -> sender j
        gr2 = (gz1 * w["c1"]).sum(-1, keepdim=True)
        gea = gea + (gz1 * w["d1"]).sum(-1, keepdim=True)
        gdhat = gx[:, :, None, :] * lay["phi"]
        gdl = gdhat * lay["inv"] - lay["dlt"] * ((gdhat * lay["dlt"]).sum(-1, keepdim=True) * lay["inv"] ** 2 / lay["nrm"])
        gdl = (gdl + 2 * lay["dlt"] * gr2) * off
        gx = gx + gdl.sum(2) - gdl.sum(1)
        gh = gh_in + gp @ w["A"] + gq @ w["B"]
    g0 = 2 * st["d0"] * gea * off
    gx = gx + g0.sum(2) - g0.sum(1)
    dU_dy = vel.reshape(B, n, 3) - wv + gx
    dfeat = st["is_t"]  # [n,2]
    dh0_dt = dfeat @ sd["egnn.embedding.weight"].T  # [n,H]
    dU_dt = (gh * dh0_dt[None]).sum((1, 2))
    return U, dU_dy.reshape(B, 3 * n), dU_dt


def energy_terms(sd, ht, x, beta, n):
    """EnergyNet: E, grad_x E, dE/dh  (energy_net.py:14-62), autograd-free."""
    B = x.shape[0]
    beta = beta * torch.ones(B, dtype=x.dtype)
    c_in = (1 + ht) ** -0.5
    c_noise = 0.125 * torch.log(ht)
    y = c_in[:, None] * x
    U, dU_dy, dU_dc = u_theta_backward(sd, c_noise, y, beta, n)
    x2 = (x * x).sum(1)
    E = x2 / (2 * (1 + ht)) - ht ** -0.5 * U
    gE = x / (1 + ht)[:, None] - (ht ** -0.5 * c_in)[:, None] * dU_dy
    dU_dh = dU_dc / (8 * ht) + (dU_dy * x).sum(1) * (-0.5) * (1 + ht) ** -1.5
    dE_dh = -x2 / (2 * (1 + ht) ** 2) + 0.5 * ht ** -1.5 * U - ht ** -0.5 * dU_dh
    return E, gE, dE_dh


def _edge_tangent(w, lay, rng, dp, dq, Dd, dea):
    """Tangent of one layer's edge function on [B,i,j] for a stack of tangents (leading dim T).
    dp:[T,B,n,1,H] dq:[T,B,1,n,H] Dd:[T,B,n,n,3] dea:[T,B,n,n,1] -> (dms [T,B,n,n,H], dtrans [T,B,n,n,3])."""
    e = lay["e"]
    dr2 = 2 * (lay["dlt"] * Dd).sum(-1, keepdim=True)
    dz1 = dp + dq + w["c1"] * dr2 + w["d1"] * dea
    da1 = e["f1"] * dz1
    dm = e["f2"] * (da1 @ w["W2"].T)
    ds = e["s"] * (1 - e["s"]) * (dm * w["wa"]).sum(-1, keepdim=True)
    dms = dm * e["s"] + e["m"] * ds
    dac = e["fc"] * (dms @ w["Wc1"].T)
    du = (dac * w["wc2"]).sum(-1, keepdim=True)
    dphi = rng * (1 - e["th"] ** 2) * du
    ddhat = Dd * lay["inv"] - lay["dlt"] * ((lay["dlt"] * Dd).sum(-1, keepdim=True) / lay["nrm"] * lay["inv"] ** 2)
    dtrans = ddhat * lay["phi"] + lay["dlt"] * lay["inv"] * dphi
    return dms, dtrans


def edge_cache(w, lay):
    """What the score/divergence kernel keeps per middle-layer edge (csrc/egnn_rows.cu, "layer-1 edge cache"): nothing here
    depends on the tangent node, so it is computed once per tile.  v = Wc1^T (wc2 * silu'(zc)) turns the coordinate-branch
    tangent  du = <wc2 * silu'(zc), Wc1 dms>  into the dot product  <v, dms>."""
    e = lay["e"]
    return dict(f1=e["f1"], m=e["m"], f2=e["f2"], att=e["s"], th=e["th"], v=(w["wc2"] * e["fc"]) @ w["Wc1"])


def _edge_tangent_cached(w, lay, cache, rng, dp, dq, Dd, dea):
    """_edge_tangent in the form the kernel evaluates it: ONE dense product per tangent row (W2), the coordinate tangent by
    a dot product with the cached v; returns dms (to be summed over senders BEFORE W3a is applied) and dtrans."""
    dr2 = 2 * (lay["dlt"] * Dd).sum(-1, keepdim=True)
    dz1 = cache["f1"] * (dp + dq + w["c1"] * dr2 + w["d1"] * dea)
    dm = cache["f2"] * (dz1 @ w["W2"].T)
    ds = cache["att"] * (1 - cache["att"]) * (dm * w["wa"]).sum(-1, keepdim=True)
    dms = dm * cache["att"] + cache["m"] * ds
    du = (cache["v"] * dms).sum(-1, keepdim=True)
    dphi = rng * (1 - cache["th"] ** 2) * du
    ddhat = Dd * lay["inv"] - lay["dlt"] * ((lay["dlt"] * Dd).sum(-1, keepdim=True) / lay["nrm"] * lay["inv"] ** 2)
    return dms, ddhat * lay["phi"] + lay["dlt"] * lay["inv"] * dphi


def trace_dxL_dy(sd, tcond, y, beta, n, cached: bool = False):
    """tr(d x_L / d y) per sample by forward-mode tangents, one tangent node k at a time, using the
    sparsity the kernels use (layer 0: edges touching k; last layer: receiver k only).  cached=True evaluates the dense
    middle layer through the per-edge cache exactly as csrc/egnn_rows.cu does."""
    B = y.shape[0]
    _, st = forward_states(sd, tcond, y, beta, n)
    L, off, rng = st["L"], st["off"], st["rng"]
    H = sd["egnn.embedding.weight"].shape[0]
    tr = torch.zeros(B, dtype=y.dtype)
    eye3 = torch.eye(3, dtype=y.dtype)
    for k in range(n):
        # tangent stack T=3: d/dy[k,a]
        dx = torch.zeros(3, B, n, 3, dtype=y.dtype)
        dx[:, :, k, :] = eye3[:, None, :]
        dh = torch.zeros(3, B, n, H, dtype=y.dtype)
        Dd0 = dx[:, :, :, None, :] - dx[:, :, None, :, :]
        dea = 2 * (st["d0"] * Dd0).sum(-1, keepdim=True)  # d edge_attr, non-zero only on edges touching k
        for l in range(L):
            lay = st["layers"][l]
            w = lay["w"]
            Dd = dx[:, :, :, None, :] - dx[:, :, None, :, :]
            dp = (dh @ w["A"].T)[:, :, :, None, :]
            dq = (dh @ w["B"].T)[:, :, None, :, :]
            if l == L - 1:  # only receiver k is needed
                sub = {kk: (vv[:, k:k + 1] if torch.is_tensor(vv) and vv.dim() == 4 and vv.shape[1] == n and vv.shape[2] == n else vv)
                       for kk, vv in lay.items() if kk not in ("e", "w")}
                sub["e"] = {kk: vv[:, k:k + 1] for kk, vv in lay["e"].items()}
                dms, dtr = _edge_tangent(w, sub, rng, dp[:, :, k:k + 1], dq, Dd[:, :, k:k + 1], dea[:, :, k:k + 1])
                dxk = dx[:, :, k, :] + (dtr * off[:, k:k + 1]).sum(3)[:, :, 0, :]  # [3,B,3]
                tr = tr + dxk.diagonal(dim1=0, dim2=2).sum(-1)
                break
            if cached and 0 < l < L - 1:  # the kernel's form of the dense middle layer(s)
                dms, dtr = _edge_tangent_cached(w, lay, edge_cache(w, lay), rng, dp, dq, Dd, dea)
            else:
                dms, dtr = _edge_tangent(w, lay, rng, dp, dq, Dd, dea)
            dagg = (dms * off).sum(3)  # summed over the sender slots first, W3a once (linearity)
            dz3 = dh @ w["W3h"].T + dagg @ w["W3a"].T
            dh = dh + (lay["f3"] * dz3) @ w["W4"].T
            dx = dx + (dtr * off).sum(3)
    return tr


def trace_dxL_dy_ranked(sd, tcond, y, beta, n):
    """Same trace for the 3-layer network with the middle layer evaluated through the RANK STRUCTURE of its inputs
    (DESIGN.md §8 item 1 — the algebra of the next kernel revision, checked here before any CUDA is written):

      after layer 0, for every node i != k the three direction tangents are rank one,  dh1_i[a] = cf_a(i) * omega_i,
      cf_a(i) = -2 (y_i - y_k)_a,  omega_i = W4 (f3_i * W3a wvec_ik);  so on an edge (i, j) with i != k and j != k

          dms_a = cf_a(i) T(f1 * A omega_i) + cf_a(j) T(f1 * B omega_j) + dr2_a T(f1 * c1),

      where T is the edge's linear tangent map (W2, silu', attention gate).  T(f1 * c1) does not depend on k (it belongs in
      the edge cache), and cf_a(i) factors out of the sum over senders, so a pass needs TWO dense products per edge and one
      accumulator for the A-part instead of three products.  Edges with i == k or j == k (own-direction vectors, edge_attr
      tangent) are evaluated directly, as the kernel will do in its per-k sections."""
    B = y.shape[0]
    _, st = forward_states(sd, tcond, y, beta, n)
    L, off, rng = st["L"], st["off"], st["rng"]
    assert L == 3
    H = sd["egnn.embedding.weight"].shape[0]
    tr = torch.zeros(B, dtype=y.dtype)
    eye3 = torch.eye(3, dtype=y.dtype)
    lay0, lay1, lay2 = st["layers"]
    w0, w1, w2 = lay0["w"], lay1["w"], lay2["w"]
    e0, e1 = lay0["e"], lay1["e"]

    def tmap(e, w, vec):  # T: R^H -> dms on [.., B, i, j, H]
        dm = e["f2"] * (vec @ w["W2"].T)
        ds = e["s"] * (1 - e["s"]) * (dm * w["wa"]).sum(-1, keepdim=True)
        return dm * e["s"] + e["m"] * ds

    cache1 = edge_cache(w1, lay1)
    Tc = tmap(e1, w1, e1["f1"] * w1["c1"])          # k-independent, cacheable
    duc = (cache1["v"] * Tc).sum(-1, keepdim=True)
    # layer 0: wvec_ij = T0(f1 * (c1 + d1)) for every edge, omega_ij = W4 (f3_i * W3a wvec_ij)
    wvec = tmap(e0, w0, e0["f1"] * (w0["c1"] + w0["d1"]))                                   # [B,i,j,H]
    omega = (lay0["f3"][:, :, None, :] * (wvec @ w0["W3a"].T)) @ w0["W4"].T                  # [B,i,j,H]  (j plays k)
    for k in range(n):
        # ---- layer 0 exactly as trace_dxL_dy does (kept dense here; it is not the subject of this function)
        dx = torch.zeros(3, B, n, 3, dtype=y.dtype)
        dx[:, :, k, :] = eye3[:, None, :]
        dh = torch.zeros(3, B, n, H, dtype=y.dtype)
        Dd0 = dx[:, :, :, None, :] - dx[:, :, None, :, :]
        dea = 2 * (st["d0"] * Dd0).sum(-1, keepdim=True)
        dms0, dtr0 = _edge_tangent(w0, lay0, rng, (dh @ w0["A"].T)[:, :, :, None, :], (dh @ w0["B"].T)[:, :, None, :, :], Dd0, dea)
        dh = dh + (lay0["f3"] * ((dms0 * off).sum(3) @ w0["W3a"].T)) @ w0["W4"].T
        dx = dx + (dtr0 * off).sum(3)
        # rank-one claim for i != k
        cf = -2 * st["d0"][:, :, k, :].permute(2, 0, 1)                                     # [3,B,n]  cf_a(i)
        om = omega[:, :, k, :]                                                              # [B,n,H]  omega_ik
        notk = torch.ones(n, dtype=torch.bool)
        notk[k] = False
        assert torch.allclose(dh[:, :, notk], cf[:, :, notk, None] * om[None, :, notk], rtol=1e-9, atol=1e-12)
        # ---- layer 1 through the rank structure
        PA, PB = om @ w1["A"].T, om @ w1["B"].T                                             # [B,n,H]
        TP = tmap(e1, w1, e1["f1"] * PA[:, :, None, :])                                     # [B,i,j,H]
        TQ = tmap(e1, w1, e1["f1"] * PB[:, None, :, :])
        Dd = dx[:, :, :, None, :] - dx[:, :, None, :, :]
        dr2 = 2 * (lay1["dlt"] * Dd).sum(-1, keepdim=True)                                  # [3,B,i,j,1]
        generic = (notk[:, None] & notk[None, :]).to(y.dtype)[None, None, :, :, None] * off
        dms = (cf[:, :, :, None, None] * TP + cf[:, :, None, :, None] * TQ + dr2 * Tc) * generic
        du = (cf[:, :, :, None, None] * (cache1["v"] * TP).sum(-1, keepdim=True)
              + cf[:, :, None, :, None] * (cache1["v"] * TQ).sum(-1, keepdim=True) + dr2 * duc) * generic
        # special edges (i == k or j == k): direct evaluation with the own-direction vectors and the edge_attr tangent
        dp = (dh @ w1["A"].T)[:, :, :, None, :]
        dq = (dh @ w1["B"].T)[:, :, None, :, :]
        dms_s = tmap(e1, w1, e1["f1"] * (dp + dq + w1["c1"] * dr2 + w1["d1"] * dea))
        du_s = (cache1["v"] * dms_s).sum(-1, keepdim=True)
        special = (1 - (notk[:, None] & notk[None, :]).to(y.dtype))[None, None, :, :, None] * off
        dms = dms + dms_s * special
        du = du + du_s * special
        dphi = rng * (1 - e1["th"] ** 2) * du
        ddhat = Dd * lay1["inv"] - lay1["dlt"] * ((lay1["dlt"] * Dd).sum(-1, keepdim=True) / lay1["nrm"] * lay1["inv"] ** 2)
        dtr1 = ddhat * lay1["phi"] + lay1["dlt"] * lay1["inv"] * dphi
        # the aggregate in the factored form: cf_a(i) * sum_j T_P  +  sum_j (cf_a(j) T_Q + dr2_a T_c)  (+ special edges)
        g2 = generic[0, :, :, :, :]
        dagg = (cf[:, :, :, None] * (TP * g2).sum(2)[None] + ((cf[:, :, None, :, None] * TQ + dr2 * Tc) * generic).sum(3)
                + (dms_s * special).sum(3))
        assert torch.allclose(dagg, dms.sum(3), rtol=1e-9, atol=1e-12)
        dz3 = dh @ w1["W3h"].T + dagg @ w1["W3a"].T
        dh = dh + (lay1["f3"] * dz3) @ w1["W4"].T
        dx = dx + (dtr1 * off).sum(3)
        # ---- layer 2: receiver k only (as trace_dxL_dy)
        Dd = dx[:, :, :, None, :] - dx[:, :, None, :, :]
        dp = (dh @ w2["A"].T)[:, :, :, None, :]
        dq = (dh @ w2["B"].T)[:, :, None, :, :]
        sub = {kk: (vv[:, k:k + 1] if torch.is_tensor(vv) and vv.dim() == 4 and vv.shape[1] == n and vv.shape[2] == n else vv)
               for kk, vv in lay2.items() if kk not in ("e", "w")}
        sub["e"] = {kk: vv[:, k:k + 1] for kk, vv in lay2["e"].items()}
        _, dtr2 = _edge_tangent(w2, sub, rng, dp[:, :, k:k + 1], dq, Dd[:, :, k:k + 1], dea[:, :, k:k + 1])
        dxk = dx[:, :, k, :] + (dtr2 * off[:, k:k + 1]).sum(3)[:, :, 0, :]
        tr = tr + dxk.diagonal(dim1=0, dim2=2).sum(-1)
    return tr


def score_and_divergence(sd, ht, x, beta, n):
    """ScoreNet.forward and tr(d score/dx)  (score_net.py:13-43; utils.py:43-51), autograd-free."""
    B = x.shape[0]
    beta = beta * torch.ones(B, dtype=x.dtype)
    c_s = 1 / (1 + ht)
    c_in = (1 + ht) ** -0.5
    c_out = ht ** 0.5 * c_in
    c_noise = 0.125 * torch.log(ht)
    y = c_in[:, None] * x
    vel, _ = forward_states(sd, c_noise, y, beta, n)
    score = ((c_s - 1)[:, None] * x + c_out[:, None] * vel) / ht[:, None]
    D = 3 * n
    div = ((c_s - 1) * D + c_out * c_in * (trace_dxL_dy(sd, c_noise, y, beta, n) - D)) / ht
    return score, div
