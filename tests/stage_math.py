"""Stage-level restatement of the fused maxdim-2 LGAE path -- TEST INFRASTRUCTURE ONLY.

The CUDA kernels in ``lgn_autoencoder_b200/csrc`` do not evaluate the reference's CG products
literally; they use the closed form of SURVEY.md appendix A.2b, rewritten as neighbour sums
``M^l_i[c] = sum_j R^l_ij[c] X^l_j[c]`` of per-node quantities, and a hand-derived adjoint.  This
module states that algebra (forward AND hand-written backward) in plain torch with complex128
tensors so that

  * the algebra itself is checked on the CPU against ``oracle/lgae_oracle.py`` (which follows the
    reference literally) and against torch autograd, and
  * each CUDA stage can be compared with its torch twin on the GPU box to localise a bug.

Conventions: S (B,N,C) complex scalars, V (B,N,C,4) complex canonical 4-vectors, y (B,N,4) the
canonical components of the particle momenta.  Gradients of a real loss wrt a complex z = x+iy are
carried as g = dL/dx + i dL/dy, so for w = a*z: g_z = conj(a) g_w.
"""
from __future__ import annotations

import math

import torch

C128 = torch.complex128
R2 = 1.0 / math.sqrt(2.0)


# ---------------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------------
def planar_to_c(x):
    """(2, ...) planar -> complex."""
    return torch.complex(x[0], x[1])


def c_to_planar(z):
    return torch.stack([z.real, z.imag], 0)


def eta(a, b):
    """Complex bilinear Minkowski form in the canonical basis."""
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 3] - a[..., 2] * b[..., 2] + a[..., 3] * b[..., 1]


def ghat(a):
    """(g a)_mu with g the canonical-basis metric: eta(a,b) = sum_mu ghat(a)_mu b_mu."""
    return torch.stack([a[..., 0], a[..., 3], -a[..., 2], a[..., 1]], -1)


def canon_real(p):
    """Real Cartesian (...,4) -> canonical complex (...,4) (zonal_functions.py:266-288)."""
    t, x, y, z = p.unbind(-1)
    zero = torch.zeros_like(t)
    return torch.stack([torch.complex(t, zero), torch.complex(R2 * x, -R2 * y), torch.complex(z, zero),
                        torch.complex(-R2 * x, -R2 * y)], -1)


def canon_cplx(p):
    """Complex Cartesian (...,4) -> canonical (p_cplx_to_rep, zonal_functions.py:292-341)."""
    t, x, y, z = p.unbind(-1)
    return torch.stack([t, R2 * (x - 1j * y), z, -R2 * (x + 1j * y)], -1)


def canon_cplx_bwd(g):
    """Adjoint of canon_cplx: g_p = M^H g."""
    g0, g1, g2, g3 = g.unbind(-1)
    return torch.stack([g0, R2 * (g1 - g3), 1j * R2 * (g1 + g3), g2], -1)


def cart_from_canon(v):
    """rep_to_p: canonical -> complex Cartesian (zonal_functions.py:344-381) = M^H v."""
    return canon_cplx_bwd(v)


def cart_from_canon_bwd(g):
    """Adjoint of rep_to_p = M g."""
    return canon_cplx(g)


def minkowski_sq_real(p):
    psq = p * p
    return 2 * psq[..., 0] - (((psq[..., 0] + psq[..., 1]) + psq[..., 2]) + psq[..., 3])


# ---------------------------------------------------------------------------------------------
# parameters of one level, extracted from a reference-style state dict
# ---------------------------------------------------------------------------------------------
def level_params(sd, lvl, encoder: bool):
    pre = f"rad_funcs.rad_funcs.{lvl}"
    mixp = f"lgn_cg.node_levels.{lvl}.cat_mix.mix_reps.weights."
    out = dict(
        a=sd[pre + ".a"].reshape(-1), b=sd[pre + ".b"].reshape(-1), c=sd[pre + ".c"].reshape(-1),
        w0=sd[pre + ".linear.0.weight"], b0=sd[pre + ".linear.0.bias"],
        w1=sd[pre + ".linear.1.weight"], b1=sd[pre + ".linear.1.bias"],
        m00=planar_to_c(sd[mixp + "(0, 0)"]), m11=planar_to_c(sd[mixp + "(1, 1)"]),
    )
    i = 0
    lin = []
    while f"lgn_cg.mlp_levels.{lvl}.linear.{i}.weight" in sd:
        lin.append((sd[f"lgn_cg.mlp_levels.{lvl}.linear.{i}.weight"], sd[f"lgn_cg.mlp_levels.{lvl}.linear.{i}.bias"]))
        i += 1
    out["mlp"] = lin
    return out


# ---------------------------------------------------------------------------------------------
# radial functions
# ---------------------------------------------------------------------------------------------
def radial_enc(norms, mask, lp):
    """Encoder radial weights R^0, R^1 (B,N,N,C) complex (position_levels.py:136-176)."""
    x = norms.unsqueeze(-1)
    d = 1.0 + (lp["c"] * x) ** 2 + 1e-16
    phi = torch.where(mask.unsqueeze(-1), lp["b"] / d + lp["a"], torch.zeros((), dtype=norms.dtype))
    r0 = phi @ lp["w0"].T + lp["b0"]
    r1 = phi @ lp["w1"].T + lp["b1"]
    return torch.complex(r0[..., 0::2], r0[..., 1::2]), torch.complex(r1[..., 0::2], r1[..., 1::2]), phi, d


def radial_dec(lp):
    """Decoder: all-zero mask => R^l[c] = bias_l[c] (1+i)."""
    return lp["b0"] * (1 + 1j), lp["b1"] * (1 + 1j)


# ---------------------------------------------------------------------------------------------
# one LGN level, forward
# ---------------------------------------------------------------------------------------------
def pair_diff(y):
    """Y_ij = y_i - y_j, (B,N,N,4)."""
    return y.unsqueeze(2) - y.unsqueeze(1)


def _bc(R, like):
    """Radial weights as (B|1, N|1, N|1, C): edge dependent (encoder) or constant (decoder)."""
    return R if R.dim() == 4 else R.view(1, 1, 1, -1)


def neighbour_sums(R0, R1, S, V, Y):
    """The four neighbour sums of one level, each over ALL j:
         A0V_i[c] = sum_j R0_ij[c] V_j[c]            A0S_i[c] = sum_j R0_ij[c] S_j[c]
         A1Y_i[c] = sum_j R1_ij[c] S_j[c] Y_ij        A1E_i[c] = sum_j R1_ij[c] eta(V_j[c], Y_ij)
    The differences Y_ij are formed first, as the reference does (zonal_functions.py:221-248): the
    Minkowski products of nearly collinear, nearly massless momenta cancel heavily and splitting
    Y_ij into y_i and y_j terms would lose ~2 digits there."""
    R0, R1 = _bc(R0, S), _bc(R1, S)
    Sj, Vj = S.unsqueeze(1), V.unsqueeze(1)                    # (B,1,N,C[,4])
    n = S.shape[1]
    A0V = (R0.unsqueeze(-1) * Vj).sum(2).expand(-1, n, -1, -1)
    A0S = (R0 * Sj).sum(2).expand(-1, n, -1)
    A1Y = ((R1 * Sj).unsqueeze(-1) * Y.unsqueeze(3)).sum(2)
    A1E = (R1 * eta(Vj, Y.unsqueeze(3))).sum(2)
    return A0V, A0S, A1Y, A1E


def assemble_cat(S, V, As):
    A0V, A0S, A1Y, A1E = As
    ag11a = (1 + 1j) * A0V
    ag11b = A1Y
    ag00a = 0.5 * A1E
    ag00b = (1 + 1j) * A0S
    sq11 = V * S.unsqueeze(-1)
    sq00a = 0.5 * eta(V, V)
    sq00b = S * S
    cat11 = torch.cat([ag11a, ag11b, V, sq11, sq11], 2)      # (B,N,5C,4)
    cat00 = torch.cat([ag00a, ag00b, S, sq00a, sq00b], 2)    # (B,N,5C)
    return cat00, cat11


def level_forward(S, V, y, R0, R1, lp):
    """-> pre-MLP scalars (B,N,C'), vectors (B,N,C',4), and the neighbour sums the backward keeps."""
    As = neighbour_sums(R0, R1, S, V, pair_diff(y))
    cat00, cat11 = assemble_cat(S, V, As)
    S_pre = torch.einsum("ok,bnk->bno", lp["m00"], cat00)
    V_out = torch.einsum("ok,bnkm->bnom", lp["m11"], cat11)
    return S_pre, V_out, As


def mlp_forward(S_pre, lin, slope=0.01):
    """CGMLP on interleaved re/im rows (lgn_levels.py:191-227). Returns output and activations."""
    x = torch.view_as_real(S_pre).reshape(*S_pre.shape[:2], -1)
    acts = [x]
    for w, b in lin[:-1]:
        x = torch.nn.functional.leaky_relu(x @ w.T + b, slope)
        acts.append(x)
    w, b = lin[-1]
    x = x @ w.T + b
    return torch.view_as_complex(x.reshape(*S_pre.shape, 2).contiguous()), acts


def mlp_backward(gS, acts, lin, slope=0.01):
    g = torch.view_as_real(gS.contiguous()).reshape(*gS.shape[:2], -1)
    grads = [None] * len(lin)
    for l in range(len(lin) - 1, -1, -1):
        w, b = lin[l]
        h_in = acts[l]
        grads[l] = (torch.einsum("bno,bnk->ok", g, h_in), g.sum((0, 1)))
        g = g @ w
        if l > 0:
            g = g * torch.where(h_in > 0, torch.ones((), dtype=g.dtype), torch.full((), slope, dtype=g.dtype))
    return torch.view_as_complex(g.reshape(*gS.shape, 2).contiguous()), grads


# ---------------------------------------------------------------------------------------------
# one LGN level, hand-written backward (what the CUDA kernel implements)
# ---------------------------------------------------------------------------------------------
def level_backward(gS_pre, gV_out, S, V, y, R0, R1, As, lp, need_gy: bool):
    """Returns gS, gV, gy (or None), gR0, gR1, g_m00, g_m11.  gR^l is (B,N,N,C) for edge-dependent
    radial weights, (C,) for the constant (decoder) case."""
    C = S.shape[2]
    const = R0.dim() == 1
    Y = pair_diff(y)
    cat00, cat11 = assemble_cat(S, V, As)
    g_m00 = torch.einsum("bno,bnk->ok", gS_pre, cat00.conj())
    g_m11 = torch.einsum("bnom,bnkm->ok", gV_out, cat11.conj())
    gcat00 = torch.einsum("ok,bno->bnk", lp["m00"].conj(), gS_pre)
    gcat11 = torch.einsum("ok,bnom->bnkm", lp["m11"].conj(), gV_out)
    g_ag11a, g_ag11b, gV, g_sq11a, g_sq11b = [gcat11[:, :, k * C:(k + 1) * C] for k in range(5)]
    g_ag00a, g_ag00b, gS, g_sq00a, g_sq00b = [gcat00[:, :, k * C:(k + 1) * C] for k in range(5)]
    gV = gV.clone()
    gS = gS.clone()
    # self product
    g_sq11 = g_sq11a + g_sq11b
    gV += S.conj().unsqueeze(-1) * g_sq11 + ghat(V).conj() * g_sq00a.unsqueeze(-1)
    gS += (V.conj() * g_sq11).sum(-1) + 2 * S.conj() * g_sq00b
    # adjoint of the neighbour sums; index i = receiving node (axis 1), j = neighbour (axis 2)
    gA0V = ((1 - 1j) * g_ag11a).unsqueeze(2)          # (B,N,1,C,4)
    gA0S = ((1 - 1j) * g_ag00b).unsqueeze(2)          # (B,N,1,C)
    gA1Y = g_ag11b.unsqueeze(2)
    gA1E = (0.5 * g_ag00a).unsqueeze(2)
    R0b, R1b = _bc(R0, S), _bc(R1, S)
    Sj, Vj, Yc = S.unsqueeze(1), V.unsqueeze(1), Y.unsqueeze(3)          # Yc (B,N,N,1,4)
    # w_ij[c] = sum_mu conj(Y_ij,mu) gA1Y_i[c,mu] ;  e_ij[c] = eta(V_j[c], Y_ij)
    w = (Yc.conj() * gA1Y).sum(-1)                                         # (B,N,N,C)
    e = eta(Vj, Yc)
    gV = gV + (R0b.conj().unsqueeze(-1) * gA0V).sum(1) + ((R1b.conj() * gA1E).unsqueeze(-1) * ghat(Yc).conj()).sum(1)
    gS = gS + (R0b.conj() * gA0S).sum(1) + (R1b.conj() * w).sum(1)
    gR0 = (Vj.conj() * gA0V).sum(-1) + Sj.conj() * gA0S
    gR1 = Sj.conj() * w + e.conj() * gA1E
    if const:
        gR0, gR1 = gR0.sum((0, 1, 2)), gR1.sum((0, 1, 2))
    gy = None
    if need_gy:
        # gY_ij = sum_c [ conj(R1 S_j) gA1Y_i + conj(R1 ghat(V_j)) gA1E_i ];  gy_i += sum_j gY_ij ; gy_j -= sum_i gY_ij
        gY = ((R1b * Sj).conj().unsqueeze(-1) * gA1Y + (R1b.unsqueeze(-1) * ghat(Vj)).conj() * gA1E.unsqueeze(-1)).sum(3)
        gy = gY.sum(2) - gY.sum(1)
    return gS, gV, gy, gR0, gR1, g_m00, g_m11


def radial_enc_backward(gR0, gR1, norms, mask, phi, d, lp):
    """Gradients of a, b, c, w0, b0, w1, b1 (the norms are data: no gradient needed)."""
    def planar(g):
        return torch.stack([g.real, g.imag], -1).reshape(*g.shape[:-1], -1)   # (...,2C): 2c -> re, 2c+1 -> im
    gr0, gr1 = planar(gR0), planar(gR1)
    g_b0, g_b1 = gr0.sum((0, 1, 2)), gr1.sum((0, 1, 2))
    g_w0 = torch.einsum("bijo,bijk->ok", gr0, phi)
    g_w1 = torch.einsum("bijo,bijk->ok", gr1, phi)
    gphi = (gr0 @ lp["w0"] + gr1 @ lp["w1"]) * mask.unsqueeze(-1)
    x2 = (norms * norms).unsqueeze(-1)
    g_a = gphi.sum((0, 1, 2))
    g_b = (gphi / d).sum((0, 1, 2))
    g_c = (gphi * (-lp["b"] / (d * d)) * 2 * lp["c"] * x2).sum((0, 1, 2))
    return dict(a=g_a, b=g_b, c=g_c, w0=g_w0, b0=g_b0, w1=g_w1, b1=g_b1)


# ---------------------------------------------------------------------------------------------
# encoder / decoder glue
# ---------------------------------------------------------------------------------------------
def enc_prepare(p4, labels=None):
    """mass, canonical momenta, pair norms and radial mask (lgn_encoder.py:338-412,
    zonal_functions.py:123-166, 221-248)."""
    mass = minkowski_sq_real(p4).abs().sqrt()
    y = canon_real(p4)
    rel = p4.unsqueeze(2) - p4.unsqueeze(1)
    s = minkowski_sq_real(rel) + 1e-16
    norms = torch.where(s != 0, s / s.abs().sqrt(), s)
    node_mask = (labels != 0) if labels is not None else (p4[..., 0] != 0)
    mask = node_mask.unsqueeze(1) & node_mask.unsqueeze(2) & (norms != 0)
    return mass, y, norms, mask


def select_minmax(cart, scal):
    """min&max arg-selection (lgn_encoder.py:540-583): indices per (re/im, b, tau)."""
    def msq(x):  # x real (B,N,tau,4)
        return x[..., 0] ** 2 - torch.norm(x[..., 1:], dim=-1) ** 2
    idx = {}
    for nm, part in (("re", lambda z: z.real), ("im", lambda z: z.imag)):
        v, s = part(cart), part(scal)
        idx[nm] = dict(vmin=msq(v).argmin(1), vmax=msq(v).argmax(1), smin=s.argmin(1), smax=(s * s).argmax(1))
    return idx


def gather_nodes(z, idx_re, idx_im):
    """z (B,N,tau,...) complex; idx (B,tau) -> (B,tau,...) with re and im parts picked separately."""
    extra = z.shape[3:]
    def g(x, idx):
        ii = idx.view(idx.shape[0], 1, idx.shape[1], *([1] * len(extra))).expand(-1, 1, -1, *extra)
        return torch.gather(x, 1, ii).squeeze(1)
    return torch.complex(g(z.real, idx_re), g(z.imag, idx_im))


def scatter_nodes(g, idx_re, idx_im, n):
    """Adjoint of gather_nodes."""
    extra = g.shape[2:]
    def s(x, idx):
        out = torch.zeros(x.shape[0], n, *x.shape[1:], dtype=x.dtype)
        ii = idx.view(idx.shape[0], 1, idx.shape[1], *([1] * len(extra))).expand(-1, 1, -1, *extra)
        return out.scatter_add(1, ii, x.unsqueeze(1))
    return torch.complex(s(g.real, idx_re), s(g.imag, idx_im))


def chamfer(x, t):
    """chamfer_loss.py:16-31: returns loss and d loss / d x."""
    diff = x.unsqueeze(2) - t.unsqueeze(1)            # (B,Nx,Nt,4)
    dist = (diff * diff).sum(-1)
    m1, j1 = dist.min(2)
    m2, i2 = dist.min(1)
    loss = 0.5 * (m1.sum() + m2.sum())
    g = torch.gather(diff, 2, j1.view(*j1.shape, 1, 1).expand(-1, -1, 1, 4)).squeeze(2)   # x_i - t_j*(i)
    d2 = torch.gather(diff, 1, i2.view(i2.shape[0], 1, i2.shape[1], 1).expand(-1, 1, -1, 4)).squeeze(1)  # x_i*(j) - t_j
    g = g.scatter_add(1, i2.unsqueeze(-1).expand(-1, -1, 4), d2)
    return loss, g


# ---------------------------------------------------------------------------------------------
# whole model (maxdim 2, 'min&max' / 'mean' latent) chained from the stages above: forward and
# hand-written backward.  This is the sequence of launches the C side performs.
# ---------------------------------------------------------------------------------------------
def _w(sd, name):
    return planar_to_c(sd[name])


def _n_levels(sd):
    n = 0
    while f"lgn_cg.node_levels.{n}.cat_mix.mix_reps.weights.(0, 0)" in sd:
        n += 1
    return n


def model_step(enc_sd, dec_sd, batch, map_to_latent="min&max", l1_lambda=0.0, with_backward=True):
    p4 = batch["p4"]
    B, N, _ = p4.shape
    grads_e = {k: torch.zeros_like(v) for k, v in enc_sd.items()}
    grads_d = {k: torch.zeros_like(v) for k, v in dec_sd.items()}
    # ---------------- encoder forward
    mass, y, norms, mask = enc_prepare(p4, batch.get("labels"))
    win00, win11 = _w(enc_sd, "input_func_node.weights.(0, 0)")[:, 0], _w(enc_sd, "input_func_node.weights.(1, 1)")[:, 0]
    S = win00 * mass.unsqueeze(-1)
    V = win11.view(1, 1, -1, 1) * y.unsqueeze(-2)
    nl = _n_levels(enc_sd)
    enc_saved = []
    for lvl in range(nl):
        lp = level_params(enc_sd, lvl, True)
        R0, R1, phi, d = radial_enc(norms, mask, lp)
        S_pre, V_new, Ms = level_forward(S, V, y, R0, R1, lp)
        S_new, acts = mlp_forward(S_pre, lp["mlp"])
        enc_saved.append((S, V, R0, R1, phi, d, Ms, acts, lp))
        S, V = S_new, V_new
    wl00, wl11 = _w(enc_sd, "mix_reps.weights.(0, 0)"), _w(enc_sd, "mix_reps.weights.(1, 1)")
    L00 = torch.einsum("tc,bnc->bnt", wl00, S)
    L11 = cart_from_canon(torch.einsum("tc,bncm->bntm", wl11, V))
    if map_to_latent == "min&max":
        idx = select_minmax(L11, L00)
        lat00 = torch.cat([gather_nodes(L00, idx["re"]["smin"], idx["im"]["smin"]),
                           gather_nodes(L00, idx["re"]["smax"], idx["im"]["smax"])], 1)
        lat11 = torch.cat([gather_nodes(L11, idx["re"]["vmin"], idx["im"]["vmin"]),
                           gather_nodes(L11, idx["re"]["vmax"], idx["im"]["vmax"])], 1)
    elif map_to_latent == "mean":
        lat00, lat11 = L00.mean(1), L11.mean(1)
    else:
        raise NotImplementedError(map_to_latent)
    # ---------------- decoder forward
    wg11 = _w(dec_sd, "latent_to_graph.weights.(1, 1)")                  # (N, tau)
    P = torch.einsum("nt,btm->bnm", wg11, lat11)
    yd = canon_cplx(P)
    din00, din11 = _w(dec_sd, "input_func_node.weights.(0, 0)")[:, 0], _w(dec_sd, "input_func_node.weights.(1, 1)")[:, 0]
    Sd = (din00 * (1 + 1j)).view(1, 1, -1).expand(B, N, -1)
    Vd = din11.view(1, 1, -1, 1) * yd.unsqueeze(-2)
    dec_saved = []
    for lvl in range(_n_levels(dec_sd)):
        lp = level_params(dec_sd, lvl, False)
        R0, R1 = radial_dec(lp)
        S_pre, V_new, Ms = level_forward(Sd, Vd, yd, R0, R1, lp)
        S_new, acts = mlp_forward(S_pre, lp["mlp"])
        dec_saved.append((Sd, Vd, R0, R1, Ms, acts, lp))
        Sd, Vd = S_new, V_new
    wo11 = _w(dec_sd, "mix_to_output.weights.(1, 1)")                     # (1, C)
    gen11 = torch.einsum("oc,bncm->bnom", wo11, Vd)[:, :, 0]
    recons = cart_from_canon(gen11)                                        # (B,N,4) complex
    x = recons.real + recons.imag
    loss, gx = chamfer(x, p4)
    if l1_lambda:
        loss = loss + l1_lambda * (sum(v.abs().sum() for v in enc_sd.values()) + sum(v.abs().sum() for v in dec_sd.values()))
    out = dict(lat00=lat00, lat11=lat11, recons=recons, loss=loss)
    if not with_backward:
        return out
    # ---------------- decoder backward
    g_rec = gx * (1 + 1j)
    g_gen11 = cart_from_canon_bwd(g_rec)
    grads_d["mix_to_output.weights.(1, 1)"] = c_to_planar(torch.einsum("bnm,bncm->c", g_gen11, Vd.conj()).unsqueeze(0))
    gV = wo11.conj()[0].view(1, 1, -1, 1) * g_gen11.unsqueeze(2)
    gS = torch.zeros_like(Sd)
    gy = torch.zeros_like(yd)
    nld = len(dec_saved)
    for lvl in range(nld - 1, -1, -1):
        Sd_in, Vd_in, R0, R1, Ms, acts, lp = dec_saved[lvl]
        gS_pre, mg = mlp_backward(gS, acts, lp["mlp"])
        for i, (gw, gb) in enumerate(mg):
            grads_d[f"lgn_cg.mlp_levels.{lvl}.linear.{i}.weight"] = gw
            grads_d[f"lgn_cg.mlp_levels.{lvl}.linear.{i}.bias"] = gb
        gS, gV, gyl, gR0, gR1, g_m00, g_m11 = level_backward(gS_pre, gV, Sd_in, Vd_in, yd, R0, R1, Ms, lp, True)
        gy = gy + gyl
        grads_d[f"lgn_cg.node_levels.{lvl}.cat_mix.mix_reps.weights.(0, 0)"] = c_to_planar(g_m00)
        grads_d[f"lgn_cg.node_levels.{lvl}.cat_mix.mix_reps.weights.(1, 1)"] = c_to_planar(g_m11)
        grads_d[f"rad_funcs.rad_funcs.{lvl}.linear.0.bias"] = gR0.real + gR0.imag
        grads_d[f"rad_funcs.rad_funcs.{lvl}.linear.1.bias"] = gR1.real + gR1.imag
    grads_d["input_func_node.weights.(0, 0)"] = c_to_planar(((1 - 1j) * gS.sum((0, 1))).unsqueeze(-1))
    grads_d["input_func_node.weights.(1, 1)"] = c_to_planar((yd.conj().unsqueeze(-2) * gV).sum((0, 1, 3)).unsqueeze(-1))
    gy = gy + (din11.conj().view(1, 1, -1, 1) * gV).sum(2)
    gP = canon_cplx_bwd(gy)
    grads_d["latent_to_graph.weights.(1, 1)"] = c_to_planar(torch.einsum("bnm,btm->nt", gP, lat11.conj()))
    g_lat11 = torch.einsum("nt,bnm->btm", wg11.conj(), gP)
    # ---------------- encoder backward
    tau = wl11.shape[0]
    if map_to_latent == "min&max":
        gL11 = (scatter_nodes(g_lat11[:, :tau], idx["re"]["vmin"], idx["im"]["vmin"], N)
                + scatter_nodes(g_lat11[:, tau:], idx["re"]["vmax"], idx["im"]["vmax"], N))
    else:
        gL11 = (g_lat11 / N).unsqueeze(1).expand(-1, N, -1, -1)
    gLc = cart_from_canon_bwd(gL11)
    grads_e["mix_reps.weights.(1, 1)"] = c_to_planar(torch.einsum("bntm,bncm->tc", gLc, V.conj()))
    gV = torch.einsum("tc,bntm->bncm", wl11.conj(), gLc)
    gS = torch.zeros_like(S)
    for lvl in range(nl - 1, -1, -1):
        S_in, V_in, R0, R1, phi, d, Ms, acts, lp = enc_saved[lvl]
        gS_pre, mg = mlp_backward(gS, acts, lp["mlp"])
        for i, (gw, gb) in enumerate(mg):
            grads_e[f"lgn_cg.mlp_levels.{lvl}.linear.{i}.weight"] = gw
            grads_e[f"lgn_cg.mlp_levels.{lvl}.linear.{i}.bias"] = gb
        gS, gV, _, gR0, gR1, g_m00, g_m11 = level_backward(gS_pre, gV, S_in, V_in, y, R0, R1, Ms, lp, False)
        grads_e[f"lgn_cg.node_levels.{lvl}.cat_mix.mix_reps.weights.(0, 0)"] = c_to_planar(g_m00)
        grads_e[f"lgn_cg.node_levels.{lvl}.cat_mix.mix_reps.weights.(1, 1)"] = c_to_planar(g_m11)
        rg = radial_enc_backward(gR0, gR1, norms, mask, phi, d, lp)
        pre = f"rad_funcs.rad_funcs.{lvl}"
        for nm, key in (("a", ".a"), ("b", ".b"), ("c", ".c")):
            grads_e[pre + key] = rg[nm].view(1, 1, 1, -1)
        grads_e[pre + ".linear.0.weight"], grads_e[pre + ".linear.0.bias"] = rg["w0"], rg["b0"]
        grads_e[pre + ".linear.1.weight"], grads_e[pre + ".linear.1.bias"] = rg["w1"], rg["b1"]
    grads_e["input_func_node.weights.(0, 0)"] = c_to_planar((mass.unsqueeze(-1) * gS).sum((0, 1)).unsqueeze(-1))
    grads_e["input_func_node.weights.(1, 1)"] = c_to_planar((y.conj().unsqueeze(-2) * gV).sum((0, 1, 3)).unsqueeze(-1))
    if l1_lambda:
        for sd, gr in ((enc_sd, grads_e), (dec_sd, grads_d)):
            for k in gr:
                gr[k] = gr[k] + l1_lambda * torch.sign(sd[k])
    out.update(grads_enc=grads_e, grads_dec=grads_d)
    return out
