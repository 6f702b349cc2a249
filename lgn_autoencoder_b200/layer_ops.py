"""Layer-level ops of the reference's generic API on the hand-written kernels of csrc/lgae_cg.cu and csrc/lgae_layers.cu:

* ``cg_pairs`` — Clebsch-Gordan product of GVec parts, channel-wise, point-wise or aggregated over the neighbour axis
  (reference lgn/cg_lib/cg_ops.py:135-298), any maxdim, with its hand-written adjoint;
* ``mix``      — per-irrep complex channel mixing (reference lgn/g_lib/cplx_lib.py:7-25) with its adjoint;
* ``scalar_irrep`` — complex scalar x irrep product with channel broadcast (edge features, cplx_lib.py:54-72);
* ``radial_functions`` — RadPolyTrig: bells + mask + one Linear per zonal degree (lgn/nn/position_levels.py:118-209);
* ``linear``   — Linear (+ LeakyReLU) on rows, the CGMLP layers (lgn/models/lgn_levels.py:191-227).

All take and return the reference's planar complex tensors (2, ..., C, d) and are ``torch.autograd.Function``s over the
C ABI (include/lgae_b200.h).  There is no CPU path: tensors must be fp64 CUDA tensors."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"lgn_autoencoder_b200 {what}: tensors must live on a CUDA device (there is no CPU path)")
    if t.dtype != torch.float64:
        raise ValueError(f"lgn_autoencoder_b200 {what}: only torch.float64 is supported (as in the reference), got {t.dtype}")


# ------------------------------------------------------------------------------------------------------------
# CG product
# ------------------------------------------------------------------------------------------------------------
class PairPlan:
    """One (irrep1, irrep2) pair of a cg_product call: its output irreps, channel offsets and device term tables."""

    def __init__(self, i1, i2, desc, tab, coef, out_slots, swap, keys=None, outs=None):
        self.keys, self.outs = keys, outs  # ((k1,n1),(k2,n2)) and the pair's output irreps
        self.i1, self.i2 = i1, i2          # indices of the operand tensors (kernel order: z1 = node side, z2 = edge side)
        self.desc, self.tab, self.coef = desc, tab, coef
        self.out_slots = out_slots         # index of every output irrep in the call's output list
        self.swap = swap


def _pair_terms(cg_dict, key1, key2, out_keys, swap):
    """Host-side term list [(component, a, d, coef)] of a pair (a indexes the node-side operand of the kernels) and the number
    of output components; cached on the CG dictionary."""
    cache = cg_dict.__dict__.setdefault("_lgae_pair_terms", {})
    ck = (key1, key2, tuple(out_keys), bool(swap))
    if ck not in cache:
        d2 = (key2[0] + 1) * (key2[1] + 1)
        terms, comp0 = [], 0
        for ok in out_keys:
            h = cg_dict[(key1, key2)][ok].detach().cpu().numpy()
            for m, idx in zip(*np.nonzero(h)):
                a, d = divmod(int(idx), d2)
                if swap:
                    a, d = d, a
                terms.append((comp0 + int(m), a, d, float(h[m, idx])))
            comp0 += h.shape[0]
        cache[ck] = (terms, comp0)
    return cache[ck]


def _term_tables(cg_dict, key1, key2, out_keys, swap, device):
    """Non-zero CG coefficients H_o[m, a*d2+d] of a pair as term lists in three orders (see include/lgae_b200.h)."""
    cache = cg_dict.__dict__.setdefault("_lgae_term_cache", {})
    ck = (key1, key2, tuple(out_keys), bool(swap), str(device))
    if ck in cache:
        return cache[ck]
    d2 = (key2[0] + 1) * (key2[1] + 1)
    terms, comp0 = _pair_terms(cg_dict, key1, key2, out_keys, swap)
    d1k = (key1[0] + 1) * (key1[1] + 1)
    dk1, dk2 = (d2, d1k) if swap else (d1k, d2)     # kernel-side d1, d2
    n = len(terms)
    tabs, coefs, starts = [], [], []
    for col, groups in ((0, comp0), (1, dk1), (2, dk2)):
        order = sorted(range(n), key=lambda t: (terms[t][col], t))
        tabs.append([[terms[t][0], terms[t][1], terms[t][2]] for t in order])
        coefs.append([terms[t][3] for t in order])
        cnt = np.bincount([terms[t][col] for t in order], minlength=groups)
        starts.append(np.concatenate(([0], np.cumsum(cnt))))
    tab = np.concatenate([np.asarray(tabs, dtype=np.int32).reshape(-1)] + [s.astype(np.int32) for s in starts])
    out = (torch.from_numpy(tab).to(device), torch.tensor(coefs, dtype=torch.float64).reshape(-1).to(device), n, comp0, dk1, dk2)
    cache[ck] = out
    return out


def plan_pairs(cg_dict, keys1, keys2, chans1, chans2, max_dim, swap, device):
    """Pairs in the reference's loop order (rep1 outer, rep2 inner, cg_ops.py:176-215) with the channel offset of every
    result inside the concatenated output irreps.  Returns (pair plans, output keys in first-appearance order,
    channels per output key)."""
    out_keys, out_ch, pairs = [], {}, []
    for i1, key1 in enumerate(keys1):
        for i2, key2 in enumerate(keys2):
            (k1, n1), (k2, n2) = key1, key2
            if max(k1, n1, k2, n2) > max_dim - 1:
                continue
            if chans1[i1] != chans2[i2]:
                raise ValueError(f"The number of fragments must be same for each part! {chans1[i1]} {chans2[i2]}")
            outs = [(k, n) for k in range(abs(k1 - k2), min(max_dim, k1 + k2 + 1), 2) for n in range(abs(n1 - n2), min(max_dim, n1 + n2 + 1), 2)]
            if not outs:
                continue
            if len(outs) > 16:
                raise NotImplementedError("more than 16 output irreps per pair")
            tab, coef, n_terms, n_comp, dk1, dk2 = _term_tables(cg_dict, key1, key2, outs, swap, device)
            desc = _lib.LgaeCgPairDesc()
            desc.d1, desc.d2, desc.channels, desc.n_out, desc.n_comp, desc.n_terms = dk1, dk2, chans1[i1], len(outs), n_comp, n_terms
            comp0, slots = 0, []
            for o, ok in enumerate(outs):
                if ok not in out_ch:
                    out_ch[ok] = 0
                    out_keys.append(ok)
                desc.out_d[o] = (ok[0] + 1) * (ok[1] + 1)
                desc.out_comp0[o] = comp0
                desc.out_coffset[o] = out_ch[ok]
                comp0 += desc.out_d[o]
                out_ch[ok] += chans1[i1]
                slots.append(out_keys.index(ok))
            pairs.append(PairPlan(i1, i2, desc, tab, coef, slots, swap, keys=(key1, key2), outs=outs))
    for p in pairs:
        for o, s in enumerate(p.out_slots):
            p.desc.out_ctotal[o] = out_ch[out_keys[s]]
    return pairs, out_keys, [out_ch[k] for k in out_keys]


class MultiPlan:
    """All pairs of an aggregated cg_product call for lgae_cg_aggregate_multi_forward (one launch)."""

    def __init__(self, cg_dict, pairs, dims1, dims2, chans, out_keys, out_ch, swap, device):
        node_d, edge_d = (dims2, dims1) if swap else (dims1, dims2)
        node_off, edge_off = np.concatenate(([0], np.cumsum(node_d))), np.concatenate(([0], np.cumsum(edge_d)))
        rows, coefs, starts, cinfo, ncomp = [], [], [0], [], 0
        for pp in pairs:
            terms, n_comp_pair = _pair_terms(cg_dict, pp.keys[0], pp.keys[1], pp.outs, swap)
            i_node, i_edge = (pp.i2, pp.i1) if swap else (pp.i1, pp.i2)
            by_comp = [[] for _ in range(n_comp_pair)]
            for comp, a, d, coef in terms:
                by_comp[comp].append((int(node_off[i_node]) + a, int(edge_off[i_edge]) + d, coef))
            for comp in range(n_comp_pair):
                o = max(k for k in range(pp.desc.n_out) if pp.desc.out_comp0[k] <= comp)
                for a_all, d_all, coef in by_comp[comp]:
                    rows.append((ncomp + comp, a_all, d_all))
                    coefs.append(coef)
                starts.append(len(rows))
                cinfo.append((pp.out_slots[o], comp - pp.desc.out_comp0[o], pp.desc.out_coffset[o]))
            ncomp += n_comp_pair
        desc = _lib.LgaeCgMultiDesc()
        desc.channels, desc.n_node, desc.n_edge, desc.n_out, desc.n_comp, desc.n_terms = chans, len(node_d), len(edge_d), len(out_keys), ncomp, len(rows)
        for i, v in enumerate(node_d):
            desc.node_d[i] = v
        for i, v in enumerate(edge_d):
            desc.edge_d[i] = v
        for i, (k, c) in enumerate(zip(out_keys, out_ch)):
            desc.out_d[i], desc.out_ctotal[i] = (k[0] + 1) * (k[1] + 1), c
        tab = np.concatenate([np.asarray(rows, dtype=np.int32).reshape(-1), np.asarray(starts, dtype=np.int32),
                              np.asarray(cinfo, dtype=np.int32).reshape(-1)])
        self.desc = desc
        self.tab = torch.from_numpy(tab).to(device)
        self.coef = torch.tensor(coefs, dtype=torch.float64, device=device)
        # adjoint order: terms sorted by cell = a_all * D2T + d_all
        self.d1t, self.d2t = int(sum(node_d)), int(sum(edge_d))
        cells = [r[1] * self.d2t + r[2] for r in rows]
        order = sorted(range(len(rows)), key=lambda t: (cells[t], t))
        cnt = np.bincount([cells[t] for t in order], minlength=self.d1t * self.d2t)
        tab_b = np.concatenate([np.asarray([rows[t][0] for t in order], dtype=np.int32), np.concatenate(([0], np.cumsum(cnt))).astype(np.int32),
                                np.asarray(cinfo, dtype=np.int32).reshape(-1)])
        self.tab_b = torch.from_numpy(tab_b).to(device)
        self.coef_b = torch.tensor([coefs[t] for t in order], dtype=torch.float64, device=device)


def plan_multi(cg_dict, pairs, parts1, parts2, out_keys, out_ch, swap):
    """MultiPlan for an aggregated call, or None when the one-launch kernel does not apply (mixed channel counts, too many
    parts / irreps, edge dimensions other than 1, 4 or 5 in total)."""
    dims1, dims2 = [int(p.shape[-1]) for p in parts1], [int(p.shape[-1]) for p in parts2]
    chans = {int(p.shape[-2]) for p in list(parts1) + list(parts2)}
    edge_total = sum(dims1 if swap else dims2)
    if len(chans) != 1 or len(parts1) > _lib.CG_MAX_PARTS or len(parts2) > _lib.CG_MAX_PARTS or len(out_keys) > _lib.CG_MAX_OUT or \
            edge_total not in (1, 4, 5):
        return None
    cache = cg_dict.__dict__.setdefault("_lgae_multi_cache", {})
    ck = (tuple(pp.keys for pp in pairs), tuple(dims1), tuple(dims2), tuple(chans), tuple(out_keys), tuple(out_ch), bool(swap), str(parts1[0].device))
    if ck not in cache:
        cache[ck] = MultiPlan(cg_dict, pairs, dims1, dims2, chans.pop(), out_keys, out_ch, swap, parts1[0].device)
    return cache[ck]


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _CGPairsFn(torch.autograd.Function):
    """All pairs of one cg_product call.  Inputs: the n1 parts of rep1 then the n2 parts of rep2."""

    @staticmethod
    def forward(ctx, cg_dict, pairs, out_keys, out_ch, n1, n_nbr, swap, *parts):
        lib = _lib.load()
        parts = [p.contiguous() for p in parts]
        for p in parts:
            _require_cuda(p, "cg_product")
        # the side without the neighbour axis fixes the output batch shape
        node_parts = parts[n1:] if swap else parts[:n1]
        batch = tuple(node_parts[0].shape[1:-2])
        rows = int(np.prod(batch)) if batch else 1
        outs = [torch.empty((2,) + batch + (c, (k[0] + 1) * (k[1] + 1)), dtype=torch.float64, device=parts[0].device)
                for k, c in zip(out_keys, out_ch)]
        st = _stream()
        multi = plan_multi(cg_dict, pairs, parts[:n1], parts[n1:], out_keys, out_ch, swap) if n_nbr > 0 else None
        if multi is not None:
            # one launch for all (node irrep, edge irrep) pairs: the edge tensor is read once
            node, edge = (parts[n1:], parts[:n1]) if swap else (parts[:n1], parts[n1:])
            rc = lib.lgae_cg_aggregate_multi_forward(C.byref(multi.desc), multi.tab.data_ptr(), multi.coef.data_ptr(), _ptr_array(node),
                                                     _ptr_array(edge), rows, n_nbr, _ptr_array(outs), st)
            if rc != -2:   # LGAE_E_UNSUPPORTED (e.g. shared memory): fall through to the per-pair launches
                _lib.check(rc, "cg_aggregate_multi_forward")
                multi = True
            else:
                multi = None
        for pp in ([] if multi else pairs):
            a, b = parts[pp.i1], parts[n1 + pp.i2]
            z1, z2 = (b, a) if swap else (a, b)
            _lib.check(lib.lgae_cg_product_forward(C.byref(pp.desc), pp.tab.data_ptr(), pp.coef.data_ptr(), z1.data_ptr(), z2.data_ptr(),
                                                   rows, n_nbr, _ptr_array([outs[s] for s in pp.out_slots]), st), "cg_product_forward")
        ctx.pairs, ctx.n1, ctx.n_nbr, ctx.swap, ctx.rows = pairs, n1, n_nbr, swap, rows
        ctx.multi = plan_multi(cg_dict, pairs, parts[:n1], parts[n1:], out_keys, out_ch, swap) if multi else None
        ctx.out_shapes = [tuple(o.shape) for o in outs]
        ctx.save_for_backward(*parts)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        lib = _lib.load()
        parts = ctx.saved_tensors
        n1, swap = ctx.n1, ctx.swap
        need = ctx.needs_input_grad[7:]
        g_outs = [g.contiguous() if g is not None else None for g in g_outs]
        grads = [None] * len(parts)
        st = _stream()
        pairs_todo = ctx.pairs
        if ctx.multi is not None:
            # one launch per operand side for all pairs (falls back to the per-pair adjoint when unsupported)
            mp = ctx.multi
            gl = [g if g is not None else torch.zeros(shp, dtype=torch.float64, device=parts[0].device) for g, shp in zip(g_outs, ctx.out_shapes)]
            node_idx = list(range(n1, len(parts))) if swap else list(range(n1))
            edge_idx = list(range(n1)) if swap else list(range(n1, len(parts)))
            node_ok = mp.d2t == 5 and mp.d1t in (5, 20) and ctx.n_nbr * mp.desc.channels <= 256
            want_node = node_ok and any(need[i] for i in node_idx)
            want_edge = any(need[i] for i in edge_idx)
            if want_node or want_edge:
                gn = [torch.empty_like(parts[i]) if (want_node and need[i]) else None for i in node_idx]
                ge = [torch.empty_like(parts[i]) if (want_edge and need[i]) else None for i in edge_idx]
                arr_n, arr_e = (C.c_void_p * len(gn))(), (C.c_void_p * len(ge))()
                for k, t in enumerate(gn):
                    arr_n[k] = t.data_ptr() if t is not None else None
                for k, t in enumerate(ge):
                    arr_e[k] = t.data_ptr() if t is not None else None
                rc = lib.lgae_cg_aggregate_multi_backward(C.byref(mp.desc), mp.tab_b.data_ptr(), mp.coef_b.data_ptr(),
                                                          _ptr_array([parts[i] for i in node_idx]), _ptr_array([parts[i] for i in edge_idx]),
                                                          ctx.rows, ctx.n_nbr, _ptr_array(gl), arr_n, arr_e, st)
                if rc == 0:
                    for i, t in zip(node_idx, gn):
                        grads[i] = t
                    for i, t in zip(edge_idx, ge):
                        grads[i] = t
                    need = list(need)
                    for i in (node_idx if want_node else []) + (edge_idx if want_edge else []):
                        need[i] = False     # done; the per-pair loop below only covers what is left
                elif rc != -2:
                    _lib.check(rc, "cg_aggregate_multi_backward")
            if not any(need):
                pairs_todo = []
        done = [g is not None for g in grads]
        for pp in pairs_todo:
            ia, ib = pp.i1, n1 + pp.i2
            gl = []
            for o, s in enumerate(pp.out_slots):
                if g_outs[s] is None:   # an output nobody used: zero gradient
                    shape = (2,) + tuple(parts[ib if swap else ia].shape[1:-2]) + (pp.desc.out_ctotal[o], pp.desc.out_d[o])
                    g_outs[s] = torch.zeros(shape, dtype=torch.float64, device=parts[0].device)
                gl.append(g_outs[s])
            acc, ptrs = {}, {}
            for idx in (ia, ib):
                if need[idx]:
                    acc[idx] = 1 if grads[idx] is not None else 0
                    if grads[idx] is None:
                        grads[idx] = torch.empty_like(parts[idx])
                    ptrs[idx] = grads[idx].data_ptr()
                else:
                    acc[idx], ptrs[idx] = 0, None
            a, b = parts[ia], parts[ib]
            (z1, i1), (z2, i2) = ((b, ib), (a, ia)) if swap else ((a, ia), (b, ib))
            _lib.check(lib.lgae_cg_product_backward(C.byref(pp.desc), pp.tab.data_ptr(), pp.coef.data_ptr(), z1.data_ptr(), z2.data_ptr(),
                                                    ctx.rows, ctx.n_nbr, _ptr_array(gl), ptrs[i1], ptrs[i2], acc[i1], acc[i2], st),
                       "cg_product_backward")
        for idx, p in enumerate(parts):   # parts that met no partner under the maxdim cut
            if need[idx] and grads[idx] is None:
                grads[idx] = torch.zeros_like(p)
        del done
        return (None,) * 7 + tuple(grads)


def cg_pairs(cg_dict, keys1, parts1, keys2, parts2, max_dim, aggregate):
    """Returns (output keys, output tensors).  aggregate: one operand has the extra neighbour axis (2,B,N,N,C,d)."""
    n_nbr, swap = 0, False
    if aggregate:
        if parts2[0].dim() == parts1[0].dim() + 1:
            swap = False
        elif parts1[0].dim() == parts2[0].dim() + 1:
            swap = True
        else:
            raise ValueError(f"Batch size error! {tuple(parts1[0].shape)} {tuple(parts2[0].shape)}")
        edge = parts1[0] if swap else parts2[0]
        node = parts2[0] if swap else parts1[0]
        if edge.dim() != 6 or tuple(edge.shape[1:4]) != (node.shape[1], node.shape[2], node.shape[2]):
            raise ValueError(f"Batch size error! {tuple(parts1[0].shape)} {tuple(parts2[0].shape)}")
        n_nbr = int(node.shape[2])
    else:
        for p, q in zip(parts1[:1], parts2[:1]):
            if p.shape[1:-2] != q.shape[1:-2]:
                raise ValueError(f"shape mismatch {tuple(p.shape)} vs {tuple(q.shape)}")
    pairs, out_keys, out_ch = plan_pairs(cg_dict, list(keys1), list(keys2), [int(p.shape[-2]) for p in parts1],
                                         [int(p.shape[-2]) for p in parts2], max_dim, swap, parts1[0].device)
    if not pairs:
        return [], []
    outs = _CGPairsFn.apply(cg_dict, pairs, out_keys, out_ch, len(parts1), n_nbr, swap, *parts1, *parts2)
    return out_keys, list(outs)


# ------------------------------------------------------------------------------------------------------------
# channel mixing
# ------------------------------------------------------------------------------------------------------------
class _MixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, part):
        lib = _lib.load()
        _require_cuda(part, "mix")
        _require_cuda(weight, "mix")
        w, x = weight.contiguous(), part.contiguous()
        c_out, c_in = int(w.shape[1]), int(w.shape[2])
        d = int(x.shape[-1])
        if x.shape[-2] != c_in:
            raise ValueError(f"mix: weight {tuple(w.shape)} does not match part {tuple(x.shape)}")
        rows = int(np.prod(x.shape[1:-2])) if x.dim() > 3 else 1
        out = torch.empty(tuple(x.shape[:-2]) + (c_out, d), dtype=torch.float64, device=x.device)
        _lib.check(lib.lgae_mix_forward(w.data_ptr(), x.data_ptr(), rows, c_in, c_out, d, out.data_ptr(), _stream()), "mix_forward")
        ctx.save_for_backward(w, x)
        ctx.dims = (rows, c_in, c_out, d)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        w, x = ctx.saved_tensors
        rows, c_in, c_out, d = ctx.dims
        g = g.contiguous()
        gw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        gx = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        part = None
        if gw is not None:
            part = torch.empty(int(lib.lgae_mix_partials_doubles(rows, c_in, c_out)), dtype=torch.float64, device=x.device)
        _lib.check(lib.lgae_mix_backward(w.data_ptr(), x.data_ptr(), g.data_ptr(), rows, c_in, c_out, d, _lib.ptr(gx), _lib.ptr(gw),
                                         _lib.ptr(part), _stream()), "mix_backward")
        return gw, gx


def mix(weight, part):
    """(2,C',C) complex weight on the channel axis of a (2,...,C,d) part."""
    if weight.dim() != 3 or weight.shape[0] != 2:
        raise ValueError(f"mix: expected a complex weight of shape (2, C_out, C_in), got {tuple(weight.shape)}")
    return _MixFn.apply(weight, part)


# ------------------------------------------------------------------------------------------------------------
# complex scalar x irrep (edge features rad (x) zonal)
# ------------------------------------------------------------------------------------------------------------
class _ScalarIrrepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scalar, part):
        lib = _lib.load()
        s, v = scalar.contiguous(), part.contiguous()
        cs, cv, d = int(s.shape[-1]), int(v.shape[-2]), int(v.shape[-1])
        edges = int(np.prod(s.shape[1:-1])) if s.dim() > 2 else 1
        out = torch.empty(tuple(v.shape[:-2]) + (max(cs, cv), d), dtype=torch.float64, device=v.device)
        _lib.check(lib.lgae_scalar_irrep_forward(s.data_ptr(), v.data_ptr(), edges, cs, cv, d, out.data_ptr(), _stream()), "scalar_irrep_forward")
        ctx.save_for_backward(s, v)
        ctx.dims = (edges, cs, cv, d)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        s, v = ctx.saved_tensors
        edges, cs, cv, d = ctx.dims
        gs = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        gv = torch.empty_like(v) if ctx.needs_input_grad[1] else None
        _lib.check(lib.lgae_scalar_irrep_backward(s.data_ptr(), v.data_ptr(), g.contiguous().data_ptr(), edges, cs, cv, d, _lib.ptr(gs),
                                                  _lib.ptr(gv), _stream()), "scalar_irrep_backward")
        return gs, gv


def scalar_irrep(scalar, part):
    """(2,...,C) x (2,...,C,d) -> (2,...,C,d); either channel count may be 1 (broadcast)."""
    _require_cuda(scalar, "scalar x irrep")
    _require_cuda(part, "scalar x irrep")
    bs, bv = tuple(scalar.shape[1:-1]), tuple(part.shape[1:-2])
    if bs != bv:   # general batch broadcasting: materialise the expanded operands
        full = torch.broadcast_shapes(bs, bv)
        scalar = scalar.expand((2,) + full + (scalar.shape[-1],))
        part = part.expand((2,) + full + tuple(part.shape[-2:]))
    cs, cv = scalar.shape[-1], part.shape[-2]
    if cs != cv and cs != 1 and cv != 1:
        raise ValueError(f"scalar x irrep: channel counts {cs} and {cv} do not broadcast")
    return _ScalarIrrepFn.apply(scalar, part)


# ------------------------------------------------------------------------------------------------------------
# RadPolyTrig
# ------------------------------------------------------------------------------------------------------------
class _RadialFn(torch.autograd.Function):
    """Inputs: norms, a, b, c, then W_0, bias_0, W_1, bias_1, ...  Outputs: one tensor per zonal degree."""

    @staticmethod
    def forward(ctx, edge_mask, planar, norms, a, b, c, *wb):
        lib = _lib.load()
        x = norms.contiguous()
        _require_cuda(x, "RadPolyTrig")
        ws = [t.contiguous() for t in wb[0::2]]
        bs = [t.contiguous() for t in wb[1::2]]
        a, b, c = a.contiguous(), b.contiguous(), c.contiguous()
        mask = (edge_mask != 0).to(torch.uint8).contiguous()
        n_l, n_out, k2 = len(ws), int(ws[0].shape[0]), int(ws[0].shape[1])
        edges, n_mask = x.numel(), mask.numel()
        if n_mask == 0 or edges % n_mask:
            raise ValueError(f"RadPolyTrig: mask {tuple(edge_mask.shape)} does not broadcast over norms {tuple(norms.shape)}")
        shape = ((2,) + tuple(x.shape) + (n_out // 2,)) if planar else (tuple(x.shape) + (n_out,))
        outs = [torch.empty(shape, dtype=torch.float64, device=x.device) for _ in range(n_l)]
        _lib.check(lib.lgae_radial_functions_forward(x.data_ptr(), mask.data_ptr(), edges, n_mask, a.data_ptr(), b.data_ptr(), c.data_ptr(), k2,
                                                     n_out, n_l, _ptr_array(ws), _ptr_array(bs), _ptr_array(outs), int(planar), _stream()),
                   "radial_functions_forward")
        ctx.save_for_backward(x, mask, a, b, c, *ws)
        ctx.dims = (edges, n_mask, k2, n_out, n_l, int(planar))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        lib = _lib.load()
        x, mask, a, b, c, *ws = ctx.saved_tensors
        edges, n_mask, k2, n_out, n_l, planar = ctx.dims
        shape = ((2,) + tuple(x.shape) + (n_out // 2,)) if planar else (tuple(x.shape) + (n_out,))
        gl = [g.contiguous() if g is not None else torch.zeros(shape, dtype=torch.float64, device=x.device) for g in g_outs]
        gx = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        gp = torch.empty(n_l * n_out * (k2 + 1) + 3 * k2, dtype=torch.float64, device=x.device)
        part = torch.empty(int(lib.lgae_radial_functions_partials_doubles(edges, k2, n_out, n_l)), dtype=torch.float64, device=x.device)
        _lib.check(lib.lgae_radial_functions_backward(x.data_ptr(), mask.data_ptr(), edges, n_mask, a.data_ptr(), b.data_ptr(), c.data_ptr(), k2,
                                                      n_out, n_l, _ptr_array(ws), _ptr_array(gl), planar, _lib.ptr(gx), gp.data_ptr(),
                                                      part.data_ptr(), _stream()), "radial_functions_backward")
        nw = n_l * n_out * k2
        gw = gp[:nw].view(n_l, n_out, k2)
        gb = gp[nw:nw + n_l * n_out].view(n_l, n_out)
        rest = gp[nw + n_l * n_out:].view(3, k2)
        wb = []
        for l in range(n_l):
            wb += [gw[l], gb[l]]
        return (None, None, gx, rest[0].view(a.shape), rest[1].view(b.shape), rest[2].view(c.shape)) + tuple(wb)


def radial_functions(norms, edge_mask, a, b, c, linears, planar):
    """linears: list of (weight (n_out, 2K), bias (n_out)).  Returns one tensor per zonal degree:
    planar (Cartesian basis): (2, *norms.shape, n_out/2); otherwise (canonical basis): (*norms.shape, n_out)."""
    wb = []
    for w, bias in linears:
        wb += [w, bias]
    return list(_RadialFn.apply(edge_mask, bool(planar), norms, a, b, c, *wb))


# ------------------------------------------------------------------------------------------------------------
# Linear (+ LeakyReLU)
# ------------------------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, slope):
        lib = _lib.load()
        _require_cuda(x, "linear")
        x2 = x.contiguous().view(-1, x.shape[-1])
        w = weight.contiguous()
        bvec = bias.contiguous() if bias is not None else None
        rows, n_in, n_out = int(x2.shape[0]), int(w.shape[1]), int(w.shape[0])
        if x2.shape[1] != n_in:
            raise ValueError(f"linear: input features {x2.shape[1]} do not match the weight {tuple(w.shape)}")
        y = torch.empty((rows, n_out), dtype=torch.float64, device=x.device)
        act = slope is not None
        _lib.check(lib.lgae_linear_forward(x2.data_ptr(), w.data_ptr(), _lib.ptr(bvec), rows, n_in, n_out, int(act), float(slope or 0.0),
                                           y.data_ptr(), _stream()), "linear_forward")
        ctx.save_for_backward(x2, w, y)
        ctx.meta = (rows, n_in, n_out, act, float(slope or 0.0), tuple(x.shape), bias is not None)
        return y.view(tuple(x.shape[:-1]) + (n_out,))

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x2, w, y = ctx.saved_tensors
        rows, n_in, n_out, act, slope, xshape, has_bias = ctx.meta
        g2 = g.contiguous().view(rows, n_out)
        gx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        gb = torch.empty(n_out, dtype=torch.float64, device=w.device) if (has_bias and ctx.needs_input_grad[2]) else None
        part = None
        if gw is not None or gb is not None:
            part = torch.empty(int(lib.lgae_linear_partials_doubles(rows, n_in, n_out)), dtype=torch.float64, device=w.device)
        _lib.check(lib.lgae_linear_backward(x2.data_ptr(), w.data_ptr(), y.data_ptr(), g2.data_ptr(), rows, n_in, n_out, int(act), slope,
                                            _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(part), _stream()), "linear_backward")
        return (gx.view(xshape) if gx is not None else None), gw, gb, None


def linear(x, weight, bias=None, leaky_slope=None):
    """y = x W^T + b on the last axis, followed by LeakyReLU(leaky_slope) when a slope is given."""
    return _LinearFn.apply(x, weight, bias, leaky_slope)
