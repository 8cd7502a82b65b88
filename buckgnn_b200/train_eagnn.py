"""Training step of the EA-GNN ("CustomGNN") variants: `GraphNetBlock` (Models/BuckGNN.py:528-566) and its
wrapper loop (:326-336, :375-387) in train mode, forward and backward on the sm_100a kernels.

Forward of one block, as the reference writes it (edges live in CSR order keyed by `row = edge_index[0]`):
    he  = relu(cat[x[row], x[col], e] W1^T + b1)  = relu((x W1a^T)[row] + (x W1b^T)[col] + e W1c^T + b1)
    e1  = he W2^T + b2                                                       (the block's edge output)
    hm  = relu(cat[x[col], e1] Wp1^T + bp1)       = relu((x Wp1a^T)[col] + e1 Wp1b^T + bp1)
    m   = hm Wp2^T + bp2;   agg = scatter_mean(m, row)
    g1  = relu(cat[x, agg] Wg1^T + bg1);  xg = g1 Wg2^T + bg2
    t   = relu(xg Wb1^T + bb1);           xo = xg + t Wb2^T + bb2
    wrapper: x' = dropout(xo [+ x]),  e' = dropout(e1 [+ e])
Every product is `bg_gemm512` (gathered rows enter as epilogue addends); unlike the eval path nothing is
folded, so each Linear keeps its own input for its weight gradient.  Backward: the same kernel on transposed
weights for input gradients, `bg_wgrad512` for weight gradients (the reduction runs over nodes or edges),
`bg_grad_mask` for ReLU / Dropout, `bg_segment_expand` for scatter_mean, and `bg_sage_aggregate` over the two
CSR orders for the gathers' scatter-adds.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import capi, engine
from .engine import Activation, _stream
from .train import (GradStore, _Saved, _f32, colsum, encoder_backward, encoder_forward_train, head_backward,
                    head_forward_train, is_node_level, layer_seed, node_head_backward, node_head_forward_train,
                    transposed_pack, weight_grad_mn)

_H = 512


def _blocks(model):
    if model.model_name == "EA_GNN_Shared":
        return [model.shared_gn_block] * model.num_layers
    return list(model.gn_blocks)


class _BlockWeights:
    """Operand-format slices of one GraphNetBlock (forward) and their transposes (backward), built on demand."""

    def __init__(self, blk, prec: str):
        self.blk, self.prec = blk, prec
        self._fwd, self._bwd = {}, {}

    def _slice(self, name):
        b = self.blk
        w1, wp1, wg1 = b.edge_mlp[0].weight, b.node_mlp_phi[0].weight, b.node_mlp_gamma[0].weight
        return {"w1a": w1[:, :_H], "w1b": w1[:, _H:2 * _H], "w1c": w1[:, 2 * _H:], "w2": b.edge_mlp[2].weight,
                "wp1a": wp1[:, :_H], "wp1b": wp1[:, _H:], "wp2": b.node_mlp_phi[2].weight,
                "wg1a": wg1[:, :_H], "wg1b": wg1[:, _H:], "wg2": b.node_mlp_gamma[2].weight,
                "wb1": b.node_mlp_beta[0].weight, "wb2": b.node_mlp_beta[2].weight}[name]

    def fwd(self, name):
        if name not in self._fwd:
            self._fwd[name] = engine.pack_linear(self._slice(name).detach().contiguous(), self.prec)
        return self._fwd[name]

    def bwd(self, name):
        if name not in self._bwd:
            self._bwd[name] = transposed_pack(self._slice(name).detach().contiguous(), self.prec)
        return self._bwd[name]


_BIAS_ORDER = ("b1", "b2", "bp1", "bp2", "bg1", "bg2", "bb1", "bb2")


def _bias_params(blk):
    return (blk.edge_mlp[0].bias, blk.edge_mlp[2].bias, blk.node_mlp_phi[0].bias, blk.node_mlp_phi[2].bias,
            blk.node_mlp_gamma[0].bias, blk.node_mlp_gamma[2].bias, blk.node_mlp_beta[0].bias, blk.node_mlp_beta[2].bias)


class EAGNNTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, edge_index, edge_attr, batch, seed, *params):
        prec = model.train_precision
        code = engine.PRECISION_FORMATS[prec][0]
        dev = x.device
        s = _stream()
        x = x.detach().to(torch.float32).contiguous()
        if edge_attr is None or not edge_attr.is_cuda:
            raise RuntimeError("buckgnn_b200: `edge_attr` must be a CUDA tensor")
        n = x.shape[0]
        p_drop = float(model.dropout.p)
        pending = engine.begin_graph_index(edge_index, batch, n, key_row=0)       # GraphNetBlock aggregates on row (:553,561)
        blocks = _blocks(model)
        L = len(blocks)
        uniq = []
        for b in blocks:
            if all(b is not q for q in uniq):
                uniq.append(b)
        # one read-back for every epilogue bias vector of the step
        bias_list = [model.node_encoder[4].bias, model.edge_encoder[4].bias] + [p for b in uniq for p in _bias_params(b)]
        biases = torch.stack([b.detach().float() for b in bias_list]).cpu()
        bias_of = {id(b): {k: biases[2 + 8 * j + i] for i, k in enumerate(_BIAS_ORDER)} for j, b in enumerate(uniq)}
        weights = {id(b): _BlockWeights(b, prec) for b in uniq}

        sv = _Saved()
        sv.model, sv.prec, sv.n, sv.x, sv.seed, sv.p_drop = model, prec, n, x, seed, p_drop
        sv.h1, sv.h2, cur = encoder_forward_train(model.node_encoder, x, prec, biases[0])       # :323
        idx = pending.finish()
        ne = idx.n_edges
        if ne == 0:
            raise NotImplementedError("buckgnn_b200: EA-GNN training needs at least one edge")
        sv.idx, sv.ne = idx, ne
        ex = engine.edge_extras(idx, prec)
        sv.ex = ex
        # edge features in CSR order (input re-ordering only), then the edge encoder (:327, :376)
        ea = edge_attr.detach().to(torch.float32).index_select(0, idx.perm[:ne].long()).contiguous()
        sv.ea = ea
        sv.eh1, sv.eh2, e = encoder_forward_train(model.edge_encoder, ea, prec, biases[1])
        G = lambda out, segs, m, **kw: engine.gemm512(segs, m, prec, out, **kw)
        seg = engine._segments
        mk = lambda rows: Activation(rows, _H, prec, dev)
        hp = lambda t: t.data_ptr()
        row_of, col = ex.row_of.data_ptr(), idx.col.data_ptr()
        agg_ws_bytes = capi.aggregate_workspace_bytes(idx.n_big)
        sv.layers = []
        for i, blk in enumerate(blocks):
            w, bz = weights[id(blk)], bias_of[id(blk)]
            with engine.TIMERS.span("gn_node_gemms"):
                P, Q, R = mk(n), mk(n), mk(n)
                G(P, seg(cur, w.fwd("w1a")), n)
                G(Q, seg(cur, w.fwd("w1b")), n)
                G(R, seg(cur, w.fwd("wp1a")), n)
            with engine.TIMERS.span("gn_edge_gemms"):
                he, e1, hm, m = mk(ne), mk(ne), mk(ne), mk(ne)
                G(he, seg(e, w.fwd("w1c")), ne, bias=hp(bz["b1"]), relu=True, gather=[(P.data.data_ptr(), row_of), (Q.data.data_ptr(), col)])
                G(e1, seg(he, w.fwd("w2")), ne, bias=hp(bz["b2"]))
                G(hm, seg(e1, w.fwd("wp1b")), ne, bias=hp(bz["bp1"]), relu=True, gather=[(R.data.data_ptr(), col)])
                G(m, seg(hm, w.fwd("wp2")), ne, bias=hp(bz["bp2"]))
            agg = mk(n)
            ws = torch.empty(agg_ws_bytes, dtype=torch.uint8, device=dev)
            with engine.TIMERS.span("gn_segment_mean"):
                capi.sage_aggregate(m.data.data_ptr(), agg.data.data_ptr(), code, n, idx.rowptr.data_ptr(), ex.iota.data_ptr(),
                                    idx.big_rows.data_ptr(), idx.n_big, capi.BG_AGGR_MEAN, ws.data_ptr(), agg_ws_bytes, s)
            agg.refresh_split()
            del m, P, Q, R
            with engine.TIMERS.span("gn_node_gemms"):
                g1, xg, t, xo = mk(n), mk(n), mk(n), mk(n)
                G(g1, seg(cur, w.fwd("wg1a")) + seg(agg, w.fwd("wg1b")), n, bias=hp(bz["bg1"]), relu=True)
                G(xg, seg(g1, w.fwd("wg2")), n, bias=hp(bz["bg2"]))
                G(t, seg(xg, w.fwd("wb1")), n, bias=hp(bz["bb1"]), relu=True)
                G(xo, seg(t, w.fwd("wb2")), n, bias=hp(bz["bb2"]), residual=xg.data.data_ptr(), ldr=_H)
            skip = 0 < i < L - 1
            x_next, e_next = mk(n), mk(ne)
            capi.dropout_residual(xo.data.data_ptr(), cur.data.data_ptr() if skip else None, x_next.data.data_ptr(), code, n,
                                  p_drop, layer_seed(seed, 2 * i), s)
            capi.dropout_residual(e1.data.data_ptr(), e.data.data_ptr() if skip else None, e_next.data.data_ptr(), code, ne,
                                  p_drop, layer_seed(seed, 2 * i + 1), s)
            x_next.refresh_split(); e_next.refresh_split()
            sv.layers.append((blk, cur, e, he, e1, hm, agg, g1, xg, t, skip))
            cur, e = x_next, e_next
        sv.weights = weights
        sv.node_level = is_node_level(model)
        pred = node_head_forward_train(model, cur, sv) if sv.node_level else head_forward_train(model, cur, idx, sv)
        ctx.sv = sv
        ctx.params = params
        return pred

    @staticmethod
    def backward(ctx, dpred):
        sv = ctx.sv
        model, prec, n, ne = sv.model, sv.prec, sv.n, sv.ne
        code = engine.PRECISION_FORMATS[prec][0]
        dev = sv.x.device
        s = _stream()
        idx, ex = sv.idx, sv.ex
        grads = GradStore(dev)
        dxn = (node_head_backward if sv.node_level else head_backward)(
            model, sv, dpred.detach().to(torch.float32).contiguous(), n, prec, grads)
        den: Optional[Activation] = None                    # the last block's edge output feeds nothing
        G = lambda out, segs, m, **kw: engine.gemm512(segs, m, prec, out, **kw)
        seg = engine._segments
        mk = lambda rows: Activation(rows, _H, prec, dev)
        # slots sorted by col: the scatter-add of a gradient that was gathered by `col` is a sum over this CSR
        ei_slots = torch.stack([ex.row_of[:ne].long(), idx.col[:ne].long()]).contiguous()
        idx_c = engine.build_graph_index(ei_slots, None, n, key_row=1)

        def sum_by(index, rowptr, big_rows, n_big, src: Activation) -> Activation:
            out = mk(n)
            nb = capi.aggregate_workspace_bytes(n_big)
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            capi.sage_aggregate(src.data.data_ptr(), out.data.data_ptr(), code, n, rowptr.data_ptr(), index.data_ptr(),
                                big_rows.data_ptr(), n_big, capi.BG_AGGR_SUM, ws.data_ptr(), nb, s)
            out.refresh_split()
            return out

        by_row = lambda src: sum_by(ex.iota, idx.rowptr, idx.big_rows, idx.n_big, src)
        by_col = lambda src: sum_by(idx_c.perm, idx_c.rowptr, idx_c.big_rows, idx_c.n_big, src)

        def masked(dy: Activation, act: Optional[Activation], rows: int, p: float = 0.0, sd: int = 0) -> Activation:
            out = mk(rows)
            capi.grad_mask(dy.data.data_ptr(), None, None if act is None else act.data.data_ptr(), out.data.data_ptr(), code, rows, p, sd, s)
            out.refresh_split()
            return out

        def linear_grads(dy: Activation, inp: Activation, rows: int, weight, col0: int, bias) -> None:
            """dW[:, col0:col0+512] += dy^T inp;  db += colsum(dy)  (bias None: a column slice that shares its bias)."""
            with engine.TIMERS.span("train_wgrad"):
                tmp = _f32((_H, _H), dev)
                weight_grad_mn(dy.data, inp.data, code, rows, tmp, False)
                grads.zeros(weight)[:, col0:col0 + _H] += tmp
                if bias is not None:
                    colsum(dy.data, code, rows, _H, _H, grads.zeros(bias), accumulate=True)

        def dgrad(dy: Activation, wt, rows: int, residual: Optional[Activation] = None) -> Activation:
            out = mk(rows)
            with engine.TIMERS.span("train_dgrad_gemm"):
                G(out, seg(dy, wt), rows, residual=None if residual is None else residual.data.data_ptr(), ldr=_H)
            return out

        L = len(sv.layers)
        for i in range(L - 1, -1, -1):
            blk, x_in, e_in, he, e1, hm, agg, g1, xg, t, skip = sv.layers[i]
            w = sv.weights[id(blk)]
            em, phi, gam, bet = blk.edge_mlp, blk.node_mlp_phi, blk.node_mlp_gamma, blk.node_mlp_beta
            dxo = masked(dxn, None, n, sv.p_drop, layer_seed(sv.seed, 2 * i))                   # Dropout backward
            de1 = None if den is None else masked(den, None, ne, sv.p_drop, layer_seed(sv.seed, 2 * i + 1))
            # xo = xg + t Wb2^T + bb2;  t = relu(xg Wb1^T + bb1)
            linear_grads(dxo, t, n, bet[2].weight, 0, bet[2].bias)
            dt = masked(dgrad(dxo, w.bwd("wb2"), n), t, n)
            linear_grads(dt, xg, n, bet[0].weight, 0, bet[0].bias)
            dxg = dgrad(dt, w.bwd("wb1"), n, residual=dxo)
            # xg = g1 Wg2^T + bg2;  g1 = relu(cat[x, agg] Wg1^T + bg1)
            linear_grads(dxg, g1, n, gam[2].weight, 0, gam[2].bias)
            dg1 = masked(dgrad(dxg, w.bwd("wg2"), n), g1, n)
            linear_grads(dg1, x_in, n, gam[0].weight, 0, gam[0].bias)
            linear_grads(dg1, agg, n, gam[0].weight, _H, None)
            dx = dgrad(dg1, w.bwd("wg1a"), n, residual=dxo if skip else None)              # + the wrapper's skip
            dagg = dgrad(dg1, w.bwd("wg1b"), n)
            # agg = scatter_mean(m, row);  m = hm Wp2^T + bp2;  hm = relu((x Wp1a^T)[col] + e1 Wp1b^T + bp1)
            dm = mk(ne)
            capi.segment_expand(dagg.data.data_ptr(), idx.rowptr.data_ptr(), n, True, dm.data.data_ptr(), code, s)
            dm.refresh_split()
            linear_grads(dm, hm, ne, phi[2].weight, 0, phi[2].bias)
            dhm = masked(dgrad(dm, w.bwd("wp2"), ne), hm, ne)
            del dm
            linear_grads(dhm, e1, ne, phi[0].weight, _H, phi[0].bias)
            de1_tot = dgrad(dhm, w.bwd("wp1b"), ne, residual=de1)
            dR = by_col(dhm)
            del dhm
            linear_grads(dR, x_in, n, phi[0].weight, 0, None)
            dx = dgrad(dR, w.bwd("wp1a"), n, residual=dx)
            # e1 = he W2^T + b2;  he = relu((x W1a^T)[row] + (x W1b^T)[col] + e W1c^T + b1)
            linear_grads(de1_tot, he, ne, em[2].weight, 0, em[2].bias)
            dhe = masked(dgrad(de1_tot, w.bwd("w2"), ne), he, ne)
            linear_grads(dhe, e_in, ne, em[0].weight, 2 * _H, em[0].bias)
            de_in = dgrad(dhe, w.bwd("w1c"), ne, residual=de1 if (skip and de1 is not None) else None)
            dP, dQ = by_row(dhe), by_col(dhe)
            del dhe, de1_tot
            linear_grads(dP, x_in, n, em[0].weight, 0, None)
            linear_grads(dQ, x_in, n, em[0].weight, _H, None)
            dx = dgrad(dP, w.bwd("w1a"), n, residual=dx)
            dx = dgrad(dQ, w.bwd("w1b"), n, residual=dx)
            dxn, den = dx, de_in
        encoder_backward(model.node_encoder, sv.x, sv.h1, sv.h2, dxn, prec, grads)
        encoder_backward(model.edge_encoder, sv.ea, sv.eh1, sv.eh2, den, prec, grads)
        ctx.sv = None
        return (None, None, None, None, None, None, *grads.for_params(ctx.params))
