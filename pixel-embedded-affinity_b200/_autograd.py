"""torch.autograd.Function shells around the fused CUDA ops.

Forward launches ONE fused pass that already produces d loss / d embedding for an upstream gradient of 1
(the upstream gradient of a scalar loss is a scalar, so the backward is known up to that factor);
backward rescales it in place by the incoming scalar (a kernel that exits immediately when the scalar is
exactly 1) and, for the 2-D functions whose returned affinities carry grad in the reference
(scripts_cvppp/loss/loss_embedding_mse.py:46), adds the gradient that arrives through ``affs``.
No host synchronisation anywhere: CUDA-graph capturable, DDP-safe (no collectives, no global state).
"""
import collections.abc

import numpy as np
import torch

from . import _ops as ops


class OpSpec:
    """Static description of one fused call (everything that is not a tensor)."""

    def __init__(self, offs3, border, loss_norm, loss_coef, affs_differentiable, impl=None):
        self.offs3 = np.ascontiguousarray(offs3, dtype=np.int32)
        self.border = border
        self.loss_norm = np.asarray(loss_norm, dtype=np.float32)
        self.loss_coef = np.asarray(loss_coef, dtype=np.float32)
        self.affs_differentiable = affs_differentiable
        self.impl = ops.DEFAULT_IMPL if impl is None else impl


class LossList(collections.abc.Sequence):
    """The reference's ``all_loss`` (a list of K Python floats, one ``.item()`` sync per offset).  Here the K
    values stay on the device and are fetched with a single copy the first time the list is read."""

    def __init__(self, device_values):
        self._dev = device_values
        self._host = None

    def _get(self):
        if self._host is None:
            self._host = [float(v) for v in self._dev.detach().cpu().tolist()]
        return self._host

    def __len__(self):
        return int(self._dev.numel())

    def __getitem__(self, i):
        return self._get()[i]

    def __repr__(self):
        return repr(self._get())

    def __eq__(self, other):
        return list(self) == list(other)

    @property
    def device_tensor(self):
        return self._dev


class FusedAffinityLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, partner, target, weight, mask, spec):
        need_e = ctx.needs_input_grad[0]
        need_u = partner is not None and ctx.needs_input_grad[1]
        loss_out, affs, grad, gpar = ops.affs_loss_fwd_bwd(
            emb, partner, spec.offs3, spec.border, target, weight, mask, spec.loss_norm, spec.loss_coef,
            want_affs=True, want_grad=need_e, want_grad_partner=need_u, impl=spec.impl)
        ctx.spec = spec
        ctx.grad_buf, ctx.gpar_buf = grad, gpar
        ctx.need = (need_e, need_u)
        ctx.has_partner = partner is not None
        if need_e or need_u:
            # inputs are kept (by reference, no copies) for the affs-gradient path and for a repeated backward
            ctx.save_for_backward(emb, partner, target, weight, mask)
        ctx.set_materialize_grads(False)
        loss = loss_out[0]
        each = loss_out[1:]
        ctx.mark_non_differentiable(each)
        if not spec.affs_differentiable:
            ctx.mark_non_differentiable(affs)
        return loss, affs, each

    @staticmethod
    def backward(ctx, g_loss, g_affs, g_each):
        need_e, need_u = ctx.need
        spec = ctx.spec
        emb, partner, target, weight, mask = ctx.saved_tensors
        ge = gu = None
        if g_loss is not None:
            ge, gu = ctx.grad_buf, ctx.gpar_buf
            if (need_e and ge is None) or (need_u and gu is None):
                # buffers already handed out by an earlier backward (retain_graph=True): recompute
                _, _, ge, gu = ops.affs_loss_fwd_bwd(emb, partner, spec.offs3, spec.border, target, weight, mask,
                                                     spec.loss_norm, spec.loss_coef, want_affs=False, want_grad=need_e,
                                                     want_grad_partner=need_u, impl=spec.impl)
            ctx.grad_buf = ctx.gpar_buf = None
            g = g_loss.to(device=emb.device, dtype=torch.float32)
            if ge is not None:
                ops.scale_inplace(ge, g)
            if gu is not None:
                ops.scale_inplace(gu, g)
        if g_affs is not None and (need_e or need_u):
            ae, au = ops.affs_bwd(emb, partner, spec.offs3, spec.border, g_affs.contiguous(), want_grad=need_e,
                                  want_grad_partner=need_u)
            ge = ae if ge is None else (ge.add_(ae) if ae is not None else ge)
            gu = au if gu is None else (gu.add_(au) if au is not None else gu)
        return (ge if need_e else None), (gu if need_u else None), None, None, None, None


class FusedAffinityLossLabels(torch.autograd.Function):
    """embedding + instance labels -> loss, affinities, per-offset losses: target / mask / weight are derived inside
    the kernel (canonical 2-D tables; ``ops.affs_loss_labels_fwd_bwd`` raises PEA_E_MODE otherwise)."""

    @staticmethod
    def forward(ctx, emb, labels_i32, wtab, ignore, spec, semantics=0):
        need_e = ctx.needs_input_grad[0]
        loss_out, affs, grad = ops.affs_loss_labels_fwd_bwd(emb, labels_i32, spec.offs3, spec.border, ignore, wtab,
                                                            spec.loss_norm, spec.loss_coef, want_affs=True, want_grad=need_e,
                                                            semantics=semantics)
        ctx.spec, ctx.ignore, ctx.grad_buf, ctx.need_e, ctx.semantics = spec, ignore, grad, need_e, semantics
        if need_e:
            ctx.save_for_backward(emb, labels_i32, wtab)
        ctx.set_materialize_grads(False)
        loss, each = loss_out[0], loss_out[1:]
        ctx.mark_non_differentiable(each)
        if not spec.affs_differentiable:
            ctx.mark_non_differentiable(affs)
        return loss, affs, each

    @staticmethod
    def backward(ctx, g_loss, g_affs, g_each):
        if not ctx.need_e:
            return None, None, None, None, None, None
        emb, labels_i32, wtab = ctx.saved_tensors
        spec = ctx.spec
        ge = None
        if g_loss is not None:
            ge = ctx.grad_buf
            if ge is None:  # buffer already handed out by an earlier backward (retain_graph=True): recompute
                _, _, ge = ops.affs_loss_labels_fwd_bwd(emb, labels_i32, spec.offs3, spec.border, ctx.ignore, wtab,
                                                        spec.loss_norm, spec.loss_coef, want_affs=False, want_grad=True,
                                                        semantics=ctx.semantics)
            ctx.grad_buf = None
            ops.scale_inplace(ge, g_loss.to(device=emb.device, dtype=torch.float32))
        if g_affs is not None:
            ae, _ = ops.affs_bwd(emb, None, spec.offs3, spec.border, g_affs.contiguous(), want_grad=True)
            ge = ae if ge is None else ge.add_(ae)
        return ge, None, None, None, None, None


class Affinities(torch.autograd.Function):
    """embedding (and optional partner) -> affinities, with the gather backward (SURVEY K1/K3 + K5)."""

    @staticmethod
    def forward(ctx, emb, partner, offs3, border, impl):
        affs = ops.affs_fwd(emb, partner, offs3, border, impl)
        ctx.offs3, ctx.border = offs3, border
        ctx.need = (ctx.needs_input_grad[0], partner is not None and ctx.needs_input_grad[1])
        if any(ctx.need):
            ctx.save_for_backward(emb, partner)
        return affs

    @staticmethod
    def backward(ctx, g_affs):
        need_e, need_u = ctx.need
        emb, partner = ctx.saved_tensors
        ge, gu = ops.affs_bwd(emb, partner, ctx.offs3, ctx.border, g_affs.contiguous(), want_grad=need_e,
                              want_grad_partner=need_u)
        return ge, gu, None, None, None
