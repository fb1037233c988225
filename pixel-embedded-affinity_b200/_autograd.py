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


def _refuse_grad(ctx, first, names):
    """The fused functions differentiate with respect to the embeddings only; the reference's torch ops would also
    propagate into target / weightmap / mask.  Nobody trains those, but a silent None would be wrong: refuse loudly."""
    for i, n in enumerate(names):
        if ctx.needs_input_grad[first + i]:
            raise RuntimeError(f"pea_b200: {n} requires grad, but the fused affinity loss is differentiable with respect to "
                               "the embeddings only (detach it, or use a criterion other than WeightedMSE)")


class FusedAffinityLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, partner, target, weight, mask, spec):
        _refuse_grad(ctx, 2, ("target", "weightmap", "mask"))
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
        # ONE call: mark_non_differentiable replaces the set of an earlier call
        ctx.mark_non_differentiable(*([each] if spec.affs_differentiable else [each, affs]))
        return loss, affs, each

    @staticmethod
    @torch.autograd.function.once_differentiable
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
        ctx.mark_non_differentiable(*([each] if spec.affs_differentiable else [each, affs]))
        return loss, affs, each

    @staticmethod
    @torch.autograd.function.once_differentiable
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


class FusedSelfAndCrossLoss(torch.autograd.Function):
    """The two full-resolution losses of a training step that share one embedding (scripts_ac3ac4/main.py:223-224): the
    one-view loss writes d/d embedding, the two-view loss (second view constant) ADDS its own into the same buffer
    (PEA_EPI_GRAD_ACCUMULATE) -- one gradient buffer and no accumulation pass by autograd.  The differentiable output is
    coef_self * L_self + coef_cross * L_cross (the coefficients are folded into the kernels' loss coefficients, so the
    buffer already holds the gradient of that sum); the two parts come back as plain values."""

    @staticmethod
    def forward(ctx, emb, partner, target, weight, spec_self, spec_cross):
        _refuse_grad(ctx, 1, ("ema_embedding", "target", "weightmap"))
        need = ctx.needs_input_grad[0]
        lo_s, affs, grad, _ = ops.affs_loss_fwd_bwd(emb, None, spec_self.offs3, spec_self.border, target, weight, None,
                                                    spec_self.loss_norm, spec_self.loss_coef, want_grad=need, impl=spec_self.impl)
        try:
            lo_c, affs_c, _, _ = ops.affs_loss_fwd_bwd(emb, partner, spec_cross.offs3, spec_cross.border, target, weight, None,
                                                       spec_cross.loss_norm, spec_cross.loss_coef, want_grad=need,
                                                       impl=spec_cross.impl, grad_into=grad if need else None)
        except ops.PeaError as ex:  # shape the tiled two-view kernel does not take: separate buffer, one add
            if ex.code != -8:
                raise
            lo_c, affs_c, gc, _ = ops.affs_loss_fwd_bwd(emb, partner, spec_cross.offs3, spec_cross.border, target, weight, None,
                                                        spec_cross.loss_norm, spec_cross.loss_coef, want_grad=need,
                                                        impl=spec_cross.impl)
            if need:
                grad.add_(gc)
        ctx.need, ctx.grad_buf = need, grad
        ctx.spec_self, ctx.spec_cross = spec_self, spec_cross
        if need:
            ctx.save_for_backward(emb, partner, target, weight)
        ctx.set_materialize_grads(False)
        total = lo_s[0] + lo_c[0]
        ctx.mark_non_differentiable(affs, affs_c, lo_s, lo_c)
        return total, affs, affs_c, lo_s, lo_c

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_total, *unused):
        if not ctx.need or g_total is None:
            return None, None, None, None, None, None
        ge = ctx.grad_buf
        if ge is None:  # handed out by an earlier backward (retain_graph=True): recompute both
            emb, partner, target, weight = ctx.saved_tensors
            ss, sc = ctx.spec_self, ctx.spec_cross
            _, _, ge, _ = ops.affs_loss_fwd_bwd(emb, None, ss.offs3, ss.border, target, weight, None, ss.loss_norm, ss.loss_coef,
                                                want_affs=False, want_grad=True, impl=ss.impl)
            _, _, gc, _ = ops.affs_loss_fwd_bwd(emb, partner, sc.offs3, sc.border, target, weight, None, sc.loss_norm, sc.loss_coef,
                                                want_affs=False, want_grad=True, impl=sc.impl)
            ge.add_(gc)
        ctx.grad_buf = None
        ops.scale_inplace(ge, g_total.to(device=ge.device, dtype=torch.float32))
        return ge, None, None, None, None, None


class MultiFusedAffinityLoss(torch.autograd.Function):
    """n small fused losses in ONE launch (ops.affs_loss_multi) and their gradient scaling in ONE launch.

    apply(specs, n, emb_0 .. emb_{n-1}, partner_0 .., target_0 .., weight_0 .., mask_0 ..) -> (loss_0, affs_0, each_0,
    loss_1, ...).  Partners / masks may be None; the second view is a constant here."""

    @staticmethod
    def forward(ctx, specs, n, *tensors):
        embs, pars, tgts, wgts, msks = (tensors[i * n:(i + 1) * n] for i in range(5))
        _refuse_grad(ctx, 2 + n, ["ema_embedding"] * n + ["target"] * n + ["weightmap"] * n + ["mask"] * n)
        need = [ctx.needs_input_grad[2 + i] for i in range(n)]
        problems = [dict(emb=embs[i], partner=pars[i], offs3=specs[i].offs3, border=specs[i].border, target=tgts[i],
                         weight=wgts[i], mask=msks[i], loss_norm=specs[i].loss_norm, loss_coef=specs[i].loss_coef,
                         want_grad=need[i]) for i in range(n)]
        outs = ops.affs_loss_multi(problems)
        if outs is None:  # no ticket word yet (first call on this device happens inside a graph capture): one by one
            outs = []
            for p in problems:
                lo, af, gr, _ = ops.affs_loss_fwd_bwd(p["emb"], p["partner"], p["offs3"], p["border"], p["target"], p["weight"],
                                                      p["mask"], p["loss_norm"], p["loss_coef"], want_grad=p["want_grad"])
                outs.append((lo, af, gr))
        ctx.n, ctx.specs, ctx.need = n, specs, need
        ctx.grads = [o[2] for o in outs]
        ctx.save_for_backward(*tensors)
        ctx.set_materialize_grads(False)
        flat, nondiff = [], []
        for i, (loss_out, affs, _) in enumerate(outs):
            each = loss_out[1:]
            nondiff.append(each)
            if not specs[i].affs_differentiable:
                nondiff.append(affs)
            flat += [loss_out[0], affs, each]
        ctx.mark_non_differentiable(*nondiff)  # ONE call: a second call would replace the set
        return tuple(flat)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        n, specs = ctx.n, ctx.specs
        tensors = ctx.saved_tensors
        embs, pars, tgts, wgts, msks = (tensors[i * n:(i + 1) * n] for i in range(5))
        out = [None] * n
        bufs, scales = [], []
        for i in range(n):
            g_loss = gs[3 * i]
            if not ctx.need[i] or g_loss is None:
                continue
            ge = ctx.grads[i]
            if ge is None:  # handed out by an earlier backward (retain_graph=True): recompute
                _, _, ge, _ = ops.affs_loss_fwd_bwd(embs[i], pars[i], specs[i].offs3, specs[i].border, tgts[i], wgts[i], msks[i],
                                                    specs[i].loss_norm, specs[i].loss_coef, want_affs=False, want_grad=True)
            ctx.grads[i] = None
            out[i] = ge
            bufs.append(ge)
            scales.append(g_loss.to(device=ge.device, dtype=torch.float32))
        if bufs:  # one launch applies every upstream scalar
            ops.scale_multi(bufs, scales)
        for i in range(n):  # gradients arriving through the returned affinities (2-D: they carry grad in the reference)
            g_affs = gs[3 * i + 1]
            if ctx.need[i] and g_affs is not None:
                ae, _ = ops.affs_bwd(embs[i], pars[i], specs[i].offs3, specs[i].border, g_affs.contiguous(), want_grad=True)
                out[i] = ae if out[i] is None else out[i].add_(ae)
        return (None, None) + tuple(out) + (None,) * (4 * n)


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
