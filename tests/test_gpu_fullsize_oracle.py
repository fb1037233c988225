"""GPU parity at BASELINE.json's FULL sizes against the CPU ORACLE (oracle/pea_oracle.py, itself pinned bit for bit to the
reference by tests/golden): cfg1 (B=1, 544x512, shifts up to 27), cfg2 (B=8, 520x696), cfg3 (B=4, 18x160x160), one view,
two views and the label-driven flavours.  The oracle needs 0.2-3 s per call on the box's host cores.

Tolerance (BASELINE.json north_star: 1e-5 relative in fp32), written as SURVEY.md 7.4-6 prescribes:
    |ours - ref| <= 1e-5 * |ref| + 1e-6 * max|ref|          element by element
Ground-truth targets, masks and weights: bit-exact.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL_REL = 1e-5, 1e-6
C = 16


def close(ours, ref, what):
    ours = ours.detach().cpu().double() if isinstance(ours, torch.Tensor) else torch.as_tensor(np.asarray(ours)).double()
    ref = ref.detach().cpu().double() if isinstance(ref, torch.Tensor) else torch.as_tensor(np.asarray(ref)).double()
    assert ours.shape == ref.shape, (what, tuple(ours.shape), tuple(ref.shape))
    scale = float(ref.abs().max())
    excess = (ours - ref).abs() - (RTOL * ref.abs() + ATOL_REL * scale)
    worst = float(excess.max())
    assert worst <= 0.0, (f"{what}: exceeds rtol {RTOL:g} + atol {ATOL_REL:g} * {scale:.3e} by {worst:.3e} "
                          f"(max |delta| {float((ours - ref).abs().max()):.3e})")


def close_s(ours, ref, what):
    ours, ref = float(ours), float(ref)
    assert abs(ours - ref) <= RTOL * max(abs(ref), 1e-30), f"{what}: {ours!r} vs {ref!r}"


@pytest.fixture(scope="module")
def P():
    import pea_b200
    return pea_b200


@pytest.fixture(scope="module")
def O():
    from oracle import pea_oracle
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    return pea_oracle


def _inputs_2d(O, B, H, W, shifts, seed):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn((B, C, H, W), generator=g)
    offsets = O.multi_offset(shifts, 4)
    lab = np.stack([O.voronoi_labels((H, W), 20, seed * 100 + b, background_frac=0.3, dtype=np.uint8) for b in range(B)])
    pairs = [O.gen_affs_ours(l, offsets, ignore=False, padding=True) for l in lab]
    tgt = np.stack([p[0] for p in pairs])
    msk = np.stack([p[1] for p in pairs])
    wgt = np.stack([np.stack([O.weight_binary_ratio(c) for c in t]) for t in tgt])
    return emb, lab, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), offsets


def _inputs_3d(O, B, D, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn((B, C, D, H, W), generator=g)
    lab = np.stack([O.voronoi_labels((D, H, W), 60, seed * 100 + b, background_frac=0.1, z_scale=8.0, dtype=np.uint16)
                    for b in range(B)])
    tgt = np.stack([O.seg_to_aff_norm5(l) for l in lab])
    wgt = np.stack([np.stack([O.weight_binary_ratio(c) for c in t]) for t in tgt])
    return emb, lab, torch.tensor(tgt), torch.tensor(wgt)


def _check_2d(P, O, B, H, W, shifts, seed, name):
    emb, lab, tgt, wgt, msk, offsets = _inputs_2d(O, B, H, W, shifts, seed)
    crit_o, crit = O.WeightedMSE(), P.WeightedMSE()
    # ---- one view
    eo = emb.clone().requires_grad_(True)
    lo, ao, allo = O.embedding_loss(eo, tgt, wgt, msk, crit_o, offsets)
    lo.backward()
    e = emb.cuda().requires_grad_(True)
    l, a, alll = P.embedding_loss(e, tgt.cuda(), wgt.cuda(), msk.cuda(), crit, offsets)
    l.backward()
    close_s(l, lo, f"{name} loss")
    for k in range(len(offsets)):
        close_s(alll[k], allo[k], f"{name} all_loss[{k}]")
    close(a, ao, f"{name} affs")
    close(e.grad, eo.grad, f"{name} grad")
    close(P.embedding2affs(emb.cuda(), offsets), ao, f"{name} embedding2affs")
    # ---- device-side GT is the oracle's, bit for bit; label-driven loss == tensor-driven loss
    t2, m2, w2 = P.make_targets_2d(torch.tensor(lab).cuda(), offsets)
    assert torch.equal(t2.cpu(), tgt) and torch.equal(m2.cpu(), msk) and torch.equal(w2.cpu(), wgt), f"{name} GT"
    return emb, lab, tgt, wgt, msk, offsets, (lo, ao, eo.grad)


def test_cfg1_full_size_vs_oracle(P, O):
    """CVPPP shape: shift 27 keeps the canonical kernel out, the table-driven axis kernel (PMAX = 28 flavour) runs."""
    _check_2d(P, O, 1, 544, 512, [1, 3, 5, 9, 27], 11, "cfg1")


def test_cfg2_full_size_vs_oracle(P, O):
    emb, lab, tgt, wgt, msk, offsets, (lo, ao, go) = _check_2d(P, O, 8, 520, 696, [1, 3, 5, 9, 11], 12, "cfg2")
    # label-driven flavour (GT never materialised) against the same oracle numbers
    e = emb.cuda().requires_grad_(True)
    l, a, _ = P.embedding_loss_labels(e, torch.tensor(lab).cuda(), P.WeightedMSE(), offsets)
    l.backward()
    close_s(l, lo, "cfg2 labels loss")
    close(a, ao, "cfg2 labels affs")
    close(e.grad, go, "cfg2 labels grad")


def test_cfg2x_full_size_two_views_vs_oracle(P, O):
    emb, lab, tgt, wgt, msk, offsets = _inputs_2d(O, 8, 520, 696, [1, 3, 5, 9, 11], 13)
    u = torch.randn(emb.shape, generator=torch.Generator().manual_seed(99))
    eo = emb.clone().requires_grad_(True)
    lo, ao = O.ema_embedding_loss(eo, u, tgt, wgt, msk, O.WeightedMSE(), offsets, affs0_weight=2)
    lo.backward()
    e = emb.cuda().requires_grad_(True)
    l, a = P.ema_embedding_loss(e, u.cuda(), tgt.cuda(), wgt.cuda(), msk.cuda(), P.WeightedMSE(), offsets, affs0_weight=2)
    l.backward()
    close_s(l, lo, "cfg2x loss")
    close(a, ao, "cfg2x affs")
    close(e.grad, eo.grad, "cfg2x grad")


def test_cfg3_full_size_vs_oracle(P, O):
    B, D, H, W = 4, 18, 160, 160
    emb, lab, tgt, wgt = _inputs_3d(O, B, D, H, W, 14)
    eo = emb.clone().requires_grad_(True)
    lo, ao = O.embedding_loss_norm5(eo, tgt, wgt, O.WeightedMSE(), affs0_weight=1.5)
    lo.backward()
    e = emb.cuda().requires_grad_(True)
    l, a = P.embedding_loss_norm5(e, tgt.cuda(), wgt.cuda(), P.WeightedMSE(), affs0_weight=1.5)
    l.backward()
    close_s(l, lo, "cfg3 loss")
    close(a, ao, "cfg3 affs")
    close(e.grad, eo.grad, "cfg3 grad")
    # forward-only entry point and the inference kernel
    with torch.no_grad():
        l2, a2 = P.embedding_loss_norm5(emb.cuda(), tgt.cuda(), wgt.cuda(), P.WeightedMSE(), affs0_weight=1.5)
    close_s(l2, lo, "cfg3 loss (no grad)")
    close(a2, ao, "cfg3 affs (no grad)")
    close(P.inf_embedding_loss_norm5(emb.cuda()), O.inf_embedding_loss_norm5(emb), "cfg3 inference affs")
    # device-side GT == oracle GT bit for bit, label-driven loss == the same oracle numbers
    t2, w2 = P.make_targets_3d(torch.tensor(lab.astype(np.int32)).cuda())
    assert torch.equal(t2.cpu(), tgt) and torch.equal(w2.cpu(), wgt), "cfg3 GT"
    e = emb.cuda().requires_grad_(True)
    l, a = P.embedding_loss_norm5_labels(e, torch.tensor(lab.astype(np.int32)).cuda(), P.WeightedMSE(), affs0_weight=1.5)
    l.backward()
    close_s(l, lo, "cfg3 labels loss")
    close(a, ao, "cfg3 labels affs")
    close(e.grad, eo.grad, "cfg3 labels grad")


def test_cfg3x_full_size_two_views_vs_oracle(P, O):
    B, D, H, W = 4, 18, 160, 160
    emb, lab, tgt, wgt = _inputs_3d(O, B, D, H, W, 15)
    u = torch.randn(emb.shape, generator=torch.Generator().manual_seed(98))
    eo = emb.clone().requires_grad_(True)
    lo, ao = O.ema_embedding_loss_norm5(eo, u, tgt, wgt, O.WeightedMSE(), affs0_weight=2)
    lo.backward()
    e = emb.cuda().requires_grad_(True)
    l, a = P.ema_embedding_loss_norm5(e, u.cuda(), tgt.cuda(), wgt.cuda(), P.WeightedMSE(), affs0_weight=2)
    l.backward()
    close_s(l, lo, "cfg3x loss")
    close(a, ao, "cfg3x affs")
    close(e.grad, eo.grad, "cfg3x grad")
