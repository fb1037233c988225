"""GPU parity: the CUDA path (through the C ABI) against (1) the committed golden vectors produced by the real
reference and (2) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): ground-truth affinities, masks and weights bit-exact; affinities, loss and
embedding gradients within 1e-5 relative in fp32, written as SURVEY.md 7.4-6 prescribes, element by element:
    |ours - ref| <= rtol * |ref| + (rtol / 10) * max|ref|        rtol = 1e-5, i.e. atol = 1e-6 * scale
(small-magnitude elements are held to a tenth of what a pure inf-norm test would allow them); scalar losses use plain
relative error.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def dev(a, grad=False):
    t = torch.tensor(np.asarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


def close_t(ours, ref, rtol=RTOL, what=""):
    ours = ours.detach().cpu().numpy() if isinstance(ours, torch.Tensor) else np.asarray(ours)
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref)
    assert ours.shape == ref.shape, (what, ours.shape, ref.shape)
    if ours.size == 0:
        return
    scale = max(float(np.abs(ref).max()), 1e-30)
    ref64 = ref.astype(np.float64)
    delta = np.abs(ours.astype(np.float64) - ref64)
    excess = float((delta - (rtol * np.abs(ref64) + 0.1 * rtol * scale)).max())
    assert excess <= 0.0, (f"{what}: |delta| exceeds {rtol:g} * |ref| + {0.1 * rtol:g} * {scale:.3e} by {excess:.3e} "
                           f"(max |delta| {float(delta.max()):.3e})")


def close_s(ours, ref, rtol=RTOL, what=""):
    ours, ref = float(ours), float(ref)
    assert abs(ours - ref) <= rtol * max(abs(ref), 1e-30), f"{what}: {ours!r} vs {ref!r}"


def same(ours, ref, what=""):
    ours = ours.detach().cpu().numpy() if isinstance(ours, torch.Tensor) else np.asarray(ours)
    assert ours.dtype == np.asarray(ref).dtype, (what, ours.dtype, np.asarray(ref).dtype)
    np.testing.assert_array_equal(ours, np.asarray(ref), err_msg=what)


@pytest.fixture(scope="module")
def P():
    import pea_b200
    return pea_b200


@pytest.fixture(scope="module")
def O():
    from oracle import pea_oracle
    return pea_oracle


# =========================================================================== golden vectors: 2D
def test_golden_embedding_loss_2d(P, golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    e = dev(g["embedding"], True)
    loss, affs, all_loss = P.embedding_loss(e, dev(g["target"]), dev(g["weightmap"]), dev(g["mask"]), P.WeightedMSE(),
                                            offsets, affs0_weight=3)
    assert loss.dim() == 0 and loss.dtype == torch.float32 and affs.requires_grad
    loss.backward()
    close_s(loss, g["loss"], what="loss")
    close_t(affs, g["affs"], what="affs")
    assert len(all_loss) == len(offsets)
    for k in range(len(offsets)):
        close_s(all_loss[k], g["all_loss"][k], what=f"all_loss[{k}]")
    # the golden batch contains one exact-zero pixel (F.normalize eps branch, gradient ~1e12 * G): check it apart
    zero = np.broadcast_to((g["embedding"] == 0).all(axis=1, keepdims=True), g["grad"].shape)
    ours = e.grad.cpu().numpy()
    close_t(np.where(zero, 0, ours), np.where(zero, 0, g["grad"]), what="grad")
    close_t(ours[zero], g["grad"][zero], rtol=1e-4, what="grad at the zero-norm pixel")
    close_t(P.embedding2affs(dev(g["embedding"]), offsets), g["embedding2affs"], what="embedding2affs")


def test_golden_embedding_loss_2d_float_mask_and_affs_grad(P, golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    crit = P.WeightedMSE()
    loss, _, _ = P.embedding_loss(dev(g["embedding"]), dev(g["target"]), dev(g["weightmap"]), dev(g["mask"]).float(),
                                  crit, offsets)
    close_s(loss, g["loss_fmask"], what="loss with float mask")
    e = dev(g["embedding"], True)
    loss, affs, _ = P.embedding_loss(e, dev(g["target"]), dev(g["weightmap"]), dev(g["mask"]), crit, offsets)
    (0.5 * loss + (affs * dev(g["affs_upstream"])).sum()).backward()
    zero = np.broadcast_to((g["embedding"] == 0).all(axis=1, keepdims=True), g["grad"].shape)
    close_t(np.where(zero, 0, e.grad.cpu().numpy()), np.where(zero, 0, g["grad_with_affs"]), what="grad through loss+affs")


def test_golden_ema_embedding_loss_2d(P, golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    e, u = dev(g["embedding"], True), dev(g["ema_embedding"], True)
    loss, affs = P.ema_embedding_loss(e, u, dev(g["target"]), dev(g["weightmap"]), dev(g["mask"]), P.WeightedMSE(),
                                      offsets, affs0_weight=3)
    loss.backward()
    close_s(loss, g["ema_loss"], what="ema loss")
    close_t(affs, g["ema_affs"], what="ema affs")
    zero = np.broadcast_to((g["embedding"] == 0).all(axis=1, keepdims=True), g["grad"].shape)
    close_t(np.where(zero, 0, e.grad.cpu().numpy()), np.where(zero, 0, g["ema_grad"]), what="ema grad")
    close_t(u.grad, g["ema_grad_partner"], what="ema grad wrt partner")


# =========================================================================== golden vectors: 3D
def test_golden_norm5_family(P, golden):
    g = golden["loss3d"]
    crit = P.WeightedMSE()
    e = dev(g["A_embedding"], True)
    loss, affs = P.embedding_loss_norm5(e, dev(g["A_target"]), dev(g["A_weightmap"]), crit, affs0_weight=2)
    assert not affs.requires_grad
    loss.backward()
    close_s(loss, g["A_n5_loss"], what="norm5 loss")
    close_t(affs, g["A_n5_affs"], what="norm5 affs")
    close_t(e.grad, g["A_n5_grad"], what="norm5 grad")

    e, u = dev(g["A_embedding"], True), dev(g["A_ema_embedding"], True)
    loss, affs = P.ema_embedding_loss_norm5(e, u, dev(g["A_target"]), dev(g["A_weightmap"]), crit, affs0_weight=2)
    loss.backward()
    close_s(loss, g["A_ema5_loss"], what="ema5 loss")
    close_t(affs, g["A_ema5_affs"], what="ema5 affs")
    close_t(e.grad, g["A_ema5_grad"], what="ema5 grad")
    close_t(u.grad, g["A_ema5_grad_partner"], what="ema5 grad wrt partner")
    close_t(P.inf_embedding_loss_norm5(dev(g["A_embedding"])), g["A_inf5"], what="inf5")


def test_golden_norm1_family(P, golden):
    g = golden["loss3d"]
    crit = P.WeightedMSE()
    T, W = dev(g["B_target"]), dev(g["B_weightmap"])
    e = dev(g["B_embedding"], True)
    loss, affs = P.embedding_loss_norm1(e, T, W, crit, affs0_weight=2)
    loss.backward()
    close_s(loss, g["B_n1_loss"], what="norm1 loss")
    close_t(affs, g["B_n1_affs"], what="norm1 affs")
    close_t(e.grad, g["B_n1_grad"], what="norm1 grad")

    e = dev(g["B_embedding"], True)
    loss, affs = P.embedding_loss_norm1(e, T, W, crit, affs0_weight=0.5, shift=2)
    loss.backward()
    close_s(loss, g["B_n1s2_loss"], what="norm1 shift 2 loss")
    close_t(affs, g["B_n1s2_affs"], what="norm1 shift 2 affs")
    close_t(e.grad, g["B_n1s2_grad"], what="norm1 shift 2 grad")

    e, u = dev(g["B_embedding"], True), dev(g["B_ema_embedding"], True)
    loss, affs = P.ema_embedding_loss_norm1(e, u, T, W, crit, affs0_weight=2)
    loss.backward()
    close_s(loss, g["B_ema1_loss"], what="ema1 loss")
    close_t(affs, g["B_ema1_affs"], what="ema1 affs")
    close_t(e.grad, g["B_ema1_grad"], what="ema1 grad")
    close_t(u.grad, g["B_ema1_grad_partner"], what="ema1 grad wrt partner")
    close_t(P.inf_embedding_loss_norm1(dev(g["B_embedding"])), g["B_inf1"], what="inf1")


# =========================================================================== golden vectors: ground truth (bit exact)
def test_golden_gen_affs_ours(P, golden):
    g = golden["gt"]
    lab, offsets = g["gt2d_labels"], g["gt2d_offsets"].tolist()
    for ig in (False, True):
        for pd in (False, True):
            a, m = P.gen_affs_ours(lab, offsets, ignore=ig, padding=pd)            # NumPy in -> NumPy out
            assert isinstance(a, np.ndarray) and isinstance(m, np.ndarray)
            same(a, g[f"gt2d_aff_i{int(ig)}_p{int(pd)}"], "aff")
            same(m, g[f"gt2d_msk_i{int(ig)}_p{int(pd)}"], "mask")
            a, m = P.gen_affs_ours(torch.tensor(lab).cuda(), offsets, ignore=ig, padding=pd)  # tensor in -> tensor out
            assert a.is_cuda and m.is_cuda
            same(a, g[f"gt2d_aff_i{int(ig)}_p{int(pd)}"], "aff (device)")
            same(m, g[f"gt2d_msk_i{int(ig)}_p{int(pd)}"], "mask (device)")
    a, m = P.gen_affs_ours(lab.astype(np.float32), offsets, ignore=False, padding=True)
    same(a, g["gt2d_aff_float_labels"])
    same(m, g["gt2d_msk_float_labels"].astype(np.uint8))
    for dt in (np.int16, np.int32, np.int64, np.uint16):
        a, m = P.gen_affs_ours(lab.astype(dt), offsets, ignore=True, padding=False)
        same(a, g["gt2d_aff_i1_p0"], f"labels as {dt.__name__}")


def test_golden_seg_to_aff(P, golden):
    g = golden["gt"]
    lab3 = g["gt3d_labels"]
    from pea_b200.data_affinity import NHOOD_NORM5
    same(P.seg_to_aff(lab3, NHOOD_NORM5, pad=''), g["gt3d_aff12"], "12-channel")
    for i in range(4):
        same(P.seg_to_aff(lab3, NHOOD_NORM5[3 * i:3 * i + 3], pad=''), g["gt3d_aff12"][3 * i:3 * i + 3], f"nhood {i}")
    same(P.seg_to_aff(lab3), g["gt3d_aff3_replicate"], "replicate")
    same(P.seg_to_aff(lab3, pad=''), g["gt3d_aff3_nopad"], "no pad")
    same(P.seg_to_aff(g["gt2d_labels"], np.asarray([[-1, 0], [0, -1]])), g["gt3d_aff_2dseg_replicate"], "2-D seg")
    same(P.seg_to_aff(torch.tensor(lab3.astype(np.int32)).cuda()), g["gt3d_aff3_replicate"], "device int32")
    same(P.seg_to_aff(lab3.astype(np.int64)), g["gt3d_aff3_replicate"], "int64")


def test_golden_weight_binary_ratio(P, golden):
    g = golden["gt"]
    for cin, cout in zip(g["wbr_in"], g["wbr_out"]):
        same(P.weight_binary_ratio(cin), cout)
    same(np.stack([P.weight_binary_ratio(c) for c in g["gt3d_aff12"]]), g["gt3d_w12"])
    same(np.stack([P.weight_binary_ratio(c) for c in g["gt2d_aff_i0_p1"]]), g["gt2d_w"])
    # batched device path: one problem per (sample, channel)
    lab3 = torch.tensor(g["gt3d_labels"].astype(np.int32)).cuda()
    tgt, w = P.make_targets_3d(torch.stack([lab3, lab3]))
    same(tgt[1], g["gt3d_aff12"])
    same(w[0], g["gt3d_w12"])
    same(w[1], g["gt3d_w12"])
    lab2 = torch.tensor(g["gt2d_labels"]).cuda()
    tgt, msk, w = P.make_targets_2d(lab2[None], g["gt2d_offsets"].tolist())
    same(tgt[0], g["gt2d_aff_i0_p1"])
    same(msk[0], g["gt2d_msk_i0_p1"])
    same(w[0], g["gt2d_w"])


# =========================================================================== CUDA vs oracle on seeded inputs
def _inputs_2d(O, B, C, H, W, offsets, seed, float_mask=False):
    rng = np.random.RandomState(seed)
    emb = rng.randn(B, C, H, W).astype(np.float32) * rng.uniform(0.2, 3.0)
    ema = rng.randn(B, C, H, W).astype(np.float32)
    K = len(offsets)
    tgt = np.zeros((B, K, H, W), np.float32)
    msk = np.zeros((B, K, H, W), np.uint8)
    wgt = np.zeros((B, K, H, W), np.float32)
    for b in range(B):
        lab = O.voronoi_labels((H, W), 9, seed * 10 + b)
        tgt[b], msk[b] = O.gen_affs_ours(lab, offsets, ignore=False, padding=True)
        wgt[b] = np.stack([O.weight_binary_ratio(c) for c in tgt[b]])
    return emb, ema, tgt, wgt, (msk.astype(np.float32) if float_mask else msk)


@pytest.mark.parametrize("B,C,H,W,shifts,nb,seed", [
    (2, 16, 37, 53, [1, 3, 5, 9, 27], 4, 1),      # odd extents, CVPPP offsets
    (1, 16, 40, 48, [1, 3, 5, 9, 11], 8, 2),      # BBBC offsets with diagonals
    (3, 8, 16, 16, [1, 20], 4, 3),                # shift larger than the image (wraps modulo)
    (1, 4, 9, 130, [2], 4, 4),
    (2, 32, 21, 33, [1, 4], 8, 5),
])
def test_oracle_2d(P, O, B, C, H, W, shifts, nb, seed):
    offsets = O.multi_offset(shifts, neighbor=nb)
    emb, ema, tgt, wgt, msk = _inputs_2d(O, B, C, H, W, offsets, seed, float_mask=(seed % 2 == 0))
    # oracle (CPU)
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao, allo = O.embedding_loss(eo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), O.WeightedMSE(), offsets)
    lo.backward()
    # CUDA
    e = dev(emb, True)
    l, a, al = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    l.backward()
    close_s(l, lo, what="loss")
    close_t(a, ao, what="affs")
    close_t(np.asarray(list(al)), np.asarray(allo), what="all_loss")
    close_t(e.grad, eo.grad, what="grad")
    # cross variant, gradient wrt both views
    eo, uo = torch.tensor(emb, requires_grad=True), torch.tensor(ema, requires_grad=True)
    lo, ao = O.ema_embedding_loss(eo, uo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), O.WeightedMSE(),
                                  offsets, affs0_weight=0.25)
    lo.backward()
    e, u = dev(emb, True), dev(ema, True)
    l, a = P.ema_embedding_loss(e, u, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets, affs0_weight=0.25)
    l.backward()
    close_s(l, lo, what="ema loss")
    close_t(a, ao, what="ema affs")
    close_t(e.grad, eo.grad, what="ema grad")
    close_t(u.grad, uo.grad, what="ema grad partner")
    # detached partner (the shipped training loops, data_consistency.py:36): only the first view gets a gradient
    e = dev(emb, True)
    l, _ = P.ema_embedding_loss(e, dev(ema), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets, affs0_weight=0.25)
    (3.0 * l).backward()
    close_t(e.grad, 3.0 * eo.grad, what="ema grad, scaled upstream")


def _criteria(P):
    """(name, oracle-side callable, drop-in criterion) for every non-fused criterion the reference ships
    (scripts_*/loss/loss.py:126-152).  The BCE flavours need predictions in [0, 1]: callers feed embeddings with
    non-negative components, whose cosine affinities are."""
    F = torch.nn.functional
    return (("MSELoss", lambda p, t, w: F.mse_loss(p, t), P.MSELoss()),
            ("BCELoss", lambda p, t, w: F.binary_cross_entropy(p, t), P.BCELoss()),
            ("WeightedBCE", lambda p, t, w: F.binary_cross_entropy(p, t, w), P.WeightedBCE()))


def test_non_fused_criterion_2d(P, O):
    """Criteria other than WeightedMSE run the CUDA affinity op + the criterion in PyTorch, like the reference loop;
    the gradient comes back through pea_affs_bwd (K5)."""
    offsets = O.multi_offset([1, 3], neighbor=4)
    emb, ema, tgt, wgt, msk = _inputs_2d(O, 2, 16, 20, 24, offsets, 7)
    emb, ema = np.abs(emb) + 0.05, np.abs(ema) + 0.05
    for name, crit_o, crit_p in _criteria(P):
        eo = torch.tensor(emb, requires_grad=True)
        lo, ao, allo = O.embedding_loss(eo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), crit_o, offsets)
        lo.backward()
        e = dev(emb, True)
        l, a, al = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), crit_p, offsets)
        l.backward()
        close_s(l, lo, what=f"{name} loss")
        close_t(a, ao, what=f"{name} affs")
        close_t(e.grad, eo.grad, what=f"{name} grad")
        close_t(np.asarray(list(al)), np.asarray(allo), what=f"{name} all_loss")
        # two views, gradient to both
        eo, uo = torch.tensor(emb, requires_grad=True), torch.tensor(ema, requires_grad=True)
        lo, ao = O.ema_embedding_loss(eo, uo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), crit_o, offsets,
                                      affs0_weight=2)
        lo.backward()
        e, u = dev(emb, True), dev(ema, True)
        l, a = P.ema_embedding_loss(e, u, dev(tgt), dev(wgt), dev(msk), crit_p, offsets, affs0_weight=2)
        l.backward()
        close_s(l, lo, what=f"{name} ema loss")
        close_t(e.grad, eo.grad, what=f"{name} ema grad")
        close_t(u.grad, uo.grad, what=f"{name} ema grad wrt partner")


def _inputs_3d(O, B, C, D, H, W, K, seed, replicate=False):
    rng = np.random.RandomState(seed)
    emb = rng.randn(B, C, D, H, W).astype(np.float32) * rng.uniform(0.2, 3.0)
    ema = rng.randn(B, C, D, H, W).astype(np.float32)
    tgt = np.zeros((B, K, D, H, W), np.float32)
    wgt = np.zeros_like(tgt)
    for b in range(B):
        lab = O.voronoi_labels((D, H, W), 12, seed * 10 + b, z_scale=4.0, dtype=np.uint16)
        tgt[b] = O.seg_to_aff_norm5(lab) if K == 12 else O.seg_to_aff(lab, pad='replicate' if replicate else '')
        wgt[b] = np.stack([O.weight_binary_ratio(c) for c in tgt[b]])
    return emb, ema, tgt, wgt


def test_non_fused_criterion_3d(P, O):
    """3-D route for criteria other than WeightedMSE (scripts_ac3ac4/loss/loss_embedding_mse.py:159-163): CUDA affinity
    op (tiled forward kernel for the 12-offset set, table-driven otherwise), per-channel cropped criterion calls in
    PyTorch, gradient through pea_affs_bwd."""
    for (B, Cc, D, H, W, K, seed) in ((1, 16, 6, 30, 32, 12, 21), (2, 16, 5, 14, 18, 3, 22)):
        emb, ema, tgt, wgt = _inputs_3d(O, B, Cc, D, H, W, K, seed)
        emb, ema = np.abs(emb) + 0.05, np.abs(ema) + 0.05
        fo, fp = (O.embedding_loss_norm5, P.embedding_loss_norm5) if K == 12 else (O.embedding_loss_norm1, P.embedding_loss_norm1)
        xo, xp = ((O.ema_embedding_loss_norm5, P.ema_embedding_loss_norm5) if K == 12
                  else (O.ema_embedding_loss_norm1, P.ema_embedding_loss_norm1))
        for name, crit_o, crit_p in _criteria(P):
            eo = torch.tensor(emb, requires_grad=True)
            lo, ao = fo(eo, torch.tensor(tgt), torch.tensor(wgt), crit_o, affs0_weight=1.5)
            lo.backward()
            e = dev(emb, True)
            l, a = fp(e, dev(tgt), dev(wgt), crit_p, affs0_weight=1.5)
            l.backward()
            assert not a.requires_grad
            close_s(l, lo, what=f"K={K} {name} loss")
            close_t(a, ao, what=f"K={K} {name} affs")
            close_t(e.grad, eo.grad, what=f"K={K} {name} grad")
            eo, uo = torch.tensor(emb, requires_grad=True), torch.tensor(ema, requires_grad=True)
            lo, ao = xo(eo, uo, torch.tensor(tgt), torch.tensor(wgt), crit_o, affs0_weight=1.5)
            lo.backward()
            e, u = dev(emb, True), dev(ema, True)
            l, a = xp(e, u, dev(tgt), dev(wgt), crit_p, affs0_weight=1.5)
            l.backward()
            close_s(l, lo, what=f"K={K} {name} ema loss")
            close_t(e.grad, eo.grad, what=f"K={K} {name} ema grad")
            close_t(u.grad, uo.grad, what=f"K={K} {name} ema grad wrt partner")


@pytest.mark.parametrize("B,C,D,H,W,seed", [(2, 16, 7, 31, 45, 11), (1, 16, 18, 40, 36, 12), (1, 8, 5, 28, 28, 13)])
def test_oracle_norm5(P, O, B, C, D, H, W, seed):
    emb, ema, tgt, wgt = _inputs_3d(O, B, C, D, H, W, 12, seed)
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao = O.embedding_loss_norm5(eo, torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(), affs0_weight=1.5)
    lo.backward()
    e = dev(emb, True)
    l, a = P.embedding_loss_norm5(e, dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=1.5)
    l.backward()
    close_s(l, lo, what="loss")
    close_t(a, ao, what="affs")
    close_t(e.grad, eo.grad, what="grad")
    close_t(P.inf_embedding_loss_norm5(dev(emb)), ao, what="inf5 == loss affs")
    eo, uo = torch.tensor(emb, requires_grad=True), torch.tensor(ema, requires_grad=True)
    lo, ao = O.ema_embedding_loss_norm5(eo, uo, torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(), affs0_weight=1.5)
    lo.backward()
    e, u = dev(emb, True), dev(ema, True)
    l, a = P.ema_embedding_loss_norm5(e, u, dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=1.5)
    l.backward()
    close_s(l, lo, what="ema loss")
    close_t(a, ao, what="ema affs")
    close_t(e.grad, eo.grad, what="ema grad")
    close_t(u.grad, uo.grad, what="ema grad partner")


@pytest.mark.parametrize("B,C,D,H,W,shift,seed", [(2, 16, 18, 10, 10, 1, 21), (2, 16, 18, 20, 20, 1, 22),
                                                   (1, 16, 6, 17, 23, 2, 23)])
def test_oracle_norm1_deep_supervision_scales(P, O, B, C, D, H, W, shift, seed):
    """The /16 and /8 deep-supervision volumes of scripts_ac3ac4/main.py:227-230 (18 x 10 x 10, 18 x 20 x 20)."""
    emb, ema, tgt, wgt = _inputs_3d(O, B, C, D, H, W, 3, seed, replicate=True)
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao = O.embedding_loss_norm1(eo, torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(), affs0_weight=2, shift=shift)
    lo.backward()
    e = dev(emb, True)
    l, a = P.embedding_loss_norm1(e, dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=2, shift=shift)
    l.backward()
    close_s(l, lo, what="loss")
    close_t(a, ao, what="affs")
    close_t(e.grad, eo.grad, what="grad")
    close_t(P.inf_embedding_loss_norm1(dev(emb), shift=shift), ao, what="inf1")


def test_oracle_gt_random_shapes(P, O):
    rng = np.random.RandomState(5)
    for trial in range(6):
        H, W = int(rng.randint(1, 70)), int(rng.randint(1, 90))
        lab = O.voronoi_labels((H, W), 6, 100 + trial)
        offsets = O.multi_offset([1, 3, 9, 27], neighbor=8) + [[int(rng.randint(-80, 80)), int(rng.randint(-80, 80))]]
        for ig in (False, True):
            a, m = P.gen_affs_ours(lab, offsets, ignore=ig, padding=True)
            ao, mo = O.gen_affs_ours(lab, offsets, ignore=ig, padding=True)
            same(a, ao)
            same(m, mo)
        for k in range(0, len(offsets), 5):
            same(P.weight_binary_ratio(ao[k]), O.weight_binary_ratio(ao[k]))
    for trial in range(3):
        D, H, W = int(rng.randint(5, 12)), int(rng.randint(28, 50)), int(rng.randint(28, 50))
        lab = O.voronoi_labels((D, H, W), 15, 200 + trial, z_scale=3.0, dtype=np.uint16)
        from pea_b200.data_affinity import NHOOD_NORM5
        same(P.seg_to_aff(lab, NHOOD_NORM5, pad=''), O.seg_to_aff_norm5(lab))
        same(P.seg_to_aff(lab), O.seg_to_aff(lab))
    # all-background and single-label volumes
    z = np.zeros((4, 6, 7), np.uint8)
    same(P.seg_to_aff(z), O.seg_to_aff(z))
    same(P.weight_binary_ratio(O.seg_to_aff(z)[0]), O.weight_binary_ratio(O.seg_to_aff(z)[0]))
    one = np.full((4, 6, 7), 3, np.uint8)
    same(P.seg_to_aff(one), O.seg_to_aff(one))
    a, m = P.gen_affs_ours(one[0], [[-1, 0], [0, 2]], padding=True)
    ao, mo = O.gen_affs_ours(one[0], [[-1, 0], [0, 2]], padding=True)
    same(a, ao)
    same(m, mo)


# =========================================================================== autograd plumbing
def test_backward_scaling_retain_graph_and_no_grad(P, O):
    offsets = O.multi_offset([1, 3], neighbor=4)
    emb, ema, tgt, wgt, msk = _inputs_2d(O, 1, 16, 24, 28, offsets, 31)
    e = dev(emb, True)
    l, _, _ = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    l.backward(retain_graph=True)
    g1 = e.grad.clone()
    e.grad = None
    (2.5 * l).backward()           # second backward through the same node: recomputed, then scaled
    close_t(e.grad, 2.5 * g1, what="retain_graph + scaled upstream")
    with torch.no_grad():
        l2, a2, _ = P.embedding_loss(dev(emb), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    assert not l2.requires_grad and not a2.requires_grad
    close_s(l2, l, rtol=1e-7)
    # determinism: the loss reduction has a fixed order
    l3, _, _ = P.embedding_loss(dev(emb), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    assert float(l3) == float(l2)


def test_3d_kernels_reproducible_graph_capturable_and_stream_safe(P, O):
    """The Tensor-Memory kernels: (a) bit-identical results run to run (fixed-order loss reduction, no atomics on
    floats), (b) capturable in a CUDA graph, (c) two streams may run them at the same time (each CTA allocates the
    SM's Tensor Memory; shared memory keeps two of them off one SM)."""
    emb, ema, tgt, wgt = _inputs_3d(O, 2, 16, 9, 52, 68, 12, 51)
    T, Wt = dev(tgt), dev(wgt)

    def run(e_np, partner=None):
        e = dev(e_np, True)
        if partner is None:
            l, a = P.embedding_loss_norm5(e, T, Wt, P.WeightedMSE())
        else:
            l, a = P.ema_embedding_loss_norm5(e, dev(partner), T, Wt, P.WeightedMSE())
        l.backward()
        return l.detach().clone(), a.clone(), e.grad.clone()

    first, second = run(emb), run(emb)
    for x, y in zip(first, second):
        assert torch.equal(x, y), "self loss: run-to-run bits differ"
    fx, sx = run(emb, ema), run(emb, ema)
    for x, y in zip(fx, sx):
        assert torch.equal(x, y), "two-view loss: run-to-run bits differ"
    # (b) graph capture and replay
    e = dev(emb, True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            l, _ = P.embedding_loss_norm5(e, T, Wt, P.WeightedMSE())
            l.backward()
            e.grad = None
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        lg, ag = P.embedding_loss_norm5(e, T, Wt, P.WeightedMSE())
        lg.backward()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(lg.detach(), first[0]) and torch.equal(ag, first[1]) and torch.equal(e.grad, first[2])
    # (c) the self kernel and the two-view kernel on two streams at once
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        r1 = run(emb)
    with torch.cuda.stream(s2):
        r2 = run(emb, ema)
    torch.cuda.synchronize()
    for x, y in zip(r1, first):
        assert torch.equal(x, y)
    for x, y in zip(r2, fx):
        assert torch.equal(x, y)


def test_half_precision_embedding_inputs(P, O):
    """bf16 / fp16 embeddings (autocast backbones): the path runs in fp32 on the upcast values -- loss and affinities
    equal the fp32 call on those values exactly; the gradient returns in the input dtype, so it matches the fp32
    gradient to the rounding of that dtype (bf16: 2^-8, fp16: 2^-11 relative; tolerance 2x that of the largest entry)."""
    emb, ema, tgt, wgt = _inputs_3d(O, 1, 16, 7, 40, 48, 12, 31)
    for dt, tol in ((torch.bfloat16, 2.0 ** -7), (torch.float16, 2.0 ** -10)):
        eh = torch.tensor(emb).cuda().to(dt).requires_grad_(True)
        e32 = eh.detach().float().requires_grad_(True)
        lh, ah = P.embedding_loss_norm5(eh, dev(tgt), dev(wgt), P.WeightedMSE())
        l32, a32 = P.embedding_loss_norm5(e32, dev(tgt), dev(wgt), P.WeightedMSE())
        lh.backward()
        l32.backward()
        assert float(lh) == float(l32) and torch.equal(ah, a32)
        assert eh.grad.dtype == dt
        close_t(eh.grad.float().cpu(), e32.grad.cpu(), rtol=tol, what=f"{dt} gradient")
    offsets = O.multi_offset([1, 3, 5], neighbor=4)
    e2, _, t2, w2, m2 = _inputs_2d(O, 1, 16, 40, 64, offsets, 32)
    eh = torch.tensor(e2).cuda().bfloat16()
    lh, ah, _ = P.embedding_loss(eh, dev(t2), dev(w2), dev(m2), P.WeightedMSE(), offsets)
    l32, a32, _ = P.embedding_loss(eh.float(), dev(t2), dev(w2), dev(m2), P.WeightedMSE(), offsets)
    assert float(lh) == float(l32) and torch.equal(ah, a32)


def test_non_contiguous_and_wrong_inputs(P, O):
    offsets = [[-1, 0], [0, -1]]
    emb, ema, tgt, wgt, msk = _inputs_2d(O, 2, 16, 12, 14, offsets, 41)
    e_nc = dev(np.ascontiguousarray(emb.transpose(0, 1, 3, 2))).transpose(2, 3)   # non-contiguous view
    assert not e_nc.is_contiguous()
    l1, a1, _ = P.embedding_loss(e_nc, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    l2, a2, _ = P.embedding_loss(dev(emb), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    assert float(l1) == float(l2) and torch.equal(a1, a2)
    with pytest.raises(ValueError):
        P.embedding_loss(dev(emb), dev(tgt[:, :1]), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    with pytest.raises(TypeError):
        P.embedding_loss(dev(emb).double(), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
    with pytest.raises(ValueError):
        P.embedding_loss(dev(emb[:, :5]), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)  # C = 5


def test_cuda_graph_capture(P, O):
    """No host synchronisation inside forward or backward: the whole step captures into a CUDA graph."""
    offsets = O.multi_offset([1, 3], neighbor=4)
    emb, ema, tgt, wgt, msk = _inputs_2d(O, 1, 16, 32, 32, offsets, 51)
    e = dev(emb, True)
    T, Wm, M = dev(tgt), dev(wgt), dev(msk)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            l, _, _ = P.embedding_loss(e, T, Wm, M, P.WeightedMSE(), offsets)
            l.backward()
    torch.cuda.current_stream().wait_stream(s)
    ref_grad = e.grad.clone() / 2
    e.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        l, _, _ = P.embedding_loss(e, T, Wm, M, P.WeightedMSE(), offsets)
        l.backward()
    e.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    close_t(e.grad, ref_grad, what="graph replay grad")


# =========================================================================== tiled kernels vs table-driven kernel vs oracle
@pytest.mark.parametrize("B,D,H,W,seed", [
    (1, 5, 28, 32, 61),      # minimal extents for the 27-shifts, D = 5 exercises ring fill/drain
    (2, 7, 37, 44, 62),      # partial tiles in y (37 = 4*8+5) and x (44 = 32+12)
    (1, 18, 64, 96, 63),     # AC3/AC4 depth
    (1, 6, 30, 28, 64),      # W < one tile
])
def test_tiled3d_matches_generic_and_oracle(P, O, B, D, H, W, seed):
    from pea_b200 import _ops as ops
    emb, ema, tgt, wgt = _inputs_3d(O, B, 16, D, H, W, 12, seed)
    out = {}
    for name, impl in (("generic", ops.IMPL_GENERIC), ("tiled", ops.IMPL_TILED)):
        ops.set_default_impl(impl)
        try:
            e = dev(emb, True)
            l, a = P.embedding_loss_norm5(e, dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=1.5)
            l.backward()
            with torch.no_grad():   # forward-only evaluation (validation loop): no gradient buffers
                l_ng, a_ng = P.embedding_loss_norm5(dev(emb), dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=1.5)
            # two-view variant, second view detached (the shipped training loop): gradient w.r.t. the first view
            e2 = dev(emb, True)
            lx, ax = P.ema_embedding_loss_norm5(e2, dev(ema), dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=1.5)
            lx.backward()
            out[name] = (float(l.detach()), a.cpu(), e.grad.cpu(), float(l_ng), a_ng.cpu())
            out[name + "_x"] = (float(lx.detach()), ax.cpu(), e2.grad.cpu())
        finally:
            ops.set_default_impl(ops.IMPL_AUTO)
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao = O.embedding_loss_norm5(eo, torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(), affs0_weight=1.5)
    lo.backward()
    ex = torch.tensor(emb, requires_grad=True)
    lxo, axo = O.ema_embedding_loss_norm5(ex, torch.tensor(ema), torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(),
                                          affs0_weight=1.5)
    lxo.backward()
    for name in ("generic", "tiled"):
        lx, ax, gx = out[name + "_x"]
        close_s(lx, lxo, what=f"{name} cross loss")
        close_t(ax, axo, what=f"{name} cross affs")
        close_t(gx, ex.grad, what=f"{name} cross grad")
    for name in ("generic", "tiled"):
        l, a, g, l_ng, a_ng = out[name]
        close_s(l, lo, what=f"{name} loss")
        close_t(a, ao, what=f"{name} affs")
        close_t(g, eo.grad, what=f"{name} grad")
        close_s(l_ng, lo, what=f"{name} loss (no grad)")
        close_t(a_ng, ao, what=f"{name} affs (no grad)")
    # border semantics are exact zeros in both
    assert float((out["tiled"][1][:, 0, 0]).abs().max()) == 0.0      # z-channel 0 is zero on the first slice
    assert float((out["tiled"][1][:, 10, :, :27]).abs().max()) == 0.0  # y-27 channel zero on the first 27 rows


def test_tiled3d_forced_on_ineligible_input_is_refused(P, O):
    from pea_b200 import _ops as ops
    emb, ema, tgt, wgt = _inputs_3d(O, 1, 16, 5, 28, 30, 12, 71)   # W = 30 is not a multiple of 4: TMA row pitch
    ops.set_default_impl(ops.IMPL_TILED)
    try:
        with pytest.raises(Exception, match="code -8"):
            P.embedding_loss_norm5(dev(emb), dev(tgt), dev(wgt), P.WeightedMSE())
    finally:
        ops.set_default_impl(ops.IMPL_AUTO)
    l, a = P.embedding_loss_norm5(dev(emb), dev(tgt), dev(wgt), P.WeightedMSE())   # AUTO falls back to the table kernel
    lo, ao = O.embedding_loss_norm5(torch.tensor(emb), torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE())
    close_s(l, lo)
    close_t(a, ao)


@pytest.mark.parametrize("B,H,W,shifts,nb,seed", [
    (2, 37, 53, [1, 3, 5, 9, 27], 4, 81),     # plus-shaped halo 27, partial tiles
    (1, 40, 48, [1, 3, 5, 9, 11], 8, 82),     # diagonals -> rectangular halo
    (3, 24, 20, [1, 30], 4, 83),              # shift larger than the image: halo wraps more than once
    (1, 130, 70, [1, 3, 5, 9, 11], 4, 84),    # several tiles in y and x
])
def test_tiled2d_matches_generic_and_oracle(P, O, B, H, W, shifts, nb, seed):
    from pea_b200 import _ops as ops
    offsets = O.multi_offset(shifts, neighbor=nb)
    emb, ema, tgt, wgt, msk = _inputs_2d(O, B, 16, H, W, offsets, seed, float_mask=(seed % 2 == 0))
    res = {}
    for name, impl in (("generic", ops.IMPL_GENERIC), ("tiled", ops.IMPL_TILED)):
        ops.set_default_impl(impl)
        try:
            e = dev(emb, True)
            l, a, al = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
            l.backward()
            e2 = dev(emb, True)
            l2, a2 = P.ema_embedding_loss(e2, dev(ema), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets,
                                          affs0_weight=0.25)
            l2.backward()
            fwd = ops.affs_fwd(dev(emb), None, ops.offsets3(offsets, 2), ops.BORDER_WRAP, impl)
            fwd0 = ops.affs_fwd(dev(emb), None, ops.offsets3(offsets, 2), ops.BORDER_ZERO, impl)
            res[name] = (float(l.detach()), a.detach().cpu(), e.grad.cpu(), np.asarray(list(al)), float(l2.detach()),
                         a2.detach().cpu(), e2.grad.cpu(), fwd.cpu(), fwd0.cpu())
        finally:
            ops.set_default_impl(ops.IMPL_AUTO)
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao, allo = O.embedding_loss(eo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), O.WeightedMSE(), offsets)
    lo.backward()
    eo2 = torch.tensor(emb, requires_grad=True)
    lo2, ao2 = O.ema_embedding_loss(eo2, torch.tensor(ema), torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk),
                                    O.WeightedMSE(), offsets, affs0_weight=0.25)
    lo2.backward()
    for name in ("generic", "tiled"):
        l, a, g, al, l2, a2, g2, fwd, fwd0 = res[name]
        close_s(l, lo, what=f"{name} loss")
        close_t(a, ao, what=f"{name} affs")
        close_t(g, eo.grad, what=f"{name} grad")
        close_t(al, np.asarray(allo), what=f"{name} all_loss")
        close_s(l2, lo2, what=f"{name} ema loss")
        close_t(a2, ao2, what=f"{name} ema affs")
        close_t(g2, eo2.grad, what=f"{name} ema grad")
        close_t(fwd, ao, what=f"{name} embedding2affs")
    # zero-border forward (D = 1 volumes through the C ABI): tiled and table-driven kernels agree
    close_t(res["tiled"][8], res["generic"][8], what="zero-border forward")


@pytest.mark.parametrize("B,H,W,shifts,seed", [
    (2, 40, 64, [1, 3, 5, 9, 11], 85),      # every CTA on the image edge (hand-filled boxes under the circular border)
    (1, 130, 128, [1, 3, 5, 9, 11], 86),    # interior CTAs (TMA boxes) and edge CTAs in one launch
    (1, 67, 96, [1, 3, 5, 9, 27], 87),      # CVPPP shifts: the 28-pixel halo variant, partial tiles in y
    (1, 150, 160, [1, 3, 5, 9, 27], 88),
])
def test_axis2d_kernel_matches_generic_and_oracle(P, O, B, H, W, shifts, seed):
    """multi_offset(shifts, neighbor=4) tables at W % 4 == 0 take the TMA-staged axis-aligned kernel (AUTO); it must
    agree with the table-driven kernel and with the oracle, for both masks, both views and with / without gradient."""
    from pea_b200 import _ops as ops
    offsets = O.multi_offset(shifts, neighbor=4)
    for float_mask in (False, True):
        emb, ema, tgt, wgt, msk = _inputs_2d(O, B, 16, H, W, offsets, seed, float_mask=float_mask)
        res = {}
        for name, impl in (("generic", ops.IMPL_GENERIC), ("auto", ops.IMPL_AUTO)):
            ops.set_default_impl(impl)
            try:
                e = dev(emb, True)
                l, a, al = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
                l.backward()
                e2 = dev(emb, True)
                l2, a2 = P.ema_embedding_loss(e2, dev(ema), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets,
                                              affs0_weight=0.25)
                l2.backward()
                with torch.no_grad():
                    l3, a3, _ = P.embedding_loss(dev(emb), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
                res[name] = (float(l.detach()), a.detach().cpu(), e.grad.cpu(), np.asarray(list(al)), float(l2.detach()),
                             a2.detach().cpu(), e2.grad.cpu(), float(l3), a3.cpu())
            finally:
                ops.set_default_impl(ops.IMPL_AUTO)
        eo = torch.tensor(emb, requires_grad=True)
        lo, ao, allo = O.embedding_loss(eo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), O.WeightedMSE(), offsets)
        lo.backward()
        eo2 = torch.tensor(emb, requires_grad=True)
        lo2, ao2 = O.ema_embedding_loss(eo2, torch.tensor(ema), torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk),
                                        O.WeightedMSE(), offsets, affs0_weight=0.25)
        lo2.backward()
        for name in ("generic", "auto"):
            l, a, g, al, l2, a2, g2, l3, a3 = res[name]
            close_s(l, lo, what=f"{name} loss")
            close_t(a, ao, what=f"{name} affs")
            close_t(g, eo.grad, what=f"{name} grad")
            close_t(al, np.asarray(allo), what=f"{name} all_loss")
            close_s(l2, lo2, what=f"{name} ema loss")
            close_t(a2, ao2, what=f"{name} ema affs")
            close_t(g2, eo2.grad, what=f"{name} ema grad")
            close_s(l3, lo, what=f"{name} loss (no grad)")
            close_t(a3, ao, what=f"{name} affs (no grad)")
    # zero border (D = 1 volumes straight through the C ABI), mixed-sign offsets, no mask: kernels agree
    offs3 = ops.offsets3([[3, 0], [-12, 0], [0, 7], [0, -1]], 2)
    g = torch.Generator().manual_seed(seed)
    e5 = torch.randn((B, 16, 1, H, W), generator=g).cuda()
    t5 = (torch.rand((B, 4, 1, H, W), generator=g) > 0.5).float().cuda()
    w5 = (torch.rand((B, 4, 1, H, W), generator=g) + 0.1).cuda()
    norm, coef = np.full(4, float(B * W), np.float32), np.asarray([1.0, 0.5, 2.0, 1.0], np.float32)
    outs = [ops.affs_loss_fwd_bwd(e5, None, offs3, ops.BORDER_ZERO, t5, w5, None, norm, coef, impl=impl)
            for impl in (ops.IMPL_GENERIC, ops.IMPL_AUTO)]
    for i, what in ((0, "loss"), (1, "affs"), (2, "grad")):
        close_t(outs[1][i].cpu(), outs[0][i].cpu(), what=f"zero-border axis kernel {what}")


KNOB_NO_CANON, KNOB_MAX_CTAS = 0, 1   # include/pea_b200.h, pea_debug_configure (test-only dispatch knobs)


def _knob(which, value):
    from pea_b200._lib import lib, check
    check("pea_debug_configure", lib.pea_debug_configure(which, int(value)))


@pytest.mark.parametrize("shifts", [[1, 3, 5, 9, 11], [1, 3, 5, 9, 27]])
@pytest.mark.parametrize("cap", [1, 3, 7])
def test_axis2d_persistent_ctas_walk_many_tiles(P, O, cap, shifts):
    """The 2-D kernels are persistent: a CTA walks tiles t, t + grid, ... and requests the TMA boxes of its next tile
    while it finishes the current one; tiles on the image edge (circular border) are filled / patched by hand in between.
    Shifts <= 12 take the fully unrolled canonical kernel (pea_affs2d_canon.cu), the CVPPP table with shift 27 the
    table-driven one (pea_affs2d_axis.cu); the two-view loss always takes the latter.  With the grid capped to a few
    CTAs (pea_debug_configure(PEA_KNOB_MAX_CTAS)) every CTA alternates between both kinds of tile many times: affinities and
    gradient must be bit-identical to the uncapped launch, the loss equal up to the summation order of the per-CTA
    partials, and everything must match the oracle."""
    offsets = O.multi_offset(shifts, neighbor=4)
    B, H, W = 2, 130, 160
    emb, ema, tgt, wgt, msk = _inputs_2d(O, B, 16, H, W, offsets, 95)
    outs = {}
    for name, val in (("full", None), ("capped", str(cap))):
        _knob(KNOB_MAX_CTAS, 0 if val is None else val)
        e = dev(emb, True)
        l, a, _ = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
        l.backward()
        e2 = dev(emb, True)
        l2, a2 = P.ema_embedding_loss(e2, dev(ema), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets)
        l2.backward()
        outs[name] = (float(l.detach()), a.detach().cpu(), e.grad.cpu(), float(l2.detach()), a2.detach().cpu(), e2.grad.cpu())
    _knob(KNOB_MAX_CTAS, 0)
    for i in (1, 2, 4, 5):
        assert torch.equal(outs["full"][i], outs["capped"][i])
    close_s(outs["capped"][0], outs["full"][0], what="loss, capped grid")
    close_s(outs["capped"][3], outs["full"][3], what="ema loss, capped grid")
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao, _ = O.embedding_loss(eo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), O.WeightedMSE(), offsets)
    lo.backward()
    close_s(outs["capped"][0], lo, what="loss vs oracle")
    close_t(outs["capped"][1], ao, what="affs vs oracle")
    close_t(outs["capped"][2], eo.grad, what="grad vs oracle")


@pytest.mark.parametrize("B,H,W,border", [(2, 130, 160, "wrap"), (1, 75, 100, "wrap"), (2, 64, 72, "zero"), (1, 131, 200, "zero")])
@pytest.mark.parametrize("mask_kind", ["u8", "f32", "none"])
def test_canonical2d_kernel_matches_table_driven_kernels(P, O, B, H, W, border, mask_kind):
    """multi_offset([1, 3, 5, 9, 11], 4) takes the fully unrolled canonical kernel (one view, and two views with the
    second one constant); with pea_debug_configure(PEA_KNOB_NO_CANON, 1) the same call takes the table-driven axis-aligned kernel, with
    IMPL_GENERIC the generic one.  All three must agree (1e-5 of the largest magnitude) on loss, affinities and
    gradient: both borders, partial tiles in x and y, every mask type, both view counts, with and without gradient."""
    from pea_b200 import _ops as ops
    offsets = O.multi_offset([1, 3, 5, 9, 11], neighbor=4)
    offs3 = ops.offsets3(offsets, 2)
    K = len(offsets)
    g = torch.Generator().manual_seed(B * 1000 + H)
    emb = torch.randn((B, 16, 1, H, W), generator=g).cuda()
    emb[0, :, 0, 3, 5] = 0.0
    par = torch.randn((B, 16, 1, H, W), generator=g).cuda()
    par[0, :, 0, 7, 9] = 0.0
    tgt = (torch.rand((B, K, 1, H, W), generator=g) > 0.5).float().cuda()
    wgt = (torch.rand((B, K, 1, H, W), generator=g) * 2 + 0.1).cuda()
    m8 = (torch.rand((B, K, 1, H, W), generator=g) > 0.2).to(torch.uint8).cuda()
    msk = {"u8": m8, "f32": m8.float() * 0.5 + 0.5, "none": None}[mask_kind]
    bd = ops.BORDER_WRAP if border == "wrap" else ops.BORDER_ZERO
    norm = np.full(K, float(B * W), np.float32)
    coef = np.linspace(0.5, 1.5, K).astype(np.float32)
    for grad, partner in ((True, None), (False, None), (True, par), (False, par)):
        outs = {}
        for name in ("canon", "axis", "generic"):
            _knob(KNOB_NO_CANON, 1 if name == "axis" else 0)
            impl = ops.IMPL_GENERIC if name == "generic" else ops.IMPL_AUTO
            lo, a, gr, _ = ops.affs_loss_fwd_bwd(emb, partner, offs3, bd, tgt, wgt, msk, norm, coef, want_affs=True,
                                                 want_grad=grad, impl=impl)
            outs[name] = (lo.cpu(), a.cpu(), gr.cpu() if grad else None)
        _knob(KNOB_NO_CANON, 0)
        for other in ("axis", "generic"):
            close_t(outs["canon"][0], outs[other][0], what=f"loss vs {other}")
            close_t(outs["canon"][1], outs[other][1], what=f"affs vs {other}")
            if grad:
                close_t(outs["canon"][2], outs[other][2], what=f"grad vs {other}")


@pytest.mark.parametrize("shifts", [[1, 3, 5, 9, 11], [1, 3, 5, 9, 27], [2], [1, 2, 4, 8, 12, 28], [3, 7, 20]])
@pytest.mark.parametrize("B,H,W", [(2, 130, 160), (1, 75, 100), (3, 33, 72)])
def test_forward_only_2d_kernel(P, O, shifts, B, H, W):
    """embedding2affs (scripts_cvppp/loss/loss_embedding_mse.py:58-66) at the axis-aligned neighbor-4 tables takes the
    unrolled forward kernel (pea_affs2d_fwd.cu: shifts <= 12 double-buffered, up to 28 single-buffered).  Against the
    oracle (circular border), the table-driven generic kernel and the older tiled kernel (PEA_KNOB_NO_CANON): both borders,
    relu folded into the store or not, 1 - 6 shifts, partial tiles in x and y, a zero-norm pixel; a grid capped to 2 CTAs
    (every CTA walks many tiles through both sets of boxes) must give bit-identical results."""
    from pea_b200 import _ops as ops
    offsets = O.multi_offset(shifts, neighbor=4)
    offs3 = ops.offsets3(offsets, 2)
    g = torch.Generator().manual_seed(B * 1000 + H)
    emb = torch.randn((B, 16, H, W), generator=g)
    emb[0, :, 3, 5] = 0.0
    emb_d = emb.cuda()
    try:
        for border in (ops.BORDER_WRAP, ops.BORDER_ZERO):
            for relu in (False, True):
                outs = {}
                for name in ("new", "new_cap2", "old", "generic"):
                    _knob(KNOB_NO_CANON, 1 if name == "old" else 0)
                    _knob(KNOB_MAX_CTAS, 2 if name == "new_cap2" else 0)
                    impl = ops.IMPL_GENERIC if name == "generic" else ops.IMPL_AUTO
                    outs[name] = ops.affs_fwd(emb_d, None, offs3, border, impl=impl, relu=relu).cpu()
                assert torch.equal(outs["new"], outs["new_cap2"]), "capped grid differs"
                close_t(outs["new"], outs["generic"], what=f"vs generic, border {border}, relu {relu}")
                close_t(outs["new"], outs["old"], what=f"vs tiled, border {border}, relu {relu}")
                if border == ops.BORDER_WRAP:
                    ao = O.embedding2affs(emb, offsets)
                    close_t(outs["new"], torch.relu(ao) if relu else ao, what=f"vs oracle, relu {relu}")
    finally:
        _knob(KNOB_NO_CANON, 0)
        _knob(KNOB_MAX_CTAS, 0)
    # the public call, with and without autograd (backward through pea_affs_bwd)
    e = emb_d.clone().requires_grad_(True)
    a = P.embedding2affs(e, offsets)
    a.square().sum().backward()
    eo = emb.clone().requires_grad_(True)
    ao = O.embedding2affs(eo, offsets)
    ao.square().sum().backward()
    close_t(a, ao, what="public call affs")
    close_t(e.grad, eo.grad, what="public call grad")


@pytest.mark.parametrize("case", ["2d_wrap", "2d_zero_two_views", "3d_norm5", "3d_diagonals_two_views"])
def test_k5_backward_routes_agree(P, O, case):
    """pea_affs_bwd (gradient through the affinities, for criteria the fused kernels do not implement): the work-list
    kernel in upstream mode against the table-driven kernel (PEA_KNOB_NO_CANON) and torch autograd through the oracle's
    affinities: 2-D and 3-D, both borders, one and two views (gradient for the anchor view), a zero-norm voxel."""
    from pea_b200 import _ops as ops
    g = torch.Generator().manual_seed(11)
    if case.startswith("2d"):
        shape, offs = (2, 16, 1, 50, 68), ops.offsets3(O.multi_offset([1, 3, 9, 27], neighbor=4), 2)
        border = ops.BORDER_WRAP if case == "2d_wrap" else ops.BORDER_ZERO
    else:
        from pea_b200.data_affinity import NHOOD_NORM5
        shape, border = (2, 16, 6, 30, 36), ops.BORDER_ZERO
        offs = np.asarray(NHOOD_NORM5, np.int32) if case == "3d_norm5" else np.asarray(
            [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, -1], [-1, 1, 1], [0, -9, -4], [0, 4, -9]], np.int32)
    two = case.endswith("two_views")
    emb = torch.randn(shape, generator=g)
    emb[0, :, 0, 3, 5] = 0.0
    par = torch.randn(shape, generator=g) if two else None
    K = offs.shape[0]
    up = torch.randn((shape[0], K) + shape[2:], generator=g)
    outs = {}
    try:
        for name in ("worklist", "table"):
            _knob(KNOB_NO_CANON, 1 if name == "table" else 0)
            ge, _ = ops.affs_bwd(emb.cuda(), par.cuda() if two else None, offs, border, up.cuda(), want_grad=True, want_grad_partner=False)
            outs[name] = ge.cpu()
    finally:
        _knob(KNOB_NO_CANON, 0)
    close_t(outs["worklist"], outs["table"], what="work-list vs table-driven kernel")
    # torch autograd through plain torch ops (normalise, shift with the border rule, dot)
    eo = emb.clone().requires_grad_(True)
    en = torch.nn.functional.normalize(eo, p=2, dim=1)
    un = torch.nn.functional.normalize(par, p=2, dim=1) if two else en
    loss = 0.0
    D, H, W = shape[2:]
    for k, (dz, dy, dx) in enumerate(offs.tolist()):
        if border == ops.BORDER_WRAP:
            sh = torch.roll(un, shifts=(-dz, -dy, -dx), dims=(2, 3, 4))
            a = (en * sh).sum(1)
        else:
            a = torch.zeros((shape[0], D, H, W))
            zs = slice(max(0, -dz), min(D, D - dz)); ys = slice(max(0, -dy), min(H, H - dy)); xs = slice(max(0, -dx), min(W, W - dx))
            zq = slice(max(0, dz), min(D, D + dz)); yq = slice(max(0, dy), min(H, H + dy)); xq = slice(max(0, dx), min(W, W + dx))
            a[:, zs, ys, xs] = (en[:, :, zs, ys, xs] * un[:, :, zq, yq, xq]).sum(1)
        loss = loss + (a * up[:, k]).sum()
    loss.backward()
    close_t(outs["worklist"], eo.grad, what="work-list kernel vs torch autograd")


@pytest.mark.parametrize("B,H,W,shifts,ldtype,ignore", [
    (2, 130, 160, [1, 3, 5, 9, 11], np.uint8, False),    # canonical table: GT derived inside the loss kernel
    (1, 75, 100, [1, 3, 5, 9, 11], np.int32, True),      # partial tiles in x and y, ignore=True
    (1, 64, 96, [1, 2, 4, 8, 12], np.uint8, False),
    (1, 67, 96, [1, 3, 5, 9, 27], np.uint8, False),      # CVPPP table: label-driven flavour of the table-driven axis kernel
    (2, 130, 160, [1, 3, 5, 9, 27], np.int16, True),     # ... interior + edge tiles, ignore=True
    (1, 64, 96, [2, 7], np.uint8, False),                # a table the canonical kernel does not know
    (1, 66, 94, [1, 3, 5, 9, 27], np.uint8, False),      # W % 4 != 0: two-step route behind the same call
])
def test_embedding_loss_labels_equals_gt_then_loss(P, O, B, H, W, shifts, ldtype, ignore):
    """embedding_loss_labels(embedding, labels, ...) == gen_affs_ours + weight_binary_ratio per (sample, channel) +
    embedding_loss: against this package's own two-step route and against the oracle's (the reference's data-provider
    recipe, scripts_cvppp/data/data_provider.py:204-225), loss / affinities / gradient within 1e-5."""
    offsets = O.multi_offset(shifts, neighbor=4)
    rng = np.random.RandomState(B * 100 + H)
    labs = np.stack([O.voronoi_labels((H, W), 9, 40 + b, background_frac=0.3, dtype=np.uint8) for b in range(B)]).astype(ldtype)
    emb = rng.randn(B, 16, H, W).astype(np.float32)
    emb[0, :, 3, 5] = 0.0
    lab_d = torch.tensor(labs).cuda()
    e = dev(emb, True)
    l, a, al = P.embedding_loss_labels(e, lab_d, P.WeightedMSE(), offsets, ignore=ignore)
    l.backward()
    # two-step route of this package
    tgt, msk, wgt = P.make_targets_2d(lab_d, offsets, ignore=ignore, padding=True)
    e2 = dev(emb, True)
    l2, a2, al2 = P.embedding_loss(e2, tgt, wgt, msk, P.WeightedMSE(), offsets)
    l2.backward()
    close_s(float(l.detach()), float(l2.detach()), what="loss vs two-step")
    close_t(a.detach().cpu(), a2.detach().cpu(), what="affs vs two-step")
    close_t(e.grad.cpu(), e2.grad.cpu(), what="grad vs two-step")
    close_t(np.asarray(list(al)), np.asarray(list(al2)), what="all_loss vs two-step")
    # oracle: the reference's host-side recipe
    to, mo, wo = [], [], []
    for b in range(B):
        t_, m_ = O.gen_affs_ours(labs[b], offsets, ignore=ignore, padding=True)
        to.append(t_)
        mo.append(m_)
        wo.append(np.stack([O.weight_binary_ratio(c) for c in t_]))
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao, _ = O.embedding_loss(eo, torch.tensor(np.stack(to)), torch.tensor(np.stack(wo)), torch.tensor(np.stack(mo)),
                                 O.WeightedMSE(), offsets)
    lo.backward()
    close_s(float(l.detach()), lo, what="loss vs oracle")
    close_t(a.detach().cpu(), ao, what="affs vs oracle")
    close_t(e.grad.cpu(), eo.grad, what="grad vs oracle")


@pytest.mark.parametrize("shifts,two_view,labels", [([1, 3, 5, 9, 11], False, False), ([1, 3, 5, 9, 11], True, False),
                                                     ([1, 3, 5, 9, 11], False, True), ([1, 3, 5, 9, 27], False, False),
                                                     ([1, 3, 5, 9, 27], True, False), ([1, 3, 5, 9, 27], False, True)])
def test_2d_kernels_write_only_their_outputs(P, O, shifts, two_view, labels):
    """Guard bands (compute-sanitizer is not available on the GPU pool): affinities, gradient, loss vector and workspace
    are carved out of larger poisoned buffers and handed to the C ABI directly; at a shape with partial tiles in x and
    y (75 x 100) the persistent 2-D kernels must leave every byte outside their outputs untouched, and write every
    element inside them."""
    import ctypes
    from pea_b200 import _ops as ops
    from pea_b200._lib import lib, check
    B, C, H, W = 2, 16, 75, 100
    offsets = O.multi_offset(shifts, neighbor=4)
    offs3 = ops.offsets3(offsets, 2)
    K = offs3.shape[0]
    g = torch.Generator().manual_seed(7)
    emb = torch.randn((B, C, H, W), generator=g).cuda()
    par = torch.randn((B, C, H, W), generator=g).cuda() if two_view else None
    tgt = (torch.rand((B, K, H, W), generator=g) > 0.5).float().cuda()
    wgt = (torch.rand((B, K, H, W), generator=g) + 0.1).cuda()
    msk = (torch.rand((B, K, H, W), generator=g) > 0.2).to(torch.uint8).cuda()
    lab = torch.tensor(np.stack([O.voronoi_labels((H, W), 7, 60 + b, dtype=np.uint8) for b in range(B)]).astype(np.int32)).cuda()
    POISON, PAD = -12345.0, 4096

    def guarded(n):
        buf = torch.full((n + 2 * PAD,), POISON, dtype=torch.float32, device="cuda")
        return buf, buf[PAD:PAD + n]

    affs_buf, affs = guarded(B * K * H * W)
    grad_buf, grad = guarded(B * C * H * W)
    loss_buf, loss = guarded(1 + K)
    nws = lib.pea_workspace_bytes(B, C, 1, H, W, K, int(two_view))
    ws_buf, ws = guarded((nws + 3) // 4)
    norm = (ctypes.c_float * K)(*([float(B * W)] * K))
    coef = (ctypes.c_float * K)(*([1.0] * K))
    offs = (ctypes.c_int32 * (3 * K))(*[int(v) for v in offs3.reshape(-1)])
    st = torch.cuda.current_stream().cuda_stream
    if labels:
        wtab = ops.gt_pair_weights(lab[:, None], offs3, 0)
        rc = lib.pea_affs_loss_labels_fwd_bwd(emb.data_ptr(), lab.data_ptr(), B, C, 1, H, W, offs, K, ops.BORDER_WRAP, 0, 0,
                                              wtab.data_ptr(), norm, coef, affs.data_ptr(), loss.data_ptr(), grad.data_ptr(),
                                              ws.data_ptr(), nws, st)
        check("pea_affs_loss_labels_fwd_bwd", rc)
    else:
        rc = lib.pea_affs_loss_fwd_bwd(emb.data_ptr(), par.data_ptr() if two_view else None, B, C, 1, H, W, offs, K,
                                       ops.BORDER_WRAP, ops.IMPL_AUTO, tgt.data_ptr(), wgt.data_ptr(), msk.data_ptr(), ops.MASK_U8,
                                       norm, coef, affs.data_ptr(), loss.data_ptr(), grad.data_ptr(), None, ws.data_ptr(), nws, st)
        check("pea_affs_loss_fwd_bwd", rc)
    torch.cuda.synchronize()
    for name, buf, n in (("affs", affs_buf, B * K * H * W), ("grad", grad_buf, B * C * H * W), ("loss", loss_buf, 1 + K),
                         ("workspace", ws_buf, (nws + 3) // 4)):
        assert bool((buf[:PAD] == POISON).all()) and bool((buf[PAD + n:] == POISON).all()), f"{name}: guard band overwritten"
    for name, out in (("affs", affs), ("grad", grad), ("loss", loss)):
        assert not bool((out == POISON).any()), f"{name}: elements left unwritten"
        assert bool(torch.isfinite(out).all()), f"{name}: non-finite values"


@pytest.mark.parametrize("B,D,H,W,seed", [(2, 7, 37, 44, 71), (1, 18, 64, 96, 72), (1, 5, 30, 28, 73), (1, 6, 30, 30, 74)])
def test_embedding_loss_norm5_labels_equals_gt_then_loss(P, O, B, D, H, W, seed):
    """embedding_loss_norm5_labels(embedding, seg, ...) == seg_to_aff(12 offsets, pad='') + weight_binary_ratio per
    (sample, channel) + embedding_loss_norm5: against this package's two-step route and the oracle's (the reference's
    data-provider recipe, data_provider_labeled_deep.py:246-285); partial tiles, D = 5 ring fill / drain, and a width
    (30) that takes the two-step route behind the same call."""
    rng = np.random.RandomState(seed)
    segs = np.stack([O.voronoi_labels((D, H, W), 9, seed + b, background_frac=0.15, z_scale=4.0, dtype=np.uint16)
                     for b in range(B)]).astype(np.int32)
    emb = rng.randn(B, 16, D, H, W).astype(np.float32)
    emb[0, :, 1, 3, 5] = 0.0
    seg_d = torch.tensor(segs).cuda()
    e = dev(emb, True)
    l, a = P.embedding_loss_norm5_labels(e, seg_d, P.WeightedMSE(), affs0_weight=2)
    l.backward()
    tgt, wgt = P.make_targets_3d(seg_d)
    e2 = dev(emb, True)
    l2, a2 = P.embedding_loss_norm5(e2, tgt, wgt, P.WeightedMSE(), affs0_weight=2)
    l2.backward()
    close_s(float(l.detach()), float(l2.detach()), what="loss vs two-step")
    close_t(a.cpu(), a2.cpu(), what="affs vs two-step")
    close_t(e.grad.cpu(), e2.grad.cpu(), what="grad vs two-step")
    to = np.stack([O.seg_to_aff_norm5(sg) for sg in segs])
    wo = np.stack([np.stack([O.weight_binary_ratio(c) for c in t_]) for t_ in to])
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao = O.embedding_loss_norm5(eo, torch.tensor(to), torch.tensor(wo), O.WeightedMSE(), affs0_weight=2)
    lo.backward()
    close_s(float(l.detach()), lo, what="loss vs oracle")
    close_t(a.cpu(), ao, what="affs vs oracle")
    close_t(e.grad.cpu(), eo.grad, what="grad vs oracle")


# =========================================================================== forward-only tiled kernel + z-shards
@pytest.mark.parametrize("B,D,H,W", [(1, 18, 64, 96), (2, 7, 37, 44), (1, 5, 30, 30)])
def test_inference_epilogue_fused(P, O, B, D, H, W):
    """relu / replicate_border of inf_embedding_loss_norm5 / norm1 == the reference callers' epilogue
    (scripts_ac3ac4/inference.py:160-165) applied with torch to the plain result, bit for bit (tiled kernel at
    W % 4 == 0, table-driven kernel otherwise)."""
    g = torch.Generator().manual_seed(D * 100 + H)
    emb = torch.randn((B, 16, D, H, W), generator=g).cuda()
    for fn in (P.inf_embedding_loss_norm5, P.inf_embedding_loss_norm1):
        plain = fn(emb)
        ref = plain.clone()
        ref[:, 1, :, :1, :] = ref[:, 1, :, 1:2, :]
        ref[:, 2, :, :, :1] = ref[:, 2, :, :, 1:2]
        ref[:, 0, :1, :, :] = ref[:, 0, 1:2, :, :]
        ref = torch.relu(ref)
        got = fn(emb, relu=True, replicate_border=True)
        assert torch.equal(got, ref)
        assert torch.equal(fn(emb, relu=True), torch.relu(plain))
        assert (plain < 0).any()  # the clamp is exercised


# (1, 5, 28, 672): 21 tile columns -- one full block of 16 and a narrower last block of the forward kernel's tile order;
# (2, 5, 28, 1056): 33 tile columns, two samples (the block walk restarts per sample)
@pytest.mark.parametrize("B,D,H,W,seed", [(1, 18, 64, 96, 91), (2, 7, 37, 44, 92), (1, 5, 28, 28, 93), (1, 5, 28, 672, 94),
                                          (2, 5, 28, 1056, 95)])
def test_tiled3d_forward_matches_generic_and_oracle(P, O, B, D, H, W, seed):
    from pea_b200 import _ops as ops
    rng = np.random.RandomState(seed)
    emb = rng.randn(B, 16, D, H, W).astype(np.float32)
    ref = O.inf_embedding_loss_norm5(torch.tensor(emb))
    outs = {}
    for name, impl in (("generic", ops.IMPL_GENERIC), ("tiled", ops.IMPL_TILED)):
        ops.set_default_impl(impl)
        try:
            outs[name] = P.inf_embedding_loss_norm5(dev(emb)).cpu()
        finally:
            ops.set_default_impl(ops.IMPL_AUTO)
        close_t(outs[name], ref, what=f"{name} inf5")
    assert float(outs["tiled"][:, 9, :4].abs().max()) == 0.0   # z-4 channel: zero border on the first four slices


@pytest.mark.parametrize("world,D", [(1, 9), (2, 11), (3, 20), (4, 18)])
def test_zshard_emulated_on_one_gpu_is_bit_identical(P, O, world, D):
    """Every rank's slab (with the halo copied in by hand instead of NCCL) through the z-range-view kernel path:
    the concatenation must equal the single-GPU whole-volume result bit for bit (same per-voxel arithmetic)."""
    from pea_b200.zshard import ZShard, sharded_affinities
    torch.manual_seed(world * 100 + D)
    vol = torch.randn(1, 16, D, 40, 64, device="cuda")
    whole = P.inf_embedding_loss_norm5(vol)
    close_t(whole, O.inf_embedding_loss_norm5(vol.cpu()), what="whole volume vs oracle")
    parts = []
    for rank in range(world):
        sh = ZShard(1, 16, D, 40, 64, rank, world, "cuda")
        sh.own.copy_(vol[:, :, sh.z0:sh.z1])
        if sh.halo:
            sh.buf[:, :, :sh.halo].copy_(vol[:, :, sh.z0 - sh.halo:sh.z0])
        sh.world = 1   # no process group in this test: the exchange was emulated by the copy above
        parts.append(sharded_affinities(sh))
    assert torch.equal(torch.cat(parts, dim=2), whole)


@pytest.mark.parametrize("world,D", [(2, 11), (3, 20), (4, 17)])
def test_zshard_fused_peer_halo_is_bit_identical(P, O, world, D):
    """The halo is not exchanged at all: each slab's kernel reads the previous slab's last four slices in place through
    the `halo` pointer of pea_affs_fwd (on a multi-GPU box that pointer is peer memory; here another local buffer)."""
    from pea_b200.zshard import shard_ranges, fused_halo_affinities, HALO
    torch.manual_seed(world * 1000 + D)
    vol = torch.randn(2, 16, D, 40, 64, device="cuda")
    whole = P.inf_embedding_loss_norm5(vol)
    ranges = shard_ranges(D, world)
    cap = max(b - a for a, b in ranges)
    slabs = []
    for z0, z1 in ranges:                      # same-size buffers like symmetric memory: slabs are z-range views
        buf = torch.full((2, 16, cap, 40, 64), float("nan"), device="cuda")
        buf[:, :, :z1 - z0].copy_(vol[:, :, z0:z1])
        slabs.append(buf[:, :, :z1 - z0])
    parts = []
    for r in range(world):
        tail = slabs[r - 1][:, :, -HALO:] if r else None
        parts.append(fused_halo_affinities(slabs[r], tail))
    assert torch.equal(torch.cat(parts, dim=2), whole)
    # a halo thinner than four slices: zero border applies beyond it (first slab of a volume cut at z = 2)
    part = fused_halo_affinities(vol[:, :, 2:], vol[:, :, :2])
    assert torch.equal(part, whole[:, :, 2:])


@pytest.mark.parametrize("shape,row0,band", [((1, 12, 64, 64), 8, 8), ((2, 9, 40, 96), 16, 16), ((1, 7, 72, 64), 0, 24),
                                             ((1, 6, 48, 64), 48, 8)])
def test_zshard_staged_halo_is_bit_identical(P, shape, row0, band):
    """Staged halo (ABI 7): a copy engine pulls the previous slab's tail band by band on a side stream while the one
    compute launch waits per tile row for its band's flag.  Emulated on one GPU (the "peer" tail is another local
    buffer), rewritten every iteration so that a stale landing buffer or a flag from the previous epoch would show."""
    from pea_b200.zshard import staged_halo_affinities, fused_halo_affinities, HALO

    class Slab:   # what staged_halo_affinities needs of a PeerZShard
        def __init__(self, own, tail):
            self.own, self.prev_tail, self._stage = own, tail, None

        def staging(self):
            if self._stage is None:
                B, C, _, H, W = self.own.shape
                self._stage = torch.full((B, C, HALO, H, W), float("nan"), device="cuda")
                self._side = torch.cuda.Stream()
            return self._stage, self._side

    B, D, H, W = shape
    buf = torch.empty(B, 16, D + HALO, H, W, device="cuda")
    slab = Slab(buf[:, :, HALO:], buf[:, :, :HALO])
    for it in range(3):
        torch.manual_seed(it)
        buf.normal_()
        whole = P.inf_embedding_loss_norm5(buf)
        out = staged_halo_affinities(slab, row0=row0, band_rows=band)
        assert torch.equal(out, whole[:, :, HALO:]), f"iteration {it}"
        assert torch.equal(out, fused_halo_affinities(slab.own, slab.prev_tail))
    if row0 < H:
        assert int(slab._flags.max()) == 3 and int(slab._flags[: -(-(H - row0) // band)].min()) == 3   # one epoch per call
        # the epoch wraps at 2^32: the kernel compares the signed distance flag - epoch
        slab._epoch = 0xFFFFFFFD
        slab._flags.fill_(-3)
        for it in range(5):
            buf.normal_()
            out = staged_halo_affinities(slab, row0=row0, band_rows=band)
            assert torch.equal(out, P.inf_embedding_loss_norm5(buf)[:, :, HALO:]), f"epoch wrap, iteration {it}"
        assert slab._epoch == 2 and int(slab._flags[0]) == 2


def test_fused_batch_gt_matches_oracle(P, O):
    """make_targets_2d / _3d (two fused passes, 4 voxels per thread) vs the oracle, per (sample, channel) weights."""
    offsets = O.multi_offset([1, 3, 5, 9, 27], neighbor=4) + [[2, 3], [0, 40]]
    labs = np.stack([O.voronoi_labels((52, 64), 9, 300 + b) for b in range(3)])
    labs[2] = 0                                     # an all-background sample: uniform channels -> weight 1
    tgt, msk, wgt = P.make_targets_2d(torch.tensor(labs).cuda(), offsets)
    for b in range(3):
        to, mo = O.gen_affs_ours(labs[b], offsets, ignore=False, padding=True)
        same(tgt[b], to, "2D target")
        same(msk[b], mo, "2D mask")
        same(wgt[b], np.stack([O.weight_binary_ratio(c) for c in to]), "2D weight")
    segs = np.stack([O.voronoi_labels((9, 40, 48), 14, 400 + b, z_scale=3.0, dtype=np.uint16) for b in range(2)])
    for dt in (np.int32, np.int64, np.uint8):
        tgt, wgt = P.make_targets_3d(torch.tensor((segs % 200).astype(dt)).cuda())
        for b in range(2):
            to = O.seg_to_aff_norm5((segs[b] % 200).astype(dt))
            same(tgt[b], to, f"3D target {dt.__name__}")
            same(wgt[b], np.stack([O.weight_binary_ratio(c) for c in to]), "3D weight")


@pytest.mark.parametrize("route", ["one_launch", "two_launch"])
def test_fused_batch_gt_routes_bit_identical(P, O, route):
    """pea_gt_targets has two routes: ONE cooperative launch (pass 1 parks its 4-bit results in shared memory, every
    sample has a completion counter, pass 2 expands the bits into weights) and the two-launch route (count + write,
    PEA_KNOB_GT_TWO_PASS) that volumes too large to be resident at once take.  Both against the oracle, bit for bit:
    8-bit labels (packed-byte path) and wider ones, rows that are / are not a multiple of eight voxels, samples
    whose channels are uniform, many slices per block (B = 5 samples of 96 x 136), replayed from a CUDA graph."""
    KNOB_GT_TWO_PASS = 3
    _knob(KNOB_GT_TWO_PASS, 1 if route == "two_launch" else 0)
    try:
        offsets = O.multi_offset([1, 3, 5, 9, 11], neighbor=4)
        for (B, H, W, dt) in ((5, 96, 136, np.uint8), (2, 52, 68, np.uint8), (3, 40, 64, np.int16), (2, 33, 36, np.int32)):
            labs = np.stack([O.voronoi_labels((H, W), 9, 500 + b, background_frac=0.3, dtype=np.uint8) for b in range(B)]).astype(dt)
            labs[B - 1] = 0
            lab_d = torch.tensor(labs).cuda()
            for ignore in (False, True):
                tgt, msk, wgt = P.make_targets_2d(lab_d, offsets, ignore=ignore, padding=True)
                for b in range(B):
                    to, mo = O.gen_affs_ours(labs[b], offsets, ignore=ignore, padding=True)
                    same(tgt[b], to, f"2D target {route} {dt.__name__} {H}x{W}")
                    same(msk[b], mo, "2D mask")
                    same(wgt[b], np.stack([O.weight_binary_ratio(c) for c in to]), "2D weight")
        segs = np.stack([O.voronoi_labels((10, 40, 48), 14, 600 + b, z_scale=3.0, dtype=np.uint16) for b in range(3)])
        for dt in (np.int32, np.uint8):
            seg_d = torch.tensor((segs % 200).astype(dt)).cuda()
            tgt, wgt = P.make_targets_3d(seg_d)
            for b in range(3):
                to = O.seg_to_aff_norm5((segs[b] % 200).astype(dt))
                same(tgt[b], to, f"3D target {route} {dt.__name__}")
                same(wgt[b], np.stack([O.weight_binary_ratio(c) for c in to]), "3D weight")
        # captured and replayed: the counters behind the counts are zeroed by the captured memset every time
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            P.make_targets_3d(seg_d)
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g):
            t2, w2 = P.make_targets_3d(seg_d)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        assert torch.equal(t2, tgt) and torch.equal(w2, wgt), "graph replay differs"
    finally:
        _knob(KNOB_GT_TWO_PASS, 0)


@pytest.mark.parametrize("route", ["one_launch", "two_launch"])
@pytest.mark.parametrize("B,D,H,W,dt,semantics", [(3, 1, 41, 72, np.uint8, 0), (2, 1, 30, 44, np.int32, 0), (2, 5, 24, 40, np.int32, 1),
                                                   (1, 3, 17, 36, np.uint8, 1)])
def test_gt_kernels_write_only_their_outputs(P, O, route, B, D, H, W, dt, semantics):
    """Guard bands around target, mask, weight and the counts scratch of pea_gt_targets (compute-sanitizer is closed on the
    GPU pool): buffers carved out of larger poisoned ones and handed to the C ABI directly, both routes, rows that are /
    are not a multiple of eight voxels, last slices only partly filled.  Nothing outside the outputs may change, every
    element inside must be written, and the results must equal the oracle's bit for bit."""
    import ctypes
    from pea_b200 import _ops as ops
    from pea_b200._lib import lib, check
    _knob(3, 1 if route == "two_launch" else 0)  # PEA_KNOB_GT_TWO_PASS
    try:
        if semantics == 0:
            offsets = O.multi_offset([1, 3, 5, 9, 11], neighbor=4)
            offs3 = ops.offsets3(offsets, 2)
            labs = np.stack([O.voronoi_labels((H, W), 7, 80 + b, background_frac=0.3, dtype=np.uint8) for b in range(B)]).astype(dt)[:, None]
        else:
            from pea_b200.gt import NHOOD_NORM5
            offs3 = np.asarray(NHOOD_NORM5, dtype=np.int32)
            labs = np.stack([O.voronoi_labels((D, H, W), 7, 90 + b, z_scale=3.0, dtype=np.uint16) for b in range(B)])
            labs = (labs % 200).astype(dt)
        K = offs3.shape[0]
        n = B * K * D * H * W
        PAD = 4096
        lab_d = torch.tensor(labs).cuda().contiguous()

        def guarded(count, dtype, poison):
            buf = torch.full((count + 2 * PAD,), poison, dtype=dtype, device="cuda")
            return buf, buf[PAD:PAD + count]

        t_buf, tgt = guarded(n, torch.float32, -12345.0)
        w_buf, wgt = guarded(n, torch.float32, -12345.0)
        m_buf, msk = guarded(n, torch.uint8, 77)
        nbytes = lib.pea_gt_counts_bytes(B, K)
        c_buf, cnt = guarded(nbytes // 8, torch.int64, -7)
        offs = (ctypes.c_int32 * (3 * K))(*[int(v) for v in np.asarray(offs3).reshape(-1)])
        rc = lib.pea_gt_targets(lab_d.data_ptr(), ops.label_code(lab_d), B, D, H, W, offs, K, semantics, 0, 1, 1.0,
                                tgt.data_ptr(), msk.data_ptr() if semantics == 0 else None, wgt.data_ptr(), cnt.data_ptr(), nbytes,
                                torch.cuda.current_stream().cuda_stream)
        check("pea_gt_targets", rc)
        torch.cuda.synchronize()
        for name, buf, poison in (("target", t_buf, -12345.0), ("weight", w_buf, -12345.0), ("counts", c_buf, -7)) + \
                ((("mask", m_buf, 77),) if semantics == 0 else ()):
            assert bool((buf[:PAD] == poison).all()) and bool((buf[-PAD:] == poison).all()), f"{name}: guard band overwritten"
        assert not bool((tgt == -12345.0).any()) and not bool((wgt == -12345.0).any()), "elements left unwritten"
        tgt5, wgt5 = tgt.view(B, K, D, H, W).cpu().numpy(), wgt.view(B, K, D, H, W).cpu().numpy()
        for b in range(B):
            if semantics == 0:
                to, mo = O.gen_affs_ours(labs[b, 0], offsets, ignore=False, padding=True)
                same(tgt5[b, :, 0], to, "target")
                same(msk.view(B, K, D, H, W)[b, :, 0], mo, "mask")
                same(wgt5[b, :, 0], np.stack([O.weight_binary_ratio(c) for c in to]), "weight")
            else:
                to = O.seg_to_aff_norm5(labs[b])
                same(tgt5[b], to, "target")
                same(wgt5[b], np.stack([O.weight_binary_ratio(c) for c in to]), "weight")
    finally:
        _knob(3, 0)


# =========================================================================== several small losses in one launch
def test_norm1_multi_matches_single_calls_and_oracle(P, O):
    """embedding_loss_norm1_multi (ONE launch for the four deep-supervision scales, scripts_ac3ac4/main.py:227-230)
    against the oracle's four calls: losses, affinities, gradients; upstream scalars other than 1 in backward."""
    embs, tgts, wgts = [], [], []
    for i, (B, D, H, W) in enumerate(((2, 6, 5, 5), (2, 6, 10, 10), (2, 6, 20, 20), (2, 6, 40, 40))):
        emb, _, tgt, wgt = _inputs_3d(O, B, 16, D, H, W, 3, 40 + i, replicate=True)
        embs.append(emb), tgts.append(tgt), wgts.append(wgt)
    scal = [1.0, 0.5, 2.0, 1.0]
    eo = [torch.tensor(e, requires_grad=True) for e in embs]
    ro = [O.embedding_loss_norm1(e, torch.tensor(t), torch.tensor(w), O.WeightedMSE(), affs0_weight=1.5)
          for e, t, w in zip(eo, tgts, wgts)]
    sum(c * r[0] for c, r in zip(scal, ro)).backward()
    e = [dev(x, True) for x in embs]
    r = P.embedding_loss_norm1_multi(e, [dev(t) for t in tgts], [dev(w) for w in wgts], P.WeightedMSE(), affs0_weight=1.5)
    sum(c * x[0] for c, x in zip(scal, r)).backward()
    for i in range(4):
        close_s(r[i][0], ro[i][0], what=f"scale {i} loss")
        close_t(r[i][1], ro[i][1], what=f"scale {i} affs")
        assert not r[i][1].requires_grad
        close_t(e[i].grad, eo[i].grad, what=f"scale {i} grad")


def test_embedding_loss_multi_2d_matches_oracle(P, O):
    """embedding_loss_multi (scripts_cvppp/main.py:284-287: coarser scales use the first 2 i offsets) against the oracle,
    including a gradient that arrives through the returned affinities and a float mask."""
    offsets = O.multi_offset([1, 3, 5, 9], neighbor=4)
    shapes = ((2, 17, 17, 2), (2, 34, 34, 4), (1, 68, 60, 6), (1, 40, 72, 8))
    ins = [_inputs_2d(O, B, 16, H, W, offsets[:K], 50 + i) for i, (B, H, W, K) in enumerate(shapes)]
    eo = [torch.tensor(x[0], requires_grad=True) for x in ins]
    ro = [O.embedding_loss(e, torch.tensor(x[2]), torch.tensor(x[3]), torch.tensor(x[4]), O.WeightedMSE(), offsets[:K])
          for e, x, (_, _, _, K) in zip(eo, ins, shapes)]
    up = torch.tensor(np.random.RandomState(3).randn(*ins[1][2].shape).astype(np.float32))
    (sum(r[0] for r in ro) + (ro[1][1] * up).sum()).backward()
    e = [dev(x[0], True) for x in ins]
    masks = [dev(x[4]) if i != 2 else dev(x[4]).float() for i, x in enumerate(ins)]
    r = P.embedding_loss_multi(e, [dev(x[2]) for x in ins], [dev(x[3]) for x in ins], masks, P.WeightedMSE(),
                               [offsets[:K] for (_, _, _, K) in shapes])
    (sum(x[0] for x in r) + (r[1][1] * up.cuda()).sum()).backward()
    for i in range(4):
        close_s(r[i][0], ro[i][0], what=f"scale {i} loss")
        close_t(r[i][1], ro[i][1], what=f"scale {i} affs")
        close_t(np.asarray(list(r[i][2])), np.asarray(ro[i][2]), what=f"scale {i} all_loss")
        close_t(e[i].grad, eo[i].grad, what=f"scale {i} grad")
    # graph capture + replay gives the same numbers (the ticket word is left zero by every launch)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        P.embedding_loss_multi(e, [dev(x[2]) for x in ins], [dev(x[3]) for x in ins], masks, P.WeightedMSE(),
                               [offsets[:K] for (_, _, _, K) in shapes])
    torch.cuda.current_stream().wait_stream(side)
    tg, wg = [dev(x[2]) for x in ins], [dev(x[3]) for x in ins]
    with torch.cuda.graph(g):
        r2 = P.embedding_loss_multi([x.detach() for x in e], tg, wg, masks, P.WeightedMSE(), [offsets[:K] for (_, _, _, K) in shapes])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for i in range(4):
        assert float(r2[i][0]) == float(r[i][0]), f"scale {i}: graph replay differs"


def test_label_pyramid_matches_oracle_recipe(P, O):
    """make_targets_pyramid_2d / _3d: all five scales' ground truth from ONE label upload, bit for bit what the
    reference's loaders build per sample (data_provider.py:198-239, data_provider_labeled_deep.py:224-285)."""
    offsets = O.multi_offset([1, 3, 5, 9, 27], neighbor=4)
    labs = np.stack([O.voronoi_labels((64, 96), 9, 70 + b) for b in range(2)])
    (t, m, w), downs = P.make_targets_pyramid_2d(torch.tensor(labs).cuda(), offsets, nb_half=2)
    for b in range(2):
        to, mo = O.gen_affs_ours(labs[b], offsets, ignore=False, padding=True)
        same(t[b], to, "full target"), same(m[b], mo, "full mask")
        same(w[b], np.stack([O.weight_binary_ratio(c) for c in to]), "full weight")
        for i, (f, c) in enumerate(zip((2, 4, 8, 16), (4, 3, 2, 1))):
            K = 2 * c
            td, md = O.gen_affs_ours(labs[b][::f, ::f], offsets[:K], ignore=False, padding=True)
            wd = np.stack([O.weight_binary_ratio(ch) for ch in td])
            same(downs[i][b], np.concatenate([td, wd, md.astype(np.float32)], axis=0), f"down{i + 1}")
    segs = np.stack([O.voronoi_labels((6, 32, 48), 10, 80 + b, z_scale=4.0, dtype=np.uint16) for b in range(2)])
    (t, w), downs = P.make_targets_pyramid_3d(torch.tensor(segs.astype(np.int32)).cuda())
    for b in range(2):
        to = O.seg_to_aff_norm5(segs[b])
        same(t[b], to, "3D full target")
        same(w[b], np.stack([O.weight_binary_ratio(c) for c in to]), "3D full weight")
        for i, f in enumerate((2, 4, 8, 16)):
            td = O.seg_to_aff(segs[b][:, ::f, ::f], pad='replicate')
            wd = np.stack([O.weight_binary_ratio(ch) for ch in td])
            same(downs[i][b], np.concatenate([td, wd], axis=0), f"3D down{i + 1}")


def test_volume_blender_matches_provider_valid(P, O):
    """Device-side add_vol / get_results (scripts_ac3ac4/data/provider_valid.py:294-340) and the Mutex Watershed's
    `1 - affs` (inference.py:196): overlapping patches straight from the inference kernel, bit for bit NumPy's result."""
    from pea_b200.provider_valid import VolumeBlender, get_weight
    K, vol, patch = 12, (10, 72, 88), (6, 40, 48)
    same(get_weight(patch), O.blend_get_weight(patch), "window")
    rng = np.random.RandomState(5)
    emb = rng.randn(1, 16, *vol).astype(np.float32)
    bl = VolumeBlender(K, vol, patch, "cuda")
    acc, wm = np.zeros((K,) + vol, np.float32), np.zeros((1,) + vol, np.float32)
    wv = O.blend_get_weight(patch)
    e = dev(emb)
    for z0 in (0, 4):
        for y0 in (0, 16, 32):
            for x0 in (0, 20, 40):
                sub = e[:, :, z0:z0 + patch[0], y0:y0 + patch[1], x0:x0 + patch[2]].contiguous()
                pred = P.inf_embedding_loss_norm5(sub, relu=True, replicate_border=True)
                bl.add_vol(pred, (z0, y0, x0))
                O.blend_add_vol(acc, wm, pred[0].cpu().numpy(), wv, (z0, y0, x0))
    same(bl.out_affs, acc, "accumulated affinities")
    same(bl.weight_map, wm, "accumulated window")
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = O.blend_get_results(acc, wm, (1, 4, 4))
    same(bl.get_results((1, 4, 4)), ref, "get_results")
    same(bl.get_results((1, 4, 4), one_minus=True), (np.float32(1) - ref).astype(np.float32), "1 - affs")


def test_consistency_flip_golden_and_oracle(P, O, golden):
    """convert_consistency_flip in one device pass: the reference's goldens (every rule combination), then the oracle on
    larger / non-multiple-of-32 / non-square (no transpose) extents, and the flip followed by the two-view loss."""
    g = golden["flip"]
    for nd in ("2d", "3d"):
        out = P.convert_consistency_flip(dev(g["in" + nd]), torch.tensor(g["rules" + nd]).cuda())
        same(out, g["out" + nd], "golden " + nd)
        assert not out.requires_grad
    rng = np.random.RandomState(9)
    for shape, nr in (((8, 16, 70, 70), 3), ((16, 4, 3, 45, 45), 4), ((4, 16, 5, 33, 64), 4), ((4, 3, 37, 50), 3)):
        x = rng.randn(*shape).astype(np.float32)
        rules = rng.randint(0, 2, (shape[0], nr))
        if shape[-1] != shape[-2]:
            rules[:, -1] = 0
        same(P.convert_consistency_flip(dev(x), torch.tensor(rules).cuda()),
             O.convert_consistency_flip(torch.tensor(x), torch.tensor(rules)).numpy(), f"oracle {shape}")
    with pytest.raises(ValueError):
        P.convert_consistency_flip(dev(rng.randn(2, 3, 8, 12).astype(np.float32)), torch.tensor([[0, 0, 1], [0, 0, 0]]).cuda())
    # flip -> two-view loss, as scripts_ac3ac4/main.py:213-222 chains them
    emb, ema, tgt, wgt = _inputs_3d(O, 2, 16, 6, 32, 32, 12, 31)
    rules = torch.tensor([[1, 0, 1, 1], [0, 1, 0, 0]])
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao = O.ema_embedding_loss_norm5(eo, O.convert_consistency_flip(torch.tensor(ema), rules), torch.tensor(tgt), torch.tensor(wgt),
                                        O.WeightedMSE(), affs0_weight=2)
    lo.backward()
    e = dev(emb, True)
    l, a = P.ema_embedding_loss_norm5(e, P.convert_consistency_flip(dev(ema), rules.cuda()), dev(tgt), dev(wgt), P.WeightedMSE(),
                                      affs0_weight=2)
    l.backward()
    close_s(l, lo, what="flip + ema loss")
    close_t(a, ao, what="flip + ema affs")
    close_t(e.grad, eo.grad, what="flip + ema grad")


def test_cosine_mode_and_masked_weights_and_wide_labels(P, O):
    """The rarely used corners of the reference's signatures: dis_mode != 'ours' (nn.CosineSimilarity on un-normalised
    embeddings, loss_embedding_mse.py:11-13), weight_binary_ratio(mask=...) (data_segmentation.py:215-225), int64 labels
    with ids beyond int32 in the label-driven losses, and embedding2affs(relu=True)."""
    offsets = O.multi_offset([1, 3], neighbor=4)
    emb, ema, tgt, wgt, msk = _inputs_2d(O, 2, 16, 20, 24, offsets, 61)
    eo = torch.tensor(emb, requires_grad=True)
    lo, ao, allo = O.embedding_loss(eo, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), O.WeightedMSE(), offsets, mode='cos')
    lo.backward()
    e = dev(emb, True)
    l, a, al = P.embedding_loss(e, dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets, mode='cos')
    l.backward()
    close_s(l, lo, what="cosine loss")
    close_t(a, ao, what="cosine affs")
    close_t(e.grad, eo.grad, what="cosine grad")
    lo2, _ = O.ema_embedding_loss(torch.tensor(emb), torch.tensor(ema), torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk),
                                  O.WeightedMSE(), offsets, affs0_weight=2, mode='cos')
    l2, _ = P.ema_embedding_loss(dev(emb), dev(ema), dev(tgt), dev(wgt), dev(msk), P.WeightedMSE(), offsets, affs0_weight=2, mode='cos')
    close_s(l2, lo2, what="cosine ema loss")
    close_t(P.embedding2affs(dev(emb), offsets, mode='cos'), O.embedding2affs(torch.tensor(emb), offsets, mode='cos'), what="cosine embedding2affs")
    close_t(P.embedding2affs(dev(emb), offsets, relu=True), torch.relu(O.embedding2affs(torch.tensor(emb), offsets)), what="relu epilogue 2D")
    # masked class-balance weights (0/1 mask: bit-exact)
    rng = np.random.RandomState(4)
    lab = (rng.rand(30, 40) > 0.7).astype(np.float32)
    m = (rng.rand(30, 40) > 0.3).astype(np.float32)
    same(P.weight_binary_ratio(lab, mask=m), O.weight_binary_ratio(lab, mask=m), "masked weights (numpy in)")
    same(P.weight_binary_ratio(torch.tensor(lab).cuda(), mask=torch.tensor(m).cuda()), O.weight_binary_ratio(lab, mask=m), "masked weights")
    # int64 labels whose ids differ by 2^32: equal after a cast to int32, different in the reference
    lab64 = np.where(O.voronoi_labels((32, 48), 6, 3) % 2 == 0, 5, 5 + (1 << 32)).astype(np.int64)
    offs = O.multi_offset([1, 3, 5, 9, 11], neighbor=4)
    x = rng.randn(1, 16, 32, 48).astype(np.float32)
    to, mo = O.gen_affs_ours(lab64, offs, ignore=False, padding=True)
    wo = np.stack([O.weight_binary_ratio(c) for c in to])
    lo3, _, _ = O.embedding_loss(torch.tensor(x), torch.tensor(to[None]), torch.tensor(wo[None]), torch.tensor(mo[None]), O.WeightedMSE(), offs)
    l3, _, _ = P.embedding_loss_labels(dev(x), torch.tensor(lab64[None]).cuda(), P.WeightedMSE(), offs)
    close_s(l3, lo3, what="int64 labels beyond int32")


def test_multi_and_flip_write_only_their_outputs(P, O):
    """Guard bands (compute-sanitizer is not available on the GPU pool): the multi-problem launch and the flip kernel get
    output buffers carved out of larger poisoned ones; every byte outside must stay untouched, every element inside
    must be written (partial 32-voxel chunks, partial 32 x 32 flip tiles)."""
    POISON, PAD = -12345.0, 2048

    def guarded(shape):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * PAD,), POISON, dtype=torch.float32, device="cuda")
        return buf, buf[PAD:PAD + n].view(shape)

    from pea_b200 import _ops as ops
    from pea_b200._lib import lib, check, LossProblem
    import ctypes
    shapes = [(1, 16, 3, 5, 7), (2, 16, 3, 9, 11)]   # 105 and 594 voxels: neither a multiple of 32
    g = torch.Generator().manual_seed(3)
    offs = (ctypes.c_int32 * 9)(-1, 0, 0, 0, -1, 0, 0, 0, -1)
    nrm = (ctypes.c_float * 3)(7.0, 7.0, 7.0)
    cf = (ctypes.c_float * 3)(1.0, 1.0, 1.0)
    arr = (LossProblem * 2)()
    keep, bufs = [], []
    for i, shp in enumerate(shapes):
        B, C, D, H, W = shp
        emb = torch.randn(shp, generator=g).cuda()
        tgt = (torch.rand((B, 3, D, H, W), generator=g) > 0.5).float().cuda()
        wgt = (torch.rand((B, 3, D, H, W), generator=g) + 0.1).cuda()
        ab, affs = guarded((B, 3, D, H, W))
        gb, grad = guarded(shp)
        lb, lout = guarded((4,))
        keep += [emb, tgt, wgt]
        bufs += [(ab, affs), (gb, grad), (lb, lout)]
        a = arr[i]
        a.emb, a.target, a.weight, a.affs, a.grad_emb, a.loss_out = (emb.data_ptr(), tgt.data_ptr(), wgt.data_ptr(), affs.data_ptr(),
                                                                     grad.data_ptr(), lout.data_ptr())
        a.offsets, a.loss_norm, a.loss_coef = offs, nrm, cf
        a.B, a.D, a.H, a.W, a.K, a.border, a.mask_dtype = B, D, H, W, 3, ops.BORDER_ZERO, ops.MASK_NONE
    nws = lib.pea_affs_loss_multi_workspace(arr, 2)
    wb, ws = guarded(((nws + 3) // 4,))
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    check("pea_affs_loss_multi", lib.pea_affs_loss_multi(arr, 2, 16, ws.data_ptr(), nws, ticket.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(ticket) == 0, "the ticket word must be left zero"
    for buf, view in bufs + [(wb, ws)]:
        n = view.numel()
        assert bool((buf[:PAD] == POISON).all()) and bool((buf[PAD + n:] == POISON).all()), "guard band overwritten"
    for buf, view in bufs:
        assert not bool((view == POISON).any()) and bool(torch.isfinite(view).all()), "output element left unwritten"
    x = torch.randn((3, 2, 2, 37, 37), generator=g).cuda()
    ob, out = guarded(tuple(x.shape))
    rules = torch.tensor([[1, 1, 0, 1], [0, 0, 1, 1], [1, 0, 0, 0]], dtype=torch.int32).cuda()
    check("pea_consistency_flip", lib.pea_consistency_flip(x.data_ptr(), rules.data_ptr(), 3, 2, 2, 37, 37, out.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert bool((ob[:PAD] == POISON).all()) and bool((ob[PAD + out.numel():] == POISON).all()), "flip: guard band overwritten"
    assert torch.equal(out.cpu(), O.convert_consistency_flip(x.cpu(), rules.cpu()))


@pytest.mark.parametrize("B,D,H,W", [(2, 7, 33, 44), (1, 18, 40, 36)])
def test_self_and_cross_share_one_gradient_buffer(P, O, B, D, H, W):
    """embedding_loss_norm5_self_and_cross == embedding_loss_norm5 + ema_embedding_loss_norm5 on the same embedding
    (scripts_ac3ac4/main.py:223-224): the two-view kernel ADDS into the one-view kernel's gradient buffer
    (PEA_EPI_GRAD_ACCUMULATE).  Against the oracle's two calls: parts, prediction, total, gradient; an upstream scalar."""
    emb, ema, tgt, wgt = _inputs_3d(O, B, 16, D, H, W, 12, 77)
    eo = torch.tensor(emb, requires_grad=True)
    l0, pred_o = O.embedding_loss_norm5(eo, torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(), affs0_weight=2)
    l1, _ = O.ema_embedding_loss_norm5(eo, torch.tensor(ema), torch.tensor(tgt), torch.tensor(wgt), O.WeightedMSE(), affs0_weight=2)
    (3.0 * (0.7 * l0 + 1.3 * l1)).backward()
    e = dev(emb, True)
    total, ls, pred, lc = P.embedding_loss_norm5_self_and_cross(e, dev(ema), dev(tgt), dev(wgt), P.WeightedMSE(), affs0_weight=2,
                                                                coef_self=0.7, coef_cross=1.3)
    (3.0 * total).backward()
    close_s(ls, l0, what="one-view part")
    close_s(lc, l1, what="two-view part")
    close_s(total, 0.7 * float(l0) + 1.3 * float(l1), what="total")
    close_t(pred, pred_o, what="prediction")
    close_t(e.grad, eo.grad, what="shared gradient buffer")
    assert not pred.requires_grad and not ls.requires_grad
    # the ABI flag is refused where no kernel implements it
    from pea_b200 import _ops as ops
    from pea_b200.loss_embedding_mse_3d import offsets_from_shifts
    with pytest.raises(ops.PeaError):
        ops.affs_loss_fwd_bwd(dev(emb), None, offsets_from_shifts([1, 1, 1]), ops.BORDER_ZERO, dev(tgt)[:, :3].contiguous(),
                              dev(wgt)[:, :3].contiguous(), None, np.ones(3, np.float32), np.ones(3, np.float32),
                              grad_into=torch.zeros_like(dev(emb)))
