"""Pin the CPU oracle (oracle/pea_oracle.py) against outputs of the real reference.

The fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs the
unmodified reference from /root/reference.  Everything here is CPU-only.
"""
import numpy as np
import pytest
import torch

from oracle import pea_oracle as O

torch.set_num_threads(1)


def T(a, grad=False):
    t = torch.tensor(np.asarray(a))
    return t.requires_grad_(True) if grad else t


def same(a, b):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


def close(a, b, rtol=1e-6, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


# ------------------------------------------------------------------ 2D loss family
def test_embedding_loss_2d(golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    e = T(g["embedding"], True)
    loss, affs, all_loss = O.embedding_loss(e, T(g["target"]), T(g["weightmap"]), T(g["mask"]), O.WeightedMSE(),
                                            offsets, affs0_weight=3)
    loss.backward()
    # same op sequence as the reference on the same torch build -> bit-identical
    same(loss.detach(), g["loss"])
    same(affs.detach(), g["affs"])
    same(np.asarray(all_loss), g["all_loss"])
    same(e.grad, g["grad"])
    same(O.embedding2affs(T(g["embedding"]), offsets), g["embedding2affs"])
    same(g["embedding2affs"], g["affs"])  # embedding2affs == embedding_loss(...)[1]


def test_embedding_loss_2d_float_mask_and_affs_grad(golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    e = T(g["embedding"], True)
    loss, _, _ = O.embedding_loss(e, T(g["target"]), T(g["weightmap"]), T(g["mask"]).float(), O.WeightedMSE(), offsets)
    same(loss.detach(), g["loss_fmask"])
    e = T(g["embedding"], True)
    loss, affs, _ = O.embedding_loss(e, T(g["target"]), T(g["weightmap"]), T(g["mask"]), O.WeightedMSE(), offsets)
    (0.5 * loss + (affs * T(g["affs_upstream"])).sum()).backward()
    close(e.grad, g["grad_with_affs"], rtol=1e-5, atol=1e-6)


def test_ema_embedding_loss_2d(golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    e, u = T(g["embedding"], True), T(g["ema_embedding"], True)
    loss, affs = O.ema_embedding_loss(e, u, T(g["target"]), T(g["weightmap"]), T(g["mask"]), O.WeightedMSE(), offsets,
                                      affs0_weight=3)
    loss.backward()
    same(loss.detach(), g["ema_loss"])
    same(affs.detach(), g["ema_affs"])
    same(e.grad, g["ema_grad"])
    same(u.grad, g["ema_grad_partner"])


def test_closed_form_backward_2d(golden):
    g = golden["loss2d"]
    offsets = g["offsets"].tolist()
    loss, affs, grad = O.closed_form_grad_2d(g["embedding"], g["target"], g["weightmap"], g["mask"], offsets)
    zero = (g["embedding"] == 0).all(axis=1, keepdims=True)  # eps path of F.normalize: huge, excluded here
    sel = np.broadcast_to(~zero, grad.shape)
    scale = np.abs(g["grad"][sel]).max()
    close(loss, g["loss"], rtol=2e-6)
    close(affs, g["affs"], rtol=0, atol=2e-6)
    close(grad[sel], g["grad"][sel], rtol=0, atol=2e-6 * scale)
    loss, affs, grad = O.closed_form_grad_2d(g["embedding"], g["target"], g["weightmap"], g["mask"], offsets,
                                             partner=g["ema_embedding"], affs0_weight=3)
    close(loss, g["ema_loss"], rtol=2e-6)
    close(grad[sel], g["ema_grad"][sel], rtol=0, atol=2e-6 * np.abs(g["ema_grad"][sel]).max())


# ------------------------------------------------------------------ 3D loss family
def test_norm5_family(golden):
    g = golden["loss3d"]
    crit = O.WeightedMSE()
    e = T(g["A_embedding"], True)
    loss, affs = O.embedding_loss_norm5(e, T(g["A_target"]), T(g["A_weightmap"]), crit, affs0_weight=2)
    loss.backward()
    same(loss.detach(), g["A_n5_loss"])
    same(affs, g["A_n5_affs"])
    same(e.grad, g["A_n5_grad"])
    assert not affs.requires_grad  # detached in the reference (:186-190)

    e, u = T(g["A_embedding"], True), T(g["A_ema_embedding"], True)
    loss, affs = O.ema_embedding_loss_norm5(e, u, T(g["A_target"]), T(g["A_weightmap"]), crit, affs0_weight=2)
    loss.backward()
    same(loss.detach(), g["A_ema5_loss"])
    same(affs, g["A_ema5_affs"])
    same(e.grad, g["A_ema5_grad"])
    same(u.grad, g["A_ema5_grad_partner"])
    same(O.inf_embedding_loss_norm5(T(g["A_embedding"])), g["A_inf5"])


def test_norm1_family(golden):
    g = golden["loss3d"]
    crit = O.WeightedMSE()
    e = T(g["B_embedding"], True)
    loss, affs = O.embedding_loss_norm1(e, T(g["B_target"]), T(g["B_weightmap"]), crit, affs0_weight=2)
    loss.backward()
    same(loss.detach(), g["B_n1_loss"])
    same(affs, g["B_n1_affs"])
    same(e.grad, g["B_n1_grad"])

    e = T(g["B_embedding"], True)
    loss, affs = O.embedding_loss_norm1(e, T(g["B_target"]), T(g["B_weightmap"]), crit, affs0_weight=0.5, shift=2)
    loss.backward()
    same(loss.detach(), g["B_n1s2_loss"])
    same(affs, g["B_n1s2_affs"])
    same(e.grad, g["B_n1s2_grad"])

    e, u = T(g["B_embedding"], True), T(g["B_ema_embedding"], True)
    loss, affs = O.ema_embedding_loss_norm1(e, u, T(g["B_target"]), T(g["B_weightmap"]), crit, affs0_weight=2)
    loss.backward()
    same(loss.detach(), g["B_ema1_loss"])
    same(affs, g["B_ema1_affs"])
    same(e.grad, g["B_ema1_grad"])
    same(u.grad, g["B_ema1_grad_partner"])
    same(O.inf_embedding_loss_norm1(T(g["B_embedding"])), g["B_inf1"])


def test_closed_form_backward_3d(golden):
    g = golden["loss3d"]
    loss, affs, grad = O.closed_form_grad_3d(g["A_embedding"], g["A_target"], g["A_weightmap"], O.SHIFTS_NORM5, 3,
                                             affs0_weight=2)
    close(loss, g["A_n5_loss"], rtol=2e-6)
    close(affs, g["A_n5_affs"], rtol=0, atol=2e-6)
    close(grad, g["A_n5_grad"], rtol=0, atol=2e-6 * np.abs(g["A_n5_grad"]).max())
    loss, affs, grad = O.closed_form_grad_3d(g["A_embedding"], g["A_target"], g["A_weightmap"], O.SHIFTS_NORM5, 3,
                                             partner=g["A_ema_embedding"], affs0_weight=2)
    close(loss, g["A_ema5_loss"], rtol=2e-6)
    close(grad, g["A_ema5_grad"], rtol=0, atol=2e-6 * np.abs(g["A_ema5_grad"]).max())


# ------------------------------------------------------------------ ground truth (bit exact)
def test_offset_tables(golden):
    g = golden["gt"]
    same(np.asarray(O.multi_offset([1, 3, 5, 9, 27], neighbor=4)), g["multi_offset_cvppp"])
    same(np.asarray(O.multi_offset([1, 2], neighbor=8)), g["multi_offset_n8"])
    same(O.mknhood3d(1), g["mknhood3d"])
    same(O.NHOOD_NORM5[:3], g["mknhood3d"])


def test_gen_affs_ours(golden):
    g = golden["gt"]
    lab, offsets = g["gt2d_labels"], g["gt2d_offsets"].tolist()
    for ig in (False, True):
        for pd in (False, True):
            a, m = O.gen_affs_ours(lab, offsets, ignore=ig, padding=pd)
            assert a.dtype == np.float32 and m.dtype == np.uint8
            same(a, g[f"gt2d_aff_i{int(ig)}_p{int(pd)}"])
            same(m, g[f"gt2d_msk_i{int(ig)}_p{int(pd)}"])
    a, m = O.gen_affs_ours(lab.astype(np.float32), offsets, ignore=False, padding=True)
    same(a, g["gt2d_aff_float_labels"])
    same(m, g["gt2d_msk_float_labels"])


def test_seg_to_aff(golden):
    g = golden["gt"]
    lab3 = g["gt3d_labels"]
    same(O.seg_to_aff_norm5(lab3), g["gt3d_aff12"])
    same(O.seg_to_aff(lab3), g["gt3d_aff3_replicate"])
    same(O.seg_to_aff(lab3, pad=''), g["gt3d_aff3_nopad"])
    same(O.seg_to_aff(g["gt2d_labels"], np.asarray([[-1, 0], [0, -1]])), g["gt3d_aff_2dseg_replicate"])


def test_weight_binary_ratio(golden):
    g = golden["gt"]
    for cin, cout in zip(g["wbr_in"], g["wbr_out"]):
        w = O.weight_binary_ratio(cin)
        assert w.dtype == np.float32
        same(w, cout)
    same(np.stack([O.weight_binary_ratio(c) for c in g["gt3d_aff12"]]), g["gt3d_w12"])
    same(np.stack([O.weight_binary_ratio(c) for c in g["gt2d_aff_i0_p1"]]), g["gt2d_w"])


def test_pyramid_downsampling_is_strided_sampling():
    """The reference's label pyramid (cv2.resize INTER_NEAREST by 1/2 .. 1/16, scripts_cvppp/data/data_provider.py:199-202;
    the (H,W,D)-transposed 3-D call, scripts_ac3ac4/data/data_provider_labeled_deep.py:224-232) equals label[::f, ::f]
    for extents divisible by f -- what pea_b200.make_targets_pyramid_* does on the device."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for (H, W) in ((544, 544), (160, 160), (704, 704), (256, 256)):
        lab = rng.randint(0, 50, (H, W)).astype(np.float32)
        for f in (2, 4, 8, 16):
            d = cv2.resize(lab, (0, 0), fx=1 / f, fy=1 / f, interpolation=cv2.INTER_NEAREST)
            assert np.array_equal(d, lab[::f, ::f]), (H, W, f)
    lb = rng.randint(0, 50, (18, 160, 160)).astype(np.uint16)
    t = np.transpose(lb, (1, 2, 0))
    for f in (2, 4, 8, 16):
        d = np.transpose(cv2.resize(t, (0, 0), fx=1 / f, fy=1 / f, interpolation=cv2.INTER_NEAREST), (2, 0, 1))
        assert np.array_equal(d, lb[:, ::f, ::f]), f


def test_oracle_consistency_flip_matches_reference_golden(golden):
    """oracle.convert_consistency_flip against the real reference's outputs for every rule combination
    (scripts_cvppp/data/data_consistency.py:34-45, scripts_ac3ac4/utils/consistency_aug.py:217-228)."""
    import torch
    from oracle import pea_oracle as O
    g = golden["flip"]
    for nd in ("2d", "3d"):
        out = O.convert_consistency_flip(torch.tensor(g["in" + nd]), torch.tensor(g["rules" + nd]))
        assert np.array_equal(out.numpy(), g["out" + nd]), nd
