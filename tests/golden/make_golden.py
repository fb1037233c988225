"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (weih527/Pixel-Embedded-Affinity).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported by path, unmodified, with the two shims SURVEY.md 8(c)
documents: ``torch.Tensor.cuda`` becomes a no-op on this GPU-less host (WeightedMSE calls
``.cuda()`` on its normaliser, scripts_cvppp/loss/loss.py:116) and stub ``skimage`` modules are
registered so ``data/data_segmentation.py`` imports.  Inputs are small seeded tensors; inputs
and reference outputs are stored together so the fixtures do not depend on any RNG stream.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PEA_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU host: see module docstring
    for n in ("skimage", "skimage.morphology", "skimage.measure", "skimage.segmentation"):
        m = types.ModuleType(n)
        m.__getattr__ = lambda attr: None
        sys.modules.setdefault(n, m)
    sys.path.insert(0, os.path.join(REF, "scripts_cvppp"))  # the reference uses cwd-relative imports
    ref = types.SimpleNamespace()
    ref.loss2d = _load("ref_loss2d", "scripts_cvppp/loss/loss_embedding_mse.py")
    ref.loss3d = _load("ref_loss3d", "scripts_ac3ac4/loss/loss_embedding_mse.py")
    ref.crit = _load("ref_crit", "scripts_cvppp/loss/loss.py")
    ref.aff2d = _load("ref_aff2d", "scripts_cvppp/utils/affinity_ours.py")
    ref.aff3d = _load("ref_aff3d", "scripts_cvppp/data/data_affinity.py")
    ref.seg = _load("ref_seg", "scripts_cvppp/data/data_segmentation.py")
    return ref


def blob_labels(shape, n_seeds, seed, dtype):
    """Small seeded nearest-seed label map with a background region (independent of the oracle)."""
    rng = np.random.RandomState(seed)
    pts = rng.rand(n_seeds, len(shape)) * np.asarray(shape)
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    d = np.stack([sum((g - p[a]) ** 2 for a, g in enumerate(grids)) for p in pts])
    lab = d.argmin(0) + 1
    lab[np.isin(lab, [1, 2])] = 0
    return lab.astype(dtype)


def np32(t):
    return t.detach().cpu().numpy()


def make_2d(ref):
    rng = np.random.RandomState(1234)
    B, C, H, W = 2, 16, 24, 20
    offsets = ref.aff2d.multi_offset([1, 3, 5], neighbor=4) + ref.aff2d.gen_offsets(2, neighbor=8)[2:] + [[3, 0], [0, 25], [-30, 2]]
    K = len(offsets)
    emb = rng.randn(B, C, H, W).astype(np.float32)
    emb[0, :, 3, 4] = 0.0  # an exact-zero pixel: F.normalize eps path
    ema = rng.randn(B, C, H, W).astype(np.float32)
    tgt = np.zeros((B, K, H, W), np.float32)
    msk = np.zeros((B, K, H, W), np.uint8)
    wgt = np.zeros((B, K, H, W), np.float32)
    labels = []
    for b in range(B):
        lab = blob_labels((H, W), 7, 50 + b, np.uint8)
        labels.append(lab)
        tgt[b], msk[b] = ref.aff2d.gen_affs_ours(lab, offsets, ignore=False, padding=True)
        for k in range(K):
            wgt[b, k] = ref.seg.weight_binary_ratio(tgt[b, k])
    crit = ref.crit.WeightedMSE()
    out = dict(offsets=np.asarray(offsets, np.int32), embedding=emb, ema_embedding=ema, target=tgt, mask=msk,
               weightmap=wgt, labels=np.stack(labels))

    e = torch.tensor(emb, requires_grad=True)
    loss, affs, all_loss = ref.loss2d.embedding_loss(e, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), crit,
                                                     offsets, affs0_weight=3)
    loss.backward()
    out.update(loss=np32(loss), affs=np32(affs), all_loss=np.asarray(all_loss, np.float64), grad=np32(e.grad))
    out["embedding2affs"] = np32(ref.loss2d.embedding2affs(torch.tensor(emb), offsets))

    # gradient when the caller also back-propagates through the returned affs
    e = torch.tensor(emb, requires_grad=True)
    up = rng.randn(B, K, H, W).astype(np.float32)
    loss, affs, _ = ref.loss2d.embedding_loss(e, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), crit, offsets)
    (0.5 * loss + (affs * torch.tensor(up)).sum()).backward()
    out.update(affs_upstream=up, grad_with_affs=np32(e.grad))

    # float32 mask (deep-supervision call sites pass a float mask, data_provider.py:236)
    e = torch.tensor(emb, requires_grad=True)
    loss, _, _ = ref.loss2d.embedding_loss(e, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk).float(), crit, offsets)
    out["loss_fmask"] = np32(loss)

    e = torch.tensor(emb, requires_grad=True)
    u = torch.tensor(ema, requires_grad=True)
    loss, affs = ref.loss2d.ema_embedding_loss(e, u, torch.tensor(tgt), torch.tensor(wgt), torch.tensor(msk), crit,
                                               offsets, affs0_weight=3)
    loss.backward()
    out.update(ema_loss=np32(loss), ema_affs=np32(affs), ema_grad=np32(e.grad), ema_grad_partner=np32(u.grad))
    np.savez_compressed(os.path.join(OUT, "loss2d.npz"), **out)


NHOODS_AC = [np.asarray(v).reshape(3, 3) for v in ([-1, 0, 0, 0, -1, 0, 0, 0, -1], [-2, 0, 0, 0, -3, 0, 0, 0, -3],
                                                    [-3, 0, 0, 0, -9, 0, 0, 0, -9], [-4, 0, 0, 0, -27, 0, 0, 0, -27])]


def _inputs_3d(ref, rng, B, D, H, W, label_seed, twelve):
    emb = rng.randn(B, 16, D, H, W).astype(np.float32)
    ema = rng.randn(B, 16, D, H, W).astype(np.float32)
    K = 12 if twelve else 3
    tgt = np.zeros((B, K, D, H, W), np.float32)
    wgt = np.zeros_like(tgt)
    labels = []
    for b in range(B):
        lab = blob_labels((D, H, W), 9, label_seed + b, np.uint16)
        labels.append(lab)
        if twelve:   # data_provider_labeled_deep.py:246-257
            tgt[b] = np.concatenate([ref.aff3d.seg_to_aff(lab, n, pad='') for n in NHOODS_AC], axis=0)
        else:        # deep-supervision scales use the default replicate pad (:234-237)
            tgt[b] = ref.aff3d.seg_to_aff(lab)
        for k in range(K):
            wgt[b, k] = ref.seg.weight_binary_ratio(tgt[b, k])
    return emb, ema, tgt, wgt, np.stack(labels)


def make_3d(ref):
    """Case A: the 12-channel family on a volume just large enough for the 27-shifts (B=1).
    Case B: a tiny B=2 volume for the 3-channel family and the batch normaliser (3-channel replicate-padded target)."""
    crit = ref.crit.WeightedMSE()
    rng = np.random.RandomState(4321)
    out = {}
    for tag, shape, seed in (("A", (1, 5, 28, 29), 70), ("B", (2, 4, 6, 7), 80)):
        emb, ema, tgt, wgt, labels = _inputs_3d(ref, rng, *shape, seed, tag == "A")
        out.update({f"{tag}_embedding": emb, f"{tag}_ema_embedding": ema, f"{tag}_target": tgt,
                    f"{tag}_weightmap": wgt, f"{tag}_labels": labels})
        T, Wm = torch.tensor(tgt), torch.tensor(wgt)

        if tag == "A":
            e = torch.tensor(emb, requires_grad=True)
            loss, affs = ref.loss3d.embedding_loss_norm5(e, T, Wm, crit, affs0_weight=2)
            loss.backward()
            out.update(A_n5_loss=np32(loss), A_n5_affs=np32(affs), A_n5_grad=np32(e.grad))

            e = torch.tensor(emb, requires_grad=True)
            u = torch.tensor(ema, requires_grad=True)
            loss, affs = ref.loss3d.ema_embedding_loss_norm5(e, u, T, Wm, crit, affs0_weight=2)
            loss.backward()
            out.update(A_ema5_loss=np32(loss), A_ema5_affs=np32(affs), A_ema5_grad=np32(e.grad),
                       A_ema5_grad_partner=np32(u.grad))
            out["A_inf5"] = np32(ref.loss3d.inf_embedding_loss_norm5(torch.tensor(emb)))
        else:
            e = torch.tensor(emb, requires_grad=True)
            loss, affs = ref.loss3d.embedding_loss_norm1(e, T, Wm, crit, affs0_weight=2)
            loss.backward()
            out.update(B_n1_loss=np32(loss), B_n1_affs=np32(affs), B_n1_grad=np32(e.grad))

            e = torch.tensor(emb, requires_grad=True)
            loss, affs = ref.loss3d.embedding_loss_norm1(e, T, Wm, crit, affs0_weight=0.5, shift=2)
            loss.backward()
            out.update(B_n1s2_loss=np32(loss), B_n1s2_affs=np32(affs), B_n1s2_grad=np32(e.grad))

            e = torch.tensor(emb, requires_grad=True)
            u = torch.tensor(ema, requires_grad=True)
            loss, affs = ref.loss3d.ema_embedding_loss_norm1(e, u, T, Wm, crit, affs0_weight=2)
            loss.backward()
            out.update(B_ema1_loss=np32(loss), B_ema1_affs=np32(affs), B_ema1_grad=np32(e.grad),
                       B_ema1_grad_partner=np32(u.grad))
            out["B_inf1"] = np32(ref.loss3d.inf_embedding_loss_norm1(torch.tensor(emb)))
    np.savez_compressed(os.path.join(OUT, "loss3d.npz"), **out)


def make_gt(ref):
    out = {}
    # 2D: offsets of both signs, diagonal, and one shift >= the axis length
    lab2 = blob_labels((31, 26), 8, 11, np.uint8)
    offsets = ref.aff2d.multi_offset([1, 3, 9], neighbor=8) + [[2, 0], [0, 4], [-40, 0], [0, 26], [5, -7]]
    out["gt2d_labels"] = lab2
    out["gt2d_offsets"] = np.asarray(offsets, np.int32)
    for ig in (False, True):
        for pd in (False, True):
            a, m = ref.aff2d.gen_affs_ours(lab2, offsets, ignore=ig, padding=pd)
            out[f"gt2d_aff_i{int(ig)}_p{int(pd)}"] = a
            out[f"gt2d_msk_i{int(ig)}_p{int(pd)}"] = m
    a, m = ref.aff2d.gen_affs_ours(lab2.astype(np.float32), offsets, ignore=False, padding=True)
    out["gt2d_aff_float_labels"] = a
    out["gt2d_msk_float_labels"] = np.asarray(m)
    out["multi_offset_cvppp"] = np.asarray(ref.aff2d.multi_offset([1, 3, 5, 9, 27], neighbor=4), np.int32)
    out["multi_offset_n8"] = np.asarray(ref.aff2d.multi_offset([1, 2], neighbor=8), np.int32)
    out["mknhood3d"] = ref.aff3d.mknhood3d(1)

    # 3D seg_to_aff with the four nhoods the AC3/AC4 provider concatenates, and the default replicate pad
    lab3 = blob_labels((7, 33, 30), 10, 12, np.uint16)
    out["gt3d_labels"] = lab3
    aff12 = np.concatenate([ref.aff3d.seg_to_aff(lab3, n, pad='') for n in NHOODS_AC], axis=0)
    out["gt3d_aff12"] = aff12
    out["gt3d_aff3_replicate"] = ref.aff3d.seg_to_aff(lab3)
    out["gt3d_aff3_nopad"] = ref.aff3d.seg_to_aff(lab3, pad='')
    out["gt3d_aff_2dseg_replicate"] = ref.aff3d.seg_to_aff(lab2, np.asarray([[-1, 0], [0, -1]]))
    out["gt3d_w12"] = np.stack([ref.seg.weight_binary_ratio(aff12[k]) for k in range(12)])

    # weight_binary_ratio corner cases: uniform channel, clip at both ends, exactly one half
    chans = []
    c = np.zeros((9, 20), np.float32); chans.append(c.copy())                 # all zero -> ones
    c[:] = 1; chans.append(c.copy())                                            # all one -> ones
    c[:] = 0; c[0, 0] = 1; chans.append(c.copy())                               # wf < 0.05 -> clip low
    c[:] = 1; c[0, 0] = 0; chans.append(c.copy())                               # wf > 0.99 -> clip high
    c[:] = 0; c[:, :10] = 1; chans.append(c.copy())                             # wf == 0.5
    c[:] = 0; c[:, :13] = 1; chans.append(c.copy())                             # wf = 0.65
    c[:] = 0; c[:, :3] = 1; chans.append(c.copy())                              # wf = 0.15
    chans = np.stack(chans)
    out["wbr_in"] = chans
    out["wbr_out"] = np.stack([ref.seg.weight_binary_ratio(ch) for ch in chans])
    out["gt2d_w"] = np.stack([ref.seg.weight_binary_ratio(ch) for ch in out["gt2d_aff_i0_p1"]])
    np.savez_compressed(os.path.join(OUT, "gt.npz"), **out)


def make_flip():
    """convert_consistency_flip of both script trees on every rule combination (square extents: the reference stacks
    the per-sample results, which a transposed non-square sample would break)."""
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), ".."))
    from oracle import ref_loader
    ref = ref_loader.load()
    rng = np.random.RandomState(77)
    out = {}
    rules2 = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], dtype=np.int64)
    g2 = rng.randn(len(rules2), 3, 6, 6).astype(np.float32)
    out["in2d"], out["rules2d"] = g2, rules2
    out["out2d"] = np32(ref.flip2d.convert_consistency_flip(torch.tensor(g2), torch.tensor(rules2)))
    rules3 = np.array([[a, b, c, d] for a in (0, 1) for b in (0, 1) for c in (0, 1) for d in (0, 1)], dtype=np.int64)
    g3 = rng.randn(len(rules3), 2, 3, 5, 5).astype(np.float32)
    out["in3d"], out["rules3d"] = g3, rules3
    out["out3d"] = np32(ref.flip3d.convert_consistency_flip(torch.tensor(g3), torch.tensor(rules3)))
    np.savez_compressed(os.path.join(OUT, "flip.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order for the stored reference outputs
    ref = load_reference()
    make_2d(ref)
    make_3d(ref)
    make_gt(ref)
    for f in ("loss2d.npz", "loss3d.npz", "gt.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
    make_flip()
