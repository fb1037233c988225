"""CPU-only checks: the C-ABI library loads and exports every symbol include/pea_b200.h declares, the host-side
logic (offset tables, normalisers, argument validation) matches the reference's conventions.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pea_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pea_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import pea_b200
    lib = ctypes.CDLL(pea_b200._lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pea_b200.h but not exported"
        assert n in pea_b200._lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(pea_b200._lib.SIGNATURES) == set(names)


def test_abi_version_and_strerror():
    import pea_b200
    lib = pea_b200._lib.lib
    assert lib.pea_abi_version() == 7
    assert lib.pea_strerror(0) == b"ok"
    assert b"channel" in lib.pea_strerror(-3)
    assert lib.pea_workspace_bytes(1, 16, 1, 8, 8, 2, 0) > 0
    assert lib.pea_workspace_bytes(1, 7, 1, 8, 8, 2, 0) == 0      # unsupported channel count
    assert lib.pea_workspace_bytes(1, 16, 1, 8, 8, 65, 0) == 0    # too many offsets
    assert lib.pea_wbr_stats_bytes(12) == 12 * 24


def test_argument_rejection_without_touching_the_device():
    """NULL pointers / bad shapes are refused before any CUDA call."""
    import pea_b200
    lib = pea_b200._lib.lib
    offs = (ctypes.c_int32 * 6)(0, -1, 0, 0, 0, -1)
    assert lib.pea_affs_fwd(None, None, 1, 16, 1, 8, 8, offs, 2, 0, 0, 0, 0, 0, None, 0, 0, None, None, 0, None) == -1
    one = ctypes.c_void_p(16)
    assert lib.pea_affs_fwd(one, None, 0, 16, 1, 8, 8, offs, 2, 0, 0, 0, 0, 0, None, 0, 0, one, one, 0, None) == -2
    assert lib.pea_affs_fwd(one, None, 1, 5, 1, 8, 8, offs, 2, 0, 0, 0, 0, 0, None, 0, 0, one, one, 0, None) == -3
    assert lib.pea_affs_fwd(one, None, 1, 16, 1, 8, 8, offs, 0, 0, 0, 0, 0, 0, None, 0, 0, one, one, 0, None) == -4
    assert lib.pea_affs_fwd(one, None, 1, 16, 1, 8, 8, offs, 2, 7, 0, 0, 0, 0, None, 0, 0, one, one, 0, None) == -8
    assert lib.pea_affs_fwd(one, None, 1, 16, 1, 8, 8, offs, 2, 0, 0, 0, 0, 0, None, 0, 0, one, one, 0, None) == -5
    assert lib.pea_gt_affs2d(one, 99, 1, 8, 8, offs, 2, 0, 1, one, one, None) == -6


def test_new_entry_points_reject_bad_arguments_without_touching_the_device():
    """ABI v4 additions: label-driven loss, weight-pair table, border replicate."""
    import pea_b200
    lib = pea_b200._lib.lib
    one = ctypes.c_void_p(16)
    offs = (ctypes.c_int32 * 6)(0, -1, 0, 0, 0, -1)
    nrm = (ctypes.c_float * 2)(1.0, 1.0)
    # pea_affs_loss_labels_fwd_bwd(emb, labels, B, C, D, H, W, offsets, K, border, semantics, ignore, wtab, norm, coef, ...)
    f = lib.pea_affs_loss_labels_fwd_bwd
    assert f(None, one, 1, 16, 1, 64, 64, offs, 2, 0, 0, 0, one, nrm, nrm, one, one, one, one, 1 << 30, None) == -1
    assert f(one, one, 0, 16, 1, 64, 64, offs, 2, 0, 0, 0, one, nrm, nrm, one, one, one, one, 1 << 30, None) == -2
    assert f(one, one, 1, 16, 1, 64, 64, offs, 2, 0, 5, 0, one, nrm, nrm, one, one, one, one, 1 << 30, None) == -8   # semantics
    assert f(one, one, 1, 16, 1, 64, 64, offs, 2, 0, 0, 0, one, nrm, nrm, one, one, one, one, 0, None) == -5          # workspace
    # a table that is not axis-aligned (a diagonal offset): the caller must take the two-step route
    diag = (ctypes.c_int32 * 6)(0, -1, -1, 0, 0, -1)
    assert f(one, one, 1, 16, 1, 64, 64, diag, 2, 0, 0, 0, one, nrm, nrm, one, one, one, one, 1 << 30, None) == -8
    g = lib.pea_gt_pair_weights
    assert g(None, 0, 1, 1, 8, 8, offs, 2, 0, 0, 1, 1.0, one, one, 64, None) == -1
    assert g(one, 0, 1, 1, 8, 8, offs, 2, 3, 0, 1, 1.0, one, one, 64, None) == -8     # semantics
    assert g(one, 0, 1, 1, 8, 8, offs, 2, 0, 0, 1, 1.0, one, one, 8, None) == -5      # counts scratch too small
    assert g(one, 0, 1, 1, 8, 6, offs, 2, 0, 0, 1, 1.0, one, one, 64, None) == -2     # W % 4 != 0
    r = lib.pea_replicate_first
    assert r(None, 1, 3, 4, 8, 8, 0, 1, None) == -1
    assert r(one, 1, 2, 4, 8, 8, 0, 1, None) == -4      # needs the three shift-1 channels
    assert r(one, 1, 3, 4, 8, 8, 0, 0, None) == -2      # shift must be positive
    assert r(one, 1, 3, 4, 8, 8, 10, 1, None) == -2     # plane smaller than D*H*W


def test_round2_entry_points_reject_bad_arguments_without_touching_the_device():
    """ABI v5 additions: multi-problem loss, gradient scaling list, patch blending, consistency flip, test knobs."""
    import pea_b200
    from pea_b200._lib import LossProblem
    lib = pea_b200._lib.lib
    one = ctypes.c_void_p(16)
    offs = (ctypes.c_int32 * 6)(0, -1, 0, 0, 0, -1)
    nrm = (ctypes.c_float * 2)(1.0, 1.0)
    pr = (LossProblem * 1)()
    pr[0].emb = pr[0].target = pr[0].weight = pr[0].loss_out = 16
    pr[0].offsets, pr[0].loss_norm, pr[0].loss_coef = offs, nrm, nrm
    pr[0].B, pr[0].D, pr[0].H, pr[0].W, pr[0].K, pr[0].border, pr[0].mask_dtype = 1, 1, 8, 8, 2, 0, 0
    f = lib.pea_affs_loss_multi
    assert lib.pea_affs_loss_multi_workspace(pr, 1) > 0 and lib.pea_affs_loss_multi_workspace(pr, 9) == 0
    assert f(pr, 1, 16, one, 1 << 20, None, None) == -1      # the ticket word is required
    assert f(pr, 0, 16, one, 1 << 20, one, None) == -2
    assert f(pr, 1, 5, one, 1 << 20, one, None) == -3
    assert f(pr, 1, 16, one, 8, one, None) == -5              # workspace too small
    pr[0].K = 13
    assert f(pr, 1, 16, one, 1 << 20, one, None) == -4        # more than PEA_MULTI_MAX_OFFSETS
    pr[0].K, pr[0].border = 2, 7
    assert f(pr, 1, 16, one, 1 << 20, one, None) == -8
    b = lib.pea_blend_add
    assert b(None, one, 12, 4, 8, 8, one, one, 8, 16, 16, 0, 0, 0, None) == -1
    assert b(one, one, 12, 4, 8, 8, one, one, 8, 16, 16, 6, 0, 0, None) == -2     # patch leaves the volume
    assert lib.pea_blend_finish(one, one, 12, 8, 16, 16, 4, 0, 0, 0, one, None) == -2   # nothing left after the crop
    c = lib.pea_consistency_flip
    assert c(None, one, 1, 16, 1, 8, 8, one, None) == -1
    assert c(one, one, 1, 16, 1, 8, 8, one, None) == -8       # in place is refused
    assert lib.pea_scale_multi(None, None, None, 1, None) == -1
    assert lib.pea_debug_configure(99, 1) == -8 and lib.pea_debug_configure(0, 0) == 0


def test_offset_tables_match_reference_convention(golden):
    import pea_b200
    g = golden["gt"]
    np.testing.assert_array_equal(np.asarray(pea_b200.multi_offset([1, 3, 5, 9, 27], neighbor=4)), g["multi_offset_cvppp"])
    np.testing.assert_array_equal(np.asarray(pea_b200.multi_offset([1, 2], neighbor=8)), g["multi_offset_n8"])
    np.testing.assert_array_equal(pea_b200.mknhood3d(1), g["mknhood3d"])
    with pytest.raises(AssertionError):
        pea_b200.gen_offsets(1, neighbor=6)
    from pea_b200.loss_embedding_mse_3d import offsets_from_shifts, SHIFTS_NORM5, _norms_3d
    from pea_b200.data_affinity import NHOOD_NORM5
    np.testing.assert_array_equal(offsets_from_shifts(SHIFTS_NORM5), NHOOD_NORM5)
    norms = _norms_3d((4, 16, 18, 160, 160), SHIFTS_NORM5)
    assert norms[0] == 4 * 17 * 160 * 160 and norms[10] == 4 * 18 * 133 * 160 and norms[11] == 4 * 18 * 160 * 133
    with pytest.raises(ValueError):
        _norms_3d((1, 16, 4, 30, 30), SHIFTS_NORM5)


def test_cpu_tensors_are_refused_loudly():
    import pea_b200
    e = torch.randn(1, 16, 8, 8)
    t = torch.zeros(1, 2, 8, 8)
    with pytest.raises(ValueError, match="no CPU path"):
        pea_b200.embedding_loss(e, t, t, t, pea_b200.WeightedMSE(), [[-1, 0], [0, -1]])
    with pytest.raises(ValueError, match="no CPU path"):
        pea_b200.embedding2affs(e, [[-1, 0]])
    with pytest.raises(TypeError):
        pea_b200.embedding2affs(np.zeros((1, 16, 8, 8), np.float32), [[-1, 0]])
    with pytest.raises(ValueError, match="no CPU path"):   # the cosine branch is composed on the device as well
        pea_b200.embedding2affs(e, [[-1, 0]], mode='cos')


def test_weighted_mse_matches_oracle_criterion():
    import pea_b200
    from oracle import pea_oracle as O
    p, t, w = torch.randn(2, 5, 7), torch.rand(2, 5, 7), torch.rand(2, 5, 7)
    assert torch.equal(pea_b200.WeightedMSE()(p, t, w), O.WeightedMSE()(p, t, w))
    p5 = torch.randn(2, 1, 3, 5, 7)
    assert torch.equal(pea_b200.WeightedMSE()(p5, p5 * 0, None), O.WeightedMSE()(p5, p5 * 0, None))


def test_staged_halo_entry_points_reject_bad_arguments_without_touching_the_device():
    """ABI 7: pea_zshard_stage_halo / pea_zshard_affs_fwd (z-shards, halo pulled by a copy engine behind the kernel)."""
    import pea_b200
    lib = pea_b200._lib.lib
    one = ctypes.c_void_p(16)
    st = lib.pea_zshard_stage_halo   # (tail, tail_plane, stage, B, C, halo_slices, H, W, row0, band_rows, flags, epoch, stream)
    assert st(None, 4 * 64, one, 1, 16, 4, 8, 8, 0, 8, one, 1, None) == -1
    assert st(one, 4 * 64, one, 1, 16, 4, 8, 8, 0, 8, None, 1, None) == -1
    assert st(one, 4 * 64, one, 1, 16, 5, 8, 8, 0, 8, one, 1, None) == -2      # at most four halo slices
    assert st(one, 3 * 64, one, 1, 16, 4, 8, 8, 0, 8, one, 1, None) == -2      # plane smaller than the tail
    assert st(one, 4 * 64 + 1, one, 1, 16, 4, 8, 8, 0, 8, one, 1, None) == -2  # plane not a whole number of slices
    assert st(one, 4 * 64, one, 1, 16, 4, 8, 8, 9, 8, one, 1, None) == -2      # row0 beyond the image
    assert st(one, 4 * 64, one, 1, 16, 4, 8, 8, 0, 0, one, 1, None) == -2
    fw = lib.pea_zshard_affs_fwd     # (emb, B, C, D, H, W, emb_plane, affs_plane, halo, n, halo_plane, stage, flags, epoch, row0, band, relu, affs, stream)
    assert fw(one, 1, 16, 6, 64, 64, 0, 0, one, 4, 4 * 4096, None, one, 1, 8, 8, 0, one, None) == -1
    assert fw(one, 1, 8, 6, 64, 64, 0, 0, one, 4, 4 * 4096, one, one, 1, 8, 8, 0, one, None) == -2     # 16 channels only
    assert fw(one, 1, 16, 6, 64, 64, 0, 0, one, 5, 5 * 4096, one, one, 1, 8, 8, 0, one, None) == -2
    assert fw(one, 1, 16, 6, 64, 64, 0, 0, one, 4, 3 * 4096, one, one, 1, 8, 8, 0, one, None) == -2
    assert fw(one, 1, 16, 6, 64, 64, 5 * 4096, 0, one, 4, 4 * 4096, one, one, 1, 8, 8, 0, one, None) == -2   # emb plane < D*H*W
    assert fw(one, 1, 16, 6, 64, 64, 0, 0, one, 4, 4 * 4096, one, one, 1, 4, 8, 0, one, None) == -2    # bands start on tile rows
    assert fw(one, 1, 16, 6, 64, 64, 0, 0, one, 4, 4 * 4096, one, one, 1, 8, 12, 0, one, None) == -2
