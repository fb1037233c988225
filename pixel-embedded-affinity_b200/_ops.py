"""Tensor-level wrappers over the C ABI: validation, workspace from torch's caching allocator, current stream.

PyTorch is plumbing here (device memory + streams); every computation is a pea_* call into libpea_b200.so.
"""
import ctypes

import numpy as np
import torch

from ._lib import lib, check, c_float_p, c_int32_p, LossProblem, MULTI_MAX_PROBLEMS, MULTI_MAX_OFFSETS, PeaError

BORDER_WRAP, BORDER_ZERO = 0, 1
MASK_NONE, MASK_U8, MASK_F32 = 0, 1, 2
IMPL_AUTO, IMPL_GENERIC, IMPL_TILED = 0, 1, 2
MAX_OFFSETS = 64

# Kernel-selection hint forwarded to the C ABI (PEA_IMPL_*).  AUTO picks the tiled kernels when the offset set /
# shape qualifies and the table-driven kernel otherwise; tests flip this to compare the two on the same inputs.
DEFAULT_IMPL = IMPL_AUTO


def set_default_impl(impl):
    global DEFAULT_IMPL
    if impl not in (IMPL_AUTO, IMPL_GENERIC, IMPL_TILED):
        raise ValueError(f"unknown impl {impl}")
    DEFAULT_IMPL = impl

_LABEL_CODES = {
    torch.uint8: 0, torch.float32: 4, torch.int8: 5, torch.int16: 6, torch.int32: 7, torch.int64: 8,
}
for _n, _c in (("uint16", 1), ("uint32", 2), ("uint64", 3)):
    if hasattr(torch, _n):
        _LABEL_CODES[getattr(torch, _n)] = _c


# label dtypes whose every value is representable in int32 (the label-driven loss kernels stage int32 planes)
INT32_SAFE_LABELS = {torch.uint8, torch.int8, torch.int16, torch.int32} | ({torch.uint16} if hasattr(torch, "uint16") else set())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _i32(values):
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.int32).reshape(-1))
    return arr, arr.ctypes.data_as(c_int32_p)


def _f32(values):
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
    return arr, arr.ctypes.data_as(c_float_p)


def check_embedding(name, t, ndim=None, make_contiguous=True):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: pea_b200 has no CPU path")
    if t.dtype in (torch.bfloat16, torch.float16):
        # autocast backbones hand over half-precision embeddings: the path computes in fp32 on the exactly representable
        # upcast values (differentiable cast; the gradient comes back rounded to the input dtype: 2^-8 / 2^-11 relative)
        t = t.float()
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (or bfloat16 / float16, upcast here), got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dimensions, got shape {tuple(t.shape)}")
    return t.contiguous() if make_contiguous else t


def check_like(name, t, shape, dtypes=(torch.float32,)):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: pea_b200 has no CPU path")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    if t.dtype not in dtypes:
        t = t.to(dtypes[0])
    return t.contiguous()


def offsets3(offsets, ndim_spatial):
    """list[[dy,dx]] or list[[dz,dy,dx]] -> (K,3) int32."""
    arr = np.asarray(offsets, dtype=np.int64)
    if arr.ndim != 2 or arr.shape[1] != ndim_spatial:
        raise ValueError(f"offsets must be a list of {ndim_spatial}-element offsets, got shape {arr.shape}")
    if not 1 <= arr.shape[0] <= MAX_OFFSETS:
        raise ValueError(f"between 1 and {MAX_OFFSETS} offsets are supported, got {arr.shape[0]}")
    if ndim_spatial == 2:
        arr = np.concatenate([np.zeros((arr.shape[0], 1), np.int64), arr], axis=1)
    return arr.astype(np.int32)


def _dims5(emb):
    if emb.dim() == 4:
        B, C, H, W = emb.shape
        return B, C, 1, H, W
    B, C, D, H, W = emb.shape
    return B, C, D, H, W


def _workspace(B, C, D, H, W, K, cross, device):
    n = lib.pea_workspace_bytes(B, C, D, H, W, K, int(cross))
    if n == 0:
        raise ValueError(f"unsupported problem: B={B} C={C} D={D} H={H} W={W} K={K} (C must be 4, 8, 16 or 32)")
    return torch.empty(n, dtype=torch.uint8, device=device), n


_TICKETS = {}       # (device index, stream handle) -> one-element int32 view
_TICKET_POOL = {}   # device index -> [int32 zeros (256,), slots handed out]


def _ticket(device):
    """One zero-initialised device word per (device, stream): the multi-problem kernel, which finalizes the losses itself,
    counts its CTAs on it and leaves it zero (include/pea_b200.h, pea_affs_loss_multi `ticket`).  Slots come out of a per-device pool created on
    the first call, so nothing is allocated per call and a captured CUDA graph keeps pointing at live memory.  Returns
    None (=> separate finalize launch) if the very first call on a device happens inside a graph capture."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(device).cuda_stream)
    t = _TICKETS.get(key)
    if t is None:
        pool = _TICKET_POOL.get(idx)
        if pool is None:
            if torch.cuda.is_current_stream_capturing():
                return None
            pool = _TICKET_POOL[idx] = [torch.zeros(256, dtype=torch.int32, device=device), 0]
            torch.cuda.current_stream(device).synchronize()  # zero-filled before any other stream counts on a slot
        if pool[1] >= pool[0].numel():
            return None
        t = _TICKETS[key] = pool[0][pool[1]:pool[1] + 1]
        pool[1] += 1
    return t


EPI_RELU = 0x100  # PEA_EPI_RELU: or-ed into the impl hint of pea_affs_fwd


def affs_fwd(emb, partner, offs3, border, impl=None, z_skip=0, out=None, halo=None, relu=False):
    """(B,C,[D,]H,W) -> (B,K,[D - z_skip,]H,W) affinities.  emb/partner already validated.

    ``emb`` may be a z-range view ``buf[:, :, z0:z1]`` of a contiguous (B,C,Dtot,H,W) buffer and ``out`` a z-range
    view of a contiguous (B,K,Dout,H,W) buffer (z-sharded inference): only the channel-plane stride differs from
    the contiguous case and it is passed through to the kernel.  ``halo`` (B,C,n<=4,H,W): the previous shard's last
    slices as a z-range view of any buffer -- including a PEER GPU's symmetric-memory buffer -- read in place."""
    impl = DEFAULT_IMPL if impl is None else impl
    B, C, D, H, W = _dims5(emb)
    K = offs3.shape[0]
    spatial = tuple(emb.shape[2:])
    emb_plane = 0
    if emb.dim() == 5 and not emb.is_contiguous():
        st = emb.stride()
        if st[4] != 1 or st[3] != W or st[2] != H * W or st[0] != C * st[1]:
            emb = emb.contiguous()
        else:
            emb_plane = st[1]
    elif not emb.is_contiguous():
        emb = emb.contiguous()
    if partner is not None:
        partner = partner.contiguous()
    out_spatial = ((D - z_skip,) + spatial[1:]) if emb.dim() == 5 else spatial
    affs_plane = 0
    if out is None:
        affs = torch.empty((B, K) + out_spatial, dtype=torch.float32, device=emb.device)
    else:
        affs = out
        if tuple(affs.shape) != (B, K) + out_spatial or affs.dtype != torch.float32:
            raise ValueError(f"out must be float32 of shape {(B, K) + out_spatial}")
        if not affs.is_contiguous():
            st = affs.stride()
            if affs.dim() != 5 or st[4] != 1 or st[3] != W or st[2] != H * W or st[0] != K * st[1]:
                raise ValueError("out must be contiguous or a z-range view of a contiguous buffer")
            affs_plane = st[1]
    halo_n, halo_plane = 0, 0
    if halo is not None:
        hs = halo.stride()
        if (halo.dim() != 5 or tuple(halo.shape[:2]) != (B, C) or tuple(halo.shape[3:]) != (H, W) or not 1 <= halo.shape[2] <= 4
                or hs[4] != 1 or hs[3] != W or hs[2] != H * W or hs[0] != C * hs[1] or halo.dtype != torch.float32):
            raise ValueError("halo must be a float32 (B,C,n<=4,H,W) z-range view of a contiguous buffer")
        halo_n, halo_plane = int(halo.shape[2]), int(hs[1])
    ws, n = _workspace(B, C, D, H, W, K, partner is not None, emb.device)
    keep, optr = _i32(offs3)
    with torch.cuda.device(emb.device):
        rc = lib.pea_affs_fwd(_ptr(emb), _ptr(partner), B, C, D, H, W, optr, K, border, impl | (EPI_RELU if relu else 0),
                              int(z_skip), emb_plane, affs_plane, _ptr(halo), halo_n, halo_plane, _ptr(affs), _ptr(ws), n,
                              _stream())
    if rc == -8 and emb_plane and impl != IMPL_TILED and halo is None:
        # z-range view that the tiled kernel cannot take (shape / offsets): fall back to a contiguous copy
        return affs_fwd(emb.contiguous(), partner, offs3, border, impl, z_skip, out, relu=relu)
    check("pea_affs_fwd", rc)
    return affs


EPI_GRAD_ACCUMULATE = 0x200  # PEA_EPI_GRAD_ACCUMULATE


def affs_loss_fwd_bwd(emb, partner, offs3, border, target, weight, mask, loss_norm, loss_coef, want_affs=True,
                      want_grad=True, want_grad_partner=False, impl=IMPL_AUTO, grad_into=None):
    """One fused launch sequence.  Returns (loss_out (1+K,), affs or None, grad_emb or None, grad_partner or None).
    ``grad_into``: an existing gradient buffer of emb's shape that this call ADDS to (PEA_EPI_GRAD_ACCUMULATE)."""
    B, C, D, H, W = _dims5(emb)
    K = offs3.shape[0]
    dev = emb.device
    spatial = tuple(emb.shape[2:])
    if mask is None:
        mdt = MASK_NONE
    elif mask.dtype == torch.uint8:
        mdt = MASK_U8
    elif mask.dtype == torch.float32:
        mdt = MASK_F32
    else:
        raise TypeError(f"mask must be uint8 or float32 here, got {mask.dtype}")
    affs = torch.empty((B, K) + spatial, dtype=torch.float32, device=dev) if want_affs else None
    if grad_into is not None:
        if tuple(grad_into.shape) != tuple(emb.shape) or grad_into.dtype != torch.float32 or not grad_into.is_contiguous():
            raise ValueError("grad_into must be a contiguous float32 tensor of the embedding's shape")
        grad, impl = grad_into, impl | EPI_GRAD_ACCUMULATE
    else:
        grad = torch.empty_like(emb) if want_grad else None
    gpar = torch.empty_like(partner) if (want_grad_partner and partner is not None) else None
    loss_out = torch.empty(1 + K, dtype=torch.float32, device=dev)
    ws, n = _workspace(B, C, D, H, W, K, partner is not None, dev)
    keep_o, optr = _i32(offs3)
    keep_n, nptr = _f32(loss_norm)
    keep_c, cptr = _f32(loss_coef)
    with torch.cuda.device(dev):
        rc = lib.pea_affs_loss_fwd_bwd(_ptr(emb), _ptr(partner), B, C, D, H, W, optr, K, border, impl, _ptr(target),
                                       _ptr(weight), _ptr(mask), mdt, nptr, cptr, _ptr(affs), _ptr(loss_out), _ptr(grad),
                                       _ptr(gpar), _ptr(ws), n, _stream())
    check("pea_affs_loss_fwd_bwd", rc)
    return loss_out, affs, grad, gpar


def multi_eligible(problems):
    """True when a list of (emb, partner, offs3, ...) problems fits one pea_affs_loss_multi launch."""
    if not 1 <= len(problems) <= MULTI_MAX_PROBLEMS:
        return False
    C = problems[0]["emb"].shape[1]
    for p in problems:
        if p["emb"].shape[1] != C or p["offs3"].shape[0] > MULTI_MAX_OFFSETS or np.abs(p["offs3"]).max() > 127:
            return False
        if p["emb"].numel() // C > (1 << 31):
            return False
    return True


def affs_loss_multi(problems, want_affs=True):
    """Several small fused losses in ONE launch (pea_affs_loss_multi).  problems: list of dicts with emb, partner,
    offs3, border, target, weight, mask, loss_norm, loss_coef, want_grad.  Returns a list of (loss_out, affs, grad), or
    None when no ticket word is available (first call on a device inside a graph capture)."""
    n = len(problems)
    dev = problems[0]["emb"].device
    ticket = _ticket(dev)
    if ticket is None:
        return None
    arr = (LossProblem * n)()
    keep, outs = [], []
    for i, p in enumerate(problems):
        emb = p["emb"]
        B, C, D, H, W = _dims5(emb)
        K = p["offs3"].shape[0]
        mask = p.get("mask")
        mdt = MASK_NONE if mask is None else (MASK_U8 if mask.dtype == torch.uint8 else MASK_F32)
        affs = torch.empty((B, K) + tuple(emb.shape[2:]), dtype=torch.float32, device=dev) if want_affs else None
        grad = torch.empty_like(emb) if p.get("want_grad", True) else None
        loss_out = torch.empty(1 + K, dtype=torch.float32, device=dev)
        ko, optr = _i32(p["offs3"])
        kn, nptr = _f32(p["loss_norm"])
        kc, cptr = _f32(p["loss_coef"])
        keep += [ko, kn, kc]
        a = arr[i]
        a.emb, a.partner, a.target, a.weight = emb.data_ptr(), (p["partner"].data_ptr() if p.get("partner") is not None else None), \
            p["target"].data_ptr(), p["weight"].data_ptr()
        a.mask = mask.data_ptr() if mask is not None else None
        a.affs = affs.data_ptr() if affs is not None else None
        a.grad_emb = grad.data_ptr() if grad is not None else None
        a.loss_out = loss_out.data_ptr()
        a.offsets, a.loss_norm, a.loss_coef = optr, nptr, cptr
        a.B, a.D, a.H, a.W, a.K, a.border, a.mask_dtype = B, D, H, W, K, p["border"], mdt
        outs.append((loss_out, affs, grad))
    nbytes = lib.pea_affs_loss_multi_workspace(arr, n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pea_affs_loss_multi(arr, n, int(problems[0]["emb"].shape[1]), _ptr(ws), nbytes, _ptr(ticket), _stream())
    check("pea_affs_loss_multi", rc)
    return outs


def scale_multi(bufs, scales):
    """bufs[i] *= scales[i] (0-dim float32 CUDA tensors) in one launch."""
    n = len(bufs)
    B = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bufs])
    L = (ctypes.c_int64 * n)(*[b.numel() for b in bufs])
    S = (ctypes.c_void_p * n)(*[s.data_ptr() for s in scales])
    with torch.cuda.device(bufs[0].device):
        rc = lib.pea_scale_multi(B, L, S, n, _stream())
    check("pea_scale_multi", rc)


def affs_bwd(emb, partner, offs3, border, grad_affs, want_grad=True, want_grad_partner=False):
    B, C, D, H, W = _dims5(emb)
    K = offs3.shape[0]
    dev = emb.device
    grad_affs = check_like("grad_affs", grad_affs, (B, K) + tuple(emb.shape[2:]))
    grad = torch.empty_like(emb) if want_grad else None
    gpar = torch.empty_like(partner) if (want_grad_partner and partner is not None) else None
    ws, n = _workspace(B, C, D, H, W, K, partner is not None, dev)
    keep_o, optr = _i32(offs3)
    with torch.cuda.device(dev):
        rc = lib.pea_affs_bwd(_ptr(emb), _ptr(partner), B, C, D, H, W, optr, K, border, _ptr(grad_affs), _ptr(grad),
                              _ptr(gpar), _ptr(ws), n, _stream())
    check("pea_affs_bwd", rc)
    return grad, gpar


def scale_inplace(buf, scale):
    """buf *= scale (0-dim float32 CUDA tensor); no-op launch when the scalar is exactly 1."""
    with torch.cuda.device(buf.device):
        rc = lib.pea_scale_inplace(_ptr(buf), buf.numel(), _ptr(scale), _stream())
    check("pea_scale_inplace", rc)
    return buf


# ------------------------------------------------------------------------------------------- ground truth
def label_code(t):
    try:
        return _LABEL_CODES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported label dtype {t.dtype}") from None


def gt_affs2d(labels, offsets2, ignore, padding, want_mask=True):
    """labels (B,H,W) CUDA tensor -> (affs (B,K,H,W) f32, mask (B,K,H,W) u8)."""
    B, H, W = labels.shape
    offs = np.asarray(offsets2, dtype=np.int32).reshape(-1, 2)
    K = offs.shape[0]
    if not 1 <= K <= MAX_OFFSETS:
        raise ValueError(f"between 1 and {MAX_OFFSETS} offsets are supported, got {K}")
    affs = torch.empty((B, K, H, W), dtype=torch.float32, device=labels.device)
    mask = torch.empty((B, K, H, W), dtype=torch.uint8, device=labels.device) if want_mask else None
    keep, optr = _i32(offs)
    with torch.cuda.device(labels.device):
        rc = lib.pea_gt_affs2d(_ptr(labels), label_code(labels), B, H, W, optr, K, int(bool(ignore)), int(bool(padding)),
                               _ptr(affs), _ptr(mask), _stream())
    check("pea_gt_affs2d", rc)
    return affs, mask


def gt_affs3d(seg, nhood3, replicate_pad):
    """seg (B,D,H,W) CUDA tensor, nhood (K,3) -> affs (B,K,D,H,W) f32.  replicate_pad in {0, 2, 3}."""
    B, D, H, W = seg.shape
    nh = np.asarray(nhood3, dtype=np.int32).reshape(-1, 3)
    K = nh.shape[0]
    if not 1 <= K <= MAX_OFFSETS:
        raise ValueError(f"between 1 and {MAX_OFFSETS} edges are supported, got {K}")
    affs = torch.empty((B, K, D, H, W), dtype=torch.float32, device=seg.device)
    keep, optr = _i32(nh)
    with torch.cuda.device(seg.device):
        rc = lib.pea_gt_affs3d(_ptr(seg), label_code(seg), B, D, H, W, optr, K, int(replicate_pad), _ptr(affs), _stream())
    check("pea_gt_affs3d", rc)
    return affs


def weight_binary_ratio(label, n_channels, alpha=1.0):
    """label: contiguous float32 CUDA tensor viewed as (n_channels, -1); one independent problem per channel."""
    n_per = label.numel() // n_channels
    weight = torch.empty_like(label)
    nbytes = lib.pea_wbr_stats_bytes(n_channels)
    stats = torch.empty(nbytes // 8, dtype=torch.int64, device=label.device)
    with torch.cuda.device(label.device):
        rc = lib.pea_weight_binary_ratio(_ptr(label), n_channels, n_per, float(alpha), _ptr(weight), _ptr(stats), nbytes,
                                         _stream())
    check("pea_weight_binary_ratio", rc)
    return weight


def gt_targets(labels4, offsets3_, semantics, ignore=False, padding=True, alpha=1.0, want_mask=True):
    """labels (B,D,H,W) CUDA tensor -> (target (B,K,D,H,W) f32, mask u8 or None, weight f32): one cooperative launch (count, then stream the stores) or the two-launch route for volumes too large to be resident."""
    B, D, H, W = labels4.shape
    offs = np.asarray(offsets3_, dtype=np.int32).reshape(-1, 3)
    K = offs.shape[0]
    if not 1 <= K <= MAX_OFFSETS:
        raise ValueError(f"between 1 and {MAX_OFFSETS} offsets are supported, got {K}")
    dev = labels4.device
    target = torch.empty((B, K, D, H, W), dtype=torch.float32, device=dev)
    weight = torch.empty_like(target)
    mask = torch.empty((B, K, D, H, W), dtype=torch.uint8, device=dev) if (want_mask and semantics == 0) else None
    nbytes = lib.pea_gt_counts_bytes(B, K)
    counts = torch.empty(nbytes // 8, dtype=torch.int64, device=dev)
    keep, optr = _i32(offs)
    with torch.cuda.device(dev):
        rc = lib.pea_gt_targets(_ptr(labels4), label_code(labels4), B, D, H, W, optr, K, int(semantics), int(bool(ignore)),
                                int(bool(padding)), float(alpha), _ptr(target), _ptr(mask), _ptr(weight), _ptr(counts),
                                nbytes, _stream())
    check("pea_gt_targets", rc)
    return target, mask, weight


def gt_pair_weights(labels4, offsets3_, semantics, ignore=False, padding=True, alpha=1.0):
    """labels (B,D,H,W) CUDA tensor -> wtab (B,K,2) f32: the value weight_binary_ratio gives the (sample, channel)'s
    weight map where its target is 1 / where it is 0, counted without materialising the target."""
    B, D, H, W = labels4.shape
    offs = np.asarray(offsets3_, dtype=np.int32).reshape(-1, 3)
    K = offs.shape[0]
    dev = labels4.device
    wtab = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    nbytes = lib.pea_gt_counts_bytes(B, K)
    counts = torch.empty(nbytes // 8, dtype=torch.int64, device=dev)
    keep, optr = _i32(offs)
    with torch.cuda.device(dev):
        rc = lib.pea_gt_pair_weights(_ptr(labels4), label_code(labels4), B, D, H, W, optr, K, int(semantics),
                                     int(bool(ignore)), int(bool(padding)), float(alpha), _ptr(wtab), _ptr(counts), nbytes,
                                     _stream())
    check("pea_gt_pair_weights", rc)
    return wtab


def affs_loss_labels_fwd_bwd(emb, labels_i32, offs3, border, ignore, wtab, loss_norm, loss_coef, want_affs=True,
                             want_grad=True, semantics=0):
    """Label-driven fused loss: semantics 0 = gen_affs_ours targets (canonical 2-D tables), 1 = seg_to_aff targets (3-D
    12-offset table); PeaError with code PEA_E_MODE for anything else.
    Returns (loss_out (1+K,), affs or None, grad_emb or None)."""
    B, C, D, H, W = _dims5(emb)
    K = offs3.shape[0]
    dev = emb.device
    affs = torch.empty((B, K) + tuple(emb.shape[2:]), dtype=torch.float32, device=dev) if want_affs else None
    grad = torch.empty_like(emb) if want_grad else None
    loss_out = torch.empty(1 + K, dtype=torch.float32, device=dev)
    ws, n = _workspace(B, C, D, H, W, K, False, dev)
    keep_o, optr = _i32(offs3)
    keep_n, nptr = _f32(loss_norm)
    keep_c, cptr = _f32(loss_coef)
    with torch.cuda.device(dev):
        rc = lib.pea_affs_loss_labels_fwd_bwd(_ptr(emb), _ptr(labels_i32), B, C, D, H, W, optr, K, border, int(semantics),
                                              int(bool(ignore)), _ptr(wtab), nptr, cptr, _ptr(affs), _ptr(loss_out),
                                              _ptr(grad), _ptr(ws), n, _stream())
    check("pea_affs_loss_labels_fwd_bwd", rc)
    return loss_out, affs, grad


def replicate_first(affs, shift=1):
    """In place: affs[:, 0, :s] = affs[:, 0, s:2s]; affs[:, 1, :, :s] = affs[:, 1, :, s:2s]; affs[:, 2, ..., :s] = ...[s:2s]
    (scripts_ac3ac4/main.py:233-236) on a contiguous (B,K>=3,D,H,W) float32 CUDA tensor."""
    if affs.dim() != 5 or not affs.is_cuda or affs.dtype != torch.float32 or not affs.is_contiguous():
        raise ValueError("affs must be a contiguous (B,K,D,H,W) float32 CUDA tensor")
    B, K, D, H, W = affs.shape
    with torch.cuda.device(affs.device):
        rc = lib.pea_replicate_first(_ptr(affs), B, K, D, H, W, 0, int(shift), _stream())
    check("pea_replicate_first", rc)
    return affs
