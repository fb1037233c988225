"""Host <-> device shuttling shared by the ground-truth drop-ins.

The reference's GT functions take and return NumPy arrays (they run inside DataLoader workers).  The drop-ins
accept either: a NumPy array is uploaded, processed by the CUDA kernel and downloaded (same types out as
the reference); a CUDA tensor stays on the device and tensors come back -- the B200-native use, which removes
the per-sample host generation and its H2D traffic altogether.
"""
import numpy as np
import torch

_NP_TO_TORCH = {
    np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16,
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
}
for _n in ("uint16", "uint32", "uint64"):
    if hasattr(torch, _n):
        _NP_TO_TORCH[np.dtype(_n)] = getattr(torch, _n)


_WORKER_HELP = (
    "pea_b200's ground-truth functions run on the GPU, and this process cannot create a CUDA context: it is a fork-started "
    "DataLoader worker (the reference calls gen_affs_ours / seg_to_aff / weight_binary_ratio inside Dataset.__getitem__, "
    "scripts_cvppp/data/data_provider.py:204-225).  Either start the workers with DataLoader(..., multiprocessing_context='spawn'), "
    "or -- the intended B200 use -- return the label map from __getitem__ and build the targets in the training loop with "
    "pea_b200.make_targets_2d / make_targets_3d / make_targets_pyramid_* (or call the *_labels losses directly).")


def device():
    """The current CUDA device; explains itself when called where CUDA cannot start (DataLoader workers)."""
    if not torch.cuda.is_available():
        raise RuntimeError("pea_b200 needs a CUDA device: there is no CPU implementation in this package")
    try:
        return torch.device("cuda", torch.cuda.current_device())
    except RuntimeError as ex:  # "Cannot re-initialize CUDA in forked subprocess"
        if "fork" in str(ex).lower():
            raise RuntimeError(_WORKER_HELP) from ex
        raise


def labels_to_device(labels):
    """-> (tensor on device with a dtype the kernels take, was_numpy)."""
    if isinstance(labels, torch.Tensor):
        if not labels.is_cuda:
            raise ValueError("label tensors must be CUDA tensors (pass a NumPy array for host data)")
        t = labels
        if t.dtype in (torch.float64, torch.float16, torch.bfloat16):
            t = t.float()
        elif t.dtype == torch.bool:
            t = t.to(torch.uint8)
        return t.contiguous(), False
    arr = np.asarray(labels)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    elif arr.dtype in (np.float64, np.float16):
        arr = arr.astype(np.float32)
    if arr.dtype not in _NP_TO_TORCH:
        raise TypeError(f"unsupported label dtype {arr.dtype}")
    arr = np.ascontiguousarray(arr)
    if arr.dtype.byteorder == '>':
        arr = arr.astype(arr.dtype.newbyteorder('='))
    try:
        t = torch.from_numpy(arr)
    except TypeError:  # torch build without this unsigned dtype: reinterpret as the signed one of equal width
        t = torch.from_numpy(arr.view(np.dtype(f"int{arr.dtype.itemsize * 8}")))
    return t.to(device(), non_blocking=False), True
