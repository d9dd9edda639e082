"""Device-memory plumbing (PyTorch tensors as buffers, CUDA streams).

The library never computes on the CPU: every helper here either hands a device
pointer to ``libacro_b200.so`` or raises.
"""
import numpy as np
import torch

from . import _native


class NoGpuError(RuntimeError):
    pass


def require_cuda():
    if not torch.cuda.is_available():
        raise NoGpuError(
            "acrobotics_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    _native.load()


def current_stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def to_device(x, dtype=torch.float64, ndim=2):
    """numpy / list / tensor -> contiguous CUDA tensor of ``dtype``.
    Returns (tensor, came_from_host)."""
    require_cuda()
    if is_tensor(x):
        host = not x.is_cuda
        t = x.to(device="cuda", dtype=dtype).contiguous()
    else:
        host = True
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        t = torch.from_numpy(a).to(device="cuda", dtype=dtype)
    while t.dim() < ndim:
        t = t.unsqueeze(0)
    return t, host


def to_host(t, host):
    return t.cpu().numpy() if host else t


def unpack_mask(words, n, host):
    """uint32 bit-mask words (int32 tensor) -> bool (n,)."""
    if host:
        w = words.cpu().numpy().view(np.uint32)
        return np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool)
    shifts = torch.arange(32, device=words.device, dtype=torch.int32)
    bits = (words.unsqueeze(1) >> shifts) & 1
    return bits.reshape(-1)[:n].to(torch.bool)


def mask_words(n):
    return (int(n) + 31) // 32
