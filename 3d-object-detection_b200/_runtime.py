"""Device-memory plumbing shared by the host mirrors: torch owns memory and streams, the kernels
get raw pointers.  Nothing here computes."""
import torch

from . import _lib

_workspaces = {}
_status = {}


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PPError("%s must be a CUDA tensor: this path has no CPU implementation" % name)


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch buffer per (device, stream, tag); 256-byte aligned by the caching allocator."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def take_workspace(device, tag):
    """Remove and return the scratch buffer of (device, current stream, tag): the caller now owns it (state that
    must survive until a backward pass); the next ``workspace`` call allocates a fresh one."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream, tag)
    return _workspaces.pop(key, None)


def status_word(device):
    """int32[1] device status word (bits PP_STATUS_*), one per device, zeroed at creation."""
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _status.get(key)
    if s is None:
        s = torch.zeros(1, dtype=torch.int32, device=dev)
        _status[key] = s
    return s


def check_status(device, what):
    """Synchronising read of the status word; raises on a set bit and clears it."""
    s = status_word(device)
    v = int(s.item())
    if v:
        s.zero_()
        msgs = []
        if v & _lib.STATUS_NEG_IOU:
            msgs.append("IOU < 0 (wrong corner winding; the reference exits the process here, "
                        "data/pillars.cpp:166-169)")
        if v & _lib.STATUS_BAD_POINT:
            msgs.append("NaN/Inf point coordinate passed the range filter (out of contract)")
        if v & _lib.STATUS_BAD_INDEX:
            msgs.append("scatter index outside the canvas")
        if v & _lib.STATUS_CAND_OVERFLOW:
            msgs.append("candidate overflow")
        if v & _lib.STATUS_RANGE:
            msgs.append("fused input path: data_mean / conv weight outside the fp16 range of the padding "
                        "pass; use pillarize + encode")
        raise _lib.PPError("%s: %s" % (what, "; ".join(msgs)))
