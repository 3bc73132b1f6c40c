"""Device-memory plumbing shared by the host mirrors: torch owns memory and streams, the kernels
get raw pointers.  Nothing here computes."""
import torch

from . import _lib

_workspaces = {}
_status = {}


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PPError("%s must be a CUDA tensor: this path has no CPU implementation" % name)


def _index(device):
    if isinstance(device, torch.device):
        idx = device.index
    elif device is None:
        idx = None
    else:
        idx = torch.device(device).index
    return torch.cuda.current_device() if idx is None else idx


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on ``device`` (raw C call: this sits on every kernel launch)."""
    return torch._C._cuda_getCurrentRawStream(_index(device))


class _NoGuard(object):
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """``torch.cuda.device(device)`` only when it is not the current device already (the context manager costs
    several microseconds per launch otherwise)."""
    idx = _index(device)
    return _NO_GUARD if torch.cuda.current_device() == idx else torch.cuda.device(idx)


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch buffer per (device, stream, tag); 256-byte aligned by the caching allocator."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    idx = _index(dev)
    key = (idx, torch._C._cuda_getCurrentRawStream(idx), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def take_workspace(device, tag):
    """Remove and return the scratch buffer of (device, current stream, tag): the caller now owns it (state that
    must survive until a backward pass); the next ``workspace`` call allocates a fresh one."""
    idx = _index(device)
    return _workspaces.pop((idx, torch._C._cuda_getCurrentRawStream(idx), tag), None)


def status_word(device):
    """int32[1] device status word (bits PP_STATUS_*), one per device, zeroed at creation."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    key = _index(dev)
    s = _status.get(key)
    if s is None:
        s = torch.zeros(1, dtype=torch.int32, device=dev)
        _status[key] = s
    return s


_polls = {}


def _raise_status(v, what):
    if v:
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


def check_status(device, what):
    """Synchronising read of the status word; raises on a set bit and clears it."""
    s = status_word(device)
    v = int(s.item())
    if v:
        s.zero_()
    st = _polls.get(_index(device if isinstance(device, torch.device) else torch.device(device)))
    if st is not None:                                   # a pending deferred copy would report the same bits again
        if st["ev"] is not None:
            st["ev"].synchronize()
        st["ev"] = None
        st["pin"].zero_()
    _raise_status(v, what)


def poll_status(device, what):
    """Deferred, non-blocking form of ``check_status`` for the asynchronous entry points (modules, target
    assignment, loss): every call copies the status word to pinned memory on the current stream and raises if the
    copy issued by an EARLIER call has completed with a bit set -- the error surfaces at the next call (or at
    ``check_status``) instead of forcing a device synchronisation into every forward."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    idx = _index(dev)
    s = status_word(dev)
    st = _polls.get(idx)
    if st is None:
        st = _polls[idx] = {"pin": torch.zeros(1, dtype=torch.int32).pin_memory(), "ev": None, "what": what}
    if st["ev"] is not None:
        if not st["ev"].query():
            return                                   # the previous copy is still in flight: look next time
        v, prev = int(st["pin"][0]), st["what"]
        st["ev"] = None
        if v:
            s.zero_()
            st["pin"].zero_()
            _raise_status(v, prev + " (reported by a later call)")
    st["pin"].copy_(s, non_blocking=True)
    st["ev"] = torch.cuda.Event()
    st["ev"].record(torch.cuda.current_stream(dev))
    st["what"] = what
