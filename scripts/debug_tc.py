"""Debug helper for the tcgen05 PFN kernel: prints raw per-pillar extremes (ext) for simple weights."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pp_b200
from pp_b200 import _lib, _runtime
import pp_b200.model as M

def run(W, b, x, gamma=None):
    C = 64
    net = M.PPFeatureNet(9, C).cuda().eval()
    with torch.no_grad():
        net.conv1.weight.copy_(torch.from_numpy(W).reshape(C, 9, 1, 1)); net.conv1.bias.copy_(torch.from_numpy(b))
        if gamma is not None: net.bn1.weight.copy_(torch.from_numpy(gamma))
        out = net(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    ws = [v for k, v in _runtime._workspaces.items() if k[2] == "pfn"][0]
    ext = ws[: x.shape[0] * x.shape[2] * 2 * C * 4].view(torch.float32).reshape(-1, 2, C).cpu().numpy()
    return out.cpu().numpy(), ext

B, P, N = 1, 4, 200
rng = np.random.default_rng(0)
x = rng.normal(0, 1, (B, 9, P, N)).astype(np.float32)
W = np.zeros((64, 9), np.float32); b = np.ones(64, np.float32)
out, ext = run(W, b, x)
print("bias-only: ext max row0", ext[0, 0, :8], "expect 1")
W = np.zeros((64, 9), np.float32); b = np.zeros(64, np.float32)
for c in range(64): W[c, c % 9] = 1.0
out, ext = run(W, b, x)
want = np.stack([x[0, c % 9].max(axis=1) for c in range(64)], 1)   # [P, 64]
print("select-d: ext row0", ext[0, 0, :10]); print("want        ", want[0, :10])
print("row1", ext[1, 0, :10]); print("want", want[1, :10])
L = _lib.load(); L.pp_set_option(b"pfn_tensor_cores", 0)
out2, ext2 = run(W, b, x)
print("simt row0   ", ext2[0, 0, :10])
