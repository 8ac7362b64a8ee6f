"""Decodes how tcgen05 interprets the A/B shared-memory images of ach_pw_conv_tc (debug aid)."""
import ctypes as C
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from achelous_b200 import _lib
from achelous_b200._lib import AchPwConv

lib = _lib.load()
st = lambda: torch.cuda.current_stream().cuda_stream


def run(X, W):
    """X (K,P), W (O,K) -> out (O,P) through pack + tc kernel."""
    K, P = X.shape
    O = W.shape[0]
    ldw = (O + 3) // 4 * 4
    wt = torch.zeros(K, ldw, device="cuda"); wt[:, :O] = W.t()
    n = lib.ach_pack_pw_tc_elems(K, O)
    hi = torch.zeros(n, device="cuda"); lo = torch.zeros(n, device="cuda")
    _lib.check(lib.ach_pack_pw_tc(wt.data_ptr(), K, O, ldw, hi.data_ptr(), lo.data_ptr(), st()))
    x = X.contiguous().cuda(); out = torch.full((O, P), -1.0, device="cuda")
    s = AchPwConv()
    s.x0, s.x0_bs, s.c0 = x.data_ptr(), K * P, K
    s.out, s.out_bs, s.B, s.O, s.P = out.data_ptr(), O * P, 1, O, P
    _lib.check(lib.ach_pw_conv_tc(C.byref(s), hi.data_ptr(), lo.data_ptr(), st()))
    torch.cuda.synchronize()
    return out.cpu()


K, O, P = 32, 32, 128
X = (torch.arange(K).view(K, 1) * 128 + torch.arange(P).view(1, P)).float()
W = torch.eye(O, K)
out = run(X, W)
print("test1: W=I, expect out[o][p] = X[o][p] = o*128+p")
bad = (out != X[:O]).sum().item()
print("  mismatches:", bad, "of", out.numel())
for o in (0, 1, 2, 3, 4, 7, 8, 9, 16, 31):
    row = out[o]
    dec = [(int(v) // 128, int(v) % 128) if v >= 0 and float(v).is_integer() else ("?", float(v)) for v in row[:12]]
    print(f"  out[{o}][0:12] decoded (k,p):", dec)
for p in (0, 1, 4, 5, 31, 32, 33, 64, 127):
    col = out[:, p]
    dec = [(int(v) // 128, int(v) % 128) if v >= 0 and float(v).is_integer() else ("?", float(v)) for v in col[:10]]
    print(f"  out[0:10][{p}] decoded (k,p):", dec)

X2 = torch.zeros(K, P)
for k in range(K):
    X2[k, k] = 1.0
W2 = (torch.arange(O).view(O, 1) * 64 + torch.arange(K).view(1, K)).float() + 1
out2 = run(X2, W2)
print("test2: X=delta(p==k), expect out[o][p] = W[o][p] for p<32 (W[o][k] = o*64+k+1)")
exp = torch.zeros(O, P); exp[:, :K] = W2
print("  mismatches:", (out2 != exp).sum().item())
for o in (0, 1, 2, 8, 9):
    dec = [((int(v) - 1) // 64, (int(v) - 1) % 64) if v > 0 else "0" for v in out2[o][:34]]
    print(f"  out2[{o}][0:34] decoded (o,k):", dec)
nz = (out2 != 0).nonzero()
print("  nonzero count", len(nz), "expected", O * K, " max p with nonzero:", int(nz[:, 1].max()) if len(nz) else None)
