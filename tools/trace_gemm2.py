"""Timeline of cluster 0 (both CTAs) of the cta_group::2 GEMM (tools only); GEMM_MODE=1 for the tf32 split."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib as L
lib = L.load()
K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ACT = int(sys.argv[3]) if len(sys.argv) > 3 else 0
C = int(sys.argv[4]) if len(sys.argv) > 4 else 14
rows = 229376 // C * C
X = torch.randn(rows, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.randn(N, device="cuda"); out = torch.empty(rows, N, device="cuda"); scratch = torch.empty(3 * N * K + 4, device="cuda")
PACKED = int(__import__("os").environ.get("GEMM_PACKED", "0"))   # 1: A operand as the packed fp16 pair (contents irrelevant for timing)
MODE = int(os.environ.get("GEMM_MODE", "0"))   # 0 fp16 split, 1 tf32 split
st = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(2 * 18 * 512, dtype=torch.int64, device="cuda")
for it in range(3):
    L.check(lib.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, C, K, N, ACT, MODE, PACKED, out.data_ptr(), scratch.data_ptr(),
                                     tr.data_ptr() if it == 2 else None, st))
torch.cuda.synchronize()
t = tr.cpu().numpy().reshape(2, 18, 512).astype(np.float64)
t0 = t[0, 0, 0]
nkb = K // (32 if MODE == 1 else 64)   # K blocks: 32 columns (tf32 split) or 64 (fp16 split)
print("kb | CTA0: prod empty_ok, issued | split full_ok, emptyA_ok, arrived | mma fullB_ok(if waited) split_ok issued || CTA1: prod empty_ok issued | split full_ok emptyA_ok arrived")
for i in list(range(0, 4)) + list(range(16, 16 + 3 * nkb)):
    a, c = t[0], t[1]
    print(f"{i:3d} | {a[0,i]-t0:8.0f} {a[1,i]-t0:8.0f} | {a[2,i]-t0:8.0f} {a[7,i]-t0:8.0f} {a[3,i]-t0:8.0f} | {a[4,i]-t0:8.0f} {a[5,i]-t0:8.0f} {a[6,i]-t0:8.0f} || "
          f"{c[0,i]-t0:8.0f} {c[1,i]-t0:8.0f} | {c[2,i]-t0:8.0f} {c[7,i]-t0:8.0f} {c[3,i]-t0:8.0f}")
a = t[0]
print("per-k-block period (mma issued):", np.diff(a[6, 16:200]).mean(), " mma issue span:", (a[6, 16:200] - a[5, 16:200]).mean(),
      " gap between batches:", (a[5, 17:200] - a[6, 16:199]).mean())
for ci in (0, 1):
    c = t[ci]
    print(f"CTA{ci}: split busy {(c[3, 16:200] - c[7, 16:200]).mean():.0f}  split wait emptyA after full {(c[7, 16:200] - c[2, 16:200]).mean():.0f}"
          f"  tma latency {(c[2, 16:200] - c[1, 16:200]).mean():.0f}  prod wait {(c[0, 17:200] - c[1, 16:199]).mean():.0f}")
    print(f"CTA{ci} splitter: emptyA->converted {(c[11, 16:200] - c[7, 16:200]).mean():.0f}  tcgen05.st+wait {(c[12, 16:200] - c[11, 16:200]).mean():.0f}  fence+arrive {(c[3, 16:200] - c[12, 16:200]).mean():.0f}")
    print(f"CTA{ci} tiles: tfull->tempty {(c[9, 2:20] - c[8, 2:20]).mean():.0f}  tempty->stored {(c[10, 2:20] - c[9, 2:20]).mean():.0f}  period {np.diff(c[8, 2:20]).mean():.0f}")

c = t[0]
if ACT == 2:
    sl = slice(2, 20)
    d = lambda x, y: (c[x, sl] - c[y, sl]).mean()
    print("epilogue warp 8, first token of each tile: tempty->STS+2 bars %.0f | loads+ss %.0f | shfl+gelu %.0f | stores %.0f | other tokens of the warp %.0f" % (
        d(13, 9), d(14, 13), d(15, 14), d(16, 15), d(10, 16)))
print("tile boundaries (leader): last kb issued | epi tfull seen (+) | tempty arrive (+) | next kb0: mma fullB_ok, split_ok(first MMA) (+ from last issue) | CTA1 tfull/tempty")
for tl in range(2, 8):
    kl = (tl + 1) * nkb - 1
    a, c1 = t[0], t[1]
    print(f"{tl}: {a[6, kl]-t0:8.0f} | +{a[8, tl]-a[6, kl]:5.0f} | +{a[9, tl]-a[6, kl]:5.0f} | +{a[4, kl+1]-a[6, kl]:5.0f} +{a[5, kl+1]-a[6, kl]:5.0f} issued +{a[6, kl+1]-a[6, kl]:5.0f} | cta1 +{c1[8, tl]-a[6, kl]:5.0f} +{c1[9, tl]-a[6, kl]:5.0f}")
