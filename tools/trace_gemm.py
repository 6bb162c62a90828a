"""Timeline of CTA 0 of the tcgen05 GEMM (tools only)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib as L
lib = L.load()
K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 768
ACT = int(sys.argv[3]) if len(sys.argv) > 3 else 0
C = int(sys.argv[4]) if len(sys.argv) > 4 else 14
rows = 229376 // C * C
X = torch.randn(rows, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.randn(N, device="cuda"); out = torch.empty(rows, N, device="cuda"); scratch = torch.empty(2 * N * K, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(11 * 512, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2: L.check(lib.psif_debug_set_trace(tr.data_ptr()))
    L.check(lib.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, C, K, N, ACT, out.data_ptr(), scratch.data_ptr(), st))
torch.cuda.synchronize(); L.check(lib.psif_debug_set_trace(None))
t = tr.cpu().numpy().reshape(11, 512).astype(np.float64)
t0 = t[0, 0]
names = ["prod:empty_ok", "prod:issued", "split:full_ok", "split:arrived", "mma:full_ok", "mma:split_ok", "mma:issued", "split:emptyA_ok", "epi:tfull", "epi:tempty", "epi:stored"]
nkb = K // 32
print("k-block timeline (cycles since start), k-blocks 16..40")
for i in list(range(0, 4)) + list(range(16, 16 + 3 * nkb)):
    print(i, " ".join(f"{names[r].split(':')[0][:3]}:{names[r].split(':')[1][:8]}={t[r, i] - t0:9.0f}" for r in (0, 1, 2, 7, 3, 5, 6)))
print("per-k-block period (mma issued):", np.diff(t[6, 16:200]).mean(), " split busy:", (t[3, 16:200] - t[7, 16:200]).mean(),
      " mma issue span:", (t[6, 16:200] - t[5, 16:200]).mean(), " tma latency (issue->full seen by splitter):", (t[2, 16:200] - t[1, 16:200]).mean())
print("mma gap between issue batches:", (t[5, 17:200] - t[6, 16:199]).mean())
print("prod wait EMPTY (from prev issue):", (t[0, 17:200] - t[1, 16:199]).mean())
print("tiles: epi tfull->tempty:", (t[9, 2:20] - t[8, 2:20]).mean(), " tempty->stored:", (t[10, 2:20] - t[9, 2:20]).mean(), " tile period:", np.diff(t[8, 2:20]).mean())
print("tile boundaries: [last K block issued -> epilogue sees TFULL -> epilogue released TEMPTY -> first K block of next tile starts]")
for tl in range(2, 8):
    kb_last = (tl + 1) * nkb - 1
    print(tl, "%.0f -> +%.0f -> +%.0f -> +%.0f" % (t[6, kb_last] - t0, t[8, tl] - t[6, kb_last], t[9, tl] - t[8, tl], t[5, kb_last + 1] - t[9, tl]))
