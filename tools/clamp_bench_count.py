"""How many of the bench's walkers carry PSIF_ST_CLAMP_SUSPECT (tools only).  usage: python tools/clamp_bench_count.py N2 Ne"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = argparse.Namespace(gpus=1, profile_mode=False)
b = bench.Bench.__new__(bench.Bench)
b.args, b.world, b.rank, b.dev = args, 1, 0, torch.device("cuda:0")
for name in sys.argv[1:] or ["N2"]:
    s = b.setup(name)
    out = s["eng"].local_energy(s["x"])
    st = out["status"]
    W = st.numel()
    flagged = ((st & 2) != 0)
    per = 1702 if name == "N2" else W
    chunks = flagged.view(-1)[: (W // per) * per].view(-1, per).sum(1) if W >= per else flagged.sum()[None]
    print(f"{name}: {W} walkers, {int(flagged.sum())} clamp-active ({100.0 * flagged.float().mean():.2f} %), per chunk max {int(chunks.max())}")
