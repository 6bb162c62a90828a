"""Two fused training steps of one BASELINE.json system (target for an ncu launch list; tools only).
usage: python tools/train_step_only.py [Be] [mh_steps_per_sample]"""
import os
import sys
from dataclasses import replace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200.config import BENCH_SYSTEMS, Train_Config  # noqa: E402
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402
from psiformer_torch_b200.train import Trainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Be"
mhs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mcfg, W, step = BENCH_SYSTEMS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = PsiFormer(mcfg).to(dev)
tcfg = Train_Config(batch_size=W, step_size=step, burn_in_steps=2, monte_carlo_length=1, mh_steps_per_sample=mhs, seed=1234,
                    train_steps=4, wand_mode="disabled", checkpoint_step=10**9)
tr = Trainer(model, tcfg, False)
for i in range(2):
    tr.train_step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
m = tr.train_step(2)
b.record()
torch.cuda.synchronize()
print(name, W, "train step ms", a.elapsed_time(b), "E", float(m["Energy"]))
