"""Walker-sharded training on N GPUs (torchrun): every rank samples its own walkers, the energy mean and the
gradients are all-reduced, and the replicas must stay bit-identical.

    torchrun --nproc-per-node 2 tools/train_ddp.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402
from psiformer_torch_b200.train import Trainer, wrapper  # noqa: E402


def main():
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)                                    # same initial weights on every rank
    mcfg, tcfg = wrapper("small", wand_mode="disabled")
    tcfg.batch_size, tcfg.monte_carlo_length, tcfg.mh_steps_per_sample, tcfg.burn_in_steps = 256, 4, 8, 32
    tcfg.train_steps, tcfg.lr, tcfg.seed = 10, 2e-3, 3
    trainer = Trainer(PsiFormer(mcfg), tcfg, False)
    trainer.train()
    flat = torch.cat([p.detach().reshape(-1) for p in trainer.model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(flat, ref))
    e = [h["Energy"] for h in trainer.history]
    x0 = trainer.mh._state[0, 0].tolist()
    print(f"rank {rank}: walker_id0 {trainer.mh.walker_id0} first walker {x0[0]:+.4f} E {e[0]:.4f} -> {e[-1]:.4f} replicas identical: {same}", flush=True)
    assert same
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
