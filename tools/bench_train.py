"""Training-step timing (BASELINE configs[4]: vocal2music.yaml training step with Band-MoE, DDP over the GPUs of one box,
synthetic latents, random-init T5 embeddings): forward (soft router, dense experts) + backward + fused AdamW.
  python tools/bench_train.py [--batch 8] [--steps 5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402  (weight recipe only)
from versband_b200 import modules, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)   # configs/vocal2music.yaml:106 — per-GPU batch
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--T", type=int, default=750)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
net = modules.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG)
net.load_state_dict(synth.synth_state_dict(synth.load_manifest("dit"), seed=1234))
net = net.to(dev).train()
# find_unused_parameters: `code_proj.*` exists in the reference's state_dict but is not on the forward path (SURVEY.md
# section 8b); pytorch_lightning 1.9's DDP strategy (main.py:657) passes find_unused_parameters=True by default
model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True) if world > 1 else net
opt = train.FusedAdamW(net.parameters(), lr=1e-5)
B, T = args.batch, args.T
g = torch.Generator(device=dev).manual_seed(100 + rank)
x1 = torch.randn(B, 20, T, device=dev, generator=g)
midi = torch.randint(0, 130, (B, 1, 2 * T), device=dev, generator=g)
beats = torch.randint(0, 3, (B, 1, 2 * T), device=dev, generator=g)
cap = torch.randn(B, 80, 1024, device=dev, generator=g)
ctx = {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}}


def step():
    t = torch.randint(0, 1000, (B,), device=dev, generator=g)
    x0 = torch.randn(B, 20, T, device=dev, generator=g)
    loss, parts = train.cfm_p_losses(lambda xn, tt: model(xn, tt, ctx), x1, t, x0)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.detach())


for _ in range(2):
    l0 = step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    l1 = step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
tm = torch.tensor([ms], device=dev)
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    flop = 3 * 135e9 * B  # forward (dense experts, SURVEY section 8d: 135 GFLOP per clip) + ~2x for backward
    print(f"train step: {float(tm):.1f} ms at per-GPU batch {B} x T={T}, {world} GPU(s); loss {l0:.4f} -> {l1:.4f}; "
          f"~{flop * world / (float(tm) * 1e-3) / 1e12:.0f} TFLOP/s aggregate (TF32 products); peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if world > 1:
    dist.destroy_process_group()
