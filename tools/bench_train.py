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
ap.add_argument("--graph", action="store_true",
                help="single GPU: capture forward + backward + optimiser into one CUDA graph (the eager step is CPU-bound: "
                     "~75 ms of Python/autograd dispatch for ~48 ms of kernels). The AdamW bias corrections are baked at "
                     "capture time, so this mode is for timing only")
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


t_buf = torch.zeros(B, device=dev, dtype=torch.int64)
x0_buf = torch.zeros(B, 20, T, device=dev)
loss_buf = torch.zeros((), device=dev)


def draw():
    t_buf.copy_(torch.randint(0, 1000, (B,), device=dev, generator=g))
    x0_buf.copy_(torch.randn(B, 20, T, device=dev, generator=g))


def body():
    loss, parts = train.cfm_p_losses(lambda xn, tt: model(xn, tt, ctx), x1, t_buf, x0_buf)
    opt.zero_grad()
    loss.backward()
    opt.step()
    loss_buf.copy_(loss.detach())


graph = None


def step():
    draw()
    if graph is not None:
        graph.replay()
    else:
        body()
    return loss_buf


for _ in range(2):
    l0 = float(step())
torch.cuda.synchronize()
if args.graph and world == 1:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()                      # one more eager step on the capture stream (allocator warm-up)
    torch.cuda.current_stream().wait_stream(side)
    opt.zero_grad()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    step()
b.record()
torch.cuda.synchronize()
l1 = float(loss_buf)
ms = a.elapsed_time(b) / args.steps
tm = torch.tensor([ms], device=dev)
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    flop = 3 * 135e9 * B  # forward (dense experts, SURVEY section 8d: 135 GFLOP per clip) + ~2x for backward
    print(f"train step{' (CUDA graph)' if graph is not None else ''}: {float(tm):.1f} ms at per-GPU batch {B} x T={T}, {world} GPU(s); loss {l0:.4f} -> {l1:.4f}; "
          f"~{flop * world / (float(tm) * 1e-3) / 1e12:.0f} TFLOP/s aggregate (TF32 products); peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if world > 1:
    dist.destroy_process_group()
