"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities): python tools/profile_train.py [--batch 8]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from versband_b200 import modules, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--T", type=int, default=750)
args = ap.parse_args()
dev = torch.device("cuda", 0)
net = modules.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG)
net.load_state_dict(synth.synth_state_dict(synth.load_manifest("dit"), seed=1234))
net = net.to(dev).train()
opt = train.FusedAdamW(net.parameters(), lr=1e-5)
B, T = args.batch, args.T
g = torch.Generator(device=dev).manual_seed(100)
x1 = torch.randn(B, 20, T, device=dev, generator=g)
midi = torch.randint(0, 130, (B, 1, 2 * T), device=dev, generator=g)
beats = torch.randint(0, 3, (B, 1, 2 * T), device=dev, generator=g)
cap = torch.randn(B, 80, 1024, device=dev, generator=g)
ctx = {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}}


def step():
    t = torch.randint(0, 1000, (B,), device=dev, generator=g)
    x0 = torch.randn(B, 20, T, device=dev, generator=g)
    loss, _ = train.cfm_p_losses(lambda xn, tt: net(xn, tt, ctx), x1, t, x0)
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
