"""In-graph timeline of one DiT evaluation: a timestamp kernel after every op, captured in a CUDA
graph together with the ops, replayed; prints mean time between consecutive marks (op + launch gap)."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from versband_b200 import _lib, ops  # noqa: E402
from versband_b200.dit import DiTCond, DiTEngine  # noqa: E402

torch.set_grad_enabled(False)
pdl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.lib().vb_set_pdl(pdl)
sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
eng = DiTEngine(sd)
B, T = 1, 750
x, midi, beats, cap_c, cap_u = [t.cuda() for t in synth.dit_inputs(B, T, seed=1)]
ac = eng.prepare_acoustic(midi, beats, T)
cond2 = DiTCond.cat([eng.prepare(None, None, cap_c, T, acoustic=ac), eng.prepare(None, None, cap_u, T, acoustic=ac)])
expo = [tuple(e.cuda() for e in blk) for blk in synth.exp_noise(4, 2 * B * T, 4, seed=2)]
t = torch.tensor([500, 500], device="cuda")
eng.forward(x, t, cond2, expo)
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
names = []
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ops.TIMELINE = (buf, names)
    ops.mark("start")
    eng.forward(x, t, cond2, expo)
    ops.TIMELINE = None
acc = defaultdict(list)
order = []
tot = []
for rep in range(20):
    g.replay()
    torch.cuda.synchronize()
    ts = buf[: len(names)].cpu().tolist()
    tot.append((ts[-1] - ts[0]) / 1e3)
    for i in range(1, len(names)):
        key = names[i]
        if rep == 0 and key not in order:
            order.append(key)
        acc[key].append((ts[i] - ts[i - 1]) / 1e3)
print(f"pdl={pdl}  one DiT evaluation (2B={2 * B}, T={T}) incl. {len(names)} timestamp kernels: {sum(tot) / len(tot):.1f} us")
for k in order:
    v = acc[k]
    n = len(v) // 20
    print(f"{k:20s} x{n:2d}  {sum(v) / len(v):7.2f} us each   {sum(v) / 20:8.1f} us per eval")
