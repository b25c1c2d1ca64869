"""Device time of ONE graph-replayed DiT evaluation (2 x 750 tokens) under engine switches.
usage: python tools/bench_eval.py [key=value ...]   e.g. use_chain=0 chain_bn=64"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from versband_b200.dit import DiTEngine

torch.set_grad_enabled(False)
sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
prec = "fp16"
kv = dict(a.split("=") for a in sys.argv[1:])
eng = DiTEngine(sd, device="cuda", precision=kv.pop("precision", prec))
B, T = 1, 750
x, midi, beats, cap_c, cap_u = synth.dit_inputs(B, T, seed=3)
cond = eng.prepare_pair(midi.cuda(), beats.cuda(), cap_c.cuda(), cap_u.cuda(), T)
expo = [tuple(e.cuda() for e in blk) for blk in synth.exp_noise(4, 2 * B * T, 4, seed=5)]
t = torch.tensor([500, 500], device="cuda")
xd = x.cuda()


def measure(label, **sw):
    for k, v in sw.items():
        setattr(eng, k, v)
    eng._ws.clear()
    for _ in range(3):
        eng.forward(xd, t, cond, expo)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            eng.forward(xd, t, cond, expo)
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"{label:40s} {a.elapsed_time(b) / 40 * 1e3:8.1f} us per evaluation")


if kv:
    measure(str(kv), **{k: int(v) for k, v in kv.items()})
else:
    measure("default engine")
