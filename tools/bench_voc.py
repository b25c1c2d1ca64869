"""Device time of the HiFi-GAN engine (1500 mel frames -> 480000 samples) and of the VAE decoder (750 -> 1500),
graph-replayed, plus parity of the vocoder against the CPU oracle. usage: python tools/bench_voc.py [noparity]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from versband_b200.vocoder import HifiGanEngine
from versband_b200.vae import VAEEngine

torch.set_grad_enabled(False)
sd = synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN)
voc = HifiGanEngine(sd, synth.HIFIGAN_CFG)
mel = synth.mel_input(1, 1500, seed=12)
md = mel.cuda()


def timeit(label, fn, reps=10):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"{label:28s} {a.elapsed_time(b) / reps * 1e3:9.1f} us", flush=True)
    return out


wav = timeit("hifigan 1500 frames", lambda: voc.forward(md))
if "noparity" not in sys.argv:
    from oracle import accompband_ref as R
    ref = R.hifigan_forward(sd, mel, synth.HIFIGAN_CFG)[:, 0]
    w = wav.float().cpu()
    print("hifigan vs oracle: rel-L2 %.3e  max/peak %.3e" % (float((w - ref).norm() / ref.norm()), float((w - ref).abs().max() / ref.abs().max())))
vsd = synth.synth_state_dict(synth.load_manifest("vae"), seed=2345)
vae = VAEEngine(vsd, synth.VAE_DDCONFIG, synth.VAE_EMBED_DIM, with_encoder=False)
z = torch.randn(1, 20, 750, generator=torch.Generator().manual_seed(5)).cuda()
timeit("vae decode T=750", lambda: vae.decode(z))
