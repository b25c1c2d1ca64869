"""Generate tests/golden/* from the LIVE reference (run in the build container only).

    python oracle/make_golden.py

Writes
  manifest_{dit,vae,hifigan}.json   state_dict key -> shape of the reference modules
  golden_dit.pt / golden_vae.pt / golden_hifigan.pt / golden_sched.pt
      outputs of the unmodified reference modules on seeded synthetic weights/inputs
      (oracle/synth.py), at sizes small enough to commit.
The reference is only imported (oracle/ref_shims.py); nothing is copied from it.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims, synth  # noqa: E402

OUT = synth.GOLDEN_DIR
torch.set_grad_enabled(False)


def save_manifest(name, module):
    man = {k: list(v.shape) for k, v in module.state_dict().items()}
    with open(os.path.join(OUT, f"manifest_{name}.json"), "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)
    return man


def golden_dit():
    m = ref_shims.load_dit_module()
    net = m.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG).eval()
    man = save_manifest("dit", net)
    sd = synth.synth_state_dict(man, seed=1234)
    net.load_state_dict(sd, strict=True)
    B, T = 2, 64
    x, midi, beats, cap_c, _ = synth.dit_inputs(B, T, seed=77)
    t = torch.tensor([125, 833])
    # record the router decisions of the unmodified MoE.gumbel_softmax
    rec = []
    orig = m.MoE.gumbel_softmax

    def spy(self, logits, temperature, hard=False):
        out = orig(self, logits, temperature, hard)
        rec.append((logits.clone(), out.clone(), hard))
        return out

    m.MoE.gumbel_softmax = spy
    torch.manual_seed(4321)  # the reference draws its Gumbel noise from the global generator
    ctx = {"c_crossattn": cap_c, "c_concat": {"midi": midi, "beats": beats, "acoustic": None}, "name": None}
    v, lb = net(x, t, ctx)
    m.MoE.gumbel_softmax = orig
    idx = {}
    for k in range(synth.DIT_CFG["depth"]):
        _, cap, _ = rec[3 * k + 1]
        _, ac, _ = rec[3 * k + 2]
        idx[f"cap_idx_{k}"] = cap.argmax(-1).to(torch.int16)
        idx[f"ac_idx_{k}"] = ac.argmax(-1).to(torch.int16)
        idx[f"cap_logits_{k}"] = rec[3 * k + 1][0]
    torch.save({"B": B, "T": T, "weight_seed": 1234, "input_seed": 77, "noise_seed": 4321, "t": t,
                "v": v.clone(), "lb_loss": lb.clone(), **idx}, os.path.join(OUT, "golden_dit.pt"))
    print("dit:", v.shape, float(v.abs().mean()), float(lb))


def golden_vae():
    m = ref_shims.load_vae_module()
    ae = m.AutoencoderKL(embed_dim=synth.VAE_EMBED_DIM, ddconfig=dict(synth.VAE_DDCONFIG),
                         lossconfig={"target": "torch.nn.Identity"}).eval()
    man = save_manifest("vae", ae)
    sd = synth.synth_state_dict(man, seed=2345)
    ae.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(1, 20, 48, generator=g)
    mel = ae.decode(z)
    x = synth.mel_input(1, 96, seed=6)
    moments = ae.quant_conv(ae.encoder(x))
    torch.save({"weight_seed": 2345, "z": z, "mel": mel.clone(), "x_seed": 6, "moments": moments.clone()},
               os.path.join(OUT, "golden_vae.pt"))
    print("vae:", mel.shape, float(mel.abs().mean()), moments.shape, float(moments.abs().mean()))


def golden_hifigan():
    m = ref_shims.load_hifigan_module()
    gen = m.HifiGanGenerator(dict(synth.HIFIGAN_CFG)).eval()
    man = save_manifest("hifigan", gen)
    sd = synth.synth_state_dict(man, seed=3456, gain=synth.HIFIGAN_GAIN)
    gen.load_state_dict(sd, strict=True)
    mel = synth.mel_input(1, 24, seed=7)
    wav = gen(mel)
    torch.save({"weight_seed": 3456, "mel_seed": 7, "wav": wav.clone()}, os.path.join(OUT, "golden_hifigan.pt"))
    print("hifigan:", wav.shape, float(wav.abs().mean()), float(wav.abs().max()))


def golden_sched():
    # integer timesteps of the 24 Euler steps (torchdyn itself is not installable)
    ts = [0, 41, 83, 125, 166, 208, 250, 291, 333, 375, 416, 458, 500, 541, 583, 625, 666, 708, 750, 791,
          833, 875, 916, 958]
    torch.save({"t_long": ts}, os.path.join(OUT, "golden_sched.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_sched()
    golden_hifigan()
    golden_vae()
    golden_dit()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
