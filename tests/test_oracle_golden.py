"""Pins the oracle (oracle/accompband_ref.py) to outputs of the LIVE reference modules that
oracle/make_golden.py stored in tests/golden/ (CPU only, no reference tree needed)."""
import os

import torch

from oracle import accompband_ref as R
from oracle import synth

G = synth.GOLDEN_DIR
torch.set_grad_enabled(False)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def test_euler_schedule_matches_survey_timesteps():
    ts = torch.load(os.path.join(G, "golden_sched.pt"))["t_long"]
    sched = R.euler_schedule(25)
    assert [t for t, _ in sched] == ts
    assert abs(sum(dt for _, dt in sched) - 1.0) < 1e-6


def test_hifigan_oracle_matches_reference_output():
    gd = torch.load(os.path.join(G, "golden_hifigan.pt"))
    sd = synth.synth_state_dict(synth.load_manifest("hifigan"), seed=gd["weight_seed"], gain=synth.HIFIGAN_GAIN)
    wav = R.hifigan_forward(sd, synth.mel_input(1, 24, seed=gd["mel_seed"]), synth.HIFIGAN_CFG)
    assert wav.shape == gd["wav"].shape
    assert rel(wav, gd["wav"]) < 1e-5


def test_vae_oracle_matches_reference_output():
    gd = torch.load(os.path.join(G, "golden_vae.pt"))
    sd = synth.synth_state_dict(synth.load_manifest("vae"), seed=gd["weight_seed"])
    mel = R.vae_decode(sd, gd["z"], synth.VAE_DDCONFIG)
    assert rel(mel, gd["mel"]) < 1e-5
    mom = R.vae_encode_moments(sd, synth.mel_input(1, 96, seed=gd["x_seed"]), synth.VAE_DDCONFIG)
    assert rel(mom, gd["moments"]) < 1e-5
    assert R.posterior_mode(mom).shape[1] == 20


def test_dit_oracle_matches_reference_output_and_router_indices():
    gd = torch.load(os.path.join(G, "golden_dit.pt"))
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=gd["weight_seed"])
    B, T = gd["B"], gd["T"]
    x, midi, beats, cap_c, _ = synth.dit_inputs(B, T, seed=gd["input_seed"])
    ctx = {"c_crossattn": cap_c, "c_concat": {"midi": midi, "beats": beats}}
    # same global-generator draws, same order, as the reference consumed
    torch.manual_seed(gd["noise_seed"])
    v, lb, aux = R.dit_forward(sd, x, gd["t"], ctx, depth=4, n_heads=8, return_aux=True)
    assert rel(v, gd["v"]) < 2e-5
    assert abs(float(lb) - float(gd["lb_loss"])) < 1e-6
    for k in range(4):
        assert torch.equal(aux[k]["cap_idx"].to(torch.int16), gd[f"cap_idx_{k}"])
        assert torch.equal(aux[k]["ac_idx"].to(torch.int16), gd[f"ac_idx_{k}"])


def test_explicit_noise_equals_generator_draws():
    """Feeding pre-drawn Exp(1) noise (what the CUDA path does) == letting the oracle draw it."""
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1)
    x, midi, beats, cap_c, _ = synth.dit_inputs(1, 32, seed=2)
    ctx = {"c_crossattn": cap_c, "c_concat": {"midi": midi, "beats": beats}}
    t = torch.tensor([500])
    torch.manual_seed(9)
    v0, _ = R.dit_forward(sd, x, t, ctx, depth=4, n_heads=8)
    torch.manual_seed(9)
    expo = [tuple(torch.empty(32, k).exponential_() for k in (2, 4, 4)) for _ in range(4)]
    v1, _ = R.dit_forward(sd, x, t, ctx, depth=4, n_heads=8, expo=expo)
    assert torch.equal(v0, v1)
