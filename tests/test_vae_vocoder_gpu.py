"""VAE decoder/encoder and HiFi-GAN engines (CUDA, C-ABI) vs the oracle and the committed golden
outputs of the live reference modules."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from versband_b200.vae import VAEEngine  # noqa: E402
from versband_b200.vocoder import HifiGanEngine  # noqa: E402

torch.set_grad_enabled(False)
G = synth.GOLDEN_DIR


def rel(a, b):
    return float((a.float().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def vae_sd():
    return synth.synth_state_dict(synth.load_manifest("vae"), seed=2345)


@pytest.fixture(scope="module")
def vae(vae_sd):
    return VAEEngine(vae_sd, synth.VAE_DDCONFIG, synth.VAE_EMBED_DIM)


@pytest.fixture(scope="module")
def voc_sd():
    return synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN)


@pytest.fixture(scope="module")
def voc(voc_sd):
    return HifiGanEngine(voc_sd, synth.HIFIGAN_CFG)


def test_vae_decode_golden(vae):
    gd = torch.load(os.path.join(G, "golden_vae.pt"))
    mel = vae.decode(gd["z"].cuda())
    torch.cuda.synchronize()
    e = rel(mel, gd["mel"])
    print("vae decode vs reference golden:", e)
    assert mel.shape == gd["mel"].shape and e < 5e-3


def test_vae_decode_vs_oracle_batch(vae, vae_sd):
    g = torch.Generator().manual_seed(3)
    z = torch.randn(2, 20, 100, generator=g)
    ref = R.vae_decode(vae_sd, z, synth.VAE_DDCONFIG)
    mel = vae.decode(z.cuda())
    torch.cuda.synchronize()
    e = rel(mel, ref)
    print("vae decode vs oracle:", e)
    assert e < 5e-3


def test_vae_encode_golden_and_roundtrip(vae, vae_sd):
    gd = torch.load(os.path.join(G, "golden_vae.pt"))
    x = synth.mel_input(1, 96, seed=gd["x_seed"])
    mom = vae.encode_moments(x.cuda())
    torch.cuda.synchronize()
    e = rel(mom, gd["moments"])
    print("vae encode vs reference golden:", e)
    assert e < 5e-3
    # config 1: encode -> mode -> decode roundtrip against the oracle
    z = R.posterior_mode(mom.cpu())
    ref = R.vae_decode(vae_sd, R.posterior_mode(gd["moments"]), synth.VAE_DDCONFIG)
    rec = vae.decode(z.cuda())
    assert rel(rec, ref) < 1e-2


def test_hifigan_golden(voc):
    gd = torch.load(os.path.join(G, "golden_hifigan.pt"))
    wav = voc.forward(synth.mel_input(1, 24, seed=gd["mel_seed"]).cuda())
    torch.cuda.synchronize()
    e = rel(wav, gd["wav"][:, 0])
    print("hifigan vs reference golden:", e)
    assert wav.shape == (1, 24 * 320) and e < 6e-3


def test_hifigan_vs_oracle_batch(voc, voc_sd):
    mel = synth.mel_input(2, 50, seed=9)
    ref = R.hifigan_forward(voc_sd, mel, synth.HIFIGAN_CFG)[:, 0]
    wav = voc.forward(mel.cuda())
    torch.cuda.synchronize()
    e = rel(wav, ref)
    print("hifigan vs oracle:", e)
    assert e < 6e-3


def test_hifigan_c_entry_point_equals_python_launch_sequence(voc):
    """vb_voc_forward (the whole generator issued from C++ through one C-ABI call) vs the same launches from Python."""
    mel = synth.mel_input(2, 64, seed=21).cuda()
    a = voc.forward(mel)
    b = voc.forward_py(mel)
    torch.cuda.synchronize()
    assert a.shape == b.shape == (2, 64 * 320)
    assert torch.equal(a, b)


def test_vae_c_entry_point_equals_python_launch_sequence(vae):
    """vb_vae_decode (post_quant_conv + Decoder1D issued from C++ through one C-ABI call) vs the same launches from
    Python, at a ragged length and at the BASELINE length."""
    for B, T in ((2, 50), (1, 750)):
        z = torch.randn(B, synth.VAE_EMBED_DIM, T, generator=torch.Generator().manual_seed(5)).cuda()
        a = vae.decode(z, scale=0.7)
        b = vae.decode_py(z, scale=0.7)
        torch.cuda.synchronize()
        assert a.shape == b.shape == (B, 80, 2 * T)
        assert torch.equal(a, b)


def test_vae_encode_c_entry_point_equals_python_launch_sequence(vae):
    """vb_vae_encode (Encoder1D + quant_conv issued from C++ through one C-ABI call) vs the same launches from Python."""
    for B, T in ((2, 96), (1, 750)):
        x = synth.mel_input(B, T, seed=31).cuda()
        a = vae.encode_moments(x)
        b = vae.encode_moments_py(x)
        torch.cuda.synchronize()
        assert a.shape == b.shape == (B, 2 * synth.VAE_EMBED_DIM, T // 2)
        assert torch.equal(a, b)
