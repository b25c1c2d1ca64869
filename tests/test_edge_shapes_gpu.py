"""Edge shapes of the hot path against the oracle: ragged lengths (T not a multiple of 8 or 128: key padding, tile
tails, bucket tails), the smallest and the largest latent length the model accepts (max_len = 1500 frames = 40 s),
short captions, odd batch sizes — DiT forward with router indices, VAE decode and vocoder."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from versband_b200.dit import DiTEngine  # noqa: E402
from versband_b200.vae import VAEEngine  # noqa: E402
from versband_b200.vocoder import HifiGanEngine  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda"


def l2(a, b):
    return float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dit_sd():
    return synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)


@pytest.fixture(scope="module")
def engine(dit_sd):
    return DiTEngine(dit_sd, device=DEV)


@pytest.mark.parametrize("B,T,L", [(1, 1, 80), (1, 7, 80), (3, 129, 80), (1, 333, 17), (2, 751, 80), (1, 1500, 80)])
def test_dit_forward_ragged_and_extreme_lengths(engine, dit_sd, B, T, L):
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=100 + T, L=L)
    t = torch.tensor([3, 500, 999][:B])
    expo = synth.exp_noise(4, B * T, 4, seed=T)
    ctx = {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}}
    v_ref, lb_ref, aux_ref = R.dit_forward(dit_sd, x, t, ctx, depth=4, n_heads=8, expo=expo, return_aux=True)
    cond = engine.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), T)
    out = torch.zeros_like(x, device=DEV)
    v, lb, aux = engine.forward(x.to(DEV), t.to(DEV), cond, [tuple(e.to(DEV) for e in blk) for blk in expo], out_v=out,
                                want_aux=True)
    out_c = torch.zeros_like(out)
    engine.forward(x.to(DEV), t.to(DEV), cond, [tuple(e.to(DEV) for e in blk) for blk in expo], out_v=out_c)  # C entry point
    torch.cuda.synchronize()
    agree = total = 0
    for k in range(4):
        for name, key in (("idx_c", "cap_idx"), ("idx_a", "ac_idx")):
            agree += int((aux[k][name].cpu().long() == aux_ref[k][key]).sum())
            total += B * T
    e = l2(out, v_ref)
    print(f"B={B} T={T} L={L}: rel-L2 {e:.3e}, router {agree}/{total}, lb {float(lb.mean()):.5f} vs {float(lb_ref):.5f}")
    assert torch.isfinite(out).all() and torch.equal(out, out_c)
    assert agree >= total - 1 and e < (1e-3 if agree == total else 5e-3)


def test_dit_rejects_lengths_beyond_max_len(engine):
    x, midi, beats, cap, _ = synth.dit_inputs(1, 1501, seed=1)
    cond = engine.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), 1501)
    expo = [tuple(e.to(DEV) for e in blk) for blk in synth.exp_noise(4, 1501, 4, seed=1)]
    with pytest.raises(Exception):  # freqs_cis[:T] of the reference ends at max_len (vocal2music_moe.py:421)
        engine.forward(x.to(DEV), torch.tensor([5], device=DEV), cond, expo)


def test_caption_longer_than_the_folded_gate_is_rejected(engine):
    _, midi, beats, cap, _ = synth.dit_inputs(1, 16, seed=1, L=97)
    with pytest.raises(ValueError):
        engine.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), 16)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 13), (1, 1500)])
def test_vae_and_vocoder_ragged_and_extreme_lengths(B, T):
    vsd = synth.synth_state_dict(synth.load_manifest("vae"), seed=2345)
    vae = VAEEngine(vsd, synth.VAE_DDCONFIG, synth.VAE_EMBED_DIM, with_encoder=False)
    z = torch.randn(B, 20, T, generator=torch.Generator().manual_seed(T))
    ref = R.vae_decode(vsd, z, synth.VAE_DDCONFIG)
    mel = vae.decode(z.cuda())
    torch.cuda.synchronize()
    ev = l2(mel, ref)
    hsd = synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN)
    voc = HifiGanEngine(hsd, synth.HIFIGAN_CFG)
    Tm = min(2 * T, 300)  # (the CPU oracle of the vocoder takes ~2 ms per frame)
    melin = synth.mel_input(B, Tm, seed=T)
    wref = R.hifigan_forward(hsd, melin, synth.HIFIGAN_CFG)[:, 0]
    wav = voc.forward(melin.cuda())
    torch.cuda.synchronize()
    eh = l2(wav, wref)
    print(f"B={B} T={T}: VAE decode rel-L2 {ev:.3e} ({tuple(mel.shape)}), HiFi-GAN {Tm} frames rel-L2 {eh:.3e}")
    assert mel.shape == (B, 80, 2 * T) and wav.shape == (B, Tm * 320)
    assert ev < 3e-3 and eh < 5e-3
