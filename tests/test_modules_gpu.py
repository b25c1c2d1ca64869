"""The drop-in nn.Modules on the GPU, used the way the reference's call sites use them
(SURVEY.md §8b): load a reference-keyed state_dict, call forward / decode / __call__."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from versband_b200 import modules  # noqa: E402

torch.set_grad_enabled(False)


def rel(a, b):
    return float((a.float().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-12))


def test_dit_module_forward_like_diffusion_wrapper():
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
    net = modules.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    B, T = 2, 48
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=31)
    t = torch.tensor([41, 958])
    ctx = {"c_crossattn": cap.cuda(), "c_concat": {"midi": midi.cuda(), "beats": beats.cuda(), "acoustic": None}, "name": None}
    torch.manual_seed(7)
    v, lb = net(x.cuda(), t.cuda(), ctx)
    # same device-generator draws, reference order, fed to the oracle
    torch.manual_seed(7)
    expo = [tuple(torch.empty(B * T, k, device="cuda").exponential_().cpu() for k in (2, 4, 4)) for _ in range(4)]
    v_ref, lb_ref = R.dit_forward(sd, x, t, {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}},
                                  depth=4, n_heads=8, expo=expo)
    assert v.shape == (B, 20, T) and lb.dim() == 0
    diff = (v.cpu() - v_ref).abs().permute(0, 2, 1).reshape(B * T, -1).max(-1)[0]
    assert float(diff.median()) / float(v_ref.abs().max()) < 1e-3   # default precision = tf32
    assert float((v.cpu() - v_ref).norm() / v_ref.norm()) < 2e-3
    assert abs(float(lb) - float(lb_ref)) < 5e-3
    # second call re-uses the memoised step-invariant conditioning and decays loss_w like the reference
    w0 = net.loss_w
    net(x.cuda(), t.cuda(), ctx)
    assert len(net._cond_cache) == 1 and net.loss_w < w0


def test_autoencoder_and_vocoder_modules():
    ae = modules.AutoencoderKL(synth.VAE_EMBED_DIM, dict(synth.VAE_DDCONFIG), {"target": "torch.nn.Identity"})
    vsd = synth.synth_state_dict(synth.load_manifest("vae"), seed=2345)
    ae.load_state_dict(vsd, strict=True)
    ae = ae.cuda().eval()
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 20, 40, generator=g)
    mel = ae.decode(z.cuda())
    assert rel(mel, R.vae_decode(vsd, z, synth.VAE_DDCONFIG)) < 3e-2
    post = ae.encode(synth.mel_input(1, 64, seed=2).cuda())
    assert post.mode().shape == (1, 20, 32) and torch.isfinite(post.sample()).all()

    hsd = synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN)
    voc = modules.HifiGAN(None, device="cuda", config=dict(synth.HIFIGAN_CFG), state_dict=hsd)
    m = synth.mel_input(1, 30, seed=4)[0].t().contiguous()  # [T, 80] like scripts/test_final.py:420
    wav = voc(m)
    ref = R.hifigan_forward(hsd, m.t()[None], synth.HIFIGAN_CFG).view(-1).numpy()
    assert isinstance(wav, np.ndarray) and wav.dtype == np.float32 and wav.shape == (30 * 320,)
    assert np.abs(wav - ref).max() / np.abs(ref).max() < 3e-2
    assert np.allclose(voc.vocode(m.t().numpy()), wav)


def test_cfm_sampler_surface():
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
    net = modules.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    B, T = 1, 32
    _, midi, beats, cap_c, cap_u = synth.dit_inputs(B, T, seed=5)
    ac = {"midi": midi.cuda(), "beats": beats.cuda(), "acoustic": None}
    c = {"caption": cap_c.cuda(), "acoustic": ac, "name": None}
    uc = {"caption": cap_u.cuda(), "acoustic": ac, "name": None}
    s = modules.CFMSampler(net, 1000)
    torch.manual_seed(0)
    z1, _ = s.sample_cfg(c, 3.0, uc, batch_size=B, shape=[20, T])
    torch.manual_seed(0)
    z2, _ = s.sample_cfg(c, 3.0, uc, batch_size=B, shape=[20, T])
    assert z1.shape == (B, 20, T) and torch.isfinite(z1).all() and torch.equal(z1, z2)
