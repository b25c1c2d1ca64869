"""Mel front-end and PCM16 back-end vs the CPU oracle (oracle/audio_ref.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import audio_ref as A  # noqa: E402
from versband_b200.audio import MelFrontEnd, loudness_pcm16  # noqa: E402

torch.set_grad_enabled(False)


def synth_wave(B, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 24000.0
    w = torch.zeros(B, n)
    for b in range(B):
        for _ in range(6):  # a few partials + noise: wide dynamic range across mel bands
            f = float(torch.rand(1, generator=g)) * 7000 + 60
            a = 10 ** (-float(torch.rand(1, generator=g)) * 3)
            w[b] += a * torch.sin(2 * np.pi * f * t + float(torch.rand(1, generator=g)) * 6.28)
        w[b] += 1e-3 * torch.randn(n, generator=g)
    return (w / w.abs().max() * 0.9).contiguous()


@pytest.mark.parametrize("B,seconds", [(1, 2.0), (2, 20.0)])
def test_mel_frontend_vs_oracle(B, seconds):
    n = int(seconds * 24000)
    wav = synth_wave(B, n, seed=int(seconds))
    ref = A.mel_spectrogram(wav)
    mel = MelFrontEnd().forward(wav.cuda()).cpu()
    torch.cuda.synchronize()
    assert mel.shape == ref.shape == (B, 80, n // 320)
    d = (mel - ref).abs()
    print(f"log10-mel B={B} {seconds}s: max abs diff {float(d.max()):.3e}, mean {float(d.mean()):.3e} (range {float(ref.min()):.2f}..{float(ref.max()):.2f})")
    assert float(d.max()) < 5e-3 and float(d.mean()) < 5e-4   # log10 units; the reference's mels span [-5, ~1.5]


def test_mel_filterbank_matches_independent_restatement():
    from versband_b200.audio import mel_filterbank
    a = mel_filterbank(24000, 1280, 80, 0, 8000)
    b = A.librosa_mel(24000, 1280, 80, 0, 8000)
    assert a.shape == b.shape == (80, 641) and float(np.abs(a - b).max()) < 1e-6


def test_loudness_pcm16_vs_oracle():
    wav = synth_wave(2, 480000, seed=5) * 0.3
    out = loudness_pcm16(wav.cuda(), -23.0).cpu().numpy()
    for b in range(2):
        ref = A.pcm16(A.normalize_loudness(wav[b].numpy(), -23))
        diff = np.abs(out[b].astype(np.int32) - ref.astype(np.int32))
        print("pcm16 max diff (LSB):", int(diff.max()), " fraction differing:", float((diff > 0).mean()))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3   # gain computed in fp32: a rounding tie may move by one LSB
