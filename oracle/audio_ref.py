"""ORACLE — test infrastructure only, never the product path.

CPU restatement of the formats either side of the AccompBand hot path:
  * log10-mel extraction — MelNet.forward, preprocess/NAT_mel.py:42-86, with the 24 kHz hyper-parameters of
    preprocess/mel_spec_24k.py:296-305. The STFT is the reference's own `torch.stft` call; the mel filterbank is
    `librosa.filters.mel` (a dependency that is NOT installed here): its published algorithm (Slaney mel scale,
    triangular filters, Slaney area normalisation) is restated below independently of the product code —
    **parity unpinned** for that sub-piece (no librosa, no reference fixture to check it against).
  * loudness normalisation + PCM16 — scripts/test_final.py:342-347 (normalize_loudness) and the
    `sf.write(..., subtype='PCM_16')` calls at :436-457 (libsndfile float -> short: x * 0x7FFF, lrint).
"""
from __future__ import annotations

import numpy as np
import torch

HP_24K = dict(audio_sample_rate=24000, audio_num_mel_bins=80, fft_size=1280, win_size=1280, hop_size=320, fmin=0, fmax=8000)


def slaney_mel_points(n_mels, fmin, fmax):
    def hz2mel(f):
        return f / (200.0 / 3) if f < 1000.0 else 15.0 + np.log(f / 1000.0) / (np.log(6.4) / 27.0)

    def mel2hz(m):
        return (200.0 / 3) * m if m < 15.0 else 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    mels = np.linspace(hz2mel(float(fmin)), hz2mel(float(fmax)), n_mels + 2)
    return np.array([mel2hz(m) for m in mels])


def librosa_mel(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel defaults: htk=False, norm='slaney' — restated filter by filter."""
    freqs = np.arange(n_fft // 2 + 1) * sr / n_fft
    pts = slaney_mel_points(n_mels, fmin, fmax)
    fb = np.zeros((n_mels, n_fft // 2 + 1))
    for m in range(n_mels):
        lo, ce, hi = pts[m], pts[m + 1], pts[m + 2]
        up = (freqs - lo) / (ce - lo)
        down = (hi - freqs) / (hi - ce)
        fb[m] = np.clip(np.minimum(up, down), 0, None) * 2.0 / (hi - lo)
    return fb.astype(np.float32)


def mel_spectrogram(y: torch.Tensor, hp=HP_24K) -> torch.Tensor:
    """MelNet.forward (center=False, complex=False), NAT_mel.py:56-79. y fp32 [B, n] -> [B, n_mels, T]."""
    n_fft, hop, win = hp["fft_size"], hp["hop_size"], hp["win_size"]
    y = y.clamp(min=-1.0, max=1.0)
    p = int((n_fft - hop) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), [p, p], mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec)
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    basis = torch.from_numpy(librosa_mel(hp["audio_sample_rate"], n_fft, hp["audio_num_mel_bins"], hp["fmin"], hp["fmax"]))
    spec = torch.matmul(basis, spec)
    return torch.log10(torch.clamp(spec, min=1e-5))


def normalize_loudness(wav: np.ndarray, target_loudness: float) -> np.ndarray:
    """scripts/test_final.py:342-347."""
    rms = np.sqrt(np.mean(wav ** 2))
    loudness = 20 * np.log10(rms)
    gain = target_loudness - loudness
    return wav * 10 ** (gain / 20)


def pcm16(wav: np.ndarray) -> np.ndarray:
    """libsndfile float -> PCM_16 with its default normalisation: lrint(x * 0x7FFF); samples beyond full scale are
    saturated here (libsndfile without SFC_SET_CLIPPING wraps them — never reached after -23 dB normalisation)."""
    return np.clip(np.rint(wav.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)
