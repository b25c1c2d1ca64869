"""The formats either side of the hot path on the GPU:

* `MelFrontEnd` — log10-mel extraction of the reference's `MelNet.forward` (preprocess/NAT_mel.py:42-86 with the
  24 kHz hyper-parameters of preprocess/mel_spec_24k.py:296-305: n_fft = win = 1280, hop 320, 80 mel bins,
  0..8000 Hz): clamp, reflect-pad (n_fft - hop)/2, STFT with a periodic Hann window (center=False),
  sqrt(re^2 + im^2 + 1e-9), mel filterbank, log10(clamp(., 1e-5)).
  B200 formulation: the windowed DFT is ONE tcgen05 GEMM whose A operand is the waveform itself viewed as
  overlapping rows (row stride = hop: a TMA tensor map, no im2col copy), against the [cos | -sin] basis with the
  window folded in; the (re, im) column pairs collapse to magnitudes in the epilogue; a second small GEMM applies
  the mel filterbank with log10 in its epilogue. fp16 tensor-core operands would cost ~66 dB of dynamic range, so
  waveform and basis are split hi + lo (three "taps": hi.W_hi + lo.W_hi + hi.W_lo), which is exact to ~1e-7.
* `loudness_pcm16` — `normalize_loudness(wav, -23)` + 24 kHz PCM16 samples as written by the reference
  (scripts/test_final.py:342-347, 436-457).

The mel filterbank is librosa's (`librosa.filters.mel`, Slaney scale + Slaney area normalisation, htk=False),
restated below because librosa is not installed here — that sub-piece is "parity unpinned" (oracle/audio_ref.py
restates it independently in numpy; the STFT itself is pinned against torch.stft, the reference's own call).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .dit import F32, H16


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney')."""
    fft_f = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


class MelFrontEnd:
    def __init__(self, *, sample_rate=24000, n_fft=1280, win_size=1280, hop_size=320, n_mels=80, fmin=0, fmax=8000,
                 device="cuda"):
        assert win_size == n_fft and n_fft % 64 == 0 and n_fft % hop_size == 0 and hop_size % 8 == 0
        self.dev = torch.device(device)
        self.n_fft, self.hop, self.n_mels = n_fft, hop_size, n_mels
        self.pad = (n_fft - hop_size) // 2
        nb = n_fft // 2 + 1
        self.nb = nb
        n = torch.arange(n_fft, dtype=torch.float64)
        win = 0.5 - 0.5 * torch.cos(2 * math.pi * n / n_fft)                      # torch.hann_window (periodic)
        ang = 2 * math.pi * torch.arange(nb, dtype=torch.float64)[:, None] * n[None, :] / n_fft
        basis = torch.stack([torch.cos(ang) * win, -torch.sin(ang) * win], 1).reshape(2 * nb, n_fft)  # (re, im) row pairs
        hi = basis.to(H16)
        lo = (basis - hi.double()).to(H16)
        # taps over the waveform planes [hi | lo | hi]  <->  weight blocks [W_hi | W_hi | W_lo]
        self.w_dft = torch.cat([hi, hi, lo], 1).contiguous().to(self.dev)        # [2*nb, 3*n_fft]
        self.nbp = (nb + 7) // 8 * 8                                             # row stride of the magnitude buffer
        mel = torch.from_numpy(mel_filterbank(sample_rate, n_fft, n_mels, fmin, fmax))
        wm = torch.zeros(n_mels, self.nbp, dtype=H16)
        wm[:, :nb] = mel.to(H16)
        self.w_mel = wm.to(self.dev)

    @torch.no_grad()
    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        """wav fp32 [B, n] (samples in [-1, 1]) -> log10-mel fp32 [B, n_mels, T], T = 1 + (n + 2*pad - n_fft) // hop."""
        dev, hop, n_fft = self.dev, self.hop, self.n_fft
        wav = wav.to(dev, F32).contiguous()
        B, n = wav.shape
        T = 1 + (n + 2 * self.pad - n_fft) // hop
        rows = T + n_fft // hop + 4                       # rows of one plane (row = hop samples)
        plane = rows * hop
        flat = torch.zeros(B * 3 * plane + n_fft, device=dev, dtype=H16)  # (+ slack read by the view's last rows)
        ops.wave_prepare(wav, flat, pad=self.pad, plane=plane)
        frames = flat.as_strided((B, 3 * rows, n_fft), (3 * plane, hop, 1))  # overlapping rows (stride = hop): no copy
        mag = torch.zeros(B * T, self.nbp, device=dev, dtype=H16)
        ops.gemm(frames, self.w_dft, M=T, N=2 * self.nb, G1=B, kpt=n_fft // 64, dil=rows, pad=0, a_c2=(1, 0),
                 swiglu=3, out_fp16=mag, outb_ld=self.nbp, outb_g=(T * self.nbp, 0))
        mel_cl = torch.empty(B * T, self.n_mels, device=dev, dtype=F32)
        ops.gemm(mag[:, : self.nb], self.w_mel[:, : self.nb], act=ops.ACT_LOG10, act_slope=1e-5, out_f32=mel_cl,
                 out_ld=self.n_mels)
        mel = torch.empty(B, self.n_mels, T, device=dev, dtype=F32)
        ops.ntc_to_nct(mel_cl.view(B, T, self.n_mels), mel, ld=self.n_mels)
        return mel


@torch.no_grad()
def loudness_pcm16(wav: torch.Tensor, target_db: float = -23.0) -> torch.Tensor:
    """wav fp32 [B, n] on the GPU -> int16 [B, n]: gain to `target_db` dB RMS, x 32767, round, saturate."""
    wav = wav.to(torch.float32).contiguous()
    out = torch.empty(wav.shape, device=wav.device, dtype=torch.int16)
    ops.loudness_pcm16(wav, out, target_db=target_db)
    return out
