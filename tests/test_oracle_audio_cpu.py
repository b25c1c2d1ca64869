"""Known-answer checks of the audio oracle (oracle/audio_ref.py) and of the host-side mel filterbank the product
builds its operands from — no GPU. librosa is not installed here, so the filterbank is pinned through the properties
its documentation states (Slaney mel scale: linear below 1 kHz at 200/3 Hz per mel, log above with step
log(6.4)/27; triangular filters; Slaney normalisation = unit area in Hz) and through the agreement of two independent
restatements (oracle/audio_ref.py:librosa_mel, versband_b200/audio.py:mel_filterbank)."""
import numpy as np
import torch

from oracle import audio_ref as A
from versband_b200.audio import mel_filterbank


def test_slaney_mel_scale_known_points():
    pts = A.slaney_mel_points(80, 0, 8000)
    assert len(pts) == 82 and pts[0] == 0.0 and abs(pts[-1] - 8000.0) < 1e-6
    assert np.all(np.diff(pts) > 0)
    # 1000 Hz is mel 15 (= 1000 / (200/3)); 6400 Hz is 27 log-steps above it
    top = 15.0 + np.log(8.0) / (np.log(6.4) / 27.0)
    mel_of = lambda f: f / (200.0 / 3) if f < 1000 else 15.0 + np.log(f / 1000.0) / (np.log(6.4) / 27.0)  # noqa: E731
    assert abs(mel_of(8000.0) - top) < 1e-12 and abs(mel_of(6400.0) - 42.0) < 1e-9
    # linear region: equally spaced in Hz
    lin = pts[pts < 1000.0]
    assert np.allclose(np.diff(lin), np.diff(lin)[0], rtol=1e-9)


def test_filterbank_properties_and_two_restatements_agree():
    sr, n_fft, n_mels = 24000, 1280, 80
    fb = A.librosa_mel(sr, n_fft, n_mels, 0, 8000)
    assert fb.shape == (n_mels, n_fft // 2 + 1) and (fb >= 0).all()
    df = sr / n_fft
    pts = A.slaney_mel_points(n_mels, 0, 8000)
    # Slaney normalisation: every triangle has unit area in Hz (sampled on the FFT grid: a few percent tolerance
    # for the narrow low-frequency filters)
    area = fb.sum(1) * df
    assert np.all(np.abs(area - 1.0) < 0.08), (area.min(), area.max())
    for m in (0, 10, 40, 79):     # support = (lower edge, upper edge), peak next to the centre frequency
        nz = np.nonzero(fb[m])[0]
        assert nz[0] * df > pts[m] - 1e-6 and nz[-1] * df < pts[m + 2] + 1e-6
        assert abs(np.argmax(fb[m]) * df - pts[m + 1]) <= df
    assert fb[:, int(8000 / df) + 1:].sum() == 0.0    # nothing above fmax
    mine = mel_filterbank(sr, n_fft, n_mels, 0, 8000)
    assert mine.shape == fb.shape and float(np.abs(mine - fb).max()) < 1e-6


def test_log_mel_of_a_sine_peaks_at_its_band():
    sr = 24000
    t = torch.arange(2 * sr) / sr
    for f0 in (440.0, 1000.0, 5000.0):
        mel = A.mel_spectrogram((0.5 * torch.sin(2 * np.pi * f0 * t))[None])
        assert mel.shape == (1, 80, 2 * sr // 320)
        pts = A.slaney_mel_points(80, 0, 8000)
        band = int(np.argmin(np.abs(pts[1:-1] - f0)))
        peak = int(mel[0, :, 10:-10].mean(-1).argmax())
        assert abs(peak - band) <= 1, (f0, peak, band)
        assert float(mel.min()) >= -5.0 - 1e-6   # log10(clamp(., 1e-5))


def test_loudness_and_pcm16_known_answers():
    x = np.array([0.0, 1.0, -1.0, 0.5, 1.4 / 32767.0, 1.6 / 32767.0, -0.4 / 32767.0, 2.0, -2.0], dtype=np.float32)
    q = A.pcm16(x)
    # x * 0x7FFF rounded to nearest (lrint), saturated at the int16 range
    assert q.tolist() == [0, 32767, -32767, 16384, 1, 2, 0, 32767, -32768] and q.dtype == np.int16
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(24000) * 0.03).astype(np.float32)
    y = A.normalize_loudness(w, -23.0)
    assert abs(20 * np.log10(np.sqrt(np.mean(y.astype(np.float64) ** 2))) + 23.0) < 1e-4
    assert np.allclose(A.normalize_loudness(3.0 * w, -23.0), y, rtol=1e-5, atol=1e-7)   # gain-invariant
