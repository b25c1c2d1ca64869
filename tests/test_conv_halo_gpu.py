"""Haloed small-C conv path of vb_gemm (activation tile fetched once, per-tap shifted UMMA descriptors)
against F.conv1d, and the full HiFi-GAN engine with the path enabled."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from versband_b200 import _lib, ops  # noqa: E402
from versband_b200.vocoder import HifiGanEngine  # noqa: E402

torch.set_grad_enabled(False)


@pytest.fixture(autouse=True)
def halo_on():
    _lib.lib().vb_set_conv_halo(1)
    yield
    _lib.lib().vb_set_conv_halo(0)


@pytest.mark.parametrize("C,k,dil", [(64, 3, 1), (64, 11, 5), (32, 7, 3), (32, 11, 1), (64, 7, 1)])
def test_halo_conv_matches_conv1d(C, k, dil):
    torch.manual_seed(C + k + dil)
    B, T = 2, 20000
    x = torch.randn(B, T, C, device="cuda").half()
    w = (torch.randn(C, C, k, device="cuda") * 0.1).half()
    wp = torch.zeros(C, k, 64, device="cuda", dtype=torch.float16)
    wp[:, :, :C] = w.permute(0, 2, 1)
    wp = wp.reshape(C, k * 64)
    bias = torch.randn(C, device="cuda")
    res = torch.randn(B, T, C, device="cuda")
    out = torch.zeros(B, T, C, device="cuda")
    nxt = torch.zeros(B, T, C, device="cuda", dtype=torch.float16)
    pad = (k * dil - dil) // 2
    n0 = _lib.launch_count()
    ops.gemm(x, wp, M=T, G1=B, kpt=1, dil=dil, pad=pad, a_c2=(1, 0), bias=bias, residual=res, out_f32=out, out_ld=C,
             out_g=(T * C, 0), out_fp16=nxt, outb_ld=C, outb_g=(T * C, 0), act_out=ops.ACT_LRELU, act_out_slope=0.1)
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=pad, dilation=dil).transpose(1, 2) + res
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    assert float((nxt.float() - F.leaky_relu(ref, 0.1)).abs().max() / ref.abs().max()) < 8e-3


def test_vocoder_with_halo_path():
    sd = synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN)
    voc = HifiGanEngine(sd, synth.HIFIGAN_CFG)
    mel = synth.mel_input(1, 400, seed=9)   # T*40 = 16000 rows at stage 3: m_tiles >= 64 -> halo kernel
    ref = R.hifigan_forward(sd, mel, synth.HIFIGAN_CFG)[:, 0]
    wav = voc.forward(mel.cuda())
    torch.cuda.synchronize()
    assert float((wav.cpu() - ref).abs().max() / ref.abs().max()) < 3e-2
