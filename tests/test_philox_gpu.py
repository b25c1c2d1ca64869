"""vb_router_noise must reproduce PyTorch's CUDA `exponential_()` stream bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402
from versband_b200 import ops  # noqa: E402
from versband_b200.dit import DiTEngine  # noqa: E402
from versband_b200.sampler import FlowSampler  # noqa: E402


@pytest.mark.parametrize("N", [32, 750, 3000])
def test_native_noise_equals_torch_stream(N):
    E, depth, steps = 4, 4, 3
    gen = torch.Generator(device="cuda")
    gen.manual_seed(987654321)
    torch.empty(5, device="cuda").normal_(generator=gen)  # start from a non-zero offset
    off0 = gen.get_offset()
    ref = [torch.zeros(steps, depth, 2 * N, w, device="cuda") for w in (2, E, E)]
    for s in range(steps):
        for br in range(2):
            for k in range(depth):
                for i in range(3):
                    ref[i][s, k, br * N:(br + 1) * N].exponential_(generator=gen)
    assert gen.get_offset() - off0 == 4 * steps * 2 * depth * 3
    mine = [torch.zeros_like(r) for r in ref]
    ops.router_noise(gen.initial_seed(), off0, N, E, depth, steps, *mine)
    torch.cuda.synchronize()
    for a, b in zip(mine, ref):
        assert torch.equal(a, b)


def test_sampler_native_noise_matches_torch_draws():
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
    eng = DiTEngine(sd)
    B, T = 1, 32
    x0, midi, beats, cap_c, cap_u = [t.cuda() for t in synth.dit_inputs(B, T, seed=3)]
    ac = eng.prepare_acoustic(midi, beats, T)
    c, u = eng.prepare(None, None, cap_c, T, acoustic=ac), eng.prepare(None, None, cap_u, T, acoustic=ac)
    s_native = FlowSampler(eng)
    assert s_native.native_noise_ok()
    s_torch = FlowSampler(eng)
    s_torch._native_noise = False
    g1 = torch.Generator(device="cuda").manual_seed(11)
    g2 = torch.Generator(device="cuda").manual_seed(11)
    a = s_native.sample_cfg(x0, c, u, 3.0, generator=g1)
    b = s_torch.sample_cfg(x0, c, u, 3.0, generator=g2)
    assert torch.equal(a, b) and g1.get_offset() == g2.get_offset()
