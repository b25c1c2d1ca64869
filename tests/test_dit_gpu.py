"""DiT engine (CUDA, through the C-ABI) vs the oracle (oracle/accompband_ref.py, CPU fp32) on the
same seeded weights, inputs and Gumbel noise; plus the committed golden fixture of the live
reference (tests/golden/golden_dit.pt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from versband_b200.dit import DiTEngine  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda"


def rel(a, b):
    return float((a.float().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def weights():
    return synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)


@pytest.fixture(scope="module")
def engine(weights):
    return DiTEngine(weights, device=DEV, precision="fp16")


def run_engine(engine, x, t, midi, beats, cap, expo, T):
    cond = engine.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), T)
    expo_d = [tuple(e.to(DEV) for e in blk) for blk in expo]
    out = torch.zeros_like(x, device=DEV)
    v, lb, aux = engine.forward(x.to(DEV), t.to(DEV), cond, expo_d, out_v=out, want_aux=True)
    torch.cuda.synchronize()
    return out, lb, aux, cond


def test_prepare_matches_oracle_conditioning(engine, weights):
    B, T = 2, 64
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=11)
    ref = R.dit_conditioning(weights, {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}})
    cond = engine.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), T)
    torch.cuda.synchronize()
    e_ac = rel(cond.acoustic.view(B, T, -1), ref["acoustic"])
    print(f"acoustic stem (exact tables + split-precision final_proj) vs oracle: {e_ac:.2e}")
    assert e_ac < 2e-5
    assert rel(cond.caption, ref["caption"]) < 2e-2
    assert rel(cond.cap_emb, ref["cap_emb"]) < 2e-2


@pytest.mark.parametrize("B,T", [(2, 64), (1, 200)])
def test_forward_vs_oracle(engine, weights, B, T):
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=21 + T)
    t = torch.tensor([125, 833][:B])
    expo = synth.exp_noise(4, B * T, 4, seed=5)
    ctx = {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}}
    v_ref, lb_ref, aux_ref = R.dit_forward(weights, x, t, ctx, depth=4, n_heads=8, expo=expo, return_aux=True)
    v, lb, aux, _ = run_engine(engine, x, t, midi, beats, cap, expo, T)
    # router decisions: identical wherever the oracle's decision margin exceeds fp16 noise
    agree = total = 0
    for k in range(4):
        for name, key in (("idx_c", "cap_idx"), ("idx_a", "ac_idx")):
            got = aux[k][name].cpu().long()
            agree += int((got == aux_ref[k][key]).sum())
            total += got.numel()
    rate = agree / total
    print(f"router agreement {agree}/{total} = {rate:.4f}; v rel err {rel(v, v_ref):.3e}")
    assert agree >= total - 1  # (exact-index assertion with margin analysis: tests/test_fullsize_gpu.py)
    # tokens whose routing agrees in every block must match to fp16 accuracy
    ok = torch.ones(B * T, dtype=torch.bool)
    for k in range(4):
        ok &= aux[k]["idx_c"].cpu().long() == aux_ref[k]["cap_idx"]
        ok &= aux[k]["idx_a"].cpu().long() == aux_ref[k]["ac_idx"]
    diff = (v.cpu() - v_ref).abs().permute(0, 2, 1).reshape(B * T, -1).max(-1)[0]
    peak = float(v_ref.abs().max())
    assert float(diff.median()) / peak < 1e-3 and rel(v, v_ref) < 2e-3


def test_golden_fixture_of_live_reference(engine, weights):
    gd = torch.load(os.path.join(synth.GOLDEN_DIR, "golden_dit.pt"))
    B, T = gd["B"], gd["T"]
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=gd["input_seed"])
    torch.manual_seed(gd["noise_seed"])  # replay the reference's CPU-generator draws
    expo = [tuple(torch.empty(B * T, k).exponential_() for k in (2, 4, 4)) for _ in range(4)]
    v, lb, aux, _ = run_engine(engine, x, gd["t"], midi, beats, cap, expo, T)
    agree = total = 0
    for k in range(4):
        agree += int((aux[k]["idx_c"].cpu().to(torch.int16) == gd[f"cap_idx_{k}"]).sum())
        agree += int((aux[k]["idx_a"].cpu().to(torch.int16) == gd[f"ac_idx_{k}"]).sum())
        total += 2 * B * T
    print(f"golden: router agreement {agree}/{total}, v rel err {rel(v, gd['v']):.3e}")
    assert agree >= total - 1
    diff = (v.cpu() - gd["v"]).abs().permute(0, 2, 1).reshape(B * T, -1).max(-1)[0]
    assert float(diff.median()) / float(gd["v"].abs().max()) < 1e-3 and rel(v, gd["v"]) < 2e-3


# ---------------------------------------------------------------------------------------------
# accuracy mode: fp32 operands multiplied as TF32 in the linear layers
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def engine_tf32(weights):
    return DiTEngine(weights, device=DEV, precision="tf32")


@pytest.mark.parametrize("B,T", [(2, 64), (1, 200)])
def test_forward_tf32_vs_oracle(engine_tf32, weights, B, T):
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=21 + T)
    t = torch.tensor([125, 833][:B])
    expo = synth.exp_noise(4, B * T, 4, seed=5)
    ctx = {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}}
    v_ref, lb_ref, aux_ref = R.dit_forward(weights, x, t, ctx, depth=4, n_heads=8, expo=expo, return_aux=True)
    v, lb, aux, _ = run_engine(engine_tf32, x, t, midi, beats, cap, expo, T)
    agree = total = 0
    for k in range(4):
        agree += int((aux[k]["idx_c"].cpu().long() == aux_ref[k]["cap_idx"]).sum())
        agree += int((aux[k]["idx_a"].cpu().long() == aux_ref[k]["ac_idx"]).sum())
        total += 2 * B * T
    d = (v.cpu() - v_ref).abs()
    peak = float(v_ref.abs().max())
    l2 = float((v.cpu() - v_ref).norm() / v_ref.norm())
    print(f"tf32: router agreement {agree}/{total}; v max {float(d.max()) / peak:.2e} median {float(d.median()) / peak:.2e} rel-L2 {l2:.2e}")
    assert agree >= total - 2
    assert float(d.median()) / peak < 1e-3 and l2 < 3e-3


@pytest.mark.parametrize("precision,B,T", [("fp16", 2, 750), ("fp16", 3, 50), ("tf32", 2, 64)])
def test_c_entry_point_equals_python_launch_sequence(weights, precision, B, T):
    """vb_dit_forward (one DiT evaluation issued from C++ through one C-ABI call, include/versband.h) vs the same
    launches issued from Python: bit-identical velocity and load-balance loss, with the time conditioning computed
    inside the call and with it precomputed, over two consecutive calls (the workspace carries the MoE bucket
    counters from one call to the next)."""
    eng = DiTEngine(weights, device=DEV, precision=precision)
    x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=3)
    t = torch.tensor([125, 833, 41][:B]).to(DEV)
    cond = eng.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), T)
    expo = [tuple(e.to(DEV) for e in blk) for blk in synth.exp_noise(4, B * T, 4, seed=5)]
    x = x.to(DEV)
    for call in range(2):
        out_c, out_p = torch.zeros_like(x), torch.zeros_like(x)
        v_c, lb_c = eng.forward(x, t, cond, expo, out_v=out_c)
        v_c, lb_c = v_c.clone(), lb_c.clone()
        v_p, lb_p = eng.forward_py(x, t, cond, expo, out_v=out_p)
        torch.cuda.synchronize()
        assert torch.equal(v_c, v_p) and torch.equal(out_c, out_p)
        assert torch.allclose(lb_c, lb_p, rtol=1e-5, atol=0)  # lb sums router probabilities with float atomics
        x = x + 0.1 * out_c
    # precomputed modulation vectors (the sampler's path)
    mod = torch.empty(B, eng.mod_ld, device=DEV)
    hl = torch.empty(B, 2 * eng.depth, device=DEV)
    eng.time_conditioning(t, cond.cap_emb, mod, hl)
    v_c = eng.forward(x, t, cond, expo, pre=(mod, hl))[0].clone()
    v_p = eng.forward_py(x, t, cond, expo, pre=(mod, hl))[0]
    torch.cuda.synchronize()
    assert torch.equal(v_c, v_p)


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_c_prepare_equals_python_launch_sequence(weights, precision):
    """vb_dit_prepare (step-invariant conditioning issued from C++ through one C-ABI call) vs the op-by-op Python
    sequence: every tensor of the conditioning bit-identical — single batch, CFG pair (acoustic stem shared by the
    two caption groups) and the length fix-up of vocal2music_moe.py:397-401 (conditioning one frame short / long)."""
    eng = DiTEngine(weights, device=DEV, precision=precision)
    B, T = 2, 48
    _, midi, beats, cap, cap_u = synth.dit_inputs(B, T, seed=8)
    midi, beats, cap, cap_u = midi.to(DEV), beats.to(DEV), cap.to(DEV), cap_u.to(DEV)
    for T_lat in (T, T + 1, T - 1):
        a = eng.prepare(midi, beats, cap, T_lat)
        b = eng.prepare_py(midi, beats, cap, T_lat)
        torch.cuda.synchronize()
        assert len(a.tensors()) == len(b.tensors()) == 4 + 5 * eng.depth
        for ta, tb in zip(a.tensors(), b.tensors()):
            assert ta.shape == tb.shape and ta.dtype == tb.dtype and torch.equal(ta, tb)
    a = eng.prepare_pair(midi, beats, cap, cap_u, T)
    ac, acl = eng.prepare_acoustic(midi, beats, T)
    b = eng.prepare_py(None, None, torch.cat([cap, cap_u], 0), T, acoustic=(torch.cat([ac, ac], 0), torch.cat([acl, acl], 0)))
    torch.cuda.synchronize()
    for ta, tb in zip(a.tensors(), b.tensors()):
        assert ta.shape == tb.shape and torch.equal(ta, tb)
