"""Training step (BASELINE configs[4]): loss and gradients of the DiT in TRAIN mode (soft router,
dense experts) on the tensor-core autograd path vs autograd through the CPU oracle with identical weights, inputs and
Gumbel noise; the fused AdamW kernel vs torch.optim.AdamW."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import accompband_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from versband_b200 import modules, train  # noqa: E402

DEV = "cuda"


@pytest.fixture(autouse=True)
def _grad_enabled():
    """Other test modules switch autograd off process-wide at import; the training tests need it on."""
    with torch.enable_grad():
        yield


def test_train_step_loss_and_grads_vs_oracle_autograd():
    torch.manual_seed(0)
    sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
    net = modules.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).train()
    B, T = 2, 32
    x1, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=41)
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(B, 20, T, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    expo = synth.exp_noise(4, B * T, 4, seed=9)
    ctx = {"c_crossattn": cap, "c_concat": {"midi": midi, "beats": beats}}
    # ---- oracle: autograd through the functional fp32 restatement (training=True: soft routing, dense experts)
    sd_ref = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    tau = 2.0 * 0.9999  # the reference decays the temperature before using it (vocal2music_moe.py:146-147)
    loss_ref, _ = train.cfm_p_losses(
        lambda xn, tt: R.dit_forward(sd_ref, xn, tt, ctx, depth=4, n_heads=8, temperature=tau, training=True, loss_w=1.0, expo=expo),
        x1, t, x0)
    loss_ref.backward()
    # ---- this repo: tensor-core autograd path on the GPU
    ctx_d = {"c_crossattn": cap.to(DEV), "c_concat": {"midi": midi.to(DEV), "beats": beats.to(DEV)}}
    expo_d = [tuple(e.to(DEV) for e in blk) for blk in expo]
    loss, parts = train.cfm_p_losses(lambda xn, tt: net._forward_train(xn, tt, ctx_d, expo=expo_d), x1.to(DEV), t.to(DEV), x0.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    print(f"loss {float(loss):.6f} vs oracle {float(loss_ref):.6f}; lb {float(parts['lb_loss']):.6f}")
    assert abs(float(loss) - float(loss_ref)) < 2e-3 * abs(float(loss_ref))
    assert abs(float(net.blocks[0].feed_forward.temperature) - tau) < 1e-9 and abs(net.loss_w - 0.9999) < 1e-9
    names = ["final_layer.linear.weight", "blocks.3.feed_forward.freq_experts.2.w2.weight", "blocks.2.attention.wq.weight",
             "blocks.1.feed_forward.caption_experts.0.w1.weight", "blocks.0.adaLN_modulation.1.weight",
             "blocks.0.feed_forward.cross_attention.in_proj_weight", "blocks.1.feed_forward.acoustic_gating_network.weight",
             "midi_proj.0.weight", "proj_in.weight", "t_embedder.mlp.0.weight", "c_embedder.mlp.0.weight",
             "blocks.2.attention.gate", "blocks.0.ffn_norm.weight", "midi_embedding.weight"]
    params = dict(net.named_parameters())
    worst = 0.0
    for n in names:
        gr, gg = sd_ref[n].grad, params[n].grad
        assert gg is not None, n
        e = float((gg.cpu() - gr).norm() / gr.norm().clamp_min(1e-20))
        worst = max(worst, e)
        print(f"  grad {n:60s} rel-L2 {e:.3e}  (|g| {float(gr.norm()):.3e})")
        assert e < 2e-2, (n, e)      # TF32 products (10-bit mantissa) in forward and backward
    print(f"worst gradient rel-L2 {worst:.3e}")


def test_fused_adamw_matches_torch():
    torch.manual_seed(1)
    p = torch.randn(1000, 37, device=DEV)
    gsteps = [torch.randn_like(p) * 0.1 for _ in range(3)]
    q = p.clone().requires_grad_(True)
    ref = torch.optim.AdamW([q], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    mine_p = torch.nn.Parameter(p.clone())
    mine = train.FusedAdamW([mine_p], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    for g in gsteps:
        q.grad = g.clone()
        ref.step()
        mine_p.grad = g.clone()
        mine.step()
    torch.cuda.synchronize()
    assert float((mine_p.data - q.data).abs().max()) < 1e-6
