"""FLAN-T5 (T5 v1.1) encoder engine vs the reference's own dependency — the `transformers` T5EncoderModel
installed in this image, run on the CPU in fp32 with the same randomly initialised weights (no checkpoint is
available offline). flan-t5-large geometry: d_model 1024, 16 heads x 64, d_ff 2816, gated-gelu, 32 buckets;
80 tokens, no attention mask (reference ldm/modules/encoders/modules.py:194-233)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

torch.set_grad_enabled(False)


def hf_model(num_layers, seed):
    from transformers import T5Config, T5EncoderModel
    cfg = T5Config(vocab_size=32128, d_model=1024, d_kv=64, d_ff=2816, num_layers=num_layers, num_heads=16,
                   relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                   layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu", is_encoder_decoder=False, use_cache=False)
    torch.manual_seed(seed)
    m = T5EncoderModel(cfg).eval()
    # HF's default init (factor 1.0) leaves the residual stream at the scale of the embeddings; draw the norm
    # weights away from 1 so that they are exercised
    for n, p in m.named_parameters():
        if "layer_norm" in n:
            p.data = 1.0 + 0.1 * torch.randn_like(p)
    return m, cfg


@pytest.mark.parametrize("layers", [2, 24])
def test_t5_encoder_vs_transformers(layers):
    from versband_b200.t5 import T5EncoderEngine
    m, cfg = hf_model(layers, seed=layers)
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, cfg.vocab_size, (2, 80), generator=g)
    ids[:, 60:] = 0  # padding tokens take part like any other token (no mask is passed, :215-218)
    ref = m(input_ids=ids).last_hidden_state
    eng = T5EncoderEngine(m.state_dict(), num_layers=layers, num_heads=cfg.num_heads, d_kv=cfg.d_kv,
                          num_buckets=cfg.relative_attention_num_buckets, max_distance=cfg.relative_attention_max_distance,
                          eps=cfg.layer_norm_epsilon)
    out = eng.forward(ids.cuda()).cpu()
    torch.cuda.synchronize()
    l2 = float((out - ref).norm() / ref.norm())
    mx = float((out - ref).abs().max() / ref.abs().max())
    print(f"T5 encoder ({layers} layers, 2 x 80 tokens) vs transformers fp32: rel-L2 {l2:.3e}, max/peak {mx:.3e}")
    assert out.shape == (2, 80, 1024)
    assert l2 < 3e-3 and mx < 1e-2


def test_relative_position_bucket_matches_transformers():
    from transformers.models.t5.modeling_t5 import T5Attention
    from versband_b200.t5 import relative_position_bucket
    ctx = torch.arange(200)
    rel = ctx[None, :] - ctx[:, None]
    ref = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(relative_position_bucket(rel, 32, 128), ref)
