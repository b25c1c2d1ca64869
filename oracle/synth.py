"""ORACLE SUPPORT — deterministic synthetic weights and inputs (SURVEY.md §8d recipe).

No checkpoint ships with the reference and its own init makes the DiT the zero function
(SURVEY.md §0.6), so every parity test and the benchmark use these seeded random weights.
Generation is shape-driven from the key→shape manifests in tests/golden/manifest_*.json
(produced from the live reference by oracle/make_golden.py), so it runs without the reference.
Pure CPU torch RNG → identical tensors in the build container and on the GPU box.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DIT_CFG = dict(in_channels=20, context_dim=768, hidden_size=768, depth=4, num_heads=8, max_len=1500,
               num_experts=4, ori_dim=1024)  # configs/vocal2music.yaml:33-43
VAE_DDCONFIG = dict(double_z=True, in_channels=80, out_ch=80, z_channels=20, kernel_size=5, ch=384,
                    ch_mult=[1, 2, 4], num_res_blocks=2, attn_layers=[3], down_layers=[0], dropout=0.0)
VAE_EMBED_DIM = 20  # configs/vocal2music.yaml:47-66
# HiFi-GAN hyper-parameters are not in the reference repo; synthetic config §8d.
HIFIGAN_CFG = dict(resblock="1", upsample_rates=[8, 5, 4, 2], upsample_kernel_sizes=[16, 15, 8, 4],
                   upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                   resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])


def load_manifest(name: str) -> Dict[str, list]:
    with open(os.path.join(GOLDEN_DIR, f"manifest_{name}.json")) as f:
        return json.load(f)


def _fan_in(shape) -> int:
    if len(shape) == 1:
        return 1
    n = 1
    for s in shape[1:]:
        n *= s
    return n


def survey_state_dict(manifest: Dict[str, list], seed: int) -> Dict[str, torch.Tensor]:
    """The weight recipe of SURVEY.md §8(d), restated from the manifest alone (names + shapes): PyTorch's default
    module inits — Linear / Conv: weight and bias ~ U(+-1/sqrt(fan_in)) (kaiming_uniform, a = sqrt 5); Embedding
    N(0, 1); MultiheadAttention in_proj xavier-uniform with zero biases — then every parameter the reference
    zero-initialises (`final_layer.linear.*`, `adaLN_modulation.1.*`) ~ N(0, 0.02), `attention.gate` ~ N(0, 0.5),
    norm weights 1 + N(0, 0.02) (norm biases stay 0); HiFi-GAN: weight_v ~ N(0, 0.01) (init_weights,
    vocoder/hifigan/modules/hifigan.py:11-14), weight_g = ||v|| (weight_norm's initial value)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    keys = sorted(manifest)
    for key in keys:
        shape = list(manifest[key])
        leaf = key.rsplit(".", 1)[-1]
        stem = key[: -len(leaf)]
        rn = lambda: torch.randn(shape, generator=g, dtype=torch.float32)  # noqa: E731
        ru = lambda b: (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * b  # noqa: E731
        if "final_layer.linear." in key or "adaLN_modulation.1." in key:
            t = 0.02 * rn()
        elif "embedding" in key:
            t = rn()
        elif leaf == "gate":
            t = 0.5 * rn()
        elif leaf == "weight_v":
            t = 0.01 * rn()
        elif leaf == "weight_g":
            t = None
        elif leaf == "in_proj_weight":
            t = ru(math.sqrt(6.0 / (shape[0] + shape[1])))
        elif leaf == "in_proj_bias" or key.endswith("out_proj.bias"):
            t = torch.zeros(shape)
        elif leaf == "bias":
            wshape = manifest.get(stem + "weight") or manifest.get(stem + "weight_v")
            if wshape is None or len(wshape) == 1:  # norm bias
                t = torch.zeros(shape)
            else:
                fan = _fan_in(list(wshape)) if not key.startswith("ups.") else wshape[1] * wshape[2]
                t = ru(1.0 / math.sqrt(fan))
        elif len(shape) == 1:  # norm scales
            t = 1.0 + 0.02 * rn()
        else:
            t = ru(1.0 / math.sqrt(_fan_in(shape)))
        sd[key] = t
    for key in keys:
        if key.endswith("weight_g"):
            v = sd[key[:-1] + "v"]
            sd[key] = v.flatten(1).norm(dim=1).view(manifest[key])
    return sd


def synth_state_dict(manifest: Dict[str, list], seed: int, gain: float = 1.0, recipe: str = "fan_in") -> Dict[str, torch.Tensor]:
    """Seeded fp32 tensors for every key of `manifest` (sorted key order → reproducible).
    recipe "fan_in" (default): every matrix ~ N(0, gain^2 / fan_in), so that activations stay O(1) through all
    layers and every branch (gates, adaLN, final layer) contributes; "survey": SURVEY.md §8(d)'s recipe."""
    if recipe == "survey":
        return survey_state_dict(manifest, seed)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key in sorted(manifest):
        shape = list(manifest[key])
        r = torch.randn(shape, generator=g, dtype=torch.float32) if shape else torch.randn((), generator=g)
        leaf = key.rsplit(".", 1)[-1]
        if "embedding" in key:
            t = r
        elif leaf == "gate":  # Attention.gate (tanh-gated text cross-attention)
            t = 0.5 * r
        elif leaf == "weight_g":
            t = None  # filled below from the matching weight_v
        elif leaf == "weight_v":
            t = r
        elif leaf == "bias" or leaf == "in_proj_bias":
            t = 0.02 * r
        elif len(shape) == 1:  # norm scales
            t = 1.0 + 0.02 * r
        else:
            t = r * (gain / math.sqrt(_fan_in(shape)))
        sd[key] = t
    for key in sorted(manifest):
        if key.endswith("weight_g"):  # weight_norm: w = g * v / ||v||  -> std(w) = g / ||v||
            v = sd[key[:-1] + "v"]
            norm = v.flatten(1).norm(dim=1).view(manifest[key])
            if key.startswith("ups."):  # ConvTranspose1d [Cin, Cout, k]: ~k/4 taps reach one output
                fan = v.shape[0] * max(v.shape[2] // 4, 1)
            else:
                fan = _fan_in(list(v.shape))
            jitter = 1.0 + 0.05 * torch.randn(manifest[key], generator=g)
            sd[key] = norm * jitter * (1.3 * gain / math.sqrt(fan))
    return sd


def dit_inputs(B: int, T: int, seed: int, ori_dim: int = 1024, L: int = 80):
    """x0, midi, beats, caption (cond / uncond) — SURVEY.md §8d."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 20, T, generator=g)
    midi = torch.randint(0, 130, (B, 1, 2 * T), generator=g)
    beats = torch.randint(0, 3, (B, 1, 2 * T), generator=g)
    cap_c = torch.randn(B, L, ori_dim, generator=g)
    cap_u = torch.randn(B, L, ori_dim, generator=g)
    return x, midi, beats, cap_c, cap_u


def exp_noise(depth: int, N: int, E: int, seed: int):
    """Exp(1) draws for the router, per block: ([N,2], [N,E], [N,E])."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(depth):
        out.append(tuple(torch.empty(N, k).exponential_(generator=g) for k in (2, E, E)))
    return out


def mel_input(B: int, T: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 80, T, generator=g) * 6.5 - 5.0  # U(-5, 1.5): log10-mel range

HIFIGAN_GAIN = 0.7  # keeps the tanh output un-saturated (|wav| ~ 0.3) so parity is meaningful
