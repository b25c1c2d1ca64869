"""CPU-side checks of the drop-in boundary: C-ABI symbols, struct ABI, state_dict key fidelity of
the nn.Module surfaces, YAML-target resolution, loud failure without a GPU."""
import ctypes
import os
import re

import pytest
import torch
import yaml

import versband_b200
from oracle import synth
from versband_b200 import _lib, modules

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "versband.h")).read()
    names = set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", hdr))
    names -= {"vb_gemm_args"}
    assert {"vb_gemm", "vb_router", "vb_norm_mod", "vb_attn_smallkv", "vb_conv_post"} <= names
    lib = _lib.lib()
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in versband.h but not exported"
    assert lib.vb_version() == 1
    assert lib.vb_sizeof(0) == ctypes.sizeof(_lib.GemmArgs)


def test_errors_are_reported_not_swallowed():
    lib = _lib.lib()
    args = _lib.GemmArgs()
    rc = lib.vb_gemm(ctypes.byref(args), None)
    assert rc == -1 and b"null operand" in lib.vb_last_error()
    with pytest.raises(_lib.VersbandError):
        _lib.check(rc, "vb_gemm")


def _check_keys(module, manifest_name):
    man = synth.load_manifest(manifest_name)
    sd = module.state_dict()
    assert set(sd.keys()) == set(man.keys()), (sorted(set(man) - set(sd))[:5], sorted(set(sd) - set(man))[:5])
    for k, shape in man.items():
        assert list(sd[k].shape) == shape, k


def test_dit_module_state_dict_matches_reference_manifest():
    net = modules.TxtFlagLargeImprovedDiTV2(**synth.DIT_CFG)
    _check_keys(net, "dit")
    # reference quirk: freshly constructed network is the zero map (final layer / adaLN zeroed)
    assert float(net.final_layer.linear.weight.abs().sum()) == 0.0
    net.load_state_dict(synth.synth_state_dict(synth.load_manifest("dit"), seed=1), strict=True)


def test_vae_module_state_dict_matches_reference_manifest():
    ae = modules.AutoencoderKL(synth.VAE_EMBED_DIM, dict(synth.VAE_DDCONFIG), {"target": "torch.nn.Identity"})
    _check_keys(ae, "vae")
    assert ae.embed_dim == 20


def test_hifigan_module_state_dict_matches_reference_manifest():
    gen = modules.HifiGanGenerator(dict(synth.HIFIGAN_CFG))
    _check_keys(gen, "hifigan")


def test_yaml_targets_resolve_to_b200_modules(tmp_path):
    versband_b200.install_compat()
    cfg = yaml.safe_load("""
unet_config:
  target: ldm.modules.diffusionmodules.vocal2music_moe.TxtFlagLargeImprovedDiTV2
  params: {in_channels: 20, ori_dim: 1024, context_dim: 768, hidden_size: 768, num_heads: 8, depth: 1, max_len: 1500, num_experts: 4}
first_stage_config:
  target: ldm.models.autoencoder1d.AutoencoderKL
  params:
    embed_dim: 20
    ddconfig: {double_z: true, in_channels: 80, out_ch: 80, z_channels: 20, kernel_size: 5, ch: 64, ch_mult: [1, 2], num_res_blocks: 1, attn_layers: [3], down_layers: [0], dropout: 0.0}
    lossconfig: {target: torch.nn.Identity}
""")
    dit = versband_b200.instantiate_from_config(cfg["unet_config"])
    ae = versband_b200.instantiate_from_config(cfg["first_stage_config"])
    assert type(dit).__module__ == "versband_b200.modules" and type(ae).__module__ == "versband_b200.modules"
    import importlib
    assert importlib.import_module("vocoder.hifigan.hifigan").HifiGAN is modules.HifiGAN


def test_no_cpu_fallback():
    net = modules.TxtFlagLargeImprovedDiTV2(20, 768, hidden_size=768, depth=1, num_heads=8, max_len=64).eval()
    x = torch.zeros(1, 20, 8)
    ctx = {"c_crossattn": torch.zeros(1, 80, 1024), "c_concat": {"midi": torch.zeros(1, 1, 16, dtype=torch.long),
                                                                 "beats": torch.zeros(1, 1, 16, dtype=torch.long)}}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, torch.zeros(1, dtype=torch.long), ctx)
    net.train()  # the training forward (versband_b200/train.py) is CUDA-only too
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, torch.zeros(1, dtype=torch.long), ctx)


def test_hifigan_wrapper_loads_checkpoint_dir(tmp_path):
    cfg = dict(synth.HIFIGAN_CFG)
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg))
    sd = synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3, gain=synth.HIFIGAN_GAIN)
    torch.save({"state_dict": {"model_gen": sd}}, tmp_path / "model_ckpt_steps_100.ckpt")
    torch.save({"state_dict": {"model_gen": {k: v * 0 for k, v in sd.items()}}}, tmp_path / "model_ckpt_steps_20.ckpt")
    voc = modules.HifiGAN(str(tmp_path), device="cpu")
    assert torch.equal(voc.model.state_dict()["conv_pre.weight_v"], sd["conv_pre.weight_v"])  # newest ckpt wins
