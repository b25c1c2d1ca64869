"""A host with no Python in it: examples/c_host/run_clip.c (plain C, gcc) runs one clip — conditioning, router
noise, 24-step CFG sampler, VAE decode, vocoder — through the model-level C-ABI entry points on weights exported by
tools/export_weights.py, and must reproduce the Python pipeline's waveform bit for bit."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tools.export_weights import export_pipeline  # noqa: E402
from versband_b200 import ops  # noqa: E402
from versband_b200.pipeline import AccompBandPipeline  # noqa: E402

torch.set_grad_enabled(False)


def _build_host(tmp_path):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "run_clip")
    libdir = os.path.join(ROOT, "versband_b200")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", os.path.join(ROOT, "examples/c_host/run_clip.c"), "-I" + os.path.join(ROOT, "include"),
           "-I" + cuda + "/include", "-L" + libdir, "-l:libversband_sm100.so", "-L" + cuda + "/lib64", "-lcudart",
           "-Wl,-rpath," + libdir, "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.mark.parametrize("T", [750])
def test_plain_c_host_reproduces_the_python_pipeline(tmp_path, T):
    sds = (synth.synth_state_dict(synth.load_manifest("dit"), seed=1234),
           synth.synth_state_dict(synth.load_manifest("vae"), seed=2345),
           synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN))
    scale_factor = 0.8
    pipe = AccompBandPipeline(*sds, dit_cfg=synth.DIT_CFG, ddconfig=synth.VAE_DDCONFIG, embed_dim=synth.VAE_EMBED_DIM,
                              voc_cfg=synth.HIFIGAN_CFG, scale_factor=scale_factor, use_graph=False)
    x0, midi, beats, cap_c, cap_u = synth.dit_inputs(1, T, seed=77)
    seed, offset, guidance = 20240917, 8, 3.0
    dit = pipe.dit
    S, d, E = len(pipe.sampler.sched), dit.depth, dit.E
    # ---- the Python pipeline, stage by stage
    cond2 = dit.prepare_pair(midi.cuda(), beats.cuda(), cap_c.cuda(), cap_u.cuda(), T)
    e0 = torch.empty(S, d, 2 * T, 2, device="cuda")
    e1, e2 = torch.empty(S, d, 2 * T, E, device="cuda"), torch.empty(S, d, 2 * T, E, device="cuda")
    ops.router_noise(seed, offset, T, E, d, S, e0, e1, e2)
    z = pipe.sampler.sample_cfg(x0.cuda(), cond2, None, guidance, noise=torch.cat([e0, e1, e2], -1))
    wav = pipe.voc.forward(pipe.vae.decode(z, scale=1.0 / scale_factor))
    torch.cuda.synchronize()
    # ---- the same clip through the C host
    info = export_pipeline(str(tmp_path / "model.vbw"), pipe)
    print("exported:", info)
    Tm, L, ori = midi.shape[-1], cap_c.shape[1], cap_c.shape[2]
    with open(tmp_path / "case.bin", "wb") as f:
        f.write(struct.pack("<6i", T, Tm, L, ori, S, 0))
        f.write(struct.pack("<2Q", seed, offset))
        f.write(struct.pack("<2f", guidance, scale_factor))
        f.write(np.asarray([t for t, _ in pipe.sampler.sched], dtype=np.int64).tobytes())
        f.write(np.asarray([dt for _, dt in pipe.sampler.sched], dtype=np.float32).tobytes())
        f.write(midi.numpy().astype(np.int64).tobytes())
        f.write(beats.numpy().astype(np.int64).tobytes())
        f.write(torch.cat([cap_c, cap_u], 0).float().numpy().tobytes())
        f.write(x0.float().numpy().tobytes())
    exe = _build_host(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "model.vbw"), str(tmp_path / "case.bin"), str(tmp_path / "out.f32")],
                       capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr
    out = torch.from_numpy(np.fromfile(tmp_path / "out.f32", dtype=np.float32))
    assert out.numel() == wav.numel() == 2 * T * 320
    assert torch.isfinite(out).all() and torch.equal(out, wav.view(-1).cpu())
