"""Host-side checks of the model-level C-ABI that need no GPU: descriptor plumbing through the workspace queries
(pure host code: no kernel is launched), their error behaviour, the relocatable weight export for non-Python hosts,
and that the plain-C host example compiles and links against the in-tree library."""
import ctypes as C
import os
import shutil
import struct
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tools import export_weights  # noqa: E402
from versband_b200 import _lib  # noqa: E402
from versband_b200.dit import DiTEngine  # noqa: E402
from versband_b200.vae import VAEEngine  # noqa: E402
from versband_b200.vocoder import HifiGanEngine  # noqa: E402


@pytest.fixture(scope="module")
def engines():
    dit = DiTEngine(synth.synth_state_dict(synth.load_manifest("dit"), seed=1), device="cpu")
    vae = VAEEngine(synth.synth_state_dict(synth.load_manifest("vae"), seed=2), synth.VAE_DDCONFIG, synth.VAE_EMBED_DIM, device="cpu")
    voc = HifiGanEngine(synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3, gain=synth.HIFIGAN_GAIN), synth.HIFIGAN_CFG,
                        device="cpu")
    return dit, vae, voc


def test_workspace_queries_and_error_strings(engines):
    dit, vae, voc = engines
    L = _lib.lib()
    d, p = dit._desc(), dit._prep_desc()
    fwd = [L.vb_dit_workspace_bytes(C.byref(d), 2, T) for T in (64, 750, 1500)]
    assert 0 < fwd[0] < fwd[1] < fwd[2] and all(v % 256 == 0 for v in fwd)
    assert L.vb_dit_workspace_bytes(C.byref(d), 8, 750) > fwd[1]
    smp = L.vb_sample_workspace_bytes(C.byref(d), 1, 750, 24)
    assert smp > fwd[1]  # the DiT workspace of the CFG pair + the sampler's own buffers
    assert L.vb_dit_prepare_workspace_bytes(C.byref(d), C.byref(p), 1, 1500, 2, 80, 750) > 0
    # beyond max_len (freqs_cis[:T], vocal2music_moe.py:421), empty batch, caption longer than the 96 key slots
    assert L.vb_dit_workspace_bytes(C.byref(d), 2, 1501) == -1 and b"max_len" in L.vb_last_error()
    assert L.vb_dit_workspace_bytes(C.byref(d), 0, 750) == -1
    assert L.vb_dit_prepare_workspace_bytes(C.byref(d), C.byref(p), 1, 1500, 2, 97, 750) == -1 and b"caption" in L.vb_last_error()
    assert L.vb_dit_prepare_workspace_bytes(C.byref(d), C.byref(p), 1, 1500, 2, 80, 760) == -1  # |T - Tm/2| > 2
    dec, enc = vae._desc(), vae._desc_enc()
    assert 0 < L.vb_vae_workspace_bytes(C.byref(dec), 1, 100) < L.vb_vae_workspace_bytes(C.byref(dec), 1, 750)
    assert 0 < L.vb_vae_encode_workspace_bytes(C.byref(enc), 1, 96) < L.vb_vae_encode_workspace_bytes(C.byref(enc), 2, 96)
    assert dec.n_ops == 13 and enc.n_ops == 10  # 3 mid + 9 resblocks + 1 upsample; 6 resblocks + 1 downsample + 3 mid
    v = voc._desc()
    assert 0 < L.vb_voc_workspace_bytes(C.byref(v), 1, 50) < L.vb_voc_workspace_bytes(C.byref(v), 1, 1500)
    v2 = type(v).from_buffer_copy(v)
    v2.n_kernels = 4
    assert L.vb_voc_workspace_bytes(C.byref(v2), 1, 50) == -1 and b"unsupported generator" in L.vb_last_error()


def test_weight_export_is_relocatable(engines, tmp_path):
    dit, vae, voc = engines
    path = str(tmp_path / "model.vbw")
    structs = [("dit", dit._desc()), ("prep", dit._prep_desc()), ("vae", vae._desc()), ("voc", voc._desc())]
    info = export_weights.export(path, structs, [dit.w, vae.w, voc.w])
    raw = open(path, "rb").read()
    assert raw[:4] == b"VBW1"
    n_structs, data_bytes = struct.unpack_from("<iq", raw, 4)
    assert n_structs == 4 and data_bytes == info["data_bytes"]
    data = raw[len(raw) - data_bytes:]
    off = 16
    for name, st in structs:
        assert raw[off:off + 16].rstrip(b"\0").decode() == name
        nbytes, n_reloc = struct.unpack_from("<qq", raw, off + 16)
        assert nbytes == C.sizeof(st) and n_reloc > 0
        body = raw[off + 32: off + 32 + nbytes]
        relocs = struct.unpack_from(f"<{n_reloc}q", raw, off + 32 + nbytes)
        for r in relocs:
            v = struct.unpack_from("<Q", body, r)[0]
            assert 1 <= v <= data_bytes and (v - 1) % 256 == 0
        off += 32 + nbytes + 8 * n_reloc
    assert off == len(raw) - data_bytes
    # one tensor back out of the blob: the first pointer field of vb_dit_desc is proj_in_w
    d = dit._desc()
    first = export_weights._pointer_fields_of(type(d))[0]
    body0 = raw[48: 48 + C.sizeof(d)]
    o = struct.unpack_from("<Q", body0, first)[0] - 1
    t = dit.w["proj_in_w"]
    assert data[o:o + t.numel() * t.element_size()] == t.contiguous().view(torch.uint8).numpy().tobytes()


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_plain_c_host_compiles_and_links(tmp_path):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no CUDA toolkit headers")
    _lib.lib()  # the in-tree library must exist (conftest builds it when missing)
    libdir = os.path.join(ROOT, "versband_b200")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples/c_host/run_clip.c"),
           "-I" + os.path.join(ROOT, "include"), "-I" + cuda + "/include", "-L" + libdir, "-l:libversband_sm100.so",
           "-L" + cuda + "/lib64", "-lcudart", "-Wl,-rpath," + libdir, "-o", str(tmp_path / "run_clip")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
