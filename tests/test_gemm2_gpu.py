"""The 2-CTA (tcgen05 cta_group::2, cluster of two CTAs) GEMM path: same contract as vb_gemm, forced
on for every legal problem and checked against the same PyTorch references."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_gemm_gpu as G  # noqa: E402
from versband_b200 import _lib, ops  # noqa: E402


@pytest.fixture(autouse=True)
def force_two_cta():
    _lib.lib().vb_set_gemm2(2)
    yield
    _lib.lib().vb_set_gemm2(0)


@pytest.mark.parametrize("M,N,K", [(1500, 2304, 768), (750, 768, 768), (300, 256, 128), (1500, 192, 512), (129, 1024, 64)])
def test_plain_fp16(M, N, K):
    G.test_plain(M, N, K, 0)


@pytest.mark.parametrize("M,N,K", [(1500, 2304, 768), (257, 192, 512), (750, 768, 768)])
def test_plain_tf32(M, N, K):
    G.test_tf32_plain(M, N, K)


def test_epilogues_and_batched():
    G.test_epilogue_bias_act_gate_residual_fp16()
    G.test_swiglu_pairs()
    G.test_qkv_rope_epilogue()
    G.test_batched_attention_shapes()


@pytest.mark.parametrize("Cin,Cout,k,dil,T", [(128, 256, 3, 1, 750), (256, 256, 11, 5, 1000), (1536, 1536, 3, 1, 750)])
def test_conv(Cin, Cout, k, dil, T):
    G.test_conv1d_implicit(Cin, Cout, k, dil, T)


def test_many_tiles_persistent_loop():
    """More tile pairs than clusters: exercises accumulator double-buffering across the pair."""
    torch.manual_seed(0)
    M, N, K = 20000, 512, 256
    a = torch.randn(M, K, device="cuda").half()
    b = torch.randn(N, K, device="cuda").half()
    out = torch.zeros(M, N, device="cuda")
    ops.gemm(a, b, out_f32=out, out_ld=N)
    G.close(out, a.float() @ b.float().t())
