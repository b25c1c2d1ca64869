"""MoE expert GEMMs of one DiT block (2x750 tokens): separate launches vs vb_gemm_chain variants, device time
per block from a CUDA graph of 20 repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from versband_b200 import ops, _lib

torch.set_grad_enabled(False)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
N, D, Fh, E, band = 1500, 768, 512, 4, 192
H16 = torch.float16
# routing: 2N rows into 8 buckets, each padded to 128 rows
idx = torch.randint(0, 2 * E, (2 * N,), device=dev, generator=g)
counts = torch.bincount(idx, minlength=2 * E)
tiles = ((counts + 127) // 128).tolist()
m_used = sum(tiles)
max_rows = ((2 * N + 127) // 128 + 2 * E) * 128
tile_group = torch.zeros(max_rows // 128, device=dev, dtype=torch.int32)
row_map = torch.full((max_rows,), -1, device=dev, dtype=torch.int32)
pos = 0
order = torch.argsort(idx)
starts = torch.cumsum(torch.tensor([0] + counts.tolist()), 0)
for e in range(2 * E):
    tile_group[pos // 128: pos // 128 + tiles[e]] = e
    n_e = int(counts[e])
    row_map[pos:pos + n_e] = order[starts[e]:starts[e] + n_e].int()
    pos += tiles[e] * 128
m_tiles = torch.tensor([m_used], device=dev, dtype=torch.int32)
row_scale = torch.rand(max_rows, device=dev, generator=g)
xs = torch.randn(max_rows, D, device=dev, generator=g).half()
w13 = (torch.randn(2 * E, 2 * Fh, D, device=dev, generator=g) / D ** 0.5).half()
w2 = (torch.randn(2 * E, D, Fh, device=dev, generator=g) / Fh ** 0.5).half()
f13 = (torch.randn(E, 2 * Fh, 2 * band, device=dev, generator=g) / band ** 0.5).half()
f2 = (torch.randn(E, band, Fh, device=dev, generator=g) / Fh ** 0.5).half()
hs = torch.zeros(max_rows, Fh, device=dev, dtype=H16)
y2 = torch.zeros(2 * N, D, device=dev, dtype=H16)
hf = torch.zeros(E, N, Fh, device=dev, dtype=H16)
gate = torch.rand(2, 6 * D, device=dev, generator=g)
hbuf = torch.randn(N, D, device=dev, generator=g)
xcur = torch.zeros(N, D, device=dev)
T = 750
yb = y2.view(2 * N, E, band).permute(1, 0, 2)


def probs(bn=0):
    return [
        ops.gemm_args(xs, w13, M=max_rows, b_c2=(0, 0, 1), tile_group=tile_group, m_tiles_used=m_tiles, swiglu=True,
                      out_fp16=hs, outb_ld=Fh, bn=bn),
        ops.gemm_args(hs, w2, M=max_rows, b_c2=(0, 0, 1), tile_group=tile_group, m_tiles_used=m_tiles, row_map=row_map,
                      rows_out=2 * N, row_scale=row_scale, out_fp16=y2, outb_ld=D, bn=bn),
        ops.gemm_args(yb, f13, M=N, G1=E, kpt=band // 64, dil=N, pad=0, a_c2=(1, 0), b_c2=(1, 0, 0), swiglu=True,
                      out_fp16=hf, outb_ld=Fh, outb_g=(N * Fh, 0), bn=bn),
        ops.gemm_args(hf, f2, M=N, G1=E, a_c2=(1, 0), b_c2=(1, 0, 0), gate=gate, gate_rows=T, gate_ld=6 * D,
                      gate_g1=band, residual=hbuf, out_f32=xcur, out_ld=D, out_g=(band, 0), bn=bn),
    ]


cnt = torch.zeros(64, 256, device=dev, dtype=torch.int32)
L = _lib.lib()


def separate(bn=0, only=None):
    def run(i):
        for j, pr in enumerate(probs(bn)):
            if only is None or j == only:
                _lib.check(L.vb_gemm(ops.C.byref(pr), ops._stream()), "vb_gemm")
    return run


def chained(bn=128, which=(0, 1, 2, 3), deps=(0, 1, 2, 1)):
    def run(i):
        ps = probs()
        ops.gemm_chain([ps[j] for j in which], [deps[j] if k else 0 for k, j in enumerate(which)], cnt[i], bn=bn)
    return run


def single_chains(bn=128):
    def run(i):
        ps = probs()
        for j in range(4):
            ops.gemm_chain([ps[j]], [0], cnt[i], bn=bn)
    return run


def timeit(label, fn, reps=20):
    for i in range(2):
        cnt.zero_()
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        cnt.zero_()
        for i in range(reps):
            fn(i)
    gr.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"{label:58s} {a.elapsed_time(b) / (5 * reps) * 1e3:7.1f} us", flush=True)


print("m_used", m_used, "tiles", tiles)
timeit("separate, auto tiles", separate())
for j, nm in enumerate(("w13", "w2", "f13", "f2")):
    timeit(f"  {nm} alone auto", separate(0, j))
    timeit(f"  {nm} alone bn=128", separate(128, j))
    timeit(f"  {nm} alone as 1-problem chain bn=128", chained(128, (j,)))
timeit("separate, bn=128", separate(128))
timeit("4 single-problem chains bn=128", single_chains(128))
timeit("chain bn=128", chained(128))
timeit("chain bn=64", chained(64))
for dbg, nm in ((1, "no deps (wrong)"), (2, "no fences (wrong)"), (4, "m-fastest"), (5, "no deps, m-fastest (wrong)")):
    L.vb_set_chain_debug(dbg)
    timeit(f"chain bn=128 dbg={dbg} {nm}", chained(128))
L.vb_set_chain_debug(0)
timeit("chain w13->w2 only", chained(128, (0, 1)))
timeit("chain f13->f2 only", chained(128, (2, 3)))
