"""Debug: compare the buffers of vb_dit_forward's workspace with the Python launch sequence (depth-1 model)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from versband_b200.dit import DiTEngine

torch.set_grad_enabled(False)
DEV = "cuda"
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, T = 2, 750
sd = synth.synth_state_dict(synth.load_manifest("dit"), seed=1234)
eng = DiTEngine(sd, device=DEV, depth=depth)
x, midi, beats, cap, _ = synth.dit_inputs(B, T, seed=3)
t = torch.tensor([125, 833]).to(DEV)
cond = eng.prepare(midi.to(DEV), beats.to(DEV), cap.to(DEV), T)
expo = [tuple(e.to(DEV) for e in blk) for blk in synth.exp_noise(4, B * T, 4, seed=5)][:depth]
x = x.to(DEV)
v_c = eng.forward(x, t, cond, expo)[0].clone()
ws = eng._workspace(B, T)
cws = ws["cws"]
v_p = eng.forward_py(x, t, cond, expo)[0]
torch.cuda.synchronize()
print("v equal:", torch.equal(v_c, v_p), float((v_c - v_p).abs().max()))
D, H, E, Fh = eng.D, eng.H, eng.E, eng.Fh
N, hd, Tp = B * T, eng.hd, (T + 7) // 8 * 8
max_rows = ((2 * N + 127) // 128 + 2 * E) * 128
off = 0
def take(nbytes, dtype, shape):
    global off
    off = (off + 255) & ~255
    tns = cws[off:off + nbytes].view(dtype).view(*shape)
    off += nbytes
    return tns
H16, F32, I32 = torch.float16, torch.float32, torch.int32
c = {}
c["bucket_count"] = take(depth * 2 * E * 4, I32, (depth, 2 * E))
c["usage"] = take(depth * 2 * E * 4, F32, (depth, 2 * E))
c["x_cl"] = take(B * T * 32 * 2, H16, (B, T, 32))
c["x"] = take(N * D * 4, F32, (N, D))
c["h"] = take(N * D * 4, F32, (N, D))
c["xa"] = take(N * D * 2, H16, (N, D))
c["q"] = take(N * D * 2, H16, (B, H, T, hd))
c["k"] = take(N * D * 2, H16, (B, H, T, hd))
c["vt"] = take(B * H * hd * Tp * 2, H16, (B, H, hd, Tp))
c["oa"] = take(N * D * 2, H16, (N, D))
c["xm"] = take(N * D * 2, H16, (N, D))
c["cap_part"] = take(N * H * E * 4, F32, (N, H, E))
for n in ("idx_c", "idx_a", "slot_c", "slot_a"):
    c[n] = take(N * 4, I32, (N,))
for n in ("scale_c", "scale_a"):
    c[n] = take(N * 4, F32, (N,))
c["bucket_start"] = take(2 * E * 4, I32, (2 * E,))
c["row_map"] = take(max_rows * 4, I32, (max_rows,))
c["row_scale"] = take(max_rows * 4, F32, (max_rows,))
c["tile_group"] = take(max_rows // 128 * 4, I32, (max_rows // 128,))
c["m_tiles"] = take(4, I32, (1,))
c["xs"] = take(max_rows * D * 2, H16, (max_rows, D))
c["hs"] = take(max_rows * Fh * 2, H16, (max_rows, Fh))
c["y2"] = take(2 * N * D * 2, H16, (2 * N, D))
c["hf"] = take(E * N * Fh * 2, H16, (E, N, Fh))
c["xf"] = take(N * D * 2, H16, (N, D))
for n in ("xa", "q", "k", "vt", "oa", "h", "xm", "cap_part", "idx_c", "idx_a", "scale_c", "scale_a", "y2", "hf", "x", "xf"):
    a, b = c[n], ws[n]
    if n == "vt":
        a, b = a[..., :T], b[..., :T]
    d = (a.float() - b.float()).abs().max()
    print(f"{n:10s} equal={torch.equal(a, b)} maxdiff={float(d):.3e}")
print("lb", ws["lb"].tolist())
for call in range(2):
    v_c, lb_c = eng.forward(x, t, cond, expo)
    v_c, lb_c = v_c.clone(), lb_c.clone()
    v_p, lb_p = eng.forward_py(x, t, cond, expo)
    torch.cuda.synchronize()
    print("call", call, torch.equal(v_c, v_p), torch.equal(lb_c, lb_p), lb_c.tolist(), lb_p.tolist())
    print(" counters C:", c["bucket_count"].tolist(), " py:", ws["bucket_count"].tolist())
