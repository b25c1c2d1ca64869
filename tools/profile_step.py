"""One eager (non-graph) pipeline step between cudaProfilerStart/Stop for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--clips B] [--t-lat T]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from versband_b200.pipeline import AccompBandPipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=1)
ap.add_argument("--t-lat", type=int, default=750)
ap.add_argument("--graph", action="store_true")
args = ap.parse_args()
torch.set_grad_enabled(False)
sds = (synth.synth_state_dict(synth.load_manifest("dit"), seed=1234),
       synth.synth_state_dict(synth.load_manifest("vae"), seed=2345),
       synth.synth_state_dict(synth.load_manifest("hifigan"), seed=3456, gain=synth.HIFIGAN_GAIN))
pipe = AccompBandPipeline(*sds, dit_cfg=synth.DIT_CFG, ddconfig=synth.VAE_DDCONFIG, embed_dim=synth.VAE_EMBED_DIM,
                          voc_cfg=synth.HIFIGAN_CFG, use_graph=args.graph)
x0, midi, beats, cap_c, cap_u = [t.cuda() for t in synth.dit_inputs(args.clips, args.t_lat, seed=1)]
for _ in range(2):
    pipe.generate(midi, beats, cap_c, cap_u, x0=x0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
pipe.generate(midi, beats, cap_c, cap_u, x0=x0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
