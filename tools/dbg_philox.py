import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from versband_b200 import ops
N,E,depth,steps=32,4,4,1
gen=torch.Generator(device="cuda"); gen.manual_seed(987654321)
off0=gen.get_offset()
ref=[torch.zeros(steps,depth,2*N,w,device="cuda") for w in (2,E,E)]
for s in range(steps):
    for br in range(2):
        for k in range(depth):
            for i in range(3):
                ref[i][s,k,br*N:(br+1)*N].exponential_(generator=gen)
mine=[torch.zeros_like(r) for r in ref]
ops.router_noise(gen.initial_seed(), off0, N, E, depth, steps, *mine)
torch.cuda.synchronize()
print("seed",gen.initial_seed(),"off0",off0,"off1",gen.get_offset())
for a,b in zip(mine,ref):
    neq=(a!=b).sum().item(); print("mismatch",neq,"of",a.numel(),"maxdiff",(a-b).abs().max().item())
print(mine[0][0,0,:4]); print(ref[0][0,0,:4])
# is it a permutation?
a=mine[0][0,0,:N].flatten().sort()[0]; b=ref[0][0,0,:N].flatten().sort()[0]
print("sorted equal:", torch.equal(a,b), (a-b).abs().max().item())
# uniform check: torch uniform with same state
gen.manual_seed(987654321)
u=torch.empty(64,device="cuda").uniform_(generator=gen)
print("u",u[:4], (-torch.log(u[:4])))
