import torch, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV='cuda:0'
B,T,U = 256,1000,200
g = torch.Generator(device=DEV).manual_seed(1)
# alignment with exactly U non-blank frames per utterance
al = torch.zeros(B, T, dtype=torch.int64, device=DEV)
for b in range(B):
    pos = torch.randperm(T-1, device=DEV, generator=g)[:U].sort().values + 1
    al[b, pos] = torch.randint(4, 5000, (U,), device=DEV, generator=g)
mask = torch.ones(B,1,T,dtype=torch.bool,device=DEV); xl = torch.full((B,),T,dtype=torch.int64,device=DEV); yls=torch.full((B,),U,dtype=torch.int64,device=DEV)
for _ in range(3): m,_,_ = ops.align_to_mask(al, yls, U, mask, xl, 0)
torch.cuda.synchronize()
s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): m,_,_ = ops.align_to_mask(al, yls, U, mask, xl, 0)
e.record(); torch.cuda.synchronize(); print("align_to_mask ms", s.elapsed_time(e)/10, "out", tuple(m.shape), m.dtype)
