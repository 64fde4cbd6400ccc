import torch, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV='cuda:0'
B,T,V,U = 256,1000,5000,200
g = torch.Generator(device=DEV).manual_seed(1234)
ys = torch.randint(4, V, (B, U), device=DEV, generator=g)
emis = torch.log_softmax(torch.randn(B, T, U + 1, device=DEV, generator=g), -1)
mask = torch.ones(B,1,T,dtype=torch.bool,device=DEV); xl = torch.full((B,),T,dtype=torch.int64,device=DEV); yls=torch.full((B,),U,dtype=torch.int64,device=DEV)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(2): va,_,_ = ops.viterbi_align_compact(emis, V, mask, xl, ys, yls, 0)
torch.cuda.synchronize()
s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(n): va,_,_ = ops.viterbi_align_compact(emis, V, mask, xl, ys, yls, 0)
e.record(); torch.cuda.synchronize(); print("viterbi compact ms", s.elapsed_time(e)/n)
