import torch, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV='cuda:0'
B,Tq,Tk,h = 64,390,500,8
d=64*h
q = torch.randn(B,Tq,d,device=DEV).bfloat16(); k = torch.randn(B,Tk,d,device=DEV).bfloat16(); v = torch.randn(B,Tk,d,device=DEV).bfloat16()
lens = torch.randint(300,501,(B,),device=DEV); mask = (torch.arange(Tk,device=DEV)[None]<lens[:,None]).unsqueeze(1)
for _ in range(3): y = ops.attention(q,k,v,mask,h)
torch.cuda.synchronize()
s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): y = ops.attention(q,k,v,mask,h)
e.record(); torch.cuda.synchronize(); print("ms per call", s.elapsed_time(e)/10, "TF/s", 4.0*B*Tq*Tk*d/(s.elapsed_time(e)/10)/1e9)
for _ in range(3): y = ops.attention(q,k,v,None,h)
s.record()
for _ in range(10): y = ops.attention(q,k,v,None,h)
e.record(); torch.cuda.synchronize(); print("no-mask ms per call", s.elapsed_time(e)/10)
