import torch, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV='cuda:0'
def run(B,Tq,Tk,h,mask=False):
    d=64*h
    q = torch.randn(B,Tq,d,device=DEV).bfloat16(); k = torch.randn(B,Tk,d,device=DEV).bfloat16(); v = torch.randn(B,Tk,d,device=DEV).bfloat16()
    m = torch.ones(B,1,Tk,dtype=torch.bool,device=DEV) if mask else None
    for _ in range(3): y = ops.attention(q,k,v,m,h)
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): y = ops.attention(q,k,v,m,h)
    e.record(); torch.cuda.synchronize(); t=s.elapsed_time(e)/10
    ctas = B*h*((Tq+127)//128)
    print(f"B{B} Tq{Tq} Tk{Tk} mask={mask}: {t*1e3:.1f} us, CTAs {ctas}, waves {ctas/148:.1f}, us per CTA-wave {t*1e3/(ctas/148):.2f}")
for Tk in (128,256,384,512): run(64,512,Tk,8)
run(64,512,512,8,True)
run(2,128,128,8)   # 16 CTAs: single-CTA latency
run(2,128,512,8)
