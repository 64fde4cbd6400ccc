import torch, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV='cuda:0'
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
for (M,N,K) in [(32000,5000,512),(32000,7168,512),(25600,2048,512),(25600,512,2048),(25600,1536,512),(25600,512,512),(8192,8192,8192)]:
    x = torch.randn(M,K,device=DEV).bfloat16(); w = torch.randn(N,K,device=DEV).bfloat16(); b=torch.randn(N,device=DEV)
    t = timeit(lambda: ops.linear_bf16(x,w,b)); 
    t16 = timeit(lambda: ops.linear_bf16(x,w,b,out_dtype=torch.bfloat16))
    tt = timeit(lambda: torch.nn.functional.linear(x,w))
    fl = 2.0*M*N*K
    print(f"{M}x{N}x{K}: f32-out {t:.3f} ms {fl/t/1e9:.0f} TF/s | bf16-out {t16:.3f} ms {fl/t16/1e9:.0f} TF/s | cublas bf16 {tt:.3f} ms {fl/tt/1e9:.0f} TF/s")
