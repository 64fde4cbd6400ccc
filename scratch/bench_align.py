import torch, time, sys
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV='cuda:0'
B,T,V,U = 256,1000,5000,200
g = torch.Generator(device=DEV).manual_seed(1234)
ys = torch.randint(4, V, (B, U), device=DEV, generator=g)
logits = torch.randn(B, T, V, device=DEV, generator=g)
t_idx = (torch.arange(U, device=DEV) * 5)[None, :, None] + torch.arange(1, 4, device=DEV)[None, None, :]
b_idx = torch.arange(B, device=DEV)[:, None, None].expand(B, U, 3)
logits[b_idx, t_idx.expand(B, U, 3), ys[:, :, None].expand(B, U, 3)] += 8.0
lp = torch.log_softmax(logits, -1); del logits
mask = torch.ones(B,1,T,dtype=torch.bool,device=DEV); xl = torch.full((B,),T,dtype=torch.int64,device=DEV); yls=torch.full((B,),U,dtype=torch.int64,device=DEV)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts)//2]
t = timeit(lambda: ops.row_stats(lp)); print(f"row_stats argmax: {t:.3f} ms  {B*T*V*4/t/1e6:.0f} GB/s")
t = timeit(lambda: ops.row_stats(lp, top2=True, peak=True)); print(f"row_stats top2: {t:.3f} ms  {B*T*V*4/t/1e6:.0f} GB/s")
t = timeit(lambda: ops.best_path_align(lp, mask, xl, 0)); print(f"best_path_align: {t:.3f} ms")
t = timeit(lambda: ops.viterbi_align(lp, mask, xl, ys, yls, 0)); S=2*U+1; lat = B*T*S*5+B*T*9
print(f"viterbi_align: {t:.3f} ms  lattice {lat/t/1e6:.0f} GB/s")
va,_,_ = ops.viterbi_align(lp, mask, xl, ys, yls, 0)
t = timeit(lambda: ops.align_to_mask(va, yls, U, mask, xl, 0)); print(f"align_to_mask: {t:.3f} ms  {(B*T*8+B*(U+1)*T)/t/1e6:.0f} GB/s")
# small training-like shape
B2,T2,U2=32,250,40
lp2 = lp[:B2,:T2].contiguous(); 
t = timeit(lambda: ops.viterbi_align(lp2, mask[:B2,:,:T2], xl[:B2].clamp(max=T2), ys[:B2,:U2], yls[:B2].clamp(max=U2), 0)); print(f"viterbi small: {t:.3f} ms")
