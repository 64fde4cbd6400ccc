import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_backward_gpu import ref_attention, rel, leaf, DEV
from cassnat_asr_public_b200 import ops
for (B,h,Tq,Tk,dk) in [(2,2,5,9,8),(1,1,5,9,8),(1,1,5,9,64),(1,1,64,64,8),(2,2,5,9,32),(2,2,5,9,64)]:
    d=h*dk
    q,k,v = leaf(B,Tq,d,seed=1), leaf(B,Tk,d,seed=2), leaf(B,Tk,d,seed=3)
    out = ops.attention(q,k,v,None,h)
    go = torch.randn(B,Tq,d,device=DEV)
    out.backward(go)
    got=[t.grad.clone() for t in (q,k,v)]
    for t in (q,k,v): t.grad=None
    ref = ref_attention(q,k,v,None,h); ref.backward(go)
    print((B,h,Tq,Tk,dk), "fwd", rel(out,ref), "dq dk dv", [rel(g,t.grad) for g,t in zip(got,(q,k,v))])
