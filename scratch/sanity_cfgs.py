import sys, time, torch
sys.path.insert(0,'.')
import bench
from types import SimpleNamespace
from cassnat_asr_public_b200.cassnat import make_model
dev='cuda:0'
def run(wname, B, over, lm=None, prec='bf16', conv=False):
    w = bench.WORKLOADS[wname]; args = bench.model_args(w, **over)
    if conv:
        args.use_conv_dec=True; args.pos_type='relative'; args.use_conv_enc=True; args.d_decff=1024; args.d_encff=1024
    torch.manual_seed(1234)
    model = make_model(80,args).to(dev).eval(); model.set_precision(prec)
    enc_h, mask, ratio, lens = bench.synth_encoder_output(w, B, 1234)
    inp = [t.to(dev) for t in (enc_h,mask,ratio)]
    for i in range(3):
        torch.manual_seed(5)
        torch.cuda.synchronize(); t0=time.perf_counter()
        with torch.no_grad(): out,_ = model.decode_from_encoder(*inp, bench.VOCAB, args, lm_model=lm)
        torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(f"{wname} B={B} {over} conv={conv} prec={prec}: {dt*1e3:.2f} ms/step, hyp0 len {len(out[0][0]['hyp'])}, mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)
    del model; torch.cuda.empty_cache()
run('cfg1', 16, {})
run('cfg1', 16, {}, prec='fp32')
class LM(torch.nn.Module):
    def __init__(s): super().__init__(); g=torch.Generator().manual_seed(1); s.e=torch.nn.Parameter(torch.randn(5000,64,generator=g)); s.o=torch.nn.Parameter(torch.randn(64,5000,generator=g)/8)
    def forward(s, ids, mask): return torch.log_softmax(s.e[ids] @ s.o, -1)
run('cfg3', 32, dict(sample_num=10, threshold=0.9), lm=LM().to(dev))
run('cfg3', 64, {}, conv=True)
run('cfg3', 64, dict(left_trigger=1, right_trigger=1, src_trigger=True))
