"""Where the host time of the eager drop-in goes (needs a B200): a TRIVIAL autograd.Function with the drop-in's
signature (no kernels), the same with custom_fwd / custom_bwd, the two raw C-ABI launches (output allocation included),
and the registered drop-in itself, forward + backward at BASELINE config 2, microseconds of host time per iteration.

    python tools/autograd_floor.py
"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xlstm_yolo_clean_b200 as pkg
import xlstm_yolo_clean_b200.backend as be
from torch.amp import custom_fwd, custom_bwd
from oracle import mlstm_oracle as O
inp = O.make_inputs(32, 4, 1600, 64, 64, seed=0, dtype=torch.float32)
t = {k: v.to(torch.bfloat16).cuda() for k, v in inp.items()}
leaves = {k: t[k].detach().requires_grad_(True) for k in ("q","k","v","i","f")}
dh = t["dh"]
def bench(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    t1=time.perf_counter(); torch.cuda.synchronize()
    return (t1-t0)/n*1e6
class Triv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q,k,v,i,f,a,b,c,d,e,g,h,j):
        ctx.save_for_backward(q,k,v,i,f); return q
    @staticmethod
    def backward(ctx, dq):
        q,k,v,i,f = ctx.saved_tensors
        return (dq, k, v, i, f) + (None,)*8
class TrivDec(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.bfloat16)
    def forward(ctx, q,k,v,i,f,a,b,c,d,e,g,h,j):
        ctx.save_for_backward(q,k,v,i,f); return q
    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dq):
        q,k,v,i,f = ctx.saved_tensors
        return (dq, k, v, i, f) + (None,)*8
def run(F):
    def f():
        for p in leaves.values(): p.grad = None
        out = F.apply(leaves["q"],leaves["k"],leaves["v"],leaves["i"],leaves["f"],None,None,None,False,64,1e-6,False,False)
        out.backward(dh)
    return f
def full():
    for p in leaves.values(): p.grad = None
    h = pkg.mlstm_chunkwise__b200(**leaves)
    h.backward(dh)
def raw():
    h, nm, _, cs = be._fw_launch(t["q"],t["k"],t["v"],t["i"],t["f"],None,None,None,None,False,64,1e-6,None,True,False,False)
    nmp = nm.data_ptr()
    be._bw_launch(t["q"],t["k"],t["v"],t["i"],t["f"],nmp,nmp+nm.stride(0)*4,dh,None,None,None,None,None,64,1e-6,None,False,cs,False,False,None)
print("trivial Function fwd+bwd: %.1f us | with custom_fwd/bwd: %.1f us | raw launches: %.1f us | full drop-in: %.1f us" % (bench(run(Triv)), bench(run(TrivDec)), bench(raw), bench(full)))
