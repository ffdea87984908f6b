import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xlstm_yolo_clean_b200 as pkg
for (B,NH,S,D) in [(16,6,6400,128),(16,6,1600,128),(16,6,400,128)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    t = {k: (0.3*torch.randn(B,NH,S,D,generator=g,device="cuda")).to(torch.bfloat16) for k in "qkv"}
    i = torch.full((B,NH,S), -8.7, device="cuda").to(torch.bfloat16); f = (3+3*torch.rand(B,NH,S,generator=g,device="cuda")).to(torch.bfloat16)
    fn = lambda: pkg.mlstm_chunkwise_fw(t["q"],t["k"],t["v"],i,f,chunk_size=16,save_states=False)
    fn(); torch.cuda.synchronize()
    side=torch.cuda.Stream(); keep=[]
    with torch.cuda.stream(side):
        gr=torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(4): keep.append(fn())
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    a,b_=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b_.record(); torch.cuda.synchronize()
    ms=a.elapsed_time(b_)/4
    byt=B*NH*S*(4*D*2+12)
    print(S, D, round(ms*1e3,1), "us", round(byt/ms/1e6), "GB/s", round(byt/ms/1e6/6535.7,3))
