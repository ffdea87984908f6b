import os, sys, torch
sys.path.insert(0, os.getcwd())
import xlstm_yolo_clean_b200 as pkg
from oracle import mlstm_oracle as O
dev = torch.device("cuda:0")
# cell output stage: ragged rows, all three head dims
for NH, D in ((8, 64), (12, 32), (6, 128)):
    H = NH * D
    h = torch.randn(3, NH, 333, D, device=dev).to(torch.bfloat16).requires_grad_(True)
    x = torch.randn(3, 333, H, device=dev).to(torch.float16).requires_grad_(True)
    w, b, sk = (torch.randn(H, device=dev).requires_grad_(True) for _ in range(3))
    y = pkg.cell_out(h, w, b, sk, x, out_dtype=torch.float16)
    y.backward(torch.randn_like(y))
# mLSTM: ragged S on the tensor path, both directions, d = 32 / 64 / 128 (block backward), with states
for D, S in ((64, 100), (32, 52), (128, 324), (64, 1004)):
    inp = O.make_inputs(2, 3, S, D, D, seed=1, dtype=torch.float32, with_states=True)
    t = {k: v.to(torch.bfloat16).to(dev) for k, v in inp.items()}
    for rev in (False, True):
        leaves = {k: t[k].detach().requires_grad_(True) for k in ("q", "k", "v", "i", "f")}
        c0 = t["c0"].detach().requires_grad_(True)
        h, (cl, nl, ml) = pkg.mlstm_chunkwise__b200(**leaves, c_initial=c0, n_initial=t["n0"], m_initial=t["m0"],
                                                    return_last_states=True, chunk_size=4, reverse=rev,
                                                    autocast_kernel_dtype=torch.bfloat16)
        torch.autograd.backward([h, cl], [t["dh"], t["dc_last"].to(cl.dtype)])
torch.cuda.synchronize()
print("probe done")
