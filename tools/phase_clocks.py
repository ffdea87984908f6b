"""Debug: per-tile phase clocks of CTA 0 for the tensor-core kernels (config 2 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["MLSTM_B200_LIB"] = os.path.join(ROOT, "xlstm_yolo_clean_b200", "lib", "libmlstm_b200_prof.so")  # built by `python __graft_entry__.py --profile`
import torch
import xlstm_yolo_clean_b200 as pkg
from xlstm_yolo_clean_b200 import _cabi
from oracle import mlstm_oracle as O

B, NH, S = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 4, 1600)))
inp = O.make_inputs(B, NH, S, 64, 64, seed=0, dtype=torch.float32)
t = {k: v.to(torch.bfloat16).cuda() for k, v in inp.items()}
lib = _cabi.load_library()
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
for _ in range(3):
    saved = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
    pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], saved[1], saved[2], t["dh"], c_states=saved[4])
lib.mlstm_b200_debug_set_clock_buffer(buf.data_ptr())
saved = pkg.mlstm_chunkwise_fw(t["q"], t["k"], t["v"], t["i"], t["f"])
pkg.mlstm_chunkwise_bw(t["q"], t["k"], t["v"], t["i"], t["f"], saved[1], saved[2], t["dh"], c_states=saved[4])
torch.cuda.synchronize()
lib.mlstm_b200_debug_set_clock_buffer(None)
v = buf.cpu().view(2, 256, 16)
NT = (S + 127) // 128
for name, k in (("fw", 0), ("bw", 1)):
    print(name, "slots: 0 top | 1 after sync(gates) | 2 after prep+sync | 3 S ready | 4 after P/W+sync | 5 dC ready | 6 dC done | 7 H/main ready | 8 epilogue done")
    e = v[k, 200]
    t0, tl = v[k, 0, 0].item(), v[k, NT - 1, 0].item()
    print(f"  kernel entry -> tile 0 top {t0 - e[0].item()} clk; last tile top -> workers done {e[1].item() - tl} clk; "
          f"-> after final CTA sync {e[2].item() - tl} clk; entry -> exit {e[2].item() - e[0].item()} clk")
    for tile in range(NT):
        r = v[k, tile]
        base = r[0].item()
        print(f"  tile {tile:2d}: " + " ".join(f"{(r[i].item() - base):6d}" for i in range(9)) + "  || ctl " +
              " ".join(f"{(r[i].item() - base):6d}" for i in range(9, 16)) +
              (f"   | next top {(v[k, tile + 1, 0].item() - base):6d}" if tile + 1 < NT else ""))
