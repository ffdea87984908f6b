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
flat = buf.cpu().view(2, 4096)
for name, k in (("fw", 0), ("bw", 1)):
    ct = flat[k, 3300:3300 + 3 * min(B * NH, 260)].view(-1, 3)
    t0 = ct[:, 0].min().item()
    st, en = (ct[:, 0] - t0).float() / 1e3, (ct[:, 1] - t0).float() / 1e3
    dur = en - st
    print(f"{name} CTA lifetimes (us, globaltimer): start min/max {st.min():.2f}/{st.max():.2f}  end min/max {en.min():.2f}/{en.max():.2f}  "
          f"duration min/median/max {dur.min():.2f}/{dur.median():.2f}/{dur.max():.2f}  cta0 {dur[0]:.2f}  distinct SMs {len(set(ct[:, 2].tolist()))}")
    order = dur.argsort()
    print("   slowest CTAs (cta, sm, start, dur):", [(int(i), int(ct[i, 2]), round(float(st[i]), 2), round(float(dur[i]), 2)) for i in order[-6:]])
    print("   fastest CTAs (cta, sm, start, dur):", [(int(i), int(ct[i, 2]), round(float(st[i]), 2), round(float(dur[i]), 2)) for i in order[:4]])
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
