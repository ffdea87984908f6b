"""fp16 <-> bf16 re-rounding pass (mlstm_b200_convert16) against Tensor.to() at the model's q/k tensor size.

    python tools/convert16_bench.py            # prints one JSON line (needs a B200)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xlstm_yolo_clean_b200 as pkg  # noqa: E402


def _time(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    out = {}
    for name, shape in (("base256 qk, S=6400, 32 img", (32, 6400, 1024)), ("base256 v, S=1600, 32 img", (32, 1600, 512))):
        x = torch.randn(*shape, device="cuda").half()
        gb = x.numel() * 4 / 1e9  # 2 bytes read + 2 bytes written per element
        us_t, us_m = _time(lambda: x.to(torch.bfloat16)), _time(lambda: pkg.convert16(x, torch.bfloat16))
        assert torch.equal(pkg.convert16(x, torch.bfloat16), x.to(torch.bfloat16))
        out[name] = {"elements": x.numel(), "torch_to_us": round(us_t, 1), "torch_to_TBps": round(gb / us_t * 1e3, 2),
                     "convert16_us": round(us_m, 1), "convert16_TBps": round(gb / us_m * 1e3, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
