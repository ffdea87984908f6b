"""PCIe probe: H2D alone, D2H alone, both at once (pinned host memory), sizes of the bench's e2e step."""
import torch
n = 105_676_800
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    for s in (s1, s2): torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s per direction")
# chunked: 48 copies per direction (the pipeline's pattern at 8 slices), no dependencies between directions
def chunked(nchunk):
    step = n // nchunk
    def fn():
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for c in range(nchunk):
            with torch.cuda.stream(s1): d_a[c * step:(c + 1) * step].copy_(h_in[c * step:(c + 1) * step], non_blocking=True)
            with torch.cuda.stream(s2): h_out[c * step:(c + 1) * step].copy_(d_b[c * step:(c + 1) * step], non_blocking=True)
    return fn
for nc in (8, 48, 96):
    ms = t(chunked(nc))
    print(f"both, {nc} chunks per direction: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s per direction")
# dependent: D2H chunk c waits for H2D chunk c (event), like the pipeline
def dependent(nchunk):
    step = n // nchunk
    def fn():
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for c in range(nchunk):
            with torch.cuda.stream(s1):
                d_a[c * step:(c + 1) * step].copy_(h_in[c * step:(c + 1) * step], non_blocking=True)
                e = torch.cuda.Event(); e.record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(e)
                h_out[c * step:(c + 1) * step].copy_(d_b[c * step:(c + 1) * step], non_blocking=True)
    return fn
for nc in (8, 48):
    ms = t(dependent(nc))
    print(f"dependent, {nc} chunks: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s per direction")
