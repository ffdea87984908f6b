"""Host-buffer entry point: fwd+bwd of the chunkwise mLSTM for tensors that live in (pinned) host
memory, pipelined over batch slices so that the H2D copy of slice j+1, the kernels of slice j and
the D2H copy of slice j-1 overlap (three CUDA streams, full-duplex PCIe), and over consecutive steps (two
alternating sets of device buffers).  The op has no cross-sample term, so slicing the batch axis is exact.
"""
from __future__ import annotations

import torch

from .backend import mlstm_chunkwise_bw, mlstm_chunkwise_fw

_IN = ("q", "k", "v", "i", "f", "dh")
_OUT = ("h", "dq", "dk", "dv", "di", "df")


class _Lane:
    """One set of device input buffers with its own copy / compute streams and captured graphs."""

    def __init__(self, shp, dtype, dev):
        self.d_in = {k: torch.empty(shp[k], dtype=dtype, device=dev) for k in _IN}
        self.s_h2d, self.s_cmp, self.s_d2h, self.launch = (torch.cuda.Stream(dev) for _ in range(4))
        self.graphs, self.keep = {}, {}


class HostFwBw:
    """Two lanes (device buffer sets) alternate between consecutive ``run`` calls, so the H2D copies of step N+1 overlap
    the kernels and the D2H copies of step N: in steady state a step costs what the busier PCIe direction needs for its
    bytes instead of fill + stream + drain.  ``run`` returns immediately; ``flush()`` makes the current stream wait for
    everything issued so far (synchronise that stream, or the device, before reading the outputs)."""

    def __init__(self, B, NH, S, DK, DV, dtype=torch.bfloat16, device="cuda:0", n_slices=8, chunk_size=64, eps=1e-6, taper=True,
                 lanes=2):
        self.dev = torch.device(device)
        self.n_slices = max(1, min(n_slices, B))
        self.chunk_size, self.eps = chunk_size, eps
        shp = dict(q=(B, NH, S, DK), k=(B, NH, S, DK), v=(B, NH, S, DV), i=(B, NH, S), f=(B, NH, S), dh=(B, NH, S, DV),
                   h=(B, NH, S, DV), dq=(B, NH, S, DK), dk=(B, NH, S, DK), dv=(B, NH, S, DV), di=(B, NH, S), df=(B, NH, S))
        self.lanes = [_Lane(shp, dtype, self.dev) for _ in range(max(1, lanes))]
        self._turn = 0
        # tapered batch slices: small first and last slices keep the pipeline's fill (first H2D, nothing else
        # running) and drain (last D2H) short, large middle slices keep the number of copies low
        n = self.n_slices
        w = [min(2.0 ** j, 2.0 ** (n - 1 - j), 8.0) for j in range(n)] if taper else [1.0] * n
        acc, bounds = 0.0, [0]
        for x in w:
            acc += x
            bounds.append(int(round(B * acc / sum(w))))
        self.slices = [slice(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.lanes[0].d_in.values())
        self.d2h_bytes = sum(torch.Size(shp[k]).numel() for k in _OUT) * torch.empty((), dtype=dtype).element_size()

    @staticmethod
    def alloc_host(B, NH, S, DK, DV, dtype=torch.bfloat16):
        shp = dict(h=(B, NH, S, DV), dq=(B, NH, S, DK), dk=(B, NH, S, DK), dv=(B, NH, S, DV), di=(B, NH, S), df=(B, NH, S))
        return {k: torch.empty(v, dtype=dtype).pin_memory() for k, v in shp.items()}

    def flush(self):
        """The current stream waits for every step issued so far (all lanes)."""
        cur = torch.cuda.current_stream(self.dev)
        for lane in self.lanes:
            cur.wait_stream(lane.launch)

    def run(self, host_in: dict, host_out: dict, use_graph: bool = True):
        """host_in: pinned q,k,v,i,f,dh; host_out: pinned h,dq,dk,dv,di,df (filled asynchronously: call ``flush()`` and
        synchronise the current stream, or synchronise the device, before reading them).

        With ``use_graph`` the whole pipeline of a step (copies and kernels on the lane's three streams) is captured once
        per lane and set of host buffers and replayed: a step costs one graph launch instead of ~30 Python-level launches
        per slice, which otherwise bound the step at more than 8 slices."""
        lane = self.lanes[self._turn]
        self._turn = (self._turn + 1) % len(self.lanes)
        cur = torch.cuda.current_stream(self.dev)
        lane.launch.wait_stream(cur)  # ordered after whatever the caller queued before (and after this lane's last step)
        with torch.cuda.stream(lane.launch):
            if not use_graph:
                self._run(lane, host_in, host_out, record=True)
                return host_out
            key = tuple(host_in[k].data_ptr() for k in _IN) + tuple(host_out[k].data_ptr() for k in _OUT)
            g = lane.graphs.get(key)
            if g is None:
                self._run(lane, host_in, host_out, record=True)  # warm-up outside capture (lazy module / attribute setup)
                lane.launch.synchronize()
                g = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(self.dev)
                cap.wait_stream(lane.launch)
                with torch.cuda.graph(g, stream=cap):
                    self._run(lane, host_in, host_out, record=False)
                lane.launch.wait_stream(cap)
                lane.graphs[key] = g
                lane.keep[key] = (host_in, host_out)  # the graph holds raw pointers into these buffers
            g.replay()
        return host_out

    def _run(self, lane, host_in: dict, host_out: dict, record: bool):
        cur = torch.cuda.current_stream(self.dev)
        for s in (lane.s_h2d, lane.s_cmp, lane.s_d2h):
            s.wait_stream(cur)
        keep = []
        for sl in self.slices:
            with torch.cuda.stream(lane.s_h2d):
                for k in _IN:
                    lane.d_in[k][sl].copy_(host_in[k][sl], non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(lane.s_h2d)
            with torch.cuda.stream(lane.s_cmp):
                lane.s_cmp.wait_event(e_in)
                d = {k: lane.d_in[k][sl] for k in _IN}
                h, n_out, m_out, _, cst = mlstm_chunkwise_fw(d["q"], d["k"], d["v"], d["i"], d["f"],
                                                             chunk_size=self.chunk_size, eps=self.eps)
                dq, dk, dv, di, df, _ = mlstm_chunkwise_bw(d["q"], d["k"], d["v"], d["i"], d["f"], n_out, m_out, d["dh"],
                                                           chunk_size=self.chunk_size, eps=self.eps, c_states=cst)
                e_c = torch.cuda.Event()
                e_c.record(lane.s_cmp)
            outs = dict(h=h, dq=dq, dk=dk, dv=dv, di=di, df=df)
            with torch.cuda.stream(lane.s_d2h):
                lane.s_d2h.wait_event(e_c)
                for k in _OUT:
                    host_out[k][sl].copy_(outs[k], non_blocking=True)
                    if record:
                        outs[k].record_stream(lane.s_d2h)
            keep.append((n_out, m_out, cst))
            if record:
                for t in (n_out, m_out, cst):
                    if t is not None:
                        t.record_stream(lane.s_cmp)
        for s in (lane.s_h2d, lane.s_cmp, lane.s_d2h):
            cur.wait_stream(s)
        return host_out
