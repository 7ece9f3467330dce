"""Whole hot-path steps as single CUDA-graph launches (host buffers in, host result out).

The reference's hot loop crosses Python -> ATen / TF once per op; here one *step* of the
sampled-action path (H2D of the minibatch, critic evaluation on the grid, per-state reduction, D2H
of the result) is captured once and replayed with one ``cudaGraphLaunch``: small configs are
launch-latency-bound, large ones still save ~10 launches and copies of host overhead.
Everything inside the graph is librlc kernels and cudaMemcpyAsync on pinned buffers."""
from __future__ import annotations

from typing import Optional

import torch

from .engine import Critic, _f32


class ForwardKLGridStep:
    """ForwardKL sampled-action step on a shared quadrature grid (forwardkl_network.py:160-194 with
    ``get_logprob`` :324-351 fused): ``loss_b, dmean, dlog_std = step(states, mean, log_std)``.

    Inputs are written into the pinned staging tensors ``s_host`` [B,S], ``mean_host`` [B,A],
    ``log_std_host`` [B,A] (or passed to ``__call__`` which copies them there); outputs land in
    the pinned ``loss_host`` [B], ``dmean_host`` [B,A], ``dlog_std_host`` [B,A].  ``q`` [B,N] stays
    on the device (``self.q``)."""

    def __init__(self, critic: Critic, grid, weights, action_scale: float, entropy_scale: float, B: int,
                 precision="auto", b_total: Optional[int] = None, use_graph: bool = True,
                 repack_each_step: bool = False, keep_q: bool = False):
        eng = critic.eng
        # one call per step (rlc_critic_eval_reduce_policy): evaluation kernel + policy-fused reduction kernel; q[B,N] is
        # library scratch unless keep_q (then ``self.q`` holds it)
        self.keep_q = bool(keep_q)
        dev = eng.device
        self.critic, self.eng = critic, eng
        self.grid, self.w = _f32(grid, dev), _f32(weights, dev).reshape(-1)
        self.N, self.A = self.grid.shape
        if self.A != critic.A or self.w.numel() != self.N:
            raise ValueError("grid / weights do not match the critic")
        self.B, self.scale, self.alpha, self.prec = int(B), float(action_scale), float(entropy_scale), precision
        self.b_total = int(b_total or B)
        # training loops change theta between steps: rebuild the tensor-core operand pack inside the step
        self.repack = bool(repack_each_step)
        # one flat pinned input buffer [s | mean | log_std] and one flat output buffer [loss | dmean | dlog_std],
        # mirrored on the device: ONE H2D and ONE D2H copy per step (each small copy costs a few us of latency)
        S, A = critic.S, self.A
        n_in, n_out = B * (S + 2 * A), B * (1 + 2 * A)
        self._in_host = torch.zeros((n_in,), dtype=torch.float32).pin_memory()   # warm-up launches read it: no garbage
        self._out_host = torch.empty((n_out,), dtype=torch.float32).pin_memory()
        self._in_dev = torch.empty((n_in,), dtype=torch.float32, device=dev)
        self._out_dev = torch.empty((n_out,), dtype=torch.float32, device=dev)

        def views(flat, shapes):
            out, off = [], 0
            for sh in shapes:
                n = int(torch.Size(sh).numel())
                out.append(flat[off:off + n].view(sh))
                off += n
            return out
        self.s_host, self.mean_host, self.log_std_host = views(self._in_host, [(B, S), (B, A), (B, A)])
        self.s, self.mean, self.log_std = views(self._in_dev, [(B, S), (B, A), (B, A)])
        self.loss_host, self.dmean_host, self.dlog_std_host = views(self._out_host, [(B,), (B, A), (B, A)])
        self._out_views = tuple(views(self._out_dev, [(B,), (B, A), (B, A)]))
        self.q = torch.empty((B, self.N), dtype=torch.float32, device=dev)
        self._graph = None
        self._stream = torch.cuda.Stream(device=dev)
        if use_graph:
            self._capture()

    def _enqueue(self):
        self._in_dev.copy_(self._in_host, non_blocking=True)
        if self.repack:
            self.critic.invalidate()
        self.critic.eval_reduce_policy(self.s, self.grid, self.w, self.scale, self.mean, self.log_std, self.alpha,
                                       b_total=self.b_total, precision=self.prec, want_q=self.q if self.keep_q else False,
                                       out=self._out_views)
        self._out_host.copy_(self._out_dev, non_blocking=True)

    def _capture(self):
        # warm up outside the capture: workspace growth, operand pack and func attributes allocate/sync
        with torch.cuda.stream(self._stream):
            for _ in range(2):
                self._enqueue()
        self._stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            self._enqueue()
        self._graph = g

    def invalidate(self):
        """Call after the critic's theta changed when ``repack_each_step`` is off: the captured graph
        does not contain the pack kernels, so run one eager step (which rebuilds the pack in place;
        the graph reads the same device buffers afterwards)."""
        self.critic.invalidate()
        with torch.cuda.stream(self._stream):
            self._enqueue()
        self._stream.synchronize()

    def launch(self):
        """Enqueue one step on the step's stream (no host synchronisation)."""
        with torch.cuda.stream(self._stream):          # CUDAGraph.replay() launches on the CURRENT stream
            if self._graph is not None:
                self._graph.replay()
            else:
                self._enqueue()

    def __call__(self, states=None, mean=None, log_std=None):
        if states is not None:
            self.s_host.copy_(torch.as_tensor(states, dtype=torch.float32).reshape(self.s_host.shape))
        if mean is not None:
            self.mean_host.copy_(torch.as_tensor(mean, dtype=torch.float32).reshape(self.mean_host.shape))
        if log_std is not None:
            self.log_std_host.copy_(torch.as_tensor(log_std, dtype=torch.float32).reshape(self.log_std_host.shape))
        self.launch()
        self._stream.synchronize()
        return self.loss_host, self.dmean_host, self.dlog_std_host


class ForwardKLGridPipeline:
    """The same step as :class:`ForwardKLGridStep` for callers that keep the device busy: ``depth`` slots, each with its
    own pinned staging and device buffers, the H2D copy of step i+1 and the D2H copy of step i-1 running on their own
    streams under the kernels of step i.

        p.submit(states, mean, log_std)     # stage + enqueue, returns at once
        loss_b, dmean, dlog_std = p.result()    # oldest outstanding step; blocks until its result is on the host

    Every step still pays its own host->device and device->host copies; what disappears is the serialisation of copy,
    compute and host-side staging (a blocking ``__call__`` costs ~0.15 ms of it per step at cfg4)."""

    def __init__(self, critic: Critic, grid, weights, action_scale: float, entropy_scale: float, B: int,
                 precision="auto", b_total: Optional[int] = None, depth: int = 2, keep_q: bool = False):
        eng = critic.eng
        self.keep_q = bool(keep_q)
        dev = eng.device
        self.critic, self.eng = critic, eng
        self.grid, self.w = _f32(grid, dev), _f32(weights, dev).reshape(-1)
        self.N, self.A = self.grid.shape
        if self.A != critic.A or self.w.numel() != self.N:
            raise ValueError("grid / weights do not match the critic")
        self.B, self.scale, self.alpha, self.prec = int(B), float(action_scale), float(entropy_scale), precision
        self.b_total = int(b_total or B)
        S, A = critic.S, self.A
        n_in, n_out = B * (S + 2 * A), B * (1 + 2 * A)
        self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))

        def views(flat, shapes):
            out, off = [], 0
            for sh in shapes:
                n = int(torch.Size(sh).numel())
                out.append(flat[off:off + n].view(sh))
                off += n
            return out
        self.slots = []
        for _ in range(max(1, int(depth))):
            sl = type("Slot", (), {})()
            sl.in_host = torch.zeros((n_in,), dtype=torch.float32).pin_memory()
            sl.out_host = torch.empty((n_out,), dtype=torch.float32).pin_memory()
            sl.in_dev = torch.empty((n_in,), dtype=torch.float32, device=dev)
            sl.out_dev = torch.empty((n_out,), dtype=torch.float32, device=dev)
            sl.h_in = views(sl.in_host, [(B, S), (B, A), (B, A)])
            sl.d_in = views(sl.in_dev, [(B, S), (B, A), (B, A)])
            sl.h_out = tuple(views(sl.out_host, [(B,), (B, A), (B, A)]))
            sl.d_out = tuple(views(sl.out_dev, [(B,), (B, A), (B, A)]))
            sl.q = torch.empty((B, self.N), dtype=torch.float32, device=dev)
            sl.ev_in, sl.ev_done, sl.ev_out = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            sl.graph = None
            self.slots.append(sl)
        # warm up (workspace growth, operand pack) and capture the kernels of every slot
        for sl in self.slots:
            with torch.cuda.stream(self.s_compute):
                for _ in range(2):
                    self._kernels(sl)
            self.s_compute.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.s_compute):
                self._kernels(sl)
            sl.graph = g
        self._next, self._oldest, self._outstanding = 0, 0, 0

    def _kernels(self, sl):
        self.critic.eval_reduce_policy(sl.d_in[0], self.grid, self.w, self.scale, sl.d_in[1], sl.d_in[2], self.alpha,
                                       b_total=self.b_total, precision=self.prec, want_q=sl.q if self.keep_q else False,
                                       out=sl.d_out)

    def submit(self, states, mean, log_std):
        if self._outstanding == len(self.slots):
            raise RuntimeError("every slot is in flight: call result() first")
        sl = self.slots[self._next]
        for dst, src in zip(sl.h_in, (states, mean, log_std)):
            dst.copy_(torch.as_tensor(src, dtype=torch.float32).reshape(dst.shape))
        with torch.cuda.stream(self.s_in):
            sl.in_dev.copy_(sl.in_host, non_blocking=True)
            sl.ev_in.record(self.s_in)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(sl.ev_in)
            sl.graph.replay()
            sl.ev_done.record(self.s_compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(sl.ev_done)
            sl.out_host.copy_(sl.out_dev, non_blocking=True)
            sl.ev_out.record(self.s_out)
        self._next = (self._next + 1) % len(self.slots)
        self._outstanding += 1

    def result(self):
        """(loss_b [B], dmean [B,A], dlog_std [B,A]) of the oldest outstanding step: pinned host tensors, valid until
        that slot is submitted again."""
        if self._outstanding == 0:
            raise RuntimeError("nothing was submitted")
        sl = self.slots[self._oldest]
        sl.ev_out.synchronize()
        self._oldest = (self._oldest + 1) % len(self.slots)
        self._outstanding -= 1
        return sl.h_out


class GridAgentUpdateStep:
    """The hot-path part of one ForwardKL / ReverseKL ``update_network`` (forwardkl_network.py:123-209,
    reversekl_network.py:130-217) as ONE CUDA graph:

        H2D (s, a, y, mean, log_std[, v])  ->  critic regression step on the B rows (grads, Adam with a
        device-side step count, operand repack)  ->  grid evaluation with the NEW theta  ->  policy
        reduction with the log-density fused  ->  D2H (q_loss, loss_b, dmean, dlog_std)

    ``y`` is the regression target the reference forms from the target V net (``r + gamma V'(s')``, :137-138)
    and ``mean, log_std`` the policy head outputs: V and pi are per-state networks outside the path.  Small
    configs (cfg1: B=32, N=62) are launch-latency-bound: one graph launch replaces ~25 launches + copies.
    Several agents (independent sweep INDEX runs, cfg5) each own a step on its own stream and Engine and
    overlap on one GPU: ``launch()`` them all, then ``wait()``."""

    def __init__(self, critic: Critic, optimizer, grid, weights, action_scale: float, entropy_scale: float,
                 B: int, kind: str = "fkl", hard: bool = False, precision="auto", use_graph: bool = True):
        if kind not in ("fkl", "rkl"):
            raise ValueError("kind must be 'fkl' or 'rkl'")
        eng, dev = critic.eng, critic.eng.device
        self.critic, self.eng, self.opt, self.kind, self.hard = critic, eng, optimizer, kind, bool(hard)
        self.grid, self.w = _f32(grid, dev), _f32(weights, dev).reshape(-1)
        self.N, self.A = self.grid.shape
        self.B, self.scale, self.alpha, self.prec = int(B), float(action_scale), float(entropy_scale), precision
        pin = lambda *sh: torch.zeros(sh, dtype=torch.float32).pin_memory()     # the warm-up launches read these: no garbage
        devt = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)
        S, A = critic.S, self.A
        self.host_in = dict(s=pin(B, S), a=pin(B, A), y=pin(B), mean=pin(B, A), log_std=pin(B, A), v=pin(B))
        self.dev_in = dict(s=devt(B, S), a=devt(B, A), y=devt(B), mean=devt(B, A), log_std=devt(B, A), v=devt(B))
        self.host_out = dict(q_loss=pin(1), loss_b=pin(B), dmean=pin(B, A), dlog_std=pin(B, A))
        self.q = devt(B, self.N)
        self._graph = None
        self._stream = torch.cuda.Stream(device=dev)
        if use_graph:
            with torch.cuda.stream(self._stream):
                for _ in range(2):
                    self._enqueue()
            self._stream.synchronize()
            # the warm-up advanced the optimiser twice; that is part of this object's contract (see reset_optimizer)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._enqueue()
            self._graph = g

    def reset_optimizer(self):
        """Zero the Adam moments and step count (e.g. after the capture warm-up updates)."""
        self.opt.m.zero_()
        self.opt.v.zero_()
        self.opt.state_dev.zero_()

    def _enqueue(self):
        d, h = self.dev_in, self.host_in
        for k in ("s", "a", "y", "mean", "log_std") + (("v",) if self.kind == "rkl" else ()):
            d[k].copy_(h[k], non_blocking=True)
        q_loss, _ = self.opt.step_graph_safe(d["s"], d["a"], d["y"])          # theta changes; pack invalidated
        self.critic.eval_into(d["s"], self.grid, self.q, self.prec)
        if self.kind == "fkl":
            loss_b, dm, ds, _ = self.eng.fkl_policy(self.q, self.w, self.grid, self.scale, d["mean"], d["log_std"],
                                                    self.alpha)
        else:
            loss_b, dm, ds, _ = self.eng.rkl_policy(self.q, d["v"], self.w, self.grid, self.scale, d["mean"],
                                                    d["log_std"], self.alpha, hard=self.hard)
        o = self.host_out
        o["q_loss"].copy_(q_loss, non_blocking=True)
        o["loss_b"].copy_(loss_b, non_blocking=True)
        o["dmean"].copy_(dm, non_blocking=True)
        o["dlog_std"].copy_(ds, non_blocking=True)

    def set_inputs(self, **arrays):
        for k, x in arrays.items():
            self.host_in[k].copy_(torch.as_tensor(x, dtype=torch.float32).reshape(self.host_in[k].shape))

    def launch(self):
        with torch.cuda.stream(self._stream):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._enqueue()

    def wait(self):
        self._stream.synchronize()
        return self.host_out

    def __call__(self, **arrays):
        self.set_inputs(**arrays)
        self.launch()
        return self.wait()
