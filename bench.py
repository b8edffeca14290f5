#!/usr/bin/env python
"""bench.py -- point clouds/s through the HPCS hot path (fwd+bwd), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload train|decode]

Workload (BASELINE.json configs[1]): ShapeNet-shaped training step, batch 32 clouds x 1024 points per
GPU, k=20, 32-d embeddings, 50 mined triplets per anchor.  One "step" is one pass of the hot path over
one batch of synthetic input, in the order a training step runs it:
    kNN(D=3)  -> edge features C=1  forward      (vn_dgcnn_partseg.py:65)
    kNN(D=63) -> edge features C=21 forward  x2  (vn_dgcnn_partseg.py:70,75)
    fused Poincare triplet loss fwd+bwd over 1.64 M mined triplets ('easy' filter in-kernel)
    edge features backward x3 (last layer first)
The dense VN/conv layers between these ops are outside the path, so the inputs of
the 63-d layers, the upstream gradients of the edge features and the embeddings are synthetic tensors.
Clouds are independent: ranks own disjoint batches (weak scaling); the only collective is the
all-reduce of the `scale` gradient, the one trainable parameter on this path.

`value`   : clouds/s with every input resident in HBM (CUDA events, max over ranks).
`e2e`     : same step through the public API from pinned HOST buffers: H2D of the step's inputs (points, the two
            63-d layer inputs, embeddings, mined triplets) and D2H of its result (loss, kept, d scale) inside the
            timed region; the input gradients stay in HBM for the caller's backward.
`roofline`: the single kernel with the largest share of the step (the persistent edge-backward gather); per-op
            figures, including that backward together with its reverse-graph build, under "ops".
`peaks`   : pipe / cache peaks measured in this process by csrc/peaks.cu (fp32 FMA, ALU, fp64, TF32 tcgen05, L2
            gathers) next to MEASURED_PEAKS.json's HBM figure: every entry of "ops" has a denominator.
`decode`  : BASELINE configs[4] sub-record (single + complete linkage, 8 clouds per GPU = 64 on 8 GPUs, and 64 per GPU).
`--impl reference`: the oracle's restatement of the reference's PyTorch CPU path, on host cores, same batch (32).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, N_PTS, K_NN, C_FEAT, D_EMB, T_PER_ANCHOR = 32, 1024, 20, 21, 32, 50
SCALE, TEMPERATURE = 1e-3, 0.05
METRIC = "point clouds/sec (1024 pts, k=20) fwd+bwd"
UNIT = "clouds/s"
# the one collective of a training step: all-reduce of the fp32 gradients -- VN_DGCNN_partseg's 1,303,850
# parameters (out=32, 16 categories) + CosFace W[32,50] + scale
GRAD_BUCKET_FLOATS = 1_303_850 + D_EMB * 50 + 1
MIN_TIMED_MS = 250.0                   # the timed region replays the step until it lasts at least this long
# thread-level fp32 instructions per mined triplet in hyp_triplet_kernel<8,1> (loss forward + gradient state), from the
# ncu capture profiles/r01_g_ncu_full.md: 303.2 FFMA + 123.5 FADD + 140.4 FMUL per triplet at the bench shape
LOSS_FLOP_PER_TRIPLET = 2 * 303.2 + 123.5 + 140.4


# DRAM bytes per op from the committed `ncu --set full` capture (profiles/r02_ncu_full.md): read + written per launch
# (edge bwd gather 333.0 + 8.8..11.2 MB; edge fwd 13.5 + 272.5 MB, below the algorithmic 343.8 MB because the tail of the
# output is still in L2 when the kernel ends; + the reverse-graph build 5.3 MB for the two-kernel backward)
NCU_TRAFFIC_BYTES = {"edge_bwd_c21": 348.3e6, "edge_bwd_gather_c21": 343.0e6, "edge_fwd_c21": 286.0e6}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# synthetic ShapeNet-shaped inputs
# ------------------------------------------------------------------------------------------------
def synth_inputs(B, seed):
    gen = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, N_PTS, 3, generator=gen)
    pts = pts - pts.mean(1, keepdim=True)
    pts = (pts / pts.norm(dim=-1).amax(1).view(B, 1, 1)).transpose(1, 2).contiguous().view(B, 1, 3, N_PTS)
    f1 = torch.randn(B, C_FEAT, 3, N_PTS, generator=gen)
    f2 = torch.randn(B, C_FEAT, 3, N_PTS, generator=gen)
    u = torch.randn(B * N_PTS, D_EMB, generator=gen)
    nrm = u.norm(dim=-1, keepdim=True)
    emb = torch.tanh(nrm.clamp(max=15)) * u / nrm                    # ExpMap(N(0,1))
    # labels: one of 16 categories per cloud, 2-6 parts each, Dirichlet(1) proportions, global ids 0..49
    labels = []
    for b in range(B):
        cat = int(torch.randint(0, 16, (1,), generator=gen))
        parts = 2 + int(torch.randint(0, 5, (1,), generator=gen))
        probs = torch._sample_dirichlet(torch.ones(parts), generator=gen)
        labels.append(torch.multinomial(probs, N_PTS, replacement=True, generator=gen) + (cat * 3) % 45)
    labels = torch.cat(labels)
    return {"pts": pts, "f1": f1, "f2": f2, "emb": emb, "labels": labels}


def clustered_features(B, seed, centres=6):
    """[B, 63, N] 'backbone-like' features for the D=63 kNN: every cloud is a mixture of a few tight clusters (points of
    one part have similar features after an EdgeConv layer), all clouds share a common offset of a few sigma and the
    channels are correlated through a random low-rank mixing."""
    gen = torch.Generator().manual_seed(seed)
    D = 3 * C_FEAT
    cen = torch.randn(B, centres, D, generator=gen)
    which = torch.randint(0, centres, (B, N_PTS), generator=gen)
    x = torch.gather(cen, 1, which.unsqueeze(-1).expand(-1, -1, D)) + 0.15 * torch.randn(B, N_PTS, D, generator=gen)
    mix = torch.eye(D) + 0.3 * torch.randn(D, 8, generator=gen) @ torch.randn(8, D, generator=gen) / 8 ** 0.5
    x = x @ mix + 2.0 * torch.randn(1, 1, D, generator=gen)
    return x.transpose(1, 2).contiguous()


def algorithmic_work(B):
    """Per-op algorithmic bytes / flops (DESIGN.md 'Kernels and rooflines')."""
    n = B * N_PTS
    E = N_PTS * K_NN
    T0 = T_PER_ANCHOR * n

    def edge(C):
        out = B * 2 * C * 3 * E * 4
        return out + B * 3 * C * N_PTS * 4 + B * E * 8, out + B * E * 8 + B * 3 * C * N_PTS * 4

    e1f, e1b = edge(1)
    e21f, e21b = edge(C_FEAT)
    return {
        "knn_d3": {"flop": B * N_PTS * N_PTS * 9.0, "bytes": B * (12 * N_PTS + 8 * E)},
        "knn_d63": {"flop": B * N_PTS * N_PTS * (2 * 63 + 3.0), "bytes": B * (252 * N_PTS + 8 * E)},
        "edge_fwd_c1": {"bytes": e1f}, "edge_bwd_c1": {"bytes": e1b},
        "edge_fwd_c21": {"bytes": e21f}, "edge_bwd_c21": {"bytes": e21b},
        "edge_bwd_gather_c21": {"bytes": e21b}, "edge_rev_build": {"bytes": B * E * 8},
        "knn_d63_clustered": {"flop": B * N_PTS * N_PTS * (2 * 63 + 3.0), "bytes": B * (252 * N_PTS + 8 * E)},
        # L2 traffic of the loss: three 128-byte rows gathered per mined triplet; HBM: 12 bytes of int32 indices per
        # triplet in, table + gradient out.  flop: measured instruction mix (LOSS_FLOP_PER_TRIPLET)
        "hyp_loss_fwd_bwd": {"flop": T0 * LOSS_FLOP_PER_TRIPLET, "l2_bytes": T0 * 3.0 * D_EMB * 4,
                             "bytes": T0 * 24.0 + 3 * n * D_EMB * 4},
    }


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons under load (nvml, one sample every 2 ms).  NVML is initialised before the
    thread starts so the first sample lands inside even a millisecond-long timed region; samples are tagged with
    the region they were taken in ("value" = the timed steps, "ops"/"e2e" = the later timed regions)."""

    NAMES = ("HwSlowdown:hw_slowdown", "HwThermalSlowdown:hw_thermal_slowdown",
             "SwThermalSlowdown:sw_thermal_slowdown", "SwPowerCap:sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, {}, set(), None
        self.region = "value"
        self._stop_evt = threading.Event()
        self._nv = self._h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._bits = {getattr(nv, "nvmlClocksThrottleReason" + n.split(":")[0]): n.split(":")[1] for n in self.NAMES}
        except Exception as exc:                                # nvml missing: report, do not fail the bench
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def sample(self):
        nv, h = self._nv, self._h
        self.samples.setdefault(self.region, []).append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for bit, name in self._bits.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if self._nv is None:
            return
        try:
            while not self._stop_evt.is_set():
                self.sample()
                self._stop_evt.wait(0.002)
        except Exception as exc:
            self.reasons.add(f"nvml_error:{type(exc).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = {k: sorted(v)[len(v) // 2] for k, v in self.samples.items() if v}
        every = sorted(x for v in self.samples.values() for x in v)
        out = {"sm_mhz": med.get("value", every[len(every) // 2] if every else None), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": {k: len(v) for k, v in self.samples.items()},
               "sm_mhz_by_region": med}
        return out


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import hpcs_b200 as hb
    from hpcs_b200 import _lib, dist as hdist
    import torch.distributed as dist

    rank, world, local = hdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = B_PER_GPU
    host = synth_inputs(B, seed=rank)
    torch.manual_seed(1000 + rank)
    trip_host = hb.get_balanced_random_triplet_indices(host["labels"], t_per_anchor=T_PER_ANCHOR, fraction=0.0)
    T0 = trip_host[0].numel()

    d = {k: v.to(dev) for k, v in host.items()}
    trip = tuple(t.to(dev) for t in trip_host)
    scale = torch.tensor([SCALE], device=dev)
    gen = torch.Generator(device=dev).manual_seed(rank)
    E = N_PTS * K_NN
    g1 = torch.randn(B, 2, 3, N_PTS, K_NN, device=dev, generator=gen)
    g2 = torch.randn(B, 2 * C_FEAT, 3, N_PTS, K_NN, device=dev, generator=gen)
    g3 = torch.randn(B, 2 * C_FEAT, 3, N_PTS, K_NN, device=dev, generator=gen)

    # gradient bucket of a training step: the backbone / CosFace gradients are produced by layers outside this
    # path, so their slots hold synthetic values; d loss / d scale, the one gradient this path produces, is the last slot
    bucket = torch.randn(GRAD_BUCKET_FLOATS, device=dev, generator=gen) if world > 1 else None

    def loss_fwd_bwd(emb_in, tr):
        emb = emb_in.detach().requires_grad_(True)
        sc = scale.detach().requires_grad_(True)            # fresh leaves: keeps autograd on the capturing stream
        loss, kept = hb.hyp_triplet_loss(emb, tr, sc, TEMPERATURE, "easy", 0.0, return_kept=True)
        ge, gs = torch.autograd.grad(loss, (emb, sc))
        return loss, kept, ge, gs

    def step(inp, tr):
        """The order a training step runs these ops in: the three layers' kNN + edge features (forward), the loss
        forward+backward on the embeddings, then the edge-feature backwards from the last layer to the first.  The
        all-reduce of d loss / d scale (the path's only parameter) is issued as soon as the loss backward has produced
        it and overlaps the edge backwards; it carries the whole 5.2 MB fp32 gradient bucket of the model (the reference's
        DDP exchange), not just that scalar."""
        xs = [inp[k].detach().requires_grad_(True) for k in ("pts", "f1", "f2")]
        ys = []
        for x in xs:
            Bc, C, _, Np = x.shape
            idx = hb.knn(x.detach().view(Bc, 3 * C, Np), K_NN)
            ys.append(hb.get_graph_feature(x, K_NN, idx=idx))
        loss, kept, ge, gs = loss_fwd_bwd(inp["emb"], tr)
        work = None
        if world > 1:
            bucket[-1:].copy_(gs.reshape(1))
            work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, async_op=True)
        gxs = [torch.autograd.grad(y, x, g)[0] for y, x, g in zip(ys[::-1], xs[::-1], (g3, g2, g1))]
        if work is not None:
            work.wait()
            gs = bucket[-1:] / world
        return loss, kept, gs, (gxs[2], gxs[1], gxs[0], ge)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def capture_fn(fn):
        """Warm up on a side stream, then capture one call of fn over static buffers into a CUDA graph."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = fn()
        return graph, outs

    def capture(inp, tr):
        return capture_fn(lambda: step(inp, tr))

    # ---- resident-input timing: K replays of the captured step ------------------------------------
    for _ in range(max(args.warmup, 3)):
        out = step(d, trip)
    launches_per_step = None
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    out = step(d, trip)
    launches_per_step = _lib.launch_count() - l0
    use_graph = not args.no_graph
    if use_graph:
        graph, out = capture(d, trip)
        run_step = graph.replay
    else:
        def run_step():
            nonlocal out
            out = step(d, trip)
    R = max(1, args.batches_per_step)               # one timed "step" = R batches back to back (a >= 250 ms region)
    for _ in range(max(args.warmup, 3) * R):
        run_step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps * R):
        run_step()
    t1.record()
    if sampler._nv is not None:
        sampler.sample()                                    # the replays are queued: the GPU is busy right now
    barrier()
    sampler.region = "ops"
    launches = launches_per_step * args.steps * R
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps                   # R batches
    ms_per_batch = ms_per_step / R
    value = world * B * R / (ms_per_step * 1e-3)
    loss_val, kept_val = float(out[0].detach()), int(out[1])

    # ---- per-op durations: every op of the step captured into its OWN CUDA graph and replayed K times
    # between CUDA events on the launching stream (no host gaps; an op = all kernels of one C-ABI call).
    from hpcs_b200 import graph as hgraph
    op_ms, op_launches = {}, {}

    def time_op(name, fn, calls_per_step):
        l0_ = _lib.launch_count()
        fn()
        op_launches[name] = (_lib.launch_count() - l0_, calls_per_step)
        g_, _ = capture_fn(fn)
        for _ in range(3):
            g_.replay()
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(args.steps * R):
            g_.replay()
        e_.record()
        torch.cuda.synchronize()
        op_ms[name] = s_.elapsed_time(e_) / (args.steps * R)

    with torch.no_grad():
        x3 = d["pts"].view(B, 3, N_PTS)
        x63 = d["f1"].view(B, 3 * C_FEAT, N_PTS)
        idx3, idx63 = hb.knn(x3, K_NN), hb.knn(x63, K_NN)
        time_op("knn_d3", lambda: hb.knn(x3, K_NN), 1)
        time_op("knn_d63", lambda: hb.knn(x63, K_NN), 2)
        # the same kNN on backbone-like features: a few tight clusters per cloud, off-centre, correlated channels (the
        # N(0,1) features above are the tensor-core path's best case: no row needs the exact redo)
        xclu = clustered_features(B, seed=100 + rank).to(dev)
        knn_stats = {}
        for tag, xin in (("gaussian", x63), ("clustered", xclu)):
            st_ = {}
            hb.knn(xin, K_NN, stats=st_)
            knn_stats[tag] = {"exact_redo_rows": st_.get("fallback_rows"), "second_chance_rows": st_.get("second_chance_rows")}
        time_op("knn_d63_clustered", lambda: hb.knn(xclu, K_NN), 0)
        time_op("edge_fwd_c1", lambda: hgraph.edge_features_forward(d["pts"], idx3), 1)
        time_op("edge_fwd_c21", lambda: hgraph.edge_features_forward(d["f1"], idx63), 2)
        time_op("edge_bwd_c1", lambda: hgraph.edge_features_backward(g1, d["pts"], idx3), 1)
        time_op("edge_bwd_c21", lambda: hgraph.edge_features_backward(g2, d["f1"], idx63), 2)
        # the same backward split into its two kernels: the reverse-graph build (idx only; inside a step it runs on a
        # second stream under the forward) and the persistent gather, the kernel with the largest share of the step
        rev63 = hgraph.build_reverse_graph(idx63, overlap=False)
        if rev63 is not None:
            time_op("edge_rev_build", lambda: hgraph.build_reverse_graph(idx63, overlap=False), 3)
            time_op("edge_bwd_gather_c21", lambda: hgraph.edge_features_backward(g2, d["f1"], idx63, prebuilt=rev63), 2)
    time_op("hyp_loss_fwd_bwd", lambda: loss_fwd_bwd(d["emb"], trip), 1)
    barrier()

    # ---- end-to-end from pinned host buffers ----------------------------------------------------------
    # Every step uploads its inputs from pinned host memory and downloads its result (loss, kept, d scale).  Two device
    # buffer sets alternate, so the upload of step i+1 and the download of step i-1 overlap the compute of step i on
    # separate streams; all of it is inside the timed region.  Two variants:
    #   "device" (the headline e2e): points, both 63-d layer inputs, embeddings and the sampling plan (label-sorted point
    #            order, 131 KB, + 1 KB of per-label segments, both derived from the labels on the host) go up;
    #            the 1.64 M triplets are drawn on the GPU inside the step (hpcs_triplet_sample_i32);
    #   "host":  as the reference does it -- triplets sampled on the host beforehand and uploaded (int32) every step.
    #   "points_only": like "device", but the layer inputs and embeddings (activations of device-resident layers in the real
    #            model, synthetic stand-ins here) are not re-uploaded: the host->device traffic of an actual training step.
    pin = {k: host[k].pin_memory() for k in ("pts", "f1", "f2", "emb")}
    # what the data loader hands a training step: un-rotated clouds [B,N,3] and (drawn on the host like the reference does) the
    # randn(B,4) of the SO(3) rotation; the rotation + transpose run on the device (row f-4)
    raw_pin = host["pts"].view(B, 3, N_PTS).transpose(1, 2).contiguous().pin_memory()
    rot_pin = torch.randn(B, 4, generator=torch.Generator().manual_seed(77 + rank)).pin_memory()
    trip_pin = tuple(t.to(torch.int32).pin_memory() for t in trip_host)       # int32 indices: half the upload
    order_host, seg_host, T0_plan = hb.triplet_plan(host["labels"], T_PER_ANCHOR, 0.0)   # from the host-side labels, like the host sampler
    order_pin, seg_pin = order_host.pin_memory(), seg_host.pin_memory()
    assert T0_plan == T0
    s_up, s_run, s_down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def flat_outs(outs):
        # the step's result as the caller reads it on the host: loss, surviving triplets, d loss / d scale.  The
        # four input gradients stay in HBM, where the backbone's backward consumes them.
        loss, kept, gs, _grads = outs
        return [loss.detach().reshape(1), kept.reshape(1), gs.reshape(1)]

    def run_e2e(variant):
        ups = dict(pin)
        if variant in ("device", "points_only"):
            ups.update(order=order_pin, seg=seg_pin)
        else:
            ups.update(ta=trip_pin[0], tp=trip_pin[1], tn=trip_pin[2])
        resident = {}
        if variant == "points_only":                     # what a real step takes from the host: the clouds, the rotation draws, the plan
            for k in ("f1", "f2", "emb"):
                resident[k] = ups.pop(k)
            ups.pop("pts")
            ups.update(pts_raw=raw_pin, rot=rot_pin)
        h2d = sum(v.numel() * v.element_size() for v in ups.values())

        samp_state = hb.sampler_state(1234 + rank, dev)     # (seed, step) on the device: every replay draws new triplets

        def step_of(inp):
            if variant == "points_only":                 # rotate + transpose on the device, then the step proper
                inp = dict(inp, pts=hb.rotate_points(inp["pts_raw"], "so3", inp["rot"]).view(B, 1, 3, N_PTS))
            if variant in ("device", "points_only"):
                tr = hb.sample_triplets_device(None, plan=(inp["order"], inp["seg"], T0_plan), state=samp_state)
            else:
                tr = (inp["ta"], inp["tp"], inp["tn"])
            return step(inp, tr)

        sets = []
        for _ in range(2):
            inp = {k: torch.empty_like(v, device=dev) for k, v in ups.items()}
            for k in ups:
                inp[k].copy_(ups[k])
            for k, v in resident.items():
                inp[k] = d[k]
            if use_graph:
                g_, outs = capture_fn(lambda: step_of(inp))
            else:
                g_, outs = None, None
            sets.append({"inp": inp, "graph": g_, "outs": outs})
        probe = flat_outs(sets[0]["outs"] if use_graph else step_of(sets[0]["inp"]))
        res_pin = [[torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in probe] for _ in range(2)]
        d2h = sum(o.numel() * o.element_size() for o in probe)
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_run = [torch.cuda.Event() for _ in range(2)]
        ev_down = [torch.cuda.Event() for _ in range(2)]

        def loop(n_steps):
            for i in range(n_steps):
                st_ = sets[i % 2]
                with torch.cuda.stream(s_up):
                    s_up.wait_event(ev_run[i % 2])               # buffer set free (its previous compute finished)
                    for k in ups:
                        st_["inp"][k].copy_(ups[k], non_blocking=True)
                    ev_up[i % 2].record(s_up)
                with torch.cuda.stream(s_run):
                    s_run.wait_event(ev_up[i % 2])
                    s_run.wait_event(ev_down[i % 2])             # previous results of this set already copied out
                    if use_graph:
                        st_["graph"].replay()
                        outs = st_["outs"]
                    else:
                        outs = step_of(st_["inp"])
                    ev_run[i % 2].record(s_run)
                with torch.cuda.stream(s_down):
                    s_down.wait_event(ev_run[i % 2])
                    for dst, src in zip(res_pin[i % 2], flat_outs(outs)):
                        dst.copy_(src, non_blocking=True)
                    ev_down[i % 2].record(s_down)

        loop(4)
        barrier()
        te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        te0.record()
        loop(args.steps * R)
        for s_ in (s_up, s_run, s_down):
            torch.cuda.current_stream().wait_stream(s_)
        te1.record()
        barrier()
        ms_e = torch.tensor([te0.elapsed_time(te1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        return {"value": round(world * B * R / (ms_e.item() / args.steps * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * R,
                "d2h_bytes_per_step": int(d2h) * R, "loss": float(res_pin[(args.steps * R - 1) % 2][0])}

    sampler.region = "peaks"
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import peaks as peak_probes
    probes = peak_probes.measure(dev)
    decode_rec = decode_subrecord(hb, dev, rank) if not args.no_decode else None
    edgeconv_rec = edgeconv_subrecord(hb, dev, B, rank) if not args.no_decode else None
    model_rec = model_step_subrecord(dev, B, rank) if not args.no_decode else None
    barrier()
    sampler.region = "e2e"
    e2e_dev = run_e2e("device")
    e2e_host = run_e2e("host")
    e2e_pts = run_e2e("points_only")
    clocks = sampler.stop()

    if rank != 0:
        return
    pk = peaks()
    work = algorithmic_work(B)
    ops = {}
    for name, avg_ms in op_ms.items():
        n_launch, calls_per_step = op_launches[name]
        w = work[name]
        sec = avg_ms * 1e-3
        ent = {"ms": round(avg_ms, 4), "calls_per_step": calls_per_step, "kernels_per_call": n_launch,
               "share": round(avg_ms * calls_per_step / ms_per_batch, 4)}
        if name == "edge_rev_build":
            # two passes over idx (read) + the reverse lists (written); the kernel is a latency chain, the HBM figure
            # only says how far from a bandwidth limit it sits
            ent.update(bound="latency (shared-memory atomics, prefix sums; one wave of 128 CTAs); HBM figure for scale",
                       achieved=round(w["bytes"] / sec / 1e9, 1), peak=pk["hbm_gbs"], unit="GB/s")
        elif name.startswith("edge"):
            ent.update(bound="hbm", achieved=round(w["bytes"] / sec / 1e9, 1), peak=pk["hbm_gbs"], unit="GB/s")
        elif name == "knn_d3":
            ent.update(bound="fp32 issue (distances: FFMA2; selection: ALU pipe)", achieved=round(w["flop"] / sec / 1e12, 3),
                       peak=probes.get("fp32_ffma2_tflops"), unit="TFLOP/s",
                       alu_note="selection-bound by design: ~900 warp instructions per row, 160 of them distance FFMA2")
        elif name.startswith("knn_d63"):
            gram = B * N_PTS * N_PTS * 2.0 * 64
            ent.update(bound="ALU pipe (selection epilogue) + L2 gathers (re-rank); tensor pipe is the stated roofline",
                       achieved=round(gram / sec / 1e12, 3), peak=probes.get("tf32_umma_tflops"), unit="TFLOP/s",
                       gram_passes=2, fallback_rows=knn_stats.get("clustered" if name.endswith("clustered") else "gaussian"))
        else:                                                 # hyp_loss_fwd_bwd
            ent.update(bound="l2 (row gathers + vector reductions)", achieved=round(w["l2_bytes"] / sec / 1e9, 1),
                       peak=probes.get("l2_gather_128B_gbs"), unit="GB/s",
                       fp32_tflops=round(w["flop"] / sec / 1e12, 3), fp32_peak_tflops=probes.get("fp32_ffma2_tflops"),
                       flop_per_triplet=round(LOSS_FLOP_PER_TRIPLET, 1),
                       flop_source="ncu instruction counters (FFMA/FADD/FMUL per mined triplet), profiles/r01_g_ncu_full.md")
            if ent["fp32_peak_tflops"]:
                ent["fp32_frac"] = round(ent["fp32_tflops"] / ent["fp32_peak_tflops"], 4)
        if ent.get("peak"):
            ent["frac"] = round(ent["achieved"] / ent["peak"], 4)
        ops[name] = ent
    # the dominant KERNEL (one launch per call) among those with an HBM roofline; "edge_bwd_c21" is that kernel plus the
    # reverse-graph build and "share" double-counts it, so multi-kernel ops are left out of the choice
    single = [n for n in ops if ops[n].get("peak") and ops[n]["kernels_per_call"] == 1 and n != "edge_rev_build"]
    top = max(single, key=lambda n: ops[n]["share"])
    roofline = {"kernel": top, "bound": ops[top]["bound"], "achieved": ops[top]["achieved"], "peak": ops[top]["peak"],
                "unit": ops[top]["unit"], "frac": ops[top]["frac"], "traffic": NCU_TRAFFIC_BYTES.get(top),
                "traffic_source": "profiles/r02_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                "algorithmic_bytes": work[top].get("bytes"), "peak_source": pk["source"]}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "shapenet_train_step_b32_n1024_k20_d32_t50 (BASELINE.json configs[1])",
                   "batches_per_step": R, "ms_per_batch": round(ms_per_batch, 4),
                   "step_definition": f"one timed step = {R} batches of {B} clouds per GPU back to back (a {round(ms_per_step * args.steps)} ms timed region); every per-batch figure below is per ONE batch",
                   "collective": (f"all-reduce of the {GRAD_BUCKET_FLOATS * 4} B fp32 gradient bucket (SURVEY 8e), async, overlapped with the edge backwards; NCCL_MAX_NCHANNELS={os.environ.get('NCCL_MAX_NCHANNELS')} NCCL_NVLS_NCHANNELS={os.environ.get('NCCL_NVLS_NCHANNELS')}" if world > 1 else "none at 1 GPU"),
                   "clouds_per_gpu": B, "points": N_PTS, "k": K_NN, "feat_channels": [1, C_FEAT, C_FEAT],
                   "emb_dim": D_EMB, "triplets_mined": T0, "triplets_kept": kept_val, "filter": "easy",
                   "scale": SCALE, "temperature": TEMPERATURE, "parallelism": f"dp{world}",
                   "l2": "per-step working set ~1.4 GB (edge-feature tensors) exceeds the 126 MB L2; no explicit flush",
                   "launch": "cuda-graph replay of one captured step" if use_graph else "eager",
                   "ops_timing": "each op captured into its own CUDA graph, replayed steps x batches_per_step times between CUDA events, after the timed region; edge_bwd_c21 = edge_rev_build + edge_bwd_gather_c21 (listed separately as well)",
                   "e2e_pipeline": "2 device buffer sets; upload / compute / download on 3 streams",
                   "overlap": "the reverse graph of each layer's backward is built on a second stream during that layer's forward"},
        "e2e": {"value": e2e_pts["value"], "unit": UNIT, "h2d_bytes_per_step": e2e_pts["h2d_bytes_per_step"],
                "d2h_bytes_per_step": e2e_pts["d2h_bytes_per_step"],
                "inputs": "what a training step takes from the host, every batch, from pinned memory: the un-rotated clouds [B,N,3], the "
                          "rotation draws randn(B,4), the sampling plan (label-sorted order + segments).  Rotation + transpose run on the "
                          "device (row f-4), the triplets are drawn on the GPU inside the step (fresh per replay); the 63-d layer inputs and "
                          "the embeddings are activations of device-resident layers in a real step (synthetic stand-ins here) and stay in "
                          "HBM.  Results read back every batch: loss, kept, d scale"},
        "e2e_all_inputs_uploaded": dict({k: e2e_dev[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
                                        note="conservative variant (round 1's headline): the layer inputs and embeddings are re-uploaded every batch too"),
        "e2e_host_sampled_triplets": {k: e2e_host[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "peaks": dict(probes, hbm_gbs=pk["hbm_gbs"], hbm_source=pk["source"]),
        "ops": ops,
        "decode": decode_rec,
        "edgeconv": edgeconv_rec,
        "model_step": model_rec,
        "loss": loss_val,
        "e2e_loss": e2e_pts["loss"],
        "e2e_loss_host_sampled": e2e_host["loss"],
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_pass(steps=1, warmup=0)        # bounded: one 32-cloud step, 10-25 s of CPU work
        try:
            line["torch_gpu_baseline"] = torch_gpu_reference_pass(dev, B)
        except RuntimeError as exc:                                         # e.g. out of memory for the dense matrices: say so
            line["torch_gpu_baseline"] = {"unavailable": str(exc)[:200]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# decode workload (BASELINE.json configs[4]): single-/complete-linkage decode of 32-d Poincare embeddings
# ------------------------------------------------------------------------------------------------
def decode_inputs(B, N, seed):
    gen = torch.Generator().manual_seed(seed)
    u = torch.randn(B, N, D_EMB, generator=gen)
    nrm = u.norm(dim=-1, keepdim=True)
    return torch.tanh(nrm.clamp(max=15)) * u / nrm                    # ExpMap(N(0,1)), SURVEY 8(d)


def decode_subrecord(hb, dev, rank, N=1024, reps=5):
    """BASELINE configs[4] inside the default bench line: decode of B clouds x N points (leaves + fp64 cosine matrix +
    linkage, one C-ABI call), for the configuration's real split (64 clouds over 8 GPUs = 8 per GPU) and for 64 per GPU,
    single (north star) and complete (what the reference ships) linkage.  CUDA events around `reps` calls."""
    pk = peaks()
    scale = torch.tensor([SCALE], device=dev)
    rec = {"points": N, "emb_dim": D_EMB, "timing": f"{reps} calls between CUDA events after 2 warm-up calls",
           "roofline": "2 x B x N(N-1)/2 x 8 B (condensed fp64 matrix written once, read once) over the HBM peak"}
    for B in (8, 64):
        x = decode_inputs(B, N, seed=rank).to(dev)
        for method in ("single", "complete"):
            for _ in range(2):
                hb.decode_linkage_batch(x, scale, method)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                hb.decode_linkage_batch(x, scale, method)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = 2.0 * B * N * (N - 1) / 2 * 8 / (ms * 1e-3) / 1e9
            rec[f"{method}_b{B}"] = {"ms": round(ms, 4), "dendrograms_per_s": round(B / (ms * 1e-3), 1),
                                     "achieved_gbs": round(gbs, 1), "frac": round(gbs / pk["hbm_gbs"], 4)}
    return rec


def edgeconv_subrecord(hb, dev, B, rank, reps=5):
    """Row f-1 inside the default bench line: the three graph layers of VN_DGCNN_partseg at the bench shape, fused
    (hpcs_b200.edgeconv: training-mode forward with its BatchNorm statistics passes, and forward+backward wrt the input and
    every parameter) against the unfused composition on the same GPU (the native edge-feature kernel followed by the same
    VN arithmetic as plain PyTorch ops over the [B,2C,3,N,k] tensor, which is what the reference's modules execute)."""
    from hpcs_b200.edgeconv import edgeconv

    class Conv(torch.nn.Module):                         # attribute layout of the reference's VNLinearLeakyReLU
        def __init__(self, cin):
            super().__init__()
            self.negative_slope = 0.2
            self.map_to_feat = torch.nn.Linear(cin, 21, bias=False)
            self.map_to_dir = torch.nn.Linear(cin, 21, bias=False)
            self.batchnorm = torch.nn.Module()
            self.batchnorm.bn = torch.nn.BatchNorm2d(21)

    def vn_torch(e, c):                                  # VNLinearLeakyReLU + VNBatchNorm (training mode) in plain PyTorch
        bn = c.batchnorm.bn
        p = c.map_to_feat(e.transpose(1, -1)).transpose(1, -1)
        d = c.map_to_dir(e.transpose(1, -1)).transpose(1, -1)
        norm = torch.norm(p, dim=2) + 1e-6
        p = p / norm.unsqueeze(2) * torch.nn.functional.batch_norm(norm, None, None, bn.weight, bn.bias, True, 0.1, bn.eps).unsqueeze(2)
        dot = (p * d).sum(2, keepdim=True)
        mask = (dot >= 0).float()
        return 0.2 * p + 0.8 * (mask * p + (1 - mask) * (p - (dot / ((d * d).sum(2, keepdim=True) + 1e-6)) * d))

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    gen = torch.Generator(device=dev).manual_seed(50 + rank)
    torch.manual_seed(50 + rank)
    rec = {"timing": f"{reps} calls between CUDA events after 2 warm-up calls (eager launches, host overhead included)",
           "unfused": "hpcs_b200.get_graph_feature + VNLinearLeakyReLU arithmetic as PyTorch ops (training-mode BatchNorm), fwd+bwd"}
    tot_f = tot_u = 0.0
    for tag, C, two in (("layer1_c1_conv1_conv2", 1, True), ("layer2_c21_conv3_conv4", C_FEAT, True), ("layer3_c21_conv5", C_FEAT, False)):
        convs = [Conv(2 * C).to(dev).train()] + ([Conv(21).to(dev).train()] if two else [])
        params = [p for c in convs for p in c.parameters()]
        x = torch.randn(B, C, 3, N_PTS, device=dev, generator=gen)
        gout = torch.randn(B, 21, 3, N_PTS, device=dev, generator=gen)
        idx = hb.knn(x.view(B, 3 * C, N_PTS), K_NN)

        def fused(bwd=True):
            xr = x.detach().requires_grad_(bwd)
            y = edgeconv(xr, K_NN, convs[0], convs[1] if two else None, idx=idx)
            if bwd:
                torch.autograd.grad((y * gout).sum(), [xr] + params)

        def unfused():
            xr = x.detach().requires_grad_(True)
            e = hb.get_graph_feature(xr, K_NN, idx=idx)
            for c in convs:
                e = vn_torch(e, c)
            torch.autograd.grad((e.mean(dim=-1) * gout).sum(), [xr] + params)
        with torch.no_grad():
            fwd_ms = timed(lambda: fused(False), reps)
        torch.cuda.reset_peak_memory_stats(dev)
        f_ms = timed(fused, reps)
        f_mem = torch.cuda.max_memory_allocated(dev)
        torch.cuda.reset_peak_memory_stats(dev)
        u_ms = timed(unfused, 2)
        u_mem = torch.cuda.max_memory_allocated(dev)
        tot_f, tot_u = tot_f + f_ms, tot_u + u_ms
        rec[tag] = {"fused_fwd_train_ms": round(fwd_ms, 4), "fused_fwd_bwd_ms": round(f_ms, 4), "unfused_fwd_bwd_ms": round(u_ms, 4),
                    "speedup": round(u_ms / f_ms, 2), "fused_peak_mem_mb": round(f_mem / 1e6), "unfused_peak_mem_mb": round(u_mem / 1e6)}
        del convs, params, x, gout, idx
        torch.cuda.empty_cache()
    rec["three_layers_fwd_bwd_ms"] = {"fused": round(tot_f, 4), "unfused": round(tot_u, 4), "speedup": round(tot_u / tot_f, 2)}
    return rec


def model_step_subrecord(dev, B, rank, reps=10, warm=3):
    """SURVEY 8d's "end-to-end clouds/s": one whole training step of the model -- VN_DGCNN_partseg + ExpMap + compute_loss,
    forward, backward and the optimizer step (RAdam, like configure_optimizers) -- through the reference-facing names with the
    binding installed, batch taken from HOST tensors as the data loader yields them.  The reference tree cannot travel to the GPU
    box, so the model is the stand-in tree of tests/fake_reference built at the reference's layer shapes (tail_width 1024 // 3:
    conv8 takes 2299 channels, 1 303 850 backbone parameters, the reference's count); its hot-path entry points raise if they
    are reached unbound.  Graph layers, kNN, loss, sampler-side filter and input pipeline are this library; the dense tail
    (conv6, the std-feature layers, conv7-11), the CosFace term and the optimizer are PyTorch library ops, outside the path.
    Eager launches, host overhead included; CUDA events around `reps` steps after `warm`."""
    if rank != 0:
        return None
    fake = os.path.join(ROOT, "tests", "fake_reference")
    saved = {m: sys.modules.pop(m) for m in list(sys.modules) if m == "hpcs" or m.startswith("hpcs.") or m == "train"}
    sys.path.insert(0, fake)
    from hpcs_b200 import _lib, patch
    rec = None
    try:
        import hpcs.models, hpcs.nn.dgcnn, hpcs.nn.pointnet, hpcs.nn.hyperbolic   # noqa: F401,E401  (import BEFORE install, like train.py)
        patch.uninstall()
        patch.install(strict=True)
        from hpcs.models import ShapeNetHypHC
        from hpcs.nn.dgcnn import VN_DGCNN_partseg
        from hpcs.nn.hyperbolic import ExpMap
        VN_DGCNN_partseg.tail_width = 1024 // 3
        torch.manual_seed(0)
        model = ShapeNetHypHC(nn_feat=VN_DGCNN_partseg(3, D_EMB, K_NN, 0.5, "mean", 16), nn_emb=ExpMap(), euclidean_size=D_EMB,
                              hyp_size=D_EMB, num_class=50, t_per_anchor=T_PER_ANCHOR, fraction=0.0, temperature=TEMPERATURE,
                              miner=True).to(dev).train()
        params = sum(p.numel() for p in model.nn_feat.parameters())
        opt = torch.optim.RAdam(model.parameters(), lr=1e-3)
        host = synth_inputs(B, 1234)
        pts = host["pts"].view(B, 3, N_PTS).transpose(1, 2).contiguous()          # [B,N,3], what the data loader yields
        targets = host["labels"].view(B, N_PTS)
        label = torch.randint(0, 16, (B, 1), generator=torch.Generator().manual_seed(5))
        out = {}

        def step(backward=True):
            losses, _ = model.forward((pts, label, targets), testing=False)
            total = losses["loss_metric"] + losses["loss_hyp"]
            if backward:
                opt.zero_grad(set_to_none=True)
                total.backward()
                opt.step()
            out["loss"] = total.detach()

        def timed(fn):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = _lib.launch_count()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps, (_lib.launch_count() - l0) // reps

        torch.cuda.reset_peak_memory_stats(dev)
        ms_ref_sampler, _ = timed(step)                      # default binding: the reference's host sampler (RNG-exact)
        patch.install(strict=True, sampler="device")         # row f-3: triplets drawn on the GPU, no host sampling / index upload
        ms_step, launches = timed(step)
        with torch.no_grad():
            ms_fwd, _ = timed(lambda: step(False))
        rec = {"ms_per_step": round(ms_step, 3), "clouds_per_s": round(B / (ms_step * 1e-3), 1), "forward_only_ms": round(ms_fwd, 3),
               "ms_per_step_host_sampler": round(ms_ref_sampler, 3), "clouds_per_s_host_sampler": round(B / (ms_ref_sampler * 1e-3), 1),
               "sampler": "device (install(sampler='device')); *_host_sampler = the default binding, the reference's CPU sampler with identical RNG draws",
               "batch": B, "native_launches_per_step": int(launches), "backbone_params": int(params),
               "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2), "loss": float(out["loss"]),
               "what": "VN_DGCNN_partseg + ExpMap + compute_loss, fwd + bwd + RAdam step, host batch in, binding installed "
                       "(hpcs_b200.patch.install(strict=True)); stand-in tree at the reference's layer shapes; eager, host overhead included",
               "optimizer": "RAdam", "timing": f"{reps} steps between CUDA events after {warm} warm-up steps"}
    except Exception as exc:                                  # the sub-record must never take the bench line down
        rec = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    finally:
        try:
            patch.uninstall()
            sys.modules["hpcs.nn.dgcnn"].VN_DGCNN_partseg.tail_width = 32
        except Exception:
            pass
        if fake in sys.path:
            sys.path.remove(fake)
        for m in [m for m in sys.modules if m == "hpcs" or m.startswith("hpcs.")]:
            del sys.modules[m]
        sys.modules.update(saved)
    return rec


def decode_cpu_baseline(x, method, budget_s=12.0):
    """The reference's decoder on host cores: normalize + project (torch CPU) then scipy linkage, cloud by cloud
    (base_hyp_hc.py:81-86,135-137), on as many clouds of the same batch as fit the time budget."""
    from oracle import hpcs_oracle as O
    done, t0 = 0, time.perf_counter()
    while done < x.shape[0] and (done == 0 or time.perf_counter() - t0 < budget_s):
        O.decode_linkage(x[done], torch.tensor([SCALE]), method)
        done += 1
    dt = time.perf_counter() - t0
    return {"value": round(done / dt, 3), "unit": "dendrograms/s", "cores": 1, "kind": "port",
            "sample": f"{done} clouds x {x.shape[1]} pts, oracle restatement of _decode_linkage (torch CPU normalize/project + "
                      f"scipy linkage(method='{method}', metric='cosine'), single-threaded like the reference)",
            "s_per_cloud": round(dt / done, 4)}


def run_decode(args):
    import hpcs_b200 as hb
    from hpcs_b200 import _lib, dist as hdist
    import torch.distributed as dist
    rank, world, local = hdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, method = args.decode_b, args.decode_n, args.method
    host = decode_inputs(B, N, seed=rank)
    x = host.to(dev)
    scale = torch.tensor([SCALE], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return hb.decode_linkage_batch(x, scale, method)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        Z = step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    Z = step()
    launches_per_step = _lib.launch_count() - l0
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        Z = step()
    t1.record()
    if sampler._nv is not None:
        sampler.sample()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps

    # end to end: embeddings from pinned host memory, Z back to pinned host memory (what fcluster consumes); two buffer
    # sets on three streams, so the upload of step i+1 and the download of step i-1 overlap the decode of step i
    sampler.region = "e2e"
    x_pin = host.pin_memory()
    z_pin = [torch.empty(Z.shape, dtype=Z.dtype).pin_memory() for _ in range(2)]
    xd = [torch.empty_like(x) for _ in range(2)]
    zd = [None, None]
    s_up, s_run, s_down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_run = [torch.cuda.Event() for _ in range(2)]
    ev_down = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n_steps):
        for i in range(n_steps):
            j = i % 2
            with torch.cuda.stream(s_up):
                s_up.wait_event(ev_run[j])                      # the previous decode of this set has read its input
                xd[j].copy_(x_pin, non_blocking=True)
                ev_up[j].record(s_up)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_up[j])
                s_run.wait_event(ev_down[j])                    # its previous result has left the device
                zd[j] = hb.decode_linkage_batch(xd[j], scale, method)
                ev_run[j].record(s_run)
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev_run[j])
                z_pin[j].copy_(zd[j], non_blocking=True)
                zd[j].record_stream(s_down)
                ev_down[j].record(s_down)

    e2e_loop(4)
    barrier()
    te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te0.record()
    e2e_loop(args.steps)
    for s_ in (s_up, s_run, s_down):
        torch.cuda.current_stream().wait_stream(s_)
    te1.record()
    barrier()
    clocks = sampler.stop()
    ms_e = torch.tensor([te0.elapsed_time(te1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    pk = peaks()
    alg_bytes = 2.0 * B * N * (N - 1) / 2 * 8                          # fp64 condensed matrix written once, read once
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    zc = z_pin[(args.steps - 1) % 2].numpy()
    line = {
        "metric": f"dendrograms/sec ({method}-linkage decode, {N} pts, 32-d)", "value": round(world * B / (ms_per_step * 1e-3), 1),
        "unit": "dendrograms/s", "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"decode_{method}_b{B}_n{N}_d{D_EMB} (BASELINE.json configs[4])", "clouds_per_gpu": B,
                   "points": N, "emb_dim": D_EMB, "method": method, "scale": SCALE, "parallelism": f"dp{world}",
                   "l2": f"fp64 distance matrices {B * N * N * 8 / 1e6:.0f} MB per step vs 126 MB L2; no explicit flush",
                   "launch": "eager (one C-ABI call per step)"},
        "e2e": {"value": round(world * B / (ms_e.item() / args.steps * 1e-3), 1), "unit": "dendrograms/s",
                "h2d_bytes_per_step": int(x_pin.numel() * 4), "d2h_bytes_per_step": int(z_pin[0].numel() * 8)},
        "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks,
        "roofline": {"kernel": "pdist_cosine + linkage", "bound": "hbm", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": None,
                     "algorithmic_bytes": alg_bytes, "peak_source": pk["source"]},
        "root_count": float(zc[0, -1, 3]),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = decode_cpu_baseline(host, method)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's restatement of the reference's PyTorch path on host cores
# ------------------------------------------------------------------------------------------------
CPU_SAMPLE_B = B_PER_GPU            # the native arm's batch: like for like (22 s per step on 8 cores, 10 GB)


def cpu_reference_pass(steps, warmup, B=None):
    from oracle import hpcs_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = CPU_SAMPLE_B if B is None else B
    host = synth_inputs(B, seed=0)
    torch.manual_seed(1000)
    trip = O.sample_triplets(host["labels"], T_PER_ANCHOR, 0.0)
    gen = torch.Generator().manual_seed(0)
    gs = [torch.randn(B, 2 * c, 3, N_PTS, K_NN, generator=gen) for c in (1, C_FEAT, C_FEAT)]
    scale = torch.tensor([SCALE], requires_grad=True)

    def one():
        for x, g in zip((host["pts"], host["f1"], host["f2"]), gs):
            xr = x.clone().requires_grad_(True)
            y = O.graph_feature(xr, K_NN)                           # reference knn + gather/cat/permute
            torch.autograd.grad(y, xr, g)
        emb = host["emb"].clone().requires_grad_(True)
        a, p, n = O.filter_triplets(emb.detach(), *trip)             # dense [n,n] similarity, like the miner
        loss = O.compute_hyp(emb, a, p, n, scale, TEMPERATURE)       # dense [n,n] again + 3 hyp_lca
        torch.autograd.grad(loss, (emb, scale))
        return float(loss.detach())

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(B / dt, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{B} clouds x {N_PTS} pts ({B * N_PTS * T_PER_ANCHOR} mined triplets) per step, "
                      f"oracle restatement of the reference PyTorch path, torch CPU threads={cores}",
            "s_per_step": round(dt, 3)}


def torch_gpu_reference_pass(dev, B, steps=2, warmup=1):
    """The reference's PyTorch composition of the path (the oracle's restatement of it: topk kNN over the [B,N,N] matrix,
    gather / cat / permute edge features, dense [n,n] similarity matrices in the miner's filter and in compute_hyp, three
    hyp_lca chains) executed ON THE SAME GPU, same batch, fp32, autograd backward.  A reported baseline next to the CPU one:
    it answers 'what would the reference's own code do on this B200' -- it is still the port, not the unmodified repo."""
    from oracle import hpcs_oracle as O
    host = synth_inputs(B, seed=0)
    torch.manual_seed(1000)
    trip = tuple(t.to(dev) for t in O.sample_triplets(host["labels"], T_PER_ANCHOR, 0.0))
    xs = [host[k].to(dev) for k in ("pts", "f1", "f2")]
    emb0 = host["emb"].to(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    gs = [torch.randn(B, 2 * c, 3, N_PTS, K_NN, device=dev, generator=gen) for c in (1, C_FEAT, C_FEAT)]
    scale = torch.tensor([SCALE], device=dev, requires_grad=True)

    def one():
        for x, g in zip(xs, gs):
            xr = x.clone().requires_grad_(True)
            torch.autograd.grad(O.graph_feature(xr, K_NN), xr, g)
        emb = emb0.clone().requires_grad_(True)
        a, p, n = O.filter_triplets(emb.detach(), *trip)
        torch.autograd.grad(O.compute_hyp(emb, a, p, n, scale, TEMPERATURE), (emb, scale))

    for _ in range(warmup):
        one()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats(dev)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    out = {"value": round(B / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 2), "kind": "port", "device": "cuda (same B200)",
           "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1),
           "sample": f"{B} clouds x {N_PTS} pts, oracle restatement of the reference's PyTorch ops run on the GPU (torch {torch.__version__}), {steps} steps"}
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, min(args.steps, 2)), max(0, min(args.warmup, 1))   # ~12-25 s per step on the host cores
    cb = cpu_reference_pass(steps, warmup, B=args.ref_clouds)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(cb["s_per_step"] * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "shapenet_train_step_b32_n1024_k20_d32_t50 (BASELINE.json configs[1])",
                   "clouds_per_step": args.ref_clouds, "same_batch_as_native_arm": args.ref_clouds == B_PER_GPU, "points": N_PTS, "k": K_NN, "emb_dim": D_EMB,
                   "note": "reference is pure Python/PyTorch with un-installable dependencies; this arm times the oracle "
                           "restatement of its CPU path on the SAME batch as the native arm (32 clouds: the dense "
                           "32768 x 32768 similarity matrix is built twice forward and once backward, as the reference does)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batches-per-step", type=int, default=16,
                    help="train workload: batches per timed step (16 x 20 steps x 0.78 ms = a 250 ms timed region)")
    ap.add_argument("--ref-clouds", type=int, default=CPU_SAMPLE_B,
                    help="reference arm: clouds per step (default 32 = the native arm's batch; smaller only for quick checks)")
    ap.add_argument("--no-decode", action="store_true", help="skip the sub-records of the train line (configs[4] decode, fused EdgeConv layers)")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", choices=["train", "decode"], default="train",
                    help="train = the headline step (configs[1]); decode = dendrogram decode (configs[4])")
    ap.add_argument("--decode-n", type=int, default=1024)
    ap.add_argument("--decode-b", type=int, default=64)
    ap.add_argument("--method", choices=["single", "complete"], default="single")
    args = ap.parse_args()
    if args.impl == "reference" and args.workload == "decode":
        if int(os.environ.get("RANK", "0")) == 0:
            cb = decode_cpu_baseline(decode_inputs(args.decode_b, args.decode_n, 0), args.method)
            print(json.dumps({"impl": "reference", "metric": f"dendrograms/sec ({args.method}-linkage decode, {args.decode_n} pts, 32-d)",
                              "value": cb["value"], "unit": cb["unit"], "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
                              "higher_is_better": True, "cpu_baseline": cb,
                              "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "decode":
        run_decode(args)
    else:
        run_native(args)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
