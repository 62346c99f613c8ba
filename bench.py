#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline: ResNet-50 int8-QAT images/s through the fake-quant hot path.

A "step" is one pass of the hot path over one batch: for every quantization node of symbol/resnet_int8.py at
batch 256 (54 activation + 54 weight nodes, 2.730 G + 25.5 M float32 elements) the
operator's forward (reduction + EMA threshold update + QDQ sweep) and then its backward (straight-through copy),
called through the reference-facing CustomOp protocol -> ctypes -> libb2q.so.  Convolutions are library code
and are not part of the path (nor of the reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (QDQ sweep): algorithmic bytes / CUDA-event time of its launches in a timed region
  kernels       the same for every kernel kind
  e2e           same step with HOST (pinned) buffers through the host-buffer C ABI: H2D + kernels + D2H per node
  cpu_baseline  oracle/c (C/OpenMP restatement of the reference's MXNet CPU op chain) on a bounded sample
  clocks        SM clock / throttle reasons sampled with NVML during the timed region
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "resnet50_int8_qat_quant_path_images_per_sec"
UNIT = "img/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="resnet50_int8")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's, 256)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--full-model-multi", action="store_true",
                    help="also run the whole-network context leg data parallel when launched on several GPUs")
    ap.add_argument("--no-full-model", action="store_true",
                    help="skip the context leg that trains the whole ResNet-50 (library convolutions + these operators)")
    ap.add_argument("--profile", action="store_true",
                    help="profiling aid: only the per-node eager step (warm-up + timed), nothing else")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks (NVML) sampled during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.05):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def make_op(op_type, is_w):
    """Operator through its registered Prop with string attributes, as MXNet would create it."""
    import b200quant
    if op_type == "GDRQ_PY":   # config/quant_attrs.py:38-63 with nbits 8 (int8)
        attrs = dict(nbits="8", group_size="-1", is_weight=str(is_w), lamda="0.001", delay_quant="0",
                     fix_alpha="False", ktimes="3")
    else:
        attrs = dict(quant_mode="minmax", is_weight=str(is_w), is_weight_perchannel="False", delay_quant="0",
                     ema_decay="0.99")
    return b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None)


def build_nodes(torch, nodes, op_type, device, host=False, seed=5):
    """Allocate every node's tensors (inputs resident before the timed region) and create its operator."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for name, kind, shape in nodes:
        is_w = kind == "weight"
        op = make_op(op_type, is_w)
        if is_w:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            x = torch.empty(shape, device=device).normal_(0.0, (2.0 / fan_in) ** 0.5, generator=g)
        else:
            x = torch.empty(shape, device=device).uniform_(-1.0, 1.0, generator=g)   # data/imagenet.py:16
        dy = torch.empty(shape, device=device).normal_(generator=g)
        out.append(dict(name=name, kind=kind, shape=shape, op=op, x=x, y=torch.empty_like(x), dy=dy,
                        dx=torch.empty_like(x), aux=torch.ones(1, device=device), n=x.numel()))
    return out


def run_step(nodes):
    for nd in nodes:                                   # forward, network order
        nd["op"].forward(True, ["write"], [nd["x"]], [nd["y"]], [nd["aux"]])
    for nd in reversed(nodes):                         # backward
        nd["op"].backward(["write"], [nd["dy"]], [nd["x"]], [nd["y"]], [nd["dx"]], [nd["aux"]])


def time_steps(torch, dist, fn, steps, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def host_step_factory(torch, nodes, ctx):
    """e2e: the same step with HOST buffers.  One pinned (x, y, dx) triple per distinct shape and staging set;
    every node copies its full input H2D and its full result D2H inside the timed region."""
    import b200quant
    pools = {}
    hnodes = []
    for i, nd in enumerate(nodes):
        key = (nd["shape"], i & 1)
        if key not in pools:
            pools[key] = tuple(torch.empty(nd["shape"], dtype=torch.float32).pin_memory() for _ in range(3))
            pools[key][0].copy_(nd["x"])
        hx, hy, hdx = pools[key]
        prop = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight=str(nd["kind"] == "weight"),
                                                          is_weight_perchannel="False")
        hnodes.append(dict(op=prop.create_operator(None, None, None), x=hx, y=hy, dx=hdx,
                           aux=torch.ones(1).pin_memory(), n=nd["n"]))

    def step():
        for h in hnodes:
            h["op"].forward(True, ["write"], [h["x"]], [h["y"]], [h["aux"]])
        for h in reversed(hnodes):
            h["op"].backward(["write"], [h["y"]], [h["x"]], [h["y"]], [h["dx"]], [h["aux"]])
        ctx.host_sync()

    bytes_in = sum(4 * h["n"] for h in hnodes) * 2       # x (forward) + dy (backward)
    bytes_out = sum(4 * h["n"] for h in hnodes) * 2      # y + dx
    return step, bytes_in, bytes_out


# ---------------------------------------------------------------------------------------------------------
# reference CPU path (oracle/c): the only place bench.py executes oracle code, and only as the timed baseline
# ---------------------------------------------------------------------------------------------------------
CPU_SAMPLE = [("act", (256, 64, 56, 56)), ("act", (256, 256, 14, 14)), ("act", (256, 2048)),
              ("weight", (512, 512, 3, 3)), ("weight", (64, 3, 7, 7))]


def full_model_leg(torch, device, batch, steps=8, warmup=3, world=1, rank=0):
    """Context only (not the metric): one SGD step of the whole ResNet-50 int8-QAT network of symbol/resnet_int8.py --
    cuDNN/cuBLAS convolutions, BatchNorm and pooling from torch, every conv/FC input and weight through this package's
    Quantization_int8_V2 nodes -- to show what share of a training step the quantization path is.  With world > 1
    (--full-model-multi) it is the data-parallel training the reference runs through Module + KVStore (train.py:34-35):
    activation thresholds through the fused peer-memory exchange, gradients through DistributedDataParallel."""
    from b200quant.harness import ResNetInt8, quant_nodes
    torch.manual_seed(11)
    model = ResNetInt8().to(device)
    net, ex = model, None
    if world > 1:
        from b200quant.dist import attach_peer_exchange
        ex = attach_peer_exchange([m.op for m in quant_nodes(model)], device)
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], broadcast_buffers=False,
                                                        gradient_as_bucket_view=True)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator(device=device).manual_seed(100 + rank)
    x = torch.randn(batch, 3, 224, 224, device=device, generator=g)
    y = torch.randint(0, 1000, (batch,), device=device, generator=g)

    def train_step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        train_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = train_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # every rank must hold the same thresholds after a data-parallel step
        aux = torch.cat([a.flatten() for m in quant_nodes(model) for a in m.aux_list()])
        lo, hi = aux.clone(), aux.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        torch.cuda.synchronize()
        dist.barrier()
        ex.close()
    else:
        same = None
    return {"images_per_sec": world * batch / (ms / 1e3), "ms_per_step": ms, "batch": batch, "n_gpus": world,
            "steps": steps, "thresholds_identical_across_ranks": same,
            "loss": float(loss.detach()), "conv_math": "torch/cuDNN fp32 (TF32 %s)" % ("on" if torch.backends.cudnn.allow_tf32 else "off"),
            "note": "whole ResNet-50 int8-QAT SGD step, eager torch autograd; convolutions/BN are library code, "
                    "quantization nodes are this package's (108 nodes); context for the metric, not the metric"}


def cpu_reference_step(sample_state, op_type="Quantization_int8_V2"):
    """One forward + backward of the sample tensors through the C restatement of the workload's operator."""
    from oracle import c_oracle as co
    if op_type == "GDRQ_PY":   # the attributes of make_op: nbits 8 -> 255 levels, ktimes 3, lamda 0.001
        for s in sample_state:
            co.gdrq_fwd(s["x"], s["y"], s["aux"], s["w"], False, True, 255.0, 3.0, 0.001)
        for s in reversed(sample_state):
            if s["w"]:
                co.ste_bwd(s["dy"], s["dx"])
            else:
                co.gdrq_bwd(s["x"], s["dy"], s["dx"], s["aux"])
        return
    variant = 1 if op_type == "ClipGrad_Quantization_int8" else 0
    for s in sample_state:
        co.minmax_quant_fwd(variant, s["x"], s["y"], s["aux"], s["w"], False, True, False, 0.99)
    for s in reversed(sample_state):
        if variant == 1 and not s["w"]:
            co.clipgrad_bwd(s["x"], s["dy"], s["dx"], s["aux"])
        else:
            co.ste_bwd(s["dy"], s["dx"])


def cpu_reference_prepare():
    import numpy as np
    from b200quant.workloads import numel
    rng = np.random.default_rng(5)
    st = []
    for kind, shape in CPU_SAMPLE:
        x = rng.uniform(-1, 1, shape).astype(np.float32) if kind == "act" else \
            (rng.standard_normal(shape) * (2.0 / numel(shape[1:])) ** 0.5).astype(np.float32)
        st.append(dict(x=x, y=np.empty_like(x), dy=rng.standard_normal(shape).astype(np.float32),
                       dx=np.empty_like(x), aux=np.ones(1, np.float32), w=(kind == "weight")))
    return st, sum(numel(s) for _, s in CPU_SAMPLE)


def cpu_baseline(total_elems, batch, steps=2, warmup=1, op_type="Quantization_int8_V2"):
    from oracle import c_oracle as co
    co.use_all_host_threads()   # whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    st, sample_elems = cpu_reference_prepare()
    for _ in range(warmup):
        cpu_reference_step(st, op_type)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(st, op_type)
    dt = (time.perf_counter() - t0) / steps
    full = dt * total_elems / sample_elems
    return dict(value=batch / full, unit=UNIT, cores=co.num_threads(), kind="port",
                sample="fwd+bwd of %s (%d of the step's %d elements) per step, time scaled by element count; "
                       "oracle/c = C/OpenMP restatement of the reference's mx.nd op chain, all host threads"
                       % (", ".join("x".join(map(str, s)) for _, s in CPU_SAMPLE), sample_elems, total_elems),
                sample_seconds=dt, ms_per_step_extrapolated=full * 1e3)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from b200quant.workloads import WORKLOADS, summary
    fn, batch, op_type = WORKLOADS[args.workload]
    if op_type == "GDRQ_Fold_BN":
        raise SystemExit("the fold-BN workload is a parity case (tests/test_gpu_configs.py), not a bench line")
    batch = args.batch or batch
    sm = summary(fn(batch))
    total = sm["act_elems"] + sm["weight_elems"]
    from oracle import c_oracle as co
    co.use_all_host_threads()   # the reference arm gets every host thread, also under torchrun (OMP_NUM_THREADS=1)
    st, sample_elems = cpu_reference_prepare()
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_step(st, op_type)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(st, op_type)
    dt = (time.perf_counter() - t0) / args.steps
    full = dt * total / sample_elems
    value = batch / full
    sample = ("each step = fwd+bwd of %s (%d of %d elements), time scaled by element count"
              % (", ".join("x".join(map(str, s)) for _, s in CPU_SAMPLE), sample_elems, total))
    line = {"impl": "reference",
            "metric": METRIC if args.workload == "resnet50_int8" else args.workload + "_quant_path_images_per_sec", "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s quant path (fwd+bwd of all %d nodes), batch %d" %
                                   (args.workload, sm["act_nodes"] + sm["weight_nodes"], batch),
                       "reference": "oracle/c C+OpenMP restatement of the MXNet CPU CustomOp chain "
                                    "(libmxnet is not installable here)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": co.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


_REAL_STDOUT = None


def claim_stdout():
    """Libraries (NCCL's version banner, torchrun notices) print to fd 1; the driver wants exactly one JSON line
    there.  Everything written to stdout from now on goes to stderr; emit() writes the JSON to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        return main_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the quantization operators have no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        # a short collective timeout: if one rank dies, the others abort instead of waiting 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=int(os.environ.get("B2Q_NCCL_TIMEOUT", "180"))))
    device = torch.device("cuda", local)

    import b200quant  # noqa: F401
    from b200quant import _lib
    from b200quant.dist import GradBucket, attach_threshold_sync
    from b200quant.workloads import WORKLOADS, summary

    fn, batch, op_type = WORKLOADS[args.workload]
    batch = args.batch or batch
    full_model = None
    if (world == 1 or args.full_model_multi) and not args.no_full_model and not args.profile \
            and args.workload == "resnet50_int8":
        try:   # context leg, run first while the device memory is still free
            full_model = full_model_leg(torch, device, batch, world=world, rank=rank)
        except Exception as e:  # pragma: no cover
            full_model = {"images_per_sec": None, "error": str(e).splitlines()[0][:200]}
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    node_list = fn(batch)
    sm = summary(node_list)
    total_elems = sm["act_elems"] + sm["weight_elems"]
    ctx = _lib.context(local)
    for k, v in os.environ.items():   # tuning experiments: B2Q_OPT_<option>=<int> (results never depend on them)
        if k.startswith("B2Q_OPT_"):
            ctx.set_option(k[len("B2Q_OPT_"):].lower(), int(v))
    if op_type == "GDRQ_Fold_BN":
        raise SystemExit("the fold-BN workload is a parity case (tests/test_gpu_configs.py), not a bench line")
    nodes = build_nodes(torch, node_list, op_type, device, seed=5 + rank)
    minmax = op_type in ("Quantization_int8_V2", "ClipGrad_Quantization_int8")
    # algorithmic bytes per element: forward 12 (4 reduce + 8 sweep); backward 8 (STE copy) or 12 (masked, activations
    # of the clip-grad / GDRQ operators)
    bwd_act = 8 if op_type == "Quantization_int8_V2" else 12
    alg_bytes_step = 12 * total_elems + bwd_act * sm["act_elems"] + 8 * sm["weight_elems"]

    bucket = None
    exchange = "none"
    if world == 1 and os.environ.get("B2Q_DEBUG_PEER_WORLD1") == "1":   # experiment: mailbox kernels without a peer
        from b200quant.dist import attach_peer_exchange
        attach_peer_exchange([nd["op"] for nd in nodes], device)
        exchange = "fused peer-memory kernels at world 1 [DEBUG experiment]"
    if world > 1:   # data parallel: thresholds max over ranks per activation node, weight grads allreduce(sum)
        exchange = "nccl allreduce(max) per activation node"
        attach_threshold_sync([nd["op"] for nd in nodes])
        if os.environ.get("B2Q_EXCHANGE", "peer") == "peer":
            try:   # fused peer-memory exchange (NVLink): no NCCL call on the forward critical path
                from b200quant.dist import attach_peer_exchange
                attach_peer_exchange([nd["op"] for nd in nodes], device)
                exchange = "fused peer-memory kernels (b2q_peer_%s_quant_fwd_f32)" % ("minmax" if minmax else "meanabs",)
            except Exception as e:  # pragma: no cover
                exchange += " (peer path unavailable: %s)" % (str(e).splitlines()[0][:100],)
        # operators without a fused exchange (the mean-based GDRQ thresholds) keep the NCCL call on their forward path
        nccl_forward = any(getattr(nd["op"], "sync", None) is not None for nd in nodes)
        if nccl_forward and exchange.startswith("fused peer"):
            exchange = "nccl allreduce(max) of the per-node statistic (no fused exchange for this operator)"
        wn = [nd for nd in nodes if nd["kind"] == "weight"]
        bucket = GradBucket([nd["shape"] for nd in wn], device)
        for nd, view in zip(wn, bucket.views):
            nd["dx"] = view
        if os.environ.get("B2Q_DEBUG_SKIP_GRAD_ALLREDUCE") == "1":   # experiment only: isolates the exchange cost
            bucket.allreduce = lambda *a, **k: None
            exchange += " [DEBUG: gradient allreduce skipped -- not a valid result]"

    def step_local():
        run_step(nodes)

    def step():
        step_local()
        if bucket is not None:
            bucket.allreduce()

    # same semantics, weight nodes batched: all 54 weight tensors in 2 launches forward / 1 launch backward
    from b200quant.multi import WeightGroup
    wnodes = [nd for nd in nodes if nd["kind"] == "weight"]
    anodes = [nd for nd in nodes if nd["kind"] == "act"]
    group = None
    if minmax or op_type == "GDRQ_PY":
        group = WeightGroup([nd["op"] for nd in wnodes], [nd["x"] for nd in wnodes], [nd["y"] for nd in wnodes],
                            [nd["aux"] for nd in wnodes], [nd["dy"] for nd in wnodes], [nd["dx"] for nd in wnodes])

    # The weight gradients are complete after group.backward(); their allreduce(sum) then runs on a side stream while
    # the activation backward sweeps continue (standard overlap of gradient communication with the backward pass).
    side = torch.cuda.Stream() if bucket is not None else None

    def part1():   # forward of everything + weight backward
        group.forward(True)
        for nd in anodes:
            nd["op"].forward(True, ["write"], [nd["x"]], [nd["y"]], [nd["aux"]])
        group.backward()

    def part2():   # activation backward
        for nd in reversed(anodes):
            nd["op"].backward(["write"], [nd["dy"]], [nd["x"]], [nd["y"]], [nd["dx"]], [nd["aux"]])

    def overlapped(p1, p2):
        p1()
        if bucket is not None:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                bucket.allreduce()
        p2()
        if bucket is not None:
            torch.cuda.current_stream().wait_stream(side)

    def step_multi_local():
        part1()
        part2()

    def step_multi():
        overlapped(part1, part2)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- main timed region (eager: every call goes through the CustomOp protocol) ----
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms_eager = time_steps(torch, dist, step, args.steps, world)
    launches = ctx.launch_count() - l0
    clocks = sampler.finish()
    mode = "eager, one CustomOp call per node"
    ms = ms_eager
    timings = {mode: ms_eager / args.steps}

    if args.profile:
        group = None
        args.no_graph = args.no_e2e = args.no_cpu = True
    if group is not None:
        for _ in range(3):
            step_multi()
        l0 = ctx.launch_count()
        ms_multi = time_steps(torch, dist, step_multi, args.steps, world)
        timings["eager, weight nodes batched (WeightGroup)"] = ms_multi / args.steps
        if ms_multi < ms:
            ms, mode, launches = ms_multi, "eager, weight nodes batched (WeightGroup)", ctx.launch_count() - l0

    # ---- the same steps replayed from a CUDA graph (single GPU; the graph holds our kernels only) ----
    ms_graph = None
    # multi-GPU: only our own kernels are captured (the peer-memory exchange needs no NCCL call); the single gradient
    # allreduce stays an eager NCCL call after each replay.  With the NCCL threshold exchange nothing is captured.
    graph_ok = world == 1 or (exchange.startswith("fused peer") and os.environ.get("B2Q_GRAPH_MULTI", "1") == "1")
    if graph_ok and not args.no_graph:
        for label, fn_ in (("cuda_graph, one CustomOp call per node", step_local),
                           ("cuda_graph, weight nodes batched (WeightGroup)", step_multi_local)):
            if fn_ is step_multi_local and group is None:
                continue
            parts = [fn_] if (bucket is None or fn_ is step_local) else [part1, part2]
            try:
                graph, err = None, None
                try:
                    graphs = []
                    captured = 0
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        for part in parts:
                            part()
                        torch.cuda.synchronize()
                        for part in parts:
                            g_ = torch.cuda.CUDAGraph()
                            l0 = ctx.launch_count()
                            with torch.cuda.graph(g_, stream=s):
                                part()
                            captured += ctx.launch_count() - l0
                            graphs.append(g_)
                    torch.cuda.current_stream().wait_stream(s)
                    graph = graphs
                except Exception as e:  # pragma: no cover
                    graph, err = None, e
                if world > 1:   # every rank must take the same path, or the peer exchange would wait forever
                    okflag = torch.tensor([0 if graph is None else 1], device="cuda")
                    dist.all_reduce(okflag, op=dist.ReduceOp.MIN)
                    if int(okflag.item()) == 0 and graph is not None:
                        graph, err = None, RuntimeError("capture failed on another rank")
                if graph is None:
                    raise err

                def replay(graphs=graph):
                    if len(graphs) == 2:
                        overlapped(graphs[0].replay, graphs[1].replay)
                    else:
                        graphs[0].replay()
                        if bucket is not None:
                            bucket.allreduce()

                for _ in range(3):
                    replay()
                t = time_steps(torch, dist, replay, args.steps, world)
                timings[label] = t / args.steps
                if label.endswith("per node"):
                    ms_graph = t
                if t < ms:
                    ms, mode, launches = t, label, captured * args.steps
            except Exception as e:  # pragma: no cover
                timings[label] = "failed: %s" % (str(e).splitlines()[0][:120],)

    value = world * batch * args.steps / (ms / 1e3)

    # ---- per-kernel timing for the roofline (second timed region, events around every flat-kernel launch) ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    ctx.set_option("timing", 1)
    ctx.timing_read(0, reset=True)
    for _ in range(0 if args.profile else min(args.steps, 5)):
        step()
    torch.cuda.synchronize()
    kinds = {1: "reduce_flat (max|x| + EMA update)", 2: "qdq_flat_hot (QDQ sweep)", 3: "bwd_flat (STE copy)",
             4: "bwd_flat (clip mask)", 5: "segmented/other"}
    kernels = {}
    LARGE = 64e6   # algorithmic bytes: the tensors the ">= 70 % of HBM peak on large tensors" target is about
    for k, name in kinds.items():
        kms, kbytes, kn = ctx.timing_read(k)
        if kn:
            kernels[name] = {"launches": kn, "ms_total": kms, "alg_bytes_total": kbytes,
                             "achieved_gbs": kbytes / kms / 1e6, "frac_of_peak": kbytes / kms / 1e6 / peak_gbs}
            lms, lbytes, ln = ctx.timing_read(k, min_bytes=LARGE)
            if ln:
                kernels[name]["large_tensors"] = {"min_alg_bytes": LARGE, "launches": ln, "ms_total": lms,
                                                  "alg_bytes_total": lbytes, "achieved_gbs": lbytes / lms / 1e6,
                                                  "frac_of_peak": lbytes / lms / 1e6 / peak_gbs}
    ctx.timing_read(0, reset=True)
    ctx.set_option("timing", 0)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = {"source": tj.get("source"), "launch_alg_bytes": tj.get("alg_bytes_of_this_launch"),
                        "launch_duration_us": tj.get("duration_us"),
                        "note": "ncu dram__bytes_read+write of one captured launch (a 256x64x56x56 activation); compare "
                                "with launch_alg_bytes, not with the average alg_bytes_per_launch"}
    except Exception:
        pass
    dom = kernels.get(kinds[2], {})
    roofline = {"bound": "hbm", "kernel": "qdq_flat_hot_kernel", "achieved": dom.get("achieved_gbs"),
                "peak": peak_gbs, "unit": "GB/s", "frac": dom.get("frac_of_peak"), "traffic": traffic,
                "peak_source": peak_src, "traffic_detail": traffic_note,
                "alg_bytes_per_launch": (dom.get("alg_bytes_total", 0) / dom["launches"]) if dom else None,
                "avg_launch_ms": (dom.get("ms_total", 0) / dom["launches"]) if dom else None}

    metric = METRIC if args.workload == "resnet50_int8" else args.workload + "_quant_path_images_per_sec"
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s quant path: fwd+bwd of all %d %s nodes (%d act + %d weight), "
                                   "per-GPU batch %d" % (args.workload, len(nodes), op_type, sm["act_nodes"],
                                                         sm["weight_nodes"], batch),
                       "elements_per_step": total_elems, "alg_bytes_per_step": alg_bytes_step,
                       "parallelism": "dp%d" % world, "threshold_exchange": exchange, "l2": "inputs larger than L2 (10.9 GB touched once per step)",
                       "mode": mode},
            "ms_per_step_by_mode": timings,
            "hbm_frac_whole_step": alg_bytes_step / (ms / args.steps / 1e3) / 1e9 / peak_gbs,
            "roofline": roofline, "kernels": kernels,
            "kernels_note": "event-timed per launch in a separate pass, all tensor sizes pooled (54 of the 108 launches "
                            "per kind are weight tensors of a few KB..MB that cost a launch each); a kernel that follows "
                            "a sweep also pays for the write-back of the output lines its predecessor left dirty in L2, "
                            "so reduce_flat reads low and qdq_flat_hot high; hbm_frac_whole_step is the unbiased figure",
            "clocks": clocks, "gpu_launches": launches}

    # ---- e2e: host buffers through the host C ABI (rank-local; N ranks run it concurrently) ----
    if not args.no_e2e and op_type == "Quantization_int8_V2":
        try:
            hstep, b_in, b_out = host_step_factory(torch, nodes, ctx)
            hstep()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                hstep()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            line["e2e"] = {"value": world * batch * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": b_in,
                           "d2h_bytes_per_step": b_out, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
                           "path": "CustomOp.forward/backward with pinned HOST tensors -> b2q_*_host_f32 (three-stream "
                                   "pipeline over a staging ring; PCIe-bound: ~46 GB/s per direction on this box)"}
        except Exception as e:  # pragma: no cover - e.g. not enough lockable host memory
            line["e2e"] = {"value": None, "unit": UNIT, "error": str(e).splitlines()[0][:200]}

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline(total_elems, batch, op_type=op_type)
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": str(e).splitlines()[0][:200]}

    if full_model is not None:
        if full_model.get("ms_per_step"):
            full_model["quant_path_share_of_step"] = (ms / args.steps) / full_model["ms_per_step"]
        line["full_model"] = full_model

    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
