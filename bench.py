#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline: ResNet-50 int8-QAT images/s through the fake-quant hot path.

A "step" is one pass of the hot path over one batch: for every quantization node of symbol/resnet_int8.py at
batch 256 (54 activation + 54 weight nodes, 2.730 G + 25.5 M float32 elements) the
operator's forward (reduction + EMA threshold update + QDQ sweep) and then its backward (straight-through copy),
called through the reference-facing CustomOp protocol -> ctypes -> libb2q.so.  Convolutions are library code
and are not part of the path (nor of the reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One JSON line on stdout (rank 0).  `value` is the DROP-IN mode -- eager, one CustomOp call per node, which is what a
CustomOp framework gets; the CUDA-graph / multi-tensor modes are reported beside it (`ms_per_step_by_mode`,
`value_best`).  Keys beyond the base contract:
  roofline      dominant kernel (QDQ sweep): algorithmic bytes / CUDA-event time of its launches in a timed region
  kernels       the same for every kernel kind
  e2e           same step with HOST (pinned) buffers through the host-buffer C ABI: H2D + kernels + D2H per node
  cpu_baseline  oracle/c (C/OpenMP restatement of the reference's MXNet CPU op chain), the full step on the host
  workloads     compact results for the other BASELINE.json configs (MobileNet GDRQ, ResNeXt-101 clip-grad,
                MobileNet fold-BN quant path)
  full_model    whole-network ResNet-50 int8-QAT SGD step (library convolutions + these operators), every N
  parity_checked  outputs of a sampled activation node bit-compared with the oracle fed the max over ranks, and
                aux identical on all ranks, after the timed region
  clocks        SM clock / throttle reasons sampled with NVML during the timed region
`--impl reference` runs the same step (every node, same config) through oracle/c on the host cores.
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
DROPIN_MODE = "eager, one CustomOp call per node"


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
    ap.add_argument("--no-workloads", action="store_true", help="skip the compact legs for the other configs")
    ap.add_argument("--no-micro", action="store_true", help="skip the second-tier / weight-kernel micro-benchmarks")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-full-model", action="store_true",
                    help="skip the leg that trains the whole ResNet-50 (library convolutions + these operators)")
    ap.add_argument("--profile", action="store_true",
                    help="profiling aid: only the per-node eager step (warm-up + timed), nothing else")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# workload description shared by BOTH arms (the driver compares the two config dicts)
# ---------------------------------------------------------------------------------------------------------
def alg_bytes_per_step(op_type, sm):
    """forward 12 B/elem (4 reduce + 8 sweep); backward 8 (STE copy) or 12 (masked: activations of the clip-grad /
    GDRQ operators).  Fold-BN quant path: data 12, weight 8 (single pass), backward = copy of the output gradient."""
    total = sm["act_elems"] + sm["weight_elems"]
    if op_type == "GDRQ_Fold_BN":
        return 12 * sm["act_elems"] + 8 * sm["weight_elems"] + 8 * sm.get("out_elems", 0)
    bwd_act = 8 if op_type == "Quantization_int8_V2" else 12
    return 12 * total + bwd_act * sm["act_elems"] + 8 * sm["weight_elems"]


def workload_config(workload, op_type, sm, batch, world):
    return {"workload": "%s quant path: fwd+bwd of all %d %s nodes (%d act + %d weight), per-GPU batch %d"
                        % (workload, sm["act_nodes"] + sm["weight_nodes"], op_type, sm["act_nodes"],
                           sm["weight_nodes"], batch),
            "elements_per_step": sm["act_elems"] + sm["weight_elems"],
            "alg_bytes_per_step": alg_bytes_per_step(op_type, sm),
            "parallelism": "dp%d" % world,
            "l2": "inputs larger than L2 (every node has its own tensors, touched once per step)",
            "mode": DROPIN_MODE}


def workload_nodes(workload, batch):
    from b200quant.workloads import WORKLOADS, summary
    fn, default_batch, op_type = WORKLOADS[workload]
    batch = batch or default_batch
    nodes = fn(batch)
    if op_type == "GDRQ_Fold_BN":   # the fold-BN operator replaces conv+BN pairs: the FC layer is not one
        nodes = [nd for nd in nodes if not nd[0].startswith("fc")]
    sm = summary(nodes)
    if op_type == "GDRQ_Fold_BN":
        sm["out_elems"] = sum(_numel(o) for o in foldbn_out_shapes(nodes))
    return nodes, sm, batch, op_type


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def foldbn_out_shapes(nodes):
    """conv output shape of every (act, weight) pair of a MobileNet-v1 node list (3x3 pad 1 / 1x1 pad 0; the stride is
    recovered from the next layer's input)."""
    acts = [s for _, k, s in nodes if k == "act"]
    wts = [s for _, k, s in nodes if k == "weight"]
    outs = []
    for i, (a, w) in enumerate(zip(acts, wts)):
        hw = acts[i + 1][2] if i + 1 < len(acts) and len(acts[i + 1]) == 4 else a[2]
        outs.append((a[0], w[0], hw, hw))
    return outs


# ---------------------------------------------------------------------------------------------------------
# clocks (NVML) sampled during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.02):
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
# our arm: nodes, steps
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


def build_nodes(torch, nodes, op_type, device, seed=5):
    """Allocate every node's tensors (inputs resident before the timed region) and create its operator."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for name, kind, shape in nodes:
        is_w = kind == "weight"
        op = make_op(op_type, is_w)
        if is_w:
            x = torch.empty(shape, device=device).normal_(0.0, (2.0 / _numel(shape[1:])) ** 0.5, generator=g)
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


class FoldBNLayer(object):
    """The quantization path of one GDRQ_Fold_BN node (symbol/fold_bn_v1_gdrq.py:53-96,113,122-129) WITHOUT its
    convolution (library code, :99-110): data path, weight path (per-channel, folded BN, bias), and the backward
    (gradient to bn_output only).  Makes the calls GDRQ_Fold_BN.forward/backward make, in the same order."""

    def __init__(self, torch, g, act_shape, w_shape, out_shape, device):
        cout = w_shape[0]
        self.x = torch.empty(act_shape, device=device).uniform_(-1.0, 1.0, generator=g)
        self.w = torch.empty(w_shape, device=device).normal_(0.0, (2.0 / _numel(w_shape[1:])) ** 0.5, generator=g)
        self.gamma = torch.empty(cout, device=device).uniform_(0.5, 1.5, generator=g)
        self.beta = torch.empty(cout, device=device).normal_(generator=g)
        self.mean = torch.empty(cout, device=device).normal_(generator=g)
        self.var = torch.empty(cout, device=device).uniform_(0.5, 1.5, generator=g)
        self.xq, self.wq = torch.empty_like(self.x), torch.empty_like(self.w)
        self.bias = torch.empty(cout, device=device)
        self.aux = [torch.ones(1, device=device), torch.ones(cout, device=device)]
        self.dy = torch.empty(out_shape, device=device).normal_(generator=g)
        self.d_bn = torch.empty_like(self.dy)
        self.init = True
        self.peer = None

    def forward(self):
        from b200quant import _kernels as K, _lib
        if self.peer is not None:
            self.peer.quantize_mean(_lib.UPD_TWICE_STORE if self.init else _lib.UPD_TWICE_EMA, self.x, self.xq,
                                    self.aux[0], 0.99, 1 - 0.99, 127.0)
        else:
            K.foldbn_data_fwd(self.x, self.xq, self.aux[0], self.init, 0.99)
        self.init = False
        K.foldbn_weight_fwd(self.w, self.wq, self.bias, self.aux[1], self.gamma, self.beta, self.mean, self.var, 1e-5,
                            True, True, True)

    def backward(self):
        from b200quant import _kernels as K
        K.assign(self.d_bn, "write", self.dy)


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
    every node copies its full input H2D and its full result D2H inside the timed region.  The straight-through
    backward of a host caller is dx <- dy between two HOST buffers: the library copies it host to host (no arithmetic,
    no PCIe round trip), so its bytes are not H2D/D2H bytes."""
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
        return float(hnodes[-1]["aux"][0])     # the step's result read on the host

    elems = sum(h["n"] for h in hnodes)
    ste_on_host = ctx.get_option("host_ste_copy") != 0
    bytes_in = 4 * elems * (1 if ste_on_host else 2)      # x (forward) [+ dy (backward) when staged through the GPU]
    bytes_out = 4 * elems * (1 if ste_on_host else 2)     # y [+ dx]
    return step, bytes_in, bytes_out, ste_on_host


# ---------------------------------------------------------------------------------------------------------
# reference CPU path (oracle/c): the only place bench.py executes oracle code -- as the timed baseline here, and as
# the checker in parity_check()
# ---------------------------------------------------------------------------------------------------------
class CpuReference(object):
    """The whole step -- forward then backward of EVERY node of the workload -- through the C/OpenMP restatement of the
    reference's operators, on all host threads.  Host buffers come from pools twice the largest tensor, handed out
    round-robin so consecutive nodes work on different memory (as the layers of a network do)."""

    def __init__(self, workload, batch):
        import numpy as np
        from oracle import c_oracle as co
        self.np, self.co = np, co
        co.use_all_host_threads()   # whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
        self.nodes, self.sm, self.batch, self.op_type = workload_nodes(workload, batch)
        self.workload = workload
        rng = np.random.default_rng(5)
        biggest = max(_numel(s) for _, _, s in self.nodes)
        pool = 2 * biggest
        self.px = rng.random(pool, dtype=np.float32)
        self.px *= 2.0
        self.px -= 1.0                                                    # uniform(-1, 1): data/imagenet.py:16
        self.pdy = rng.standard_normal(pool, dtype=np.float32)
        self.py = np.empty(pool, np.float32)
        self.pdx = np.empty(pool, np.float32)
        cur = 0
        self.state = []
        fold = self.op_type == "GDRQ_Fold_BN"
        outs = foldbn_out_shapes(self.nodes) if fold else None
        acts = [nd for nd in self.nodes if nd[1] == "act"]
        wts = [nd for nd in self.nodes if nd[1] == "weight"]
        order = self.nodes if not fold else [(a, w, o) for a, w, o in zip(acts, wts, outs)]
        for item in order:
            if fold:
                (_, _, ashape), (_, _, wshape), oshape = item
                n, no = _numel(ashape), _numel(oshape)
                if cur + max(n, no) > pool:
                    cur = 0
                cout = wshape[0]
                w = (rng.standard_normal(wshape) * (2.0 / _numel(wshape[1:])) ** 0.5).astype(np.float32)
                self.state.append(dict(x=self.px[cur:cur + n].reshape(ashape), y=self.py[cur:cur + n].reshape(ashape),
                                       dy=self.pdy[cur:cur + no], dx=self.pdx[cur:cur + no], w=w, wq=np.empty_like(w),
                                       bias=np.empty(cout, np.float32), aux0=np.ones(1, np.float32),
                                       aux1=np.ones(cout, np.float32),
                                       gamma=rng.uniform(0.5, 1.5, cout).astype(np.float32),
                                       beta=rng.standard_normal(cout).astype(np.float32),
                                       mean=rng.standard_normal(cout).astype(np.float32),
                                       var=rng.uniform(0.5, 1.5, cout).astype(np.float32), init=True))
                cur += max(n, no)
                continue
            name, kind, shape = item
            n = _numel(shape)
            if kind == "weight":
                x = (rng.standard_normal(shape) * (2.0 / _numel(shape[1:])) ** 0.5).astype(np.float32)
                st = dict(x=x, y=np.empty_like(x), dy=rng.standard_normal(shape).astype(np.float32),
                          dx=np.empty_like(x))
            else:
                if cur + n > pool:
                    cur = 0
                st = dict(x=self.px[cur:cur + n].reshape(shape), y=self.py[cur:cur + n].reshape(shape),
                          dy=self.pdy[cur:cur + n].reshape(shape), dx=self.pdx[cur:cur + n].reshape(shape))
                cur += n
            st.update(aux=np.ones(1, np.float32), w=(kind == "weight"), init=True)
            self.state.append(st)

    def step(self):
        co, op_type = self.co, self.op_type
        if op_type == "GDRQ_Fold_BN":
            for s in self.state:
                co.foldbn_data_fwd(s["x"], s["y"], s["aux0"], s["init"], 0.99)
                s["init"] = False
                co.foldbn_weight_fwd(s["w"], s["wq"], s["bias"], s["aux1"], s["gamma"], s["beta"], s["mean"], s["var"],
                                     1e-5, True, True, True)
            for s in reversed(self.state):
                co.ste_bwd(s["dy"], s["dx"])
            return
        if op_type == "GDRQ_PY":   # the attributes of make_op: nbits 8 -> 255 levels, ktimes 3, lamda 0.001
            for s in self.state:
                co.gdrq_fwd(s["x"], s["y"], s["aux"], s["w"], False, True, 255.0, 3.0, 0.001)
            for s in reversed(self.state):
                if s["w"]:
                    co.ste_bwd(s["dy"], s["dx"])
                else:
                    co.gdrq_bwd(s["x"], s["dy"], s["dx"], s["aux"])
            return
        variant = 1 if op_type == "ClipGrad_Quantization_int8" else 0
        for s in self.state:
            co.minmax_quant_fwd(variant, s["x"], s["y"], s["aux"], s["w"], False, True,
                                variant == 1 and s["init"] and not s["w"], 0.99)
            s["init"] = False
        for s in reversed(self.state):
            if variant == 1 and not s["w"]:
                co.clipgrad_bwd(s["x"], s["dy"], s["dx"], s["aux"])
            else:
                co.ste_bwd(s["dy"], s["dx"])

    def run(self, steps, warmup):
        """-> (seconds per step: total / steps, median, all step times)"""
        for _ in range(warmup):
            self.step()
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            self.step()
            times.append(time.perf_counter() - t0)
        s = sorted(times)
        return sum(times) / len(times), s[len(s) // 2], times


def cpu_baseline(workload, batch, steps=3, warmup=1):
    ref = CpuReference(workload, batch)
    mean, median, times = ref.run(steps, warmup)
    return dict(value=ref.batch / median, unit=UNIT, cores=ref.co.num_threads(), kind="port",
                sample="the full step: fwd+bwd of all %d nodes (%d elements), %d timed steps after %d warm-up, median; "
                       "oracle/c = C/OpenMP restatement of the reference's mx.nd op chain, all host threads"
                       % (len(ref.nodes), ref.sm["act_elems"] + ref.sm["weight_elems"], steps, warmup),
                ms_per_step=median * 1e3, ms_per_step_mean=mean * 1e3)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(args.workload, args.batch)
    mean, median, times = ref.run(max(args.steps, 1), max(args.warmup, 1))
    value = ref.batch / median
    world = max(args.gpus, 1)
    sample = ("every step = fwd+bwd of ALL %d nodes of the workload (%d elements, nothing extrapolated); value from the "
              "median step of %d" % (len(ref.nodes), ref.sm["act_elems"] + ref.sm["weight_elems"], len(times)))
    line = {"impl": "reference",
            "metric": METRIC if args.workload == "resnet50_int8" else args.workload + "_quant_path_images_per_sec",
            "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": median * 1e3, "ms_per_step_mean": mean * 1e3,
            "ms_per_step_min_max": [min(times) * 1e3, max(times) * 1e3],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, ref.op_type, ref.sm, ref.batch, world),
            "reference": "oracle/c: C + OpenMP restatement of the reference's MXNet CPU CustomOp chain, one pass and one "
                         "temporary per mx.nd call (libmxnet is not installable here); runs on rank 0's host cores",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.co.num_threads(), "kind": "port",
                             "sample": sample},
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


# ---------------------------------------------------------------------------------------------------------
# whole-network leg
# ---------------------------------------------------------------------------------------------------------
def full_model_leg(torch, device, batch, steps=8, warmup=3, world=1, rank=0):
    """One SGD step of the whole ResNet-50 int8-QAT network of symbol/resnet_int8.py (what core/solver.py:98-163 +
    train.py:209-241 run): cuDNN/cuBLAS convolutions, BatchNorm and pooling from torch, every conv/FC input and weight
    through this package's Quantization_int8_V2 nodes, SGD momentum 0.9 wd 1e-4 (edict_config.py:58-60).  With world > 1
    it is the data-parallel training the reference runs through Module + KVStore (train.py:34-35): activation
    thresholds through the fused peer-memory exchange, gradients through DistributedDataParallel (NCCL)."""
    from b200quant.harness import ResNetInt8, quant_nodes
    torch.manual_seed(11)
    torch.backends.cudnn.benchmark = True     # library convolutions: let cuDNN pick its fastest algorithms
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
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = train_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
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
            "note": "whole ResNet-50 int8-QAT SGD step (momentum 0.9, wd 1e-4), eager torch autograd, weak scaling; "
                    "convolutions/BN are library code, the 108 quantization nodes are this package's"}


# ---------------------------------------------------------------------------------------------------------
# parity check inside the bench (after the timed region): a sampled activation node against the oracle
# ---------------------------------------------------------------------------------------------------------
def parity_check(torch, dist, nodes, op_type, world, rank, exchange):
    """Fresh operator (attached to the same cross-rank exchange as the timed nodes) on one sampled activation node:
    every rank's aux must be identical and equal the reference update fed the max over ranks of the statistic, and every
    rank's output must be bit-identical to oracle/c given that threshold.  The oracle is the checker, not the product."""
    import numpy as np
    from oracle import c_oracle as co
    co.use_all_host_threads()
    F = np.float32
    cand = [nd for nd in nodes if nd["kind"] == "act" and 8e6 <= nd["n"] <= 60e6] or \
           [nd for nd in nodes if nd["kind"] == "act"]
    nd = cand[len(cand) // 2]
    op = make_op(op_type, False)
    if exchange is not None:
        op.peer, op.sync = exchange, None
    elif world > 1:
        from b200quant.dist import ThresholdSync
        op.sync = ThresholdSync()
    x, y, aux = nd["x"], torch.empty_like(nd["x"]), torch.ones(1, device=nd["x"].device)
    op.forward(True, ["write"], [x], [y], [aux])
    torch.cuda.synchronize()
    hx = x.cpu().numpy()
    a_old = F(1.0)
    if op_type == "GDRQ_PY":
        stat = F(F(np.abs(hx).sum(dtype=np.float64)) / F(hx.size))                       # GDRQ.py:67-72
    else:
        stat = F(np.abs(hx).max())
    if world > 1:
        t = torch.tensor([float(stat)], device=x.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stat = F(t.item())
    yr = np.empty_like(hx)
    got_aux = aux.cpu().numpy().astype(F)
    a = got_aux.copy()          # the oracle sweeps with the threshold every rank holds (checked against `want` below)
    if op_type == "GDRQ_PY":
        thr = F(F(3.0) * stat)
        want = F(a_old + F(F(0.001) * F(a_old - thr)))                                   # GDRQ.py:76
        co.gdrq_fwd(hx, yr, a, False, True, True, 255.0, 3.0, 0.001)                     # fix_alpha: threshold given
        # mean-based threshold: fp64 sum rounded once on both sides -> 1e-6 relative (BASELINE.json north_star)
        ok_aux = bool(abs(float(got_aux[0]) - float(want)) <= 1e-6 * abs(float(want)))
    else:
        if op_type == "ClipGrad_Quantization_int8":
            want = stat                                                                  # first batch: clip_grad...py:42-44
            co.minmax_quant_fwd(1, hx, yr, a, False, False, False, False, 0.99)
        else:
            want = F(F(a_old * F(0.99)) + F(stat * F(1 - 0.99)))                         # quant_ops.py:37
            co.minmax_quant_fwd(0, hx, yr, a, False, False, False, False, 0.99)
        ok_aux = bool(got_aux.view(np.uint32)[0] == np.array([want], F).view(np.uint32)[0])   # bit-exact
    ok_out = bool(torch.equal(y.view(torch.int32), torch.from_numpy(yr).to(y.device).view(torch.int32)))
    same = True
    if world > 1:
        lo, hi = aux.clone(), aux.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        flag = torch.tensor([1 if (ok_aux and ok_out) else 0], device=x.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all = bool(flag.item())
    else:
        ok_all = ok_aux and ok_out
    return {"checked": True, "ok": bool(ok_all and same), "node": nd["name"], "shape": list(nd["shape"]),
            "aux_identical_across_ranks": same, "aux_equals_oracle": ok_aux, "output_bit_equal_oracle": ok_out,
            "ranks": world, "oracle": "oracle/c fed the max over ranks of the per-rank statistic"}


# ---------------------------------------------------------------------------------------------------------
# one workload on our arm
# ---------------------------------------------------------------------------------------------------------
def kernel_table(ctx, peak_gbs):
    kinds = {1: "reduce_flat (max|x| + EMA update)", 2: "qdq_flat_hot (QDQ sweep)", 3: "bwd_flat (STE copy)",
             4: "bwd_flat (clip mask)", 5: "segmented/other", 6: "fused single-launch forward (resident tensors)"}
    kernels = {}
    LARGE = 64e6   # algorithmic bytes: the tensors the ">= 70 % of HBM peak on large tensors" target is about
    for k, name in kinds.items():
        try:
            kms, kbytes, kn = ctx.timing_read(k)
        except Exception:
            continue
        if kn:
            kernels[name] = {"launches": kn, "ms_total": kms, "alg_bytes_total": kbytes,
                             "achieved_gbs": kbytes / kms / 1e6, "frac_of_peak": kbytes / kms / 1e6 / peak_gbs}
            lms, lbytes, ln = ctx.timing_read(k, min_bytes=LARGE)
            if ln:
                kernels[name]["large_tensors"] = {"min_alg_bytes": LARGE, "launches": ln, "ms_total": lms,
                                                  "alg_bytes_total": lbytes, "achieved_gbs": lbytes / lms / 1e6,
                                                  "frac_of_peak": lbytes / lms / 1e6 / peak_gbs}
    return kernels, kinds


def forward_by_size(ctx, sizes, peak_gbs):
    """forward (reduction + sweep, 12 B/element algorithmic) of every distinct activation size, from the event records of
    the timed pass: where the launch-latency regime ends.  Uses the two kernels' records of that exact byte count."""
    out = {}
    for n in sorted(set(sizes), reverse=True):
        t_ms, launches = 0.0, 0
        for kind, per_elem in ((1, 4.0), (2, 8.0), (6, 12.0)):
            b = per_elem * n
            try:
                ms, _, cnt = ctx.timing_read(kind, min_bytes=b * 0.9999, max_bytes=b * 1.0001)
            except Exception:
                continue
            if cnt and kind != 6:
                t_ms += ms / cnt
                launches += 1
            elif cnt:
                t_ms, launches = ms / cnt, 2
                break
        if launches == 2 and t_ms > 0:
            out["%d" % n] = {"mb": round(4 * n / 1e6, 1), "us": round(t_ms * 1e3, 2),
                             "frac_of_peak_on_12B_per_elem": round(12.0 * n / t_ms / 1e6 / peak_gbs, 4)}
    return out


def forward_by_size_in_stream(torch, nodes, peak_gbs, reps=3):
    """The same table measured WITHOUT events between the kernels: all activation nodes of one size run their forward
    back to back (programmatic dependent launch intact, as in the step), two events around the whole group.  Nodes of a
    size are distinct tensors (groups above the L2 size by construction for all but the smallest sizes).  Single-rank
    kernels only (at N > 1 the exchange makes every node wait for its peers, which is not a property of a size)."""
    by_size = {}
    for nd in nodes:
        if nd["kind"] == "act":
            by_size.setdefault(nd["n"], []).append(nd)
    out = {}
    for n in sorted(by_size, reverse=True):
        group = by_size[n]

        def run():
            for nd in group:
                nd["op"].forward(True, ["write"], [nd["x"]], [nd["y"]], [nd["aux"]])
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(group))
        out["%d" % n] = {"mb": round(4 * n / 1e6, 1), "nodes": len(group), "us": round(us, 2),
                         "frac_of_peak_on_12B_per_elem": round(12.0 * n / us / 1e3 / peak_gbs, 4)}
    return out


def measure_kernels(torch, ctx, step, reps, peak_gbs, act_sizes=None):
    ctx.set_option("timing", 1)
    ctx.timing_read(0, reset=True)
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    kernels, kinds = kernel_table(ctx, peak_gbs)
    if act_sizes:
        kernels["forward_by_activation_size"] = forward_by_size(ctx, act_sizes, peak_gbs)
    ctx.timing_read(0, reset=True)
    ctx.set_option("timing", 0)
    return kernels, kinds


def load_peaks():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return peak_gbs, src


def compact_workload(torch, dist, ctx, name, device, world, rank, steps, peak_gbs, exchange):
    """ms/step, img/s and per-kernel fractions of one of the other BASELINE.json configs (eager, drop-in mode)."""
    import gc
    nodes_l, sm, batch, op_type = workload_nodes(name, 0)
    free, _ = torch.cuda.mem_get_info()
    need = 16 * (sm["act_elems"] + sm["weight_elems"]) + (8 * sm.get("out_elems", 0))
    if need > 0.9 * free:   # ResNeXt-101 at batch 256 needs 86 GB of tensors
        nodes_l, sm, batch, op_type = workload_nodes(name, batch // 2)
    out = {"op_type": op_type, "batch": batch, "nodes": sm["act_nodes"] + sm["weight_nodes"]}
    g = torch.Generator(device=device).manual_seed(17 + rank)
    if op_type == "GDRQ_Fold_BN":
        acts = [s for _, k, s in nodes_l if k == "act"]
        wts = [s for _, k, s in nodes_l if k == "weight"]
        layers = [FoldBNLayer(torch, g, a, w, o, device) for a, w, o in zip(acts, wts, foldbn_out_shapes(nodes_l))]
        if exchange is not None:
            for l in layers:
                l.peer = exchange

        def step():
            for l in layers:
                l.forward()
            for l in reversed(layers):
                l.backward()
        out["path"] = "data path (2*mean|x| + EMA + clip/QDQ) + per-channel weight path (fold, 2*mean|w'|, clip, QDQ, " \
                      "bias) + backward copy to bn_output of all 27 conv layers; convolution excluded (library code)"
    else:
        nodes = build_nodes(torch, nodes_l, op_type, device, seed=17 + rank)
        if exchange is not None:
            for nd in nodes:
                if nd["kind"] == "act":
                    nd["op"].peer, nd["op"].sync = exchange, None

        def step():
            run_step(nodes)
    for _ in range(3):
        step()
    l0 = ctx.launch_count()
    ms = time_steps(torch, dist, step, steps, world) / steps
    out["gpu_launches_per_step"] = (ctx.launch_count() - l0) // steps
    alg = alg_bytes_per_step(op_type, sm)
    out.update(ms_per_step=ms, images_per_sec=world * batch / (ms / 1e3), alg_bytes_per_step=alg,
               hbm_frac_whole_step=alg / (ms / 1e3) / 1e9 / peak_gbs, mode=DROPIN_MODE)
    kernels, _ = measure_kernels(torch, ctx, step, 2, peak_gbs)
    out["kernels_large_tensor_frac"] = {k: round(v["large_tensors"]["frac_of_peak"], 4)
                                        for k, v in kernels.items() if "large_tensors" in v}
    out["kernels_pooled_frac"] = {k: round(v["frac_of_peak"], 4) for k, v in kernels.items()}
    del step
    if op_type == "GDRQ_Fold_BN":
        del layers
    else:
        del nodes
    gc.collect()
    torch.cuda.empty_cache()
    return out


def micro_leg(torch, ctx, peak_gbs):
    """Second-tier operators and the weight kernels on the micro-benchmark inputs of SURVEY.md 8d: achieved GB/s on
    ALGORITHMIC bytes (CUDA events, 2^26-element inputs, L2 flushed by the input size: 256 MB per tensor)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import microbench2
        return microbench2.run(torch, ctx, peak_gbs, quick=True)
    except Exception as e:  # pragma: no cover
        return {"error": str(e).splitlines()[0][:200]}


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

    node_list, sm, batch, op_type = workload_nodes(args.workload, args.batch)
    primary = args.workload == "resnet50_int8"
    peak_gbs, peak_src = load_peaks()
    full_model = None
    if not args.no_full_model and not args.profile and primary:
        try:   # whole-network leg, run first while the device memory is still free
            full_model = full_model_leg(torch, device, batch, world=world, rank=rank)
        except Exception as e:  # pragma: no cover
            full_model = {"images_per_sec": None, "error": str(e).splitlines()[0][:200]}
            if world > 1:
                raise
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    total_elems = sm["act_elems"] + sm["weight_elems"]
    ctx = _lib.context(local)
    for k, v in os.environ.items():   # tuning experiments: B2Q_OPT_<option>=<int> (results never depend on them)
        if k.startswith("B2Q_OPT_"):
            ctx.set_option(k[len("B2Q_OPT_"):].lower(), int(v))
    if op_type == "GDRQ_Fold_BN":   # as a --workload it only has the compact treatment
        res = compact_workload(torch, dist, ctx, args.workload, device, world, rank, args.steps, peak_gbs, None)
        line = {"metric": args.workload + "_quant_path_images_per_sec", "value": res["images_per_sec"], "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.workload, op_type, sm, batch, world), "detail": res,
                "gpu_launches": res["gpu_launches_per_step"] * args.steps}
        if rank == 0:
            emit(line)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    nodes = build_nodes(torch, node_list, op_type, device, seed=5 + rank)
    minmax = op_type in ("Quantization_int8_V2", "ClipGrad_Quantization_int8")
    alg_bytes_step = alg_bytes_per_step(op_type, sm)

    bucket = None
    exchange = "none"
    grad_exchange = "none"
    peer_ex = None
    if world == 1 and os.environ.get("B2Q_DEBUG_PEER_WORLD1") == "1":   # experiment: mailbox kernels without a peer
        from b200quant.dist import attach_peer_exchange
        peer_ex = attach_peer_exchange([nd["op"] for nd in nodes], device)
        exchange = "fused peer-memory kernels at world 1 [DEBUG experiment]"
    if world > 1:   # data parallel: thresholds max over ranks per activation node, weight grads allreduce(sum)
        exchange = "nccl allreduce(max) per activation node"
        attach_threshold_sync([nd["op"] for nd in nodes])
        if os.environ.get("B2Q_EXCHANGE", "peer") == "peer":
            try:   # fused peer-memory exchange (NVLink): no NCCL call on the forward critical path
                from b200quant.dist import attach_peer_exchange
                peer_ex = attach_peer_exchange([nd["op"] for nd in nodes], device)
                exchange = "fused peer-memory kernels (b2q_peer_%s_quant_fwd_f32)" % ("minmax" if minmax else "meanabs",)
            except Exception as e:  # pragma: no cover
                exchange += " (peer path unavailable: %s)" % (str(e).splitlines()[0][:100],)
        # operators without a fused exchange (grouped GDRQ thresholds) keep the NCCL call on their forward path
        nccl_forward = any(getattr(nd["op"], "sync", None) is not None for nd in nodes)
        if nccl_forward and exchange.startswith("fused peer"):
            exchange = "nccl allreduce(max) of the per-node statistic (no fused exchange for this operator)"
        wn = [nd for nd in nodes if nd["kind"] == "weight"]
        grad_exchange = "nccl allreduce(avg), one flat bucket"
        # measured (profiles/r02q_gradient_exchange.md): NCCL's allreduce (in-switch reduction at 8 GPUs) beats the
        # peer-memory slice-owner kernel by 0.4 % of the step at 2 GPUs and 1.5 % at 8, so it stays the default here
        if peer_ex is not None and os.environ.get("B2Q_GRAD_EXCHANGE", "nccl") == "peer":
            from b200quant.dist import PeerGradBucket   # slice-owner kernel over NVLink: no NCCL call in the step at all
            bucket = PeerGradBucket([nd["shape"] for nd in wn], peer_ex)
            grad_exchange = "peer-memory slice-owner kernel (b2q_peer_allreduce_sum_f32), one flat bucket"
        else:
            bucket = GradBucket([nd["shape"] for nd in wn], device)
        for nd, view in zip(wn, bucket.views):
            nd["dx"] = view
        if os.environ.get("B2Q_DEBUG_SKIP_GRAD_ALLREDUCE") == "1":   # experiment only: isolates the exchange cost
            bucket.allreduce = lambda *a, **k: None
            exchange += " [DEBUG: gradient allreduce skipped -- not a valid result]"

    # The weight gradients are complete once every weight node's backward has run; their allreduce(sum) then runs on a
    # side stream while the activation backward sweeps continue (standard overlap of gradient communication with the
    # backward pass).  In the per-node (drop-in) order the weight backward calls come first for that reason.
    side = torch.cuda.Stream(priority=-1) if bucket is not None else None   # the collective's blocks are dispatched first
    wnodes = [nd for nd in nodes if nd["kind"] == "weight"]
    anodes = [nd for nd in nodes if nd["kind"] == "act"]

    def step_local():
        run_step(nodes)

    def step():
        if bucket is None:
            return run_step(nodes)
        for nd in nodes:
            nd["op"].forward(True, ["write"], [nd["x"]], [nd["y"]], [nd["aux"]])
        for nd in reversed(wnodes):
            nd["op"].backward(["write"], [nd["dy"]], [nd["x"]], [nd["y"]], [nd["dx"]], [nd["aux"]])
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            bucket.allreduce()
        for nd in reversed(anodes):
            nd["op"].backward(["write"], [nd["dy"]], [nd["x"]], [nd["y"]], [nd["dx"]], [nd["aux"]])
        torch.cuda.current_stream().wait_stream(side)

    # same semantics, weight nodes batched: all 54 weight tensors in 2 launches forward / 1 launch backward
    from b200quant.multi import WeightGroup
    group = None
    if minmax or op_type == "GDRQ_PY":
        group = WeightGroup([nd["op"] for nd in wnodes], [nd["x"] for nd in wnodes], [nd["y"] for nd in wnodes],
                            [nd["aux"] for nd in wnodes], [nd["dy"] for nd in wnodes], [nd["dx"] for nd in wnodes])

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

    # ---- main timed region: the drop-in mode (eager, every call goes through the CustomOp protocol) ----
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms_eager = time_steps(torch, dist, step, args.steps, world)
    launches = ctx.launch_count() - l0
    clocks = sampler.finish()
    timings = {DROPIN_MODE: ms_eager / args.steps}
    best_ms, best_mode = ms_eager, DROPIN_MODE

    if args.profile:
        group = None
        args.no_graph = args.no_e2e = args.no_cpu = args.no_workloads = args.no_micro = True
    if group is not None:
        for _ in range(3):
            step_multi()
        ms_multi = time_steps(torch, dist, step_multi, args.steps, world)
        timings["eager, weight nodes batched (WeightGroup)"] = ms_multi / args.steps
        if ms_multi < best_ms:
            best_ms, best_mode = ms_multi, "eager, weight nodes batched (WeightGroup)"

    # ---- the same steps replayed from a CUDA graph (the graph holds our kernels only) ----
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
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        for part in parts:
                            part()
                        torch.cuda.synchronize()
                        for part in parts:
                            g_ = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(g_, stream=s):
                                part()
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
                if t < best_ms:
                    best_ms, best_mode = t, label
                del graph, graphs, replay
            except Exception as e:  # pragma: no cover
                timings[label] = "failed: %s" % (str(e).splitlines()[0][:120],)

    value = world * batch * args.steps / (ms_eager / 1e3)

    # ---- per-kernel timing for the roofline (second timed region, events around every flat-kernel launch) ----
    kernels, kinds = measure_kernels(torch, ctx, step, 0 if args.profile else min(args.steps, 5), peak_gbs,
                                     act_sizes=[nd["n"] for nd in nodes if nd["kind"] == "act"])
    fwd_by_size = kernels.pop("forward_by_activation_size", None)
    fwd_by_size_stream = None
    if world == 1 and not args.profile:
        try:
            fwd_by_size_stream = forward_by_size_in_stream(torch, nodes, peak_gbs)
        except Exception as e:  # pragma: no cover
            fwd_by_size_stream = {"error": str(e).splitlines()[0][:160]}
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = {"source": tj.get("source"), "launch_alg_bytes": tj.get("alg_bytes_of_this_launch"),
                        "launch_duration_us": tj.get("duration_us"),
                        "note": "ncu dram__bytes_read+write of one captured launch; compare with launch_alg_bytes, not "
                                "with the average alg_bytes_per_launch"}
    except Exception:
        pass
    dom = kernels.get(kinds[2], {})
    roofline = {"bound": "hbm", "kernel": "qdq_flat_hot_kernel", "achieved": dom.get("achieved_gbs"),
                "peak": peak_gbs, "unit": "GB/s", "frac": dom.get("frac_of_peak"), "traffic": traffic,
                "peak_source": peak_src, "traffic_detail": traffic_note,
                "alg_bytes_per_launch": (dom.get("alg_bytes_total", 0) / dom["launches"]) if dom else None,
                "avg_launch_ms": (dom.get("ms_total", 0) / dom["launches"]) if dom else None}

    metric = METRIC if primary else args.workload + "_quant_path_images_per_sec"
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_eager / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, op_type, sm, batch, world),
            "value_dropin": value, "mode": DROPIN_MODE,
            "value_best": world * batch * args.steps / (best_ms / 1e3), "mode_best": best_mode,
            "threshold_exchange": exchange,
            "gradient_exchange": grad_exchange if world > 1 else "none",
            "ms_per_step_by_mode": timings,
            "hbm_frac_whole_step": alg_bytes_step / (ms_eager / args.steps / 1e3) / 1e9 / peak_gbs,
            "hbm_frac_whole_step_best_mode": alg_bytes_step / (best_ms / args.steps / 1e3) / 1e9 / peak_gbs,
            "roofline": roofline, "kernels": kernels, "forward_by_activation_size": fwd_by_size,
            "forward_by_activation_size_in_stream": fwd_by_size_stream,
            "forward_by_size_note": "forward_by_activation_size: from the per-launch event records (the events break the "
                                    "programmatic dependent launch between reduction and sweep, so small and mid sizes read "
                                    "low); ..._in_stream: all nodes of a size back to back with no event in between, as in "
                                    "the step",
            "kernels_note": "event-timed per launch in a separate pass, all tensor sizes pooled (54 of the launches per "
                            "kind are weight tensors of a few KB..MB that cost a launch each); a kernel that follows a "
                            "sweep also pays for the write-back of the output lines its predecessor left dirty in L2 "
                            "(profiles/r02_instep_reduce.md), so reduce_flat reads low and qdq_flat_hot high; "
                            "hbm_frac_whole_step is the unbiased figure",
            "clocks": clocks, "gpu_launches": launches}

    # ---- parity of what was just timed: a sampled node against the oracle, aux identical on every rank ----
    if not args.profile:
        try:
            line["parity"] = parity_check(torch, dist, nodes, op_type, world, rank, peer_ex)
            line["parity_checked"] = bool(line["parity"]["ok"])
        except Exception as e:  # pragma: no cover
            line["parity"] = {"checked": False, "error": str(e).splitlines()[0][:200]}
            line["parity_checked"] = False
            if world > 1:
                raise

    # ---- e2e: host buffers through the host C ABI (rank-local; N ranks run it concurrently) ----
    if not args.no_e2e and op_type == "Quantization_int8_V2":
        try:
            host_cores = None
            if world > 1:   # every rank its own host cores (its GPU's NUMA node when the platform reports one)
                from b200quant.dist import pin_rank_to_host_cores
                host_cores = pin_rank_to_host_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)), local)
            hstep, b_in, b_out, ste_on_host = host_step_factory(torch, nodes, ctx)
            hstep()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                hstep()
            torch.cuda.synchronize()
            dt_local = time.perf_counter() - t0
            dt = dt_local
            per_rank = [dt_local]
            if world > 1:
                t = torch.tensor([dt], device="cuda")
                allt = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allt, t)
                per_rank = [float(a.item()) for a in allt]
                dt = max(per_rank)
            line["e2e"] = {"value": world * batch * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": b_in,
                           "d2h_bytes_per_step": b_out, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
                           "per_rank_gbs_each_direction": [round(b_in * args.e2e_steps / t_ / 1e9, 2) for t_ in per_rank],
                           "host_cores_per_rank": (len(host_cores) if host_cores else os.cpu_count()),
                           "ste_backward": ("host-to-host copy inside the library (no PCIe round trip, no arithmetic)"
                                            if ste_on_host else "staged through the GPU (H2D dy, D2H dx)"),
                           "path": "CustomOp.forward/backward with pinned HOST tensors -> b2q_*_host_f32 (three-stream "
                                   "pipeline over a staging ring; PCIe-bound)"}
            del hstep
        except Exception as e:  # pragma: no cover - e.g. not enough lockable host memory
            line["e2e"] = {"value": None, "unit": UNIT, "error": str(e).splitlines()[0][:200]}
            if world > 1:
                raise

    # ---- the other configs, compact ----
    if primary and not args.no_workloads:
        import gc
        if group is not None:
            group.close()
        del nodes, wnodes, anodes, group, bucket
        gc.collect()
        torch.cuda.empty_cache()
        line["workloads"] = {}
        for name in ("mobilenet_v1_gdrq", "resnext101_clipgrad", "mobilenet_v1_foldbn"):
            try:
                line["workloads"][name] = compact_workload(torch, dist, ctx, name, device, world, rank,
                                                           min(args.steps, 10), peak_gbs, peer_ex)
            except Exception as e:  # pragma: no cover
                line["workloads"][name] = {"error": str(e).splitlines()[0][:200]}
                if world > 1:
                    raise
    if primary and rank == 0 and world == 1 and not args.no_micro:
        line["micro"] = micro_leg(torch, ctx, peak_gbs)

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline(args.workload, batch)
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": str(e).splitlines()[0][:200]}

    if full_model is not None:
        if full_model.get("ms_per_step"):
            full_model["quant_path_share_of_step"] = (ms_eager / args.steps) / full_model["ms_per_step"]
        line["full_model"] = full_model

    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        if peer_ex is not None:
            torch.cuda.synchronize()
            peer_ex.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
