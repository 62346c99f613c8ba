"""CPU checks of the benchmark plumbing: the node inventories behind the bench workloads, the
reference arm's JSON contract, the harness topology, and that the product arm refuses to run without a CUDA device."""
import collections
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_node_inventories_match_the_symbol_files():
    from b200quant.workloads import WORKLOADS, summary
    want = {   # nodes (act + weight), activation elements, weight elements at the workload's batch
        "simple_cifar": (3, 3, 164096, 872),
        "resnet50_int8": (54, 54, 2730098688, 25502912),
        "mobilenet_v1_gdrq": (28, 28, 1316880384, 4209088),
        "mobilenet_v1_foldbn": (28, 28, 1316880384, 4209088),
        "resnext101_clipgrad": (105, 105, 5318377472, 44038848),
    }
    for name, (fn, batch, _op) in WORKLOADS.items():
        sm = summary(fn(batch))
        assert (sm["act_nodes"], sm["weight_nodes"], sm["act_elems"], sm["weight_elems"]) == want[name], name


def test_resnet50_activation_shapes_are_appendix_b():
    from b200quant.workloads import resnet50_nodes
    acts = collections.Counter(tuple(shape) for _, kind, shape in resnet50_nodes(256) if kind == "act")
    assert acts == collections.Counter({
        (256, 256, 56, 56): 4, (256, 128, 56, 56): 1, (256, 512, 28, 28): 5, (256, 64, 56, 56): 8, (256, 256, 28, 28): 1,
        (256, 1024, 14, 14): 7, (256, 3, 224, 224): 1, (256, 128, 28, 28): 7, (256, 512, 14, 14): 1,
        (256, 2048, 7, 7): 2, (256, 256, 14, 14): 11, (256, 512, 7, 7): 5, (256, 2048): 1})
    weights = collections.Counter(tuple(shape) for _, kind, shape in resnet50_nodes(256) if kind == "weight")
    assert weights[(512, 512, 3, 3)] == 3 and weights[(1000, 2048)] == 1 and weights[(64, 3, 7, 7)] == 1
    assert sum(weights.values()) == 54


def test_harness_resnet_int8_matches_the_inventory():
    from b200quant.harness import ResNetInt8, quant_nodes
    from b200quant.workloads import resnet50_nodes
    model = ResNetInt8()
    names = [n for n, m in model.named_modules() if m in set(quant_nodes(model))]
    assert len(names) == 108
    weight_shapes = collections.Counter(tuple(p.shape) for n, p in model.named_parameters()
                                        if n.endswith(".weight") and p.dim() in (2, 4) and "bn" not in n)
    want = collections.Counter(tuple(shape) for _, kind, shape in resnet50_nodes(256) if kind == "weight")
    assert weight_shapes == want


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--batch", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "resnet50_int8_qat_quant_path_images_per_sec"
    assert d["unit"] == "img/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm runs EVERY node of the workload (nothing extrapolated) and describes it exactly as the product arm does
    import bench
    nodes, sm, batch, op_type = bench.workload_nodes("resnet50_int8", 8)
    assert d["config"] == bench.workload_config("resnet50_int8", op_type, sm, batch, 1)
    assert d["config"]["elements_per_step"] == sm["act_elems"] + sm["weight_elems"] and len(nodes) == 108
    assert "ALL 108 nodes" in d["cpu_baseline"]["sample"] and d["ms_per_step"] > 0
    # other ranks of a multi-rank launch exit 0 without work
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--batch", "8"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_reference_arm_runs_the_other_workloads_too():
    """MobileNet GDRQ, ResNeXt clip-grad and the fold-BN quant path through oracle/c (small batch: CPU suite budget)."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    for wl, nodes in (("mobilenet_v1_gdrq", 56), ("resnext101_clipgrad", 210), ("mobilenet_v1_foldbn", 54)):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                              "--warmup", "1", "--batch", "2", "--workload", wl], capture_output=True, text=True,
                             timeout=600, cwd=ROOT, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads(out.stdout.strip().splitlines()[-1])
        assert d["metric"] == wl + "_quant_path_images_per_sec" and d["value"] > 0
        assert "all %d " % nodes in d["config"]["workload"]
