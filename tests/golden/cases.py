"""Golden-vector case list shared by the generator (reference-over-shim), the oracle tests and the CUDA tests.

A case = one operator instance (op_type + the *string* attributes the reference Props take) driven through
a sequence of steps.  Each step is ``(is_train, req, do_backward)``; inputs are regenerated per step from a
seeded RNG with a per-step scale so EMA / first-batch-init / delay_quant state machines are exercised.
"""
import numpy as np

F = np.float32


def make_input(kind, shape, rng, step):
    scale = F([1.0, 1.7, 0.6, 2.3][step % 4])
    n = int(np.prod(shape))
    if kind == "normal":
        return (rng.standard_normal(shape).astype(F) * scale).astype(F)
    if kind == "uniform":  # data/imagenet.py:16 synthetic iterator distribution
        return (rng.uniform(-1, 1, size=shape).astype(F) * scale).astype(F)
    if kind == "relu":
        return np.maximum(rng.standard_normal(shape).astype(F) * scale, F(0)).astype(F)
    if kind == "pos":
        return (rng.uniform(0.5, 1.5, size=shape).astype(F)).astype(F)
    if kind == "ties":
        # absmax = 127/8 -> quant unit exactly 1/8; every other value sits exactly on a (k+0.5)/8 tie
        k = rng.integers(-126, 126, size=n).astype(F)
        v = ((k + F(0.5)) * F(0.125)).astype(F)
        v[::2] = (k[::2] * F(0.125) + rng.uniform(-0.06, 0.06, size=v[::2].shape).astype(F)).astype(F)
        v[0] = F(127.0 / 8.0)
        v[-1] = F(-0.0)
        return v.reshape(shape)
    if kind == "unit":  # values in (-1.2, 1.2): exercises QIL prune / clip regions
        return rng.uniform(-1.2, 1.2, size=shape).astype(F)
    raise KeyError(kind)


T, E = True, False
W, A, N = "write", "add", "null"

_QI8 = dict(quant_mode="minmax", ema_decay="0.99")


def _qi8(**kw):
    d = dict(_QI8)
    d.update({k: str(v) for k, v in kw.items()})
    return d


_FOLD = dict(quant_mode="minmax", ema_decay="0.99", name="fold_bn", no_bias="True", eps="1e-05",
             momentum="0.9", fix_gamma="False", dilate="(1, 1)")


def _fold(**kw):
    d = dict(_FOLD)
    d.update({k: str(v) for k, v in kw.items()})
    return d


def _fold_inputs(n, cin, h, w, cout, group, k, stride, pad):
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    return [("uniform", (n, cin, h, w)), ("normal", (cout, cin // group, k, k)), ("normal", (n, cout, oh, ow)),
            ("pos", (cout,)), ("normal", (cout,)), ("normal", (cout,)), ("pos", (cout,))]


CASES = [
    # ---- Quantization_int8_V2 (symbol/quant_ops.py) ------------------------------------------------
    dict(id="qi8v2_weight_pertensor", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=True, is_weight_perchannel=False, delay_quant=0),
         inputs=[("normal", (8, 3, 3, 3))], aux_init=[1.0],
         steps=[(T, W, True), (T, A, True), (E, W, False)]),
    dict(id="qi8v2_weight_perchannel", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=True, is_weight_perchannel=True, delay_quant=0),
         inputs=[("normal", (6, 4, 3, 3))], aux_init=[1.0],
         steps=[(T, W, True), (E, W, False), (T, A, True)]),
    dict(id="qi8v2_weight_depthwise", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=True, is_weight_perchannel=True, delay_quant=0),
         inputs=[("normal", (16, 1, 3, 3))], aux_init=[1.0],
         steps=[(T, W, True)]),
    dict(id="qi8v2_weight_ties", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=True, is_weight_perchannel=False, delay_quant=0),
         inputs=[("ties", (5, 7, 3, 3))], aux_init=[1.0],
         steps=[(T, W, False)]),
    dict(id="qi8v2_act_ema", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=0),
         inputs=[("uniform", (4, 3, 8, 8))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True), (E, W, False), (T, A, True)]),
    dict(id="qi8v2_act_delay", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=2),
         inputs=[("relu", (2, 5, 7, 7))], aux_init=[1.0],
         steps=[(T, W, True), (E, W, False), (T, W, True), (T, W, True)]),
    dict(id="qi8v2_act_unclipped", op_type="Quantization_int8_V2",   # |x| > aux -> codes beyond +-127
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=0),
         inputs=[("normal", (3, 4, 5, 5))], aux_init=[0.25],
         steps=[(E, W, False), (T, W, False)]),
    # ---- ClipGrad_Quantization_int8 (symbol/clip_grad_quantization_int8.py) ------------------------
    dict(id="clipgrad_weight_pertensor", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=True, is_weight_perchannel=False, delay_quant=0),
         inputs=[("normal", (8, 3, 3, 3))], aux_init=[1.0],
         steps=[(T, W, True), (E, W, False), (T, A, True)]),
    dict(id="clipgrad_weight_perchannel", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=True, is_weight_perchannel=True, delay_quant=0),
         inputs=[("normal", (6, 2, 3, 3))], aux_init=[1.0],
         steps=[(T, W, True), (E, W, False)]),
    dict(id="clipgrad_act", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=0),
         inputs=[("normal", (4, 3, 8, 8))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True), (E, W, False), (T, W, True)]),
    dict(id="clipgrad_act_delay", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=1),
         inputs=[("relu", (2, 6, 5, 5))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True), (T, W, True)]),
    dict(id="clipgrad_act_ties", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=0),
         inputs=[("ties", (2, 4, 6, 6))], aux_init=[1.0],
         steps=[(T, W, True)]),
    # ---- GDRQ_Fold_BN (symbol/fold_bn_v1_gdrq.py) ---------------------------------------------------
    dict(id="foldbn_pertensor", op_type="GDRQ_Fold_BN",
         attrs=_fold(is_weight_perchannel=False, delay_quant=0, num_filter=4, num_group=1, kernel=(3, 3),
                     stride=(1, 1), pad=(0, 0), quantize_flag=True),
         inputs=_fold_inputs(2, 3, 5, 5, 4, 1, 3, 1, 0), aux_init=[1.0, 1.0],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="foldbn_perchannel_depthwise", op_type="GDRQ_Fold_BN",
         attrs=_fold(is_weight_perchannel=True, delay_quant=0, num_filter=6, num_group=6, kernel=(3, 3),
                     stride=(2, 2), pad=(1, 1), quantize_flag=True),
         inputs=_fold_inputs(2, 6, 8, 8, 6, 6, 3, 2, 1), aux_init=[1.0, 1.0],
         steps=[(T, W, True), (T, A, True)]),
    dict(id="foldbn_perchannel_pointwise_delay", op_type="GDRQ_Fold_BN",
         attrs=_fold(is_weight_perchannel=True, delay_quant=1, num_filter=8, num_group=1, kernel=(1, 1),
                     stride=(1, 1), pad=(0, 0), quantize_flag=True),
         inputs=_fold_inputs(2, 4, 6, 6, 8, 1, 1, 1, 0), aux_init=[1.0, 1.0],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="foldbn_noquant", op_type="GDRQ_Fold_BN",
         attrs=_fold(is_weight_perchannel=False, delay_quant=0, num_filter=4, num_group=1, kernel=(3, 3),
                     stride=(1, 1), pad=(1, 1), quantize_flag=False),
         inputs=_fold_inputs(1, 3, 5, 5, 4, 1, 3, 1, 1), aux_init=[1.0, 1.0],
         steps=[(T, W, False), (E, W, False)]),
    dict(id="foldbn_eval_nameerror", op_type="GDRQ_Fold_BN",
         attrs=_fold(is_weight_perchannel=False, delay_quant=0, num_filter=4, num_group=1, kernel=(3, 3),
                     stride=(1, 1), pad=(0, 0), quantize_flag=True),
         inputs=_fold_inputs(1, 3, 5, 5, 4, 1, 3, 1, 0), aux_init=[1.0, 1.0],
         steps=[(E, W, False)]),
    # ---- GDRQ_PY / CLIP_RELU_PY (core/operator/GDRQ.py) --------------------------------------------
    dict(id="gdrq_weight_pertensor", op_type="GDRQ_PY",
         attrs=dict(nbits="4", group_size="-1", is_weight="True", lamda="0.001", delay_quant="0",
                    fix_alpha="False", ktimes="3"),
         inputs=[("normal", (8, 4, 3, 3))], aux_init=[0.5],
         steps=[(T, W, True), (T, A, True), (E, W, False)]),
    dict(id="gdrq_act_pertensor", op_type="GDRQ_PY",
         attrs=dict(nbits="8", group_size="-1", is_weight="False", lamda="0.001", delay_quant="1",
                    fix_alpha="False", ktimes="3"),
         inputs=[("normal", (4, 6, 5, 5))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True), (E, W, True)]),
    dict(id="gdrq_act_fixalpha", op_type="GDRQ_PY",
         attrs=dict(nbits="4", group_size="-1", is_weight="False", lamda="0.001", delay_quant="0",
                    fix_alpha="True", ktimes="3"),
         inputs=[("normal", (2, 6, 5, 5))], aux_init=[1.0],
         steps=[(T, W, True), (T, A, True)]),
    dict(id="gdrq_weight_grouped", op_type="GDRQ_PY",
         attrs=dict(nbits="4", group_size="2", is_weight="True", lamda="0.001", delay_quant="0",
                    fix_alpha="False", ktimes="2"),
         inputs=[("normal", (8, 3, 3, 3))], aux_init=[0.5],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="gdrq_act_grouped", op_type="GDRQ_PY",
         attrs=dict(nbits="5", group_size="3", is_weight="False", lamda="0.01", delay_quant="1",
                    fix_alpha="False", ktimes="3"),
         inputs=[("normal", (3, 6, 4, 5))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True), (T, A, True)]),
    dict(id="clip_relu", op_type="CLIP_RELU_PY", attrs=dict(nbits="8", threshold="8.0"),
         inputs=[("normal", (2, 4, 6, 6))], aux_init=[],
         steps=[(T, W, True), (T, A, True)], input_gain=5.0),
    # ---- second tier: core/operator/PACT.py, WNQ.py, QIL*.py -----------------------------------------
    dict(id="quant_ste", op_type="QUANT_STE_PY", attrs=dict(nbits="8"),
         inputs=[("normal", (4, 5, 3, 3))], aux_init=[],
         steps=[(T, W, True), (T, A, True)]),
    dict(id="pact", op_type="PACT_PY", attrs=dict(nbits="4"),
         inputs=[("relu", (2, 4, 6, 6)), ("const", (1,), 1.5)], aux_init=[],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="pact_v2", op_type="PACT_V2_PY", attrs=dict(nbits="4"),
         inputs=[("normal", (2, 4, 6, 6)), ("const", (1,), 1.25)], aux_init=[],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="dorefa", op_type="DoReFa_PY", attrs=dict(nbits="4"),
         inputs=[("normal", (6, 3, 3, 3))], aux_init=[],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="wnq_pertensor", op_type="WNQ_PY", attrs=dict(nbits="4", is_perchannel="False"),
         inputs=[("normal", (6, 3, 3, 3))], aux_init=[],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="wnq_perchannel", op_type="WNQ_PY", attrs=dict(nbits="4", is_perchannel="True"),
         inputs=[("normal", (6, 3, 3, 3))], aux_init=[],
         steps=[(T, W, True)]),
    dict(id="qil_v1", op_type="QIL_PY", attrs=dict(is_weight="True", fix_gamma="True", nbits="4"),
         inputs=[("unit", (4, 3, 3, 3)), ("const", (1,), 0.1), ("const", (1,), 0.9), ("const", (1,), 1.0)],
         aux_init=[], steps=[(T, W, True), (T, W, True)]),
    dict(id="qil_v1_clamp", op_type="QIL_PY", attrs=dict(is_weight="False", fix_gamma="True", nbits="3"),
         inputs=[("unit", (2, 3, 4, 4)), ("const", (1,), -0.2), ("const", (1,), 1.3), ("const", (1,), 1.0)],
         aux_init=[], steps=[(T, W, True)]),
    dict(id="qil_v2", op_type="QIL_V2_PY", attrs=dict(is_weight="True", fix_gamma="True", nbits="4"),
         inputs=[("unit", (4, 3, 3, 3)), ("const", (1,), 0.5), ("const", (1,), 0.4), ("const", (1,), 1.0)],
         aux_init=[], steps=[(T, W, True)]),
    dict(id="qil_v3", op_type="QIL_V3_PY", attrs=dict(is_weight="True", fix_gamma="True", nbits="4"),
         inputs=[("unit", (4, 3, 3, 3)), ("const", (1,), -2.0), ("const", (1,), -0.3), ("const", (1,), 1.0)],
         aux_init=[], steps=[(T, W, True)]),
    # ---- second batch: FC shapes, eval-first / add / delay combinations, grouped convolution, other bit widths --------
    dict(id="qi8v2_fc_weight_perchannel", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=True, is_weight_perchannel=True, delay_quant=0),
         inputs=[("normal", (10, 33))], aux_init=[1.0],
         steps=[(T, W, True), (T, A, True), (E, W, False)]),
    dict(id="qi8v2_fc_act", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=0, ema_decay=0.9),
         inputs=[("relu", (7, 33))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True), (T, W, True), (E, A, False)]),
    dict(id="qi8v2_weight_delay", op_type="Quantization_int8_V2",
         attrs=_qi8(is_weight=True, is_weight_perchannel=False, delay_quant=1),
         inputs=[("normal", (4, 2, 3, 3))], aux_init=[1.0],
         steps=[(T, W, True), (E, W, False), (T, W, True)]),
    dict(id="clipgrad_fc_weight_eval_first", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=True, is_weight_perchannel=False, delay_quant=0),
         inputs=[("normal", (12, 17))], aux_init=[0.75],
         steps=[(E, W, False), (T, W, True), (E, W, False)]),
    dict(id="clipgrad_weight_perchannel_eval_first", op_type="ClipGrad_Quantization_int8",
         attrs=_qi8(is_weight=True, is_weight_perchannel=True, delay_quant=0),
         inputs=[("normal", (5, 3, 3, 3))], aux_init=[0.5],
         steps=[(E, W, False), (T, A, True)]),
    dict(id="clipgrad_act_add_req", op_type="ClipGrad_Quantization_int8",   # the op writes with [:]= whatever req says
         attrs=_qi8(is_weight=False, is_weight_perchannel=False, delay_quant=0),
         inputs=[("normal", (3, 5, 4, 4))], aux_init=[1.0],
         steps=[(T, A, True), (T, A, True), (E, A, True)]),
    dict(id="gdrq_weight_fixalpha", op_type="GDRQ_PY",
         attrs=dict(nbits="8", group_size="-1", is_weight="True", lamda="0.001", delay_quant="0",
                    fix_alpha="True", ktimes="3"),
         inputs=[("normal", (6, 4, 3, 3))], aux_init=[0.3],
         steps=[(T, W, True), (E, A, True)]),
    dict(id="gdrq_act_delay2_add", op_type="GDRQ_PY",
         attrs=dict(nbits="6", group_size="-1", is_weight="False", lamda="0.05", delay_quant="2",
                    fix_alpha="False", ktimes="2"),
         inputs=[("relu", (3, 4, 6, 6))], aux_init=[2.0],
         steps=[(T, W, True), (T, A, True), (T, W, True), (E, W, True)]),
    dict(id="gdrq_weight_grouped_delay", op_type="GDRQ_PY",
         attrs=dict(nbits="3", group_size="4", is_weight="True", lamda="0.001", delay_quant="1",
                    fix_alpha="False", ktimes="3"),
         inputs=[("normal", (8, 2, 3, 3))], aux_init=[0.5],
         steps=[(T, W, True), (T, W, True), (E, W, False)]),
    dict(id="gdrq_fc_act_grouped", op_type="GDRQ_PY",
         attrs=dict(nbits="8", group_size="8", is_weight="False", lamda="0.001", delay_quant="0",
                    fix_alpha="False", ktimes="3"),
         inputs=[("normal", (5, 24))], aux_init=[1.0],
         steps=[(T, W, True), (T, W, True)]),
    dict(id="foldbn_grouped_conv_stride2", op_type="GDRQ_Fold_BN",
         attrs=_fold(is_weight_perchannel=True, delay_quant=0, num_filter=6, num_group=2, kernel=(3, 3),
                     stride=(2, 2), pad=(1, 1), quantize_flag=True),
         inputs=_fold_inputs(2, 4, 9, 9, 6, 2, 3, 2, 1), aux_init=[1.0, 1.0],
         steps=[(T, W, True), (T, W, True), (T, W, True)]),
    dict(id="clip_relu_4bit", op_type="CLIP_RELU_PY", attrs=dict(nbits="4", threshold="6.0"),
         inputs=[("normal", (3, 5, 5))], aux_init=[],
         steps=[(T, W, True), (E, A, True)], input_gain=4.0),
    dict(id="quant_ste_4bit", op_type="QUANT_STE_PY", attrs=dict(nbits="4"),
         inputs=[("uniform", (9, 13))], aux_init=[],
         steps=[(T, W, True), (E, W, True)]),
    dict(id="pact_8bit", op_type="PACT_PY", attrs=dict(nbits="8"),
         inputs=[("relu", (4, 3, 5, 5)), ("const", (1,), 0.8)], aux_init=[],
         steps=[(T, W, True)]),
    dict(id="dorefa_2bit", op_type="DoReFa_PY", attrs=dict(nbits="2"),
         inputs=[("normal", (5, 4, 3, 3))], aux_init=[],
         steps=[(T, W, True)]),
    dict(id="wnq_8bit_perchannel_fc", op_type="WNQ_PY", attrs=dict(nbits="8", is_perchannel="True"),
         inputs=[("normal", (7, 19))], aux_init=[],
         steps=[(T, W, True), (T, W, True)]),
]

CASE_BY_ID = {c["id"]: c for c in CASES}


def case_seed(case):
    return 5 + sum(ord(ch) for ch in case["id"])  # echoes np.random.seed(5), fold_bn_v1_gdrq.py:330


def build_step_inputs(case, rng, step):
    arrs = []
    for spec in case["inputs"]:
        if spec[0] == "const":
            arrs.append(np.full(spec[1], spec[2], dtype=F))
        else:
            a = make_input(spec[0], spec[1], rng, step)
            if "input_gain" in case:
                a = (a * F(case["input_gain"])).astype(F)
            arrs.append(a)
    return arrs
