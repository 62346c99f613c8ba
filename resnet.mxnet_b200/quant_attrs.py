"""Drop-in for ``config/quant_attrs.py``: the schema of ``config.quantize_setting`` entries.

The reference file is one docstring of examples (quant_attrs.py:1-81); here the same schema is data, with a validator,
so a mis-typed attribute fails at graph-construction time instead of inside an operator.  Every attribute value is a
STRING, exactly as MXNet hands operator attributes to a ``CustomOpProp``.

    {"weight": {"quantize_op_name": ..., "init_value": ..., "attrs": {...}}, "act": {...}}
"""

# quantize_op_name -> {attr: example}, as documented in config/quant_attrs.py
SCHEMA = {
    "Quantization_int8": {   # quant_attrs.py:5-35 (fork C++ op; mapped to Quantization_int8_V2 by graph_optimize)
        "nbits": "4", "quant_mode": "minmax", "is_weight": "True", "is_weight_perchannel": "False",
        "delay_quant": "0", "ema_decay": "0.99", "grad_mode": "ste", "fix_act_scale": "False"},
    "GDRQ": {                # quant_attrs.py:37-64 (GDRQ_PY; GDRQ_CXX is the fork's C++ twin)
        "nbits": "4", "fix_alpha": "False", "group_size": "-1", "is_weight": "True", "lamda": "0.001",
        "delay_quant": "0", "ktimes": "3"},
    "DoReFa_PY": {"nbits": "4"},   # quant_attrs.py:66-80
    "PACT": {"nbits": "4"},
    "QIL": {"is_weight": "False", "fix_gamma": "True", "nbits": "4"},
    "WNQ": {"nbits": "4", "is_perchannel": "False"},
}
SCHEMA["GDRQ_CXX"] = SCHEMA["GDRQ"]
SCHEMA["DoReFa_CXX"] = SCHEMA["DoReFa_PY"]
SCHEMA["PACT_CXX"] = SCHEMA["PACT"]

DEFAULT_INIT = {"Quantization_int8": 0, "QIL": 1.0, "PACT": 8.0, "PACT_CXX": 8.0, "GDRQ": 1.0, "GDRQ_CXX": 1.0}
# core/graph_optimize.py:166,170,181,185,189,193


def validate(setting):
    """Check one ``weight`` / ``act`` entry; returns it unchanged.  Raises ValueError with the offending key."""
    get = setting.get if isinstance(setting, dict) else (lambda k, d=None: getattr(setting, k, d))
    name = get("quantize_op_name")
    if name not in SCHEMA:
        raise ValueError("unknown quantize_op_name %r (known: %s)" % (name, ", ".join(sorted(SCHEMA))))
    attrs = get("attrs", {}) or {}
    for k, v in attrs.items():
        if k not in SCHEMA[name]:
            raise ValueError("%s does not take attribute %r (takes: %s)" % (name, k, ", ".join(sorted(SCHEMA[name]))))
        if not isinstance(v, str):
            raise ValueError("attribute %s=%r must be a string (MXNet passes operator attributes as strings)" % (k, v))
    return setting


def validate_quantize_setting(quantize_setting):
    for side in ("weight", "act"):
        if side not in quantize_setting:
            raise ValueError("quantize_setting needs a %r entry" % side)
        validate(quantize_setting[side])
    return quantize_setting
