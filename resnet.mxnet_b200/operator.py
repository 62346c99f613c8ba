"""The MXNet CustomOp protocol the reference operators are written against.

When ``mxnet`` is importable the real ``mx.operator.CustomOp`` / ``CustomOpProp`` / ``register`` are used, so
``mx.sym.Custom(..., op_type="Quantization_int8_V2")`` instantiates these classes exactly as it instantiates
the reference's (symbol/quant_ops.py:44, :89-94).  Otherwise an identical local protocol is provided and the
classes are driven by ``harness.Custom`` (a torch autograd wrapper).
"""
import ast

try:  # pragma: no cover - mxnet is not installable in the build image
    import mxnet as _mx
    if getattr(_mx, "__is_b2q_shim__", False):
        raise ImportError("test shim, not MXNet")
    HAVE_MXNET = True
except Exception:
    _mx = None
    HAVE_MXNET = False

REGISTRY = {}

if HAVE_MXNET:  # pragma: no cover
    CustomOp = _mx.operator.CustomOp
    CustomOpProp = _mx.operator.CustomOpProp

    def _synced(method):
        """MXNet's engine cannot see the library's launches: do not hand control back before they are done."""
        import functools

        @functools.wraps(method)
        def call(self, *args, **kwargs):
            try:
                return method(self, *args, **kwargs)
            finally:
                from .dlpack import sync_foreign
                sync_foreign()
        call.__b2q_syncs__ = True
        return call

    def register(reg_name):
        mx_deco = _mx.operator.register(reg_name)

        def deco(prop_cls):
            REGISTRY[reg_name] = prop_cls
            make = prop_cls.create_operator

            def create_operator(self, ctx, in_shapes, in_dtypes):
                op = make(self, ctx, in_shapes, in_dtypes)
                cls = type(op)
                if not getattr(cls.forward, "__b2q_syncs__", False):
                    cls.forward = _synced(cls.forward)
                    cls.backward = _synced(cls.backward)
                return op
            prop_cls.create_operator = create_operator
            return mx_deco(prop_cls)
        return deco
else:
    class CustomOp(object):
        """Base class for operators [upstream python/mxnet/operator.py: class CustomOp]."""

        def forward(self, is_train, req, in_data, out_data, aux):
            pass

        def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
            pass

        def assign(self, dst, req, src):
            """dst (req) src with req in {'null','write','inplace','add'}; a CUDA copy / accumulate kernel."""
            from . import _kernels
            _kernels.assign(dst, req, src)

    class CustomOpProp(object):
        """Base class for operator properties [upstream python/mxnet/operator.py: class CustomOpProp]."""

        def __init__(self, need_top_grad=True):
            self.need_top_grad_ = need_top_grad

        def infer_shape(self, in_shape):
            return in_shape, [in_shape[0]] * len(self.list_outputs()), []

        def infer_type(self, in_type):
            return in_type, [in_type[0]] * len(self.list_outputs()), \
                [in_type[0]] * len(self.list_auxiliary_states())

        def list_outputs(self):
            return ["output"]

        def list_arguments(self):
            return ["data"]

        def list_auxiliary_states(self):
            return []

        def declare_backward_dependency(self, out_grad, in_data, out_data):
            deps = []
            if self.need_top_grad_:
                deps.extend(out_grad)
            deps.extend(in_data)
            deps.extend(out_data)
            return deps

        def create_operator(self, ctx, in_shapes, in_dtypes):
            return CustomOp()

    def register(reg_name):
        """Register a CustomOpProp subclass under ``op_type`` [upstream mx.operator.register]."""
        def deco(prop_cls):
            REGISTRY[reg_name] = prop_cls
            return prop_cls
        return deco


def get_prop(op_type):
    if op_type not in REGISTRY:
        from . import ops  # noqa: F401  (registers everything)
    return REGISTRY[op_type]


# -- attribute parsing: MXNet hands every attribute to the Prop as a *string*  ----------
def py_literal(v):
    """What the reference's ``eval(attr)`` yields for the literals it is given ("True", "(3, 3)", "8")."""
    if isinstance(v, str):
        return ast.literal_eval(v.strip())
    return v


def py_bool(v):
    return bool(py_literal(v))
