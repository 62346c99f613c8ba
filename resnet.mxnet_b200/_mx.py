"""Lazy access to the ``mxnet`` module for the symbol-level builders (quant_conv, clipgrad_quant_*, GDRQ_fold_bn,
graph_optimize's symbol path).  The operators themselves never need MXNet; the builders construct ``mx.sym`` graphs and
do.  Resolution happens at call time, so importing this package never requires MXNet, and a test may install
``oracle/mxshim`` (a symbolic graph recorder, test infrastructure) as ``mxnet`` to execute the builders side by side with
the reference's."""
import sys


class LazyMx(object):
    def module(self):
        m = sys.modules.get("mxnet")
        if m is None:
            try:
                import mxnet as m   # noqa: F401
            except Exception:
                raise RuntimeError("this function builds mx.sym graphs and needs MXNet; under torch use "
                                   "b200quant.harness (QuantConv2d / QuantLinear / QuantDeconv2d / QuantAdd / "
                                   "QuantConcat / QuantData: same node and parameter names) or "
                                   "b200quant.graph_optimize.attach_quantize_node on a torch model")
        if not hasattr(m, "sym"):
            raise RuntimeError("the installed mxnet stand-in has no symbolic API")
        return m

    def __getattr__(self, item):
        return getattr(self.module(), item)


mx = LazyMx()
