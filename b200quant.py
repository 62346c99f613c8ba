"""Import alias: ``import b200quant`` loads the package that lives in ``resnet.mxnet_b200/`` (a directory name
the task fixes but Python cannot import directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resnet.mxnet_b200")
_spec = importlib.util.spec_from_file_location("b200quant", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200quant"] = _mod
_spec.loader.exec_module(_mod)
