"""Quantization-node inventories of the reference's int8 networks (shapes only), derived by walking the
topologies in symbol/resnet_int8.py, symbol/mobilenet_int8*.py, symbol/resnext.py and symbol/simple.py.
They define the benchmark workloads (BASELINE.json configs) and the shapes the parity tests use.

Each entry: (name, kind, shape) with kind "act" (the tensor entering a conv / FC) or "weight".
Node names follow the wrappers' contract: ``<layer>_data`` / ``<layer>_weight`` (symbol/int8_api.py:31-36).
"""


def _conv_out(hw, k, s, p):
    return (hw + 2 * p - k) // s + 1


class _Walk(object):
    def __init__(self, batch):
        self.batch = batch
        self.nodes = []

    def conv(self, name, cin, hw, cout, k, s, p, group=1):
        self.nodes.append((name + "_data", "act", (self.batch, cin, hw, hw)))
        self.nodes.append((name + "_weight", "weight", (cout, cin // group, k, k)))
        return _conv_out(hw, k, s, p)

    def fc(self, name, cin, cout):
        self.nodes.append((name + "_data", "act", (self.batch, cin)))
        self.nodes.append((name + "_weight", "weight", (cout, cin)))


def resnet50_nodes(batch=256, units=(3, 4, 6, 3), filters=(64, 256, 512, 1024, 2048), num_classes=1000):
    """symbol/resnet_int8.py:69-131 (imagenet stem, bottleneck units :16-44): 54 act + 54 weight nodes."""
    w = _Walk(batch)
    hw = w.conv("conv0", 3, 224, filters[0], 7, 2, 3)
    hw = _conv_out(hw, 3, 2, 1)                                    # max pool (:95)
    cin = filters[0]
    for i, n_units in enumerate(units):
        cout = filters[i + 1]
        stride = 1 if i == 0 else 2
        for j in range(n_units):
            name = "stage%d_unit%d" % (i + 1, j + 1)
            s = stride if j == 0 else 1
            mid = int(cout * 0.25)
            w.conv(name + "_conv1", cin, hw, mid, 1, 1, 0)
            hw2 = w.conv(name + "_conv2", mid, hw, mid, 3, s, 1)
            w.conv(name + "_conv3", mid, hw2, cout, 1, 1, 0)
            if j == 0:
                w.conv(name + "_sc", cin, hw, cout, 1, s, 0)     # shortcut quantises act1 again (:38)
            hw, cin = hw2, cout
    w.fc("fc1", cin, num_classes)
    return w.nodes


def resnext101_nodes(batch=256, units=(3, 4, 23, 3), filters=(64, 256, 512, 1024, 2048), num_group=32,
                     num_classes=1000):
    """symbol/resnext.py:72-103 with every conv/FC wrapped by clipgrad_quant_conv/fc (symbol/int8_api.py:19-71):
    105 act + 105 weight nodes."""
    w = _Walk(batch)
    hw = w.conv("conv0", 3, 224, filters[0], 7, 2, 3)
    hw = _conv_out(hw, 3, 2, 1)
    cin = filters[0]
    for i, n_units in enumerate(units):
        cout = filters[i + 1]
        stride = 1 if i == 0 else 2
        for j in range(n_units):
            name = "stage%d_unit%d" % (i + 1, j + 1)
            s = stride if j == 0 else 1
            mid = int(cout * 0.5)
            w.conv(name + "_conv1", cin, hw, mid, 1, 1, 0)
            hw2 = w.conv(name + "_conv2", mid, hw, mid, 3, s, 1, group=num_group)
            w.conv(name + "_conv3", mid, hw2, cout, 1, 1, 0)
            if j == 0:
                w.conv(name + "_sc", cin, hw, cout, 1, s, 0)
            hw, cin = hw2, cout
    w.fc("fc1", cin, num_classes)
    return w.nodes


def mobilenet_v1_nodes(batch=256, num_classes=1000):
    """MobileNet-v1 (alpha = 1) as in symbol/mobilenet_int8*.py: conv 3x3/2, 13 depthwise + pointwise pairs, FC:
    28 act + 28 weight nodes."""
    w = _Walk(batch)
    hw = w.conv("conv1", 3, 224, 32, 3, 2, 1)
    cin = 32
    cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2)] + [(512, 1)] * 5 + [(1024, 2), (1024, 1)]
    for i, (cout, s) in enumerate(cfg):
        hw = w.conv("conv%d_dw" % (i + 2), cin, hw, cin, 3, s, 1, group=cin)
        w.conv("conv%d_pw" % (i + 2), cin, hw, cout, 1, 1, 0)
        cin = cout
    w.fc("fc", cin, num_classes)
    return w.nodes


def simple_nodes(batch=32, num_classes=10):
    """symbol/simple.py:10-18 (two conv-bn-relu stages, global pool, FC) at CIFAR 32x32: 3 + 3 nodes."""
    w = _Walk(batch)
    hw = 32
    hw = w.conv("stage1_conv", 3, hw, 8, 3, 2, 1)
    hw = w.conv("stage2_conv", 8, hw, 8, 3, 2, 1)
    w.fc("fc1", 8, num_classes)
    return w.nodes


WORKLOADS = {
    "simple_cifar": (simple_nodes, 32, "Quantization_int8_V2"),
    "resnet50_int8": (resnet50_nodes, 256, "Quantization_int8_V2"),
    "mobilenet_v1_foldbn": (mobilenet_v1_nodes, 256, "GDRQ_Fold_BN"),
    "mobilenet_v1_gdrq": (mobilenet_v1_nodes, 256, "GDRQ_PY"),
    "resnext101_clipgrad": (resnext101_nodes, 256, "ClipGrad_Quantization_int8"),
}


def numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def summary(nodes):
    acts = [numel(s) for _, k, s in nodes if k == "act"]
    wts = [numel(s) for _, k, s in nodes if k == "weight"]
    return dict(act_nodes=len(acts), weight_nodes=len(wts), act_elems=sum(acts), weight_elems=sum(wts))
