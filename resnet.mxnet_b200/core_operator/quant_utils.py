"""Drop-in for ``core/operator/quant_utils.py`` (print_info, quantizeK)."""


def print_info(auto_grad, cal_grad, name):
    to_np = lambda a: a.asnumpy() if hasattr(a, "asnumpy") else a.detach().cpu().numpy()
    print("{} autograd:\n{}\n cal grad:\n{}".format(name, to_np(auto_grad), to_np(cal_grad)))
    print("{} autograd - cal_grad:\n{}".format(name, to_np(auto_grad) - to_np(cal_grad)))


def quantizeK(data, nbits):
    """round(L*x)/L with L = 2^nbits - 1 (quant_utils.py:11-13); a tiny helper, evaluated with the framework's
    own ops (only used outside the hot path)."""
    L = 2 ** nbits - 1
    try:
        import torch
        if isinstance(data, torch.Tensor):
            t = data * L
            r = torch.trunc(t)
            r = r + torch.where((t - r).abs() >= 0.5, torch.sign(t), torch.zeros_like(t))   # roundf
            return r / L
    except ImportError:  # pragma: no cover
        pass
    import mxnet as mx  # pragma: no cover
    return mx.nd.round(L * data) / L
