"""oracle/cppn.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Functional fp32 (or fp64) torch-CPU restatement of the reference MLP
(/root/reference/model/CPPN.py:166-222, construction :96-131) restricted to the
configuration the driver uses: relu, no skip, no view directions, pos_enc in
{'none', 'fourier', 'barf'}.  Parameters are passed as a dict with the reference's own
state-dict key names (``early_pts_layers.{0,2,..}.{weight,bias}``, ``output_linear.0.*``,
``fourier_coefficients``), so a reference checkpoint's ``['model']`` dict works as is.
Pinned against the reference class by tests/golden/make_golden.py.
"""
import numpy as np
import torch


def n_hidden_linears(params) -> int:
    return len([k for k in params if k.startswith("early_pts_layers.") and k.endswith(".weight")])


def fourier_features(x, coeff, basis):
    """CPPN.pos_enc + fourier_pos_enc, /root/reference/model/CPPN.py:207-222."""
    v = torch.cat(basis * [x], dim=-1)                 # index j <-> coord j % 3, freq j // 3
    a = 2 * np.pi * v * coeff                          # ((2*pi) * v) * coeff, left to right
    return torch.cat([x, torch.sin(a), torch.cos(a)], dim=-1)


def barf_weights(barf_alpha, basis, n_channels=3):
    """CPPN.barf_coefficients, /root/reference/model/CPPN.py:242-259, quirks included (3.1415, alpha - k + 1)."""
    k_values = torch.repeat_interleave(torch.arange(0., basis), n_channels)
    w = []
    for k in k_values:
        barf_k = barf_alpha - (k + 1)
        if barf_k < 0:
            w.append(0)
        elif barf_k < 1:
            w.append((1 - torch.cos((barf_alpha - k + 1) * 3.1415)) / 2)
        else:
            w.append(1)
    return torch.Tensor(w)


def barf_features(x, weights, basis):
    """CPPN.pos_enc + barf_pos_enc, /root/reference/model/CPPN.py:207-234: freq = float32(2^k * pi), one multiply."""
    k_values = torch.repeat_interleave(torch.arange(0., basis), x.shape[-1])
    freq = torch.Tensor(2 ** k_values * np.pi)
    v = torch.cat(basis * [x], dim=-1)
    a = freq * v
    return torch.cat([x, weights * torch.sin(a), weights * torch.cos(a)], dim=-1)


def cppn_forward(params, x, pos_enc="none", basis=5):
    """x[S,3] -> raw logit [S,1].  /root/reference/model/CPPN.py:166-205."""
    h = x
    if pos_enc == "fourier" and basis > 0:
        h = fourier_features(x, params["fourier_coefficients"], basis)
    elif pos_enc == "barf" and basis > 0:
        h = barf_features(x, params["barf_weights"], basis)
    elif pos_enc != "none":
        raise ValueError("oracle restates pos_enc in {'none','fourier','barf'} only")
    for i in range(n_hidden_linears(params)):
        w = params[f"early_pts_layers.{2 * i}.weight"]
        b = params[f"early_pts_layers.{2 * i}.bias"]
        h = torch.relu(torch.nn.functional.linear(h, w, b))
    return torch.nn.functional.linear(h, params["output_linear.0.weight"], params["output_linear.0.bias"])


def init_params(num_layers, width, pos_enc="none", basis=5, sigma=5.0, seed=0, dtype=torch.float32):
    """Random parameters with torch.nn.Linear's default init, keyed like the reference state dict."""
    g = torch.Generator().manual_seed(seed)
    d_in = 3 + (3 * 2 * basis if pos_enc != "none" else 0)
    p = {}
    if pos_enc == "fourier":
        p["fourier_coefficients"] = (torch.randn(3 * basis, generator=g) * sigma).to(dtype)

    def lin(i, o):
        bound = 1.0 / np.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        b = (torch.rand(o, generator=g) * 2 - 1) * bound
        return w.to(dtype), b.to(dtype)

    dims = [d_in] + [width] * (num_layers + 1)
    for li in range(num_layers + 1):
        w, b = lin(dims[li], dims[li + 1])
        p[f"early_pts_layers.{2 * li}.weight"] = w
        p[f"early_pts_layers.{2 * li}.bias"] = b
    w, b = lin(width, 1)
    p["output_linear.0.weight"] = w
    p["output_linear.0.bias"] = b
    return p
