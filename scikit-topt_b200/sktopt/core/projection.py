"""Smoothed Heaviside projection (reference ``core/projection.py``).

rho_hat = (tanh(beta eta) + tanh(beta (rho - eta))) / (tanh(beta eta) +
tanh(beta (1 - eta)) + 1e-12); note the +1e-12 in the denominator (:74,:115).
NumPy inputs are evaluated with NumPy (API edge / small arrays); CUDA tensors
go through the ``sktb_heaviside`` kernel."""
import numpy as np


def _is_dev(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _denominator(beta, eta):
    return np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta)) + 1e-12


def heaviside_projection(rho, beta, eta=0.5):
    return (np.tanh(beta * eta) + np.tanh(beta * (rho - eta))) / _denominator(beta, eta)


def heaviside_projection_derivative(rho, beta, eta=0.5):
    return beta * (1.0 - np.tanh(beta * (rho - eta)) ** 2) / _denominator(beta, eta)


def heaviside_projection_inplace(rho, beta, eta=0.5, out=None):
    """Writes H_beta(rho) to ``out`` (allocated when None) and returns it."""
    if _is_dev(rho):
        import torch
        from sktopt._b200 import device as dev
        if out is None:
            out = torch.empty_like(rho)
        dev.heaviside(rho, beta, eta, out=out, dH=None)
        return out
    if out is None:
        out = np.empty_like(rho)
    np.subtract(rho, eta, out=out)
    out *= beta
    np.tanh(out, out=out)
    out += np.tanh(beta * eta)
    out /= _denominator(beta, eta)
    return out


def heaviside_projection_derivative_inplace(rho, beta, eta=0.5, out=None):
    """Writes dH/drho = beta sech^2(beta (rho - eta)) / denom to ``out``."""
    if _is_dev(rho):
        import torch
        from sktopt._b200 import device as dev
        if out is None:
            out = torch.empty_like(rho)
        dev.heaviside(rho, beta, eta, out=None, dH=out)
        return out
    if out is None:
        out = np.empty_like(rho)
    np.subtract(rho, eta, out=out)
    out *= beta
    np.cosh(out, out=out)
    np.square(out, out=out)
    np.reciprocal(out, out=out)
    out *= beta
    out /= _denominator(beta, eta)
    return out
