"""Helpers for the -m gpu tests: thin callers of the C ABI on torch tensors."""
import torch

from tpxl_b200 import _lib


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dev():
    return torch.device("cuda:0")


def st():
    return _lib.stream_ptr()


def linear(A, W, bias=None, act=0, post_scale=1.0, tile_n=0):
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=torch.float16, device=A.device)
    _lib.check(_lib.lib().tpx_linear(A.data_ptr(), A.stride(0), W.data_ptr(), _lib.ptr(bias), out.data_ptr(), N, M, N, K, act, post_scale, tile_n, st()), "tpx_linear")
    return out


def linear_ref(A, W, bias=None, act=0, post_scale=1.0):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias.float()
    y = y.half().float()
    if act:
        y = torch.nn.functional.gelu(y, approximate="tanh")
    elif post_scale != 1.0:
        y = y * post_scale
    return y.half()
