"""CPU: the plain-C restatement (oracle/msig_oracle.c, direct loops, double accumulation) against the
torch-based oracle's primitives on small shapes — the oracle's arithmetic is pinned by code that
shares nothing with PyTorch."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

ODIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def C():
    subprocess.run(["make", "-C", ODIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(ODIR, "_build", "liboracle.so"))
    lib.orc_l1.restype = ctypes.c_float
    lib.orc_mse.restype = ctypes.c_float
    return lib


def P(a):
    return a.ctypes.data_as(FP) if a is not None else None


def A(t):
    return np.ascontiguousarray(t.detach().numpy().astype(np.float32))


def close(a, b, tol=2e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("k,stride,pad,reflect", [(3, 1, 1, 0), (4, 2, 1, 0), (7, 1, 3, 1), (1, 1, 0, 0)])
def test_conv2d_fwd_bwd(C, k, stride, pad, reflect):
    g = torch.Generator().manual_seed(k)
    n, cin, h, w, cout = 2, 3, 10, 12, 4
    x = torch.randn(n, cin, h, w, generator=g, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, generator=g, requires_grad=True)
    b = torch.randn(cout, generator=g, requires_grad=True)
    if reflect:
        y = O.reflect_conv7(x, wt, b)
    else:
        y = F.conv2d(x, wt, b, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g)
    (y * dy).sum().backward()
    oh, ow = y.shape[2], y.shape[3]
    yc = np.zeros(tuple(y.shape), np.float32)
    C.orc_conv2d(P(A(x)), P(A(wt)), P(A(b)), P(yc), n, cin, h, w, cout, k, stride, pad, pad, oh, ow, reflect)
    assert close(yc, A(y))
    dx, dw, db = np.zeros(tuple(x.shape), np.float32), np.zeros(tuple(wt.shape), np.float32), np.zeros(cout, np.float32)
    C.orc_conv2d_bwd(P(A(x)), P(A(wt)), P(A(dy)), P(dx), P(dw), P(db), n, cin, h, w, cout, k, stride, pad, pad, oh, ow, reflect)
    assert close(dx, A(x.grad)) and close(dw, A(wt.grad)) and close(db, A(b.grad))


def test_discriminator_head_asymmetric_padding(C):
    """ZeroPad2d((1,0,1,0)) + Conv2d(4, padding=1) == top/left pad 2 (model.py:182-183)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 5, 6, 6, generator=g)
    wt = torch.randn(1, 5, 4, 4, generator=g)
    y = F.conv2d(F.pad(x, (1, 0, 1, 0)), wt, None, padding=1)
    yc = np.zeros(tuple(y.shape), np.float32)
    C.orc_conv2d(P(A(x)), P(A(wt)), None, P(yc), 1, 5, 6, 6, 1, 4, 1, 2, 2, y.shape[2], y.shape[3], 0)
    assert y.shape[2:] == (6, 6) and close(yc, A(y))


def test_conv_transpose(C):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 5, 6, generator=g)
    wt = torch.randn(3, 4, 4, 4, generator=g)
    b = torch.randn(4, generator=g)
    y = F.conv_transpose2d(x, wt, b, stride=2, padding=1)
    yc = np.zeros(tuple(y.shape), np.float32)
    C.orc_conv_transpose2d(P(A(x)), P(A(wt)), P(A(b)), P(yc), 2, 3, 5, 6, 4)
    assert close(yc, A(y))


def test_adain_fwd_bwd(C):
    g = torch.Generator().manual_seed(2)
    n, c, h, w = 2, 6, 5, 7
    x = (torch.randn(n, c, h, w, generator=g) * 2 + 1).requires_grad_(True)
    style = torch.randn(n, 8, generator=g)
    lw = torch.randn(2 * c, 8, generator=g)
    lb = torch.randn(2 * c, generator=g)
    gb = (style @ lw.t() + lb).detach().requires_grad_(True)
    y = gb[:, :c].reshape(n, c, 1, 1) * O.instance_norm(x) + gb[:, c:].reshape(n, c, 1, 1)
    assert torch.allclose(y, O.adain(x, style, lw, lb), atol=1e-6)
    dy = torch.randn(y.shape, generator=g)
    (y * dy).sum().backward()
    gam, bet = A(gb[:, :c]), A(gb[:, c:])
    yc = np.zeros(tuple(y.shape), np.float32)
    C.orc_adain(P(A(x)), P(gam), P(bet), P(yc), n, c, h * w, ctypes.c_float(1e-5))
    assert close(yc, A(y))
    dx, dg, dbt = np.zeros(tuple(x.shape), np.float32), np.zeros((n, c), np.float32), np.zeros((n, c), np.float32)
    C.orc_adain_bwd(P(A(x)), P(gam), P(A(dy)), P(dx), P(dg), P(dbt), n, c, h * w, ctypes.c_float(1e-5))
    assert close(dx, A(x.grad), 1e-4) and close(dg, A(gb.grad[:, :c]), 1e-4) and close(dbt, A(gb.grad[:, c:]), 1e-4)


def test_gram_losses_pools(C):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 6, 6, generator=g)
    gc = np.zeros((8, 8), np.float32)
    C.orc_gram(P(A(x)), P(gc), 8, 36)
    assert close(gc, A(O.gram(x)))
    a, b = torch.randn(50, generator=g), torch.randn(50, generator=g)
    assert abs(C.orc_l1(P(A(a)), P(A(b)), ctypes.c_size_t(50)) - float(F.l1_loss(a, b))) < 1e-6
    assert abs(C.orc_mse(P(A(a)), P(A(b)), ctypes.c_size_t(50)) - float(F.mse_loss(a, b))) < 1e-6
    mp = np.zeros((2, 4, 3, 3), np.float32)
    C.orc_maxpool2(P(A(x)), P(mp), 8, 6, 6)
    assert close(mp, A(F.max_pool2d(x, 2)), 0)
    ap = np.zeros((2, 4), np.float32)
    C.orc_avgpool(P(A(x)), P(ap), 8, 36)
    assert close(ap, A(x.mean((2, 3))))
