"""Runs the HBM-bound InstanceNorm / AdaIN kernels at the residual-block shape (B=32, 64x64x256) for
`ncu --set full -k regex:norm_act|nc_reduce|epi_stats`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import lib as L  # noqa: E402
from msig_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
ops.ensure_init(dev)
B, h, c = 32, 64, 256
x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
dy = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
res = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
y = torch.empty_like(x)
st = ops.in_stats(x)                                  # nc_reduce_kernel<0>
ops.norm_act_fwd(x, st, L.ACT_RELU, out=y)            # norm_act_fwd_kernel
ops.norm_act_fwd(x, st, L.ACT_NONE, residual=res, out=y)
ops.norm_act_bwd(dy, x, st, L.ACT_RELU, out=y)        # nc_reduce_kernel<1> + norm_act_bwd_kernel
es = ops.epi_stats(B, h, h, c, dev)
es.buf.zero_()
ops.norm_bwd_from(es, dy, x, st, out=y)               # epi_stats_finalize_kernel<1> + norm_act_bwd_kernel (fused-reduction form)
# the pad-fused pair around the generator's final conv: [B,256,256,64]
x64 = torch.randn(B, 256, 256, 64, device=dev).to(torch.bfloat16)
st64 = ops.in_stats(x64)
yp = ops.norm_act_fwd_pad(x64, st64, L.ACT_RELU, 3)   # norm_act_fwd_kernel<PADOUT>
dyp = torch.randn(B, 262, 262, 64, device=dev).to(torch.bfloat16)
ops.norm_act_bwd_pad(dyp, x64, st64, L.ACT_RELU, 3)   # nc_reduce_kernel<1, FOLD> + norm_act_bwd_kernel<FOLD>
torch.cuda.synchronize()
print("ok")
