#!/usr/bin/env python
"""The kernels behind bench.py's roofline objects, one launch each after a warm-up, for `ncu --set full`:
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/kernels \
      python tools/profile_kernels.py
Launch order inside the profiled range: chan_stats, gn_apply2 (two-pass form), gn_apply2 (in-step form), gemm<256> FF1/GEGLU, gemm<128> (N=640 out-proj with
residual), gemm M8192 N1280 K1280 + residual, conv3x3 (8,64,64,640->640), flash_attn (B8 N4096 h10), cross_ip_attn (B8 N4096 h10), layernorm, the 4-link GEMM chain of a level-2 transformer block."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import diffsensei_b200 as ds
from diffsensei_b200.weights import pack_conv3x3

ops = ds.ops
dev = torch.device("cuda:0")
bf = torch.bfloat16
r = lambda *s: torch.randn(*s, device=dev).to(bf)
x = r(8, 128, 128, 320)
ga, be, st = torch.ones(320, device=dev), torch.zeros(320, device=dev), torch.empty(ops.groupnorm_scratch_floats(8, 320), device=dev)
a1, w1, b1 = r(8192, 1280), r(10240, 1280) * 0.03, torch.zeros(10240, device=dev)
a2, w2, b2, res2 = r(32768, 640), r(640, 640) * 0.04, torch.zeros(640, device=dev), r(32768, 640)
a3, w3, b3, res3 = r(8192, 1280), r(1280, 1280) * 0.03, torch.zeros(1280, device=dev), r(8192, 1280)
xc, wc, bc = r(8, 64, 64, 640), pack_conv3x3(torch.randn(640, 640, 3, 3, device=dev) * 0.013), torch.zeros(640, device=dev)
temb = torch.randn(8, 640, device=dev)
qkv = r(8, 4096, 1920)
q, kvt, kvi = r(8, 4096, 640), r(8, 77, 1280), r(8, 80, 1280)
bbox = torch.tensor([[[0.0] * 4] * 4] * 4 + [[[.05, .10, .50, .95], [.50, .15, .95, .90], [0.0] * 4, [0.0] * 4]] * 4, device=dev)
ln_g, ln_b = torch.ones(640, device=dev), torch.zeros(640, device=dev)


# the 4-link chain of a level-2 BasicTransformerBlock (attn2.to_out -> ff.net.0 -> ff.net.2 -> next to_qkv), LayerNorms folded
hch = r(8192, 1280)
w_ff2, b_ff2 = r(1280, 5120) * 0.014, torch.zeros(1280, device=dev)
w_qkv, b_qkv = r(3840, 1280) * 0.03, torch.zeros(3840, device=dev)
cs1, csq = torch.randn(10240, device=dev), torch.randn(3840, device=dev)
stc = [torch.zeros(2 * 8192, dtype=torch.float64, device=dev) for _ in range(3)]


def chain():
    return ops.gemm_chain([
        ((a3, w3, b3), dict(residual=hch, out=hch, row_stats_out=stc[0], row_stats_zeroed=True)),
        ((None, w1, b1), dict(epilogue=ops.EPI_GEGLU, ln_stats=stc[0], ln_colsum=cs1, zero_rows=stc[2])),
        ((None, w_ff2, b_ff2), dict(residual=hch, out=hch, row_stats_out=stc[1], row_stats_zeroed=True)),
        ((None, w_qkv, b_qkv), dict(ln_stats=stc[1], ln_colsum=csq, zero_rows=stc[0]))])


cst = ops.channel_stats(x)
cst_conv = torch.zeros(8, 640, 2, dtype=torch.float64, device=dev)


def run():
    ops.groupnorm_silu(x, ga, be, 32, 1e-5, True, stats=st)           # chan_stats_kernel + gn_apply2_kernel
    ops.groupnorm_apply(x, cst, ga, be, 32, 1e-5, True)               # the in-step form: apply only
    ops.gemm(a1, w1, b1, epilogue=ops.EPI_GEGLU)
    ops.gemm(a2, w2, b2, residual=res2)
    ops.gemm(a3, w3, b3, residual=res3)
    ops.conv3x3(xc, wc, bc, rowbias=temb, residual=xc)
    ops.conv3x3(xc, wc, bc, rowbias=temb, residual=xc, chan_stats=cst_conv)   # with the statistics epilogue
    ops.attention_self(qkv, 10)
    ops.attention_cross_ip(q, kvt, kvi, bbox, 10, 1.0, 0.6, 16, 16)
    ops.layernorm(a2, ln_g, ln_b)
    stc[1].zero_()
    chain()


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
