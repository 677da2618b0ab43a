#!/usr/bin/env python
"""One-shot tuning sweep of the GroupNorm apply kernel's launch knobs (each configuration in its own process: the
knobs are read once per process).  python tools/gn_sweep.py  ->  table of us / GB/s per (threads, unroll, occ, hint)."""
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import diffsensei_b200 as ds
ops = ds.ops
dev = torch.device("cuda:0")
bf = torch.bfloat16
def timed(calls, rounds=6):
    for c in calls: c()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(rounds):
        for c in calls: c()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / (rounds * len(calls))
res = []
for (B, H, W, C) in ((8, 128, 128, 320), (8, 64, 64, 640), (8, 32, 32, 1280), (8, 128, 128, 640)):
    ga, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    n = max(2, int(700e6 // (B * H * W * C * 4)) + 1)
    sets = []
    for i in range(min(n, 12)):
        x = torch.randn(B, H, W, C, device=dev).to(bf)
        sets.append((x, torch.empty_like(x), ops.channel_stats(x)))
    ms = timed([(lambda s=s: ops.groupnorm_apply(s[0], s[2], ga, be, 32, 1e-5, True, out=s[1])) for s in sets])
    gb = 2 * B * H * W * C * 2 / 1e9
    res.append("%%dx%%dx%%dx%%d %%.1fus %%.0fGB/s" %% (B, H, W, C, ms * 1e3, gb / (ms * 1e-3)))
print(" | ".join(res))
''' % ROOT
for threads, unroll, occ, hint in [(256, 8, 8, 1), (256, 4, 8, 1), (512, 8, 8, 1), (512, 4, 8, 1), (128, 8, 16, 1),
                                   (256, 8, 8, 0), (384, 8, 8, 1), (1024, 4, 8, 1), (256, 8, 1, 1)]:
    env = dict(os.environ, DS_GN_THREADS=str(threads), DS_GN_UNROLL=str(unroll), DS_GN_OCC=str(occ),
               DS_GN_L2HINT=str(hint))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"threads={threads:4d} unroll={unroll} occ<={occ:2d} hint={hint}: {r.stdout.strip() or r.stderr[-300:]}", flush=True)
