"""Isolated timing of the BasicTransformerBlock linear chains: separate launches vs ops.gemm_chain.

    python tools/chain_bench.py [C M]          (default 1280 8192 and 640 32768: the two SDXL transformer levels at cfg2)

Each variant is captured into a CUDA graph that walks `SETS` different weight sets (208 MB at C=1280: more than L2, as in
the real step, where every linear's weights come from HBM) and replayed; times are CUDA-event means per block.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from diffsensei_b200 import ops  # noqa: E402

bf16, DEV, SETS = torch.bfloat16, "cuda", 4


def make(C, M, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    r = lambda *s, k=1.0: (torch.randn(*s, device=DEV, generator=g) * k)
    a, h = r(M, C).to(bf16), r(M, C).to(bf16)
    st = [torch.zeros(2 * M, dtype=torch.float64, device=DEV) for _ in range(3)]
    qkv = torch.empty(M, 3 * C, dtype=bf16, device=DEV)
    q2 = torch.empty(M, C, dtype=bf16, device=DEV)
    W = dict(wo=r(C, C, k=C ** -0.5).to(bf16), bo=r(C), w1=r(8 * C, C, k=C ** -0.5).to(bf16), b1=r(8 * C),
             cs1=r(8 * C), w2=r(C, 4 * C, k=(4 * C) ** -0.5).to(bf16), b2=r(C), wq=r(3 * C, C, k=C ** -0.5).to(bf16),
             bq=r(3 * C), csq=r(3 * C), wq2=r(C, C, k=C ** -0.5).to(bf16), bq2=r(C), csq2=r(C))
    four = [((a, W["wo"], W["bo"]), dict(residual=h, out=h, row_stats_out=st[0], row_stats_zeroed=True)),
            ((None, W["w1"], W["b1"]), dict(epilogue=ops.EPI_GEGLU, ln_stats=st[0], ln_colsum=W["cs1"], zero_rows=st[2])),
            ((None, W["w2"], W["b2"]), dict(residual=h, out=h, row_stats_out=st[1], row_stats_zeroed=True)),
            ((None, W["wq"], W["bq"]), dict(ln_stats=st[1], ln_colsum=W["csq"], zero_rows=st[0], out=qkv))]
    two = [((a, W["wo"], W["bo"]), dict(residual=h, out=h, row_stats_out=st[0], row_stats_zeroed=True)),
           ((None, W["wq2"], W["bq2"]), dict(ln_stats=st[0], ln_colsum=W["csq2"], zero_rows=st[2], out=q2))]
    return four, two, [a, h, torch.empty(M, 4 * C, dtype=bf16, device=DEV), h]


def separate(links, chain_geometry=False):
    prev = None
    for args, kw in links:
        args = (prev,) + args[1:] if args[0] is None else args
        prev = ops.gemm_chain([(args, kw)], min_links=1)[0] if chain_geometry else ops.gemm(*args, **kw)


def capture(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def timed_all(graphs: dict, reps=10, rounds=7) -> dict:
    """Interleaved rounds (power-capped boxes drift by several % within a second): median of per-round means."""
    out = {k: [] for k in graphs}
    for _ in range(rounds):
        for k, g in graphs.items():
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            out[k].append(e0.elapsed_time(e1) / reps * 1e3)
    return {k: sorted(v)[len(v) // 2] for k, v in out.items()}


def main():
    shapes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(1280, 8192), (640, 32768)]
    ops.gemm_chain_prepare()
    for C, M in shapes:
        sets = [make(C, M, s) for s in range(SETS)]
        for name, idx in (("4-link attn2.to_out>ff.net.0>ff.net.2>to_qkv", 0), ("2-link attn1.to_out>attn2.to_q", 1)):
            graphs = {
                "separate (default tiles)": capture(lambda: [separate(s[idx]) for s in sets]),
                "separate (<256,2> tiles, no mixed tail)": capture(lambda: [separate(s[idx], True) for s in sets]),
                "chain": capture(lambda: [ops.gemm_chain(s[idx]) for s in sets])}
            if idx == 0:
                for j, nm in enumerate(("to_out", "ff.net.0", "ff.net.2", "to_qkv")):
                    def one(j=j):
                        for s in sets:
                            args, kw = s[0][j]
                            ops.gemm(*((s[2][j],) + args[1:]), **kw)
                    graphs[f"  alone: {nm}"] = capture(one)
            res = {k: v / SETS for k, v in timed_all(graphs).items()}
            print(f"C={C} M={M}  {name}")
            for k, v in res.items():
                print(f"    {k:<44s} {v:9.1f} us")


if __name__ == "__main__":
    main()
