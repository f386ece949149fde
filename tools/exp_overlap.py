#!/usr/bin/env python
"""EXPERIMENT: the day in C chunks of whole buckets on two streams — the solver of chunk c+1 beside the splines and candles of
chunk c (the solver is issue-bound, the splines latency/memory-bound).  Prints the one-stream time and the two-stream time."""
import sys, os, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iv_interpolation_b200 import surface as SF, tapes

dev = torch.device("cuda:0")
n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 7200
spec = SF.ChainSpec(n_t, 30, 2000, 512, 64, bucket_ticks=60 * tapes.US, min_count=10)
chain = tapes.make_quote_chain(0, spec, dev)
P = spec.n_exp * spec.n_strike


def chunked(C):
    ct = n_t // C
    assert ct % 60 == 0
    out = []
    for c in range(C):
        a, b = c * ct, (c + 1) * ct
        sp = dataclasses.replace(spec, n_t=ct)
        q = SF.QuoteChain(sp, chain.ts[a:b], chain.price[a * P:b * P], chain.S[a * P:b * P], chain.K[a * P:b * P], chain.T[a * P:b * P],
                          chain.r[a * P:b * P], chain.is_call[a * P:b * P], chain.grid_k, chain.grid_m)
        out.append((SF.SurfacePipeline(sp, dev), q))
    return out


def timed(fn, reps=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


whole = SF.SurfacePipeline(spec, dev)
t_whole = timed(lambda: whole.run(chain))
print(f"n_t {n_t}: one launch sequence for the whole day: {t_whole:.3f} ms")
del whole
for C in (2, 4, 6):
    if (n_t // C) % 60:
        continue
    parts = chunked(C)
    t_seq = timed(lambda: [p.run(q) for p, q in parts])
    s_iv, s_rest = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)

    def overlapped():
        cur = torch.cuda.current_stream()
        s_iv.wait_stream(cur); s_rest.wait_stream(cur)
        for p, q in parts:
            with torch.cuda.stream(s_iv):
                p.run(q, part="iv")
                ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(s_rest):
                s_rest.wait_event(ev)
                p.run(q, part="rest")
        cur.wait_stream(s_iv); cur.wait_stream(s_rest)

    t_ov = timed(overlapped)
    print(f"  {C} chunks: one stream {t_seq:.3f} ms, two streams {t_ov:.3f} ms  ({t_ov / t_whole:.3f} of the whole-day sequence)")
    del parts
    torch.cuda.empty_cache()
