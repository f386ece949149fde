"""Where does the strike spline's time go on a tape with missing quotes?  general kernel / ladder-verify kernel / banded + list."""
import torch
from iv_interpolation_b200 import device as D, surface as SF, tapes
dev = torch.device("cuda:0")
spec = SF.ChainSpec(1200, 30, 2000, 512, 64, bucket_ticks=60 * tapes.US, min_count=1)
for drop in (0.0, 0.02):
    chain = tapes.make_quote_chain(5, spec, dev, drop_frac=drop)
    pipe = SF.SurfacePipeline(spec, dev, max_buckets=40)
    pipe.run(chain)
    iv = pipe.iv
    def timed(fn, n=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    out = torch.empty_like(pipe.smile)
    t_gen = timed(lambda: D.spline_fit_eval(pipe.knot_off, chain.K, iv, chain.grid_k, max_knots=2000, out=out))
    t_ver = timed(lambda: D.spline_fit_eval(pipe.knot_off, chain.K, iv, chain.grid_k, max_knots=2000, out=out, x_off=pipe._q_x_off, lu=pipe._q_lu,
                                            slice_ladder=pipe._q_slice_ladder, x_own=chain.K, ladder_off=pipe._q_ladder_off))
    t_band = timed(lambda: D.spline_fit_eval_banded(pipe.knot_off, pipe._q_x_off, pipe._q_slice_ladder, pipe._order, pipe._lad_slice_off, spec.n_t,
                                                    pipe._q_ladder_off, chain.K, pipe._q_lu, iv, chain.grid_k, pipe._q_band, max_knots=2000,
                                                    x_own=chain.K, out=out, fb=pipe._fb))
    print(f"drop {drop}: general {t_gen:.3f} ms, ladder-verify {t_ver:.3f} ms, banded+list {t_band:.3f} ms, listed {int(pipe._fb[1].item())} of {spec.n_t * 30}")
    del chain, pipe
