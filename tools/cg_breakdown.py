"""Where a CR CG iteration (BASELINE configs[2]) spends its time: Legendre kernels, ring-FFT stages (library event
timers) and the remainder (vector passes, dot products, host synchronisation)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import argparse
    import torch
    import bench
    from commander_b200 import sharp
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)

    class Args:
        pass
    bench.run_cg_metric(Args(), None, dev, 1)          # warm-up: plans, tables
    sharp.set_profiling(True)
    sharp.last_legendre_ms()
    r = bench.run_cg_metric(Args(), None, dev, 1)
    ent = sharp.last_legendre_ms()
    sharp.set_profiling(False)
    n = 11 + 3                                          # matmulA calls of the timed solve + warm-up solve inside run_cg_metric
    leg = sum(ms for s, d, ms in ent if s < 100)
    fft = sum(ms for s, d, ms in ent if 100 <= s < 200)
    calls = len([1 for s, d, ms in ent if s < 100]) / 4.0
    print(f"ms_per_iter {r['ms_per_iter']:.3f}  legendre {leg / calls:.3f}  ringfft {fft / calls:.3f}  "
          f"rest {r['ms_per_iter'] - (leg + fft) / calls:.3f}  (A applications profiled: {calls:.0f})")


if __name__ == "__main__":
    main()
