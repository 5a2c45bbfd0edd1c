"""Timeline of one time step of the persistent sweep kernel from the global-timer stamps of
Engine.sweep_trace (gpu_perf_probe.py with MMMFE_TRACE_ALL=<file.npz>): per tree level and direction, when the
first / median / last CTA starts and ends, run and dependency-wait time per CTA.  Usage: python gpu_trace_report.py f.npz"""
import sys

import numpy as np


def main(path):
    d = np.load(path)
    cb, kh, st = d["cb"], d["kh"], d["st"]
    ok = st[:, 0] > 0
    nc = len(cb) - 1
    t0 = st[ok, 0].min()
    print(f"step span {(st[ok, 2].max() - t0) / 1e3:.1f} us, {nc} CTAs, {int(ok.sum())} entries")
    cta_of = np.zeros(len(st), dtype=int)
    for c in range(nc):
        cta_of[cb[c]:cb[c + 1]] = c
    tot_run = np.zeros(nc)
    tot_wait = np.zeros(nc)

    def level(kind, h):
        m = ok & (kh[:, 0] == kind) & (kh[:, 1] == h)
        if not m.any():
            return
        s, c = st[m], cta_of[m]
        run = np.bincount(c, weights=s[:, 2] - s[:, 1], minlength=nc) / 1e3
        wait = np.bincount(c, weights=s[:, 1] - s[:, 0], minlength=nc) / 1e3
        cnt = np.bincount(c, minlength=nc)
        fs = np.full(nc, np.inf)
        le = np.zeros(nc)
        np.minimum.at(fs, c, (s[:, 0] - t0) / 1e3)
        np.maximum.at(le, c, (s[:, 2] - t0) / 1e3)
        a = cnt > 0
        tot_run[:] += run
        tot_wait[:] += wait
        name = ["sub fwd", "top fwd", "top bwd", "sub bwd"][kind]
        print(f"{name} h{h:2d} n {int(m.sum()):5d} | start first {fs[a].min():7.1f} med {np.median(fs[a]):7.1f} | end med "
              f"{np.median(le[a]):7.1f} last {le[a].max():7.1f} | run/cta mean {run[a].mean():6.1f} max {run[a].max():6.1f} "
              f"| wait mean {wait[a].mean():6.1f} max {wait[a].max():6.1f} | entries/cta {cnt[a].mean():4.1f} max {cnt.max()}")

    hs = sorted(set(kh[ok & (kh[:, 0] == 1), 1]))
    level(0, 0)
    for h in hs:
        level(1, h)
    for h in sorted(set(kh[ok & (kh[:, 0] == 2), 1]), reverse=True):
        level(2, h)
    level(3, 0)
    print(f"per CTA: run mean {tot_run.mean():.1f} min {tot_run.min():.1f} max {tot_run.max():.1f} us | dependency wait mean "
          f"{tot_wait.mean():.1f} min {tot_wait.min():.1f} max {tot_wait.max():.1f} us")
    for kind, name in enumerate(["sub fwd", "top fwd", "top bwd", "sub bwd"]):
        m = ok & (kh[:, 0] == kind)
        r = (st[m, 2] - st[m, 1]) / 1e3
        print(f"  {name}: {int(m.sum() / nc):4d} entries/cta, run per entry p10 {np.percentile(r, 10):.2f} p50 {np.percentile(r, 50):.2f} "
              f"p90 {np.percentile(r, 90):.2f} us, run/cta {r.sum() / nc:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
