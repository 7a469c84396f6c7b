"""Remerge pair test + support flags at scale on the GPU, with the CPU oracle on a sample."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from limap_b200.config import LINKER3D_DEFAULTS, make_linker
from limap_b200.engine import MergeEngine
from limap_b200.synth import make_track_lines, make_tracks

LK = dict(score_th=0.5, th_angle=5.0, th_overlap=0.001, th_smartoverlap=0.1, th_smartangle=1.0, th_perp=1.0, th_innerseg=1.0)
eng = MergeEngine()
for T in (20000, 100000, 300000):
    L = make_track_lines(T, dup_frac=0.3, seed=1, extent=60.0)
    act = np.ones(T, np.uint8)
    for it in range(3):
        t0 = time.perf_counter()
        lab, ng, ne = eng.remerge_labels(L, act, make_linker(LINKER3D_DEFAULTS, LK))
        dt = time.perf_counter() - t0
    st = eng.stats()
    print(f"T={T}: groups {ng} edges {ne} gated {st['n_pairs_gated']} kernel {st['last_remerge_kernel_ms']:.3f} ms "
          f"device {st['last_remerge_ms']:.3f} ms wall {dt*1e3:.2f} ms pairs/s {T*(T-1)/2/(st['last_remerge_kernel_ms']*1e-3):.3e}")
if "--cpu" in sys.argv:
    from oracle import oracle as orc
    T = 20000
    L = make_track_lines(T, dup_frac=0.3, seed=1, extent=60.0)
    t0 = time.perf_counter()
    lab, ng, ne = orc.remerge_labels(L, np.ones(T, np.uint8), LK, threads=orc.usable_cpus())
    dt = time.perf_counter() - t0
    print(f"cpu T={T}: groups {ng} edges {ne} {dt:.2f} s pairs/s {T*(T-1)/2/dt:.3e}")
ts = make_tracks(T=100000, S=30, seed=2)
views, first = np.unique(ts.img_ids, return_index=True)
remap = np.zeros(int(views.max()) + 1, np.int32); remap[views] = np.arange(len(views))
for it in range(3):
    t0 = time.perf_counter()
    f = eng.support_flags(None, ts.kvec[first], ts.qvec[first], ts.tvec[first], ts.sup_off, remap[ts.img_ids], ts.segs, ts.line_init)
    dt = time.perf_counter() - t0
st = eng.stats()
print(f"flags: {len(f)} supports kernel {st['last_flags_kernel_ms']:.3f} ms device {st['last_flags_ms']:.3f} ms wall {dt*1e3:.2f} ms; bits {[(int((f & b).astype(bool).sum())) for b in (1,2,4)]}")
