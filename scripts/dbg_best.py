import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from limap_b200.config import DEFAULT_YAML_TRIANGULATION
from limap_b200.synth import make_scene
from parity_utils import run_both
sc = make_scene(V=12, L=300, N=8, K=10, seed=12)
cfg = dict(DEFAULT_YAML_TRIANGULATION); cfg["debug_mode"] = True
eng, orc = run_both(sc, cfg)
worst = 0
for i in sc.img_ids:
    i = int(i)
    gl, gng, gnc = eng.get_best(i); ol, ong, onc = orc.get_best(i)
    bad = np.where((gng != ong).any(1) & (onc > 0))[0]
    for l in bad[:3]:
        cl, cng = eng.get_cands_node(i, l); rl, rng_ = orc.get_cands_node(i, l)
        d = np.abs(cl[:, 9] - rl[:, 9])
        print("img", i, "line", l, "ncand", onc[l], "best gpu", gng[l], gl[l, 9], "orc", ong[l], ol[l, 9], "max score diff", d.max(), "at", d.argmax(), cl[d.argmax(), 9], rl[d.argmax(), 9])
        top = np.argsort(-rl[:, 9])[:3]
        print("   top oracle scores", rl[top, 9], "gpu same idx", cl[top, 9])
    for l in range(len(onc)):
        if onc[l] == 0: continue
        cl, cng = eng.get_cands_node(i, l); rl, rng_ = orc.get_cands_node(i, l)
        worst = max(worst, np.abs(cl[:, 9] - rl[:, 9]).max())
print("worst score diff over all candidates", worst)
