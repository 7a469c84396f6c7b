import sys
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from limap_b200.synth import make_scene
from limap_b200.config import DEFAULT_YAML_TRIANGULATION
from parity_utils import run_both
sc = make_scene(V=8, L=100, N=5, K=4, seed=14, scale=100.0, id_stride=7, shuffle_rows=True)
cfg = dict(DEFAULT_YAML_TRIANGULATION); cfg['debug_mode']=True
eng, orc = run_both(sc, cfg)
gt, ot = eng.build_tracks(), orc.build_tracks()
np.savez(os.path.join(R,'gpurun_out','dbg_tracks.npz'), **{'g_'+k:v for k,v in gt.items()}, **{'o_'+k:v for k,v in ot.items()})
print("saved")
