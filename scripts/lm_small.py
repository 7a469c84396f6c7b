"""A small line-BA solve (sanitizer target): 64 tracks x 12 supports, 30 iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limap_b200.engine import BAEngine
from limap_b200.synth import make_tracks

o = BAEngine().solve_trackset(make_tracks(T=64, S=12, V=30, seed=9), max_num_iterations=30)
print("lm", o["stats"]["total_iterations"])
