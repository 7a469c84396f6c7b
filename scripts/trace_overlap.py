"""With an engine built with -DLM_TRACE: where the 3 ms between 'device run' and 'add+run' go (groups = 5)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limap_b200.config import DEFAULT_YAML_TRIANGULATION
from limap_b200.engine import TriEngine
from limap_b200.synth import CONFIGS, make_scene
sc = make_scene(**CONFIGS["hypersim100"])
bsrc, bng, boff, bpairs = sc.bulk_matches()
tp = torch.empty(bpairs.shape, dtype=torch.int32, pin_memory=True); tp.numpy()[...] = bpairs; pp = tp.numpy()
eng = TriEngine(dict(DEFAULT_YAML_TRIANGULATION))
eng.set_pipeline_groups(5)
for it in range(4):
    eng.upload(sc); eng.set_ranges(*sc.ranges); torch.cuda.synchronize()
    t = time.perf_counter()
    eng.add_matches_bulk(bsrc, bng, boff, pp)
    t1 = time.perf_counter()
    st = eng.run()
    torch.cuda.synchronize()
    print(f"iter {it}: add returned {1e3*(t1-t):.3f} ms, add+run {1e3*(time.perf_counter()-t):.3f} ms, device run {st['last_run_ms']:.3f}", file=sys.stderr)
