"""Runner-level drop-in (BASELINE.json configs[0]: the reference plumbing on a 10-image scene; SURVEY.md §8d config 1:
V=10, L=800, N=9, K=10 through limap.runners.line_triangulation with load_det / load_match artefacts).

  * CPU, where /root/reference exists: the REFERENCE'S OWN runner file, src/limap/runners/line_triangulation.py, is
    loaded by path and executed UNMODIFIED against this repository's `limap` package (base, triangulation, merging,
    optimize, vplib, util.io, runners, visualize). No GPU is needed because the three engine classes are replaced by
    oracle-backed stand-ins for the duration of the test. Its result must equal the result of this repository's own
    runner mirror (limap_b200/runners.py) on the same artefacts: same tracks, same refined lines, same files on
    disk, same [Track Report].
  * GPU: the runner mirror with the real CUDA engines against the same mirror on the oracle stand-ins: track
    membership bit-exact, refined endpoints 1e-4, [Track Report] equal.
Together: reference runner == mirror (same surface), mirror on CUDA == mirror on the oracle (same arithmetic)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from limap_b200.config import default_runner_config
from limap_b200.synth import CONFIGS, make_scene

from runner_utils import imagecols_of, install_oracle_backend, summarize, write_artifacts

REF_RUNNER = "/root/reference/src/limap/runners/line_triangulation.py"


def _scene(small):
    if small:
        return make_scene(V=8, L=120, N=5, K=6, seed=51)
    return make_scene(**CONFIGS["hypersim10"])


def _cfg(tmp, sc, **over):
    cfg = default_runner_config(output_dir=str(tmp / "out"), load_dir=str(tmp / "artefacts"), n_neighbors=9, **over)
    write_artifacts(sc, cfg, cfg["load_dir"])
    return cfg


def _run_mirror(cfg, sc):
    import copy
    import limap.runners as runners
    return runners.line_triangulation(copy.deepcopy(cfg), imagecols_of(sc), neighbors=dict(sc.neighbors), ranges=sc.ranges)


def _report(tracks):
    import limap.visualize as vis
    return vis.Open3DTrackVisualizer(tracks).track_report()


@pytest.mark.skipif(not os.path.exists(REF_RUNNER), reason="the reference tree is only present in the authoring container")
def test_reference_runner_file_runs_unmodified_on_this_surface(tmp_path, monkeypatch, capsys):
    import copy
    install_oracle_backend(monkeypatch)
    # the only names the reference runner imports that are not part of the hot path: pycolmap (logging) -- stubbed
    pyc = types.ModuleType("pycolmap")
    pyc.logging = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, error=lambda *a, **k: None)
    monkeypatch.setitem(sys.modules, "pycolmap", pyc)
    monkeypatch.setitem(sys.modules, "pycolmap.logging", pyc.logging)
    import limap  # noqa: F401  (alias package: limap.X -> limap_b200.X)
    spec = importlib.util.spec_from_file_location("reference_line_triangulation", REF_RUNNER)
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    with open(REF_RUNNER) as f:
        assert "def line_triangulation(cfg, imagecols, neighbors=None, ranges=None):" in f.read()

    sc = _scene(small=True)
    cfg = _cfg(tmp_path, sc)
    cfg_ref = copy.deepcopy(cfg)
    cfg_ref["output_dir"] = str(tmp_path / "out_ref")
    ref_tracks = ref_mod.line_triangulation(cfg_ref, imagecols_of(sc), neighbors=dict(sc.neighbors), ranges=sc.ranges)
    out = capsys.readouterr().out
    assert "[Track Report]" in out
    my_tracks = _run_mirror(cfg, sc)
    (m_ref, l_ref), (m_my, l_my) = summarize(ref_tracks), summarize(my_tracks)
    assert len(m_ref) > 20 and m_ref == m_my
    assert np.array_equal(l_ref, l_my)  # same backend, same call sequence: bit-identical
    assert _report(ref_tracks) == _report(my_tracks)
    # the files a user finds afterwards
    for rel in ("image_list.txt", "imagecols.npy", "metainfos.txt", "alltracks.txt", "finaltracks/track_0.txt",
                "triangulated_lines_nv4.obj"):
        a, b = tmp_path / "out_ref" / rel, tmp_path / "out" / rel
        assert a.exists() and b.exists(), rel
        if rel.endswith(".txt") or rel.endswith(".obj"):
            assert a.read_text() == b.read_text(), rel
    import limap.util.io as limapio
    back, cfg_back, ic_back, segs_back = limapio.read_folder_linetracks_with_info(str(tmp_path / "out" / "finaltracks"))
    assert len(back) == len(my_tracks) and ic_back.NumImages() == len(sc.img_ids) and len(segs_back) == len(sc.img_ids)


@pytest.mark.gpu
def test_runner_mirror_cuda_equals_oracle_backend_hypersim10(tmp_path, monkeypatch):
    sc = _scene(small=False)  # V=10, L=800, N=9, K=10: the configs[0] stand-in
    cfg = _cfg(tmp_path, sc)
    gpu_tracks = _run_mirror(cfg, sc)
    rep_gpu = _report(gpu_tracks)
    with monkeypatch.context() as mp:
        install_oracle_backend(mp)
        cfg2 = dict(cfg, output_dir=str(tmp_path / "out_cpu"))
        cpu_tracks = _run_mirror(cfg2, sc)
    (m_g, l_g), (m_c, l_c) = summarize(gpu_tracks), summarize(cpu_tracks)
    assert len(m_c) > 100 and m_g == m_c
    d = np.minimum(np.abs(l_g - l_c).max(1), np.abs(l_g - l_c[:, [3, 4, 5, 0, 1, 2]]).max(1))
    assert d.max() <= 1e-4, d.max()
    assert rep_gpu == _report(cpu_tracks) and rep_gpu[0] == len(m_c)
    assert (tmp_path / "out" / "finaltracks" / "track_0.txt").exists()


REF_TEST_LINEBASE = "/root/reference/tests/base/test_linebase.py"


@pytest.mark.skipif(not os.path.exists(REF_TEST_LINEBASE), reason="the reference tree is only present in the authoring container")
def test_reference_own_linebase_test_passes_on_the_mirror():
    """The one test of the reference's own suite that touches a hot-path type (tests/base/test_linebase.py), loaded by
    path and run unmodified against `import limap` = this repository's alias package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_test_linebase", REF_TEST_LINEBASE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import limap
    assert mod.limap is limap and limap.base.__name__ == "limap_b200.base"
    ran = 0
    for name in dir(mod):
        if name.startswith("test_"):
            getattr(mod, name)()
            ran += 1
    assert ran >= 1
