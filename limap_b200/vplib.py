"""limap.vplib operator surface: VPResult, JLinkage detector and get_vp_detector over the CUDA J-Linkage.

Mirrors src/limap/vplib/{vpbase.h:18-47, bindings.cc:20-37, JLinkage/bindings.cc:14-45,
JLinkage/JLinkage.py, base_vp_detector.py, register_vp_detector.py}. detect_vp_all_images sends every image
to the GPU in one batched call instead of the reference's joblib fan-out (base_vp_detector.py:46-78).
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import Context, check, lib, ptr


class VPResult:
    """vplib/vpbase.h:18-47"""

    def __init__(self, labels=None, vps=None):
        if isinstance(labels, dict):
            labels, vps = labels["labels"], labels["vps"]
        elif isinstance(labels, VPResult):
            labels, vps = labels.labels, labels.vps
        self.labels = [int(x) for x in (labels if labels is not None else [])]
        self.vps = [np.asarray(v, dtype=np.float64) for v in (vps if vps is not None else [])]

    def as_dict(self):
        return {"labels": list(self.labels), "vps": [v.copy() for v in self.vps]}

    def count_lines(self):
        return len(self.labels)

    def count_vps(self):
        return len(self.vps)

    def GetVPLabel(self, line_id):
        return self.labels[line_id]

    def GetVPbyCluster(self, vp_id):
        return self.vps[vp_id]

    def HasVP(self, line_id):
        return self.labels[line_id] >= 0

    def GetVP(self, line_id):
        if not self.HasVP(line_id):
            raise RuntimeError("THROW_CHECK_EQ(HasVP(line_id), true)")
        return self.vps[self.labels[line_id]]


VP_DEFAULTS = dict(min_length=40.0, inlier_threshold=1.0, min_num_supports=5, th_perp_supports=3.0)


def _segs_of(lines):
    arr = getattr(lines, "array", None)
    if arr is not None and len(arr) == len(lines):
        return np.asarray(arr, np.float64)
    return (np.array([[l.start[0], l.start[1], l.end[0], l.end[1]] for l in lines], dtype=np.float64)
            if len(lines) else np.zeros((0, 4)))


class JLinkageDetector:
    """_vplib.JLinkage (vplib/JLinkage/JLinkage.h:24-43). `seed` replaces the library's unseeded RNG."""

    def __init__(self, cfg=None, device=0, seed=0, n_models=5000):
        self.config_ = dict(VP_DEFAULTS)
        self.config_.update({k: v for k, v in (cfg or {}).items() if k in VP_DEFAULTS})
        # JLinkage shadows BaseVPDetector::config_, so count_valid_supports_2d always sees the default 3.0
        # (SURVEY.md §8 a18); mirror that unless the caller overrides it explicitly on this object
        self.th_perp_supports_effective = VP_DEFAULTS["th_perp_supports"]
        self.seed, self.n_models = int(seed), int(n_models)
        self._ctx = Context(device)

    def as_dict(self):
        return dict(self.config_)

    def stats(self):
        st = _cabi.VPStats()
        check(lib().lm_vp_get_stats(self._ctx.handle, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _cabi.VPStats._fields_}

    def detect_batch(self, segs_list, image_index=None):
        """image_index[i]: position of image i in the full image list of the scene (seeds its hypotheses); a rank
        detecting a share of the images passes it to get the labels of the single-call run."""
        idx = None if image_index is None else np.ascontiguousarray(image_index, np.int64)
        off = np.zeros(len(segs_list) + 1, np.int64)
        for i, s in enumerate(segs_list):
            off[i + 1] = off[i] + len(s)
        segs = np.ascontiguousarray(np.concatenate(segs_list, 0) if len(segs_list) else np.zeros((0, 4)), np.float64)
        c = self.config_
        cfg = _cabi.VPConfig(c["min_length"], c["inlier_threshold"], self.th_perp_supports_effective,
                             int(c["min_num_supports"]), self.n_models, self.seed)
        labels = np.full(int(off[-1]), -1, np.int32)
        vp_off = np.zeros(len(segs_list) + 1, np.int64)
        cap = 64 * max(len(segs_list), 1)
        vps = np.zeros((cap, 3))
        n = check(lib().lm_vp_detect_indexed(self._ctx.handle, len(segs_list), ptr(off), ptr(segs), C.byref(cfg), ptr(idx),
                                             ptr(labels), ptr(vp_off), ptr(vps), cap))
        if n > cap:
            vps = np.zeros((n, 3))
            check(lib().lm_vp_detect_indexed(self._ctx.handle, len(segs_list), ptr(off), ptr(segs), C.byref(cfg), ptr(idx),
                                             ptr(labels), ptr(vp_off), ptr(vps), n))
        return [VPResult(labels[off[i]:off[i + 1]], vps[vp_off[i]:vp_off[i + 1]]) for i in range(len(segs_list))]

    def ComputeVPLabels(self, lines):
        return self.AssociateVPs(lines).labels

    def AssociateVPs(self, lines):
        return self.detect_batch([_segs_of(lines)])[0]

    def AssociateVPsParallel(self, all_lines):
        keys = list(all_lines.keys())
        res = self.detect_batch([_segs_of(all_lines[k]) for k in keys])
        return dict(zip(keys, res))


class BaseVPDetectorOptions:
    def __init__(self, n_jobs=1):
        self.n_jobs = n_jobs

    def _replace(self, **kw):
        return BaseVPDetectorOptions(**{"n_jobs": self.n_jobs, **kw})


class JLinkage:
    """vplib/JLinkage/JLinkage.py: Python-level detector object returned by get_vp_detector."""

    def __init__(self, cfg_jlinkage, options=None, seed=0):
        self.n_jobs = getattr(options, "n_jobs", 1)
        self.detector = JLinkageDetector(cfg_jlinkage, seed=seed)

    def get_module_name(self):
        return "JLinkage"

    def detect_vp(self, lines, camview=None):
        return self.detector.AssociateVPs(lines)

    def detect_vp_all_images(self, all_lines, camviews=None):
        return self.detector.AssociateVPsParallel(all_lines)


def get_vp_detector(cfg_vp_detector, n_jobs=1):  # register_vp_detector.py:4-24
    method = cfg_vp_detector["method"]
    if method == "jlinkage":
        return JLinkage(cfg_vp_detector, BaseVPDetectorOptions(n_jobs))
    raise NotImplementedError(f"VP detector '{method}' is outside the hot path")
