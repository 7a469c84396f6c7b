"""ctypes wrapper of oracle/_ref/liblimap_ref.so: the REFERENCE'S OWN hot-path sources, compiled unchanged from
/root/reference (oracle/Makefile target `ref`, header shims in oracle/ref_shim/). TEST INFRASTRUCTURE ONLY: used by
tests/ to pin the restatement (oracle/*.h) to the reference's compiled code, function by function and as a whole
pipeline. The .so is built in the authoring container (where /root/reference exists), is git-ignored and travels to the
GPU box with the snapshot; nothing here reads /root/reference at run time."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _orc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liblimap_ref.so")
REFERENCE_SRC = "/root/reference/src"
_P = C.c_void_p
_lib = None


def can_build():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "limap"))


def build(force=False):
    """Compile the reference's sources where they lie (only possible where /root/reference exists)."""
    if not can_build():
        return LIB_PATH if os.path.exists(LIB_PATH) else None
    env = dict(os.environ)
    env.pop("CXX", None)
    cmd = ["make", "-C", _HERE, "-j", str(max(1, _orc.usable_cpus())), "ref"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
    return LIB_PATH


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/liblimap_ref.so is missing (built only where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_tri_create.restype = _P
        L.ref_tri_create.argtypes = [_P]
        L.ref_tri_destroy.argtypes = [_P]
        L.ref_tri_init.argtypes = [_P, C.c_int] + [_P] * 7
        L.ref_tri_set_ranges.argtypes = [_P, _P, _P]
        L.ref_tri_unset_ranges.argtypes = [_P]
        L.ref_tri_triangulate_image.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P]
        L.ref_tri_triangulate_image_exhaustive.argtypes = [_P, C.c_int, C.c_int, _P]
        L.ref_tri_rows_tested.restype = C.c_longlong
        L.ref_tri_rows_tested.argtypes = [_P]
        L.ref_tri_get_best.argtypes = [_P, C.c_int, _P, _P, _P]
        L.ref_tri_get_valid_edges.restype = C.c_longlong
        L.ref_tri_get_valid_edges.argtypes = [_P, C.c_int, _P, _P]
        L.ref_tri_get_tris_node.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, _P]
        L.ref_tri_compute_tracks.argtypes = [_P, _P]
        L.ref_tri_get_tracks.argtypes = [_P] * 7
        L.ref_set_num_threads.argtypes = [C.c_int]
        for name in ("ref_line2d_length", "ref_compute_epipolar_IoU", "ref_score_3d", "ref_score_2d",
                     "ref_line3d_sensitivity", "ref_line3d_uncertainty"):
            getattr(L, name).restype = C.c_double
        L.ref_line2d_length.argtypes = [_P]
        L.ref_line2d_direction.argtypes = [_P, _P]
        L.ref_compute_epipolar_IoU.argtypes = [_P] * 4
        L.ref_triangulate_line.argtypes = [_P, _P, _P, _P, C.c_int, _P]
        L.ref_triangulate_line_with_direction.argtypes = [_P] * 6
        L.ref_project_point.argtypes = [_P, _P, _P]
        L.ref_ray_direction.argtypes = [_P, _P, _P]
        L.ref_line3d_sensitivity.argtypes = [_P, _P]
        L.ref_line3d_uncertainty.argtypes = [_P, _P, C.c_double]
        L.ref_score_3d.argtypes = [_P, _P, _P]
        L.ref_score_2d.argtypes = [_P, _P, _P]
        L.ref_check_connection_3d.argtypes = [_P, _P, _P]
        L.ref_aggregate_lines.argtypes = [C.c_int64, _P, _P, _P, C.c_int32, _P]
        L.ref_minimal_from_line.argtypes = [_P, _P]
        L.ref_infinite_from_minimal.argtypes = [_P, _P, _P]
        L.ref_segment_from_minimal.argtypes = [_P, _P, C.c_int64, C.c_int, _P]
        L.ref_track_support_flags.argtypes = [C.c_int32, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P] + [C.c_double] * 4 + [_P]
        L.ref_geometric_residual.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_double, _P, _P]
        L.ref_vp_residual.argtypes = [C.c_int, _P, _P, _P, _P, _P, _P]
        L.ref_sfm_rank_neighbors.argtypes = [C.c_int, _P, _P, C.c_int64, _P, _P, _P, C.c_int, C.c_double, C.c_int, _P, _P]
        L.ref_sfm_robust_ranges.argtypes = [C.c_int64, _P, C.c_double, C.c_double, C.c_double, _P]
        L.ref_vp_set_context.argtypes = [C.c_uint64, C.c_uint64]
        L.ref_vp_associate.restype = C.c_longlong
        L.ref_vp_associate.argtypes = [C.c_int64, _P, C.c_double, C.c_double, C.c_int, C.c_double, _P, _P, C.c_longlong]
        L.ref_linetrack_write.argtypes = [C.c_char_p, _P, C.c_int64, _P, _P, _P, _P, _P, _P]
        L.ref_linetrack_read.restype = C.c_int64
        L.ref_linetrack_read.argtypes = [C.c_char_p, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P]
        L.ref_line_weights.argtypes = [C.c_int64, _P, _P]
        L.ref_track_filter.restype = C.c_int64
        L.ref_track_filter.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _P, C.c_int32, _P, _P, _P, _P, C.c_int64] + [_P] * 18
        L.ref_remerge_groups.restype = C.c_int64
        L.ref_remerge_groups.argtypes = [C.c_int64, _P, _P, _P, _P]
        _lib = L
    return _lib


class RefTri(_orc.OracleTri):
    """The reference's GlobalLineTriangulator itself (compiled from /root/reference), behind the array-level interface
    of OracleTri / limap_b200.engine.TriEngine."""
    _PREFIX = "ref_"

    @staticmethod
    def _library():
        return lib()

    def __init__(self, cfg=None, threads=None):
        super().__init__(cfg, threads=threads, node_parallel=False)


_p, _f64 = _orc._p, _orc._f64


def linker_cfg(d):
    base = dict(score_th=0.5, th_angle=10.0, th_overlap=0.01, th_smartoverlap=0.1, th_smartangle=1.0, th_perp=0.02,
                th_innerseg=0.02, th_scaleinv=0.01, use_angle=1, use_overlap=1, use_smartangle=1, use_perp=0,
                use_innerseg=1, use_scaleinv=0)
    base.update({k: v for k, v in d.items() if k in base})
    return _orc.OrcLinkerCfg(*[float(base[n]) if t is C.c_double else int(bool(base[n])) for n, t in _orc.OrcLinkerCfg._fields_])


def aggregate_lines(off, lines, scores, num_outliers):
    off = np.ascontiguousarray(off, np.int64)
    lines, scores = _f64(lines), _f64(scores)
    out = np.zeros((len(off) - 1, 7))
    lib().ref_aggregate_lines(len(off) - 1, _p(off), _p(lines), _p(scores), int(num_outliers), _p(out))
    return out


def track_support_flags(model_ids, kvec, qvec, tvec, sup_off, sup_view, segs, track_line, th_angular_2d=8.0,
                        th_perp_2d=5.0, th_sv_angular_3d=75.0, th_overlap=0.5):
    sup_off = np.ascontiguousarray(sup_off, np.int64)
    sup_view = np.ascontiguousarray(sup_view, np.int32)
    model_ids = None if model_ids is None else np.ascontiguousarray(model_ids, np.int32)
    kvec, qvec, tvec, segs, track_line = map(_f64, (kvec, qvec, tvec, segs, track_line))
    flags = np.zeros(int(sup_off[-1]), np.uint8)
    lib().ref_track_support_flags(len(kvec), _p(model_ids), _p(kvec), _p(qvec), _p(tvec), len(sup_off) - 1, _p(sup_off),
                                  _p(sup_view), _p(segs), _p(track_line), th_angular_2d, th_perp_2d, th_sv_angular_3d,
                                  th_overlap, _p(flags))
    return flags


def remerge_groups(track_line, active, linker):
    """Partition of the tracks produced by the reference's RemergeLineTracks: group[t] = smallest index merged with t."""
    track_line = _f64(track_line)
    active = np.ascontiguousarray(active, np.uint8)
    cfg = linker_cfg(linker)
    group = np.arange(len(track_line), dtype=np.int32)
    n = lib().ref_remerge_groups(len(track_line), _p(track_line), _p(active), C.byref(cfg), _p(group))
    return group, int(n)


def colmap_float_centres(R, T):
    """Projection centres as colmap::mvs::Model::ComputeTriangulationAngles derives them from an image built by
    CreateSfmImage: R and T rounded to float, C = -R^T T evaluated in float, widened to double."""
    R = np.asarray(R, np.float64).reshape(-1, 3, 3).astype(np.float32)
    T = np.asarray(T, np.float64).reshape(-1, 3).astype(np.float32)
    c = np.empty((len(R), 3), np.float32)
    for i in range(3):
        c[:, i] = -((R[:, 0, i] * T[:, 0] + R[:, 1, i] * T[:, 1]) + R[:, 2, i] * T[:, 2])
    return c.astype(np.float64)


def sfm_rank_neighbors(R, T, xyz, track_off, track_img, num_images, min_triangulation_angle=1.0, mode=0):
    """limap::pointsfm::SfmModel::{GetMaxIoUImages, GetMaxDiceCoeffImages, GetMaxOverlapImages} of the reference's compiled
    sfm_model.cc (COLMAP's mvs::Model statistics come from oracle/ref_shim)."""
    L = lib()
    R = np.ascontiguousarray(np.asarray(R, np.float64).reshape(-1, 9))
    T = np.ascontiguousarray(np.asarray(T, np.float64).reshape(-1, 3))
    xyz = np.ascontiguousarray(xyz, np.float64)
    track_off = np.ascontiguousarray(track_off, np.int64)
    track_img = np.ascontiguousarray(track_img, np.int32)
    n = len(R)
    out = np.full((n, int(num_images)), -1, np.int32)
    cnt = np.zeros(n, np.int32)
    L.ref_sfm_rank_neighbors(n, _p(R), _p(T), len(xyz), _p(xyz), _p(track_off), _p(track_img), int(num_images),
                             float(min_triangulation_angle), int(mode), _p(out), _p(cnt))
    return out, cnt


def sfm_robust_ranges(xyz, q_lo, q_hi, kstretch):
    L = lib()
    xyz = np.ascontiguousarray(xyz, np.float64)
    out = np.zeros(6)
    L.ref_sfm_robust_ranges(len(xyz), _p(xyz), float(q_lo), float(q_hi), float(kstretch), _p(out))
    return out[:3].copy(), out[3:].copy()


def vp_associate(segs, seed=0, image_index=0, min_length=40.0, inlier_threshold=1.0, min_num_supports=5,
                 th_perp_supports=3.0):
    """vplib::JLinkage::JLinkage::AssociateVPs of the reference's compiled wrapper for ONE image (5000 hypotheses, as the
    wrapper hard-codes); the JLinkage library underneath it is the oracle's restatement, seeded by (seed, image_index)."""
    L = lib()
    segs = np.ascontiguousarray(segs, np.float64).reshape(-1, 4)
    labels = np.full(len(segs), -1, np.int32)
    vps = np.zeros((64, 3))
    L.ref_vp_set_context(int(seed), int(image_index))
    n = L.ref_vp_associate(len(segs), _p(segs), float(min_length), float(inlier_threshold), int(min_num_supports),
                           float(th_perp_supports), _p(labels), _p(vps), 64)
    return labels, vps[:n]
