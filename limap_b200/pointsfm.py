"""limap.pointsfm as far as the triangulation path needs it: the sparse-point model that yields the
visual neighbours and the robust 3D ranges a runner passes to line_triangulation. Running COLMAP itself
(`run_colmap_sfm*`, reading COLMAP folders) stays out of scope; the model is filled with addImage / addPoint exactly
like `pointsfm/colmap_reader.py` / `bundler_reader.py` fill the reference's `_pointsfm.SfmModel`.

  SfmModel.{addImage, addPoint, GetMaxIoUImages, GetMaxDiceCoeffImages, GetMaxOverlapImages, ComputeRanges,
            ComputeNumPoints, GetImageNames}      pointsfm/sfm_model.{h,cc}
  compute_neighbors, compute_metainfos, filter_by_cam_id                                 pointsfm/functions.py:6-55
The ranking runs on the GPU (lm_sfm_rank_neighbors: one triangulation angle per (point, image pair), CUB sorts and a
run-length encoding for the per-pair percentile and counts)."""
import ctypes as C

import numpy as np

from ._cabi import Context, check, lib, ptr


class SfmImage:
    """colmap::mvs::Image as CreateSfmImage builds it (sfm_model.cc:13-23): name, size, K, R, T (row-major)."""

    def __init__(self, filename, width, height, K, R, T):
        self.filename, self.width, self.height = filename, int(width), int(height)
        # colmap::mvs::Image keeps K, R, T as float (CreateSfmImage converts the doubles it is given)
        self.K = np.asarray(K, np.float64).reshape(3, 3).astype(np.float32)
        self.R = np.asarray(R, np.float64).reshape(3, 3).astype(np.float32)
        self.T = np.asarray(T, np.float64).reshape(3).astype(np.float32)

    def centre(self):
        """Projection centre as colmap::mvs::Model::ComputeTriangulationAngles uses it: C = -R^T T evaluated in float
        (ComputeProjectionCenter), widened to double."""
        R, T = self.R, self.T
        c = [-((R[0, i] * T[0] + R[1, i] * T[1]) + R[2, i] * T[2]) for i in range(3)]
        return np.asarray(c, np.float32).astype(np.float64)


def CreateSfmImage(filename, width, height, K, R, T):
    return SfmImage(filename, width, height, K, R, T)


class SfmModel:
    _MODES = {"iou": 0, "dice": 1, "overlap": 2}

    def __init__(self, device=0):
        self.images, self.reg_image_ids = [], []
        self._xyz, self._tracks = [], []
        self._device = device
        self._ctx = None

    def addImage(self, image, img_id=-1):  # sfm_model.cc:48-57
        self.images.append(image)
        if img_id == -1:
            if self.reg_image_ids and self.reg_image_ids[-1] != len(self.reg_image_ids) - 1:
                raise RuntimeError("THROW_CHECK_EQ(reg_image_ids.back(), reg_image_ids.size() - 1)")
            self.reg_image_ids.append(len(self.reg_image_ids))
        else:
            self.reg_image_ids.append(int(img_id))

    def addPoint(self, x, y, z, image_ids):  # sfm_model.cc:25-33 (image_ids are image INDICES, as in the readers)
        # colmap::mvs::Model::Point keeps float coordinates: the triangulation angles and ranges see the rounded values
        self._xyz.append((float(np.float32(x)), float(np.float32(y)), float(np.float32(z))))
        self._tracks.append(np.asarray(image_ids, np.int32).reshape(-1))

    def GetImageNames(self):
        return [im.filename for im in self.images]

    def ComputeNumPoints(self):
        n = np.zeros(len(self.images), np.int64)
        for t in self._tracks:
            np.add.at(n, t, 1)
        return n.tolist()

    def _arrays(self):
        xyz = np.ascontiguousarray(np.asarray(self._xyz, np.float64).reshape(-1, 3))
        off = np.zeros(len(self._tracks) + 1, np.int64)
        if self._tracks:
            off[1:] = np.cumsum([len(t) for t in self._tracks])
        img = np.ascontiguousarray(np.concatenate(self._tracks) if self._tracks else np.zeros(0, np.int32), np.int32)
        return xyz, off, img

    def _rank(self, num_images, min_triangulation_angle, mode):
        if self._ctx is None:
            self._ctx = Context(self._device)
        n = len(self.images)
        centres = np.ascontiguousarray(np.stack([im.centre() for im in self.images]) if n else np.zeros((0, 3)))
        xyz, off, img = self._arrays()
        out = np.full((n, int(num_images)), -1, np.int32)
        cnt = np.zeros(n, np.int32)
        check(lib().lm_sfm_rank_neighbors(self._ctx.handle, n, ptr(centres), len(xyz), ptr(xyz), ptr(off), ptr(img),
                                          int(num_images), float(min_triangulation_angle), int(mode), ptr(out), ptr(cnt)))
        # neighbors_vec_to_map (sfm_model.cc:75-86)
        return {self.reg_image_ids[i]: [self.reg_image_ids[j] for j in out[i, :cnt[i]]] for i in range(n)}

    def GetMaxIoUImages(self, num_images, min_triangulation_angle):
        return self._rank(num_images, min_triangulation_angle, 0)

    def GetMaxDiceCoeffImages(self, num_images, min_triangulation_angle):
        return self._rank(num_images, min_triangulation_angle, 1)

    def GetMaxOverlapImages(self, num_images, min_triangulation_angle):
        return self._rank(num_images, min_triangulation_angle, 2)

    def ComputeRanges(self, range_robust, kstretch):  # sfm_model.cc:245-261
        if self._ctx is None:
            self._ctx = Context(self._device)
        xyz, _, _ = self._arrays()
        out = np.zeros(6)
        check(lib().lm_sfm_robust_ranges(self._ctx.handle, len(xyz), ptr(xyz), float(range_robust[0]), float(range_robust[1]),
                                         float(kstretch), ptr(out)))
        return out[:3].copy(), out[3:].copy()


def filter_by_cam_id(cam_id, prev_imagecols, prev_neighbors):  # pointsfm/functions.py:6-17
    assert prev_imagecols.NumImages() == len(prev_neighbors)
    keep = [i for i in prev_imagecols.get_img_ids() if prev_imagecols.camimage(i).cam_id == cam_id]
    from . import base
    imagecols = base.ImageCollection({c: prev_imagecols.cam(c) for c in prev_imagecols.get_cam_ids()},
                                     {i: prev_imagecols.camimage(i) for i in keep})
    return imagecols, imagecols.update_neighbors(prev_neighbors)


def compute_neighbors(model, n_neighbors, min_triangulation_angle=1.0, neighbor_type="iou"):  # functions.py:20-39
    if neighbor_type == "iou":
        return model.GetMaxIoUImages(n_neighbors, min_triangulation_angle)
    if neighbor_type == "overlap":
        return model.GetMaxOverlapImages(n_neighbors, min_triangulation_angle)
    if neighbor_type == "dice":
        return model.GetMaxDiceCoeffImages(n_neighbors, min_triangulation_angle)
    raise NotImplementedError


def compute_metainfos(cfg, model, n_neighbors=20):  # functions.py:42-55
    neighbors = compute_neighbors(model, n_neighbors, min_triangulation_angle=cfg["min_triangulation_angle"],
                                  neighbor_type=cfg["neighbor_type"])
    ranges = model.ComputeRanges(cfg["ranges"]["range_robust"], cfg["ranges"]["k_stretch"])
    return neighbors, ranges


def _needs_colmap(*a, **k):
    raise NotImplementedError("running / reading COLMAP is outside the hot path; fill an SfmModel with addImage / addPoint")


check_exists_colmap_model = run_colmap_sfm_with_known_poses = run_colmap_sfm = read_infos_colmap = _needs_colmap
