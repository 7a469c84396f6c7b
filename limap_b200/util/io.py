"""limap.util.io — on-disk formats at the boundary of the hot path.

Same file layouts as the reference so that artefacts written by either side load in the other:
  matches_{id}.npy        pickled dict {ng_img_id: (M, 2) int array}     line2d/base_matcher.py:80-117, util/io.py:39-49
  segments_{id}.txt       "n\\n" + n rows "x1 y1 x2 y2"                   util/io.py:441-474
  metainfos.txt           neighbors + ranges                             util/io.py:87-131
  alltracks.txt           all tracks in one file                         util/io.py:259-293
  finaltracks/track_k.txt LineTrack::Write / Read                        base/linetrack.cc:133-270, util/io.py:296-346
  imagecols.npy, all_2d_segs.npy, config.npy next to the tracks          util/io.py:323-346
Only numpy and the standard library are used (tqdm / pycolmap logging of the reference are dropped).
"""
import os
import shutil

import numpy as np

from .. import base


def _require_parent(fname):
    parent = os.path.dirname(fname)
    if parent and not os.path.isdir(parent):
        raise ValueError(f"Error! Base directory {parent} does not exist!")


def _require(path):
    if not os.path.exists(path):
        raise ValueError(f"Error! File {path} does not exist!")


check_directory, check_path = _require_parent, _require  # names of util/io.py:11-26


def check_makedirs(folder):
    os.makedirs(folder, exist_ok=True)


def delete_folder(folder):
    shutil.rmtree(folder, ignore_errors=True)


def save_npy(fname, nparray):
    """Pickled object array, as util/io.py:39-42 writes it (np.save of np.array(obj, dtype=object))."""
    _require_parent(fname)
    payload = np.array(nparray, dtype=object)
    with open(fname, "wb") as fh:
        np.save(fh, payload)


def read_npy(fname):
    _require(fname)
    with open(fname, "rb") as fh:
        return np.load(fh, allow_pickle=True)


def save_npz(fname, dic):
    _require_parent(fname)
    np.savez(fname, **dic)


def read_npz(fname):
    _require(fname)
    return np.load(fname, allow_pickle=True)


# ---- matches: matches_{id}.npy = pickled {ng_img_id: (M, 2) int array} (line2d/base_matcher.py:80-117) ----------
def get_match_filename(matches_folder, idx):
    return os.path.join(matches_folder, "matches_%s.npy" % idx)


def save_match(matches_folder, idx, matches):
    save_npy(get_match_filename(matches_folder, idx), matches)


def read_match(matches_folder, idx):
    return read_npy(get_match_filename(matches_folder, idx)).item()


# ---- metainfos.txt: neighbors + ranges (util/io.py:87-131) -----------------------------------------------------
def save_txt_metainfos(fname, neighbors, ranges):
    _require_parent(fname)
    lo, hi = ranges
    rows = [f"number of images, {len(neighbors)}"]
    rows += [f"{axis}-range, {lo[k]}, {hi[k]}" for k, axis in enumerate("xyz")]
    rows += [", ".join([f"image {img_id}"] + [str(n) for n in ngs]) for img_id, ngs in neighbors.items()]
    with open(fname, "w") as fh:
        fh.write("\n".join(rows) + "\n")


def read_txt_metainfos(fname):
    _require(fname)
    with open(fname) as fh:
        rows = [r.strip() for r in fh]
    fields = lambda row: [x for x in row.split(",")[1:]]
    n_images = int(fields(rows[0])[0])
    lo, hi = np.zeros(3), np.zeros(3)
    for k in range(3):
        lo[k], hi[k] = (float(x) for x in fields(rows[1 + k])[:2])
    neighbors = {}
    for row in rows[4:4 + n_images]:
        head = row.split(",")[0]
        neighbors[int(head[len("image "):])] = [int(x) for x in fields(row)]
    return neighbors, (lo, hi)


# ---- segments_{id}.txt: count, then "x1 y1 x2 y2" per row (util/io.py:436-474) -----------------------------------
def _segments_file(folder, img_id):
    return os.path.join(folder, "segments_%s.txt" % img_id)


def exists_txt_segments(folder, img_id):
    return os.path.exists(_segments_file(folder, img_id))


def save_txt_segments(folder, img_id, segs):
    segs = np.asarray(segs)
    body = "".join("%s %s %s %s\n" % (s[0], s[1], s[2], s[3]) for s in segs)
    with open(_segments_file(folder, img_id), "w") as fh:
        fh.write("%d\n%s" % (segs.shape[0], body))


def read_txt_segments(folder, img_id):
    _require(folder)
    with open(_segments_file(folder, img_id)) as fh:
        rows = fh.read().splitlines()
    n = int(rows[0])
    if n + 1 != len(rows):
        raise AssertionError("segment count does not match the number of rows")
    return np.array([[float(x) for x in r.split(" ")] for r in rows[1:]])


def read_all_segments_from_folder(folder):
    out = {}
    for name in os.listdir(folder):
        img_id = int(name[len("segments_"):-len(".txt")])
        out[img_id] = read_txt_segments(folder, img_id)
    return out


# ---- line tracks (util/io.py:259-346) ----------------------------------------------------------------------------
def save_txt_linetracks(fname, linetracks, n_visible_views=4):
    """alltracks.txt. A 3D point row is "x <18 blanks>y <18 blanks>z": the reference's f-string is continued over
    source lines, which leaves the source indentation inside the row."""
    parent = os.path.dirname(fname)
    if parent:
        os.makedirs(parent, exist_ok=True)
    kept = [t for t in linetracks if t.count_images() >= n_visible_views]
    gap = " " + " " * 18
    point = lambda p: gap.join("%.10f" % float(v) for v in p)
    ids = lambda seq: "".join("%s " % v for v in seq)
    out = [str(len(kept))]
    for k, t in enumerate(kept):
        out += ["%d %d %d" % (k, t.count_lines(), t.count_images()), point(t.line.start), point(t.line.end),
                ids(t.image_id_list), ids(t.line_id_list)]
    with open(fname, "w") as fh:
        fh.write("\n".join(out) + "\n")


def read_txt_linetracks(fname):
    """Inverse of save_txt_linetracks: [(line (2,3), image_id_list, line_id_list)] (whitespace tolerant)."""
    _require(fname)
    with open(fname) as fh:
        tok = fh.read().split()
    pos = 1
    out = []
    for _ in range(int(tok[0])):
        n_lines = int(tok[pos + 1])
        pos += 3
        line = np.array(tok[pos:pos + 6], dtype=np.float64).reshape(2, 3)
        pos += 6
        img = [int(x) for x in tok[pos:pos + n_lines]]
        lid = [int(x) for x in tok[pos + n_lines:pos + 2 * n_lines]]
        pos += 2 * n_lines
        out.append((line, img, lid))
    return out


def save_folder_linetracks(folder, linetracks):
    delete_folder(folder)
    os.makedirs(folder)
    for k, t in enumerate(linetracks):
        t.Write(os.path.join(folder, "track_%d.txt" % k))


def read_folder_linetracks(folder):
    _require(folder)
    n_tracks = len([f for f in os.listdir(folder) if f.startswith("track") and f.endswith(".txt")])
    out = []
    for k in range(n_tracks):
        t = base.LineTrack()
        t.Read(os.path.join(folder, "track_%d.txt" % k))
        out.append(t)
    return out


_INFO_FILES = ("config.npy", "imagecols.npy", "all_2d_segs.npy")


def save_folder_linetracks_with_info(folder, linetracks, config=None, imagecols=None, all_2d_segs=None):
    save_folder_linetracks(folder, linetracks)
    payload = (config, None if imagecols is None else imagecols.as_dict(), all_2d_segs)
    for name, obj in zip(_INFO_FILES, payload):
        if obj is not None:
            save_npy(os.path.join(folder, name), obj)


def read_folder_linetracks_with_info(folder):
    linetracks = read_folder_linetracks(folder)
    got = []
    for name in _INFO_FILES:
        f = os.path.join(folder, name)
        got.append(read_npy(f).item() if os.path.isfile(f) else None)
    cfg, ic, segs = got
    return linetracks, cfg, (base.ImageCollection(ic) if ic is not None else None), segs


def save_txt_imname_dict(fname, imname_dict):
    """image_list.txt: a count line, then `img_id, image_name` per image (util/io.py:157-162)."""
    _require_parent(fname)
    rows = [f"number of images, {len(imname_dict)}"] + [f"{i}, {n}" for i, n in imname_dict.items()]
    with open(fname, "w") as f:
        f.write("\n".join(rows) + "\n")


def read_txt_imname_dict(fname):
    _require(fname)
    with open(fname) as f:
        rows = [r.rstrip("\n") for r in f.readlines()]
    n = int(rows[0].split(",")[1])
    out = {}
    for r in rows[1:1 + n]:
        k, v = r.split(",", 1)
        out[int(k)] = v.strip()
    return out


def save_obj(fname, lines):
    """Wavefront .obj of 3D segments (vertices + `l` elements, util/io.py:181-199)."""
    arr = [np.asarray(l if isinstance(l, np.ndarray) else l.as_array(), float).reshape(2, 3) for l in lines]
    with open(fname, "w") as f:
        for a in arr:
            for v in a:
                f.write(f"v {v[0]} {v[1]} {v[2]}\n")
        for k in range(len(arr)):
            f.write(f"l {2 * k + 1} {2 * k + 2}\n")
