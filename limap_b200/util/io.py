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


def check_directory(fname):  # util/io.py:12-15
    d = os.path.dirname(fname)
    if d and not os.path.exists(d):
        raise ValueError(f"Error! Base directory {d} does not exist!")


def check_path(fname):  # util/io.py:18-20
    if not os.path.exists(fname):
        raise ValueError(f"Error! File {fname} does not exist!")


def check_makedirs(folder):
    if not os.path.exists(folder):
        os.makedirs(folder)


def delete_folder(folder):
    if os.path.exists(folder):
        shutil.rmtree(folder)


def save_npy(fname, nparray):  # util/io.py:39-42
    check_directory(fname)
    with open(fname, "wb") as f:
        np.save(f, np.array(nparray, dtype=object))


def read_npy(fname):  # util/io.py:45-49
    check_path(fname)
    with open(fname, "rb") as f:
        return np.load(f, allow_pickle=True)


def save_npz(fname, dic):
    check_directory(fname)
    np.savez(fname, **dic)


def read_npz(fname):
    check_path(fname)
    return np.load(fname, allow_pickle=True)


# ---- matches (line2d/base_matcher.py:80-117) ---------------------------------------------------------------
def get_match_filename(matches_folder, idx):
    return os.path.join(matches_folder, f"matches_{idx}.npy")


def save_match(matches_folder, idx, matches):
    save_npy(get_match_filename(matches_folder, idx), matches)


def read_match(matches_folder, idx):
    return read_npy(get_match_filename(matches_folder, idx)).item()


# ---- neighbors + ranges (util/io.py:87-131) -----------------------------------------------------------------
def save_txt_metainfos(fname, neighbors, ranges):
    check_directory(fname)
    with open(fname, "w") as f:
        f.write(f"number of images, {len(neighbors)}\n")
        f.write(f"x-range, {ranges[0][0]}, {ranges[1][0]}\n")
        f.write(f"y-range, {ranges[0][1]}, {ranges[1][1]}\n")
        f.write(f"z-range, {ranges[0][2]}, {ranges[1][2]}\n")
        for img_id, neighbor in neighbors.items():
            str_ = f"image {img_id}"
            for ng_idx in neighbor:
                str_ += f", {ng_idx}"
            f.write(str_ + "\n")


def read_txt_metainfos(fname):
    check_path(fname)
    with open(fname) as f:
        txt_lines = f.readlines()
    n_images = int(txt_lines[0].strip().split(",")[1])
    ranges = (np.zeros(3), np.zeros(3))
    for axis in range(3):
        k = txt_lines[1 + axis].strip().split(",")[1:]
        ranges[0][axis], ranges[1][axis] = float(k[0]), float(k[1])
    neighbors = {}
    for row in txt_lines[4:4 + n_images]:
        k = row.strip().split(",")
        neighbors[int(k[0][6:])] = [int(kk) for kk in k[1:]]
    return neighbors, ranges


# ---- 2D segments (util/io.py:436-474) ------------------------------------------------------------------------
def exists_txt_segments(folder, img_id):
    return os.path.exists(os.path.join(folder, f"segments_{img_id}.txt"))


def save_txt_segments(folder, img_id, segs):
    fname = os.path.join(folder, f"segments_{img_id}.txt")
    segs = np.asarray(segs)
    with open(fname, "w") as f:
        f.write(f"{segs.shape[0]}\n")
        for line in segs:
            f.write(f"{line[0]} {line[1]} {line[2]} {line[3]}\n")


def read_txt_segments(folder, img_id):
    check_path(folder)
    fname = os.path.join(folder, f"segments_{img_id}.txt")
    with open(fname) as f:
        txt_lines = f.readlines()
    n_segments = int(txt_lines[0].strip())
    assert n_segments + 1 == len(txt_lines)
    segs = [[float(kk) for kk in row.strip().split(" ")] for row in txt_lines[1:]]
    return np.array(segs)


def read_all_segments_from_folder(folder):
    all_2d_segs = {}
    for fname in os.listdir(folder):
        img_id = int(fname[9:-4])
        all_2d_segs[img_id] = read_txt_segments(folder, img_id)
    return all_2d_segs


# ---- line tracks (util/io.py:259-346) ----------------------------------------------------------------------------
def save_txt_linetracks(fname, linetracks, n_visible_views=4):
    d = os.path.dirname(fname)
    if d and not os.path.exists(d):
        os.makedirs(d)
    linetracks = [track for track in linetracks if track.count_images() >= n_visible_views]
    with open(fname, "w") as f:
        f.write(f"{len(linetracks)}\n")
        for track_id, track in enumerate(linetracks):
            f.write(f"{track_id} {track.count_lines()} {track.count_images()}\n")
            # the reference's f-string continues over source lines, which leaves the indentation inside the row
            pad = " " * 18
            f.write(f"{track.line.start[0]:.10f} {pad}{track.line.start[1]:.10f} {pad}{track.line.start[2]:.10f}\n")
            f.write(f"{track.line.end[0]:.10f} {pad}{track.line.end[1]:.10f} {pad}{track.line.end[2]:.10f}\n")
            f.write("".join(f"{i} " for i in track.image_id_list) + "\n")
            f.write("".join(f"{i} " for i in track.line_id_list) + "\n")


def read_txt_linetracks(fname):
    """Inverse of save_txt_linetracks: [(line (2,3), image_id_list, line_id_list)] (whitespace tolerant)."""
    check_path(fname)
    with open(fname) as f:
        tok = f.read().split()
    n, p, out = int(tok[0]), 1, []
    for _ in range(n):
        n_lines = int(tok[p + 1])
        p += 3
        line = np.array([float(x) for x in tok[p:p + 6]]).reshape(2, 3)
        p += 6
        img = [int(x) for x in tok[p:p + n_lines]]
        p += n_lines
        lid = [int(x) for x in tok[p:p + n_lines]]
        p += n_lines
        out.append((line, img, lid))
    return out


def save_folder_linetracks(folder, linetracks):
    if os.path.exists(folder):
        shutil.rmtree(folder)
    os.makedirs(folder)
    for track_id, track in enumerate(linetracks):
        track.Write(os.path.join(folder, f"track_{track_id}.txt"))


def read_folder_linetracks(folder):
    check_path(folder)
    n_tracks = sum(1 for fname in os.listdir(folder) if fname[-4:] == ".txt" and fname[:5] == "track")
    linetracks = []
    for track_id in range(n_tracks):
        track = base.LineTrack()
        track.Read(os.path.join(folder, f"track_{track_id}.txt"))
        linetracks.append(track)
    return linetracks


def save_folder_linetracks_with_info(folder, linetracks, config=None, imagecols=None, all_2d_segs=None):
    save_folder_linetracks(folder, linetracks)
    if config is not None:
        save_npy(os.path.join(folder, "config.npy"), config)
    if imagecols is not None:
        save_npy(os.path.join(folder, "imagecols.npy"), imagecols.as_dict())
    if all_2d_segs is not None:
        save_npy(os.path.join(folder, "all_2d_segs.npy"), all_2d_segs)


def read_folder_linetracks_with_info(folder):
    linetracks = read_folder_linetracks(folder)
    cfg, imagecols, all_2d_segs = None, None, None
    if os.path.isfile(os.path.join(folder, "config.npy")):
        cfg = read_npy(os.path.join(folder, "config.npy")).item()
    if os.path.isfile(os.path.join(folder, "imagecols.npy")):
        imagecols = base.ImageCollection(read_npy(os.path.join(folder, "imagecols.npy")).item())
    if os.path.isfile(os.path.join(folder, "all_2d_segs.npy")):
        all_2d_segs = read_npy(os.path.join(folder, "all_2d_segs.npy")).item()
    return linetracks, cfg, imagecols, all_2d_segs
