"""On-disk formats at the boundary of the hot path: byte layout of the text files as the
reference writes them (util/io.py, base/linetrack.cc:133-213) and round trips through the readers."""
import os

import numpy as np

import limap.base as base
import limap.util.io as limapio


def _track():
    t = base.LineTrack()
    t.line = base.Line3d(np.array([1.0, 2.5, -3.0]), np.array([4.0, 5.0, 6.125]))
    t.image_id_list, t.line_id_list = [3, 7, 7], [10, 0, 2]
    t.line2d_list = [base.Line2d(np.array([1.0, 2.0]), np.array([3.0, 4.5])) for _ in range(3)]
    t.node_id_list = [5, 6, 9]
    t.score_list = [0.5, 1.0, 2.25]
    t.line3d_list = [base.Line3d(np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 1.0])) for _ in range(3)]
    return t


def test_linetrack_write_layout_and_read(tmp_path):
    t = _track()
    f = str(tmp_path / "track_0.txt")
    t.Write(f)
    txt = open(f).read().split("\n")
    assert txt[0] == "1.0000000000 2.5000000000 -3.0000000000 4.0000000000 5.0000000000 6.1250000000 "
    assert txt[1] == "3 2"
    assert txt[2] == "image_id_list 3 7 7 "
    assert txt[3] == "line_id_list 10 0 2 "
    assert txt[4] == "line2d_list"
    assert txt[5] == "1.0000000000 2.0000000000 3.0000000000 4.5000000000 "
    assert txt[8] == "node_id_list 5 6 9 "
    assert txt[9] == "score_list 0.5000000000 1.0000000000 2.2500000000 "
    assert txt[10] == "line3d_list"
    assert txt[11] == "0.0000000000 0.0000000000 1.0000000000 1.0000000000 0.0000000000 1.0000000000 "
    assert txt[14] == "END"
    r = base.LineTrack()
    r.Read(f)
    assert r.image_id_list == t.image_id_list and r.line_id_list == t.line_id_list
    assert r.node_id_list == t.node_id_list and r.score_list == t.score_list
    assert np.allclose(r.line.start, t.line.start) and np.allclose(r.line.end, t.line.end)
    assert np.allclose(r.line2d_list[2].end, [3.0, 4.5]) and np.allclose(r.line3d_list[1].end, [1.0, 0.0, 1.0])
    # NaN endpoints are written as zeros (linetrack.cc:139-153); a track without aux lists ends after line2d_list
    t2 = base.LineTrack()
    t2.line = base.Line3d(np.array([np.nan, 0, 0]), np.array([1.0, 1, 1]))
    t2.image_id_list, t2.line_id_list = [1], [2]
    t2.line2d_list = [base.Line2d(np.array([0.0, 0.0]), np.array([1.0, 1.0]))]
    t2.Write(f)
    lines = open(f).read().split("\n")
    assert lines[0].startswith("0.0000000000 0.0000000000 0.0000000000 1.0000000000")
    assert lines[6] == "END"
    r2 = base.LineTrack()
    r2.Read(f)
    assert r2.count_lines() == 1 and r2.line_id_list == [2]


def test_folder_and_single_file_tracks(tmp_path):
    tracks = [_track(), _track()]
    tracks[1].image_id_list = [1, 2, 3]
    folder = str(tmp_path / "finaltracks")
    cams = {0: base.Camera("PINHOLE", [500.0, 510.0, 320.0, 240.0], 0, (480, 640))}
    imgs = {i: base.CameraImage(0, base.CameraPose(np.array([1.0, 0, 0, 0]), np.array([0.1 * i, 0, 0])), f"{i}.png")
            for i in (1, 2, 3, 7)}
    imagecols = base.ImageCollection(cams, imgs)
    segs = {1: np.array([[0.0, 1, 2, 3]]), 2: np.zeros((0, 4))}
    limapio.save_folder_linetracks_with_info(folder, tracks, config={"a": 1}, imagecols=imagecols, all_2d_segs=segs)
    assert sorted(os.listdir(folder)) == ["all_2d_segs.npy", "config.npy", "imagecols.npy", "track_0.txt", "track_1.txt"]
    lt, cfg, ic, s2 = limapio.read_folder_linetracks_with_info(folder)
    assert len(lt) == 2 and cfg == {"a": 1} and ic.get_img_ids() == [1, 2, 3, 7]
    assert np.allclose(ic.campose(3).tvec, [0.3, 0, 0]) and np.array_equal(s2[1], segs[1])
    assert lt[1].image_id_list == [1, 2, 3]
    # alltracks.txt: the n_visible_views filter and the row layout of util/io.py:259-293
    f = str(tmp_path / "out" / "alltracks.txt")
    limapio.save_txt_linetracks(f, tracks, n_visible_views=3)
    rows = open(f).read().split("\n")
    assert rows[0] == "1" and rows[1] == "0 3 3"
    assert rows[2] == "1.0000000000 " + " " * 18 + "2.5000000000 " + " " * 18 + "-3.0000000000"
    assert rows[4] == "1 2 3 " and rows[5] == "10 0 2 "
    back = limapio.read_txt_linetracks(f)
    assert len(back) == 1 and back[0][1] == [1, 2, 3] and np.allclose(back[0][0][1], [4.0, 5.0, 6.125])


def test_segments_metainfos_matches(tmp_path):
    d = str(tmp_path)
    segs = np.array([[0.5, 1.25, 100.0, 200.75], [3.0, 4.0, 5.0, 6.0]])
    limapio.save_txt_segments(d, 12, segs)
    assert open(os.path.join(d, "segments_12.txt")).read() == "2\n0.5 1.25 100.0 200.75\n3.0 4.0 5.0 6.0\n"
    assert limapio.exists_txt_segments(d, 12) and not limapio.exists_txt_segments(d, 13)
    assert np.array_equal(limapio.read_txt_segments(d, 12), segs)
    os.makedirs(os.path.join(d, "segs"))
    limapio.save_txt_segments(os.path.join(d, "segs"), 3, segs[:1])
    limapio.save_txt_segments(os.path.join(d, "segs"), 40, segs)
    allsegs = limapio.read_all_segments_from_folder(os.path.join(d, "segs"))
    assert sorted(allsegs) == [3, 40] and allsegs[40].shape == (2, 4)
    neighbors = {0: [1, 2], 1: [0], 2: []}
    ranges = (np.array([-1.0, -2.0, -3.5]), np.array([1.0, 2.0, 3.5]))
    f = os.path.join(d, "metainfos.txt")
    limapio.save_txt_metainfos(f, neighbors, ranges)
    assert open(f).read() == ("number of images, 3\nx-range, -1.0, 1.0\ny-range, -2.0, 2.0\nz-range, -3.5, 3.5\n"
                              "image 0, 1, 2\nimage 1, 0\nimage 2\n")
    n2, r2 = limapio.read_txt_metainfos(f)
    assert n2 == neighbors and np.array_equal(r2[0], ranges[0]) and np.array_equal(r2[1], ranges[1])
    m = {1: np.array([[0, 1], [2, 3]], np.int32), 5: np.zeros((0, 2), np.int32)}
    limapio.save_match(d, 0, m)
    assert os.path.basename(limapio.get_match_filename(d, 0)) == "matches_0.npy"
    back = limapio.read_match(d, 0)
    assert sorted(back) == [1, 5] and np.array_equal(back[1], m[1]) and back[5].shape == (0, 2)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")


def test_golden_files_written_by_the_reference(tmp_path):
    """tests/golden/io/* were written by the reference's own util/io.py (tests/golden/make_io_golden.py): this
    repository's writers reproduce them byte for byte from the same inputs, and its readers parse them."""
    z = np.load(os.path.join(GOLDEN, "inputs.npz"), allow_pickle=True)
    d = str(tmp_path)
    for i in z["seg_ids"]:
        i = int(i)
        limapio.save_txt_segments(d, i, z[f"segs_{i}"])
        assert open(os.path.join(d, f"segments_{i}.txt")).read() == open(os.path.join(GOLDEN, f"segments_{i}.txt")).read()
        got = limapio.read_txt_segments(GOLDEN, i)
        if len(z[f"segs_{i}"]):
            assert np.array_equal(got, z[f"segs_{i}"])
    nb, rg = limapio.read_txt_metainfos(os.path.join(GOLDEN, "metainfos.txt"))
    assert nb == {3: [11, 12], 11: [3], 12: []}
    assert np.array_equal(rg[0], z["range_lo"]) and np.array_equal(rg[1], z["range_hi"])
    limapio.save_txt_metainfos(os.path.join(d, "metainfos.txt"), nb, rg)
    assert open(os.path.join(d, "metainfos.txt")).read() == open(os.path.join(GOLDEN, "metainfos.txt")).read()
    tracks = []
    for L, img, lid in zip(z["track_lines"], z["track_img"], z["track_lid"]):
        t = base.LineTrack()
        t.line = base.Line3d(L[:3], L[3:])
        t.image_id_list, t.line_id_list = list(img), list(lid)
        t.line2d_list = [base.Line2d() for _ in img]
        tracks.append(t)
    for nv in (1, 3):
        f = os.path.join(d, f"alltracks_nv{nv}.txt")
        limapio.save_txt_linetracks(f, tracks, n_visible_views=nv)
        assert open(f).read() == open(os.path.join(GOLDEN, f"alltracks_nv{nv}.txt")).read()
    back = limapio.read_txt_linetracks(os.path.join(GOLDEN, "alltracks_nv1.txt"))
    assert len(back) == 3 and back[2][1] == list(z["track_img"][2])
    m = limapio.read_match(GOLDEN, 3)
    assert sorted(m) == [11, 12] and m[11].shape == (7, 2) and m[12].shape == (0, 2)
