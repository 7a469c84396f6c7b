"""limap.visualize as far as the triangulation runner needs it: the track report (the reference's scalar fingerprint of
a run, visualize/trackvis/base.py:20-52) and the line arrays for the .obj dump. Rendering (Open3D / pyvista) is out
of scope."""
import logging

import numpy as np

_log = logging.getLogger("limap_b200")


class BaseTrackVisualizer:
    def __init__(self, tracks):
        self.tracks = list(tracks)
        self.counts = [t.count_images() for t in self.tracks]
        self.counts_lines = [t.count_lines() for t in self.tracks]
        self.lines = [t.line for t in self.tracks]

    def track_report(self):
        """(N2, N4, N6, N8, N10, N20, N50): tracks supported by at least k images."""
        c = np.asarray(self.counts, dtype=np.int64)
        return tuple(int((c >= k).sum()) for k in (2, 4, 6, 8, 10, 20, 50))

    def report_stats(self):
        msg = "[Track Report] (N2, N4, N6, N8, N10, N20, N50) = ({})".format(", ".join(map(str, self.track_report())))
        _log.info(msg)
        print(msg)
        return msg

    def report_avg_supports(self, n_visible_views=4):
        c, cl = np.asarray(self.counts), np.asarray(self.counts_lines)
        sel = c >= n_visible_views
        if sel.any():
            _log.info("average supporting images (>= %d): %d / %d = %.2f", n_visible_views, c[sel].sum(), sel.sum(), c[sel].mean())
            _log.info("average supporting lines (>= %d): %d / %d = %.2f", n_visible_views, cl[sel].sum(), sel.sum(), cl[sel].mean())

    def report(self):
        self.report_stats()
        self.report_avg_supports(3)
        self.report_avg_supports(4)

    def get_counts_np(self):
        return np.asarray(self.counts)

    def get_lines_np(self, n_visible_views=0):
        out = [l.as_array() for l, c in zip(self.lines, self.counts) if c >= n_visible_views]
        return np.asarray(out) if out else np.zeros((0, 2, 3))

    def vis_all_lines(self, *a, **k):
        raise NotImplementedError("rendering is outside the hot path (SURVEY.md §2)")

    vis_reconstruction = vis_all_lines


Open3DTrackVisualizer = PyVistaTrackVisualizer = BaseTrackVisualizer


def visualize_line_track(*a, **k):
    raise NotImplementedError("rendering is outside the hot path (SURVEY.md §2)")
