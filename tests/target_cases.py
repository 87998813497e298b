"""Inputs of the target-assignment golden cases that are a pure function of the level table (kept out of the .npz):
the FCOS location grid of compute_locations_per_level (fcos/utils.py:78-92: arange(0, w*s, s) + s // 2, x fastest),
the per-level sizes of interest of FCOSRepPoints.get_ground_truth (fcos_rpd_s1_topk.py:322-337) and the strides."""
import numpy as np

INF = 100000000
SIZES = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, INF]]


def locations_of(levels):
    out = []
    for h, w, s in levels:
        ys, xs = np.meshgrid(np.arange(0, h * s, s, dtype=np.float32), np.arange(0, w * s, s, dtype=np.float32), indexing="ij")
        out.append(np.stack([xs.reshape(-1), ys.reshape(-1)], 1).astype(np.float32) + np.float32(s // 2))
    return out


def soi_of(levels):
    return np.concatenate([np.tile(np.array(SIZES[i], np.float32)[None], (h * w, 1)) for i, (h, w, s) in enumerate(levels)])


def strides_of(levels):
    return [int(s) for _, _, s in levels]


def num_points_of(levels):
    return [int(h * w) for h, w, _ in levels]
