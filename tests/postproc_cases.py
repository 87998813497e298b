"""Inputs of the post-processing golden cases that are a function of the level table: the point centres
(reference: shifts = arange * stride, x fastest) and strides."""
import numpy as np

TRANSFORMS = ["minmax", "partial_minmax", "moment"]


def centers_of(levels):
    out = []
    for h, w, s in levels:
        ys, xs = np.meshgrid(np.arange(h, dtype=np.float32) * np.float32(s), np.arange(w, dtype=np.float32) * np.float32(s),
                             indexing="ij")
        out.append(np.stack([xs.reshape(-1), ys.reshape(-1)], 1).astype(np.float32))
    return out


def unpack(c):
    lv = [tuple(int(v) for v in r) for r in c["levels"]]
    K, topk, det = [int(v) for v in c["cfg"]]
    thr, nms, mt0, mt1 = [float(v) for v in c["fcfg"]]
    return dict(levels=lv, K=K, topk=topk, max_det=det, score_thresh=thr, nms_thresh=nms, moment_transfer=(mt0, mt1),
                transform=TRANSFORMS[int(c["transform"])], image_size=tuple(int(v) for v in c["image_size"]),
                cls=[c["cls%d" % l] for l in range(len(lv))], pts=[c["pts%d" % l] for l in range(len(lv))],
                centers=centers_of(lv), strides=[float(s) for _, _, s in lv])
