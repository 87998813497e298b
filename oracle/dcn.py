"""ctypes front-end for oracle/dcn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``forward`` / ``backward`` take and return float32 numpy arrays in the reference's
layouts (see dcn_oracle.c header; detectron2/layers/deform_conv.py:15-135, :179-301
for the argument meaning).  ``build()`` compiles the C file on demand.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_dcn.so")
_lib = None


def build(force=False):
    """Compile dcn_oracle.c -> liboracle_dcn.so (gcc, OpenMP if available)."""
    src = os.path.join(_HERE, "dcn_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    base = ["-O3", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", _SO, src, "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True, cwd=_HERE)
                return _SO
            except (subprocess.CalledProcessError, FileNotFoundError) as e:  # try next
                last = e
    raise RuntimeError("could not build the DCN oracle: %r" % (last,))


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def out_hw(H, W, KH, KW, stride, padding, dilation):
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    return ((H + 2 * ph - (dh * (KH - 1) + 1)) // sh + 1, (W + 2 * pw - (dw * (KW - 1) + 1)) // sw + 1)


def forward(x, offset, weight, mask=None, bias=None, stride=1, padding=0, dilation=1, groups=1,
            deformable_groups=1):
    x, offset, weight, mask, bias = map(_f32, (x, offset, weight, mask, bias))
    N, C, H, W = x.shape
    O, _, KH, KW = weight.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    Ho, Wo = out_hw(H, W, KH, KW, stride, padding, dilation)
    out = np.empty((N, O, Ho, Wo), np.float32)
    rc = _load().dcn_oracle_forward(_p(x), _p(offset), _p(mask), _p(weight), _p(bias), _p(out),
                                    N, C, H, W, O, KH, KW, sh, sw, ph, pw, dh, dw, groups,
                                    deformable_groups)
    if rc != 0:
        raise ValueError("dcn_oracle_forward rc=%d" % rc)
    return out


def backward(x, offset, weight, grad_out, mask=None, with_bias=False, stride=1, padding=0,
             dilation=1, groups=1, deformable_groups=1):
    """Returns dict(grad_x, grad_offset, grad_mask|None, grad_weight, grad_bias|None)."""
    x, offset, weight, mask, grad_out = map(_f32, (x, offset, weight, mask, grad_out))
    N, C, H, W = x.shape
    O, _, KH, KW = weight.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    gx = np.empty_like(x)
    go = np.empty_like(offset)
    gm = np.empty_like(mask) if mask is not None else None
    gw = np.empty_like(weight)
    gb = np.empty((O,), np.float32) if with_bias else None
    rc = _load().dcn_oracle_backward(_p(x), _p(offset), _p(mask), _p(weight), _p(grad_out), _p(gx),
                                     _p(go), _p(gm), _p(gw), _p(gb), N, C, H, W, O, KH, KW, sh, sw,
                                     ph, pw, dh, dw, groups, deformable_groups)
    if rc != 0:
        raise ValueError("dcn_oracle_backward rc=%d" % rc)
    return dict(grad_x=gx, grad_offset=go, grad_mask=gm, grad_weight=gw, grad_bias=gb)
