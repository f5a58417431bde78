"""Golden layers of processors::Basic through OpenCV itself (cv2, the library the reference calls): run in the build
container (cv2 4.13 present):  python oracle/make_golden_basic.py  -> tests/golden/basic_masks.npz"""
import hashlib, os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from art_planner_b200 import synth  # noqa: E402
from oracle import basic_oracle as bo  # noqa: E402
import cases  # noqa: E402


def cv_kernel(size):   # getCircularKernel, utils.cpp:106-111
    k = np.zeros((size, size), np.uint8)
    if size > 0:
        cv2.circle(k, (size // 2, size // 2), size // 2, (255, 255, 255), -1)
    return k


def cv_morph(fn):
    def f(mat, size):
        img = np.ascontiguousarray(np.asarray(mat, np.float32).T)      # cols x rows row-major view of the column-major matrix
        k = cv_kernel(size)
        out = fn(img, k if k.size else None)
        return np.asfortranarray(out.T)
    return f


def main():
    out = {}
    for name, mk, res_scale, p in cases.BASIC_CASES:
        m = cases.MAPS[mk]()
        trav, obs = synth.make_traversability(m, seed=13)
        masked, thr = bo.masked_elevation(m.elevation, trav, obs, m.res * res_scale, p, morph=(cv_morph(cv2.erode), cv_morph(cv2.dilate)))
        out[name + "/finite"] = np.packbits(np.isfinite(masked).ravel(order="F"))
        out[name + "/thr"] = np.packbits((thr > 0.5).ravel(order="F"))
        h = hashlib.sha256(); [h.update(np.ascontiguousarray(a).tobytes()) for a in (m.elevation, trav, obs)]
        out[name + "/sha"] = np.array(h.hexdigest())
        print(name, "traversable fraction", float(np.isfinite(masked).mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "basic_masks.npz"), **out)


if __name__ == "__main__":
    main()
