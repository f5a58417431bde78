"""CPU-only: the restatement of processors::Basic (oracle/basic_oracle.py) against OpenCV itself where cv2 is importable
(this build container), against the golden layers generated through cv2 (tests/golden/basic_masks.npz), and the library's
structuring elements (host code of libartp.so) against both."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

import cases
from art_planner_b200 import synth
from oracle import basic_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "basic_masks.npz"))


def test_circular_kernels_match_opencv():
    cv2 = pytest.importorskip("cv2")
    for size in range(0, 65):
        k = np.zeros((size, size), np.uint8)
        if size:
            cv2.circle(k, (size // 2, size // 2), size // 2, (255, 255, 255), -1)
        assert np.array_equal(bo.circular_kernel(size), k), size


def test_library_structuring_elements():
    from art_planner_b200 import capi
    lib = capi.load()
    for size in range(0, 65):
        buf = np.zeros(64 * 64, np.uint8)
        n = lib.artp_debug_circular_kernel(size, buf.ctypes.data)
        want = bo.circular_kernel(size)
        if size == 0:
            assert n == 3 and buf[:9].all()
        else:
            assert n == size and np.array_equal(buf[: size * size].reshape(size, size), (want > 0).astype(np.uint8)), size


def test_morphology_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.standard_normal((57, 43)).astype(np.float32))
    for size in (0, 1, 2, 3, 4, 7, 8, 15, 16):
        k = np.zeros((size, size), np.uint8)
        if size:
            cv2.circle(k, (size // 2, size // 2), size // 2, (255, 255, 255), -1)
        img = np.ascontiguousarray(a.T)
        assert np.array_equal(bo.erode(a, size), cv2.erode(img, k if size else None).T), size
        assert np.array_equal(bo.dilate(a, size), cv2.dilate(img, k if size else None).T), size


@pytest.mark.parametrize("case", cases.BASIC_CASES, ids=[c[0] for c in cases.BASIC_CASES])
def test_oracle_matches_opencv_golden(case, gold, maps):
    name, mk, scale, p = case
    m = maps(mk)
    trav, obs = synth.make_traversability(m, seed=13)
    h = hashlib.sha256(); [h.update(np.ascontiguousarray(x).tobytes()) for x in (m.elevation, trav, obs)]
    assert h.hexdigest() == str(gold[name + "/sha"]), "generator drift"
    masked, thr = bo.masked_elevation(m.elevation, trav, obs, m.res * scale, p)
    n = masked.size
    assert np.array_equal(np.isfinite(masked).ravel(order="F"), np.unpackbits(gold[name + "/finite"])[:n].astype(bool))
    assert np.array_equal((thr > 0.5).ravel(order="F"), np.unpackbits(gold[name + "/thr"])[:n].astype(bool))
    fin = np.isfinite(masked)
    assert np.array_equal(masked[fin], m.elevation[fin]) and 0 < fin.mean() < 1
