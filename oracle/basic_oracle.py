"""CPU oracle of processors::Basic::setMaskedElevationAndTraversability (TEST INFRASTRUCTURE ONLY).

Restates art_planner/src/map/processors/basic.cpp:42-106 on numpy arrays (inputs: the INPAINTED elevation /
traversability layers and the "observed" layer) with the morphology helpers of art_planner/src/utils.cpp:106-209.
OpenCV is not in the reference tree (version unpinned): `circular_kernel` restates cv::circle's integer
midpoint rasterisation (imgproc/src/drawing.cpp) and `erode` / `dilate` cv::erode / cv::dilate with the default anchor
(element centre) and border (outside cells do not take part). Where cv2 is importable (this build container: 4.13)
tests/test_basic_cpu.py pins both against the library itself, and oracle/make_golden_basic.py generates the committed
golden layers through cv2 -- the same calls the reference makes.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass(frozen=True)
class BasicParams:
    """params.h:23-35 with the shipped values of art_planner_ros/config/params.yaml:8-18."""
    traversability_thres: float = 0.15
    unknown_space_untraversable: bool = True
    foothold_margin: float = 0.3
    foothold_margin_max_hole_size: float = 0.3
    foothold_margin_max_drop: float = 0.3
    foothold_margin_max_drop_search_radius: float = 0.16
    foothold_margin_min_step: float = 0.3
    foothold_size: float = 0.1


def circular_kernel(size: int) -> np.ndarray:
    """getCircularKernel (utils.cpp:106-111): size x size uint8, filled circle of radius size//2 about (size//2, size//2)."""
    if size <= 0:
        return np.zeros((0, 0), np.uint8)
    k = np.zeros((size, size), np.uint8)
    r = size // 2
    cx = cy = r

    def span(y, x0, x1):
        if 0 <= y < size:
            x0, x1 = max(x0, 0), min(x1, size - 1)
            if x0 <= x1:
                k[y, x0:x1 + 1] = 255
    err, dx, dy, plus, minus = 0, r, 0, 1, (r << 1) - 1
    while dx >= dy:
        span(cy - dy, cx - dx, cx + dx); span(cy + dy, cx - dx, cx + dx)
        span(cy - dx, cx - dy, cx + dy); span(cy + dx, cx - dy, cx + dy)
        dy += 1; err += plus; plus += 2
        mask = -1 if err > 0 else 0          # (err <= 0) - 1
        err -= minus & mask; dx += mask; minus -= mask & 2
    return k


def _morph(mat: np.ndarray, size: int, dilate_: bool) -> np.ndarray:
    """cv::erode / cv::dilate of the cols x rows row-major VIEW of a column-major rows x cols matrix (utils.cpp:120-123)
    with getCircularKernel(size); an empty element means OpenCV's 3 x 3 box."""
    k = circular_kernel(size)
    if k.size == 0:
        k = np.full((3, 3), 255, np.uint8)
    a = k.shape[0] // 2
    img = np.ascontiguousarray(mat.T)          # image row = j, column = i
    fill = -np.inf if dilate_ else np.inf
    out = np.full(img.shape, fill, np.float32)
    H, W = img.shape
    for kr in range(k.shape[0]):
        for kc in range(k.shape[1]):
            if not k[kr, kc]:
                continue
            dr, dc = kr - a, kc - a            # dst(r, c) takes src(r + dr, c + dc)
            r0, r1 = max(0, -dr), min(H, H - dr)
            c0, c1 = max(0, -dc), min(W, W - dc)
            if r0 >= r1 or c0 >= c1:
                continue
            src = img[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
            dst = out[r0:r1, c0:c1]
            np.maximum(dst, src, out=dst) if dilate_ else np.minimum(dst, src, out=dst)
    return np.asfortranarray(out.T)


def erode(mat, size):
    return _morph(mat, size, False)


def dilate(mat, size):
    return _morph(mat, size, True)


def masked_elevation(elevation, traversability, observed, res: float, p: BasicParams = BasicParams(), morph=None):
    """basic.cpp:42-106 after the inpainting. Returns (elevation_masked, traversability_thresholded), float32 F-order.
    morph: optional (erode, dilate) pair, e.g. the cv2-backed one of make_golden_basic.py."""
    er, di = morph if morph else (erode, dilate)
    E = np.asfortranarray(elevation, dtype=np.float32)
    trav = np.asfortranarray(traversability, dtype=np.float32)
    if p.unknown_space_untraversable:
        trav = np.where(np.asarray(observed, dtype=np.float32) > np.float32(0.5), trav, np.float32(0))
    T0 = np.where(trav > np.float32(p.traversability_thres), np.float32(1), np.float32(0)).astype(np.float32)
    foothold = int(math.ceil(p.foothold_size / res))
    margin = int(math.ceil(2 * p.foothold_margin / res))
    hole = int(math.floor(p.foothold_margin_max_hole_size / res))
    search = int(math.ceil(2 * p.foothold_margin_max_drop_search_radius / res))
    S = er(di(T0, hole), hole)                                           # dilateAndErode: close holes
    hole_mask = (E - er(E, search)) > np.float32(p.foothold_margin_max_drop)
    S = np.where(hole_mask, T0, S)
    wall_mask = (di(E, margin) - E) > np.float32(p.foothold_margin_min_step)
    S = np.where(wall_mask, np.float32(1), S).astype(np.float32)
    S = er(S, margin)
    S = np.where((T0 < np.float32(0.5)) | wall_mask, T0, S).astype(np.float32)
    S = di(er(S, foothold), foothold)                                    # erodeAndDilate: remove small patches
    S = np.where(T0 < np.float32(0.5), T0, S).astype(np.float32)
    masked = np.where(S > np.float32(0.5), E, np.float32(-np.inf)).astype(np.float32)
    return np.asfortranarray(masked), np.asfortranarray(S)
