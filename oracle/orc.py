"""ctypes binding of the CPU oracle libraries (TEST INFRASTRUCTURE ONLY).

    Oracle("port")       -> oracle/liborc_port.so     (C restatement, artp_oracle.c)
    Oracle("reference")  -> oracle/_ref/liborc_ref.so (the reference's own compiled ODE + ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liborc_port.so")
REF_SO = os.path.join(HERE, "_ref", "liborc_ref.so")


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "torso_length", "torso_width", "torso_height", "torso_off_x", "torso_off_y", "torso_off_z",
        "feet_off_x", "feet_off_y", "feet_off_z", "reach_x", "reach_y", "reach_z")] + [
        ("unknown_space_untraversable", C.c_int), ("use_directional_cost", C.c_int),
        ("max_lon_vel", C.c_double), ("max_lat_vel", C.c_double), ("max_ang_vel", C.c_double)]


class OrcSamplerMap(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("elevation", "normal_x", "normal_y", "normal_z", "plane_fit_std_dev",
                                          "cum_prob", "cum_prob_rowwise")] + [
        ("rows", C.c_int), ("cols", C.c_int), ("res", C.c_double), ("cx", C.c_double), ("cy", C.c_double)]


class OrcSamplerParams(C.Structure):
    _fields_ = [("max_roll_pert", C.c_double), ("max_pitch_pert", C.c_double), ("sample_from_distribution", C.c_int),
                ("low", C.c_double * 2), ("high", C.c_double * 2), ("reach_z", C.c_double)]


def sample_states(m, layers, sp, reach_z: float, u):
    """SE3FromSE2Sampler::sampleUniform restated (artp_oracle.c, port library only): u [n, 6] uniforms ->
    (states [n, 7], rowcol [n, 2])."""
    if not os.path.exists(PORT_SO):
        build("port")
    lib = C.CDLL(PORT_SO)
    lib.orc_sample_states.argtypes = [C.POINTER(OrcSamplerMap), C.POINTER(OrcSamplerParams), C.c_void_p, C.c_size_t,
                                      C.c_void_p, C.c_void_p]
    keep = [np.asfortranarray(a, dtype=np.float32) for a in (
        m.elevation, layers.normal_x, layers.normal_y, layers.normal_z, layers.plane_fit_std_dev, layers.cum_prob)]
    keep.append(np.ascontiguousarray(layers.cum_prob_rowwise, dtype=np.float32))
    sm = OrcSamplerMap(*[a.ctypes.data for a in keep], m.rows, m.cols, float(m.res), float(m.cx), float(m.cy))
    pp = OrcSamplerParams(float(sp.max_roll_pert), float(sp.max_pitch_pert), int(sp.sample_from_distribution),
                          (C.c_double * 2)(*sp.low), (C.c_double * 2)(*sp.high), float(reach_z))
    uu = np.ascontiguousarray(u, dtype=np.float64)
    n = uu.shape[0]
    states = np.zeros((n, 7), np.float64)
    rc = np.zeros((n, 2), np.int32)
    assert lib.orc_sample_states(C.byref(sm), C.byref(pp), uu.ctypes.data, n, states.ctypes.data, rc.ctypes.data) == 0
    return states, rc


def estimate_normals(m, estimation_radius: float):
    """art_planner::estimateNormals restated (artp_oracle.c): returns (normal_x, normal_y, normal_z, plane_fit_std_dev),
    float32 Fortran-order [rows, cols]."""
    if not os.path.exists(PORT_SO):
        build("port")
    lib = C.CDLL(PORT_SO)
    lib.orc_estimate_normals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    e = np.asfortranarray(m.elevation, dtype=np.float32)
    outs = [np.zeros(e.shape, np.float32, order="F") for _ in range(4)]
    rc = lib.orc_estimate_normals(e.ctypes.data, e.shape[0], e.shape[1], float(m.res), float(m.cx), float(m.cy),
                                  float(estimation_radius), *[a.ctypes.data for a in outs])
    assert rc == 0
    return tuple(outs)


def compute_cdf(prob):
    """computeCumulativeProbabilityDistribution restated (artp_oracle.c): (cum_prob [rows, cols] F-order, cum_row [rows])."""
    if not os.path.exists(PORT_SO):
        build("port")
    lib = C.CDLL(PORT_SO)
    lib.orc_compute_cdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    p = np.asfortranarray(prob, dtype=np.float32)
    cum = np.zeros(p.shape, np.float32, order="F")
    row = np.zeros(p.shape[0], np.float32)
    with np.errstate(all="ignore"):
        assert lib.orc_compute_cdf(p.ctypes.data, p.shape[0], p.shape[1], cum.ctypes.data, row.ctypes.data) == 0
    return cum, row


def build(kind: str = "port", quiet: bool = True) -> None:
    """Compile the oracle library with oracle/Makefile (building the checker is not using it)."""
    target = "port" if kind == "port" else "ref"
    subprocess.run(["make", "-s", "-C", HERE, "-j8", target], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def available(kind: str) -> bool:
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self, params, kind: str = "port"):
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            if kind == "port" or os.path.isdir("/root/reference/ode"):
                build(kind)
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        self.kind = kind
        lib = C.CDLL(path)
        self.lib = lib
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(OrcParams)]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_set_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                    C.c_double, C.c_double]
        lib.orc_box_collide.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_void_p]
        lib.orc_check_poses.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.orc_check_motions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                          C.c_void_p, C.c_void_p]
        lib.orc_path_length_cost.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orc_check_edge_interiors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]
        lib.orc_check_poses_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        lib.orc_check_motions_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
        lib.orc_kind.restype = C.c_char_p
        p = OrcParams()
        for name, _ in OrcParams._fields_:
            v = getattr(params, name)
            setattr(p, name, int(v) if name in ("unknown_space_untraversable", "use_directional_cost") else float(v))
        self.h = lib.orc_create(C.byref(p))
        assert lib.orc_kind().decode() == ("port" if kind == "port" else "reference")

    def close(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, m) -> None:
        e = np.asfortranarray(m.elevation, dtype=np.float32)
        k = np.asfortranarray(m.elevation_masked, dtype=np.float32)
        self._keep = (e, k)
        rc = self.lib.orc_set_map(self.h, e.ctypes.data, k.ctypes.data, e.shape[0], e.shape[1],
                                  float(m.res), float(m.cx), float(m.cy))
        assert rc == 0

    def box_collide(self, which: int, origins, rots, want_zone=False):
        o = np.ascontiguousarray(origins, dtype=np.float32)
        r = np.ascontiguousarray(rots, dtype=np.float32)
        n = o.shape[0]
        hit = np.zeros(n, np.uint8)
        zv = np.zeros(n, np.uint32)
        rc = self.lib.orc_box_collide(self.h, which, o.ctypes.data, r.ctypes.data, n, hit.ctypes.data,
                                      zv.ctypes.data)
        assert rc == 0
        return (hit, zv) if want_zone else hit

    def check_poses(self, states, want_zone=False):
        s = _f64(states)
        n = s.shape[0]
        valid = np.zeros(n, np.uint8)
        zv = np.zeros(n, np.uint32)
        rc = self.lib.orc_check_poses(self.h, s.ctypes.data, n, valid.ctypes.data, zv.ctypes.data)
        assert rc == 0
        return (valid, zv) if want_zone else valid

    def check_poses_mt(self, states, n_threads: int):
        s = _f64(states)
        n = s.shape[0]
        valid = np.zeros(n, np.uint8)
        rc = self.lib.orc_check_poses_mt(self.h, s.ctypes.data, n, valid.ctypes.data, int(n_threads))
        assert rc == 0
        return valid

    def check_motions(self, s1, s2, n_steps: int, want_zone=False):
        a, b = _f64(s1), _f64(s2)
        n = a.shape[0]
        valid = np.zeros(n, np.uint8)
        zv = np.zeros(n, np.uint32)
        rc = self.lib.orc_check_motions(self.h, a.ctypes.data, b.ctypes.data, n, int(n_steps),
                                        valid.ctypes.data, zv.ctypes.data)
        assert rc == 0
        return (valid, zv) if want_zone else valid

    def check_motions_mt(self, s1, s2, n_steps: int, n_threads: int):
        a, b = _f64(s1), _f64(s2)
        n = a.shape[0]
        valid = np.zeros(n, np.uint8)
        rc = self.lib.orc_check_motions_mt(self.h, a.ctypes.data, b.ctypes.data, n, int(n_steps), valid.ctypes.data,
                                           int(n_threads))
        assert rc == 0
        return valid

    def pose_box_stats(self, states):
        """Port library only: per (pose, box) exit stage of the collider WITHOUT the pose-level short-circuits
        (ORC_ST_*: 0 aabb, 1 above, 2 under, 3 span, 4 single plane, 5 vertex, 6 plane hit, 7 fall-through, 255 outside
        the map), hit flag and zone vertex count; arrays of shape [n, 5] (box 0 = torso, 1..4 = feet)."""
        assert self.kind == "port"
        s = _f64(states)
        n = s.shape[0]
        st = np.zeros((n, 5), np.uint8); hit = np.zeros((n, 5), np.uint8); zv = np.zeros((n, 5), np.uint32)
        f = self.lib.orc_port_pose_box_stats
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        assert f(self.h, s.ctypes.data, n, st.ctypes.data, hit.ctypes.data, zv.ctypes.data) == 0
        return st, hit, zv

    def valid_segment_count(self, low, high, s1, s2, frac=0.01):
        a, b = _f64(s1), _f64(s2)
        n = a.shape[0]
        nd = np.zeros(n, np.int32)
        f = self.lib.orc_valid_segment_count
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lo, hi = _f64(low), _f64(high)
        assert f(lo.ctypes.data, hi.ctypes.data, float(frac), a.ctypes.data, b.ctypes.data, n, nd.ctypes.data) == 0
        return nd

    def check_motions_segments(self, s1, s2, nd):
        """DiscreteMotionValidator::checkMotion(s1, s2, lastValid) per edge -> (valid, lastValid.second)."""
        a, b = _f64(s1), _f64(s2)
        n = a.shape[0]
        seg = np.ascontiguousarray(nd, dtype=np.int32)
        valid = np.zeros(n, np.uint8)
        t = np.zeros(n, np.float64)
        f = self.lib.orc_check_motions_segments
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        assert f(self.h, a.ctypes.data, b.ctypes.data, n, seg.ctypes.data, valid.ctypes.data, t.ctypes.data) == 0
        return valid, t

    def edge_matrix(self, s_start, s_target):
        a, b = _f64(s_start), _f64(s_target)
        out = np.zeros((a.shape[0], 6), np.float32)
        f = self.lib.orc_edge_matrix
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        assert f(a.ctypes.data, b.ctypes.data, a.shape[0], out.ctypes.data) == 0
        return out

    def check_edge_interiors(self, s1, s2, n_interp=None, max_lateral=0.5):
        a, b = _f64(s1), _f64(s2)
        n = a.shape[0]
        out = np.zeros(n, np.int32)
        ni = None if n_interp is None else np.ascontiguousarray(n_interp, dtype=np.int32)
        rc = self.lib.orc_check_edge_interiors(self.h, a.ctypes.data, b.ctypes.data, n,
                                               None if ni is None else ni.ctypes.data, float(max_lateral), out.ctypes.data)
        assert rc == 0
        return out

    def path_length_cost(self, s1, s2):
        a, b = _f64(s1), _f64(s2)
        n = a.shape[0]
        cost = np.zeros(n, np.float64)
        rc = self.lib.orc_path_length_cost(self.h, a.ctypes.data, b.ctypes.data, n, cost.ctypes.data)
        assert rc == 0
        return cost
