"""CPU-only: the sampler oracle (oracle/artp_oracle.c: orc_sample_states, restating sampler.cpp:40-131) and the
test-side Philox restatement the GPU stream is compared with."""
import numpy as np
import pytest

import cases
import philox_ref
from art_planner_b200 import synth


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = philox_ref.philox4x32_10(np.array([ctr], np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_uniform_stream_is_uniform_and_counter_based():
    u = philox_ref.sampler_uniforms(1234, 0, 20000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    assert np.array_equal(philox_ref.sampler_uniforms(1234, 5000, 10), u[5000:5010])     # random access
    assert not np.array_equal(philox_ref.sampler_uniforms(1235, 0, 10), u[:10])


@pytest.fixture(scope="module")
def setup(maps):
    m = maps("fbm_rough")
    return m, synth.make_sampler_layers(m, seed=7)


def test_oracle_cdf_scan_semantics(setup, port_lib):
    """sampler.cpp:66-71: first index whose CDF value exceeds the variate, last index as fallback; NaN rows fall
    through to the last column."""
    m, L = setup
    sp = synth.sampler_params_for(m)
    u = philox_ref.sampler_uniforms(3, 0, 20000)
    st, rc = port_lib.sample_states(m, L, sp, 0.2, u)
    row_want = np.minimum(np.searchsorted(L.cum_prob_rowwise[:-1].astype(np.float64), u[:, 1], side="right"), m.rows - 1)
    assert np.array_equal(rc[:, 0], row_want)
    for i in range(0, 2000):
        c = L.cum_prob[rc[i, 0], :-1].astype(np.float64)
        col = np.searchsorted(c, u[i, 0], side="right") if not np.isnan(c[0]) else m.cols - 1
        assert rc[i, 1] == min(col, m.cols - 1)
    # mass-less cells are never drawn (apart from the last-index fallbacks)
    inner = (rc[:, 0] < m.rows - 1) & (rc[:, 1] < m.cols - 1)
    assert (L.sample_probability[rc[inner, 0], rc[inner, 1]] > 0).all()
    # a variate beyond the row CDF's last-but-one entry selects the last row, whose CDF is NaN -> last column
    u2 = u[:4].copy()
    u2[:, 1] = np.nextafter(1.0, 0.0)
    _, rc2 = port_lib.sample_states(m, L, sp, 0.2, u2)
    assert (rc2[:, 0] == m.rows - 1).all() and (rc2[:, 1] == m.cols - 1).all()


def test_oracle_state_construction(setup, port_lib):
    m, L = setup
    sp = synth.sampler_params_for(m)
    u = philox_ref.sampler_uniforms(4, 0, 5000)
    st, rc = port_lib.sample_states(m, L, sp, 0.2, u)
    x, y = m.cell_xy()
    r, c = rc[:, 0], rc[:, 1]
    # position = cell centre + normal * pert, |pert| <= min(std, 0.5) * reach_z   (sampler.cpp:93-107)
    pert_max = np.minimum(L.plane_fit_std_dev[r, c], 0.5).astype(np.float64) * 0.2
    d = st[:, :3] - np.stack([x[r], y[c], m.elevation[r, c].astype(np.float64)], axis=1)
    n = np.stack([L.normal_x[r, c], L.normal_y[r, c], L.normal_z[r, c]], axis=1).astype(np.float64)
    assert (np.linalg.norm(d, axis=1) <= pert_max * np.linalg.norm(n, axis=1) + 1e-12).all()
    assert np.allclose(np.cross(d, n), 0, atol=1e-12)
    assert np.allclose(np.linalg.norm(st[:, 3:], axis=1), 1.0, atol=1e-14)
    # with zero perturbation the body z axis tilts with the terrain normal: roll/pitch from the normal (:123-126)
    sp0 = synth.SamplerParams(0.0, 0.0, True, sp.low, sp.high)
    st0, _ = port_lib.sample_states(m, L, sp0, 0.2, u)
    qx, qy, qz, qw = st0[:, 3], st0[:, 4], st0[:, 5], st0[:, 6]
    yaw = np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))
    assert np.allclose(np.abs(np.angle(np.exp(1j * (yaw - np.pi * (1 - 2 * u[:, 5]))))), 0, atol=1e-9)


def test_oracle_uniform_mode_rejects_outside(setup, port_lib):
    m, L = setup
    sp = synth.sampler_params_for(m, from_distribution=False)     # bounds = map +- one length (planner.cpp:148-156)
    u = philox_ref.sampler_uniforms(5, 0, 20000)
    st, rc = port_lib.sample_states(m, L, sp, 0.2, u)
    rej = rc[:, 0] < 0
    assert 0.6 < rej.mean() < 0.9                                  # ~ 1 - 1/4 .. edges
    assert np.isnan(st[rej]).all() and not np.isnan(st[~rej]).any()
    lx, ly = m.length
    xs = sp.low[0] + (sp.high[0] - sp.low[0]) * u[:, 0]
    inside = np.abs(xs - m.cx) < 0.5 * lx - 1e-9
    assert not (rej & inside & (np.abs(sp.low[1] + (sp.high[1] - sp.low[1]) * u[:, 1] - m.cy) < 0.5 * ly - 1e-9)).any()


def test_oracle_estimate_normals(maps, port_lib):
    """utils.cpp:213-324 restated: unit normals tilted against the slope, zero vector where no neighbour pair fits,
    plane_fit_std_dev = largest |dz| over the visited neighbours."""
    m = maps("ramp")
    nx, ny, nz, sd = port_lib.estimate_normals(m, 0.49)
    n2 = nx.astype(np.float64) ** 2 + ny.astype(np.float64) ** 2 + nz.astype(np.float64) ** 2
    zero = n2 == 0
    assert np.allclose(n2[~zero], 1.0, atol=1e-6)
    # the only cells without any neighbour pair are the two corners (rows-1, 0) and (0, cols-1)... and their like:
    # every loop needs either (i+o, j+o) or (i-o, j-o) style room
    assert zero.sum() <= 4 and zero[m.rows - 1, 0] and zero[0, m.cols - 1]
    assert (nz[~zero] > 0).all()
    inner = (slice(16, -16), slice(16, -16))
    e = m.elevation.astype(np.float64)
    r = int(0.49 / m.res)
    # std layer == max |dz| along the axis / diagonal offsets actually visited (interior cells see all of them)
    want = np.zeros_like(e[inner])
    for o in range(1, r):
        for di, dj in ((o, 0), (0, o), (-o, 0), (0, -o)):
            want = np.maximum(want, np.abs(np.roll(e, (-di, -dj), (0, 1))[inner] - e[inner]))
    rd = int(0.49 * 0.70710678118 / m.res)
    for o in range(1, rd):
        for di, dj in ((o, o), (-o, o), (-o, -o), (o, -o)):
            want = np.maximum(want, np.abs(np.roll(e, (-di, -dj), (0, 1))[inner] - e[inner]))
    assert np.allclose(sd[inner], want, atol=1e-6)
    # flat map: every defined normal is exactly +z
    f = maps("flat")
    fx, fy, fz, fs = port_lib.estimate_normals(f, 0.49)
    ok = fz != 0
    assert (fz[ok] == 1.0).all() and (fx[ok] == 0).all() and (fy[ok] == 0).all() and (fs == 0).all()


def test_oracle_cdf_layers(setup, port_lib):
    """probability_distribution.cpp:20-46 restated: row-normalised column cumulation, row distribution, NaN rows."""
    m, L = setup
    cum, row = port_lib.compute_cdf(L.sample_probability)
    p = L.sample_probability.astype(np.float64)
    mass = p.sum(axis=1)
    dead = mass == 0
    assert dead.any() and np.isnan(cum[dead]).all() and not np.isnan(cum[~dead]).any()
    want = np.cumsum(p[~dead] / mass[~dead, None], axis=1)
    assert np.abs(cum[~dead] - want).max() < 5e-6
    assert np.abs(row - np.cumsum(mass / mass.sum())).max() < 5e-6
    assert (np.diff(cum[~dead], axis=1) >= 0).all() and (np.diff(row) >= 0).all()
    # numpy stand-in used by synth.make_sampler_layers agrees to summation-order noise
    assert np.nanmax(np.abs(cum - L.cum_prob)) < 5e-6
