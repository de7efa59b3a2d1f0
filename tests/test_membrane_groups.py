"""Host model of the membrane-group grid generator and permittivity map (BASELINE config 5, the myelinated
axon): Ax1GridGenerator::setupXPartition / intervalVectorToGridVector (ax1_gridgenerator.hh:1021-1312),
Physics::initPermittivities (acme2_cyl_physics.hh:2424-2606), processor-boundary guard (:903-942).
The reference ships no golden x vector for a multi-group grid (all shipped states are 1-column equilibria),
so these are property tests of the restated algorithm on the shipped myelin configuration
(src/config/acme2_cyl_par_fiona_myelin_48mmx10mm_48nodes_...config)."""
import numpy as np
import pytest

from dune_ax1_b200 import api, setups


@pytest.fixture(scope="module")
def axlib_built():
    import __graft_entry__ as ge
    ge.build()


def test_partition_48_nodes(axlib_built):
    mg = setups.myelin_groups()
    node, myel = mg.names.index("node_of_ranvier"), mg.names.index("myelin")
    assert len(mg.ival_group) == 97 and np.sum(mg.ival_group == node) == 48
    assert mg.ival_start[0] == 0.0 and mg.ival_end[-1] == 48.0e6
    assert np.array_equal(mg.ival_start[1:], mg.ival_end[:-1])            # complete, non-overlapping
    assert np.array_equal(mg.ival_group[::2], np.full(49, myel))           # default group fills the gaps
    ns = mg.ival_start[mg.ival_group == node]
    assert np.allclose(ns, 500.0e3 + 1000.0e3 * np.arange(48)) and np.allclose(mg.ival_end[mg.ival_group == node] - ns, 1.0e3)


def test_x_vector_properties(axlib_built):
    mg = setups.myelin_groups()
    x = mg.x_vector()
    dx = np.diff(x)
    assert x[0] == 0.0 and x[-1] == 48.0e6 and np.all(dx > 0)
    # every group border is a grid point
    for b in mg.ival_end:
        assert np.min(np.abs(x - b)) < 1e-6
    xc = 0.5 * (x[:-1] + x[1:])
    k = 7  # a node in the middle
    ns, ne = 500.0e3 + 1.0e6 * k, 501.0e3 + 1.0e6 * k
    assert np.allclose(dx[(xc > ns) & (xc < ne)], 100.0) and np.sum((xc > ns) & (xc < ne)) == 10
    # n_transition = 10 cells of dx_transition = 100 on the myelin side of both borders
    assert np.allclose(dx[(xc > ns - 1000.0) & (xc < ns)], 100.0) and np.allclose(dx[(xc > ne) & (xc < ne + 1000.0)], 100.0)
    # geometric ramp 100 -> 10e3 (q = 0.7): monotone, ratio of neighbours <= 1/0.7 (+ rounding slack of the fit)
    right = dx[(xc > ne + 1000.0) & (xc < ne + 60.0e3)]
    assert np.all(np.diff(right) >= -1e-6) and np.max(right[1:] / right[:-1]) < 1.0 / 0.7 + 0.05
    assert abs(right[-1] - 10.0e3) < 1e-6 and dx.max() < 10.0e3 * (1 + 1e-9)
    # mirror symmetry of one myelin interval
    seg = dx[(xc > ne) & (xc < ns + 1.0e6)]
    assert np.allclose(seg, seg[::-1], rtol=1e-9)


def test_membrane_arrays(axlib_built):
    mg = setups.myelin_groups()
    x = mg.x_vector()
    eps, gl, gn, gk, grp = mg.membrane_arrays(x)
    node = mg.names.index("node_of_ranvier")
    xc = 0.5 * (x[:-1] + x[1:])
    assert np.all(eps[grp == node] == 2.0) and np.all(gn[grp == node] == 1200.0) and np.all(gk[grp == node] == 360.0)
    assert np.all(gl[grp != node] == 0.0) and np.all(gn[grp != node] == 0.0)
    far = (grp != node) & (np.abs(((xc - 500.5e3) % 1.0e6 + 0.5e6) % 1.0e6 - 0.5e6) > 2.0e3)
    assert np.allclose(eps[far], 6.0 * 5.0 / 500.0)
    # smooth step x^2 (3 - 2x) over n_transition * dx_transition = 1000 nm on the myelin side of each border
    ne = 501.0e3
    m = (xc > ne) & (xc < ne + 1000.0)
    xs = (xc[m] - ne) / 1000.0
    assert np.allclose(eps[m], 2.0 + (0.06 - 2.0) * xs * xs * (3 - 2 * xs), rtol=1e-12)
    ns = 1500.0e3
    m = (xc > ns - 1000.0) & (xc < ns)
    xs = (xc[m] - (ns - 1000.0)) / 1000.0
    assert np.allclose(eps[m], 0.06 + (2.0 - 0.06) * xs * xs * (3 - 2 * xs), rtol=1e-12)
    # a rank-local slice gives the same arrays as the global run (the state machine starts mid-axon)
    i0, i1 = 2000, 2600
    e2, *_ = mg.membrane_arrays(x[i0:i1 + 1])
    assert np.array_equal(e2, eps[i0:i1])


def test_processor_boundary_guard(axlib_built):
    mg = setups.myelin_groups()
    x = mg.x_vector()
    nx = len(x) - 1
    for n in (2, 4, 8):                      # the reference's own run uses 8 ranks (docker-compose-myelin.yml:6)
        for r in range(n - 1):
            mg.check_processor_boundary(x[api.slab(nx, n, r)[1]])
    with pytest.raises(api.Ax1GpuError):
        mg.check_processor_boundary(500.5e3)     # inside a node of Ranvier
    with pytest.raises(api.Ax1GpuError):
        mg.check_processor_boundary(501.3e3)     # inside the transition zone next to it


def test_uniform_single_group_is_equidistant(axlib_built):
    groups = {"axon": dict(leak=0.5, Nav=120.0, Kv=36.0)}
    mg = api.MembraneGroups(groups, "axon", 10.0e6, 100.0e3)
    x = mg.x_vector()
    assert len(x) == 101 and np.allclose(np.diff(x), 100.0e3) and x[-1] == 10.0e6
    eps, gl, gn, gk, grp = mg.membrane_arrays(x)
    assert np.all(eps == 2.0) and np.all(gn == 120.0)


def test_myelin_setup_shapes(axlib_built):
    s = setups.make_myelin_setup(xmax=2.0e6)
    assert s["u0"].shape == (4 * (s["nx"] + 1) * 181,)
    u = s["u0"].reshape(181, s["nx"] + 1, 4)
    node = s["groups"].names.index("node_of_ranvier")
    i = int(np.nonzero(s["cell_group"] == node)[0][0])
    assert abs(u[30, i, 0] - 11.9033551) < 1e-6 and abs(u[30, 0, 0] - 12.0025723) < 1e-6   # node / myelin columns
    s2 = setups.make_myelin_setup(xmax=2.0e6, rank=1, nranks=2)
    assert s2["left_pb"] and not s2["right_pb"] and s2["x"][-1] == 2.0e6
