"""State files of the reference (Ax1SimulationState ASCII format) and doExtrapolateXData: the loader is
checked against the state file the reference ships (copied verbatim as a data fixture) and against
the independently parsed JSON fixture; the writer round-trips bit for bit. CPU only."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DAT = os.path.join(ROOT, "tests", "golden", "yasp_square1mm_p0.dat")
G1 = json.load(open(os.path.join(ROOT, "tests", "golden", "yasp_square1mm.json")))


def test_load_shipped_state(axlib):
    st = axlib.load_state(DAT)
    assert st["elements"] == 90 and st["time"] == 1000.0 and st["dt"] == 10.0 and st["timestep"] == 100
    assert st["mpi_rank"] == 0 and st["mpi_size"] == 1 and st["nvectors"] == 5
    for k in ("x", "y", "channelStates", "uold", "unew"):
        assert np.array_equal(st[k], np.array(G1[k])), k
    assert len(st["unew"]) == 2 * 91 * 4


def test_save_round_trip(axlib, tmp_path):
    st = axlib.load_state(DAT)
    p = str(tmp_path / "chk_p0.dat")
    vecs = {k: st[k] for k in ("x", "y", "channelStates", "uold", "unew")}
    axlib.save_state(p, st["elements"], st["time"], st["dt"], vecs, timestep=st["timestep"])
    back = axlib.load_state(p)
    for k in vecs:
        assert np.array_equal(back[k], st[k])       # 17 significant digits round-trip exactly
    assert back["time"] == st["time"] and back["timestep"] == 100
    # same header keys, in the reference's order
    head = [ln.split()[0] for ln in open(p).read().splitlines()[:7]]
    assert head == ["#elements", "timestep", "time", "dt", "mpi_rank", "mpi_size", "#vectors"]


def test_extrapolate_x(axlib):
    st = axlib.load_state(DAT)
    x = np.linspace(0.0, st["x"][-1], 7)
    u = axlib.extrapolate_state(st["x"], st["y"], st["unew"], x, st["y"]).reshape(91, 7, 4)
    src = st["unew"].reshape(91, 2, 4)
    assert np.array_equal(u[:, 0, :], src[:, 0, :]) and np.array_equal(u[:, -1, :], src[:, 1, :])
    # interior columns lie between the two stored columns (Q1 in x), y nodes are reproduced exactly
    lo, hi = np.minimum(src[:, 0, :], src[:, 1, :]), np.maximum(src[:, 0, :], src[:, 1, :])
    tol = 1e-15 * np.maximum(np.abs(lo), np.abs(hi))
    assert np.all(u[:, 3, :] >= lo - tol) and np.all(u[:, 3, :] <= hi + tol)
    # refinement in y: linear between nodes
    y2 = np.sort(np.concatenate([st["y"], 0.5 * (st["y"][:-1] + st["y"][1:])]))
    u2 = axlib.extrapolate_state(st["x"], st["y"], st["unew"], st["x"], y2).reshape(len(y2), 2, 4)
    assert np.array_equal(u2[0::2], src)
    assert np.allclose(u2[1::2], 0.5 * (src[:-1] + src[1:]), rtol=1e-13, atol=1e-13)
