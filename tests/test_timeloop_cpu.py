"""The time-step controller of Acme2CylSimulation::run as restated in dune_ax1_b200/timeloop.py (and, line for line,
csrc/timeloop.cu) replayed on the oracle's full action-potential trace (tests/golden/ap_trace_paper_grid.json): fed
the Newton iteration counts and membrane-potential maxima of the trace it must reproduce the trace's dt sequence --
x 1.1 / : 1.2 from the iteration counts (no first-step exception), clamps, the action-potential limiter, the reset at
the start of the injection; what remains are the halvings after a failed Newton solve (the trace was recorded with
Newton tolerances at the rounding floor, where that happens in one step out of twelve). CPU only."""
import json
import os

import numpy as np

from helpers import trace_restarts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_controller_reproduces_the_dt_sequence_of_the_oracle_trace():
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ap_trace_paper_grid.json")))
    rows = fx["steps"]
    restarts = trace_restarts(fx)          # asserts the sequence step by step
    assert len(restarts) == len(rows) > 600
    assert sum(1 for n in restarts if n) <= 0.1 * len(rows) and max(restarts) <= 3
    dts = np.array([r[1] for r in rows])
    assert rows[0][1] == 1.0 * 1.1                     # first adaptive step: 1.1 dt0
    assert dts.max() <= 10.0 + 1e-12 and dts.min() >= 0.05   # AP limiter while stimulating / V_m > -50 mV
    assert np.any(dts == 1.0)                          # reset to dtstart when the injection starts
