import sys, os, numpy as np
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root); sys.path.insert(0, root + "/tests")
from helpers import small_setup
from dune_ax1_b200 import setups
outs = []
for kw in (dict(nx=40, dx=100.0), dict(nx=5, stimulation=True), dict(nx=17, dx=100.0e3)):
    for flags in ((0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)):
        s = small_setup(**kw)
        s["params"] = dict(s["params"], use_elec_jacobian_volume=flags[0], use_memb_jacobian_volume=flags[1], use_time_jacobian_volume=flags[2])
        dev = setups.make_device(s)
        u0 = s["u0"] * (1.0 + 1e-3 * np.random.default_rng(3).standard_normal(s["u0"].shape))
        dev.set_time(0.0, 1.0); dev.set_uold(s["u0"]); dev.update_state(s["u0"])
        outs.append(np.asarray(dev.jacobian(u0)).ravel())
        dev.close()
np.save(sys.argv[1], np.concatenate(outs))
