import sys, time, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from oracle import pyoracle as O
from dune_ax1_b200 import setups
from dune_ax1_b200.timeloop import TimeLoop
from helpers import oracle_problem
nx=int(sys.argv[1]); tend=float(sys.argv[2])
s=setups.make_setup(nx, dx=100e3, stimulation=True, perturbation=0.0)
P=oracle_problem(O,s)
opts=O.default_newton_opts(reduction=1e-5, abs_limit=1e-5, min_lin_reduction=1e-5, ilu_fill=1, maxit=50)
tl=TimeLoop(P, s['u0'], dtstart=1.0, newton_opts=opts, do_stimulation=True, lower_limit=10, upper_limit=30)
t0=time.time()
while tl.time < tend:
    r=tl.step()
    tr=tl.trace[-1]
    if len(tl.trace)%10==0 or tr[3]>0: print("t=%8.2f dt=%6.3f its=%d rst=%d vm[min,max]=[%.3f, %.3f] lin=%d  wall %.0fs"%(tr[0],tr[1],tr[2],tr[3],tr[4],tr[5],r.lin_iterations,time.time()-t0), flush=True)
