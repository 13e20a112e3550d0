import sys, numpy as np, tempfile
sys.path.insert(0,'/root/repo')
from mag2d_b200 import decks
from mag2d_b200.api import Sim
d=decks.deck("c4",tempfile.mkdtemp(),n_particles=100000000)
sim=Sim(d["config"],d["species_conf"])
sim.run_initscript(d["initscript"])
for s in (1,2): sim.sort(s)
sim.set_sort_interval(8)
sim.advance_init()
for cyc in (1,2,3,4,-1,-2,-3,-4):
    sim.set_solver(cycles_per_step=cyc,tol=1e-12,max_cycles=60)
    sim.advance(4); sim.solver_stats()
    sim.advance(12)
    print('cycles/step',cyc,'max resid over 12 steps',sim.solver_stats()['resid'])
