import sys, numpy as np, tempfile
sys.path.insert(0,'/root/repo')
from mag2d_b200 import decks
from mag2d_b200.api import Sim
d=decks.deck("c4",tempfile.mkdtemp(),n_particles=20000000)
sim=Sim(d["config"],d["species_conf"])
sim.run_initscript(d["initscript"])
for s in (1,2): sim.sort(s)
sim.advance_init()
sim.set_solver(cycles_per_step=0,tol=1e-14,max_cycles=60)
sim.advance(3)
# exact solution for the current rho
sim.species_accumulate  # noqa
info=sim.solve(rf=False,tol=1e-15,max_cycles=80); uex=sim.get_field('u'); print('exact',info, np.abs(uex).max())
# perturb: restart from a slightly stale field (previous step) and converge to various tolerances
for tol in (1e-6,1e-8,1e-9,1e-10,1e-11,1e-12):
    sim.set_field('u', uex*(1+1e-3*np.sin(np.arange(uex.size).reshape(uex.shape)*0.01)))
    info=sim.solve(rf=False,tol=tol,max_cycles=80); u=sim.get_field('u')
    print('tol',tol,info,'rel err',np.abs(u-uex).max()/np.abs(uex).max())
# warm-start behaviour: fixed cycles per step, error vs converged each step
for cyc in (1,2,3,4):
    sim.set_solver(cycles_per_step=cyc,tol=1e-14,max_cycles=60)
    sim.advance(6)
    u=sim.get_field('u')                 # field used by the last push = solve of rho from previous step
    rho_fixed=[sim.rho_fixed(s) for s in (1,2)]
    # emulate: what would the converged solve of the *previous* rho be? not available; instead measure residual now
    info=sim.solve(rf=False,tol=1e-15,max_cycles=80)
    print('cycles/step',cyc,'-> extra cycles to 1e-15',info['cycles'])
