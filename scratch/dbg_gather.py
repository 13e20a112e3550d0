import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from mag2d_b200 import decks
from mag2d_b200.api import Sim
G=np.load('/root/repo/tests/golden/reference_v1.npz')
d=decks.deck("c2","/tmp/dbg",n_particles=10,geometry="RF_8PT",x_sampl=41,z_sampl=41,Bt=0.01,Bz=0.02,Br=0.005)
sim=Sim(d["config"],d["species_conf"],presolve=False)
k="c2_RF_8PT_"
sim.set_field("u",G[k+"u"]); sim.set_field("uRF",G[k+"uRF"])
x,z=G[k+"E_xz"]
out={}
out['u_back']=sim.get_field('u'); out['urf_back']=sim.get_field('uRF')
for t in (0.0,3.3e-8):
    ex,ez=sim.field_E(x,z,t); out['E_%g'%t]=np.stack([ex,ez])
print('rf',sim.grid.rf,'amp',sim.grid.rf_amplitude,'omega',sim.grid.rf_omega, sim.param['rf'])
np.savez('/root/repo/gpurun_out/dbg_gather.npz',**out)
