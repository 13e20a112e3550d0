import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from mag2d_b200 import decks
from mag2d_b200.api import Sim
d=decks.deck("c2","/tmp/dbg",n_particles=10,geometry="RF_8PT",x_sampl=41,z_sampl=41)
sim=Sim(d["config"],d["species_conf"],presolve=False)
M=N=41
ii,jj=np.indices((M,N)).astype(float)
x=np.array([0.0,0.25,0.5,0.75,1.0,1.5,5.3,19.487,39.9,40.0])*5e-4
z=np.array([3.2,0.0,0.5,7.75,1.0,1.5,5.3,28.024,39.9,40.0])*5e-4
for name,u,urf in (('i',ii,ii*0),('j',jj,ii*0),('i*i',ii*ii,ii*0),('rf=i',ii*0,ii),('100i+j',100*ii+jj,ii*0)):
    sim.set_field("u",u); sim.set_field("uRF",urf)
    ex,ez=sim.field_E(x,z,0.0)
    print(name,'Ex/idx',np.round(ex/2000,4),'Ez/idz',np.round(ez/2000,4))
