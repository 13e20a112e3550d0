import sys, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np
from mag2d_b200 import decks
from mag2d_b200.api import Sim
import bench
tmp = tempfile.mkdtemp()
n = 125_000_000
d = bench.make_deck("c5", n, 1, tmp)
sim = Sim(d["config"], d["species_conf"])
e = sim.species_index("ELECTRON")
sim.generate(e, "everywhere", n)
sim.sort(e)
sim.set_sort_interval(-1)
sim.advance_init()
for k in range(14):
    sim.advance(2)
    sim.sync()
    print(k, sim.store_stats(e), sim.count(e), flush=True)
