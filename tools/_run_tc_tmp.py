import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minesweeper_b200 import synth
from minesweeper_b200.engine import Engine
prec = sys.argv[1]
net = synth.make_phot_net(h=128, seed=1)
mist = synth.make_mist(ne=40, nm=12, nf=5, na=3, seed=0)
eng = Engine(mist, net, None, precision=prec)
rng = np.random.default_rng(0)
n = 1 << 20
pp = np.stack([rng.uniform(3000, 14000, n), rng.uniform(-0.5, 5.3, n), rng.uniform(-3.9, 0.4, n),
               rng.uniform(-0.2, 0.6, n), rng.uniform(-1.5, 2.5, n), rng.uniform(1.0, 5000.0, n), rng.uniform(0.0, 3.0, n)], 1)
m = eng.phot_sed(pp)
torch.cuda.synchronize()
pt = torch.from_numpy(pp).cuda()
for _ in range(3): eng.phot_sed(pt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.phot_sed(pt)
e1.record(); torch.cuda.synchronize()
print('phot_sed ms', prec, e0.elapsed_time(e1) / 10)
