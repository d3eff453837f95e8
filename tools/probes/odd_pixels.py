import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, golden_util as gu
from solat_optics_b200 import maps
tgf = maps.ff_tele_geo(140, 1.32)
th, phi = gu.ff_fan(50)
x, y, e = gu.ff_grid("nf51")
plan = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0)
for odd in ([[3e-4, gu.EL0], [-3e-4 - 1e-15, gu.EL0]], [[-5e-4, gu.EL0]], [[2e2, gu.EL0]], [[np.nan, gu.EL0]], [[3e-4, gu.EL0], [-3e-4 - 1e-15, gu.EL0], [-5e-4, gu.EL0], [2e2, gu.EL0], [np.nan, gu.EL0]]):
    try:
        print(odd[-1], plan.run(np.array(odd)))
    except Exception as ex:
        print(odd[-1], 'ERR', ex)
