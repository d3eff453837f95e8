import sys; sys.path.insert(0, "/root/repo")
import bench
from solat_optics_b200 import maps
x, y, e = bench.e_rec_config5()
px, amp, ph = maps.ff_map(bench.FREQ, bench.RES_ARCMIN, bench.NUM, bench.OFFSET, x, y, e, n_scan=1000, device=0)
print("map done", amp.shape, float(amp.sum()))
