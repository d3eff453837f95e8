"""Config 5 (401x401 px x 10^6 rays) with the mirror-pair walk on and off: az >= 0 bit-identical, az < 0 to 1e-9 (dev probe, GPU)."""
import os, sys, subprocess, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) > 1:
    from solat_optics_b200 import maps
    ax = np.linspace(-5, 5, 101)
    yy, xx = np.meshgrid(ax, ax, indexing="ij")
    e5 = np.exp(-(xx ** 2 + yy ** 2) / (2 * 1.5 ** 2)) * np.exp(1j * 0.3 * (xx ** 2 + yy ** 2))
    t0 = time.time()
    px, a, p = maps.ff_map(140, 0.05, int(os.environ.get("NUM", "200")), 1.32, xx, yy, e5, n_scan=1000, device=0)
    print(sys.argv[1], "%.2f s" % (time.time() - t0), flush=True)
    np.savez(sys.argv[1], px=px, a=a, p=p)
else:
    subprocess.check_call([sys.executable, __file__, "/tmp/c5_on.npz"])
    subprocess.check_call([sys.executable, __file__, "/tmp/c5_off.npz"], env=dict(os.environ, SOLAT_FF_PAIRS="0"))
    on, off = np.load("/tmp/c5_on.npz"), np.load("/tmp/c5_off.npz")
    pos = on["px"][:, 0] >= 0
    wrap = lambda v: (v + np.pi) % (2 * np.pi) - np.pi
    print("pixels", len(pos), "az>=0 bit-identical:", np.array_equal(on["a"][pos], off["a"][pos]) and np.array_equal(on["p"][pos], off["p"][pos]))
    print("az>=0: max |dA|/A %.2e  max |dphi| %.2e (the two modes use different CTA shapes for a 101x101 grid, i.e. another summation order)"
          % (np.max(np.abs(on["a"][pos] - off["a"][pos]) / off["a"][pos]), np.max(np.abs(wrap(on["p"][pos] - off["p"][pos])))))
    print("az<0: max |dA|/A %.2e  max |dphi| %.2e" % (np.max(np.abs(on["a"][~pos] - off["a"][~pos]) / off["a"][~pos]),
                                                     np.max(np.abs(wrap(on["p"][~pos] - off["p"][~pos])))))
