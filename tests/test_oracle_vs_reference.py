"""Direct cross-check of the CPU oracle against the UNMODIFIED reference, run live.  Only possible in the
build container (the reference tree does not travel to the GPU box): skipped when /root/reference is absent.
The committed fixtures (tests/golden) pin the same thing without the reference; this test shows the fixtures
are not stale and covers a case that is not in them."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SOLAT_REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "solat_optics")), reason="reference tree not present")

_SNIPPET = r"""
import json, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r)
import ref_loader as rl
ref = rl.load()
amp, phi = rl.nf_pixel(ref, 220.0, -18.0, 33.0, 14, p_rx=(20.0, 0.0, -35.0))
tg = rl.nf_tele_geo(ref, 220.0, 14)
so = ref.rti.source_to_lyot_model([-18.0, tg.y_ap, 33.0], tg, rl.PlotOpts).output()
print(json.dumps({"amp": float(amp), "phi": float(phi), "opl": [float(v) for v in so[3]]}))
"""


def test_oracle_matches_live_reference(oracle):
    out = subprocess.run([sys.executable, "-c", _SNIPPET % os.path.join(ROOT, "oracle")], capture_output=True, text=True,
                         cwd="/tmp", timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    ref = json.loads(out.stdout.strip().splitlines()[-1])
    import golden_util as gu

    y_ap, fw = gu.nf_band(220.0)
    so = oracle.nf_source_trace([-18.0, y_ap, 33.0], 14)
    np.testing.assert_allclose(so[3], np.array(ref["opl"]), rtol=0, atol=1e-10)
    rx = oracle.nf_rx_trace([20.0, 0.0, -35.0], fw, 14)
    amp, ph = oracle.nf_map([-18.0, 33.0], y_ap, 14, rx[3], rx[4] * gu.lyot_mask(14), [gu.nf_k(220.0)])
    assert abs(amp[0, 0] ** 2 - ref["amp"] ** 2) <= 1e-9 * ref["amp"] ** 2
    assert abs(gu.wrap(ph[0, 0] - ref["phi"])) <= 1e-9
