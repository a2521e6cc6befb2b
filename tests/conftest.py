import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_c0_config(**extra):
    """BASELINE config[0] as the plugin would hand it to the simulator (fixture produced by
    tests/golden/make_golden.py from the reference's shipped fax_config + test overrides)."""
    from wfsim_b200 import config as wcfg
    with open(os.path.join(GOLDEN, 'c0_config.json')) as f:
        fax = json.load(f)
    return wcfg.plugin_config(fax, overrides=extra or None, to_pe=np.full(494, 0.008))


def load_det_case(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    extra = json.loads(str(z['cfg_extra']))
    cfg = load_c0_config(**extra)
    cfg['gains'] = z['gains'].copy()
    noise = None
    if 'noise_x2' in z.files:
        noise = z['noise_x2'].astype(np.float64) / 2.0
        cfg['enable_noise'] = True
    from wfsim_b200.dtypes import raw_record_dtype
    rr = z['rr'].view(raw_record_dtype())
    rr_he = z['rr_he'].view(raw_record_dtype())
    return dict(cfg=cfg, pcall=z['pcall'], channel=z['channel'], t=z['t'], gain=z['gain'],
                group_of=z['group_of'], noise=noise, ix_rand=z['ix_rand'], rr=rr, rr_he=rr_he)


@pytest.fixture(scope='session')
def c0_config():
    return load_c0_config()


DET_CASES = ['det_basic', 'det_he_thr', 'det_noise']
