"""-m gpu: every kernel variant that is selectable only through a profiling knob stays bit-exact.

The knobs ($ACGPU_FORCE_RAGGED, $ACGPU_FLAT420, $ACGPU_TMA, $ACGPU_WAVES; DESIGN.md section 4, profiles/r1_experiments.md)
are read once per process, so each setting runs tests/knob_worker.py in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("env,tier", [
    ({"ACGPU_FORCE_RAGGED": "0"}, 0),      # row-pair kernels for RGB -> 4:2:0 too (the default walks it flat)
    ({"ACGPU_FORCE_RAGGED": "1"}, 0),      # flat one-row-per-unit kernels for every 4:2:0 pair, aligned widths included
    ({"ACGPU_FLAT420": "0"}, 0),           # no flat walk for 4:2:0 -> RGB at widths that idle lanes
    ({"ACGPU_WAVES": "3"}, 0),             # a different grid shape
    ({"ACGPU_TMA": "0"}, 3),               # tier 3: bulk (TMA) stores
    ({"ACGPU_TMA": "1"}, 3),               # tier 3: bulk-async staged loads
    ({"ACGPU_TMA": "2"}, 3),               # tier 3: staged loads + bulk stores
    ({"ACGPU_TMA": "3"}, 3),               # tier 3: 2-D tensor-map stores (UTMASTG)
    ({"ACGPU_TMA": "4"}, 3),               # tier 3: tensor-map loads (UTMALDG) and stores
    ({"ACGPU_TMA": "5"}, 3),               # tier 3: tensor-map loads, LDS + STG stores
    ({"ACGPU_TMA": "6"}, 3),               # tier 3: three-stage tensor-map loads, tensor-map stores
    ({"ACGPU_TMA": "7"}, 3),               # tier 3: three-stage tensor-map loads, LDS + STG stores
    ({"ACGPU_TMA": "8"}, 3),               # tier 3: four-stage tensor-map loads, LDS + STG stores
    ({"ACGPU_TMA_FLAT": "0"}, 0),          # no flat tensor-map form: widths that do not fill whole warps stay on tier 2 / the row-pair form
    ({"ACGPU_TMA_FLAT": "2"}, 0),          # the flat tensor-map form for every width it can take (1920 and 4128 too)
    ({"ACGPU_TMA_FLAT": "2", "ACGPU_TMA_FLAT_BLOCK": "256"}, 0),
    ({"ACGPU_TMA_AUTO": "0"}, 0),          # the automatic path without the tensor-map form (tier 2 everywhere)
    ({"ACGPU_TMA_AUTO": "7"}, 0),          # tensor-map staged loads for EVERY YUV source -> RGB24 / BGR24 (two rows per trip)
], ids=lambda v: "-".join(f"{k[6:]}{x}" for k, x in v.items()) if isinstance(v, dict) else f"tier{v}")
def test_knob_variants_match_the_checker(env, tier):
    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = os.pathsep.join([ROOT, HERE, e.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, os.path.join(HERE, "knob_worker.py"), str(tier)], capture_output=True, text=True,
                       env=e, cwd=ROOT, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), (env, r.stdout[-400:], r.stderr[-400:])
