"""Every formulation of the fused layer kernel must pass the same parity checks; the choice is
made once per process from the environment, so each variant runs in a subprocess with RC_FLOW_KERNEL forcing it on
every launch, however small."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("kernel", ["strip", "tile"])
def test_fused_layer_variant(kernel):
    env = dict(os.environ, RC_FLOW_KERNEL=kernel, RC_STRIP_MINPX="0")
    out = subprocess.run([sys.executable, os.path.join(HERE, "variant_check.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "variant ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
