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


@pytest.mark.parametrize("val", ["tma", "packed"])
def test_expansion_kernel_variant(val):
    """The opt-in forms of the polynomial-expansion kernel (RC_POLYEXP=tma: cp.async.bulk.tensor staging; =packed: FFMA2 pair
    accumulators) pass the same golden / oracle / batched checks as the default."""
    env = dict(os.environ, RC_POLYEXP=val)
    out = subprocess.run([sys.executable, os.path.join(HERE, "variant_check.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "variant ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("var,val", [("RC_PYR", "separate"), ("RC_FLOW_KERNEL", "strip"), ("RC_POLYEXP", "tma")])
def test_alternative_kernels_give_identical_bits(tmp_path, var, val):
    """pyr3_kernel (layers 0-2 in one pass over the frame) against the per-layer pyramid kernels, and the strip form of
    the fused flow layer (forced on every launch) against the tile form these small launches select by default: same
    arithmetic, so the flows must agree bit for bit -- results do not depend on the batch size that picks the kernel.
    The TMA form of the expansion kernel differs from the default only in how its input tile reaches shared memory."""
    import numpy as np
    outs = []
    for mode in ("default", "alt"):
        env = dict(os.environ)
        if mode == "alt":
            env[var] = val
            if var == "RC_FLOW_KERNEL":
                env["RC_STRIP_MINPX"] = "0"
        path = str(tmp_path / (mode + ".npz"))
        out = subprocess.run([sys.executable, os.path.join(HERE, "pyr_dump.py"), path], env=env, capture_output=True, text=True,
                             timeout=600)
        assert out.returncode == 0 and "dumped" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
        outs.append(np.load(path))
    assert sorted(outs[0].files) == sorted(outs[1].files) and len(outs[0].files) == 5
    for k in outs[0].files:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_march_tma_staging_gives_identical_bits(tmp_path):
    """The marching kernel (winsize > 3) stages M with one TMA tensor copy per interior step; RC_MARCH_TMA=0 selects the
    per-element cp.async copies everywhere.  Same data in the same shared-memory layout, so every flow must agree bit for bit
    (six configurations: half-widths 2 / 5 / 10, box and Gaussian, ragged sizes, layers narrower than the TMA box)."""
    import numpy as np
    outs = []
    for mode in ("default", "alt"):
        env = dict(os.environ)
        if mode == "alt":
            env["RC_MARCH_TMA"] = "0"
        path = str(tmp_path / (mode + ".npz"))
        out = subprocess.run([sys.executable, os.path.join(HERE, "march_dump.py"), path], env=env, capture_output=True, text=True,
                             timeout=600)
        assert out.returncode == 0 and "dumped" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
        outs.append(np.load(path))
    assert sorted(outs[0].files) == sorted(outs[1].files) and len(outs[0].files) == 6
    for k in outs[0].files:
        assert np.isfinite(outs[0][k]).all() and np.abs(outs[0][k]).max() > 0.1, k
        assert np.array_equal(outs[0][k], outs[1][k]), k
