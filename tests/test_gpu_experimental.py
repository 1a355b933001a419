"""The switchable kernel paths that are OFF by default stay under test: each case re-runs a few parity tests of
``tests/test_gpu_kernels.py`` in a child process with the switch set (the library reads its switches once per process).

* ``B200SEG_LINE_CONV=1 B200SEG_LINE_W128=1`` -- the line-tiled tcgen05 kernel (``csrc/tc_line.cu``) takes the 16-channel
  3x3x3 layers whose rows are 32 / 64 / 128 voxels (fprop, dgrad, fused statistics, fused InstanceNorm-backward sums);
* ``B200SEG_SLIDE_PERSIST=1`` -- persistent CTAs in the sliding conv kernel (several items per CTA);
* ``B200SEG_SLIDE_MINB3=0`` / ``1`` -- the 16->16 sliding kernel forced to its two- / three-CTAs-per-SM build;
* ``B200SEG_NORM_CLUSTER=0`` -- InstanceNorm backward without thread-block clusters (the r1 / r2 kernels).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# test ids of tests/test_gpu_kernels.py: TC_GEOMS sp13..sp18 = sliding-kernel shapes (sp18 = 2 x 40 x 128 x 128: long
# sweeps, several items per persistent CTA), sp19..sp22 = line-kernel shapes (rows of 32 / 64 / 128 voxels)
SLIDE = "(tcgen05_conv and (sp13 or sp14 or sp15 or sp16 or sp17 or sp18))"
LINE = "(tcgen05_conv and (sp18 or sp19 or sp20 or sp21 or sp22))"
CASES = [
    ({"B200SEG_LINE_CONV": "1", "B200SEG_LINE_W128": "1"}, LINE + " or partials or fused_with"),
    ({"B200SEG_LINE_CONV": "1", "B200SEG_LINE_W128": "1", "B200SEG_LINE_EG": "2"}, LINE + " or fused_with"),
    ({"B200SEG_SLIDE_PERSIST": "1"}, SLIDE + " or partials or fused_with"),
    ({"B200SEG_SLIDE_MINB3": "0"}, SLIDE + " or fused_with"),
    ({"B200SEG_SLIDE_MINB3": "1"}, SLIDE + " or fused_with"),
    ({"B200SEG_NORM_CLUSTER": "0"}, "test_instnorm_prelu"),
]


@pytest.mark.parametrize("env,kexpr", CASES, ids=["+".join(f"{k[8:]}={v}" for k, v in e.items()) for e, _ in CASES])
def test_switched_paths(env, kexpr):
    child_env = dict(os.environ)
    child_env.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"), "-m", "gpu",
                        "-x", "-q", "-k", kexpr, "-p", "no:cacheprovider"],
                       cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout or "")[-3000:] + (r.stderr or "")[-1000:]
    assert r.returncode == 0, f"switched path {env} failed:\n{tail}"
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, tail
