"""CTA-pair mode of conv_halo_kernel (tcgen05 cta_group::2: two CTAs of a cluster compute one M256 x N tile pair, each
holding its own activation tile and half of the weight rows).  It is not the default routing (csrc/conv_halo.cu,
profiles/r02_pair_mode.txt), so the conv parity tests and the launch stress test are re-run here with the policy forced
(`I2R_HALO_PAIR=2`: every problem whose shape allows it, resident and streamed weights, fp16 and split-operand) in a
child process -- the policy is read once per process."""
import os
import subprocess
import sys

import pytest

import paths

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("target", ["tests/test_kernels_gpu.py", "tests/test_halo_stress_gpu.py"])
def test_parity_and_stress_suites_pass_in_forced_pair_mode(target):
    env = dict(os.environ, I2R_HALO_PAIR="2")
    proc = subprocess.run([sys.executable, "-m", "pytest", target, "-m", "gpu", "-q", "-x"], cwd=paths.REPO, env=env,
                          capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-1500:]
    assert proc.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
