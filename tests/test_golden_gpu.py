"""The committed golden fixtures of the scaled BASELINE shapes (tests/golden/c3_*, c4_*, c5_*.npz, produced in the build
container by tests/golden/make_golden.py from the reference's L2 sources over the oracle) checked on the GPU by a process
that loads neither the oracle nor the reference: tests/golden/check_gpu.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cuda_path_reproduces_scaled_baseline_goldens_without_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "check_gpu.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "golden ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_fixtures_exist_and_are_small():
    g = os.path.join(ROOT, "tests", "golden")
    for f in ("c3_n2_m256_v27.npz", "c4_link_m512_qam256.npz", "c5_n2_m512_qam64.npz"):
        assert 1000 < os.path.getsize(os.path.join(g, f)) < 600000
