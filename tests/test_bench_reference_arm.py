"""CPU: `bench.py --impl reference` drives the UNMODIFIED reference modules staged under baseline/_ref (no GPU, none of this
package imported) and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "lib", "diffusion", "sampling.py")),
                    reason="baseline/_ref not staged (python baseline/install_reference.py where /root/reference exists)")
def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, MDB_CPU_THREADS=str(min(8, os.cpu_count() or 1)), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "sample-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the arm must not have pulled this repository's package (and with it the native library) into the process
    probe = subprocess.run([sys.executable, "-c",
                            "import sys; sys.argv=['bench.py','--impl','reference','--steps','0']; sys.path.insert(0, %r); "
                            "from baseline import reference_arm; reference_arm.load('cpu'); "
                            "print('PKG', any(m.startswith('meshdiffusion_b200') for m in sys.modules))" % ROOT],
                           capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert "PKG False" in probe.stdout, probe.stdout + probe.stderr[-1000:]
