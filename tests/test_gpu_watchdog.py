"""The engine's wait watchdog (csrc/c2b_engine.cu): a blocking wait for the GPU that lasts longer than C2B_WATCHDOG_S ends the
process with exit code 70 and a message instead of hanging it.  Forced here with a 10 ms limit on 4 Mi-read launches (each
wait lasts ~55 ms); the same command with the default limit finishes normally (every other GPU test)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_a_wait_beyond_the_limit_ends_the_process_loudly():
    env = dict(os.environ, C2B_WATCHDOG_S="0.01")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--reads", str(1 << 22), "--steps", "4", "--warmup", "3",
                        "--no-api", "--no-cpu-baseline", "--no-gate"], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 70, (p.returncode, p.stderr[-500:])
    assert "a wait for the GPU has lasted more than" in p.stderr
    assert p.stdout.strip() == ""                           # no bench line from a run that was cut short
