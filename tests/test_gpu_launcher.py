"""The shipped launcher (crispresso2_b200/launcher.py) against the UNMODIFIED reference CLI on a real GPU: the reference's own
`CRISPResso` main() (baseline/_ref, pip-installed from /root/reference by __graft_entry__.build(); it travels to the GPU box) is
run twice on the same FASTQ -- as it is (CPU), and through `python -m crispresso2_b200.launcher` (process_fastq, filterFastqs
and the table around the cut re-bound to the engine, sm_100a library) -- and every file of the two output folders must be
byte-identical.  The CPU twin of this test (warp-emulator engine) is tests/test_cli_dropin.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import golden_util as G  # noqa: E402
from test_cli_dropin import _info_stats, _snapshot  # noqa: E402

pytestmark = pytest.mark.gpu

REF_MAIN = ("import sys; sys.path.insert(0, %r); from baseline import ref_shim; CORE = ref_shim.load_core(); "
            "sys.argv = ['CRISPResso'] + sys.argv[1:]; CORE.main()" % ROOT)


def _fastq(tmp_path, case):
    rec = G.load(case)
    fq = tmp_path / (case + ".fastq")
    with open(fq, "w") as fh:
        for k, s in enumerate(rec["reads"]):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    return rec, str(fq)


@pytest.mark.parametrize("case,extra", [("fanc_cas9", ["--write_detailed_allele_table"]),
                                        ("synth_hdr", []),
                                        ("synth_single", ["--ignore_substitutions", "-w", "10"])])
def test_launcher_output_folder_equals_the_reference(case, extra, tmp_path):
    from baseline import ref_shim
    if not ref_shim.available():
        pytest.skip("baseline/_ref (the pip-installed reference) did not travel")
    rec, fq = _fastq(tmp_path, case)
    names = rec["ref_names"]
    amp = rec["refs"][names[0]]["sequence"]
    guide = "GGAATCCCTTCTGCAGCACC" if case.startswith("fanc") else amp[110:130]
    argv = ["-r1", fq, "-a", amp, "-g", guide, "--suppress_plots", "--suppress_report"] + extra
    if len(names) > 1:
        argv += ["-e", rec["refs"][names[1]]["sequence"]]
    env = dict(os.environ, PYTHONPATH=ROOT)
    outs = {}
    for mode, cmd in (("ref", [sys.executable, "-c", REF_MAIN]), ("b200", [sys.executable, "-m", "crispresso2_b200.launcher"])):
        out = str(tmp_path / mode)
        os.makedirs(out)
        p = subprocess.run(cmd + argv + ["-o", out], capture_output=True, text=True, timeout=900, env=env, cwd=out)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
        outs[mode] = out
    a, b = _snapshot(outs["ref"]), _snapshot(outs["b200"])
    assert sorted(a) == sorted(b)
    diff = [k for k in a if a[k] != b[k]]
    assert not diff, diff
    assert len(a) >= 10
    assert _info_stats(outs["ref"]) == _info_stats(outs["b200"])
