"""The drop-in end to end on a GPU: the reference's `parafrost` CLI (its parser, CDCL, local search, model
extension and verification) linked against libsigma_b200 through integration/sigma_shim.cpp
(`oracle/_ref/parafrost_sigma`, built by integration/Makefile in the build container).  The final answer must
equal the unmodified reference GPU binary's, and every SAT model - extended over the engine's witness stack by
the reference's own MODEL::extend - must pass the reference's own `-modelverify` against the input file."""
import os
import re
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "parafrost_sigma")
REF = os.path.join(ROOT, "oracle", "_ref", "parafrost_gpu")
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def run(binary, cnf, *flags, timeout=40):
    r = subprocess.run([binary, cnf] + list(flags), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    out = ANSI.sub("", r.stdout)
    ans = [l for l in out.splitlines() if l.startswith("s ")]
    return (ans[-1][2:].strip() if ans else None), out


def cnf_file(tmp_path, fam, seed, args):
    path = str(tmp_path / f"{fam}_{seed}.cnf")
    helpers.gen_cnf(fam, seed, args, dimacs_path=path)
    return path


def test_dropin_cli_solves_through_the_engine(tmp_path):
    """the run recorded in profiles/r01_dropin_cli_v54.log"""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/parafrost_sigma not built")
    ans, out = run(BIN, cnf_file(tmp_path, "ksat", 12, [300, 900, 3]))
    assert ans == "SATISFIABLE", out[-2000:]
    m = re.search(r"Removed variables\s*:\s*(\d+)", out)
    assert m and int(m.group(1)) > 0, out[-2000:]     # the simplifier did run


CASES = {
    "k3_sat": ("ksat", 12, [300, 900, 3]),
    "k3_dense": ("ksat", 15, [150, 690, 3]),
    "parity": ("parity", 41, [300]),
    "mult6": ("mult", 31, [6]),
    "multpar": ("multpar", 51, [5, 120]),
}


@pytest.mark.parametrize("name", list(CASES))
def test_dropin_answer_and_model_match_reference(tmp_path, name):
    if not (os.path.exists(BIN) and os.path.exists(REF)):
        pytest.skip("oracle/_ref binaries not built")
    cnf = cnf_file(tmp_path, *CASES[name])
    ans, out = run(BIN, cnf, "-model", "-modelverify")
    # --ereminthreads=32: the unmodified reference's ERE launch writes past its shared memory on small instances
    # (tests/golden/make_golden.py, ERE_LAUNCH_FIX); the option makes its launch shape consistent
    ref_ans, ref_out = run(REF, cnf, "--ereminthreads=32", "-model", "-modelverify")
    both = "---- parafrost_sigma ----\n" + out[-2500:] + "\n---- parafrost_gpu (reference) ----\n" + ref_out[-2500:]
    assert ans is not None, both
    assert ref_ans is not None, both
    assert ans == ref_ans, both
    if ans == "SATISFIABLE":
        assert "model VERIFIED" in out, both
