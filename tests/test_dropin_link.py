"""The drop-in boundary at link level: the reference's host objects + integration/sigma_shim.cpp +
libsigma_b200.so link into a `parafrost` CLI without any of the reference's CUDA translation units.
Needs the reference tree (build container only); on a box without it the prebuilt binary is checked."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "parafrost_sigma")
HAVE_REF = os.path.isdir("/root/reference/src/gpu") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "obj_gpu"))


def sh(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)


@pytest.fixture(scope="module")
def dropin():
    if HAVE_REF:
        from parafrost_b200.build import build
        build()
        r = sh(["make", "-s", "-f", os.path.join(ROOT, "integration", "Makefile")])
        assert r.returncode == 0, r.stdout[-3000:]
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/parafrost_sigma not built (needs /root/reference)")
    return BIN


def test_shim_defines_the_seven_symbols_and_imports_only_the_c_abi(dropin):
    defined = sh(["nm", "-C", "--defined-only", dropin]).stdout
    for sym in ("ParaFROST::Solver::simplify(bool const&)", "ParaFROST::Solver::optSimp()", "ParaFROST::Solver::freeSimp()",
                "ParaFROST::Solver::newBeginning()", "ParaFROST::cuMM::cuMM()", "ParaFROST::CACHER::destroy()",
                "ParaFROST::GOPTION::GOPTION()"):
        assert re.search(r" [TW] " + re.escape(sym) + r"$", defined, re.M), sym
    # none of the reference's simplifier kernels or their launchers are in the binary
    for gone in ("ve_k_1", "ParaFROST::sub_k", "ParaFROST::ere_k", "ParaFROST::bce_k", "create_ot_k", "prep_cnf_k", "Solver::awaken",
                 "Solver::LCVE", "Solver::sortOT"):
        assert gone not in defined, gone
    undefined = sh(["nm", "-C", "--undefined-only", dropin]).stdout
    abi = sorted(set(re.findall(r"\bU (sigma_[a-z_]+)", undefined)))
    assert abi, "the binary does not import the C ABI"
    hdr = open(os.path.join(ROOT, "include", "sigma.h")).read()
    for s in abi:
        assert re.search(r"\b" + s + r"\s*\(", hdr), f"{s} is not declared in include/sigma.h"
    # the reference's own host mirror in and out (pinned), the loop one round at a time, the per-prop trail ranges
    assert {"sigma_create", "sigma_load_sclauses", "sigma_begin", "sigma_round", "sigma_finish", "sigma_store_compact", "sigma_store_sclauses",
            "sigma_trail_info", "sigma_copy_trail", "sigma_pinned_alloc", "sigma_destroy"} <= set(abi)


def test_cli_flags_of_the_simplifier_still_parse(dropin):
    out = sh([dropin, "-h"]).stdout
    for flag in ("--resolventmax", "-vefunction", "--xormaxarity", "--eremaxoccurs", "-profilegpu", "--phases", "-proof"):
        assert flag in out, flag
