"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a
GPU, exports every symbol include/sigma.h declares, and its host-only entry points (option
defaults / normalisation, argument validation) behave like the reference's option handling
(src/gpu/options.cpp:168-300).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import parafrost_b200
    from parafrost_b200 import sigma
    parafrost_b200.build()
    return sigma.lib()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sigma.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sigma_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    import parafrost_b200
    from parafrost_b200 import sigma
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libsigma_b200.so does not export {s}"
    assert sorted(sigma.SYMBOLS) == syms
    # nothing torch-typed or C++-mangled in the public surface
    out = subprocess.run(["nm", "-D", "--defined-only", parafrost_b200.lib_path()], stdout=subprocess.PIPE, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(syms) <= exported


def test_library_is_sm100a_only():
    import parafrost_b200
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", parafrost_b200.lib_path()], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_touch_the_oracle():
    """The product path must never import / link / call anything under oracle/."""
    pkg = os.path.join(ROOT, "parafrost_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "sigma_oracle" not in txt and "oracle/" not in txt and "oracle_" not in txt, f
    import parafrost_b200
    out = subprocess.run(["ldd", parafrost_b200.lib_path()], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out


def test_default_and_normalised_options(lib):
    from parafrost_b200 import sigma
    o = sigma.make_opts()
    # defaults of src/gpu/options.cpp:24-43 and options.cu:36-60
    assert (o.phases, o.ve_en, o.ve_plus_en, o.sub_en, o.bce_en, o.ere_en) == (5, 1, 1, 1, 0, 1)
    assert (o.mu_pos, o.mu_neg, o.lcve_min_vars, o.lcve_max_occurs, o.lcve_clause_max) == (32, 32, 2, 3000, 30000)
    assert (o.phase_lits_min, o.shrink_rate, o.ve_clause_max, o.xor_max_arity) == (500, 2, 100, 10)
    assert (o.ere_clause_max, o.sh_max_bve_out1) == (250, 250)
    # derivations of options.cpp:291-296
    o = sigma.make_opts(ve_en=0, ve_plus_en=1)
    assert o.ve_en == 1
    o = sigma.make_opts(ve_en=0, ve_plus_en=0, sub_en=0, bce_en=0)
    assert o.phases == 0
    o = sigma.make_opts(all_en=1)
    assert o.bce_en == 1 and o.ere_en == 1 and o.ve_en == 1
    o = sigma.make_opts(ve_en=0, ve_plus_en=0, phases=4)
    assert o.phases == 1
    o = sigma.make_opts(ere_clause_max=100000)
    assert o.ere_clause_max == 250
    # the reference's CLI spelling
    d = sigma.opts_from_flags(["--phases=3", "-no-ere", "-bce", "--mupos=16", "-no-lcvefast"])
    assert d == {"phases": 3, "ere_en": 0, "bce_en": 1, "mu_pos": 16, "lcve_fast": 0}
    assert sigma.opts_from_flags(["-lcvefast"]) == {"lcve_fast": 1} and sigma.make_opts().lcve_fast == 0
    with pytest.raises(KeyError):
        sigma.opts_from_flags(["-nonsense"])


def test_null_arguments_are_errors_not_crashes(lib):
    assert lib.sigma_create(0, None, None) != 0
    assert lib.sigma_destroy(None) == 0
    assert lib.sigma_run(None, None) != 0
    assert lib.sigma_num_rounds(None) == 0
    assert lib.sigma_last_error(None) == b"null context"
    assert b"sm_100a" in lib.sigma_version()


def test_no_gpu_means_loud_failure(lib):
    """Without a device the product raises; it never falls back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from parafrost_b200 import sigma
    with pytest.raises(sigma.SigmaError):
        sigma.Simplifier(0)


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """include/sigma.h is the FFI surface: it must be valid C99 by itself (no C++, no CUDA, no torch types) and a C
    program must link against the library with nothing else on the link line."""
    import parafrost_b200
    hdr = os.path.join(ROOT, "include", "sigma.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "sigma.h"\n#include <stdio.h>\n'
        "static void sink(void* u, const uint8_t* b, uint64_t n) { (void)u; (void)b; (void)n; }\n"
        "int main(void) {\n"
        "  sigma_opts o; sigma_ctx* c = 0; int rc;\n"
        "  sigma_default_opts(&o); o.proof_en = 1; sigma_normalize_opts(&o);\n"
        "  rc = sigma_create(0, &o, &c);\n"
        '  printf("%d %d %d\\n", rc, (int)sizeof o, o.phases);\n'
        "  if (!rc) { sigma_set_proof_sink(c, sink, 0); sigma_destroy(c); }\n"
        "  return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(parafrost_b200.lib_path())
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                        "-L", libdir, "-lsigma_b200", "-Wl,-rpath," + libdir], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout.split()
    from parafrost_b200 import sigma
    assert int(out[1]) == C.sizeof(sigma.SigmaOpts)      # the ctypes mirror has the layout the C compiler sees
    assert int(out[2]) == 5


def test_cpp_example_builds_parses_and_fails_loudly_without_a_gpu(lib, tmp_path):
    """examples/sigma_cli.cpp: the C ABI from a plain g++ program (no nvcc, no reference headers)."""
    import parafrost_b200
    libdir = os.path.dirname(parafrost_b200.lib_path())
    exe = tmp_path / "sigma_cli"
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "sigma_cli.cpp"), "-L", libdir, "-lsigma_b200", "-Wl,-rpath," + libdir, "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    cnf = tmp_path / "t.cnf"
    cnf.write_text("c tiny\np cnf 3 3\n1 -2 0\n2 3 0\n-1 -3 0\n")
    r = subprocess.run([str(exe), str(cnf)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "3 variables, 3 clauses, 6 literals" in r.stdout
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CPU path" in r.stdout      # loud failure, no fallback
    else:
        assert r.returncode == 0 and "\ns " in r.stdout
