"""The reference-side binding (staarpipeline_b200/rshim/) must compile with the SAME 13 signatures the reference's
generated glue declares (src/RcppExports.cpp).  R/Rcpp are absent from the image, so it is compiled against the
stand-in headers (oracle/ref_stub) and linked against the C-ABI library; the signatures are compared textually with
the declarations in RcppExports.cpp when /root/reference is present."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "staarpipeline_b200", "rshim", "staar_b200_shim.cpp")


def test_shim_compiles_and_links(tmp_path):
    from staarpipeline_b200 import build
    lib = build.build()
    out = tmp_path / "libshim.so"
    cmd = ["g++", "-std=c++17", "-fPIC", "-shared", "-DARMA_64BIT_WORD=1", "-I", os.path.join(ROOT, "oracle", "ref_stub"),
           "-I", os.path.join(ROOT, "include"), SHIM, "-o", str(out), lib, "-L" + os.path.join(ROOT, "oracle"),
           "-l:liboracle.so", "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-DC", str(out)], capture_output=True, text=True).stdout
    for fn in ("matrix_flip_mean", "matrix_flip_minor", "Individual_Score_Test_SPA", "Individual_Score_Test_cond_multi"):
        assert fn + "(" in syms


def _norm(sig: str) -> str:
    sig = re.sub(r"\s*=\s*\d+", "", sig)              # default arguments live only in the definition
    return re.sub(r"\s+", " ", sig).strip()


@pytest.mark.skipif(not os.path.exists("/root/reference/src/RcppExports.cpp"), reason="reference tree not present")
def test_signatures_match_rcppexports():
    glue = open("/root/reference/src/RcppExports.cpp").read()
    decl = {m.group(2): _norm(m.group(1) + " " + m.group(2) + "(" + m.group(3) + ")")
            for m in re.finditer(r"^(List|arma::vec) (\w+)\(([^;{]*)\);", glue, flags=re.M)}
    assert len(decl) == 13
    shim = open(SHIM).read()
    mine = {m.group(2): _norm(m.group(1) + " " + m.group(2) + "(" + m.group(3) + ")")
            for m in re.finditer(r"^(List|arma::vec) (\w+)\(([^{]*)\)\s*\{", shim, flags=re.M | re.S)}
    for name, sig in decl.items():
        assert name in mine, name
        assert mine[name] == sig, (mine[name], sig)
