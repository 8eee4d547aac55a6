"""Builds and runs tests/cpp/test_api.cpp: the reference's tests restated against the C++ mirror (include/jubjub_schnorr.hpp)."""
import json
import os
import subprocess

import pytest

from tests.test_oracle_kat import legacy_double_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arr(name, b):
    return "static const uint8_t %s[%d] = {%s};\n" % (name, len(b), ", ".join(str(x) for x in b))


def test_cpp_api(tmp_path):
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))["multisig_kat"]
    pkd, sgd, _ = legacy_double_fixture()
    hdr = "#include <stdint.h>\nstatic const uint8_t KAT_PUBLIC_KEYS[3][32] = {%s};\n" % ", ".join(
        "{" + ", ".join(str(x) for x in bytes.fromhex(h)) + "}" for h in k["PUBLIC_KEYS"])
    hdr += _arr("KAT_AGGREGATE_PUBLIC_KEY", bytes.fromhex(k["AGGREGATE_PUBLIC_KEY"])) + _arr("KAT_SIGNATURE", bytes.fromhex(k["SIGNATURE"]))
    hdr += _arr("KAT_LEGACY_DOUBLE_PK", pkd) + _arr("KAT_LEGACY_DOUBLE_SIG", sgd)
    (tmp_path / "kat.h").write_text(hdr)
    exe = str(tmp_path / "test_api")
    libdir = os.path.join(ROOT, "jubjub_schnorr_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", str(tmp_path), "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_api.cpp"),
                           "-L", libdir, "-ljjschnorr_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "cpp api ok" in out.stdout
