"""The C header is valid C, and the reference-side shim compiles against the mocked PFEM3D interfaces."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                           os.path.join(ROOT, "include", "pfem_b200.h")])


def test_shim_compiles_against_mock_reference():
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "shim", "mock"), os.path.join(ROOT, "shim", "compile_check.cpp")])


def test_shim_compiles_against_the_real_reference_headers():
    """Where /root/reference exists: the shim against the reference's own Equation/Solver/Problem/Mesh headers (third-party
    includes answered by oracle/refbuild/standin)."""
    import pytest
    if not os.path.isdir("/root/reference/srcs"):
        pytest.skip("no /root/reference here")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-fopenmp", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "shim"), "-I", os.path.join(ROOT, "oracle", "refbuild", "standin"),
                           "-I", os.path.join(ROOT, "oracle", "refbuild"), "-I", "/root/reference/srcs",
                           os.path.join(ROOT, "shim", "compile_check_reference.cpp")])
