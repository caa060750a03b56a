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
