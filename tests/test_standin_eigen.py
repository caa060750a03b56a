"""The stand-in Eigen that the reference sources are compiled against (oracle/refbuild/standin) has the semantics the
hot path relies on: comma initialiser with blocks, column-major storage, left-to-right products, Triplet default,
setFromTriplets summing duplicates in triplet order into sorted columns with explicit zeros kept, InnerIterator column
walks, a direct solve and a Jacobi-CG.  Values are worked out by hand in oracle/refbuild/standin_selftest.cpp."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_standin_eigen_selftest():
    exe = os.path.join(tempfile.mkdtemp(), "standin_selftest")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "oracle", "refbuild", "standin"),
                           os.path.join(ROOT, "oracle", "refbuild", "standin_selftest.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "standin selftest: ok" in out.stdout, out.stdout + out.stderr
