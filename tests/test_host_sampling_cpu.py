"""CPU test of host logic that ships in the product: the library's initial sampling (acvd_b200/csrc/host_sampling.hpp, the
flat-ring version with prefetching) equals the plain restatement of ComputeInitialRandomSampling in the same header on the
same rings -- compiled here with g++, no GPU involved."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flat_ring_sampling_equals_plain_restatement(tmp_path):
    exe = tmp_path / "host_sampling_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "acvd_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "host_sampling_check.cpp"), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
