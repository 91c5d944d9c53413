"""The drop-in claim, taken literally: a program written against the REFERENCE's header-only C++ wrapper
(include/charls/charls.hpp, SURVEY.md 8b "what calls it") is compiled with the reference's headers and linked with
charls_b200/lib/libcharls.so.3.  The compile step needs /root/reference (this container); the binary lands in
tests/abi_consumer/build/ and travels to the GPU box, where the round trip through charls::jpegls_encoder /
jpegls_decoder runs on the kernels.  (Last in file order on purpose: `pytest -x` reaches it after everything else.)"""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.support import REF_LIB, REFERENCE_ROOT, ROOT, s_smooth

SOURCE = os.path.join(ROOT, "tests", "abi_consumer", "consumer.cpp")
BINARY = os.path.join(ROOT, "tests", "abi_consumer", "build", "consumer")
INCLUDE = os.path.join(REFERENCE_ROOT, "include")


def build_consumer(library_dir, library_name, output):
    os.makedirs(os.path.dirname(output), exist_ok=True)
    rpath = os.path.relpath(library_dir, os.path.dirname(output))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + INCLUDE, SOURCE, "-o", output, "-L" + library_dir,
                           "-l:" + library_name, "-Wl,-rpath,$ORIGIN/" + rpath])
    return output


@pytest.fixture(scope="module")
def consumer(product):
    if os.path.isdir(INCLUDE) and shutil.which("g++"):
        return build_consumer(os.path.dirname(product.path), os.path.basename(product.path), BINARY)
    if os.path.exists(BINARY):
        return BINARY
    pytest.skip("reference headers not here and no prebuilt consumer")


def test_reference_cpp_wrapper_links_and_parses_headers(consumer, oracle, tmp_path):
    """Host-only calls through the reference's C++ classes: same answers as the same program on the reference library."""
    image = s_smooth(16, 24, 8, 3, layout="interleaved")
    path = tmp_path / "frame.jls"
    path.write_bytes(oracle.encode_image(image, 8, ilv=2, ri=1))
    ours = subprocess.run([consumer, "header", str(path)], capture_output=True, text=True)
    assert ours.returncode == 0, ours.stderr
    assert ours.stdout.split() == ["24", "16", "8", "3", "0", "2", str(16 * 24 * 3)]
    if os.path.isdir(INCLUDE) and os.path.exists(REF_LIB):
        theirs_binary = build_consumer(os.path.dirname(REF_LIB), os.path.basename(REF_LIB), str(tmp_path / "consumer_ref"))
        theirs = subprocess.run([theirs_binary, "header", str(path)], capture_output=True, text=True)
        assert (theirs.returncode, theirs.stdout) == (0, ours.stdout)


def test_reference_cpp_wrapper_reports_the_missing_gpu(consumer):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is here: see the round trip below")
    run = subprocess.run([consumer, "roundtrip"], capture_output=True, text=True)
    assert run.returncode == 2 and "jpegls_error 200" in run.stderr, (run.returncode, run.stderr)  # loud, no CPU fallback


@pytest.mark.gpu
def test_reference_cpp_wrapper_round_trip(consumer):
    run = subprocess.run([consumer, "roundtrip"], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert int(run.stdout.split()[0]) > 0
