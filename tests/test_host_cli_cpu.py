"""CPU-only checks of the C++ host drivers: command-line behaviour identical to the reference's
Options class (FATAL + exit code 1 on bad input), and loud failure -- not emulation -- when the CUDA
device is missing."""
import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "gvamp_b200", "bin", "main_real")


@pytest.fixture(scope="module", autouse=True)
def built():
    from gvamp_b200 import build
    build.build_all(verbose=False)
    assert os.path.exists(EXE)


def run(*args):
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=60)


def test_unknown_flag_is_fatal():
    r = run("--no-such-flag", "1")
    assert r.returncode == 1 and 'FATAL: option "--no-such-flag" unknown' in r.stdout


def test_missing_value_is_fatal():
    r = run("--bed-file")
    assert r.returncode == 1 and 'missing argument for last option "--bed-file"' in r.stdout


def test_range_checks_match_reference_messages():
    r = run("--bed-file", "x.bed", "--iterations", "0")
    assert r.returncode == 1 and "option --iterations has to be a strictly positive integer! (0 was passed)" in r.stdout
    r = run("--bed-file", "x.bed", "--N-test", "0")
    assert r.returncode == 1 and "option --N_test has to be" in r.stdout
    r = run("--bed-file", "x.bed", "--seed", "-3")
    assert r.returncode == 1 and "option --seed has to be a non-negative integer" in r.stdout


def test_no_bed_file_is_fatal():
    r = run("--run-mode", "infere")
    assert r.returncode == 1 and "no bed file provided" in r.stdout


def test_missing_phen_file_is_fatal(tmp_path):
    r = run("--bed-file", "x.bed", "--phen-files", str(tmp_path / "nope.phen"))
    assert r.returncode == 1 and "not found" in r.stdout


def test_options_echo_and_out_dir(tmp_path):
    out = tmp_path / "newdir"
    r = run("--bed-file", "x.bed", "--out-dir", str(out), "--rho", "0.25", "--probs", "0.9,0.1", "--run-mode", "nothing")
    assert r.returncode == 0
    assert "ardyh command line options:" in r.stdout and "--rho 0.25" in r.stdout and "--probs 0.9,0.1" in r.stdout
    assert out.is_dir()


def test_infere_without_gpu_fails_loudly(tmp_path):
    """No CPU fallback: with no CUDA device the driver must stop with a FATAL line."""
    from gvamp_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    phen = tmp_path / "a.phen"
    phen.write_text("".join(f"{i} {i} {0.1 * i}\n" for i in range(8)))
    bed = tmp_path / "a.bed"
    bed.write_bytes(bytes([0x6C, 0x1B, 0x01]) + bytes(2 * 4))
    r = run("--run-mode", "infere", "--bed-file", str(bed), "--phen-files", str(phen), "--N", "8", "--Mt", "4", "--probs", "0.5,0.5",
            "--vars", "0,0.1", "--out-dir", str(tmp_path) + "/", "--out-name", "t")
    assert r.returncode != 0 and "FATAL" in r.stdout and "CUDA" in r.stdout
