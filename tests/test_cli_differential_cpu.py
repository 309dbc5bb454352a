"""Differential test of the command line: the B200 driver (gvamp_b200/bin/main_real, host/options.cpp) against the
UNMODIFIED reference executable (oracle/_ref/main_real_scalar.exe = main_real.cpp + options.cpp of the reference, built by
oracle/Makefile with the single-rank MPI shim).  Every option of options.cpp:18-429 is fed valid, out-of-range and malformed
values; the two programs must print the same text (option echo, FATAL messages) and return the same exit code.  No run
mode that touches the GPU is used (`--run-mode nothing` stops after the option echo in both programs), so this runs on CPU.

Two lines of the reference's echo are typos and are deliberately not reproduced:
  options.cpp:342  `--probit-var` is echoed without the blank before its value;
  options.cpp:356  `--EM-max-iter` echoes the value of --EM-err-thr instead of its own.
"""
import os
import subprocess

import pytest

from conftest import ROOT

OURS = os.path.join(ROOT, "gvamp_b200", "bin", "main_real")
REF = os.path.join(ROOT, "oracle", "_ref", "main_real_scalar.exe")

NUMERIC = ["--C", "--CG-max-iter", "--CV", "--EM-err-thr", "--EM-max-iter", "--Mt", "--Mt-test", "--N", "--N-test", "--alpha-scale", "--gam1-init",
           "--gamma-damp", "--gamw-init", "--h2", "--init-est", "--iterations", "--learn-vars", "--num-mix-comp", "--probit-var", "--rho", "--seed",
           "--stop-criteria-thr", "--store-pvals", "--use-XXT-denoiser", "--use-freeze", "--use-lmmse-damp", "--verbosity", "--red", "--meth-imp",
           "--predict"]
VALUES = ["0", "-1", "3", "0.5", "2.5", "abc"]

OTHER = [
    "--bed-file x.bed --model linear --run-mode nothing",
    "--bed-file x.bed --model bin_class --run-mode nothing",
    "--bed-file x.bed --model foo --run-mode nothing",
    "--bed-file x.bed --run-mode foo",
    "--bed-file x.bed --test-iter-range 3,7 --run-mode nothing",
    "--bed-file x.bed --test-iter-range 3 --run-mode nothing",
    "--bed-file x.bed --test-iter-range a,b --run-mode nothing",
    "--bed-file x.bed --probs 0.5,0.3,0.2 --vars 0,0.01,0.1 --run-mode nothing",
    "--bed-file x.bed --probs 0.5,,0.2 --run-mode nothing",
    "--bed-file x.bed --probs abc --run-mode nothing",
    "--bed-file x.bed --phen-files a.phen,b.phen --run-mode nothing",
    "--bed-file x.bed --true-signal-files t1,t2 --run-mode nothing",
    "--bed-file x.bed --bed-file-test y.bed --phen-files-test p.phen --run-mode nothing",
    "--bed-file x.bed --bim-file x.bim --ref-bim-file r.bim --run-mode nothing",
    "--bed-file x.bed --cov-file c.cov --cov-estimate-file ce --dim-file d --run-mode nothing",
    "--bed-file x.bed --estimate-file e.bin --freeze-index-file f --group-index-file g --group-mixture-file gm --run-mode nothing",
    "--bed-file x.bed --out-name nm --run-mode nothing",
    "--bed-file x.bed --iterations",
    "--bed-file x.bed --model",
    "--bed-file",
    "--bed-file x.bed --bed-file y.bed --run-mode nothing",
    "--run-mode nothing",
    "--run-mode infere",
    "--help",
    "-h",
    "--no-such-flag 1",
    "--bed-file x.bed extra --run-mode nothing",
]

# the reference's echo typos (see the module docstring): what it prints -> what a correct echo prints
def _reference_echo_typos_fixed(text, args):
    lines = text.splitlines()
    out = []
    for ln in lines:
        if ln.startswith("--probit-var") and not ln.startswith("--probit-var "):
            ln = "--probit-var " + ln[len("--probit-var"):]
        elif ln.startswith("--EM-max-iter ") and "--EM-max-iter" in args:
            ln = "--EM-max-iter " + str(int(float(args[args.index("--EM-max-iter") + 1])))
        out.append(ln)
    return out


@pytest.fixture(scope="module", autouse=True)
def built():
    from gvamp_b200 import build
    build.build_all(verbose=False)
    if not os.path.exists(REF):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        else:
            pytest.skip("oracle/_ref (the compiled reference) is not present on this machine")
    assert os.path.exists(OURS) and os.path.exists(REF)


def _both(args, cwd):
    res = []
    for exe in (REF, OURS):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=60, cwd=cwd)
        res.append((r.returncode, r.stdout))
    return res


def test_numeric_options_behave_like_the_reference(tmp_path):
    out_dir = str(tmp_path / "od") + "/"
    n = 0
    for opt in NUMERIC:
        for val in VALUES:
            args = ["--bed-file", "x.bed", opt, val, "--run-mode", "nothing", "--out-dir", out_dir]
            (rc_ref, txt_ref), (rc_ours, txt_ours) = _both(args, str(tmp_path))
            assert rc_ref == rc_ours, (opt, val, rc_ref, rc_ours)
            assert _reference_echo_typos_fixed(txt_ref, args) == txt_ours.splitlines(), (opt, val)
            n += 1
    assert n == len(NUMERIC) * len(VALUES)


@pytest.mark.parametrize("line", OTHER)
def test_string_and_list_options_behave_like_the_reference(tmp_path, line):
    args = line.split()
    if "--run-mode" in args and "nothing" in args:
        args += ["--out-dir", str(tmp_path / "od") + "/"]
    (rc_ref, txt_ref), (rc_ours, txt_ours) = _both(args, str(tmp_path))
    assert rc_ref == rc_ours, (line, rc_ref, rc_ours)
    assert _reference_echo_typos_fixed(txt_ref, args) == txt_ours.splitlines(), line
