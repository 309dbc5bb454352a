"""Checker script (not a pytest module): config 1 against the reference golden run with the Onsager warm start on and off;
prints the margins quoted in DESIGN.md section 7.  Uses the oracle only to regenerate the synthetic bed / phen files."""
import os, sys, subprocess, tempfile, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
from oracle import oracle as O
g = np.load(os.path.join(ROOT, "tests/golden/config1.npz"))
N, M, iters = int(g["N"]), int(g["M"]), int(g["iterations_done"])
bed = O.synth_bed(int(g["seed"]), 0, M, N)
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
with tempfile.TemporaryDirectory() as tmp:
    bedp, phenp = tmp + "/c1.bed", tmp + "/c1.phen"
    O.write_bed(bedp, bed); O.write_phen(phenp, g["y"])
    for mode in ("0", "1"):
        outd = f"{tmp}/o{mode}/"
        args = ["--run-mode", "infere", "--model", "linear", "--bed-file", bedp, "--phen-files", phenp, "--N", str(N), "--Mt", str(M), "--out-dir", outd, "--out-name", "c1"]
        extra = [str(a) for a in g["args"]]
        for k in range(0, len(extra), 2):
            if extra[k] not in {"--N", "--Mt", "--out-dir", "--out-name"}: args += [extra[k], extra[k + 1]]
        r = subprocess.run([ROOT + "/gvamp_b200/bin/main_real"] + args, capture_output=True, text=True, env=dict(os.environ, GVB_ONSAGER_WARM=mode))
        log = r.stdout
        x = np.fromfile(outd + f"c1_it_{iters}.bin")
        a2 = np.array([float(l.split("=")[1]) for l in log.splitlines() if l.startswith("alpha2 = ")])
        gw = np.array([float(l.split("=")[1]) for l in log.splitlines() if l.startswith("gamw = ")])
        sw = [int(l.split("=")[1]) for l in log.splitlines() if l.startswith("bed sweeps this iteration")]
        tt = [float(l.split("=")[1]) for l in log.splitlines() if l.startswith("total iteration time")]
        print("warm", mode, "x1_last relerr", rel(x, g["x1_last"]), "alpha2 maxrel", np.max(np.abs(a2 / g["alpha2_log"] - 1)), "gamw maxrel", np.max(np.abs(gw / g["gamw_log"] - 1)), "sweeps", sw, "iter_s", np.round(tt, 4).tolist())
