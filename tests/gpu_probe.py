"""Stand-alone GPU probe: runs every igemm parity case in its own subprocess (a device trap in one
case must not poison the others) and writes one JSON line per case to gpurun_out/probe.jsonl.

    python tests/gpu_probe.py [case_substring ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _fmt(v):
    if isinstance(v, float):
        return round(v, 6)
    if isinstance(v, (tuple, list)):
        return [_fmt(x) for x in v]
    if isinstance(v, dict):
        return {k: _fmt(x) for k, x in v.items()}
    return v


def run_one(name):
    import torch  # noqa: F401
    cases = all_cases()
    t0 = time.time()
    r = cases[name]()
    if isinstance(r[0], dict):     # network-level case: (details, ok)
        rec = {"case": name, "ok": bool(r[1]), "s": round(time.time() - t0, 2)}
        rec.update({k: _fmt(v) for k, v in r[0].items()})
    else:
        err, tol = r
        rec = {"case": name, "err": err, "tol": tol, "ok": bool(err <= tol), "s": round(time.time() - t0, 2)}
    print(json.dumps(rec))


def all_cases():
    import igemm_cases
    import ops_cases
    import net_cases
    import fullsize_cases
    import layerwise_cases
    d = dict(igemm_cases.CASES)
    d.update(ops_cases.CASES)
    d.update(net_cases.CASES)
    d.update(fullsize_cases.CASES)
    d.update(layerwise_cases.CASES)
    return d


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        return
    pats = sys.argv[1:]
    names = [n for n in all_cases() if not pats or any(p in n for p in pats)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "probe.jsonl"), "a")
    nfail = 0
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                rec = json.loads(line[-1])
            else:
                rec = {"case": n, "ok": False, "rc": r.returncode, "stderr": r.stderr[-1500:]}
        except subprocess.TimeoutExpired:
            rec = {"case": n, "ok": False, "timeout": True}
        nfail += 0 if rec.get("ok") else 1
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        out.flush()
    print(f"probe: {len(names) - nfail}/{len(names)} ok")
    sys.exit(1 if nfail else 0)


if __name__ == "__main__":
    main()
