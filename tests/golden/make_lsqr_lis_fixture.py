"""Extracts the numbers of the reference's committed golden log into a JSON fixture.

Run in the build container (where /root/reference exists):
    python tests/golden/make_lsqr_lis_fixture.py
It reads /root/reference/test/LSQR.LIS (the output of test/lsqrtest_module.f90 for the 18
LSTP problems) and the known answers of README.md:55-58, and writes
    tests/golden/lsqr_lis.json
which the CPU tests compare the oracle against.  Only numbers are kept (headers, the
iteration rows for itn <= 10, the exit block, the xcheck report, the first 8 solution
components and the verdict); the GPU box never reads /root/reference.
"""
import json
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/LSQR.LIS"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lsqr_lis.json")

NUM = r"[-+]?\d\.\d+E[-+]\d+"


def parse_problem(block, first_line_no):
    p = {"line": first_line_no}
    m = re.search(r"P\(\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(" + NUM + r")\s*\)", block)
    p.update(m=int(m[1]), n=int(m[2]), nduplc=int(m[3]), npower=int(m[4]), damp=float(m[5]))
    m = re.search(r"Condition no\. =\s*(" + NUM + r")\s+Residual function =\s*(" + NUM + ")", block)
    p.update(gen_acond=float(m[1]), gen_rnorm=float(m[2]))
    m = re.search(r"aprod seems OK\.\s+Relative error =\s*(" + NUM + ")", block)
    p["acheck_relerr"] = float(m[1]) if m else None
    m = re.search(r"atol\s+=\s*(" + NUM + r")\s+conlim =\s*(" + NUM + ")", block)
    p.update(atol=float(m[1]), conlim=float(m[2]))
    m = re.search(r"itnlim =\s*(\d+)", block)
    p["itnlim"] = int(m[1])
    rows = []
    for line in block.splitlines():
        mm = re.match(r"^\s*(\d+)((?:\s+" + NUM + r"){4,10})\s*$", line)
        if mm:
            vals = [float(t) for t in mm[2].split()]
            itn = int(mm[1])
            if itn <= 10:
                rows.append([itn] + vals)
    p["rows"] = rows    # [itn, x1, rnorm, test1, test2, (anorm, acond, phi, dknorm, dxk, alfopt)]
    m = re.search(r"istop\s+=\s*(\d+)\s+itn\s+=\s*(\d+)", block)
    p.update(istop=int(m[1]), itn=int(m[2]))
    for key in ("anorm", "acond", "bnorm", "xnorm", "rnorm", "arnorm"):
        m = re.search(r"Exit  LSQR\..*?\b" + key + r"\s*=\s*(" + NUM + ")", block)
        p[key] = float(m[1])
    m = re.search(r"inform\s+=\s*(\d+)", block)
    p["xcheck_inform"] = int(m[1])
    for key, pat in (("rho1", r"norm\(r\)\s+=\s*"), ("sigma1", r"norm\(A'r\)\s+=\s*"),
                     ("xtest1", r"test1\s+=\s*"), ("xtest2", r"test2\s+=\s*"), ("xtest3", r"test3\s+=\s*")):
        m = re.search(pat + "(" + NUM + ")", block)
        p[key] = float(m[1])
    sol = re.search(r"Solution  x:\s*\n((?:.*\n){2})", block)
    xs = re.findall(r"\d+\s+([-+]?\d*\.\d+(?:E[-+]\d+)?)", sol[1])
    p["x_head"] = [float(t) for t in xs]
    m = re.search(r"LSQR  appears to (be successful|have failed)\.\s+Relative error in  x  =\s*(" + NUM + ")", block)
    p["success"] = m[1] == "be successful"
    p["enorm"] = float(m[2])
    return p


def main():
    text = open(SRC).read()
    lines = text.splitlines(keepends=True)
    starts = [i for i, l in enumerate(lines) if "Least-Squares Test Problem" in l]
    problems = []
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else len(lines)
        problems.append(parse_problem("".join(lines[s:e]), s + 1))
    out = {
        "source": "test/LSQR.LIS of jacobwilliams/LSQR (numbers only), made with fourpi = real32(4.0*3.141592)",
        "fourpi": 12.566368103027344,
        "problems": problems,
        "readme_ez": {   # README.md:55-58
            "istop": 1, "x": [1.242424E+00, -6.060606E-02, -4.040404E-02],
        },
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {OUT}: {len(problems)} problems")


if __name__ == "__main__":
    main()
