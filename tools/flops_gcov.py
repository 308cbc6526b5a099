#!/usr/bin/env python3
"""Pin the algorithmic flop count of one BSIM4 instance-evaluation.

Input: the `b4ld.c.gcov` file gcov writes after a run of the reference compiled with `--coverage`
(oracle/flops_gcov.sh builds that binary from /root/reference and runs one Monte-Carlo sample of
config 3).  For every executed statement of BSIM4load the floating-point operators in the SOURCE
TEXT are counted -- add, subtract, multiply, divide (compound assignments included), and calls of
sqrt / exp / log / pow / fabs / MAX / MIN count one each -- and multiplied by the execution count.
Index arithmetic, comparisons, pointer dereferences, unary minus and `->` are not operators here.
The sum divided by the number of instance evaluations is F_alg.

usage: tools/flops_gcov.py oracle/_ref/gcov/b4ld.c.gcov
"""
import re
import sys

CALLS = ("sqrt", "exp", "log", "pow", "fabs", "MAX", "MIN", "atan", "tanh")


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", lambda m: re.sub(r"[^\n]", " ", m.group(0)), text, flags=re.S)


def count_ops(stmt):
    """(add/sub, mul, div, calls{}) of one statement's text"""
    s = stmt
    s = re.sub(r"->", "  ", s)
    s = re.sub(r"\+\+|--", "  ", s)
    s = re.sub(r"[<>=!]=", "  ", s.replace("+=", " + ").replace("-=", " - ").replace("*=", " * ").replace("/=", " / "))
    calls = {c: len(re.findall(r"\b%s\s*\(" % c, s)) for c in CALLS}
    # numeric literals with exponent signs (1.0e-3) must not count as subtractions
    s = re.sub(r"(\d\.?\d*)[eE][+-]?\d+", "1", s)
    add = mul = div = 0
    prev = None                     # previous significant character
    for i, ch in enumerate(s):
        if ch in " \t\n":
            continue
        if ch in "+-":
            if prev is not None and (prev.isalnum() or prev in ")]_."):
                add += 1                # binary
        elif ch == "*":
            if prev is not None and (prev.isalnum() or prev in ")]_."):
                mul += 1                # binary (a dereference follows an operator or '(')
        elif ch == "/":
            div += 1
        prev = ch
    return add, mul, div, calls


def main(path):
    rows = []                       # (count or None, lineno, text)
    for ln in open(path, errors="replace"):
        m = re.match(r"\s*([^:]+):\s*(\d+):(.*)$", ln.rstrip("\n"))
        if not m:
            continue
        c, no, text = m.group(1).strip(), int(m.group(2)), m.group(3)
        if no == 0:
            continue
        c = c.rstrip("*")
        cnt = None if c == "-" else (0 if c in ("#####", "=====") else int(c))
        rows.append((cnt, no, text))
    src = strip_comments("\n".join(r[2] for r in rows)).split("\n")
    # restrict to BSIM4load's body: from its first line to the start of BSIM4polyDepletion's definition
    start = next(i for i, t in enumerate(src) if re.match(r"\s*BSIM4load\s*\(", t) or "BSIM4LoadOMP(" in t and "int" in t)
    # statements: split on ';' while tracking the lines they cover
    totals = {"add": 0, "mul": 0, "div": 0}
    totals.update({c: 0 for c in CALLS})
    evals = None
    stmt, lines = "", []
    for i in range(start, len(src)):
        t = src[i]
        if re.match(r"\s*#", t):
            continue
        for piece in re.split(r"(;|\{|\})", t):
            if piece in (";", "{", "}"):
                cnts = [rows[j][0] for j in lines if rows[j][0] is not None]
                n = max(cnts) if cnts else 0
                if n and stmt.strip():
                    a, m_, d, calls = count_ops(stmt)
                    # control headers (`for (...)`, `if (...)`): comparisons only, their arithmetic is index work
                    if not re.match(r"\s*(for|while)\b", stmt):
                        totals["add"] += a * n; totals["mul"] += m_ * n; totals["div"] += d * n
                        for c, k in calls.items():
                            totals[c] += k * n
                    if evals is None and "Check = Check1 = Check2 = 1" in stmt:
                        evals = n
                stmt, lines = "", []
            else:
                stmt += piece + " "
                if piece.strip():
                    lines.append(i)
    if not evals:
        raise SystemExit("instance-evaluation marker line not found")
    per = {k: v / evals for k, v in totals.items()}
    flop = per["add"] + per["mul"] + per["div"] + per["sqrt"] + per["exp"] + per["log"] + per["pow"]
    print(f"instance evaluations: {evals}")
    for k in ("add", "mul", "div", "sqrt", "exp", "log", "pow", "fabs", "MAX", "MIN"):
        print(f"  {k:5s} per evaluation: {per[k]:9.1f}")
    print(f"F_alg (add+mul+div+sqrt+exp+log+pow, each 1): {flop:.0f} flop per BSIM4 instance-evaluation")
    return flop


if __name__ == "__main__":
    main(sys.argv[1])
