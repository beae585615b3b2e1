"""Writes profiles/<tag>_sass_score_bounds.txt: instruction-mix summary and the hot loop of
tv5::score_bounds<false,4> as it sits in deep-sfm-revisited_b200/libtv5.so (cuobjdump -sass), so that
FFMA2 / LDS.128 / UBLKCP (1-D TMA bulk copy) / SYNCS (mbarrier) can be checked without disassembling.
    python tools/sass_excerpt.py r2
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = os.path.join(ROOT, "deep-sfm-revisited_b200", "libtv5.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(txt) if "Function :" in l]
out = []
for want, label in (("score_boundsILb0ELi4E", "score_bounds<false,4>  (the roofline kernel: one-sided bound, 4 hypotheses per thread)"),
                    ("solve_frontILi32E", "solve_front<32>")):
    i0 = next(i for i in start if want in txt[i])
    i1 = min([j for j in start if j > i0] + [len(txt)])
    body = [l for l in txt[i0:i1] if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in body:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(2)] += 1
    out.append(f"== {label}: {len(body)} SASS instructions ==")
    out.append("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
    if "score_bounds" in want:
        # the hot loop: the longest run of lines between two backward branches that is dense in FFMA2
        idx = [k for k, l in enumerate(body) if "FFMA2" in l]
        lo, hi = idx[0], idx[-1]
        # widen to the enclosing loop: first LDS before lo, first BRA after hi
        while lo > 0 and "LDS" not in body[lo]:
            lo -= 1
        while hi < len(body) - 1 and " BRA" not in body[hi]:
            hi += 1
        loop = body[lo:hi + 1]
        lops = collections.Counter(re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l).group(2) for l in loop)
        out.append(f"-- region from the first tile load to the loop's back edge: {len(loop)} instructions: "
                   + ", ".join(f"{k} {v}" for k, v in lops.most_common(12)))
        out.append("-- TMA / mbarrier instructions of the kernel:")
        out += ["   " + l.strip() for l in body if re.search(r"UBLKCP|SYNCS|ARRIVE|UTMA", l)][:12]
        out.append("-- first 60 instructions of that region:")
        out += ["   " + l.strip() for l in loop[:60]]
    out.append("")
dst = os.path.join(ROOT, "profiles", f"{tag}_sass_score_bounds.txt")
open(dst, "w").write("\n".join(out) + "\n")
print(dst, len(out), "lines")
