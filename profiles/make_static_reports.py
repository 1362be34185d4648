"""Static evidence of the built library, no GPU needed: per-kernel registers / spills / shared memory from the ptxas log of
the last build (profiles/r2_ptxas.txt) and a SASS opcode histogram per kernel from cuobjdump (profiles/r2_sass_opcodes.md):
packed FP32 (FFMA2 / FADD2 / FMUL2), bulk async copies (UBLKCP), REDUX, DFMA; no tensor-core opcodes (HMMA / UTC*MMA) by design.
usage: python profiles/make_static_reports.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dsp-speech-recognition_b200")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = []
    for o in out:
        o = o.replace("(anonymous namespace)::", "").replace("dspfe::", "")
        o = re.sub(r"^void ", "", o)
        res.append(re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", o))     # drop the parameter list
    return res


def ptxas():
    log = open(os.path.join(PKG, "libdspfe.build.log")).read()
    rows = []
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", log):
        rows.append((m.group(1), int(m.group(5)), int(m.group(3)), int(m.group(4)), m.group(7) or "0"))
    names = demangle([r[0] for r in rows])
    with open(os.path.join(ROOT, "profiles", "r2_ptxas.txt"), "w") as f:
        f.write("kernel | registers | spill stores B | spill loads B | static smem B   (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v)\n")
        for n, r in sorted(zip(names, rows)):
            f.write(f"{n} | {r[1]} | {r[2]} | {r[3]} | {r[4]}\n")
    return len(rows)


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "libdspfe.so")], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1); per[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    names = demangle(list(per))
    keys = ["FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "REDUX", "POPC", "LOP3",
            "HMMA", "UTCHMMA", "UTCQMMA", "LDTM"]
    with open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.md"), "w") as f:
        f.write("# SASS opcode histogram of libdspfe.so (static instruction counts per kernel, `cuobjdump -sass`)\n\n")
        f.write("| kernel | total | " + " | ".join(keys) + " |\n|---|---|" + "---|" * len(keys) + "\n")
        for n, (_, c) in sorted(zip(names, per.items())):
            if sum(c.values()) < 200:
                continue
            f.write(f"| {n} | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in keys) + " |\n")
        tot = collections.Counter()
        for c in per.values():
            tot.update(c)
        f.write("\nWhole library: " + ", ".join(f"{k} {tot.get(k, 0)}" for k in keys) + ".\n")
        f.write("No tensor-core opcodes (HMMA / UTC*MMA / LDTM): by `north_star` the FFTs stay off the tensor cores and the mel / DCT contractions on FP32 FFMA "
                "(packed as FFMA2).  `UBLKCP` + `SYNCS` = `cp.async.bulk` 1-D TMA copies with mbarrier completion (K1's PCM staging).\n")
    return len(per)


if __name__ == "__main__":
    print(ptxas(), "kernels in r2_ptxas.txt;", sass(), "functions in r2_sass_opcodes.md")
