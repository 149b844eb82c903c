"""Writes profiles/r2_fused_sass.md: what `cuobjdump -sass` shows for the kernels of libdenseflow_b200.so that use TMA /
packed fp32 — the TMA and mbarrier instructions, the opcode mix of the lean inner loop of k_tvl1_pair, and the register /
spill summary from ptxas.  No GPU needed."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "denseflow_b200", "lib", "obj_default")


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            funcs[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    return funcs


def opcode(l):
    t = l.split()
    op = t[1] if not t[1].startswith("@") else t[2]
    return op.split(".")[0]


def ptxas_info(src, pattern):
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-I", os.path.join(ROOT, "include"),
           "-Xptxas", "-v", "-c", os.path.join(ROOT, "denseflow_b200", "csrc", src), "-o", "/dev/null"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr.splitlines()
    res = []
    for i, l in enumerate(err):
        if "Compiling entry function" in l and re.search(pattern, l):
            res += [x.strip() for x in err[i + 1:i + 4] if "Used" in x or "spill" in x]
    return res


lines = ["# r2 SASS evidence (cuobjdump -sass of denseflow_b200/lib/obj_default/*.o, sm_100a)", ""]
f = sass("tvl1_fused.o")
name, body = next((k, v) for k, v in f.items() if "k_tvl1_pair" in k)
lines += ["## k_tvl1_pair (%d SASS instructions)" % len(body), "", "ptxas: " + "; ".join(ptxas_info("tvl1_fused.cu", "k_tvl1_pair")), "",
          "TMA / mbarrier instructions (UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier):", "```"]
lines += [l for l in body if re.search(r"UTMALDG|SYNCS|UTMAPF|FENCE\.VIEW", l)][:24]
lines += ["```", ""]
# lean inner loop = the backward branch whose body has the fewest instructions among bodies containing 48 MUFU
best = None
for i, l in enumerate(body):
    m = re.search(r"@P\d BRA (0x[0-9a-f]+)", l)
    if m:
        tgt = int(m.group(1), 16)
        adr = int(re.search(r"/\*([0-9a-f]+)\*/", l).group(1), 16)
        if tgt < adr:
            j = next(k for k, x in enumerate(body) if int(re.search(r"/\*([0-9a-f]+)\*/", x).group(1), 16) >= tgt)
            seg = body[j:i + 1]
            nm = sum(1 for x in seg if "MUFU" in x)
            if nm >= 40 and (best is None or len(seg) < len(best)):
                best = seg
if best:
    h = collections.Counter(opcode(l) for l in best)
    lines += ["Lean inner loop (interior tile, no error sum): %d instructions per iteration of a warp = 16 pixels -> %.1f per pixel-iteration"
              % (len(best), len(best) / 16.0), "", "| opcode | count |", "|---|---|"]
    lines += ["| %s | %d |" % kv for kv in h.most_common(16)]
    lines += ["", "first packed-arithmetic instructions of the loop:", "```"] + [l for l in best if re.search(r"FFMA2|FADD2|FMUL2", l)][:6] + ["```", ""]
f = sass("farneback.o")
for k, body in f.items():
    if "k_box_solve_update_tma" in k:
        lines += ["## k_box_solve_update_tma<6> (%d SASS instructions)" % len(body), "", "ptxas: " + "; ".join(ptxas_info("farneback.cu", "k_box_solve_update_tma")),
                  "", "```"] + [l for l in body if re.search(r"UTMALDG|SYNCS", l)][:12] + ["```", ""]
open(os.path.join(ROOT, "profiles", "r2_fused_sass.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
