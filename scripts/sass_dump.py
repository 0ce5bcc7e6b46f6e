"""Per-kernel SASS listings + mnemonic histogram of libmobgt.so (run here, no GPU): python scripts/sass_dump.py profiles/sass"""
import collections, os, re, subprocess, sys
out = sys.argv[1] if len(sys.argv) > 1 else "profiles/sass"
os.makedirs(out, exist_ok=True)
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mobgt_b200", "lib", "libmobgt.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
summary = []
for p in parts[1:]:
    mangled = p.split("\n", 1)[0].strip()
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("mobgt::", "")
    short = re.sub(r"[^\w]+", "_", short).strip("_")
    body = [l for l in p.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
    ins = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l) for l in body]
    ins = [re.sub(r"\s*/\*.*", "", l).strip() for l in ins]
    ins = [l for l in ins if l]
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l).split()[0].rstrip(";") for l in ins)
    with open(os.path.join(out, short + ".sass"), "w") as f:
        f.write(f"// {name}\n// {len(ins)} SASS instructions (cuobjdump -sass, sm_100a)\n")
        f.write("\n".join(ins) + "\n")
    key = {k: v for k, v in hist.items() if re.match(r"(UTC|UTMA|UBLKCP|SYNCS|LDTM|STTM|UTCBAR|UTCHMMA|UTCMMA|UTCCP|LDGSTS|REDG|RED|ATOM|BAR|LDS|STS|LDG|STG|VIMNMX|VMNMX|HFMA2|FFMA|MUFU|ELECT|CCTL|UCGABAR|MEMBAR)", k)}
    summary.append((short, len(ins), dict(sorted(key.items(), key=lambda kv: -kv[1]))))
with open(os.path.join(out, "SUMMARY.txt"), "w") as f:
    for s, n, k in summary:
        f.write(f"{s}: {n} instr; " + ", ".join(f"{a}={b}" for a, b in k.items()) + "\n")
print(open(os.path.join(out, "SUMMARY.txt")).read())
