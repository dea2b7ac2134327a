"""Which Blackwell-native instructions each kernel of libnrf_b200.so contains (cuobjdump -sass; no GPU needed).
    python scripts/sass_summary.py > profiles/r02_sass_opcodes.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "real-robot-nerf-actor_b200", "libnrf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS", "UCGABAR",
         "HMMA", "FFMA", "FADD2", "F2FP", "HMNMX2", "HADD2", "REDG", "RED.", "ATOMG", "MUFU", "SHFL", "LDGSTS", "ELECT", "USETMAXREG"]
kern = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "").replace("nrf::", "")
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WATCH:
            if op.startswith(w) or (w.endswith(".") and op.startswith(w[:-1] + ".")):
                counts[kern][op] += 1
print("# SASS opcode summary of `libnrf_b200.so` (sm_100a, `cuobjdump -sass`): Blackwell-native instructions per kernel\n")
print("`UTCHMMA` = tcgen05.mma (kind::f16; `.2CTA` = cta_group::2), `LDTM` / `STTM` = tcgen05.ld / tcgen05.st (tensor memory), "
      "`UTMALDG` / `UTMASTG` = TMA tensor load / store (cp.async.bulk.tensor), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, "
      "`USETMAXREG` = setmaxnreg, `FADD2` = add.f32x2, `F2FP` = packed fp32 -> bf16 / fp16 conversion, `REDG` = red.global.\n")
print("| kernel | SASS instructions | tensor / TMEM / TMA | other |\n|---|---:|---|---|")
for k, c in counts.items():
    if c["_total"] < 40:
        continue
    tc = {o: n for o, n in c.items() if o != "_total" and re.match(r"UTC|LDTM|STTM|UTMA|SYNCS|UCGABAR|USETMAXREG|ELECT", o)}
    other = {o: n for o, n in c.items() if o != "_total" and o not in tc}
    fold = lambda d: ", ".join(f"{o} x{n}" for o, n in sorted(d.items(), key=lambda kv: -kv[1])[:16]) or "-"
    print(f"| `{k[:90]}` | {c['_total']} | {fold(tc)} | {fold(other)} |")
