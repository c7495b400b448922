"""Per-kernel SASS evidence of the shipped library: which tensor / TMA / barrier instructions each kernel of
qodeapplications_b200/libxr_b200.so contains.   python tools/sass_summary.py > profiles/r02_sass_summary.txt
(cuobjdump -sass of the in-tree .so; run where it was built).  sm_100a has no f64 kind in tcgen05, so the FP64 tensor
instruction is DMMA (mma.sync.m8n8k4.f64); UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk (1-D
bulk TMA), SYNCS = mbarrier operations; HMMA/IMMA/UTC*MMA must be absent."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "qodeapplications_b200", "libxr_b200.so")
elf = subprocess.run(["cuobjdump", "-lelf", lib], stdout=subprocess.PIPE, text=True).stdout
archs = collections.Counter(re.findall(r"sm_\d+a?", elf))
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
WATCH = ["DMMA", "DFMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "LDTM", "ATOMG", "REDG", "RED."]
per, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        per[name] = collections.Counter()
        continue
    if name:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            per[name]["total"] += 1
            for w in WATCH:
                if m.group(1).startswith(w):
                    per[name][w.rstrip(".")] += 1
print("libxr_b200.so: cubins", dict(archs))
cols = [w.rstrip(".") for w in WATCH]
print("%-64s %7s " % ("kernel", "instrs") + " ".join("%7s" % c for c in cols))
tot = collections.Counter()
for k, c in per.items():
    print("%-64s %7d " % (k[:64], c["total"]) + " ".join("%7d" % c[w] for w in cols))
    tot.update(c)
print("%-64s %7d " % ("ALL KERNELS", tot["total"]) + " ".join("%7d" % tot[w] for w in cols))
