"""Opcode inventory of the shipped library per kernel (cuobjdump -sass): which kernels carry tcgen05 / TMA / bulk-copy code.
Usage: python tools/sass_inventory.py > profiles/r2_sass_inventory.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "neurecon_b200", "lib", "libneurecon_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
keys = [("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
        ("UTCBAR", r"\bUTCBAR"), ("MUFU.TANH", r"MUFU\.TANH"), ("FFMA2", r"\bFFMA2"), ("DISCARD", r"\bCCTL\.[A-Z.]*DISCARD|\bDISCARD"),
        ("LDG.256", r"LDG\.E\.[A-Z.]*256"), ("STG.256", r"STG\.E\.[A-Z.]*256")]
print("SASS inventory of neurecon_b200/lib/libneurecon_b200.so (end of round 2), cuobjdump -sass, instruction counts per kernel")
print("(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,")
print(" MUFU.TANH = tanh.approx, FFMA2 = packed fp32x2 FMA, LDG.256 / STG.256 = 256-bit global accesses); kernels without any of these by name only\n")
blocks = re.split(r"Function : \S+", sass)[1:]
plain = []
for nm, body in zip(names, blocks):
    nm = re.sub(r"\(anonymous namespace\)::", "", nm)
    c = collections.OrderedDict((k, len(re.findall(p, body))) for k, p in keys)
    n_inst = len(re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/", body, re.M))
    tags = " ".join("%s=%d" % (k, v) for k, v in c.items() if v)
    if tags:
        print("%-104s instr=%-5d %s" % (nm[:104], n_inst, tags))
    else:
        plain.append(nm.split("(")[0])
print("\nwithout tensor / bulk-copy instructions: " + ", ".join(sorted(set(plain))))
