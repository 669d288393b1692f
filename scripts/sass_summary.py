"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library -> profiles/sass_summary.txt
(tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, legacy mma.sync = HMMA).  No GPU needed."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multiscale_variational_autoencoder_b200", "libmvae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pats = collections.OrderedDict([("UTCHMMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
                                ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("HMMA", r"\bHMMA\b"), ("RED/ATOM", r"\b(RED|ATOM[GS]?)\b")])
kern, rows, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if kern:
        for k, p in pats.items():
            if re.search(p, line):
                rows[kern][k] += 1
with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
    f.write(f"# cuobjdump -sass multiscale_variational_autoencoder_b200/libmvae_b200.so  (arch: {', '.join(sorted(arch))}); scripts/sass_summary.py\n")
    f.write(f"# instruction counts per kernel; kernels without any of these mnemonics are listed by name only at the end\n")
    f.write(f"{'kernel':100s} " + " ".join(f"{k:>8s}" for k in pats) + "\n")
    tot = collections.Counter()
    plain = []
    for k, c in rows.items():
        name = re.sub(r"\(.*", "", demangle(k))[:100]
        if sum(c[p] for p in pats if p != "RED/ATOM") == 0:
            plain.append(name)
            continue
        tot.update(c)
        f.write(f"{name:100s} " + " ".join(f"{c[p]:8d}" for p in pats) + "\n")
    f.write(f"{'TOTAL':100s} " + " ".join(f"{tot[p]:8d}" for p in pats) + "\n")
    f.write(f"# {len(plain)} CUDA-core kernels without tensor-core / TMA instructions: " + ", ".join(sorted(set(plain))) + "\n")
print(open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read()[:3000])
