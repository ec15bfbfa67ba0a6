"""cuobjdump -sass opcode summary of the tensor-core / TMEM / TMA kernels of libll_b200.so (profiles/r02_sass_tensor_kernels.txt)."""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                  "imagecompressionlearnedliftingandlearnedtreebasedmodels_b200", "libll_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", sass), filt))
KEY = re.compile(r"^(UTC\w*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|SYNCS|UCGABAR|FFMA2|LDGSTS|STG\.E\.ENL2\.256|UTCCP)")
print("cuobjdump -sass of libll_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a): tensor-core / TMEM / TMA / mbarrier opcodes per kernel")
print("(the PTX names never appear in SASS: tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, cp.async.bulk.tensor = UTMALDG, tcgen05.commit = UTCBAR,")
print(" cta_group::2 = the .2CTA suffix, mbarrier = SYNCS)\n")
for blk in sass.split("Function : ")[1:]:
    mangled = blk.split("\n", 1)[0].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][\w.]*)", blk, re.M)
    cnt = collections.Counter(o for o in ops if KEY.match(o))
    if not any(o.startswith(("UTC", "LDTM", "STTM", "UTMALDG")) for o in cnt) and "dwt97" not in mangled:
        continue
    print("== " + names.get(mangled, mangled))
    print(f"   {len(ops)} instructions; " + ", ".join(f"{o} x{n}" for o, n in cnt.most_common()))
    shown = 0
    for line in blk.split("\n"):
        if re.search(r"\b(UTC\w*MMA|UTMALDG|LDTM)", line) and shown < 4:
            print("     " + line.strip()[:150]); shown += 1
    print()
