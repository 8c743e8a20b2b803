"""SASS instruction histograms of the contraction / panel kernels in lla_b200/libllacuda.so (cuobjdump -sass), the evidence
B200_PROFILING.md asks for: DMMA for the FP64 tensor path, UTC*MMA / LDTM / UTMALDG for tcgen05 + TMEM + TMA.
usage: python tools/sass_hist.py > profiles/r02_sass_histograms.md"""
import collections
import re
import subprocess
import sys

so = "lla_b200/libllacuda.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["lla_dgemm_kernel", "lla_zgemm_kernel", "lla_sgemm_tcgen05_kernel", "lla_dgetf2_panel_kernel", "lla_dpotrf128_kernel",
        "lla_dgesv_batched_smem_kernel", "lla_xgemm_simt_kernel", "lla_xgetf2_panel_kernel", "lla_convert_kernel", "lla_copy_blocks_kernel",
        "lla_dlaswp_dense_kernel", "lla_swap_pack_kernel"]
interesting = ["DMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "REDUX", "DFMA", "DMUL", "DADD",
               "FFMA", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "SYNCS", "MUFU"]
cur, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(hist, dem))
print("# Round 2: SASS instruction histograms (cuobjdump -sass lla_b200/libllacuda.so, sm_100a)\n")
print("Static instruction counts per kernel instantiation; only the mnemonics that identify the execution path are listed.\n")
print("| kernel | total | " + " | ".join(interesting) + " |")
print("|---|---|" + "---|" * len(interesting))
seen = set()
for f, h in hist.items():
    d = names[f]
    if not any(w in d for w in want):
        continue
    short = re.sub(r"^void ", "", d)
    short = re.sub(r"\(.*$", "", short).replace("llacuda::", "")
    if short in seen:
        continue
    seen.add(short)
    tot = sum(h.values())
    if tot < 50:
        continue
    print(f"| `{short[:90]}` | {tot} | " + " | ".join(str(sum(v for k, v in h.items() if k.startswith(i))) for i in interesting) + " |")
