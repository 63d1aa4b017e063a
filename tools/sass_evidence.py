"""Disassemble the built library and summarise, per kernel, the SASS mnemonics that prove what it runs on
(tcgen05 = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG/UTMASTG, cp.async = LDGSTS, packed fp32 = FFMA2, ...).

    python tools/sass_evidence.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oscillink_b200", "_lib", "libosc_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "LDGSTS",
        "FFMA2", "FMUL2", "FADD2", "DFMA", "LDS.128", "LDG.E.128", "STG.E.128", "SHFL", "BAR.SYNC", "REDUX", "ATOM",
        "RED.", "LDL", "STL", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts, excerpts, n_inst = None, {}, {}, collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            counts[fn], excerpts[fn] = collections.Counter(), {}
            continue
        if fn is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1).strip()
        n_inst[fn] += 1
        for k in KEYS:
            if k == "HMMA" and "UTCHMMA" in ins:
                continue  # legacy mma.sync only
            if k in ins:
                counts[fn][k] += 1
                excerpts[fn].setdefault(k, ins)
    print(f"# SASS evidence for {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions")
    print("# per kernel, the line after a mnemonic is its first occurrence\n")
    want = ("knn_tc", "knn_rescore", "knn_exact", "batched_ms_kernel<5, 2, 256", "batched_ms_kernel<4, 2, 320",
            "batched_ms_kernel<2, 2, 608", "batched_settle_kernel<4, 2, 320, 2, false", "batched_pack", "pcg_spmm_kernel<4",
            "pcg_update_kernel<4", "pcg_pupdate_kernel<4", "halo_pull_kernel<4", "receipt_full", "normalize_rows",
            "assemble_", "pcg_decide")
    for fn in sorted(counts):
        if not any(w in fn for w in want):
            continue
        c = counts[fn]
        print(f"== {fn}   [{n_inst[fn]} instructions]")
        print("   " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
        for k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "LDTM", "LDGSTS", "FFMA2", "DFMA"):
            if k in excerpts[fn]:
                print(f"     {k:8s} {excerpts[fn][k]}")
        print()


if __name__ == "__main__":
    sys.exit(main())
