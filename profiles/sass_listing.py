"""Regenerates the SASS evidence under profiles/r02/ from the built library (no GPU needed):

    python profiles/sass_listing.py

sass_mnemonics.txt: per-kernel counts of the mnemonics that prove tcgen05 / TMEM / TMA / mbarrier / cp.async / IMMA use;
sass_<kernel>.txt: full listings of the kernels the bench line's rooflines name."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vectorragquantization_b200", "libvrq.so")
OUT = os.path.join(ROOT, "profiles", "r02")
KEYS = ["UTCOMMA.2CTA", "UTCOMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "IMMA.16832.S8.S8", "USETMAXREG",
        "FMNMX3", "POPC", "LOP3", "DADD", "DFMA", "IDP.4A", "IDP.2A", "ATOMS"]
LISTINGS = {
    "sass_hamming_scan_mma_pair_dense.txt": r"hamming_scan_mma_kernel<4, 2, false, 4>",
    "sass_hamming_scan_mma_wide64.txt": r"hamming_scan_mma_wide_kernel<64>",
    "sass_rescore_binary_imma.txt": r"rescore_binary_imma_kernel[<(]",
    "sass_rescore_int8cos_imma_w12_s1.txt": r"rescore_int8cos_imma_kernel<12, 1, false>",
}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    total = collections.Counter()
    lines = []
    for name, blk in zip(names, blocks):
        name = re.sub(r"vrq::\(anonymous namespace\)::", "", name)
        ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
        c = collections.Counter()
        for m in ins:
            for k in KEYS:
                if m == k or m.startswith(k + "."):
                    c[k] += 1
                    break
        total.update(c)
        lines.append(f"{name}\n    {len(ins)} instructions: " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
        for fn, pat in LISTINGS.items():
            if re.search(pat, name):
                # drop the encoding words: one line per instruction
                body = "\n".join(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln).rstrip() for ln in blk.split("\n")
                                 if not re.match(r"^\s*/\* 0x[0-9a-f]{16} \*/\s*$", ln))
                with open(os.path.join(OUT, fn), "w") as f:
                    f.write(f"{name}\n(cuobjdump -sass vectorragquantization_b200/libvrq.so, sm_100a; encoding words dropped)\n{body}\n")
    head = ("cuobjdump -sass vectorragquantization_b200/libvrq.so (sm_100a), mnemonic counts per kernel; built from HEAD of round 2 (profiles/sass_listing.py).\n"
            "Proof of: tcgen05 MMA (UTCOMMA / UTCOMMA.2CTA / UTCIMMA), tensor memory (LDTM / STTM), TMA (UTMALDG, UBLKCP), mbarriers (SYNCS, UTCBAR),\n"
            "register reallocation between warp roles (USETMAXREG), cp.async (LDGSTS), mma.sync s8 (IMMA.16832.S8.S8).  No BMMA: b1 xor.popc is\n"
            "emulated on sm_100a, hence the +-1 formulation.\n\n")
    with open(os.path.join(OUT, "sass_mnemonics.txt"), "w") as f:
        f.write(head + "TOTAL  " + "  ".join(f"{k}={v}" for k, v in sorted(total.items())) + "\n\n" + "\n".join(lines) + "\n")
    print("TOTAL", dict(total))


if __name__ == "__main__":
    main()
