"""Static instruction mix of the kernels in a built libgenfft_cuda.so (no GPU needed): per kernel the SASS instruction
count by class, and the issue-slot floor it implies for a pass (one-shot grid: every thread runs the kernel body once
per tile, so instructions per point ~ body / P).   usage: python tools/sass_mix.py <lib.so> <regex on demangled name>"""
import collections, re, subprocess, sys

FP = {"FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "DADD", "DMUL", "DFMA"}
INT = {"IMAD", "IADD3", "LEA", "LOP3", "SHF", "ISETP", "SEL", "VIMNMX", "MOV", "PRMT", "IABS", "UMOV", "UIMAD", "UIADD3", "ULEA", "USHF", "ULOP3"}
MEM = {"LDG", "STG", "LDS", "STS", "LD", "ST", "LDL", "STL", "LDC", "LDCU", "UBLKCP", "ATOMG", "RED", "SYNCS"}


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, res = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            res[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            res[name][m.group(1)] += 1
    for n, c in sorted(res.items()):
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().replace("genfft_cuda::", "")
        if not re.search(pat, d):
            continue
        tot = sum(c.values()) - c["NOP"]
        fp = sum(v for k, v in c.items() if k in FP)
        it = sum(v for k, v in c.items() if k in INT)
        mem = sum(v for k, v in c.items() if k in MEM)
        print(f"{d[:120]}\n   {tot:5d} instructions: fp {fp} (packed {c['FADD2'] + c['FMUL2'] + c['FFMA2']}), integer/address {it}, memory {mem}"
              f" (LDG {c['LDG']} STG {c['STG']} LDS {c['LDS']} STS {c['STS']} local {c['LDL'] + c['STL']}), other {tot - fp - it - mem}")


if __name__ == "__main__":
    main()
