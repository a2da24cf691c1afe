"""Per-basic-block instruction counts of one kernel from an ncu source page
(`ncu -i x.ncu-rep --page source --csv --print-source sass [| gzip]`): which stretches of SASS the executed
warp-instructions sit in.  usage: python tools/ncu_blocks.py profiles/r02_ncu_k_eval_voxels_source_page.csv.gz [top]"""
import csv
import gzip
import io
import sys


def load(path):
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if "Address" in r and "Source" in r)
    data = rows[rows.index(hdr) + 1:]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][ia], 16)
    return rows[0][1] if len(rows[0]) > 1 else "", [(int(r[ia], 16) - base, r[isrc].strip(), int(r[iex]), int(r[ismp])) for r in data]


def blocks(instrs):
    out, cur = [], []
    for ins in instrs:
        cur.append(ins)
        s = ins[1]
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        if op.startswith(("BRA", "BRXU", "BRX", "EXIT", "RET", "JMP")) and not s.startswith("@"):
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    name, instrs = load(path)
    total = sum(i[2] for i in instrs)
    print(f"kernel: {name}\nexecuted warp-instructions: {total}\n")
    print("| SASS offsets | instructions in block | executions of its first / hottest instruction | executed in block | share | what it is (arithmetic it contains) |")
    print("|---|---|---|---|---|---|")
    rows = []
    for b in blocks(instrs):
        t = sum(i[2] for i in b)
        ops = [i[1].split()[1] if i[1].startswith("@") else i[1].split()[0] for i in b]
        key = [o for o in ops if o.startswith(("FADD", "FMUL", "FMNMX", "MUFU", "FFMA", "FCHK", "LDS", "LDTM", "STS", "STTM", "LDL", "STL", "LDG", "STG", "ATOM", "VOTE", "SHFL", "BRXU", "LDCU", "R2UR"))]
        rows.append((t, b[0][0], b[-1][0], len(b), b[0][2], max(i[2] for i in b), " ".join(key[:9])))
    for t, a, z, n, first, hot, key in sorted(rows, reverse=True)[:top]:
        print(f"| {a:#x}-{z:#x} | {n} | {first / 1e6:.2f} M / {hot / 1e6:.2f} M | {t / 1e6:.1f} M | {100 * t / total:.1f} % | {key} |")


if __name__ == "__main__":
    main()
