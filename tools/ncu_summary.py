"""Turns an `ncu --set full` report (gpurun_out/*.ncu-rep) into the small committed summary
profiles/<name>.csv: one `metric,unit,value` line per metric of the first captured launch.
usage: python tools/ncu_summary.py gpurun_out/r2prof/prof_voxels.ncu-rep profiles/r02_ncu_k_eval_voxels.csv"""
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for h, u, v in zip(hdr, units, vals):
            w.writerow([h, u, v])
    print(out, len(hdr), "metrics")


if __name__ == "__main__":
    main()
