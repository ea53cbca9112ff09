"""Top stall-sample SASS lines of one kernel from an .ncu-rep (source page).
    python tools/ncu_hot_sass.py gpurun_out/decode.ncu-rep gemm_bf16_kernel [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, launch=0, top=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else kernel
    hdr = rows[1]
    si, ai, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":   # next launch's table
            break
        if len(r) > ai and r[ai].isdigit():
            body.append(r)
    total = sum(int(r[ai] or 0) for r in body) or 1
    print(f"`{name[:90]}` — {total} stall samples")
    print()
    print("| % of samples | SASS | instr executed |")
    print("|---|---|---|")
    for r in sorted(body, key=lambda r: -int(r[ai] or 0))[:top]:
        print(f"| {100.0 * int(r[ai] or 0) / total:.1f} | `{r[si].strip()[:70]}` | {r[ii]} |")


if __name__ == "__main__":
    a = sys.argv
    main(a[1], a[2], int(a[3]) if len(a) > 3 else 0, int(a[4]) if len(a) > 4 else 12)
