"""Per-role instruction and stall-sample accounting of tc_gemm_kernel from an `ncu --set full --import-source on`
report (reads the report with `ncu -i ... --page source --csv`; no GPU needed).

    python scripts/ncu_role_sections.py gpurun_out/r1b_tc_gemm.ncu-rep [n_ctas n_k_blocks]

The kernel's three roles execute disjoint code ranges; this finds them by their signature instructions
(TMA producer: UTMALDG; MMA issuer: UTCHMMA/UTCBAR; workers: STTM/LDTM) and reports warp-instructions per CTA per
k-block and where the warp-state samples of each role fall (mbarrier waits, tensor-queue back-pressure, the rest).
This is the analysis behind profiles/r1b_mma_warp_source_samples.txt.
"""
import csv
import subprocess
import sys


def load(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    return hdr, body


def main():
    report = sys.argv[1]
    n_ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    n_kb = int(sys.argv[3]) if len(sys.argv) > 3 else 41
    hdr, body = load(report)
    si, ni, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda x: int(x) if x.isdigit() else 0
    idx = lambda pat: [i for i, r in enumerate(body) if pat in r[si]]
    tma, mma, bar = idx("UTMALDG"), idx("UTCHMMA"), idx("UTCBAR")
    sttm, ldtm = idx("STTM"), idx("LDTM")
    if not (tma and mma and bar and sttm):
        sys.exit("signature instructions not found: is this a tc_gemm_kernel report?")
    # role ranges: from the previous role's last signature instruction to this role's last one (+ loop tail)
    sections = {"tma producer": (max(0, tma[0] - 60), tma[-1] + 30),
                "workers (split + drain)": (tma[-1] + 30, max(ldtm[:2] + sttm) + 60),
                "mma issuer": (mma[0] - 150, bar[-1] + 25)}
    total_samples = sum(num(r[ni]) for r in body)
    print("kernel: %d warp-state samples, %d warp-instructions per CTA" % (total_samples, sum(num(r[ei]) for r in body) // n_ctas))
    for name, (a, b) in sections.items():
        sec = body[max(a, 0):b]
        samples = sum(num(r[ni]) for r in sec)
        inst = sum(num(r[ei]) for r in sec)
        wait = sum(num(r[ni]) for r in sec if "BRA" in r[si] and any(hdr[i] == "stall_long_sb" and num(r[i]) for i in stall))
        at_mma = sum(num(r[ni]) for r in sec if "UTCHMMA" in r[si])
        print("%-24s %5d samples (%4.1f%%)  %7.1f warp-instr / CTA / k-block   mbarrier waits %d, at UTCHMMA %d, other %d"
              % (name, samples, 100.0 * samples / max(total_samples, 1), inst / n_ctas / n_kb, wait, at_mma, samples - wait - at_mma))


if __name__ == "__main__":
    main()
