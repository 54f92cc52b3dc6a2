"""Summarises a DGRM_TRACE=1 timeline (stderr of a bench / tool run): per rank, how busy the MMA stream was between the
first and the last contraction launch, and how much of the decode ran underneath it.
    DGRM_TRACE=1 torchrun ... bench.py --gpus 8 --steps 3 --warmup 1 --no-e2e --no-check 2> trace.err
    python tools/trace_summary.py trace.err > profiles/r2_trace_n8_summary.txt"""
import collections
import re
import sys

pat = re.compile(r"\[dgrm trace r(\d+)\]\s+([\d.]+) ms\s+\+\s*([\d.]+)\s+(\w+)")


def main(path):
    per = collections.defaultdict(list)
    for line in open(path, errors="replace"):
        m = pat.search(line)
        if m:
            per[int(m.group(1))].append((float(m.group(2)), float(m.group(3)), m.group(4)))
    print(f"# {path}: DGRM_TRACE=1 launch timeline; the last timeline each rank printed (= the timed steps)")
    for r in sorted(per):
        ev = per[r]
        # every dgrm_kernel_times call prints one timeline sorted by start time, with its own time base: keep the last
        cuts = [i for i in range(1, len(ev)) if ev[i][0] < ev[i - 1][0]]
        if cuts:
            ev = ev[cuts[-1]:]
        mma = sorted((t, d) for t, d, k in ev if k in ("syrk", "paircnt"))
        dec = sorted((t, d) for t, d, k in ev if k in ("expand", "gather", "count"))
        if not mma:
            continue
        # split into GRM builds at gaps > 5 ms on the MMA stream
        builds, cur = [], [mma[0]]
        for a, b in zip(mma, mma[1:]):
            if b[0] - (a[0] + a[1]) > 5.0:
                builds.append(cur)
                cur = []
            cur.append(b)
        builds.append(cur)
        span = sum(b[-1][0] + b[-1][1] - b[0][0] for b in builds)
        busy = sum(d for b in builds for _, d in b)
        gaps = [b2[0] - (b1[0] + b1[1]) for bl in builds for b1, b2 in zip(bl, bl[1:])]
        dec_total = sum(d for _, d in dec)
        # decode time that overlaps an MMA launch
        ov = 0.0
        j = 0
        for t, d in dec:
            for mt, md in mma:
                lo, hi = max(t, mt), min(t + d, mt + md)
                if hi > lo:
                    ov += hi - lo
        print(f"rank {r}: {len(builds)} builds, MMA stream span {span:8.2f} ms, busy {busy:8.2f} ms ({100 * busy / span:5.1f} %), "
              f"largest gap {max(gaps) if gaps else 0:6.3f} ms, mean gap {sum(gaps) / max(len(gaps), 1):6.3f} ms; "
              f"decode kernels {dec_total:8.2f} ms, of which {100 * ov / max(dec_total, 1e-9):5.1f} % under an MMA launch")


if __name__ == "__main__":
    main(sys.argv[1])
