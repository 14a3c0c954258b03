"""Classify `compute-sanitizer --tool racecheck --racecheck-report hazard` output for the per-thread TMA rollout kernel.

  python tools/racecheck_classify.py gpurun_out/r2_racecheck_tma_hazards.log <PDP_TSTRIDE> <PDP_TXS> <PDP_TUS> <PDP_TLS>

Every thread of pdp_k_rollout_costate_tma owns a private shared-memory slot (PDP_TSTRIDE doubles) with two private
mbarriers and issues its own cp.async.bulk copies into it.  For each reported hazard this prints which thread's slot the
address lies in (dynamic shared memory starts 1 KB into the window on sm_100) next to the threads racecheck names."""
import collections
import re
import sys


def main(path, tstride, txs, tus, tls, base=1024):
    txt = open(path).read()
    stride = tstride * 8
    pat = re.compile(r"Potential (\w+) hazard detected \(([^)]*)\) at __shared__ (0x[0-9a-f]+) in block \((\d+),0,0\) :\n"
                     r"=========     (\w+) Thread \((\d+),0,0\) at (\w+).*?\n=========     (\w+) Thread \((\d+),0,0\) at (\w+)")
    edges = [("x/u slot 0", txs + tus), ("x/u slot 1", 2 * (txs + tus)), ("out slots", 2 * (txs + tus) + tls + tus),
             ("mbarriers", tstride)]
    c = collections.Counter()
    for m in pat.finditer(txt):
        kind, why, addr, _, a1, t1, f1, a2, t2, f2 = m.groups()
        a = int(addr, 16) - base
        owner, off = a // stride, (a % stride) // 8
        region = next(name for name, end in edges if off < end)
        c[(kind, why, "%s in %s attributed to thread %s" % (a1, f1, t1), "%s in %s by thread %s" % (a2, f2, t2),
           "address in the slot of thread %d (%s)" % (owner, region))] += 1
    for k, v in sorted(c.items(), key=lambda kv: -kv[1]):
        print("%5d  %s" % (v, " | ".join(k)))
    s = re.search(r"RACECHECK SUMMARY: .*", txt)
    print(s.group(0) if s else "no summary line")


if __name__ == "__main__":
    main(sys.argv[1], *[int(a) for a in sys.argv[2:6]])
