"""Summarise an EKB200_GEMM_TRACE file (one line per engine GEMM: stage m n k flags tri_keep splitk ms) by stage and
by shape class: seconds, FLOPs, TFLOP/s, launches -- which shapes of a stage run below the engine's large-product rate."""
import collections
import sys


def main(path, top=14):
    rows = collections.defaultdict(lambda: [0.0, 0.0, 0])
    stage_tot = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for ln in open(path):
        w = ln.split()
        if len(w) != 8:
            continue
        st, m, n, k, fl, tri, sk, ms = w[0], int(w[1]), int(w[2]), int(w[3]), int(w[4]), int(w[5]), int(w[6]), float(w[7])
        elems = m * n if tri < 0 else m * n - 0.5 * min(m, n) * (min(m, n) - 1)
        fl_ = 2.0 * k * elems

        def bucket(x):  # power-of-two class
            b = 1
            while b * 2 <= x:
                b *= 2
            return b
        key = (st, bucket(m), bucket(n), bucket(k), fl & 7, tri >= 0, sk)
        for d, kk in ((rows, key), (stage_tot, st)):
            d[kk][0] += ms * 1e-3
            d[kk][1] += fl_
            d[kk][2] += 1
    for st, (s, f, c) in sorted(stage_tot.items(), key=lambda x: -x[1][0]):
        print(f"== {st}: {s:.3f} s, {f / max(s, 1e-12) / 1e12:.1f} TF, {c} launches")
        sub = [(k, v) for k, v in rows.items() if k[0] == st]
        for k, (s2, f2, c2) in sorted(sub, key=lambda x: -x[1][0])[:top]:
            print(f"   m>={k[1]:6d} n>={k[2]:6d} k>={k[3]:6d} flags={k[4]} tri={int(k[5])} splitk={k[6]}: "
                  f"{s2 * 1e3:8.2f} ms {f2 / max(s2, 1e-12) / 1e12:6.1f} TF {c2:5d} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
