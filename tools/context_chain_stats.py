"""How serial is a line of the restart-interval-1 lossless coder?  For rows of the bench's S_smooth frame (cfg2, 8 bits) this counts,
per sample: which of the five contexts |Q(-Ra)| it uses, and how many samples of the SAME context lie between two changes of that
context's bias C (T.87 A.13) - the events at which a group of lanes that speculates on (B, C) has to stop and start over
(the warp-per-line encoder that was not built).  Plain Python restatement of the context update
(reference src/regular_mode_context.hpp:45-94); CPU only."""
import sys

import numpy as np


def s_smooth_rows(rows, w=4096, h=4096, seed=1234):
    rng = np.random.default_rng(seed)
    mx = 255
    x = np.arange(w)[None, :]
    y = np.arange(h)[:, None]
    base = 0.8 * mx * (0.5 + 0.25 * np.sin(x / 97.0) + 0.25 * np.cos(y / 131.0))
    img = np.clip(base + rng.normal(0, 0.01 * mx, (h, w)), 0, mx).astype(np.uint8)
    return img[rows]


def line_stats(row, t1=3, t2=7, t3=21, reset=64):
    ctx = [dict(a=4, b=0, c=0, n=1) for _ in range(5)]  # A = max(2, (RANGE + 32) / 64) = 4 for 8 bits
    use = [0] * 5
    gaps = []          # samples of one context between two changes of its C
    since = [0] * 5
    ra = 0
    for x in row.tolist():
        q = (ra >= t3) + (ra >= t2) + (ra >= t1) + (ra > 0)
        if q == 0:
            ra = x  # run mode (Ra == 0): rare in this image, not modelled
            continue
        c = ctx[q]
        use[q] += 1
        px = min(max(ra - c["c"], 0), 255)          # sign is negative for every q != 0 (Rb = Rc = Rd = 0)
        e = ((px - x + 128) & 255) - 128            # modulo 256, sign applied
        c["a"] += abs(e)
        c["b"] += e
        if c["n"] == reset:
            c["a"] >>= 1
            c["b"] >>= 1
            c["n"] >>= 1
        c["n"] += 1
        before = c["c"]
        if c["b"] + c["n"] <= 0:
            c["b"] = max(c["b"] + c["n"], 1 - c["n"])
            c["c"] = max(c["c"] - 1, -128)
        elif c["b"] > 0:
            c["b"] = min(c["b"] - c["n"], 0)
            c["c"] = min(c["c"] + 1, 127)
        since[q] += 1
        if c["c"] != before:
            gaps.append(since[q])
            since[q] = 0
        ra = x
    return use, gaps


def main():
    rows = [int(a) for a in sys.argv[1:]] or [0, 500, 1000, 2000, 3000, 4000]
    img = s_smooth_rows(rows)
    total_use, all_gaps = np.zeros(5, np.int64), []
    for r in img:
        use, gaps = line_stats(r)
        total_use += np.array(use)
        all_gaps += gaps
    share = total_use / total_use.sum()
    g = np.array(all_gaps)
    print("share of samples per context |Q(-Ra)| = 1..4:", np.round(share[1:], 4).tolist())
    print(f"changes of C: {len(g)} in {int(total_use.sum())} samples; samples of the context between two changes: "
          f"mean {g.mean():.2f}, median {np.median(g):.0f}, 90th percentile {np.percentile(g, 90):.0f}")
    for lanes in (8, 16, 32):
        # a group of `lanes` consecutive samples of one context commits up to and including its first C change
        p_clean = np.mean(g > lanes)
        rounds = lanes / np.minimum(g, lanes).mean()
        print(f"  groups of {lanes}: {100 * p_clean:.1f} % of the stretches between two changes are longer than the group; "
              f"about {rounds:.1f} commit rounds per group")


if __name__ == "__main__":
    main()
