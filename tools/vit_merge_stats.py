"""How often does the speculative conv-coded decode have to fall back to the exact decoder?

A numpy model of the K = 7, r = 1/2 hard-decision Viterbi decoder (generators 0x6d / 0x4f, as csrc/packet.cu) on a binary
symmetric channel, CPU only.  For each frame it runs the exact decoder, then
  * the thread-parallel traceback: slices of the trellis walked from state 0 starting OV steps above the slice
    (128 threads x 80 steps and 32 threads x 304 steps for a 1200-byte payload, OV = 128 as VIT_TB_OV), and
  * the segmented recursion: four segments, each started SEG_OV = 160 steps early from equal metrics and run 160 steps past
    its end (VIT_OV), then traced back the same way,
and counts the frames in which any kept bit differs from the exact decoder's -- the frames whose CRC check would send them
to the exact decoder (frames the exact decoder gets wrong as well fail the CRC either way).

    python tools/vit_merge_stats.py [frames per error rate]
"""
import sys
import numpy as np

NS = 64
PREV0 = np.arange(NS) >> 1                 # predecessor with the oldest bit clear
PREV1 = PREV0 | 32
BIT = np.arange(NS) & 1


def parity(v):
    v = v ^ (v >> 4); v = v ^ (v >> 2); v = v ^ (v >> 1)
    return v & 1


def branch_out(prev, bit):
    reg = (prev << 1) | bit
    return (parity(reg & 0x6d) << 1) | parity(reg & 0x4f)


OUT0 = branch_out(PREV0, BIT)              # expected pair on the branch PREV0[s] -> s
OUT1 = branch_out(PREV1, BIT)
POPC2 = np.array([0, 1, 1, 2])


def encode(bits):
    sr = 0
    out = np.empty(len(bits), np.int64)
    for t, b in enumerate(bits):
        sr = ((sr << 1) | int(b)) & 0x7f
        out[t] = (parity(np.int64(sr & 0x6d)) << 1) | parity(np.int64(sr & 0x4f))
    return out


def acs(rx, t0, t1, metrics):
    """recursion over steps [t0, t1): returns the decisions (t1 - t0, 64) and the final metrics"""
    dec = np.empty((t1 - t0, NS), np.uint8)
    m = metrics.copy()
    for t in range(t0, t1):
        c0 = m[PREV0] + POPC2[OUT0 ^ rx[t]]
        c1 = m[PREV1] + POPC2[OUT1 ^ rx[t]]
        d = c1 < c0
        dec[t - t0] = d
        m = np.where(d, c1, c0)
    return dec, m


def traceback(dec, base, hi, lo, state):
    """walk steps hi-1 .. lo from `state` (the state after step hi-1); returns the input bits of steps [lo, hi)"""
    bits = np.empty(hi - lo, np.uint8)
    s = state
    for t in range(hi - 1, lo - 1, -1):
        bits[t - lo] = s & 1
        s = (s >> 1) | (32 if dec[t - base, s] else 0)
    return bits


def sliced(dec, base, lo, keep_hi, hi, nth, ov):
    """the thread-parallel traceback of csrc/packet.cu viterbi27_traceback_spec; returns the bits of [lo, keep_hi)"""
    sl = (((keep_hi - lo) + nth - 1) // nth + 7) & ~7
    out = np.empty(keep_hi - lo, np.uint8)
    for th in range(nth):
        a = lo + th * sl
        if a >= keep_hi:
            break
        b = min(keep_hi, a + sl)
        S = min(hi, b + ov)
        bits = traceback(dec, base, S, a, 0)
        out[a - lo:b - lo] = bits[:b - a]
    return out


def main():
    nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n = 1204                                   # 1200-byte payload + CRC-32
    nbits = 8 * n + 6
    TB_OV, SEG_OV = 128, 160
    rng = np.random.default_rng(7)
    print("bit error rate | exact decoder wrong | fallbacks: 128-thread traceback | 32-thread traceback | segmented recursion   (of %d frames)" % nframes)
    for ber in (0.0, 0.01, 0.03, 0.05, 0.07, 0.10):
        bad_exact = fb128 = fb32 = fbseg = 0
        for _ in range(nframes):
            msg = np.concatenate([rng.integers(0, 2, 8 * n), np.zeros(6, np.int64)])
            rx = encode(msg)
            flips = (rng.random(nbits) < ber).astype(np.int64) | ((rng.random(nbits) < ber).astype(np.int64) << 1)
            rx = rx ^ flips
            start = np.full(NS, 63, np.int64); start[0] = 0
            dec, _ = acs(rx, 0, nbits, start)
            exact = traceback(dec, 0, nbits, 0, 0)
            bad_exact += int(np.any(exact[:8 * n] != msg[:8 * n]))
            fb128 += int(np.any(sliced(dec, 0, 0, nbits, nbits, 128, TB_OV)[:8 * n] != exact[:8 * n]))
            fb32 += int(np.any(sliced(dec, 0, 0, nbits, nbits, 32, TB_OV)[:8 * n] != exact[:8 * n]))
            # four segments (viterbi27_decode_par)
            seg = ((nbits + 3) // 4 + 79) // 80 * 80
            out = np.empty(nbits, np.uint8)
            for wp in range(4):
                a = wp * seg
                if a >= nbits:
                    break
                b = min(nbits, a + seg)
                s0 = 0 if wp == 0 else a - SEG_OV
                e0 = nbits if (b == nbits or b + SEG_OV >= nbits) else b + SEG_OV
                m0 = start if wp == 0 else np.zeros(NS, np.int64)
                d, _ = acs(rx, s0, e0, m0)
                out[a:b] = sliced(d, s0, a, b, e0, 32, TB_OV)
            fbseg += int(np.any(out[:8 * n] != exact[:8 * n]))
        print("   %5.2f       |       %4d          |            %4d                 |        %4d         |       %4d" % (ber, bad_exact, fb128, fb32, fbseg))


if __name__ == "__main__":
    main()
