"""CPU analysis behind K8 v5's list walk: set-bit list lengths of ORB descriptors, the majority flip, the load
balance of warps / CTAs when a query costs what its list is long, and one flip vector per pool cluster.  cv2 only, no GPU.
usage: python tools/k8_list_stats.py [pages=50] [frames=64]      (also reads the reference's fixture PNGs under tests/golden/ref_fixtures)
"""
import glob
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2

import synth

_ORB = None


def _orb():
    global _ORB
    if _ORB is None:
        _ORB = cv2.ORB_create(2000, 1.2, 8, 62, 0, 2, cv2.ORB_FAST_SCORE, 62, 20)     # feature_extractor.rs:13-23
    return _ORB


def _canonical(img):
    """descriptors in the library's canonical order (octave, y, x) -- what K8 sees as consecutive queries"""
    k, d = _orb().detectAndCompute(img, None)
    if d is None:
        return np.zeros((0, 32), np.uint8)
    key = np.array([(kp.octave, round(kp.pt[1] / 1.2 ** kp.octave), round(kp.pt[0] / 1.2 ** kp.octave)) for kp in k])
    return d[np.lexsort((key[:, 2], key[:, 1], key[:, 0]))]


def _job(a):
    kind, i, npg = a
    return _canonical(synth.make_page(i) if kind == "page" else synth.make_frame(i, npg))


def majority(pool_bits):
    n_s = min(len(pool_bits), 8192)                                   # knn5_flip_kernel's sample
    return (2 * pool_bits[(np.arange(n_s, dtype=np.int64) * len(pool_bits)) // n_s].sum(0) > n_s).astype(np.uint8)


def list_len(q_bits, m):
    a = (q_bits ^ m).sum(1)
    return np.minimum(a, 256 - a)


def report(name, pool, q):
    pb, qb = np.unpackbits(pool, axis=1), np.unpackbits(q, axis=1)
    p = pb.mean(0)
    print(f"== {name}: pool {len(pool)}, queries {len(q)}; sum over bits of min(p, 1 - p) = {np.minimum(p, 1 - p).sum():.1f} (128 for unbiased bits)")
    zero, maj = np.zeros(256, np.uint8), majority(pb)
    for label, m in (("plain", zero), ("majority flip", maj)):
        n = list_len(qb, m)
        for gran, nh in (("blocks of 16", 2 * ((n + 15) // 16)), ("half blocks of 8", (n + 7) // 8)):
            nt_ = len(nh) // 128
            if nt_ < 148:
                print(f"  {label:14s} {gran:16s}: mean list {n.mean():6.1f} entries, {nh.mean() / 2:5.2f} blocks of 16 per query")
                continue
            tiles = nh[:nt_ * 128].reshape(nt_, 128)
            t = tiles.sum(1).astype(float)
            waves = nt_ // 148
            fixed = 0.3 * 16 * 128                                     # survivor handling etc.: ~30 % of a dense tile
            contiguous = (t + fixed)[:waves * 148].reshape(waves, 148).sum(0)
            strided = nh[:nt_ * 128].reshape(128, nt_).T.sum(1).astype(float)      # tile t = rows ql * n_tiles + t
            even = (strided + fixed)[:waves * 148].reshape(waves, 148).sum(0)
            natural = tiles.reshape(nt_, 16, 8).sum(2)
            srt = -np.sort(-tiles, axis=1)
            dealt = np.zeros((nt_, 16))
            for r in range(8):
                seg = srt[:, 16 * r:16 * r + 16]
                dealt += seg if r % 2 == 0 else seg[:, ::-1]
            print(f"  {label:14s} {gran:16s}: mean list {n.mean():6.1f} entries, {nh.mean() / 2:5.2f} blocks of 16 per query | "
                  f"tile-to-tile cv {t.std() / t.mean():.3f}, CTA max/mean over {waves} waves: contiguous tiles {contiguous.max() / contiguous.mean():.3f}, "
                  f"even tiles {even.max() / even.mean():.3f} | warp max/mean inside a tile: natural {natural.max(1).mean() / natural.mean():.3f}, "
                  f"dealt {dealt.max(1).mean() / dealt.mean():.3f}")
    # one flip vector per cluster of the pool (k-majority clustering): cost = sum over clusters of weight x mean blocks against its vector
    rng = np.random.default_rng(0)
    blocks = lambda m: ((list_len(qb, m) + 15) // 16).mean()
    line = [f"1 vector {blocks(maj):.2f}"]
    for K in (2, 4, 8):
        if len(pb) < 64 * K:
            break
        cent = pb[rng.choice(len(pb), K, replace=False)].copy()
        for _ in range(12):
            d = np.stack([(pb ^ c).sum(1) for c in cent], 1)
            lab = np.minimum(d, 256 - d).argmin(1)
            for k in range(K):
                sel = pb[lab == k]
                if len(sel):
                    sel = sel ^ (d[lab == k, k] > 128)[:, None].astype(np.uint8)
                    cent[k] = (2 * sel.sum(0) > len(sel)).astype(np.uint8)
        w = np.bincount(lab, minlength=K) / len(pb)
        line.append(f"{K} clusters {sum(w[k] * blocks(cent[k]) for k in range(K)):.2f}")
    print("  blocks of 16 per (query, slab) with one flip vector per pool cluster: " + ", ".join(line))


def main():
    npg = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    with Pool(min(16, os.cpu_count() or 1)) as pool:
        pages = pool.map(_job, [("page", p, npg) for p in range(npg)])
        frames = pool.map(_job, [("frame", f, npg) for f in range(nf)])
    report(f"synthetic deck, {npg} pages x {nf} frames (bench workload)", np.concatenate(pages), np.concatenate(frames))
    fx = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_fixtures", "*.png")))
    slides = [f for f in fx if "slide" in os.path.basename(f)]
    shots = [f for f in fx if "slide" not in os.path.basename(f)]
    if slides and shots:
        report("the reference's fixture PNGs (data/matchings/test1)", np.concatenate([_canonical(cv2.imread(f)) for f in slides]),
               np.concatenate([_canonical(cv2.imread(f)) for f in shots]))
    rng = np.random.default_rng(7)
    report("uniform random descriptors (the matcher sweep)", rng.integers(0, 256, (20000, 32), dtype=np.uint8),
           rng.integers(0, 256, (148 * 128 * 2, 32), dtype=np.uint8))


if __name__ == "__main__":
    main()
