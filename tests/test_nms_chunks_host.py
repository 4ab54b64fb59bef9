"""k_nms (watsor_b200/csrc/kernels_post.cu) visits a class's candidates in chunks picked from a score histogram instead
of sorting them all.  The claim that makes this exact -- "the chunks, each sorted, concatenate to the full descending
key order" -- is a property of the bin function and of the cut search, checked here with a line-by-line numpy model
of both (warp scan included) on adversarial score distributions: uniform, sigmoid-shaped, all equal, two clusters at
the ends of the range, sub-threshold tails.  The kernel itself is compared with the oracle in tests/test_gpu_stages.py."""
import numpy as np
import pytest

NMS_BINS, CHUNK0, CHUNK = 1024, 96, 384


def nms_bin(keys):
    """sign + exponent + 5 mantissa bits of the score, counted from 2^-27 (kernels_post.cu: nms_bin)"""
    return np.clip((keys >> np.uint64(50)).astype(np.int64) - (100 << 5), 0, NMS_BINS - 1)


def cut_search(hist, hi_bin, want):
    """warp 0 of k_nms: 32 bins per iteration from hi_bin - 1 downwards, inclusive prefix over lanes, first lane that
    reaches `want` wins; nothing left -> cut 0"""
    acc, cut, top = 0, 0, hi_bin - 1
    while top >= 0 and acc < want:
        b = top - np.arange(32)
        incl = np.cumsum(np.where(b >= 0, hist[np.maximum(b, 0)], 0))
        reach = np.nonzero(acc + incl >= want)[0]
        if len(reach):
            return top - int(reach[0])
        acc += int(incl[31])
        cut = max(top - 31, 0)
        top -= 32
    return cut


def scores(kind, n, rng):
    if kind == 'uniform':
        return rng.random(n).astype(np.float32)
    if kind == 'sigmoid':
        return (1 / (1 + np.exp(-rng.normal(-4.5, 1.0, n)))).astype(np.float32)
    if kind == 'equal':
        return np.full(n, 0.0123, np.float32)
    if kind == 'two_clusters':
        return np.concatenate([np.full(n // 2, 1.0, np.float32), (rng.random(n - n // 2) * 1e-7 + 1.1e-8).astype(np.float32)])
    if kind == 'tiny':
        return (rng.random(n) * 1e-9).astype(np.float32) + np.float32(1e-12)
    raise ValueError(kind)


@pytest.mark.parametrize('kind', ['uniform', 'sigmoid', 'equal', 'two_clusters', 'tiny'])
def test_chunks_concatenate_to_the_descending_key_order(kind):
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(641, 1918))               # the chunked path is taken above 640 candidates
        sc = scores(kind, n, rng)
        keys = (sc.view(np.uint32).astype(np.uint64) << np.uint64(32)) | \
               (np.uint64(0xFFFFFFFF) - np.arange(n, dtype=np.uint64))      # score_bits << 32 | ~anchor
        bins = nms_bin(keys)
        assert np.all(np.diff(bins[np.argsort(keys)]) >= 0)                  # the bin function is monotone in the key
        hist = np.bincount(bins, minlength=NMS_BINS)
        hi, chunk, visited = NMS_BINS, 0, []
        while hi > 0:
            want = CHUNK0 if chunk == 0 else CHUNK
            cut = cut_search(hist, hi, want)
            sel = keys[(bins >= cut) & (bins < hi)]
            assert len(sel) <= 2048                                          # fits the kernel's sort buffer
            assert len(sel) >= want or cut == 0                              # short only when nothing is left
            visited.append(np.sort(sel)[::-1])
            hi, chunk = cut, chunk + 1
        visited = np.concatenate(visited)
        assert len(visited) == n and np.array_equal(visited, np.sort(keys)[::-1])
